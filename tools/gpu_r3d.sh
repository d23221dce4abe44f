#!/bin/bash
# Round 2, call 3d: launch times of the superpixel kernels with k_sp_pixels / k_sp_pixels4 (ncu, durations only; not bench values).
TAG=${1:-r3d}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for c in 0 1; do
  MSL_SP_PIX4=$c MSL_DIAG=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sp_ --csv --log-file $OUT/${TAG}_sp_c$c.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_ncu_c$c.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/${TAG}_sp_c$c.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
d = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    d[r[ki]][r[mi]].append(float(r[vi].replace(",", "")))
print("pix4=$c")
for k, m in d.items():
    t = m["gpu__time_duration.sum"]
    print("  %-14s n=%3d avg %.1f us  inst %.2fM  issue %.0f%%  warps %.0f%%" % (k[:14], len(t), sum(t) / len(t) / 1000, sum(m["smsp__inst_executed.sum"]) / len(t) / 1e6, sum(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / len(t), sum(m["sm__warps_active.avg.pct_of_peak_sustained_active"]) / len(t)))
PY
done
