#!/bin/bash
# Round 2, call 3s: float-threshold comparisons in the superpixel kernels (no float->double conversion for a compare), parity + timing.
TAG=${1:-r3s}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q -k "not two_kernel and not stream and not pipe_" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -8
for i in 1 2; do
  MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_diag_$i.json
  python -c "import json;j=json.load(open('$OUT/${TAG}_diag_$i.json'));print('superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_bench_$i.json
  python -c "
import json
j=json.load(open('$OUT/${TAG}_bench_$i.json'))
print('value %.0f ms/step %.3f e2e %.0f frac %.3f iso %.3f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['isolated']['frac']))
"
done
MSL_DIAG=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_sp_ --csv --log-file $OUT/${TAG}_sp.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/${TAG}_sp.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value")
d = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    d[r[ki]][r[mi]].append(float(r[vi].replace(",", "")))
for k, m in d.items():
    t = m["gpu__time_duration.sum"]
    print("  %-40s n=%3d avg %.1f us  inst %.2fM  issue %.0f%%" % (k[-40:], len(t), sum(t) / len(t) / 1000, sum(m["smsp__inst_executed.sum"]) / len(t) / 1e6, sum(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / len(t)))
PY
tail -c 300 $OUT/${TAG}_ab.err
