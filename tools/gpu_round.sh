#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, the bench line, scan-kernel A/B over its tuning knobs,
# the ncu launch list of a short bench run and ncu --set full captures of the two fuse kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a'
# Everything lands in gpurun_out/<tag>_*; nothing printed under ncu is a bench value.
TAG=${1:-run}
shift
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

if [[ " $* " != *" notest "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi

timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 2500 $OUT/${TAG}_bench.json

if [[ " $* " == *" ab "* ]]; then
  for cfg in "1 3" "2 3" "2 5" "3 2" "4 2"; do
    set -- $cfg
    MSL_SCAN_STAGES=$1 MSL_SCAN_CTAS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
      > $OUT/${TAG}_ab_s$1_c$2.json 2>> $OUT/${TAG}_bench.err
    python - "$OUT/${TAG}_ab_s$1_c$2.json" "$cfg" <<'EOF'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    r = j["roofline"]
    print("AB stages,ctas=%s  value %.0f  scan %.1f us (isolated %.1f)  chain iso %s" % (
        sys.argv[2], j["value"], 1e3 * r["avg_launch_ms"], 1e3 * r["isolated"]["avg_launch_ms"],
        {k: round(v, 1) for k, v in r["isolated"]["chain_us_per_frame"].items()}))
except Exception as e:
    print("AB", sys.argv[2], "failed", e)
EOF
  done
fi

if [[ " $* " != *" noncu "* ]]; then
  # launch list (per-launch times are cold-cache and serialised: shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
  head -30 $OUT/${TAG}_launches_summary.txt
  # full captures of the two fuse kernels (one launch each, a frame in the steady state of the stream)
  for kn in k_fuse_scan k_fuse_apply; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 40 -c 1 -f -o $OUT/${TAG}_$kn \
      python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline >> $OUT/${TAG}_ncu_bench.log 2>&1
    python tools/ncu_brief.py $OUT/${TAG}_$kn.ncu-rep > $OUT/${TAG}_${kn}_brief.txt 2>&1
    cat $OUT/${TAG}_${kn}_brief.txt
  done
fi
