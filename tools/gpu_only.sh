#!/bin/bash
# Diagnostic: which stage bounds the step?  bench.py with subsets of the stages (not bench values).
#   gpurun -- 'bash tools/gpu_only.sh TAG surfel orb,match,plane orb ...'
TAG=$1
shift
mkdir -p gpurun_out
for o in "$@"; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --only $o > gpurun_out/${TAG}_only_${o//,/_}.json 2>> gpurun_out/${TAG}_only.err
  python - "$o" gpurun_out/${TAG}_only_${o//,/_}.json <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[2]))
    print("ONLY %-24s ms_per_step %.3f  chain %s" % (sys.argv[1], j["ms_per_step"], {k: round(v, 1) for k, v in j["roofline"]["chain_us_per_frame"].items()}))
except Exception as e:
    print("ONLY", sys.argv[1], "failed:", e)
PY
done
