#!/bin/bash
# Round 2, call 3h: CTAs per k_fuse_pipe launch (MSL_STREAM_GRID) swept between two and three per SM.
TAG=${1:-r3h}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_$name.json 2>> $OUT/${TAG}_ab.err
  python -c "
import json
j=json.load(open('$OUT/${TAG}_$name.json'))
r=j['roofline']
print('%-28s value %.0f ms/step %.3f e2e %.0f fuse in-step %.1f us (frac %.3f) iso %.1f us (%.3f)' % ('$name', j['value'], j['ms_per_step'], j['e2e']['value'], r['avg_launch_ms']*1000, r['frac'], r['isolated']['avg_launch_ms']*1000, r['isolated']['frac']))
"
}
for g in 222 259 296 333 370 407 444 296; do run grid$g MSL_STREAM_GRID=$g; done
tail -c 300 $OUT/${TAG}_ab.err
