#!/bin/bash
# Round 2, call 3t: CTA count of the chain's launches re-swept after the superpixel stage got lighter.
TAG=${1:-r3t}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
j=json.load(open('$OUT/${TAG}_$name.json'))
r=j['roofline']
print('%-12s value %.0f ms/step %.3f e2e %.0f fuse in-step %.1f us (frac %.3f)' % ('$name', j['value'], j['ms_per_step'], j['e2e']['value'], r['avg_launch_ms']*1000, r['frac']))
"
}
for g in 296 314 333 352 296 314 333; do run grid${g}_$RANDOM MSL_STREAM_GRID=$g; done
tail -c 200 $OUT/${TAG}_ab.err
