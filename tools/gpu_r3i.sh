#!/bin/bash
# Round 2, call 3i: in the two-CTAs-per-SM geometry a warp has more registers and shared memory to spend on hiding its own
# latency: PRE (first fuse round's gathers an iteration ahead, 128-register instantiation) and carried rounds, one run each.
TAG=${1:-r3i}
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
print('%-22s value %.0f ms/step %.3f e2e %.0f fuse in-step %.1f us (frac %.3f) iso %.1f us (%.3f) grid %d' % ('$name', j['value'], j['ms_per_step'], j['e2e']['value'], r['avg_launch_ms']*1000, r['frac'], r['isolated']['avg_launch_ms']*1000, r['isolated']['frac'], r['launch']['grid']))
"
}
run default A=1
run pre_regs2 MSL_STREAM_PRE=1 MSL_STREAM_REGS=2
run carry MSL_FUSE_CARRY=1
run pf3 MSL_STREAM_PF=3
run pf5 MSL_STREAM_PF=5
run default_again A=1
tail -c 300 $OUT/${TAG}_ab.err
