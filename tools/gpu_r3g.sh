#!/bin/bash
# Round 2, call 3g: co-scheduling knobs of the step (stream priorities, fuse CTAs per SM, split superpixel stage), one run each.
TAG=${1:-r3g}
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
print('%-28s value %.0f ms/step %.3f e2e %.0f fuse in-step %.1f us iso %.1f us' % ('$name', j['value'], j['ms_per_step'], j['e2e']['value'], r['avg_launch_ms']*1000, r['isolated']['avg_launch_ms']*1000))
"
}
run default A=1
run sp_prio2 MSL_SP_PRIO=2
run sp_prio1 MSL_SP_PRIO=1
run wave2 MSL_STREAM_WAVE=2
run wave2_prio2 MSL_STREAM_WAVE=2 MSL_SP_PRIO=2
run split MSL_SP_SPLIT=1
run wave2_split MSL_STREAM_WAVE=2 MSL_SP_SPLIT=1
run wave1 MSL_STREAM_WAVE=1
run default_again A=1
tail -c 300 $OUT/${TAG}_ab.err
