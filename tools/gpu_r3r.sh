#!/bin/bash
# Round 2, call 3r: three staged segments per warp in the two-CTA geometry (MSL_PIPE_NB=3), parity + A/B.
TAG=${1:-r3r}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "pipe_nb3" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
MSL_PIPE_NB=3 timeout 900 python -m pytest tests/test_surfel_gpu.py -m gpu -q >> $OUT/${TAG}_pytest.log 2>&1
echo "pytest nb3 exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -8
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
j=json.load(open('$OUT/${TAG}_$name.json'))
r=j['roofline']
print('%-22s value %.0f ms/step %.3f e2e %.0f fuse in-step %.1f us (frac %.3f) iso %.1f us (%.3f) grid %d' % ('$name', j['value'], j['ms_per_step'], j['e2e']['value'], r['avg_launch_ms']*1000, r['frac'], r['isolated']['avg_launch_ms']*1000, r['isolated']['frac'], r['launch']['grid']))
"
}
run nb2 A=1
run nb3 MSL_PIPE_NB=3
run nb2_b A=1
run nb3_b MSL_PIPE_NB=3
run nb3_grid333 MSL_PIPE_NB=3 MSL_STREAM_GRID=333
run nb3_grid259 MSL_PIPE_NB=3 MSL_STREAM_GRID=259
tail -c 300 $OUT/${TAG}_ab.err
