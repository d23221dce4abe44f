#!/bin/bash
# Round 2, call 3y: k_resize4 (four output pixels per thread, packed column table), ORB parity + A/B (MSL_ORB_RESIZE4).
TAG=${1:-r3y}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_orb_gpu.py tests/test_v_reference_golden_gpu.py tests/test_glue_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -6
run() {
  name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
j=json.load(open('$OUT/${TAG}_$name.json'))
print('%-14s %s value %.0f ms/step %.3f e2e %.0f' % ('$name', '$wl'[:12], j['value'], j['ms_per_step'], j['e2e']['value']))
"
}
run orb_r1 orb_match_640x480_b64 MSL_ORB_RESIZE4=0
run orb_r4 orb_match_640x480_b64 MSL_ORB_RESIZE4=1
run orb_r1_b orb_match_640x480_b64 MSL_ORB_RESIZE4=0
run orb_r4_b orb_match_640x480_b64 MSL_ORB_RESIZE4=1
run def_r1 frontend_640x480_b64_map5M MSL_ORB_RESIZE4=0
run def_r4 frontend_640x480_b64_map5M MSL_ORB_RESIZE4=1
tail -c 300 $OUT/${TAG}_ab.err
