#!/bin/bash
# Round 2, call 3j: programmatic dependent launch along the fuse chain (MSL_FUSE_PDL), parity + A/B.
TAG=${1:-r3j}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -8
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
run pdl0 MSL_FUSE_PDL=0
run pdl1 MSL_FUSE_PDL=1
run pdl0_b MSL_FUSE_PDL=0
run pdl1_b MSL_FUSE_PDL=1
run pdl1_grid333 MSL_FUSE_PDL=1 MSL_STREAM_GRID=333
run pdl1_wave3 MSL_FUSE_PDL=1 MSL_STREAM_WAVE=3
tail -c 300 $OUT/${TAG}_ab.err
