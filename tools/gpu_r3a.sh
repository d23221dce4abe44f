#!/bin/bash
# Round 2, call 3a: k_fuse_pipe with carried partial fuse rounds (MSL_FUSE_CARRY), parity + A/B.
TAG=${1:-r3a}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "pipe_carry or default" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest s8 exit $?" >> $OUT/${TAG}_pytest.log
MSL_FUSE_CARRY=1 timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q >> $OUT/${TAG}_pytest.log 2>&1
echo "pytest carry exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -12
for c in 0 1 0 1; do
  MSL_FUSE_CARRY=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_c$c.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_bench_c$c.json "[carry=$c]"
  python -c "
import json
j=json.load(open('$OUT/${TAG}_bench_c$c.json'))
print('   ms/step %.3f e2e %.0f frac %.3f iso %.3f' % (j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['isolated']['frac']))
"
done
tail -c 300 $OUT/${TAG}_ab.err
