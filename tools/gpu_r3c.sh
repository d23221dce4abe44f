#!/bin/bash
# Round 2, call 3c: k_sp_pixels4 (four pixels per thread), parity + A/B against k_sp_pixels (MSL_SP_PIX4=0).
TAG=${1:-r3c}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q -k "not two_kernel and not stream and not one" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -8
for c in 0 1 0 1; do
  MSL_SP_PIX4=$c MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag_c$c.json 2>> $OUT/${TAG}_ab.err
  python -c "import json;j=json.load(open('$OUT/${TAG}_diag_c$c.json'));print('pix4=$c superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
  MSL_SP_PIX4=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench_c$c.json 2>> $OUT/${TAG}_ab.err
  python -c "
import json
j=json.load(open('$OUT/${TAG}_bench_c$c.json'))
print('pix4=$c value %.0f ms/step %.3f e2e %.0f frac %.3f iso %.3f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['isolated']['frac']))
"
done
tail -c 300 $OUT/${TAG}_ab.err
