#!/bin/bash
# Round 2, call y: k_sp_pixels with the CTA's 18 candidate seed records in shared memory.
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1.json'));print('DIAG1 superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_ab.err
python tools/ab_line.py $OUT/${TAG}_bench.json "[default]"
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench.json'))
print('   ms/step %.3f e2e %.0f' % (j['ms_per_step'], j['e2e']['value']))
"
tail -c 300 $OUT/${TAG}_ab.err
