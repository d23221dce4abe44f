#!/bin/bash
# Round 2, call s: k_sp_seeds3 (the group's region staged in shared memory): parity, stage time, bench A/B.
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_v_reference_golden_gpu.py tests/test_s8_bench_scale_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
for v in 1 2; do
MSL_SP_V2=$v MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1_$v.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1_$v.json'));print('DIAG1 MSL_SP_V2=$v superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
MSL_SP_V2=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$v.json 2>> $OUT/${TAG}_ab.err
python tools/ab_line.py $OUT/${TAG}_ab$v.json "[MSL_SP_V2=$v]"
python -c "
import json
j=json.load(open('$OUT/${TAG}_ab$v.json'))
print('   ms/step %.3f e2e %.0f' % (j['ms_per_step'], j['e2e']['value']))
"
done
tail -c 300 $OUT/${TAG}_ab.err
