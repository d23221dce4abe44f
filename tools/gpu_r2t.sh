#!/bin/bash
# Round 2, call t: the superpixel stage of a batch as two half batches on two streams (MSL_SP_SPLIT=1).
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
MSL_SP_SPLIT=1 timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
for v in 0 1; do
MSL_SP_SPLIT=$v MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1_$v.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1_$v.json'));print('DIAG1 MSL_SP_SPLIT=$v superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
MSL_SP_SPLIT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$v.json 2>> $OUT/${TAG}_ab.err
python tools/ab_line.py $OUT/${TAG}_ab$v.json "[MSL_SP_SPLIT=$v]"
python -c "
import json
j=json.load(open('$OUT/${TAG}_ab$v.json'))
print('   ms/step %.3f e2e %.0f' % (j['ms_per_step'], j['e2e']['value']))
"
done
tail -c 300 $OUT/${TAG}_ab.err
