#!/bin/bash
# Round 2, call q: FAST without per-pixel divisions + zero-bordered NMS, blur with four outputs per thread, k_sp_pixels with
# its loads issued together: parity (ORB / surfel / reference goldens), isolated stage times, bench.
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_surfel_gpu.py tests/test_v_reference_golden_gpu.py tests/test_track_batch_gpu.py tests/test_s8_bench_scale_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/ab_line.py $OUT/${TAG}_bench.json "[default]"
python -c "
import json
j=json.load(open('$OUT/${TAG}_bench.json'))
print('   ms/step %.3f e2e %.0f host_calls %.0f' % (j['ms_per_step'], j['e2e']['value'], j['e2e_host_calls']['value']))
"
for o in orb,match,plane; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only $o > $OUT/${TAG}_only_${o//,/_}.json 2>> $OUT/${TAG}_ab.err
  python -c "import json;j=json.load(open('$OUT/${TAG}_only_${o//,/_}.json'));print('ONLY $o ms_per_step %.3f' % j['ms_per_step'])"
done
MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1.json'));print('DIAG1 superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
tail -c 300 $OUT/${TAG}_bench.err
