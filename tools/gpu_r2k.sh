#!/bin/bash
# Round 2, call k: k_fuse_pipe with the packed {depth, index} gather, seed-record base in registers, and the PRE form (first
# fuse round's gathers issued one iteration ahead) at 3 and 2 CTAs per SM.
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_s8_bench_scale_gpu.py -m gpu -q -k "not two_kernel and not stream" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
MSL_STREAM_PRE=1 timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py tests/test_surfel_gpu.py -m gpu -q -k "default or pipe or stream_of or panning" > $OUT/${TAG}_s8pre.log 2>&1
echo "s8 pre exit $?" >> $OUT/${TAG}_s8pre.log
tail -4 $OUT/${TAG}_s8pre.log
i=0
for cfg in "MSL_STREAM_PRE=0" "MSL_STREAM_PRE=1" "MSL_STREAM_PRE=1 MSL_STREAM_REGS=2 MSL_STREAM_WAVE=2" "MSL_STREAM_PRE=0 MSL_STREAM_WAVE=2"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
  python -c "
import json
j=json.load(open('$OUT/${TAG}_ab$i.json'))
print('   ms/step %.3f e2e %.0f host_calls %.0f' % (j['ms_per_step'], j['e2e']['value'], j['e2e_host_calls']['value']))
"
done
tail -c 600 $OUT/${TAG}_ab.err
