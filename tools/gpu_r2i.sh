#!/bin/bash
# Round 2, call i: frame-set e2e (one upload per step), FAST compass-test compaction, k_peac_frame warp selection + 4-way
# region-grow batching, k_fuse_pipe prefetch knobs (MSL_STREAM_PF bits 1 and 2).
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_glue_gpu.py tests/test_x_peac_gpu.py tests/test_v_reference_golden_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
MSL_STREAM_PF=7 timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "default or pipe" > $OUT/${TAG}_s8.log 2>&1
echo "s8 exit $?" >> $OUT/${TAG}_s8.log
tail -4 $OUT/${TAG}_s8.log
MSL_PEAC_THREADS=512 timeout 300 python tools/peac_time.py 64 3 > $OUT/${TAG}_peac_t512_b64.json 2>> $OUT/${TAG}_peac.err
python -c "
import json
j=json.load(open('$OUT/${TAG}_peac_t512_b64.json'))
print('PEAC threads 512 batch 64 ms/batch %.2f equal %s prof %s' % (j['ms_per_batch_min'], j['equals_oracle_first_frames'], j['profile']))
"
i=0
for cfg in "MSL_STREAM_PF=1" "MSL_STREAM_PF=3" "MSL_STREAM_PF=5" "MSL_STREAM_PF=7"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
  python -c "
import json
j=json.load(open('$OUT/${TAG}_ab$i.json'))
print('   ms/step %.3f e2e %.0f (h2d %d) host_calls %.0f' % (j['ms_per_step'], j['e2e']['value'], j['e2e']['h2d_bytes_per_step'], j['e2e_host_calls']['value']))
"
done
tail -c 600 $OUT/${TAG}_ab.err
