#!/bin/bash
# Round 2, call e: superpixel kernels with row-vectorised loads (parity, ncu timing, A/B), k_peac_frame after the merge-step
# rework (parity, phase profile at 256 / 512 / 1024 threads, batch 8 and 64).
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_x_peac_gpu.py tests/test_v_reference_golden_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "default or superpixels_v1" > $OUT/${TAG}_s8.log 2>&1
tail -4 $OUT/${TAG}_s8.log
i=0
for cfg in "MSL_SP_V2=0" "MSL_SP_V2=1" "MSL_SP_V2=1 MSL_STREAM_WAVE=2"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
for v in 0 1; do
MSL_SP_V2=$v MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1_$v.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1_$v.json'));print('DIAG1 SP_V2=$v superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
done
NCU_FUSE_FRAMES=2 timeout 600 ncu --set full --clock-control none -k regex:'k_sp_' -f -o /tmp/${TAG}_sp python tools/ncu_kernels.py > $OUT/${TAG}_ncu_sp.log 2>&1
python tools/ncu_brief.py /tmp/${TAG}_sp.ncu-rep > $OUT/${TAG}_sp_brief.txt 2>&1
grep -E "captured|gpu__time_duration|dram__bytes|issue_active|warps_active|stalls" $OUT/${TAG}_sp_brief.txt | sed 's/  */ /g' | awk '/captured/{print ""; printf "%s | ", $0; next} {printf "%s | ", $0} END{print ""}' | cut -c1-420
for t in 256 512 1024; do
  for b in 8 64; do
    MSL_PEAC_THREADS=$t timeout 300 python tools/peac_time.py $b 3 > $OUT/${TAG}_peac_t${t}_b$b.json 2>> $OUT/${TAG}_peac.err
    python -c "
import json
j=json.load(open('$OUT/${TAG}_peac_t${t}_b$b.json'))
print('PEAC threads $t batch $b ms/batch %.2f equal %s prof %s' % (j['ms_per_batch_min'], j['equals_oracle_first_frames'], j['profile']))
"
  done
done
MSL_PEAC_FLOOD_SERIAL=1 timeout 300 python tools/peac_time.py 64 3 > $OUT/${TAG}_peac_serial_b64.json 2>> $OUT/${TAG}_peac.err
python -c "
import json
j=json.load(open('$OUT/${TAG}_peac_serial_b64.json'))
print('PEAC serial flood batch 64 ms/batch %.2f prof %s' % (j['ms_per_batch_min'], j['profile']))
"
du -sh $OUT
