#!/bin/bash
# First GPU evidence for the code written after round 1's GPU budget was spent: the GPU tests that
# have never run, one by one (no -x: one failure must not hide the others), f2 timings (level vs FIFO region grow, two
# batch sizes) and an ncu capture of k_peac_frame.
#   gpurun --timeout 900 -- 'bash tools/gpu_peac.sh r02a'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for t in tests/test_v_reference_golden_gpu.py tests/test_widened_mappoint_gpu.py tests/test_x_peac_gpu.py tests/test_y_reference_mapping_gpu.py; do
  timeout 600 python -m pytest $t -m gpu -q > $OUT/${TAG}_$(basename $t .py).log 2>&1
  echo "== $t: exit $?"; tail -3 $OUT/${TAG}_$(basename $t .py).log
done
for b in 16 148 256; do
  timeout 300 python tools/peac_time.py $b 5 | tee $OUT/${TAG}_peac_levels_b$b.json
done
MSL_PEAC_FLOOD_SERIAL=1 timeout 600 python tools/peac_time.py 16 3 | tee $OUT/${TAG}_peac_fifo_b16.json
for t in 128 512; do MSL_PEAC_THREADS=$t timeout 300 python tools/peac_time.py 148 5 | tee $OUT/${TAG}_peac_levels_b148_t$t.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_peac_frame -c 1 -f -o $OUT/${TAG}_k_peac_frame \
  python tools/peac_time.py 16 1 > $OUT/${TAG}_ncu_peac.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_peac_frame.ncu-rep > $OUT/${TAG}_k_peac_frame_brief.txt 2>&1
cat $OUT/${TAG}_k_peac_frame_brief.txt
timeout 300 python bench.py --widened-only peac | tee $OUT/${TAG}_widened_peac.json
