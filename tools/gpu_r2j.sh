#!/bin/bash
# Round 2, call j: ncu --set full captures with source of k_fuse_pipe (bench, 40th launch) and k_peac_frame (batch 8); the
# reports come back in gpurun_out/ for line-level reading here.
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o $OUT/${TAG}_k_fuse_pipe \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_fuse_pipe.ncu-rep > $OUT/${TAG}_k_fuse_pipe_brief.txt 2>&1
head -22 $OUT/${TAG}_k_fuse_pipe_brief.txt
MSL_PEAC_THREADS=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_peac_frame -s 1 -c 1 -f -o $OUT/${TAG}_k_peac_frame \
  python tools/peac_time.py 8 1 > $OUT/${TAG}_ncu_peac.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_peac_frame.ncu-rep > $OUT/${TAG}_k_peac_frame_brief.txt 2>&1
head -22 $OUT/${TAG}_k_peac_frame_brief.txt
ls -la $OUT/*.ncu-rep
