#!/bin/bash
# Round 2, call 3b: ncu --set full of k_fuse_pipe with and without carried partial rounds (same launch index of the same run).
TAG=${1:-r3b}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for c in 0 1; do
  MSL_FUSE_CARRY=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o $OUT/${TAG}_k_fuse_pipe_c$c \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu_c$c.log 2>&1
  python tools/ncu_brief.py $OUT/${TAG}_k_fuse_pipe_c$c.ncu-rep > $OUT/${TAG}_k_fuse_pipe_c${c}_brief.txt 2>&1
  cat $OUT/${TAG}_k_fuse_pipe_c${c}_brief.txt
done
