#!/bin/bash
# Round 2, call x: ncu of k_sp_bin / k_sp_seeds4 (batch 16)
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for k in k_sp_bin k_sp_seeds4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$k \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu_$k.log 2>&1
  python tools/ncu_brief.py $OUT/${TAG}_$k.ncu-rep 2>&1 | head -24
done
