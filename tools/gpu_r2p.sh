#!/bin/bash
# Round 2, call p: ncu --set full captures with source of the ORB / superpixel kernels (one launch each, batch 16).
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for k in k_fast_cells k_blur k_sp_pixels k_sp_seeds2 k_sp_fit2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 1 -c 1 -f -o $OUT/${TAG}_$k \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu_$k.log 2>&1
  python tools/ncu_brief.py $OUT/${TAG}_$k.ncu-rep 2>&1 | head -24
done
ls -la $OUT/*.ncu-rep
