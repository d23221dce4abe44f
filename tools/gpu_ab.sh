#!/bin/bash
# Quick A/B of environment knobs on the bench (no tests, no ncu):  gpurun -- 'bash tools/gpu_ab.sh TAG "K1=V1 K2=V2" "K1=V3" ...'
TAG=$1
shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ab$i.json 2>> gpurun_out/${TAG}_ab.err
  python tools/ab_line.py gpurun_out/${TAG}_ab$i.json "[$cfg]"
done
