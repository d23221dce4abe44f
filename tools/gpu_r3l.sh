#!/bin/bash
# Round 2, call 3l (2 GPUs): where the N = 2 step's extra 0.35 ms comes from -- with / without the count-table collective, and in
# the three-CTA geometry.
TAG=${1:-r3l}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run2() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-16s N=2 value %.0f ms %.3f e2e %.0f per_rank %s' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value'], b.get('per_rank')))
"
}
run1() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-16s N=1 value %.0f ms %.3f e2e %.0f' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value']))
"
}
run1 n1 A=1
run1 n1_gpu1 CUDA_VISIBLE_DEVICES=1
run2 n2 A=1
run2 n2_nogather MSL_BENCH_NO_GATHER=1
run2 n2_wave3 MSL_STREAM_WAVE=3
run1 n1_wave3 MSL_STREAM_WAVE=3
tail -c 300 $OUT/${TAG}.err
