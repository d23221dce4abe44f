#!/bin/bash
# Round 2, call 3m (2 GPUs): is the clock sampler (one nvidia-smi -lms 20 per rank) what slows the second rank?
TAG=${1:-r3m}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run2() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-22s N=2 value %.0f ms %.3f e2e %.0f per_rank %s clocks %s' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value'], [round(x,3) for x in b['per_rank']['ms_per_step']], b['clocks']))
"
}
run2 default A=1
run2 clocks_off MSL_BENCH_CLOCKS=off
run2 clocks_off_nogather MSL_BENCH_CLOCKS=off MSL_BENCH_NO_GATHER=1
run2 clocks_rank0_50ms MSL_BENCH_CLOCKS=rank0 MSL_BENCH_CLOCKS_MS=50
run2 clocks_100ms MSL_BENCH_CLOCKS_MS=100
tail -c 300 $OUT/${TAG}.err
