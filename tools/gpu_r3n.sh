#!/bin/bash
# Round 2, call 3n (2 GPUs): host time to enqueue a step, per rank.
TAG=${1:-r3n}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nproc; lscpu | grep -E "^CPU\(s\)|NUMA|Model name" ; nvidia-smi topo -m 2>/dev/null | head -8
run2() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-22s N=2 value %.0f ms %.3f e2e %.0f per_rank %s' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value'], {k:[round(x,3) for x in v] for k,v in b['per_rank'].items()}))
"
}
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_n1.json
python -c "
import json
b=json.load(open('$OUT/${TAG}_n1.json'))
print('N=1 value %.0f ms %.3f host enqueue %.3f ms per step' % (b['value'], b['ms_per_step'], b['host_enqueue_ms_per_step']))
"
run2 default A=1
run2 nogather MSL_BENCH_NO_GATHER=1
tail -c 200 $OUT/${TAG}.err
