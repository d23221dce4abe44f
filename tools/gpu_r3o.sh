#!/bin/bash
# Round 2, call 3o (2 GPUs): is the slower second rank its scene?  N = 1 on scene 1, then N = 2 with the same scene on both
# ranks (the new default) and with a scene per rank.
TAG=${1:-r3o}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
run2() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-22s N=2 value %.0f ms %.3f e2e %.0f per_rank %s' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value'], {k:[round(x,3) for x in v] for k,v in b['per_rank'].items()}))
"
}
run1() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}.err | grep '^{' > $OUT/${TAG}_$name.json
  python -c "
import json
b=json.load(open('$OUT/${TAG}_$name.json'))
print('%-22s N=1 value %.0f ms %.3f e2e %.0f' % ('$name', b['value'], b['ms_per_step'], b['e2e']['value']))
"
}
run1 n1_scene0 A=1
run1 n1_scene1 MSL_BENCH_SCENE=1
run2 n2_same_scene A=1
run2 n2_rank_scenes MSL_BENCH_RANK_SCENES=1
tail -c 200 $OUT/${TAG}.err
