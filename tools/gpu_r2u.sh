#!/bin/bash
# Round 2, call u (8 GPUs): the bench at N=8 on the round-2 code (count table all-gathered on its own stream).
TAG=${1:-r2u}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_n8.json 2> $OUT/${TAG}_n8.err
python - <<PY
import json
try:
    line = [l for l in open("$OUT/${TAG}_n8.json") if l.startswith("{")][-1]
    j = json.loads(line)
    print("N=8 value %.0f ms/step %.3f e2e %.0f frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"]))
    print("per_rank", j.get("per_rank"))
except Exception as e:
    print("N=8 failed: %s" % e)
PY
tail -c 600 $OUT/${TAG}_n8.err
