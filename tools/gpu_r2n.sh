#!/bin/bash
# Round 2, call n (2 GPUs): surfel parity after the float reciprocal in k_sp_pixels, matcher batches (8 / 32) with device
# times, then the bench at N=1 and N=2 on the same box (count table all-gathered on its own stream).
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_node_search_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --widened-only matcher > $OUT/${TAG}_widened.json 2> $OUT/${TAG}_widened.err
python - <<PY
import json
j = json.load(open("$OUT/${TAG}_widened.json"))
for k, v in j.items():
    print(k, v if not isinstance(v, dict) else {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -c 400 $OUT/${TAG}_widened.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_n2.json 2> $OUT/${TAG}_n2.err
python - <<PY
import json
for n in (1, 2):
    try:
        j = json.load(open("$OUT/${TAG}_n%d.json" % n))
        print("N=%d value %.0f ms/step %.3f e2e %.0f frac %.3f per_rank %s" % (n, j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], j.get("per_rank")))
    except Exception as e:
        print("N=%d failed: %s" % (n, e))
PY
tail -c 500 $OUT/${TAG}_n2.err
