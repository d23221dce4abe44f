#!/bin/bash
# Round 2, call 3k (2 GPUs): the new S8 variants, then N = 1 and N = 2 of the default bench line back to back.
TAG=${1:-r3k}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "full_wave or grid_259 or pdl or pixels_1px or api_ctas" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
grep -E "passed|failed|exit|Error|assert" $OUT/${TAG}_pytest.log | tail -6
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_n1.json 2>> $OUT/${TAG}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_n2.json 2>> $OUT/${TAG}.err
python - <<PY
import json
a = json.load(open("$OUT/${TAG}_n1.json")); b = json.load(open("$OUT/${TAG}_n2.json"))
print("N=1 value %.0f ms %.3f e2e %.0f" % (a["value"], a["ms_per_step"], a["e2e"]["value"]))
print("N=2 value %.0f ms %.3f e2e %.0f  efficiency %.3f  e2e efficiency %.3f  per_rank %s" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["value"] / (2 * a["value"]), b["e2e"]["value"] / (2 * a["e2e"]["value"]), b.get("per_rank")))
PY
tail -c 400 $OUT/${TAG}.err
