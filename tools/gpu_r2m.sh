#!/bin/bash
# Round 2, call m: register budget of k_fuse_pipe against co-residency (64 registers: room for the other streams' CTAs),
# matcher batches with pre-packed feature vectors.
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
i=0
for cfg in "MSL_STREAM_REGS=4 MSL_STREAM_WAVE=3" "MSL_STREAM_REGS=4 MSL_STREAM_WAVE=4" "MSL_STREAM_REGS=3 MSL_STREAM_WAVE=3"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
  python -c "
import json
j=json.load(open('$OUT/${TAG}_ab$i.json'))
print('   ms/step %.3f e2e %.0f host_calls %.0f' % (j['ms_per_step'], j['e2e']['value'], j['e2e_host_calls']['value']))
"
done
timeout 300 python bench.py --widened-only matcher > $OUT/${TAG}_widened.json 2> $OUT/${TAG}_widened.err
python - <<PY
import json
j = json.load(open("$OUT/${TAG}_widened.json"))
for k, v in j.items():
    print(k, v if not isinstance(v, dict) else {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -c 400 $OUT/${TAG}_widened.err
