#!/bin/bash
# Round 2, call o: full default bench (per-kernel table, parity replay, widened), stage-subset diagnostics.
TAG=${1:-r2o}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f host_calls %.0f frac %.3f iso %.3f parity %s cpu %.2f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_host_calls"]["value"], r["frac"], r["isolated"]["frac"], j["parity_check"] and j["parity_check"]["check"], j["cpu_baseline"]["value"]))
    for k in (r.get("kernels") or []):
        print("  %-18s n %4d avg_us %9.1f us/step %9.1f" % (k["kernel"], k["launches"], k["avg_us"], k["total_us_per_step"]))
    print("  dropin", {k: v for k, v in (j.get("e2e_dropin") or {}).items() if k != "what"})
except Exception as e:
    print("bench parse failed", e)
PY
for o in surfel orb,match,plane; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only $o > $OUT/${TAG}_only_${o//,/_}.json 2>> $OUT/${TAG}_ab.err
  python -c "import json;j=json.load(open('$OUT/${TAG}_only_${o//,/_}.json'));print('ONLY $o ms_per_step %.3f' % j['ms_per_step'])"
done
MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1.json'));print('DIAG1 superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
tail -c 500 $OUT/${TAG}_bench.err
