#!/bin/bash
# Round 2, call h: whole GPU suite, default bench, the peac workload, peac phase profile after the merge-step rework.
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
for t in 512 256; do
  MSL_PEAC_THREADS=$t timeout 300 python tools/peac_time.py 64 3 > $OUT/${TAG}_peac_t${t}_b64.json 2>> $OUT/${TAG}_peac.err
  python -c "
import json
j=json.load(open('$OUT/${TAG}_peac_t${t}_b64.json'))
print('PEAC threads $t batch 64 ms/batch %.2f equal %s prof %s' % (j['ms_per_batch_min'], j['equals_oracle_first_frames'], j['profile']))
"
done
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f frac %.3f iso %.3f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r["frac"], r["isolated"]["frac"], j["parity_check"] and j["parity_check"]["check"]))
    print("  dropin", {k: v for k, v in (j.get("e2e_dropin") or {}).items() if k != "what"})
    print("  peac", (j.get("widened") or {}).get("plane_detect_640x480"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 python bench.py --workload frontend_peac_640x480_b64_map5M --steps 10 > $OUT/${TAG}_bench_peac.json 2>> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench_peac.json"))
    r = j["roofline"]
    print("BENCH-PEAC value %.0f ms %.3f e2e %.0f frac %s parity %s cpu %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r.get("frac"), j["parity_check"] and j["parity_check"]["check"], j["cpu_baseline"] and j["cpu_baseline"]["value"]))
    for k in (r.get("kernels") or [])[:8]:
        print("  %-18s n %4d avg_us %9.1f us/step %9.1f" % (k["kernel"], k["launches"], k["avg_us"], k["total_us_per_step"]))
except Exception as e:
    print("bench-peac parse failed", e)
PY
tail -c 600 $OUT/${TAG}_bench.err
du -sh $OUT
