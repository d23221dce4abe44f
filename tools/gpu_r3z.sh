#!/bin/bash
# Round 2 (third session), final evidence call: whole GPU suite + smoke, default bench line, reference arm, every workload, ncu launch list
# of the bench command and the ncu --set full capture of the dominant kernel.  Nothing printed under ncu is a bench value.
TAG=${1:-r3z}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
MSL_TIMELINE=$OUT/${TAG}_trace.json timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/timeline.py $OUT/${TAG}_trace.json > $OUT/${TAG}_timeline.txt 2>&1; rm -f $OUT/${TAG}_trace.json
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f host_calls %.0f frac %.3f iso %.3f traffic %s parity %s cpu %.2f launches %d" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e_host_calls"]["value"], r["frac"], r["isolated"]["frac"], r.get("traffic"), j["parity_check"] and j["parity_check"]["check"], j["cpu_baseline"]["value"], j["gpu_launches"]))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
cut -c1-300 $OUT/${TAG}_bench_reference_arm.json
for w in frontend_peac_640x480_b64_map5M orb_match_640x480_b64 plane_640x480_b256 surfel_640x480_b64_map5M frontend_1280x960_b64_map5M; do
  timeout 900 python bench.py --workload $w --steps 10 > $OUT/${TAG}_bench_$w.json 2>> $OUT/${TAG}_bench.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench_$w.json"))
    print("WL $w value %.0f ms %.3f e2e %.0f cpu %s parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["cpu_baseline"] and round(j["cpu_baseline"]["value"], 1), j.get("parity_check") and j["parity_check"]["check"]))
except Exception as e:
    print("WL $w failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
head -32 $OUT/${TAG}_launches_summary.txt
rm -f $OUT/${TAG}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o $OUT/${TAG}_k_fuse_pipe \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras >> $OUT/${TAG}_ncu_bench.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_fuse_pipe.ncu-rep > $OUT/${TAG}_k_fuse_pipe_brief.txt 2>&1
cat $OUT/${TAG}_k_fuse_pipe_brief.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_sp_pixels4 -s 1 -c 1 -f -o $OUT/${TAG}_k_sp_pixels4 \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras >> $OUT/${TAG}_ncu_bench.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_sp_pixels4.ncu-rep > $OUT/${TAG}_k_sp_pixels4_brief.txt 2>&1
cat $OUT/${TAG}_k_sp_pixels4_brief.txt
tail -c 400 $OUT/${TAG}_bench.err
