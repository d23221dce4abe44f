#!/bin/bash
# Round 2, call f: batched SearchByProjection (parity), k_sp_fit2 with 5-byte entries (parity, ncu), full GPU suite,
# the bench line with the new stage, A/B lines.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_track_batch_gpu.py tests/test_surfel_gpu.py -m gpu -q > $OUT/${TAG}_new.log 2>&1
echo "new tests exit $?" >> $OUT/${TAG}_new.log
tail -15 $OUT/${TAG}_new.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_track_batch_gpu.py --deselect tests/test_surfel_gpu.py > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 400 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f frac %.3f iso %.3f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r["frac"], r["isolated"]["frac"], j["parity_check"] and j["parity_check"]["check"]))
    for k in (r.get("kernels") or [])[:32]:
        print("  %-18s n %4d avg_us %9.1f us/step %9.1f gbs %s frac %s" % (k["kernel"], k["launches"], k["avg_us"], k["total_us_per_step"], k["achieved_gbs"] and round(k["achieved_gbs"]), k["frac"] and round(k["frac"], 3)))
    print("cpu", j["cpu_baseline"]["value"])
    w = j.get("widened") or {}
    for k, v in w.items():
        if isinstance(v, dict): print("  widened", k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
except Exception as e:
    print("bench parse failed", e)
PY
i=0
for cfg in "MSL_SP_V2=0" "MSL_SP_V2=1" "MSL_SP_V2=1 MSL_STREAM_WAVE=2"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1.json'));print('DIAG1 superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only orb,match,track,plane > $OUT/${TAG}_only_orb.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_only_orb.json'));print('ONLY orb,match,track,plane ms_per_step %.3f' % j['ms_per_step'])"
NCU_FUSE_FRAMES=2 timeout 600 ncu --set full --clock-control none -k regex:'k_sp_|k_track|k_search|k_keypoint' -f -o /tmp/${TAG}_sp python tools/ncu_kernels.py > $OUT/${TAG}_ncu_sp.log 2>&1
python tools/ncu_brief.py /tmp/${TAG}_sp.ncu-rep > $OUT/${TAG}_sp_brief.txt 2>&1
grep -E "captured|gpu__time_duration|dram__bytes|issue_active|warps_active|stalls" $OUT/${TAG}_sp_brief.txt | sed 's/  */ /g' | awk '/captured/{print ""; printf "%s | ", $0; next} {printf "%s | ", $0} END{print ""}' | cut -c1-400
MSL_FUSE_ONE=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o /tmp/${TAG}_k_fuse_pipe \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_brief.py /tmp/${TAG}_k_fuse_pipe.ncu-rep > $OUT/${TAG}_k_fuse_pipe_brief.txt 2>&1
python tools/ncu_hotspots.py /tmp/${TAG}_k_fuse_pipe.ncu-rep 45 > $OUT/${TAG}_k_fuse_pipe_hotspots.txt 2>&1
du -sh $OUT
