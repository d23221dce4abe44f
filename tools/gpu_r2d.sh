#!/bin/bash
# Round 2, call d: superpixel kernels v2 (parity first), matcher scratch growth, the new bench.py (default + reference arm +
# other workloads), A/B of the superpixel forms, per-kernel ncu briefs of every kernel (text only: the report is large).
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_plane_match_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "default or superpixels_v1 or pipe_64 or two_kernel" > $OUT/${TAG}_s8.log 2>&1
echo "s8 exit $?" >> $OUT/${TAG}_s8.log
tail -8 $OUT/${TAG}_s8.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f frac %.3f iso %.3f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r["frac"], r["isolated"]["frac"], j["parity_check"]))
    for k in (r.get("kernels") or [])[:30]:
        print("  %-18s n %4d avg_us %9.1f us/step %9.1f gbs %s frac %s" % (k["kernel"], k["launches"], k["avg_us"], k["total_us_per_step"], k["achieved_gbs"] and round(k["achieved_gbs"]), k["frac"] and round(k["frac"], 3)))
    print("kernels_source", r.get("kernels_source"))
    print("cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["sample"][:200])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cut -c1-400 $OUT/${TAG}_bench_ref.json
i=0
for cfg in "MSL_SP_V2=0" "MSL_SP_V2=1" "MSL_SP_V2=1 MSL_FUSE_ONE=2 MSL_STREAM_REGS=4" "MSL_SP_V2=1 MSL_STREAM_WAVE=2"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
for o in surfel orb,match,plane; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only $o > $OUT/${TAG}_only_${o//,/_}.json 2>> $OUT/${TAG}_ab.err
  python -c "import json;j=json.load(open('$OUT/${TAG}_only_${o//,/_}.json'));print('ONLY $o ms_per_step %.3f' % j['ms_per_step'])"
done
MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel > $OUT/${TAG}_diag1.json 2>> $OUT/${TAG}_ab.err
python -c "import json;j=json.load(open('$OUT/${TAG}_diag1.json'));print('DIAG1 superpixel stage alone ms_per_step %.3f' % j['ms_per_step'])"
for wl in orb_match_640x480_b64 plane_640x480_b256 surfel_640x480_b64_map5M frontend_1280x960_b64_map5M; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > $OUT/${TAG}_wl_$wl.json 2>> $OUT/${TAG}_bench.err
  python -c "
import json
try:
    j=json.load(open('$OUT/${TAG}_wl_$wl.json')); r=j['roofline'] or {}
    print('WL $wl value %.0f ms %.3f e2e %.0f kernel %s frac %s cpu %s parity %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], r.get('kernel'), r.get('frac'), j['cpu_baseline'] and round(j['cpu_baseline']['value'],2), j.get('parity_check') and j['parity_check']['check']))
except Exception as e: print('WL $wl failed', e)
"
done
NCU_FUSE_FRAMES=2 timeout 900 ncu --set full --clock-control none -k regex:'^k_|k_[a-z_0-9]+' -f -o /tmp/${TAG}_all_kernels \
  python tools/ncu_kernels.py > $OUT/${TAG}_ncu_all.log 2>&1
tail -2 $OUT/${TAG}_ncu_all.log
python tools/ncu_brief.py /tmp/${TAG}_all_kernels.ncu-rep > $OUT/${TAG}_all_kernels_brief.txt 2>&1
grep -c "captured" $OUT/${TAG}_all_kernels_brief.txt
ls -la /tmp/${TAG}_all_kernels.ncu-rep
du -sh $OUT
