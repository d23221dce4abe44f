#!/bin/bash
# Round 2, call g: k_fuse_pipe with three staged segments (parity at bench scale, timing, ncu), dirty download, e2e_dropin,
# peac sub-phase profile, fp64 latency microbenchmark.
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "default or pipe or two_kernel or stream" > $OUT/${TAG}_s8.log 2>&1
echo "s8 exit $?" >> $OUT/${TAG}_s8.log
tail -6 $OUT/${TAG}_s8.log
timeout 900 python -m pytest tests/test_surfel_gpu.py tests/test_x_peac_gpu.py tests/test_v_reference_golden_gpu.py tests/test_y_reference_mapping_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench.json"))
    r = j["roofline"]
    print("BENCH value %.0f ms %.3f e2e %.0f frac %.3f iso %.3f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], r["frac"], r["isolated"]["frac"], j["parity_check"] and j["parity_check"]["check"]))
    print("  dropin", j.get("e2e_dropin"))
    print("  peac", (j.get("widened") or {}).get("plane_detect_640x480"))
except Exception as e:
    print("bench parse failed", e)
PY
i=0
for cfg in "MSL_STREAM_WAVE=3" "MSL_STREAM_WAVE=2" "MSL_STREAM_EARLY=0" "MSL_FUSE_ONE=2 MSL_STREAM_REGS=4"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o /tmp/${TAG}_k_fuse_pipe \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-extras > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_brief.py /tmp/${TAG}_k_fuse_pipe.ncu-rep > $OUT/${TAG}_k_fuse_pipe_brief.txt 2>&1
cat $OUT/${TAG}_k_fuse_pipe_brief.txt | head -22
python tools/ncu_hotspots.py /tmp/${TAG}_k_fuse_pipe.ncu-rep 45 > $OUT/${TAG}_k_fuse_pipe_hotspots.txt 2>&1
for t in 256 512; do
  MSL_PEAC_THREADS=$t timeout 300 python tools/peac_time.py 64 3 > $OUT/${TAG}_peac_t${t}_b64.json 2>> $OUT/${TAG}_peac.err
  python -c "
import json
j=json.load(open('$OUT/${TAG}_peac_t${t}_b64.json'))
print('PEAC threads $t batch 64 ms/batch %.2f equal %s prof %s' % (j['ms_per_batch_min'], j['equals_oracle_first_frames'], j['profile']))
"
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I manhattanslam_b200/csrc -o /tmp/fp64_latency tools/fp64_latency.cu > /dev/null 2>&1 && /tmp/fp64_latency | tee $OUT/${TAG}_fp64_latency.txt
du -sh $OUT
