#!/bin/bash
# Round 2, first GPU call: the new bench-scale S8 parity tests first (own timeout: k_fuse_stream has never run),
# then the whole GPU suite, the bench line, A/B lines of the fuse-kernel forms, one ncu capture of k_fuse_stream.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q > $OUT/${TAG}_s8.log 2>&1
echo "s8 exit $?" >> $OUT/${TAG}_s8.log
tail -15 $OUT/${TAG}_s8.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_s8_bench_scale_gpu.py > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench.json
i=0
for cfg in "MSL_FUSE_ONE=1" "MSL_FUSE_ONE=1 MSL_ONE_EARLY=1" "MSL_STREAM_EARLY=0" "MSL_STREAM_REGS=4" "MSL_STREAM_WAVE=2" \
           "MSL_STREAM_REGS=4 MSL_STREAM_WAVE=2" "MSL_STREAM_PF=0" "MSL_STREAM_REGS=4 MSL_STREAM_EARLY=0"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_stream -s 40 -c 1 -f -o $OUT/${TAG}_k_fuse_stream \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_fuse_stream.ncu-rep > $OUT/${TAG}_k_fuse_stream_brief.txt 2>&1
cat $OUT/${TAG}_k_fuse_stream_brief.txt
