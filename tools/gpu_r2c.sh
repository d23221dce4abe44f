#!/bin/bash
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_s8_bench_scale_gpu.py -m gpu -q -k "pipe or default" > $OUT/${TAG}_s8.log 2>&1
echo "s8 exit $?" >> $OUT/${TAG}_s8.log
tail -15 $OUT/${TAG}_s8.log
timeout 300 python -m pytest tests/test_surfel_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
i=0
for cfg in "MSL_FUSE_ONE=2" "MSL_FUSE_ONE=4" "MSL_FUSE_ONE=4 MSL_STREAM_REGS=4" "MSL_FUSE_ONE=4 MSL_STREAM_REGS=4 MSL_STREAM_WAVE=4" \
           "MSL_FUSE_ONE=4 MSL_STREAM_WAVE=2" "MSL_FUSE_ONE=4 MSL_STREAM_EARLY=0" "MSL_FUSE_ONE=4 MSL_STREAM_REGS=4 MSL_STREAM_WAVE=4 MSL_STREAM_EARLY=0"; do
  i=$((i + 1))
  env $cfg timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ab$i.json 2>> $OUT/${TAG}_ab.err
  python tools/ab_line.py $OUT/${TAG}_ab$i.json "[$cfg]"
done
for v in "3" "4"; do
MSL_FUSE_ONE=4 MSL_STREAM_REGS=$v MSL_STREAM_WAVE=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_pipe -s 40 -c 1 -f -o $OUT/${TAG}_k_fuse_pipe$v \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_brief.py $OUT/${TAG}_k_fuse_pipe$v.ncu-rep > $OUT/${TAG}_k_fuse_pipe${v}_brief.txt 2>&1
cat $OUT/${TAG}_k_fuse_pipe${v}_brief.txt
python tools/ncu_hotspots.py $OUT/${TAG}_k_fuse_pipe$v.ncu-rep 40 > $OUT/${TAG}_k_fuse_pipe${v}_hotspots.txt 2>&1
done
NCU_FUSE_FRAMES=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_|k_[a-z_0-9]+' -f -o $OUT/${TAG}_all_kernels \
  python tools/ncu_kernels.py > $OUT/${TAG}_ncu_all.log 2>&1
tail -3 $OUT/${TAG}_ncu_all.log
python tools/ncu_brief.py $OUT/${TAG}_all_kernels.ncu-rep > $OUT/${TAG}_all_kernels_brief.txt 2>&1
grep -c "captured" $OUT/${TAG}_all_kernels_brief.txt
