#!/bin/bash
# Round 2, call 3v: same-box A/B of the float-threshold comparisons against the reference's double comparisons
# (scratch/libmsl_double_compares.so = the same sources with -DMSL_DOUBLE_COMPARES).
TAG=${1:-r3v}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
LIB=manhattanslam_b200/libmsl_frontend.so
cp $LIB scratch/libmsl_float_compares.so
run() {
  name=$1
  MSL_DIAG=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --only surfel 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_diag_$name.json
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>> $OUT/${TAG}_ab.err | grep '^{' > $OUT/${TAG}_bench_$name.json
  python -c "
import json
d=json.load(open('$OUT/${TAG}_diag_$name.json')); j=json.load(open('$OUT/${TAG}_bench_$name.json'))
print('%-10s superpixel stage alone %.3f ms   step %.3f ms  %.0f frames/s  e2e %.0f' % ('$name', d['ms_per_step'], j['ms_per_step'], j['value'], j['e2e']['value']))
"
}
for i in 1 2; do
  cp scratch/libmsl_double_compares.so $LIB; run double_$i
  cp scratch/libmsl_float_compares.so $LIB; run float_$i
done
tail -c 200 $OUT/${TAG}_ab.err
