#!/bin/bash
# Round 2, call 3f: CUPTI timeline of five steps of the default workload (diagnostic; not a bench value).
TAG=${1:-r3f}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
MSL_TIMELINE=$OUT/${TAG}_trace.json timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/timeline.py $OUT/${TAG}_trace.json > $OUT/${TAG}_timeline.txt 2>&1
head -60 $OUT/${TAG}_timeline.txt
gzip -f $OUT/${TAG}_trace.json
tail -c 300 $OUT/${TAG}_bench.err
