#!/bin/bash
# Round 2, call l: matcher with packed single-copy uploads and deferred batches (msl_matcher_batch_begin / _end).
TAG=${1:-r2l}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_node_search_gpu.py tests/test_plane_match_gpu.py tests/test_widened_mappoint_gpu.py tests/test_track_batch_gpu.py tests/test_v_reference_golden_gpu.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --widened-only matcher > $OUT/${TAG}_widened.json 2> $OUT/${TAG}_widened.err
python - <<PY
import json
j = json.load(open("$OUT/${TAG}_widened.json"))
for k, v in j.items():
    print(k, v if not isinstance(v, dict) else {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -c 400 $OUT/${TAG}_widened.err
