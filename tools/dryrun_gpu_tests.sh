#!/bin/bash
# Dry-run the GPU test scripts that have never met a GPU against an oracle-backed fake of the Python mirror
# (tools/dryrun_gpu_tests/conftest.py): catches mistakes in the test code itself, says nothing about the kernels.
#   bash tools/dryrun_gpu_tests.sh [test files...]
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
FILES=${@:-tests/test_v_reference_golden_gpu.py tests/test_widened_mappoint_gpu.py tests/test_x_peac_gpu.py tests/test_y_reference_mapping_gpu.py}
cp "$ROOT/tools/dryrun_gpu_tests/conftest.py" "$TMP/"
mkdir -p "$TMP/golden" && cp "$ROOT"/tests/golden/*.npz "$ROOT"/tests/golden/*.py "$TMP/golden/"
for f in $FILES; do cp "$ROOT/$f" "$TMP/"; done
cd "$TMP" && MSL_REPO="$ROOT" python -m pytest -q -p no:cacheprovider .
rc=$?
rm -rf "$TMP"
exit $rc
