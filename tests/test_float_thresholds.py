"""manhattanslam_b200/csrc/float_thresholds.h: the superpixel kernels compare a float against a double literal ((double)x < 0.4,
src/SurfelFusion.cpp:116, :495, :716 ...) in float against the literal's neighbouring floats.  The header is compiled for the host
and the equivalence is checked on every 61st float bit pattern plus every pattern within 2^21 ulps of a literal (NaNs, infinities,
zeros, subnormals included); MSL_EXHAUSTIVE=1 checks all 2^32 patterns (0 mismatches, ~4 CPU-minutes; run when the header
changes)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_float_threshold_comparisons_equal_double_comparisons(tmp_path):
    exe = str(tmp_path / "float_thresholds_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "manhattanslam_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "host_emul", "float_thresholds_check.cpp")])
    stride = "1" if os.environ.get("MSL_EXHAUSTIVE") == "1" else "61"
    r = subprocess.run([exe, str(min(16, os.cpu_count() or 1)), stride], capture_output=True, text=True, timeout=1200)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout + r.stderr
