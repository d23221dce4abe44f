"""CPU-side checks: the C-ABI library loads and exports every symbol include/msl_frontend.h declares
(no compute calls without a GPU), fails loudly without a device, and the host logic of the multi-GPU
path (frame sharding + count all-gather) works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "msl_frontend.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(msl_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(msl):
    from manhattanslam_b200._lib import LIB_PATH
    lib = ctypes.CDLL(LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) > 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback(msl):
    """Without a CUDA device every create call must fail loudly (MSL_ERR_CUDA), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    for ctor in (lambda: msl.ORBextractor(), lambda: msl.SurfelFusion(), lambda: msl.PlaneDetection(),
                 lambda: msl.ORBmatcher()):
        with pytest.raises(msl.MslError) as e:
            ctor()
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under manhattanslam_b200/ or include/ may reference it."""
    bad = []
    for base in ("manhattanslam_b200", "include", "adapters"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|msl_oracle\.h|libmsl_oracle|orc_[a-z_]+\(", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_struct_layouts_match_header():
    from manhattanslam_b200 import KP_DTYPE, SURFEL_DTYPE, SEED_DTYPE, BLOCK_DTYPE, GEOM_DTYPE
    assert KP_DTYPE.itemsize == 28 and SURFEL_DTYPE.itemsize == 56 and SEED_DTYPE.itemsize == 72
    assert BLOCK_DTYPE.itemsize == 72 and GEOM_DTYPE.itemsize == 116


def test_shard_plan():
    from manhattanslam_b200.sharding import shard_frames
    assert shard_frames(64, 0, 1) == (0, 64)
    parts = [shard_frames(65, r, 8) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 65
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from manhattanslam_b200.sharding import shard_frames, gather_counts
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
B = 10
lo, hi = shard_frames(B, rank, world)
# per-frame (keypoints, new surfels, updated surfels) of this rank's chunk
counts = torch.tensor([[100 * f + 1, 100 * f + 2, 100 * f + 3] for f in range(lo, hi)], dtype=torch.int32)
table = gather_counts(counts, B, rank, world)
assert table.shape == (B, 3)
assert all(int(table[f, 0]) == 100 * f + 1 and int(table[f, 2]) == 100 * f + 3 for f in range(B)), table
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_count_allgather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_adapter_bindings_typecheck():
    """every binding in adapters/ compiles (syntax + types, -Wall -Wextra clean) against stand-ins of the reference headers
    that copy the names, types, constness, declaration order and access levels of the members the bindings use
    (tools/adapter_stubs/, written from /root/reference/include/*.h; the build image has no OpenCV / Eigen)."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        import pytest
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["bash", os.path.join(root, "tools", "check_adapters.sh")], capture_output=True, text=True)
    # seven bindings against the stand-in headers (+ three of them against the reference's own headers where /root/reference exists)
    n_ok = out.stdout.count("ok ")
    assert out.returncode == 0 and n_ok == (10 if os.path.isdir("/root/reference/include") else 7), out.stdout + out.stderr
    assert "FAILED" not in out.stdout and "warning" not in out.stderr, out.stdout + out.stderr


def test_header_is_plain_c(tmp_path):
    """include/msl_frontend.h is the drop-in boundary: it must compile as C99 (plain pointers and sizes, no C++ types)
    and as C++11 (the reference's language) without warnings."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "msl_frontend.h"\nint main(void) { return MSL_OK; }\n')
    inc = os.path.join(root, "include")
    for cmd in (["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-fsyntax-only", "-I", inc, str(src)],
                ["g++", "-std=c++11", "-Wall", "-Wextra", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)]):
        out = subprocess.run(cmd, capture_output=True, text=True)
        assert out.returncode == 0 and not out.stderr.strip(), out.stderr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) prints ONE JSON line with the keys the driver reads;
    run here on a tiny configuration."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "4", "--surfels", "60000", "--cpu-frames", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rgbd_frontend_frames_per_s" and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    # "reference" where oracle/_ref/libsurfel_ref_threads.so exists (the reference's own SurfelFusion as the dominant stage)
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert (j["cpu_baseline"]["kind"] == "reference") == os.path.exists(os.path.join(root, "oracle", "_ref", "libsurfel_ref_threads.so"))
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"] == "frontend_640x480_b64_map5M_DIAGNOSTIC" and j["config"]["batch_per_gpu"] == 4
    assert j["config"]["surfels_per_gpu"] == 60000 and j["config"]["stages"] == ["orb", "hamming_match", "search_by_projection", "plane_prestage", "surfel_fuse"]
    # both arms print the same `config` dict (the driver compares them): it is built by one function from the arguments alone
    import bench
    sys_argv = sys.argv
    try:
        sys.argv = ["bench.py", "--batch", "4", "--surfels", "60000"]
        assert bench.config_of(bench.parse()) == j["config"]
    finally:
        sys.argv = sys_argv
