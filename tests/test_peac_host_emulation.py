"""The f2 device algorithm (manhattanslam_b200/csrc/peac_frame.cuh: what the k_peac_frame kernel runs) compiled for the
HOST -- one "thread", no-op barriers, tests/host_emul/peac_host.cpp -- against the oracle's restatement of ahCluster +
refineDetails (which tests/test_oracle_ref.py pins to the reference's own peac).  This checks the algorithmic content of
the kernel on CPU-only machines: slot reuse, mask adjacency, libstdc++ heap order, tie rules, region-grow seeds, flood
fill, the final merge and the plane-id remap.  What it cannot check -- the barriers between the parallel phases and the
launch plumbing -- is what tests/test_x_peac_gpu.py is for.  A test harness: nothing in the product links it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

HERE = os.path.dirname(os.path.abspath(__file__))
PL = np.dtype([("normal", "<f8", 3), ("center", "<f8", 3), ("N", "<i4"), ("rid", "<i4"), ("vertices", "<i4"), ("pad", "<i4")])


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "host_emul", "peac_host.cpp")
    hdr = os.path.join(HERE, "..", "manhattanslam_b200", "csrc", "peac_frame.cuh")
    out_dir = os.path.join(HERE, "host_emul", "build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libpeac_host.so")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.peac_host_frame.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_void_p] * 5 +
                                  [C.c_int, C.c_void_p, C.c_int, C.c_int])
    return L


def _run(emu, oracle, d16, K=S.K_DEFAULT, fac=1.0, flood_serial=0):
    h, w = d16.shape
    d16 = np.ascontiguousarray(d16, np.uint16)
    _, blocks, seed, edges = oracle.plane_prestage(d16, K=K, depth_map_factor=fac)
    mem = np.zeros(((h + 1) // 2, (w + 1) // 2), np.int32)
    pl, err = np.zeros(128, PL), np.zeros(1, np.int32)
    n = emu.peac_host_frame(d16.ctypes.data, w, h, w, K[0], K[1], K[2], K[3], fac, blocks.ctypes.data, seed.ctypes.data,
                            edges.ctypes.data, mem.ctypes.data, pl.ctypes.data, 128, err.ctypes.data, 4 * mem.size, flood_serial)
    assert n >= 0 and err[0] == 0
    return mem, pl[:n]


def _same(emu, oracle, d16, K=S.K_DEFAULT, fac=1.0):
    mem, pl = _run(emu, oracle, d16, K, fac)            # region grow by levels (the default)
    mem_s, pl_s = _run(emu, oracle, d16, K, fac, 1)     # ... and as the FIFO on thread 0
    assert np.array_equal(mem, mem_s) and pl.tobytes() == pl_s.tobytes()
    mo, po = oracle.plane_detect(d16, K=K, depth_map_factor=fac, cap=128)
    assert len(pl) == len(po["N"])
    assert np.array_equal(mem, mo)
    for f in ("N", "rid", "vertices"):
        assert np.array_equal(pl[f], po[f]), f
    assert pl["normal"].tobytes() == po["normal"].tobytes() and pl["center"].tobytes() == po["center"].tobytes()
    return mem, pl


def test_shared_memory_fits_one_cta(emu):
    assert emu.peac_host_shared_bytes() <= 227 * 1024


@pytest.mark.parametrize("seed0", [0, 20, 40])
def test_device_algorithm_equals_oracle(emu, oracle, seed0):
    counts = []
    for seed in range(seed0, seed0 + 20):
        d16, _ = S.depth_frame(seed)
        _, pl = _same(emu, oracle, d16)
        counts.append(len(pl))
    assert max(counts) >= 3


def test_device_algorithm_metres_sizes_and_degenerate_depth(emu, oracle):
    for seed in (1, 2, 3):
        _same(emu, oracle, S.depth_frame(seed)[0], fac=1.0 / 5000.0)
    for (w, h) in ((320, 240), (646, 486)):
        K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
        _same(emu, oracle, S.depth_frame(70 + w % 7, w, h, K=K)[0], K=K)
    mem, pl = _same(emu, oracle, np.zeros((480, 640), np.uint16))
    assert len(pl) == 0 and (mem == -1).all()
    # one perfect plane: every block has mse == 0 -- the queue's and the merge choice's exact-tie rules decide everything
    mem, pl = _same(emu, oracle, np.full((480, 640), 1500, np.uint16))
    assert len(pl) == 1 and pl["N"][0] == 76800
    two = np.full((480, 640), 1500, np.uint16)
    two[:, 320:] = 2500
    _, pl = _same(emu, oracle, two)
    assert len(pl) == 2
    holes = np.full((480, 640), 1500, np.uint16)
    holes[::20, ::20] = 0
    _same(emu, oracle, holes)


def test_global_memory_instance_on_larger_frames(emu, oracle):
    """more than 768 blocks: the same code on SharedT<3072> (in global memory on the device) -- 800x600 and 1280x960"""
    assert emu.peac_host_shared_big_bytes() < 2 << 20
    for (w, h, seed) in ((800, 600, 4), (1280, 960, 2)):
        K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
        d16, _ = S.depth_frame(seed, w, h, K=K)
        _, pl = _same(emu, oracle, d16, K=K)
        assert len(pl) >= 2


def test_more_than_16_planes_of_equal_size(emu, oracle):
    """std::sort of extractedPlanes is not stable and libstdc++ only insertion-sorts up to 16 elements: with 22-23 planes, many
    of equal size, the plane ids in membershipImg are whatever its introsort leaves -- which the kernel source restates"""
    w, h = 1280, 960
    K = tuple(k * 2 for k in S.K_DEFAULT)
    for gx, gy in ((5, 5), (6, 4)):
        d16 = np.zeros((h, w), np.uint16)
        for i in range(gy):
            for j in range(gx):
                d16[i * h // gy:(i + 1) * h // gy, j * w // gx:(j + 1) * w // gx] = 1000 + (i * gx + j) * 150
        _, pl = _same(emu, oracle, d16, K=K)
        assert len(pl) > 16 and len(set(pl["N"].tolist())) < len(pl)
        if oracle.build_ref(name="libplane_ref.so"):  # ... and the oracle's std::sort agrees with the reference's own
            mo, po = oracle.plane_detect(d16, K=K, depth_map_factor=1.0, cap=128)
            mr, pr = oracle.ref_plane_run(d16, K=K, depth_map_factor=1.0, cap=128)
            assert np.array_equal(mo, mr) and po["center"].tobytes() == pr["center"].tobytes()


def test_frames_beyond_3072_blocks_are_refused(emu, oracle):
    d16 = np.full((1500, 2000), 1500, np.uint16)
    _, blocks, seed, edges = oracle.plane_prestage(d16)
    mem, pl, err = np.zeros((750, 1000), np.int32), np.zeros(4, PL), np.zeros(1, np.int32)
    assert emu.peac_host_frame(d16.ctypes.data, 2000, 1500, 2000, 525.0, 525.0, 319.5, 239.5, 1.0, blocks.ctypes.data,
                               seed.ctypes.data, edges.ctypes.data, mem.ctypes.data, pl.ctypes.data, 4, err.ctypes.data, 16, 0) == -1


# ---- the CTA's barriers: several real threads + a pthread barrier under ThreadSanitizer (tests/host_emul/peac_host_mt.cpp)
@pytest.fixture(scope="module")
def emu_mt():
    src = os.path.join(HERE, "host_emul", "peac_host_mt.cpp")
    hdr = os.path.join(HERE, "..", "manhattanslam_b200", "csrc", "peac_frame.cuh")
    out_dir = os.path.join(HERE, "host_emul", "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "peac_host_mt")
    if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(exe):
        r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-ffp-contract=off", "-pthread", "-o", exe, src],
                           capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("ThreadSanitizer build not available: " + r.stderr[-200:])
    return exe


@pytest.mark.parametrize("seed,threads,w,h", [(2, 8, 640, 480), (4, 5, 640, 480), (33, 16, 640, 480), (2, 8, 800, 600)])
def test_cta_phases_are_race_free_under_tsan(emu_mt, oracle, tmp_path, seed, threads, w, h):
    """every hand-over through shared memory (queue pop -> candidate fits -> selection -> mask update -> push, membership
    -> seeds -> flood fill -> final merge -> remap) must be separated by a barrier: ThreadSanitizer reports none missing,
    and the threaded run equals the oracle.  (Dropping one PEAC_SYNC() makes this test fail with TSAN reports.)"""
    K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
    d16, _ = S.depth_frame(seed, w, h, K=K)
    _, blocks, sd, ed = oracle.plane_prestage(d16, K=K, depth_map_factor=1.0)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    w2, h2 = (w + 1) // 2, (h + 1) // 2
    with open(fin, "wb") as f:
        f.write(np.array([w, h, 64], np.int32).tobytes())
        f.write(np.array(list(K) + [1.0], np.float32).tobytes())
        f.write(d16.tobytes() + blocks.tobytes() + sd.tobytes() + ed.tobytes())
    r = subprocess.run([emu_mt, str(threads), fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    raw = open(fout, "rb").read()
    cnt, err = np.frombuffer(raw[:8], np.int32)
    mem = np.frombuffer(raw[8:8 + 4 * h2 * w2], np.int32).reshape(h2, w2)
    pl = np.frombuffer(raw[8 + 4 * h2 * w2:], PL)[:cnt]
    mo, po = oracle.plane_detect(d16, K=K, depth_map_factor=1.0)
    assert err == 0 and cnt == len(po["N"]) and np.array_equal(mem, mo)
    assert np.array_equal(pl["N"], po["N"]) and np.array_equal(pl["vertices"], po["vertices"])
    assert pl["normal"].tobytes() == po["normal"].tobytes() and pl["center"].tobytes() == po["center"].tobytes()
