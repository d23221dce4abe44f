"""GPU parity of fuseSurfelsKernel (src/SurfelFusion.cpp:167-283) + the fuseMap tail (src/SurfelMapping.cpp:366-391) on the
configuration bench.py measures: the bench generator's map (make_inputs), 64-frame batches through
msl_surfel_fuse_batch_dev back to back with device-resident inputs and no host synchronisation in between (the superpixel
stage of batch k+1 overlaps the fuse chain of batch k on its own stream, double-buffered), and a map large enough that the
persistent kernel's warps draw many segments each.  The downloaded map must equal the CPU oracle's record for record.

Bar (north_star): integer fields bit-exact, float fields within 1e-4 relative; the observed difference is 0 and asserted.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOAT_SURFEL = ["px", "py", "pz", "nx", "ny", "nz", "size", "color", "weight"]
INT_SURFEL = ["r", "g", "b", "updateTimes", "lastUpdate"]
W, H = 640, 480
REF0 = 100

# environment knobs of the fuse kernels (read by msl_surfel_create): every launch form the library can be switched to
VARIANTS = {
    "default": {},                                                  # k_fuse_pipe: TMA-staged segments, scan(s+1) / fuse(s) interleaved per warp
    "pipe_full_wave": {"MSL_STREAM_WAVE_BATCH": "3"},                # three CTAs per SM inside a batch too (the default there is two)
    "pipe_grid_259": {"MSL_STREAM_GRID": "259"},                     # an uneven share of CTAs per SM
    "pipe_nb3": {"MSL_PIPE_NB": "3"},                                # three staged segments per warp (two CTAs per SM)
    "pipe_pdl": {"MSL_FUSE_PDL": "1"},                               # programmatic dependent launch along the chain
    "pipe_pixels_1px": {"MSL_SP_PIX4": "0"},                         # updatePixels with one pixel per thread
    "pipe_carry": {"MSL_FUSE_CARRY": "1"},                           # full fuse rounds only, partial rounds carried in registers
    "pipe_64regs_wave4": {"MSL_STREAM_REGS": "4", "MSL_STREAM_WAVE": "4"},
    "pipe_late_loads_wave2": {"MSL_STREAM_EARLY": "0", "MSL_STREAM_WAVE": "2"},
    "pipe_no_prefetch": {"MSL_STREAM_PF": "0"},
    "stream": {"MSL_FUSE_ONE": "2"},                                # TMA-staged, phases in sequence
    "stream_64regs_late": {"MSL_FUSE_ONE": "2", "MSL_STREAM_REGS": "4", "MSL_STREAM_EARLY": "0"},
    "one": {"MSL_FUSE_ONE": "1"},                                   # round 1's kernel (direct loads)
    "one_early_wave4": {"MSL_FUSE_ONE": "1", "MSL_ONE_EARLY": "1", "MSL_ONE_WAVE": "0"},
    "two_kernel_chain": {"MSL_FUSE_ONE": "0"},
    "superpixels_v1": {"MSL_SP_V2": "0"},                           # round 1's superpixel kernels (local-memory lists)
}
KNOBS = sorted({k for v in VARIANTS.values() for k in v})

_cache = {}


def _inputs(n_surfels, batch):
    import bench
    key = (n_surfels, batch)
    if key not in _cache:
        gray, depth, mem, poses, surfels = bench.make_inputs(0, batch, n_surfels)
        _cache.clear()  # one configuration at a time: a 5 M-surfel map is 315 MB
        _cache[key] = {"in": (gray, depth, mem, poses, surfels)}
    return _cache[key]


def _oracle_stream(oracle, entry, batch, n_batches):
    """the CPU oracle over n_batches x batch frames, frame by frame: fuse (scan on all host threads: the result does not
    depend on the slicing) + the fuseMap tail, in one growing buffer (no per-frame copies of a 300 MB map)"""
    if "ref" in entry:
        return entry["ref"]
    gray, depth, mem, poses, surfels = entry["in"]
    L = oracle.lib()
    n = len(surfels)
    buf = np.zeros(n + n_batches * batch * 4800, oracle.SURFEL_DTYPE)
    buf[:n] = surfels
    so = oracle.SurfelOracle(W, H)
    threads = min(16, os.cpu_count() or 1)
    per_frame = []
    for k in range(n_batches * batch):
        b = k % batch
        before_ut = buf["updateTimes"][:n].copy() if k == n_batches * batch - 1 else None
        new = np.ascontiguousarray(so.fuse(REF0 + k, gray[b], depth[b], mem[b], poses[b], buf[:n], threads=threads))
        if before_ut is not None:
            per_frame.append(int((buf["updateTimes"][:n] == before_ut + 1).sum()))
        n = L.orc_surfel_compact(oracle._p(buf), n, oracle._p(new), len(new))
    entry["ref"] = (buf[:n].copy(), per_frame)
    return entry["ref"]


def _gpu_stream(msl, entry, batch, n_batches, env, ctas=None):
    import torch
    gray, depth, mem, poses, surfels = entry["in"]
    saved = {k: os.environ.get(k) for k in KNOBS}
    try:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        sf = msl.SurfelFusion(W, H, max_surfels=len(surfels) + (n_batches + 1) * batch * 4800)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if ctas:  # msl_surfel_set_fuse_ctas_per_sm: the launch geometry is a performance knob, never a result
        sf.set_fuse_ctas_per_sm(*ctas)
    sf.upload_map(surfels)
    dev = torch.device("cuda", 0)
    d_gray = torch.from_numpy(gray).to(dev)
    d_depth = torch.from_numpy(depth).to(dev)
    d_mem = torch.from_numpy(mem).to(dev)
    torch.cuda.synchronize()
    for k in range(n_batches):  # back to back: no host sync between the batches, exactly like bench.py's step_dev
        sf.fuse_batch_dev(REF0 + k * batch, d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), d_mem.data_ptr(), poses, batch, True)
    stats = sf.read_stats()
    got = sf.download_map()
    info = sf.launch_info()
    sf.close()
    return got, stats, info


def _check(got, ref, what):
    assert len(got) == len(ref), (what, len(got), len(ref))
    for f in INT_SURFEL:
        assert np.array_equal(got[f], ref[f]), "%s.%s" % (what, f)
    for f in FLOAT_SURFEL:
        assert np.allclose(got[f], ref[f], rtol=1e-4, atol=1e-6, equal_nan=True), "%s.%s" % (what, f)
    # the kernels keep the reference's operation order: the records are identical bit for bit (regression guard)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), what + ": not bit-identical"


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_fuse_stream_5M_two_batches(oracle, msl, variant):
    """bench.py's map (5.0 M surfels in the steady state), 2 x 64 frames: every warp of the persistent kernel draws >= 8
    segments, the per-slot counters are reset by the last CTA 128 times, the second batch's superpixels overlap the first
    batch's chain."""
    batch, n_batches, n_surfels = 64, 2, 5_000_000
    entry = _inputs(n_surfels, batch)
    ref, fused_last = _oracle_stream(oracle, entry, batch, n_batches)
    got, stats, info = _gpu_stream(msl, entry, batch, n_batches, VARIANTS[variant])
    _check(got, ref, "5M map after %d frames (%s)" % (batch * n_batches, variant))
    assert stats[3] == len(ref)
    assert 4_500_000 < len(ref) < 5_700_000
    assert fused_last[0] > 1_000_000  # the projective update really fuses ~29 % of the map per frame
    if info["kernels"] == 1 and info["persistent"]:
        draws = info["segments"] / (info["grid"] * info["warps_per_cta"])
        assert draws >= (8 if info["grid"] <= 444 else 5), info  # the persistent multi-draw loop is what is being pinned here


@pytest.mark.parametrize("variant", ["default", "two_kernel_chain", "api_ctas"])
def test_fuse_stream_1M_three_batches(oracle, msl, variant):
    batch, n_batches, n_surfels = 64, 3, 1_000_000
    entry = _inputs(n_surfels, batch)
    ref, _ = _oracle_stream(oracle, entry, batch, n_batches)
    got, stats, _ = _gpu_stream(msl, entry, batch, n_batches, VARIANTS.get(variant, {}), ctas=(1, 2) if variant == "api_ctas" else None)
    _check(got, ref, "1M map after %d frames (%s)" % (batch * n_batches, variant))
    assert stats[3] == len(ref)
