"""GPU parity: CUDA SurfelFusion (through the C ABI) vs the CPU oracle on identical inputs.

Bar: integer fields (superpixel index, use/stable/fused flags, r/g/b, updateTimes, lastUpdate, counts)
bit-exact; float fields (seed position/normal/depth/size, surfel position/normal/size/weight/colour)
within 1e-4 relative (north_star).  The CUDA path keeps the reference's float/double operation order,
so the observed difference is 0 -- asserted as such where noted so regressions show up.
"""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu

FLOAT_SEED = ["x", "y", "size", "normX", "normY", "normZ", "posX", "posY", "posZ", "viewCos", "meanDepth",
              "meanIntensity"]
INT_SEED = ["r", "g", "b", "stable", "use"]
FLOAT_SURFEL = ["px", "py", "pz", "nx", "ny", "nz", "size", "color", "weight"]
INT_SURFEL = ["r", "g", "b", "updateTimes", "lastUpdate"]
RTOL = 1e-4


def _cmp(a, b, ffields, ifields, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    for f in ifields:
        assert np.array_equal(a[f], b[f]), "%s.%s" % (what, f)
    worst = 0.0
    for f in ffields:
        assert np.allclose(a[f], b[f], rtol=RTOL, atol=1e-6), "%s.%s" % (what, f)
        worst = max(worst, float(np.abs(a[f].astype(np.float64) - b[f]).max()) if len(a) else 0.0)
    return worst


def _frame(seed, plane_fraction=0.0):
    g = S.gray_frame(seed)
    _, d = S.depth_frame(seed)
    m = S.membership(seed, plane_fraction=plane_fraction)
    return g, d, m


@pytest.mark.parametrize("seed,pf", [(1, 0.0), (2, 0.4), (3, 0.0)])
def test_superpixels_match_oracle(oracle, msl, seed, pf):
    g, d, m = _frame(seed, pf)
    o = oracle.SurfelOracle()
    o.fuse(0, g, d, m, np.eye(4, dtype=np.float32), np.zeros(0, oracle.SURFEL_DTYPE))
    sf = msl.SurfelFusion(max_surfels=1024)
    seeds, index = sf.superpixels(g, d, m)
    assert np.array_equal(index[0], o.index()), "superpixelIndex"
    so = o.seeds()
    worst = _cmp(so, seeds[0], FLOAT_SEED, INT_SEED, "seed")
    assert worst == 0.0, worst  # same operation order => bit-exact in practice (regression guard)
    assert so["stable"].sum() > 100  # the stable/fixed-point path was exercised


def test_superpixels_batch(oracle, msl):
    B = 4
    fr = [_frame(10 + b, 0.2 * (b % 2)) for b in range(B)]
    sf = msl.SurfelFusion(max_surfels=1024)
    seeds, index = sf.superpixels(np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr]),
                                  np.stack([f[2] for f in fr]))
    o = oracle.SurfelOracle()
    for b in range(B):
        o.fuse(0, fr[b][0], fr[b][1], fr[b][2], np.eye(4, dtype=np.float32), np.zeros(0, oracle.SURFEL_DTYPE))
        assert np.array_equal(index[b], o.index())
        _cmp(o.seeds(), seeds[b], FLOAT_SEED, INT_SEED, "seed[%d]" % b)


def test_superpixels_degenerate_inputs(oracle, msl):
    """all-zero depth (no geometry), all-plane membership (no seed used), constant image."""
    sf = msl.SurfelFusion(max_surfels=1024)
    o = oracle.SurfelOracle()
    g, d, m = _frame(5)
    cases = [(g, np.zeros_like(d), m), (g, d, np.zeros_like(m)), (np.full_like(g, 9), d, m)]
    for gg, dd, mm in cases:
        o.fuse(0, gg, dd, mm, np.eye(4, dtype=np.float32), np.zeros(0, oracle.SURFEL_DTYPE))
        seeds, index = sf.superpixels(gg, dd, mm)
        assert np.array_equal(index[0], o.index())
        _cmp(o.seeds(), seeds[0], FLOAT_SEED, INT_SEED, "seed")


@pytest.mark.parametrize("n", [0, 1, 3, 1000, 200003])
def test_fuse_matches_oracle(oracle, msl, n):
    g, d, m = _frame(1)
    T = S.pose_walk(0, 3)[2]
    local = S.surfel_map(0, n, d, T, ref_index=100)
    lo = local.copy()
    o = oracle.SurfelOracle()
    new_o = o.fuse(100, g, d, m, T, lo)
    sf = msl.SurfelFusion(max_surfels=max(n, 16))
    sf.upload_map(local)
    new_g, stats = sf.fuseInitializeMap(100, g, d, m, T, compact=False)
    got = sf.download_map()
    w1 = _cmp(lo, got, FLOAT_SURFEL, INT_SURFEL, "localSurfels")
    w2 = _cmp(new_o, new_g, FLOAT_SURFEL, INT_SURFEL, "newSurfels")
    assert w1 == 0.0 and w2 == 0.0
    assert np.array_equal(o.seeds()["fused"], sf.debug_seeds()["fused"])
    n_upd = int((lo["updateTimes"] == local["updateTimes"] + 1).sum())
    n_del = int(((lo["updateTimes"] == 0) & (local["updateTimes"] != 0)).sum())
    assert stats == (len(new_o), n_upd, n_del, n)
    if n >= 100000:
        assert n_upd > n // 50 and n_del > 0


@pytest.mark.parametrize("n", [0, 5, 5000, 100001])
def test_fuse_with_compaction_stream(oracle, msl, n):
    """Several consecutive keyframes with the fuseMap tail (refill deleted slots / swap-remove) on the device."""
    K = 4
    T = S.pose_walk(3, K)
    fr = [_frame(20 + k, 0.3 if k == 2 else 0.0) for k in range(K)]
    local = S.surfel_map(1, n, fr[0][1], T[0], ref_index=50)
    o = oracle.SurfelOracle()
    lo = local.copy()
    for k in range(K):
        new = o.fuse(50 + k, fr[k][0], fr[k][1], fr[k][2], T[k], lo)
        lo = oracle.surfel_compact(lo, new)
    # frame-by-frame host API
    sf = msl.SurfelFusion(max_surfels=n + K * 4800)
    sf.upload_map(local)
    for k in range(K):
        _, stats = sf.fuseInitializeMap(50 + k, fr[k][0], fr[k][1], fr[k][2], T[k], compact=True)
    got = sf.download_map()
    assert stats[3] == len(lo) == len(got)
    assert _cmp(lo, got, FLOAT_SURFEL, INT_SURFEL, "map after stream") == 0.0
    # batched stream API (superpixels batched, fuse in order)
    sf2 = msl.SurfelFusion(max_surfels=n + K * 4800)
    sf2.upload_map(local)
    st2 = sf2.fuse_batch(50, np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr]), np.stack([f[2] for f in fr]), T)
    got2 = sf2.download_map()
    assert st2[3] == len(lo)
    assert _cmp(lo, got2, FLOAT_SURFEL, INT_SURFEL, "map after batched stream") == 0.0


def test_compaction_many_deleted(oracle, msl):
    """More deleted slots than new surfels (swap-remove path with chained tail moves)."""
    g, d, m = _frame(7)
    T = S.pose_walk(1, 1)[0]
    local = S.surfel_map(2, 30000, d, T, ref_index=100)
    r = np.random.default_rng(0)
    local["updateTimes"][r.random(len(local)) < 0.6] = 0  # 60 % already dead, many at the tail
    local["updateTimes"][-500:] = 0
    lo = local.copy()
    o = oracle.SurfelOracle()
    new = o.fuse(100, g, d, m, T, lo)
    lo = oracle.surfel_compact(lo, new)
    sf = msl.SurfelFusion(max_surfels=40000)
    sf.upload_map(local)
    _, stats = sf.fuseInitializeMap(100, g, d, m, T, compact=True)
    got = sf.download_map()
    assert len(got) == len(lo) == stats[3]
    assert _cmp(lo, got, FLOAT_SURFEL, INT_SURFEL, "compacted map") == 0.0


def test_upload_download_roundtrip(msl):
    _, d = S.depth_frame(1)
    local = S.surfel_map(0, 12345, d, np.eye(4, dtype=np.float32))
    sf = msl.SurfelFusion(max_surfels=20000)
    sf.upload_map(local)
    assert sf.map_size() == 12345
    assert np.array_equal(sf.download_map(), local)


def _pose_stamped_map(seed, n, depth, T, ref, n_poses=12):
    """A local map whose lastUpdate values spread over n_poses keyframe indices, with some dead slots."""
    r = np.random.default_rng(seed)
    local = S.surfel_map(seed, n, depth, T, ref_index=ref)
    local["lastUpdate"] = ref - r.integers(0, n_poses, n)
    local["updateTimes"][r.random(n) < 0.03] = 0
    return local


@pytest.mark.parametrize("n", [3000, 70001])
def test_move_add_surfels_matches_oracle(oracle, msl, n):
    """SurfelMapping::moveAddSurfels (src/SurfelMapping.cpp:194-304): local map and mvInactiveSurfels, record for
    record and in order, over a sequence of move-out / move-in steps interleaved with fuse + compaction."""
    img = S.gray_frame(5)
    _, depth = S.depth_frame(5)
    mem = S.membership(5)
    T = S.pose_walk(5, 1)[0]
    ref = 40
    local = _pose_stamped_map(5, n, depth, T, ref)
    mo = oracle.SurfelMappingOracle()
    so = oracle.SurfelOracle()
    sf = msl.SurfelFusion(max_surfels=3 * n + 20000)
    sf.upload_map(local)
    lo = local.copy()
    steps = [([29, 31, 33], []), ([30], [31]), ([], []), ([32, 34], [29, 30]), ([], [33, 32, 34])]
    for k, (rem, add) in enumerate(steps):
        lo = mo.move_add(lo, rem, add)
        out, inn, size = sf.moveAddSurfels(rem, add)
        got = sf.download_map()
        assert size == len(lo) == len(got)
        assert np.array_equal(got.view(np.uint8), lo.view(np.uint8)), "local map differs after step %d" % k
        assert np.array_equal(sf.download_inactive().view(np.uint8), mo.inactive().view(np.uint8)), "inactive differs after step %d" % k
        if rem:
            assert out > 0
        # a keyframe in between: fuse + compaction refill the dead slots the move-out left behind
        new = so.fuse(ref + 1 + k, img, depth, mem, T, lo)
        lo = oracle.surfel_compact(lo, new)
        sf.fuseInitializeMap(ref + 1 + k, img, depth, mem, T, compact=True)
        got = sf.download_map()
        assert len(got) == len(lo)
        for f in got.dtype.names:
            if got.dtype[f].kind == "i":
                assert np.array_equal(got[f], lo[f]), f
            else:
                assert np.allclose(got[f], lo[f], rtol=1e-4, atol=1e-6), f
        lo = got.copy()  # continue from identical bits on both sides


def test_move_add_many_poses_and_errors(oracle, msl):
    _, depth = S.depth_frame(6)
    T = S.pose_walk(6, 1)[0]
    local = _pose_stamped_map(6, 20000, depth, T, 100, n_poses=40)
    mo = oracle.SurfelMappingOracle()
    sf = msl.SurfelFusion(max_surfels=80000)
    sf.upload_map(local)
    rem = list(range(61, 101, 1))[:37]  # more than one pass of 16 poses
    lo = mo.move_add(local.copy(), rem, [])
    sf.moveAddSurfels(rem, [])
    assert np.array_equal(sf.download_map().view(np.uint8), lo.view(np.uint8))
    assert np.array_equal(sf.download_inactive().view(np.uint8), mo.inactive().view(np.uint8))
    with pytest.raises(Exception):
        sf.moveAddSurfels([], [5])          # never moved out
    with pytest.raises(Exception):
        sf.moveAddSurfels([rem[0]], [])     # already inactive
    back = rem[::-3]
    lo = mo.move_add(lo, [], back)
    out, inn, size = sf.moveAddSurfels([], back)
    assert out == 0 and inn > 0 and size == len(lo)
    assert np.array_equal(sf.download_map().view(np.uint8), lo.view(np.uint8))
    assert np.array_equal(sf.download_inactive().view(np.uint8), mo.inactive().view(np.uint8))
    # empty call is a no-op
    assert sf.moveAddSurfels([], []) == (0, 0, len(lo))


def test_scan_division_is_ieee(msl):
    """The scan's shared-reciprocal division equals div.rn.f32 bit for bit over the operand ranges it sees."""
    import ctypes as C
    sf = msl.SurfelFusion(max_surfels=2000)
    L = msl.lib()
    for seed, (lo, hi, amax) in enumerate([(0.5, 30.0, 3.0e4), (0.5, 30.0, 50.0), (0.01, 100.0, 1.0e6), (0.5, 2.0, 1e-3)]):
        bad = C.c_int64(-1)
        rc = L.msl_surfel_selftest_div(sf._h, C.c_int64(200_000_000), C.c_uint64(seed + 1), C.c_float(lo), C.c_float(hi),
                                       C.c_float(amax), C.byref(bad))
        assert rc == 0 and bad.value == 0, (lo, hi, amax, bad.value)


def test_fuse_stream_with_panning_camera(oracle, msl):
    """A camera that pans away and comes back: surfels leave and re-enter the frustum, new surfels are spawned at the
    borders, the unstable-drop rule fires on the ones left behind -- every frame's map must equal the oracle's."""
    img = S.gray_frame(21)
    frames = [S.depth_frame(21 + k, scene=21)[1] for k in range(3)]
    mem = S.membership(21)
    T0 = S.pose_walk(21, 1)[0].astype(np.float64)
    local = S.surfel_map(21, 90001, frames[0], T0.astype(np.float32), ref_index=50)
    sf = msl.SurfelFusion(max_surfels=200000)
    sf.upload_map(local)
    so = oracle.SurfelOracle()
    lo = local.copy()
    for k, yaw in enumerate([0, 12, 25, 40, 40, 25, 12, 0, -15, -30, 0, 0]):  # degrees about the camera's y axis
        a = np.deg2rad(yaw)
        R = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]])
        T = (T0 @ R).astype(np.float32)
        ref = 51 + k
        new = so.fuse(ref, img, frames[k % 3], mem, T, lo)
        lo = oracle.surfel_compact(lo, new)
        _, stats = sf.fuseInitializeMap(ref, img, frames[k % 3], mem, T, compact=True)
        got = sf.download_map()
        assert len(got) == len(lo) == stats[3], "frame %d" % k
        for f in got.dtype.names:
            if got.dtype[f].kind == "i":
                assert np.array_equal(got[f], lo[f]), (k, f)
            else:
                # this scene has two seeds whose plane fit degenerates to NaN (0/0 in the reference's arithmetic as well):
                # the NaNs must sit in the same records
                assert np.allclose(got[f], lo[f], rtol=1e-4, atol=1e-6, equal_nan=True), (k, f)
        lo = got.copy()


def test_membership_values_other_than_minus_one(oracle, msl):
    """planeMembershipImg comes out of peac's refineDetails with plane ids >= 0 AND with values <= -2 (floodFill's
    bookkeeping `trail -= 1`, include/peac/AHCPlaneFitter.hpp:463-467); SurfelFusion only tests `!= -1`
    (src/SurfelFusion.cpp:366, 543), so every value but -1 means "in a plane"."""
    g, d, m = _frame(5, 0.4)
    r = np.random.default_rng(5)
    m = np.where(m == 0, r.choice(np.array([-9, -2, 0, 3, 17], np.int32), m.shape), m).astype(np.int32)
    assert (m <= -2).any() and (m > 0).any() and (m == -1).any()
    o = oracle.SurfelOracle()
    o.fuse(0, g, d, m, np.eye(4, dtype=np.float32), np.zeros(0, oracle.SURFEL_DTYPE))
    sf = msl.SurfelFusion(max_surfels=1024)
    seeds, index = sf.superpixels(g, d, m)
    assert np.array_equal(index[0], o.index())
    assert _cmp(o.seeds(), seeds[0], FLOAT_SEED, INT_SEED, "seed") == 0.0
    # the same frame with every plane pixel set to 0 gives the same superpixels: only `!= -1` matters
    m0 = np.where(m != -1, 0, -1).astype(np.int32)
    seeds0, index0 = sf.superpixels(g, d, m0)
    assert np.array_equal(index0, index) and np.array_equal(seeds0.view(np.uint8), seeds.view(np.uint8))


@pytest.mark.parametrize("w,h", [(1280, 960), (328, 248)])
def test_fuse_stream_other_frame_sizes(oracle, msl, w, h):
    """BASELINE.json config 5 runs the front-end at 1280x960 (19,200 seeds per frame, K scaled x2): superpixel index, seeds
    and the map after a three-keyframe stream with compaction must equal the oracle at that size as well (and at a size
    whose seed grid is not a multiple of the kernels' seed groups)."""
    s = w / 640.0
    K = (525.0 * s, 525.0 * s, 319.5 * s + (s - 1) * 0.5, 239.5 * s + (s - 1) * 0.5)
    frames = [(S.gray_frame(40 + k, w, h), S.depth_frame(40 + k, w, h, K, scene=40)[1], S.membership(40 + k, w, h, plane_fraction=0.3 * (k == 1)))
              for k in range(3)]
    T = S.pose_walk(40, 3)
    local = S.surfel_map(40, 150001, frames[0][1], T[0], K, ref_index=60, w=w, h=h)
    so = oracle.SurfelOracle(w, h, *K)
    sf = msl.SurfelFusion(w, h, *K, max_surfels=len(local) + 4 * (w // 8) * (h // 8))
    seeds, index = sf.superpixels(frames[0][0], frames[0][1], frames[0][2])
    lo = local.copy()
    so.fuse(60, *frames[0], T[0], lo.copy())
    assert np.array_equal(index[0], so.index()), "superpixelIndex"
    assert _cmp(so.seeds(), seeds[0], FLOAT_SEED, INT_SEED, "seed") == 0.0
    sf.upload_map(local)
    for k in range(3):
        new = so.fuse(60 + k, *frames[k], T[k], lo)
        lo = oracle.surfel_compact(lo, new)
    st = sf.fuse_batch(60, np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), np.stack([f[2] for f in frames]), T)
    got = sf.download_map()
    assert st[3] == len(lo) == len(got)
    assert np.array_equal(got.view(np.uint8), lo.view(np.uint8)), "map after the stream is not bit-identical"


@pytest.mark.parametrize("n", [0, 7, 150003])
def test_dirty_download_equals_full_download(oracle, msl, n):
    """The exact drop-in keeps the host vector authoritative: after a fuseInitializeMap call (no compaction tail) the host copy
    patched with msl_surfel_download_changed must equal the full download -- and the oracle's map."""
    g, d, m = _frame(9)
    T = S.pose_walk(9, 1)[0]
    local = S.surfel_map(9, n, d, T, ref_index=77)
    sf = msl.SurfelFusion(max_surfels=max(n, 16))
    sf.upload_map(local)
    sf.fuseInitializeMap(77, g, d, m, T, compact=False)
    idx, rec = sf.download_changed(77)
    full = sf.download_map()
    host = local.copy()
    host[idx] = rec
    assert np.array_equal(host.view(np.uint8), full.view(np.uint8))
    assert np.all(np.diff(idx) > 0)
    lo = local.copy()
    oracle.SurfelOracle().fuse(77, g, d, m, T, lo)
    assert np.array_equal(lo.view(np.uint8), host.view(np.uint8))
    if n > 1000:
        assert 0.05 * n < len(idx) < 0.7 * n
