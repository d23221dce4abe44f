"""GPU parity: plane pre-stage and matcher (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _cmp_blocks(bo, bg):
    assert np.array_equal(bo["N"], bg["N"]) and np.array_equal(bo["nouse"], bg["nouse"])
    ok = bo["N"] >= 4
    assert np.array_equal(np.isnan(bo["mse"]), np.isnan(bg["mse"]))
    worst = 0.0
    for f in ("center", "normal", "mse", "curvature"):
        a, b = bo[f][ok], bg[f][ok]
        assert np.allclose(a, b, rtol=1e-4, atol=1e-12), f  # north_star tolerance for float fields
        worst = max(worst, float(np.abs(a - b).max()) if a.size else 0.0)
    return worst


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_plane_prestage_matches_oracle(oracle, msl, seed):
    d16, _ = S.depth_frame(seed)
    co, bo, so, eo = oracle.plane_prestage(d16)
    pd = msl.PlaneDetection(max_batch=1)
    cg, bg, sg, eg = pd.prestage(d16)
    assert np.array_equal(co, cg[0])  # cloud: same fp64 operations -> bit-exact
    worst = _cmp_blocks(bo, bg[0])
    assert worst < 1e-12
    assert np.array_equal(so, sg[0]) and np.array_equal(eo, eg[0])
    assert so.sum() > 300


def test_plane_prestage_batch_and_holes(oracle, msl):
    B = 6
    ds = np.stack([S.depth_frame(30 + b)[0] for b in range(B)])
    ds[1] = 0  # no depth at all: every block rejected
    ds[2, ::7, ::5] = 0  # scattered holes
    pd = msl.PlaneDetection(max_batch=B)
    cg, bg, sg, eg = pd.prestage(ds, want_cloud=False)
    for b in range(B):
        co, bo, so, eo = oracle.plane_prestage(ds[b])
        _cmp_blocks(bo, bg[b])
        assert np.array_equal(so, sg[b]) and np.array_equal(eo, eg[b])
    assert sg[1].sum() == 0


def test_plane_other_size(oracle, msl):
    d16, _ = S.depth_frame(4, 1280, 960)
    K = (1050.0, 1050.0, 639.5, 479.5)
    co, bo, so, eo = oracle.plane_prestage(d16, K)
    cg, bg, sg, eg = msl.PlaneDetection(1280, 960).prestage(d16, K)
    assert np.array_equal(co, cg[0]) and np.array_equal(so, sg[0]) and np.array_equal(eo, eg[0])
    _cmp_blocks(bo, bg[0])


def test_hamming_best2_and_all_pairs(oracle, msl):
    r = np.random.default_rng(0)
    B, nq, nt = 3, 777, 1003
    q = r.integers(0, 256, (B, nq, 32), dtype=np.uint8)
    t = r.integers(0, 256, (B, nt, 32), dtype=np.uint8)
    t[0, 5] = q[0, 9]  # exact match
    t[0, 700] = q[0, 9]  # duplicate: tie -> lowest index
    m = msl.ORBmatcher(max_queries=1024, max_train=1024, max_batch=B)
    d = m.hamming_all_pairs(q, t)
    ref = np.unpackbits(q[:, :, None, :] ^ t[:, None, :, :], axis=3).sum(3).astype(np.uint16)
    assert np.array_equal(d, ref)
    assert d[0, 9, 5] == 0 and oracle.descriptor_distance(q[1, 3], t[1, 4]) == ref[1, 3, 4]
    bi, bd, sd = m.hamming_best2(q, t)
    assert np.array_equal(bi, ref.argmin(2)) and np.array_equal(bd, ref.min(2))
    assert np.array_equal(sd, np.sort(ref, 2)[:, :, 1])
    assert bi[0, 9] == 5 and sd[0, 9] == 0


@pytest.mark.parametrize("seed,th", [(1, 15.0), (2, 30.0), (3, 7.0), (4, 15.0)])
def test_search_by_projection_frame(oracle, msl, seed, th):
    cur, last, mps, Tc, Tl = S.match_scene(seed)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    for check in (True, False):
        m.mbCheckOrientation = check
        n_o, cm_o = oracle.search_by_projection_frame(g, Tc, Tl, th, check, last, cur)
        n_g, cm_g = m.SearchByProjectionFrame(g, Tc, Tl, th, last, cur)
        assert n_o == n_g and np.array_equal(cm_o, cm_g)
        assert n_o > 100


@pytest.mark.parametrize("seed,th", [(1, 1.0), (2, 3.0), (5, 5.0)])
def test_search_by_projection_points(oracle, msl, seed, th):
    cur, last, mps, Tc, Tl = S.match_scene(seed, collide=0.5)
    g = msl.frame_geom()
    m = msl.ORBmatcher(nnratio=0.8)
    n_o, cm_o = oracle.search_by_projection_points(g, th, 0.8, mps, cur)
    n_g, cm_g = m.SearchByProjectionPoints(g, th, mps, cur)
    assert n_o == n_g and np.array_equal(cm_o, cm_g)
    assert n_o > 50


def test_search_edge_cases(oracle, msl):
    cur, last, mps, Tc, Tl = S.match_scene(9, n_cur=50, n_last=40)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    # no usable query at all
    last2 = dict(last)
    last2["has_mp"] = np.zeros_like(last["has_mp"])
    assert m.SearchByProjectionFrame(g, Tc, Tl, 15.0, last2, cur)[0] == oracle.search_by_projection_frame(g, Tc, Tl, 15.0, True, last2, cur)[0] == 0
    # every slot occupied on entry
    cur2 = dict(cur)
    cur2["occupied"] = np.ones_like(cur["occupied"])
    n, cm = m.SearchByProjectionFrame(g, Tc, Tl, 15.0, last, cur2)
    assert n == 0 and (cm == -2).all()


LSF = float(np.float32(np.log(np.float64(np.float32(1.2)))))  # Frame::mfLogScaleFactor (src/Frame.cc:82)


@pytest.mark.parametrize("seed,th,orb_dist", [(1, 15.0, 100), (2, 10.0, 64), (3, 3.0, 100), (4, 30.0, 50)])
def test_search_by_projection_keyframe(oracle, msl, seed, th, orb_dist):
    """Relocalisation overload (src/ORBmatcher.cc:680-797): bit-exact match table and count."""
    cur, kf, Tc = S.reloc_scene(seed, collide=0.4)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    for check in (True, False):
        m.mbCheckOrientation = check
        n_o, cm_o = oracle.search_by_projection_keyframe(g, Tc, th, orb_dist, check, LSF, kf, cur)
        n_g, cm_g = m.SearchByProjectionKeyFrame(g, Tc, th, orb_dist, kf, cur, LSF)
        assert n_o == n_g and np.array_equal(cm_o, cm_g)
        assert n_o > 50


def test_search_keyframe_edge_cases(oracle, msl):
    cur, kf, Tc = S.reloc_scene(9, n_cur=60, n_kf=50)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    kf2 = dict(kf)
    kf2["valid"] = np.zeros_like(kf["valid"])  # every point bad / already found
    assert m.SearchByProjectionKeyFrame(g, Tc, 15.0, 100, kf2, cur, LSF)[0] == 0
    cur2 = dict(cur)
    cur2["occupied"] = np.ones_like(cur["occupied"])  # every slot already holds a MapPoint
    n, cm = m.SearchByProjectionKeyFrame(g, Tc, 15.0, 100, kf, cur2, LSF)
    assert n == 0 and (cm == -2).all()
    # single-level pyramid: PredictScale clamps to level 0 and the level window is [-1, 1]
    g1 = msl.frame_geom(scale_factors=[1.0])
    n_o, cm_o = oracle.search_by_projection_keyframe(g1, Tc, 15.0, 100, True, LSF, kf, cur)
    n_g, cm_g = m.SearchByProjectionKeyFrame(g1, Tc, 15.0, 100, kf, cur, LSF)
    assert n_o == n_g and np.array_equal(cm_o, cm_g)
    with pytest.raises(Exception):
        m.SearchByProjectionKeyFrame(g, Tc, 15.0, 300, kf, cur, LSF)  # ORBdist out of range


def test_hamming_best2_ragged_device(oracle, msl):
    """Ragged batch (per-entry row counts, as produced by msl_orb_extract_dev) through the device entry point."""
    import torch
    r = np.random.default_rng(3)
    B, rows = 5, 300
    desc = r.integers(0, 256, (B, rows, 32), dtype=np.uint8)
    counts = np.array([300, 17, 0, 123, 256], np.int32)
    d_desc = torch.from_numpy(desc).cuda()
    d_counts = torch.from_numpy(counts).cuda()
    d_bi = torch.full((B, rows), -7, dtype=torch.int32, device="cuda")
    d_bd, d_sd = torch.zeros_like(d_bi), torch.zeros_like(d_bi)
    m = msl.ORBmatcher(max_queries=rows, max_train=rows, max_batch=B)
    # entry b: frame b (query) vs frame b+1 (train)
    m.hamming_best2_counts_dev(d_desc.data_ptr(), d_desc.data_ptr() + rows * 32, rows, d_counts.data_ptr(),
                               d_counts.data_ptr() + 4, B - 1, d_bi.data_ptr(), d_bd.data_ptr(), d_sd.data_ptr())
    m._L.msl_matcher_sync(m._h)
    bi, bd, sd = d_bi.cpu().numpy(), d_bd.cpu().numpy(), d_sd.cpu().numpy()
    for b in range(B - 1):
        nq, nt = counts[b], counts[b + 1]
        if nq == 0:
            continue
        if nt == 0:
            assert (bi[b, :nq] == -1).all() and (bd[b, :nq] == 256).all()
            continue
        obi, obd, osd = oracle.hamming_best2(desc[b, :nq], desc[b + 1, :nt])
        assert np.array_equal(bi[b, :nq], obi) and np.array_equal(bd[b, :nq], obd) and np.array_equal(sd[b, :nq], osd)
        assert (bi[b, nq:] == -7).all()  # rows beyond the count are not touched


def test_searches_with_more_map_points_than_the_handle_was_created_for(oracle, msl):
    """The reference passes the whole of mvpLocalMapPoints to SearchByProjection (src/Tracking.cc:1693) and the map points of
    every neighbour keyframe to Fuse (src/LocalMapping.cc:569): far more than a frame's keypoints.  The query side of a
    search grows the handle's scratch on demand; only the per-frame keypoint count is limited."""
    g = msl.frame_geom()
    cur, last, mps, Tc, Tl = S.match_scene(7, n_cur=1200, n_last=9000, collide=0.5)
    m = msl.ORBmatcher(nnratio=0.8, max_queries=512, max_train=2048)  # deliberately small
    n_o, cm_o = oracle.search_by_projection_points(g, 3.0, 0.8, mps, cur)
    n_g, cm_g = m.SearchByProjectionPoints(g, 3.0, mps, cur)
    assert n_o == n_g and np.array_equal(cm_o, cm_g) and n_o > 100
    n_o, cm_o = oracle.search_by_projection_frame(g, Tc, Tl, 7.0, True, last, cur)
    n_g, cm_g = m.SearchByProjectionFrame(g, Tc, Tl, 7.0, last, cur)
    assert n_o == n_g and np.array_equal(cm_o, cm_g)
    mpf, kfs, Tcw, ils = S.fuse_scene(3, n_mp=12000, n_kf=1500)
    n_c, bi_c, bd_c = oracle.fuse_search(g, Tcw, 3.0, LSF, ils, mpf, kfs)
    n_f, bi_f, bd_f = m.Fuse(g, Tcw, mpf, kfs, ils, th=3.0, log_scale_factor=LSF)
    assert n_c == n_f and np.array_equal(bi_c, bi_f) and np.array_equal(bd_c, bd_f) and n_c > 100
