"""GPU parity: the vocabulary-node searches (SearchByBoW, SearchForTriangulation) and the search part of Fuse, through
the C ABI, vs the CPU oracle (src/ORBmatcher.cc:146-255, 257-406, 408-519).  All outputs are indices: bit-exact."""
import os

import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu
LSF = float(np.float32(np.log(np.float64(np.float32(1.2)))))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("seed,nnratio,shuffle", [(1, 0.7, False), (2, 0.9, True), (3, 0.6, False)])
def test_search_by_bow(oracle, msl, seed, nnratio, shuffle):
    kf, f = S.bow_scene(seed, shuffle=shuffle)
    m = msl.ORBmatcher(nnratio=nnratio)
    for check in (True, False):
        m.mbCheckOrientation = check
        n_o, fm_o = oracle.search_by_bow(nnratio, check, kf, f)
        n_g, fm_g = m.SearchByBoW(kf, f)
        assert n_o == n_g and np.array_equal(fm_o, fm_g)
        assert n_o > 150


def test_search_by_bow_edge_cases(oracle, msl):
    kf, f = S.bow_scene(7, n_kf=80, n_f=70, n_nodes=9)
    m = msl.ORBmatcher(nnratio=0.7)
    # no valid KeyFrame map point at all
    kf2 = dict(kf)
    kf2["valid"] = np.zeros_like(kf["valid"])
    n, fm = m.SearchByBoW(kf2, f)
    assert n == 0 and (fm == -1).all()
    # disjoint vocabularies: no common node
    kf3 = dict(kf)
    kf3["featvec"] = {nd + 1000000: v for nd, v in kf["featvec"].items()}
    n, fm = m.SearchByBoW(kf3, f)
    assert n == oracle.search_by_bow(0.7, True, kf3, f)[0] == 0 and (fm == -1).all()
    # one big node on both sides (more candidates than a warp has lanes; every query competes for the same slots)
    kf4, f4 = dict(kf), dict(f)
    kf4["featvec"] = {5: list(range(len(kf["valid"])))}
    f4["featvec"] = {5: list(range(len(f["angle"])))}
    n_o, fm_o = oracle.search_by_bow(0.7, True, kf4, f4)
    n_g, fm_g = m.SearchByBoW(kf4, f4)
    assert n_o == n_g and np.array_equal(fm_o, fm_g)
    # empty Frame
    f5 = {"featvec": {}, "desc": np.zeros((0, 32), np.uint8), "angle": np.zeros(0, np.float32)}
    n, fm = m.SearchByBoW(kf, f5)
    assert n == 0 and len(fm) == 0
    # malformed feature vector: index out of range
    kf6 = dict(kf)
    kf6["featvec"] = {3: [0, 100000]}
    with pytest.raises(Exception):
        m.SearchByBoW(kf6, f)


@pytest.mark.parametrize("seed", [1, 2])
def test_search_for_triangulation(oracle, msl, seed):
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(seed)
    m = msl.ORBmatcher()
    for only_stereo in (False, True):
        for check in (True, False):
            m.mbCheckOrientation = check
            n_o, m_o = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, check, sf, ls, kf1, kf2)
            n_g, m_g = m.SearchForTriangulation(kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls, bOnlyStereo=only_stereo)
            assert n_o == n_g and np.array_equal(m_o, m_g)
            assert n_o > (40 if only_stereo else 150)


def test_search_for_triangulation_edge_cases(oracle, msl):
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(5, n=120, n_nodes=6)
    m = msl.ORBmatcher()
    # degenerate fundamental matrix: den == 0 for every keypoint -> CheckDistEpipolarLine is false everywhere
    n, mm = m.SearchForTriangulation(kf1, kf2, np.zeros(9, np.float32), Cw1, Tcw2, K2, sf, ls)
    assert n == 0 and (mm == -1).all()
    # every keypoint of KeyFrame 1 already has a MapPoint
    k1 = dict(kf1)
    k1["has_mp"] = np.ones_like(kf1["has_mp"])
    assert m.SearchForTriangulation(k1, kf2, F12, Cw1, Tcw2, K2, sf, ls)[0] == 0
    # many equal candidates: KeyFrame 2 holds the same descriptor everywhere in one node -> the LAST passing one wins
    k2 = dict(kf2)
    k2["desc"] = np.tile(kf2["desc"][:1], (len(kf2["angle"]), 1))
    k1b = dict(kf1)
    k1b["desc"] = np.tile(kf2["desc"][:1], (len(kf1["angle"]), 1))
    k1b["featvec"] = {9: list(range(len(kf1["angle"])))}
    k2["featvec"] = {9: list(range(len(kf2["angle"])))}
    n_o, m_o = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, k1b, k2)
    n_g, m_g = m.SearchForTriangulation(k1b, k2, F12, Cw1, Tcw2, K2, sf, ls)
    assert n_o == n_g and np.array_equal(m_o, m_g) and n_o > 10
    with pytest.raises(Exception):
        bad = dict(kf2)
        bad["octave"] = kf2["octave"] + 100  # octave outside the scale tables
        m.SearchForTriangulation(kf1, bad, F12, Cw1, Tcw2, K2, sf, ls)


@pytest.mark.parametrize("seed,th", [(1, 3.0), (2, 5.0), (3, 1.5)])
def test_fuse_search(oracle, msl, seed, th):
    mps, kf, Tcw, ils = S.fuse_scene(seed)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    n_o, bi_o, bd_o = oracle.fuse_search(g, Tcw, th, LSF, ils, mps, kf)
    n_g, bi_g, bd_g = m.Fuse(g, Tcw, mps, kf, ils, th=th, log_scale_factor=LSF)
    assert n_o == n_g and np.array_equal(bi_o, bi_g) and np.array_equal(bd_o, bd_g)
    assert n_o > 200


def test_fuse_search_edge_cases(oracle, msl):
    mps, kf, Tcw, ils = S.fuse_scene(8, n_mp=90, n_kf=80)
    g = msl.frame_geom()
    m = msl.ORBmatcher()
    mp2 = dict(mps)
    mp2["valid"] = np.zeros_like(mps["valid"])
    n, bi, bd = m.Fuse(g, Tcw, mp2, kf, ils)
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    # single-level pyramid: PredictScale clamps to level 0
    g1 = msl.frame_geom(scale_factors=[1.0])
    kf1 = dict(kf)
    kf1["octave"] = np.zeros_like(kf["octave"])
    n_o, bi_o, bd_o = oracle.fuse_search(g1, Tcw, 3.0, LSF, ils[:1], mps, kf1)
    n_g, bi_g, bd_g = m.Fuse(g1, Tcw, mps, kf1, ils[:1], log_scale_factor=LSF)
    assert n_o == n_g and np.array_equal(bi_o, bi_g) and np.array_equal(bd_o, bd_g)
    # no keypoints in the KeyFrame
    kf0 = {"xy": np.zeros((0, 2), np.float32), "octave": np.zeros(0, np.int32), "uright": np.zeros(0, np.float32),
           "desc": np.zeros((0, 32), np.uint8)}
    n, bi, bd = m.Fuse(g, Tcw, mps, kf0, ils)
    assert n == 0 and (bi == -1).all()


def test_golden_node_searches_gpu(msl):
    gold = np.load(os.path.join(GOLD, "node_searches.npz"))
    m = msl.ORBmatcher(nnratio=0.7)
    kf, f = S.bow_scene(3)
    n, fm = m.SearchByBoW(kf, f)
    assert n == int(gold["bow_n"]) and np.array_equal(fm, gold["bow_match"])
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(3)
    n, mm = m.SearchForTriangulation(kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls)
    assert n == int(gold["tri_n"]) and np.array_equal(mm, gold["tri_match"])
    mps, kfs, Tcw, ils = S.fuse_scene(3)
    n, bi, bd = m.Fuse(msl.frame_geom(), Tcw, mps, kfs, ils, th=3.0, log_scale_factor=LSF)
    assert n == int(gold["fuse_n"]) and np.array_equal(bi, gold["fuse_idx"]) and np.array_equal(bd, gold["fuse_dist"])


def test_deferred_batch_of_mixed_searches(oracle, msl):
    """msl_matcher_batch_begin / _end: eight SearchByBoW, SearchForTriangulation, Fuse and SearchByProjection calls recorded,
    one upload, one CTA per call -- every call's result equals the oracle's (and the immediate call's); a small handle whose
    arena overflows mid-batch executes the recorded part early and still delivers everything at the end."""
    geom = msl.frame_geom()
    for cap in (4096, 64):  # 64: the arena holds about one call -> early execution inside the batch
        m = msl.ORBmatcher(nnratio=0.7, max_queries=cap, max_train=cap)
        bows = [S.bow_scene(10 + k, shuffle=bool(k & 1)) for k in range(3)]
        tris = [S.triangulation_scene(20 + k) for k in range(3)]
        fus = [S.fuse_scene(30 + k) for k in range(2)]
        cur, last, mps2, Tc, Tl = S.match_scene(41)
        m.set_timing(True)
        with m.batch():
            rb = [m.SearchByBoW(kf, f) for kf, f in bows]
            rt = [m.SearchForTriangulation(k1, k2, F12, Cw1, Tcw2, K2, sf, ls) for k1, k2, F12, Cw1, Tcw2, K2, sf, ls in tris]
            rf = [m.Fuse(geom, Tcw, mps, kfs, ils, th=3.0, log_scale_factor=LSF) for mps, kfs, Tcw, ils in fus]
            rp = m.SearchByProjectionFrame(geom, Tc, Tl, 7.0, last, cur)
            with pytest.raises(RuntimeError):
                rb[0].get()
        dev_ms, calls = m.last_execution()
        assert dev_ms > 0 and 1 <= calls <= 9  # (the last execution: an arena that fills up runs the recorded part early)
        for (kf, f), r in zip(bows, rb):
            n_o, fm_o = oracle.search_by_bow(0.7, True, kf, f)
            n_g, fm_g = r.get()
            assert n_o == n_g and np.array_equal(fm_o, fm_g)
        for (k1, k2, F12, Cw1, Tcw2, K2, sf, ls), r in zip(tris, rt):
            n_o, m_o = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, k1, k2)
            n_g, m_g = r.get()
            assert n_o == n_g and np.array_equal(m_o, m_g)
        for (mps, kfs, Tcw, ils), r in zip(fus, rf):
            n_o, bi_o, bd_o = oracle.fuse_search(geom, Tcw, 3.0, LSF, ils, mps, kfs)
            n_g, bi_g, bd_g = r.get()
            assert n_o == n_g and np.array_equal(bi_o, bi_g) and np.array_equal(bd_o, bd_g)
        n_o, cm_o = oracle.search_by_projection_frame(geom, Tc, Tl, 7.0, True, last, cur)
        n_g, cm_g = rp.get()
        assert n_o == n_g and np.array_equal(cm_o, cm_g)
        # the handle is back in immediate mode
        n_g, fm_g = m.SearchByBoW(*bows[0])
        assert (n_g, fm_g.tobytes()) == (rb[0].get()[0], rb[0].get()[1].tobytes())
        m.close()
