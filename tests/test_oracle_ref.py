"""The oracle's SurfelFusion restatement against the REFERENCE'S OWN SOURCE.

oracle/_ref/libsurfel_ref.so is /root/reference/src/SurfelFusion.cpp compiled where it lies, unmodified, against stand-in
headers for OpenCV / Eigen / <thread> (oracle/ref_shim_cv/, see oracle/ref_wrap.cpp): the reference's control flow and scalar
arithmetic line for line; Eigen's products / 4x4 inverse evaluated as this repository assumes; the ten thread slices run in
order.  Built here (where /root/reference exists); elsewhere the prebuilt library is used if it travelled, else the tests
skip.  Everything is compared BIT FOR BIT (NaNs as NaNs): superpixel index, every seed field, the local map, the new surfels."""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref/libsurfel_ref.so not built and /root/reference absent")
    return oracle


def _same(a, b):
    """field by field: identical bits, except that any NaN equals any NaN (sign / payload of a NaN is a code-generation
    artefact -- x86 propagates the first operand's -- not something the reference defines)"""
    if a.shape != b.shape:
        return False
    for f in a.dtype.names:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            if not ((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))).all():
                return False
        elif not np.array_equal(x, y):
            return False
    return True


@pytest.mark.parametrize("seed,pf,n", [(3, 0.0, 30000), (4, 0.4, 50000), (5, 0.2, 10000), (6, 1.0, 5000), (7, 0.0, 0)])
def test_single_frame_matches_reference_source(ref, seed, pf, n):
    g = S.gray_frame(seed)
    _, d = S.depth_frame(seed)
    m = S.membership(seed, plane_fraction=pf)
    T = S.pose_walk(seed, 1)[0]
    local = S.surfel_map(seed, n, d, T, ref_index=20) if n else np.zeros(0, ref.SURFEL_DTYPE)
    lo, lr = local.copy(), local.copy()
    o, r = ref.SurfelOracle(), ref.RefSurfelFusion()
    new_o = o.fuse(20, g, d, m, T, lo)
    new_r = r.fuse(20, g, d, m, T, lr)
    assert np.array_equal(o.index(), r.index())
    assert _same(o.seeds(), r.seeds())
    assert _same(lo, lr) and _same(new_o, new_r)
    if pf < 1.0:
        assert len(new_o) > 50
    if n:
        assert (lo["lastUpdate"] == 20).sum() > n // 10  # a real fuse happened


def test_keyframe_stream_matches_reference_source(ref):
    """six keyframes of one scene with the SurfelMapping::fuseMap tail in between (src/SurfelMapping.cpp:366-391), incl.
    the scene whose degenerate seeds produce NaN surfels (getWeight's std::min keeps the NaN)"""
    img = S.gray_frame(21)
    frames = [S.depth_frame(21 + k, scene=21)[1] for k in range(3)]
    mem = S.membership(21)
    T0 = S.pose_walk(21, 1)[0].astype(np.float64)
    lo = S.surfel_map(21, 40001, frames[0], T0.astype(np.float32), ref_index=50)
    lr = lo.copy()
    o, r = ref.SurfelOracle(), ref.RefSurfelFusion()
    saw_nan = False
    for k, yaw in enumerate([0, 12, 25, 25, 12, 0]):
        a = np.deg2rad(yaw)
        R = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]])
        T = (T0 @ R).astype(np.float32)
        new_o = o.fuse(51 + k, img, frames[k % 3], mem, T, lo)
        new_r = r.fuse(51 + k, img, frames[k % 3], mem, T, lr)
        assert np.array_equal(o.index(), r.index()) and _same(o.seeds(), r.seeds()), k
        assert _same(lo, lr) and _same(new_o, new_r), k
        saw_nan |= bool(np.isnan(lo["weight"]).any() or np.isnan(new_o["weight"]).any())
        lo = ref.surfel_compact(lo, new_o)
        lr = lo.copy()
    assert saw_nan


def test_small_and_odd_image_sizes_match_reference_source(ref):
    """320x240 and a size whose width / 8 leaves a remainder (the clamped windows and the slice arithmetic differ)"""
    for (w, h, seed) in ((320, 240, 31), (328, 248, 32)):
        K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
        g = S.gray_frame(seed, w, h)
        _, d = S.depth_frame(seed, w, h)
        m = S.membership(seed, w, h, plane_fraction=0.2)
        T = S.pose_walk(seed, 1)[0]
        local = S.surfel_map(seed, 8000, d, T, K=K, ref_index=9, w=w, h=h)
        lo, lr = local.copy(), local.copy()
        o = ref.SurfelOracle(w, h, K[0], K[1], K[2], K[3])
        r = ref.RefSurfelFusion(w, h, K[0], K[1], K[2], K[3])
        new_o = o.fuse(10, g, d, m, T, lo)
        new_r = r.fuse(10, g, d, m, T, lr)
        assert np.array_equal(o.index(), r.index()) and _same(o.seeds(), r.seeds())
        assert _same(lo, lr) and _same(new_o, new_r)


# ---------------------------------------------------------------------------------------------------------------------
# ORBextractor: the oracle restatement against the reference's own src/ORBextractor.cc (oracle/_ref/liborb_ref.so,
# oracle/ref_orb_wrap.cpp).  OpenCV's five algorithms are the oracle's cv2-pinned primitives on both sides; the
# constructor tables, the cell loop, DistributeOctTree / DivideNode, IC_Angle, computeOrbDescriptor and the level
# bookkeeping are the reference's own code.  Everything is compared bit for bit.

@pytest.fixture(scope="module")
def ref_orb(oracle):
    if oracle.build_ref(name="liborb_ref.so") is None:
        pytest.skip("oracle/_ref/liborb_ref.so not built and /root/reference absent")
    return oracle


def _orb_same(o, r, img, nlevels):
    ko, do = o(img)
    kr, dr = r(img)
    assert len(ko) == len(kr), (len(ko), len(kr))
    for f in ko.dtype.names:
        a, b = ko[f], kr[f]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f
    assert np.array_equal(do, dr)
    for l in range(nlevels):
        assert np.array_equal(o.level_image(l), r.level_image(l)), "pyramid level %d" % l
    return ko


@pytest.mark.parametrize("nf,sf,nl,ini,mn", [(1000, 1.2, 8, 20, 7), (500, 1.2, 6, 20, 7), (2000, 1.2, 8, 20, 7),
                                             (1000, 1.5, 4, 20, 7), (1500, 1.1, 12, 15, 5), (300, 2.0, 3, 30, 10)])
def test_orb_constructor_tables_match_reference_source(ref_orb, nf, sf, nl, ini, mn):
    """mvScaleFactor (float x double member), sigma^2, inverses, mnFeaturesPerLevel (cvRound), umax
    (src/ORBextractor.cc:412-468)"""
    o = ref_orb.OrbOracle(nf, sf, nl, ini, mn)
    r = ref_orb.RefOrbExtractor(nf, sf, nl, ini, mn)
    got = o.scale_factors() + [o.features_per_level(), o.umax()]
    for a, b in zip(got, r.tables()):
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_orb_frame_matches_reference_source(ref_orb, seed):
    k = _orb_same(ref_orb.OrbOracle(), ref_orb.RefOrbExtractor(), S.gray_frame(seed), 8)
    assert len(k) > 900 and len(set(k["octave"].tolist())) == 8


@pytest.mark.parametrize("kind", ["flat", "noise", "sparse", "gradient"])
def test_orb_edge_images_match_reference_source(ref_orb, kind):
    """the same edge images as tests/test_orb_gpu.py::test_orb_edge_images"""
    r = np.random.default_rng(5)
    if kind == "flat":
        img = np.full((480, 640), 77, np.uint8)
    elif kind == "noise":
        img = r.integers(0, 256, (480, 640), dtype=np.uint8)
    elif kind == "sparse":
        img = np.full((480, 640), 100, np.uint8)
        for _ in range(12):
            x, y = r.integers(40, 600), r.integers(40, 440)
            img[y:y + 5, x:x + 5] += np.uint8(r.integers(9, 18))
    else:
        img = (np.add.outer(np.arange(480), np.arange(640)) % 256).astype(np.uint8)
    k = _orb_same(ref_orb.OrbOracle(), ref_orb.RefOrbExtractor(), img, 8)
    if kind == "flat":
        assert len(k) == 0


@pytest.mark.parametrize("w,h,nf,nl,sf", [(752, 480, 1200, 8, 1.2), (320, 240, 500, 6, 1.2), (1280, 960, 2000, 8, 1.2),
                                           (640, 480, 1000, 4, 1.5), (161, 123, 200, 3, 1.2)])
def test_orb_other_sizes_match_reference_source(ref_orb, w, h, nf, nl, sf):
    img = S.gray_frame(40 + w % 7, w, h)
    _orb_same(ref_orb.OrbOracle(nf, sf, nl), ref_orb.RefOrbExtractor(nf, sf, nl), img, nl)


def test_orb_strided_input_matches_reference_source(ref_orb):
    """a view into a wider buffer (cv::Mat step > cols)"""
    big = S.gray_frame(9, 800, 600)
    view = big[50:530, 70:710]
    assert not view.flags["C_CONTIGUOUS"]
    o, r = ref_orb.OrbOracle(), ref_orb.RefOrbExtractor()
    ko, do = o(np.ascontiguousarray(view))
    buf = big.copy()
    sub = buf[50:530, 70:710]
    n_cap = 1000 + 64 + 64
    kps, desc = np.zeros(n_cap, ref_orb.KP_DTYPE), np.zeros((n_cap, 32), np.uint8)
    n = r.L.ref_orb_extract(*r.params, sub.ctypes.data, 640, 480, buf.strides[0], kps.ctypes.data, desc.ctypes.data, n_cap,
                            None, None)
    assert n == len(ko) and kps[:n].tobytes() == ko.tobytes() and np.array_equal(desc[:n], do)


def test_orb_heap_address_tiebreak_is_the_only_freedom(ref_orb):
    """DistributeOctTree sorts (key count, node pointer) pairs (src/ORBextractor.cc:654): equal counts are ordered by heap
    address.  On the process allocator the reference therefore picks a few different keypoints than on the bump arena
    (= creation order, what the oracle restates) -- but only there: pyramid, per-level counts of the levels that never
    reach the largest-first phase, and every keypoint both runs share are identical."""
    img = S.gray_frame(2)
    ka, da = ref_orb.RefOrbExtractor(arena=True)(img)
    km, dm = ref_orb.RefOrbExtractor(arena=False)(img)
    sa = {bytes(k) + bytes(d) for k, d in zip(ka, da)}
    sm = {bytes(k) + bytes(d) for k, d in zip(km, dm)}
    # the bulk of the selection does not depend on the allocator ...
    assert len(sa & sm) > 0.95 * len(sa)
    # ... and when it does, only whole keypoints are swapped (a keypoint present in both has identical fields)
    assert abs(len(ka) - len(km)) <= 16


# ---------------------------------------------------------------------------------------------------------------------
# Plane pre-stage: the oracle restatement against the reference's own src/PlaneExtractor.cpp + include/peac/
# (oracle/_ref/libplane_ref.so, oracle/ref_plane_wrap.cpp).  OpenCV is a pixel container there; the one Eigen algorithm
# (SelfAdjointEigenSolver<Matrix3d>) is the oracle's Jacobi solver on both sides.  Bit for bit.

@pytest.fixture(scope="module")
def ref_plane(oracle):
    if oracle.build_ref(name="libplane_ref.so") is None:
        pytest.skip("oracle/_ref/libplane_ref.so not built and /root/reference absent")
    return oracle


def _plane_same(B, d16, K, fac):
    co, bo, so, eo = B.plane_prestage(d16, K=K, depth_map_factor=fac)
    cr, br, sr, er = B.ref_plane_prestage(d16, K=K, depth_map_factor=fac)
    assert co.tobytes() == cr.tobytes()  # readDepthImage, src/PlaneExtractor.cpp:44-76
    assert np.array_equal(bo["N"], br["N"]) and np.array_equal(bo["nouse"], br["nouse"])
    v = br["N"] >= 4  # center / normal of a rejected block are indeterminate in the reference
    for f in ("center", "normal", "mse", "curvature"):
        assert np.array_equal(bo[f][v].view(np.uint64), br[f][v].view(np.uint64)), f
    assert np.isnan(br["mse"][~v]).all() and np.isnan(bo["mse"][~v]).all()
    assert not (er & 16).any()  # only 4-neighbour edges exist after initGraph
    assert np.array_equal(so, sr) and np.array_equal(eo, er)
    return so, eo


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("fac", [1.0 / 5000.0, 1.0])
def test_plane_prestage_matches_reference_source(ref_plane, seed, fac):
    """fac = 1/5000 is TUM's DepthMapFactor (metres: every block passes the millimetre-scaled T_mse); fac = 1 keeps the
    synthetic depth in millimetres, the unit peac's thresholds assume, so that the seed test and T_ang really decide"""
    d16, _ = S.depth_frame(seed)
    so, eo = _plane_same(ref_plane, d16, S.K_DEFAULT, fac)
    assert 100 < so.sum() < 768 and (eo > 0).sum() > 100


@pytest.mark.parametrize("w,h", [(320, 240), (646, 486), (1280, 960), (652, 492)])
def test_plane_prestage_other_sizes_match_reference_source(ref_plane, w, h):
    """odd half sizes: ceil(cols / 2) columns, blocks only where a whole 10x10 window fits, neighbours in the partial rim"""
    K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
    d16, _ = S.depth_frame(60 + w % 5, w, h, K=K)
    _plane_same(ref_plane, d16, K, 1.0)


def test_plane_prestage_degenerate_depth_matches_reference_source(ref_plane):
    """all-zero depth (every block rejected), a constant plane, and a frame with a hole in every block"""
    z = np.zeros((480, 640), np.uint16)
    so, _ = _plane_same(ref_plane, z, S.K_DEFAULT, 1.0)
    assert so.sum() == 0
    so, eo = _plane_same(ref_plane, np.full((480, 640), 1500, np.uint16), S.K_DEFAULT, 1.0)
    assert so.all()
    holes = np.full((480, 640), 1500, np.uint16)
    holes[::20, ::20] = 0
    so, _ = _plane_same(ref_plane, holes, S.K_DEFAULT, 1.0)
    assert so.sum() == 0


def test_reference_peac_membership_has_trail_counters(ref_plane):
    """the reference's own ahCluster + refineDetails on a synthetic room: plane ids >= 0, -1, and floodFill's <= -2
    trail counters (include/peac/AHCPlaneFitter.hpp:463-467) all occur -- the value set SurfelFusion's `!= -1` test
    (src/SurfelFusion.cpp:541) has to treat as 'in a plane'"""
    d16, _ = S.depth_frame(2)
    mem, planes = ref_plane.ref_plane_run(d16, depth_map_factor=1.0)
    assert mem.shape == (240, 320) and len(planes["N"]) >= 2
    assert (np.diff(planes["N"]) <= 0).all()  # sorted by size, descending
    vals = set(np.unique(mem).tolist())
    assert -1 in vals and 0 in vals and min(vals) <= -2 and min(vals) >= -5
    # plane_vertices_[i] = the pixels labelled i
    for i, n in enumerate(planes["vertices"]):
        assert (mem == i).sum() == n
    assert np.allclose(np.linalg.norm(planes["normal"], axis=1), 1.0, atol=1e-9)


# ---------------------------------------------------------------------------------------------------------------------
# ORBmatcher: the oracle restatements against the reference's own src/ORBmatcher.cc (oracle/_ref/libmatch_ref.so,
# oracle/ref_match_wrap.cpp): compiled unmodified on top of data-only MapPoint / KeyFrame / Frame stand-ins
# (oracle/ref_shim_match/slam_standins.hpp) whose grid query is the oracle's; the cv::Mat pose products are the oracle's
# cv2-pinned gemm primitives.  Projections, windows, level rules, ratio tests, slot blocking, rotation histograms, the
# epipolar test and the chi-square gates are the reference's own code.  Match tables and counts must be identical.

@pytest.fixture(scope="module")
def ref_match(oracle):
    if oracle.build_ref(name="libmatch_ref.so") is None:
        pytest.skip("oracle/_ref/libmatch_ref.so not built and /root/reference absent")
    return oracle


def _null(a):
    """the reference stores NULL where the oracle records -3 (assigned, then reset by the rotation check)"""
    a = a.copy()
    a[a == -3] = -1
    return a


_LSF = float(np.float32(np.log(np.float64(np.float32(1.2)))))


def test_descriptor_distance_matches_reference_source(ref_match):
    r = np.random.default_rng(1)
    d = r.integers(0, 256, (200, 32), dtype=np.uint8)
    d[0], d[1] = 0, 255
    for i in range(0, 200, 2):
        assert ref_match.ref_descriptor_distance(d[i], d[i + 1]) == ref_match.descriptor_distance(d[i], d[i + 1])
    assert ref_match.ref_descriptor_distance(d[0], d[1]) == 256


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 11])
def test_search_by_projection_frame_matches_reference_source(ref_match, seed):
    """src/ORBmatcher.cc:548-678; the seeds cover bForward, bBackward and neither"""
    from manhattanslam_b200.matcher import frame_geom
    B = ref_match
    cur, last, _, Tc, Tl = S.match_scene(seed)
    for th in (15.0, 7.0):
        for chk in (False, True):
            n, cm = B.search_by_projection_frame(frame_geom(), Tc, Tl, th, chk, last, cur)
            with B.reference_matcher():
                nr, cr = B.search_by_projection_frame(frame_geom(), Tc, Tl, th, chk, last, cur)
            assert n == nr and np.array_equal(_null(cm), cr)
            assert n > 100


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_search_by_projection_points_matches_reference_source(ref_match, seed):
    """src/ORBmatcher.cc:40-124 (th == 1 takes the bFactor == false branch)"""
    from manhattanslam_b200.matcher import frame_geom
    B = ref_match
    cur, _, mps, _, _ = S.match_scene(seed)
    for th, ratio in ((1.0, 0.8), (3.0, 0.8), (5.0, 0.6)):
        n, cm = B.search_by_projection_points(frame_geom(), th, ratio, mps, cur)
        with B.reference_matcher():
            nr, cr = B.search_by_projection_points(frame_geom(), th, ratio, mps, cur)
        assert n == nr and np.array_equal(_null(cm), cr)
        assert n > 100


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_search_by_projection_keyframe_matches_reference_source(ref_match, seed):
    """src/ORBmatcher.cc:680-797 incl. MapPoint::PredictScale and the distance-invariance range"""
    from manhattanslam_b200.matcher import frame_geom
    B = ref_match
    cur, kf, Tc = S.reloc_scene(seed)
    for th, od in ((15.0, 100), (10.0, 64)):
        for chk in (False, True):
            n, cm = B.search_by_projection_keyframe(frame_geom(), Tc, th, od, chk, _LSF, kf, cur)
            with B.reference_matcher():
                nr, cr = B.search_by_projection_keyframe(frame_geom(), Tc, th, od, chk, _LSF, kf, cur)
            assert n == nr and np.array_equal(_null(cm), cr)
            assert n > 100


@pytest.mark.parametrize("seed,shuffle", [(0, False), (1, True), (2, False), (32, True)])
def test_search_by_bow_matches_reference_source(ref_match, seed, shuffle):
    """src/ORBmatcher.cc:146-255 with the real DBoW2::FeatureVector (Thirdparty/DBoW2) built from the CSR arrays"""
    B = ref_match
    kf, f = S.bow_scene(seed, shuffle=shuffle)
    for ratio in (0.7, 0.9):
        for chk in (False, True):
            n, fm = B.search_by_bow(ratio, chk, kf, f)
            with B.reference_matcher():
                nr, fr = B.search_by_bow(ratio, chk, kf, f)
            assert n == nr and np.array_equal(_null(fm), fr)
            assert n > 100


@pytest.mark.parametrize("seed", [0, 1, 2, 41])
def test_search_for_triangulation_matches_reference_source(ref_match, seed):
    """src/ORBmatcher.cc:257-406 + CheckDistEpipolarLine :127-144"""
    B = ref_match
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(seed)
    for only_stereo in (False, True):
        for chk in (False, True):
            n, m = B.search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, chk, sf, ls, kf1, kf2)
            with B.reference_matcher():
                nr, mr = B.search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, chk, sf, ls, kf1, kf2)
            assert n == nr and np.array_equal(_null(m), mr)
            assert n > 30


@pytest.mark.parametrize("seed", [0, 1, 2, 51])
def test_fuse_matches_reference_source(ref_match, seed):
    """src/ORBmatcher.cc:408-519: the reference only exposes which keypoint a map point was fused at (AddObservation) --
    the oracle's (best_idx if best_dist <= TH_LOW) -- and the count"""
    from manhattanslam_b200.matcher import frame_geom
    B = ref_match
    mps, kf, Tcw, ils = S.fuse_scene(seed)
    for th in (3.0, 5.0):
        n, bi, bd = B.fuse_search(frame_geom(), Tcw, th, _LSF, ils, mps, kf)
        nr, fi = B.ref_fuse(frame_geom(), Tcw, th, _LSF, ils, mps, kf)
        assert n == nr and np.array_equal(np.where(bd <= 50, bi, -1), fi)
        assert n > 100


# ---------------------------------------------------------------------------------------------------------------------
# f2: the oracle's restatement of ahCluster + refineDetails (oracle/peac_oracle.inc) against the
# reference's own peac (ref_plane_run: src/PlaneExtractor.cpp + include/peac/ compiled unmodified, inside the bump arena).

def _peac_same(B, d16, K=S.K_DEFAULT, fac=1.0):
    mo, po = B.plane_detect(d16, K=K, depth_map_factor=fac)
    mr, pr = B.ref_plane_run(d16, K=K, depth_map_factor=fac)
    assert np.array_equal(mo, mr)                      # PlaneFitter::membershipImg, trail counters included
    assert np.array_equal(po["N"], pr["N"]) and np.array_equal(po["vertices"], pr["vertices"])
    assert po["normal"].tobytes() == pr["normal"].tobytes() and po["center"].tobytes() == pr["center"].tobytes()
    return mo, po


@pytest.mark.parametrize("seed0", [0, 10, 20, 30])
def test_peac_cluster_and_refine_match_reference_source(ref_plane, seed0):
    nplanes = []
    for seed in range(seed0, seed0 + 10):
        d16, _ = S.depth_frame(seed)
        mem, planes = _peac_same(ref_plane, d16)
        nplanes.append(len(planes["N"]))
        assert mem.min() >= -6
    assert max(nplanes) >= 2  # multi-plane scenes were among them


def test_peac_in_metres_and_other_sizes_match_reference_source(ref_plane):
    """depth in metres (TUM's factor: every block passes the millimetre-scaled thresholds, one huge plane or none) and
    image sizes whose half resolution is not a multiple of the 10x10 window"""
    for seed in (1, 2, 3):
        d16, _ = S.depth_frame(seed)
        _peac_same(ref_plane, d16, fac=1.0 / 5000.0)
    for (w, h) in ((320, 240), (646, 486), (1280, 960)):
        K = tuple(k * (w / 640.0) for k in S.K_DEFAULT)
        d16, _ = S.depth_frame(70 + w % 7, w, h, K=K)
        _peac_same(ref_plane, d16, K=K)


def test_peac_degenerate_depth_matches_reference_source(ref_plane):
    """no valid block at all; one perfect plane (mse == 0 everywhere: the exact-tie rules of the queue and of the merge
    candidate choice decide); a frame with a hole in every block"""
    mem, planes = _peac_same(ref_plane, np.zeros((480, 640), np.uint16))
    assert len(planes["N"]) == 0 and (mem == -1).all()
    mem, planes = _peac_same(ref_plane, np.full((480, 640), 1500, np.uint16))
    assert len(planes["N"]) == 1 and planes["N"][0] == 76800
    holes = np.full((480, 640), 1500, np.uint16)
    holes[::20, ::20] = 0
    _peac_same(ref_plane, holes)
    two = np.full((480, 640), 1500, np.uint16)  # two fronto-parallel planes with a depth step: two components
    two[:, 320:] = 2500
    mem, planes = _peac_same(ref_plane, two)
    assert len(planes["N"]) == 2


# ---------------------------------------------------------------------------------------------------------------------
# f1: SurfelMapping::moveAddSurfels + fuseMap.  oracle/_ref/libmapping_ref.so is the reference's own
# src/SurfelMapping.cpp + src/SurfelFusion.cpp compiled unmodified (oracle/ref_mapping_wrap.cpp; Map / pcl / cv::FileStorage
# stand-ins in oracle/ref_shim_map/).  Keyframes go through InsertKeyFrame + ProcessNewKeyFrame as in SurfelMapping::Run; the
# oracle is driven next to it with the pose lists the reference's getAddRemovePoses produced.

def test_surfel_mapping_stream_matches_reference_source(oracle):
    """a chain of 26 keyframes (poses fall out of the 10-level drift-free window: their surfels move to mvInactiveSurfels),
    then links back into the old part of the graph (poses and their surfels move back in, the inactive vector is spliced):
    Map::mvLocalSurfels and Map::mvInactiveSurfels record for record after every keyframe"""
    B = oracle
    if B.build_ref(name="libmapping_ref.so") is None:
        pytest.skip("oracle/_ref/libmapping_ref.so not built and /root/reference absent")
    w, h = 320, 240
    K = tuple(k * 0.5 for k in S.K_DEFAULT)
    r = B.RefSurfelMapping(w, h, *K)
    o, mo = B.SurfelOracle(w, h, *K), B.SurfelMappingOracle()
    local = np.zeros(0, B.SURFEL_DTYPE)
    refs = [0] + [i - 1 for i in range(1, 26)] + [3, 26, 4, 28, 27, 2, 30, 31]
    poses = S.pose_walk(5, len(refs))
    g = S.gray_frame(5, w, h)
    mem = S.membership(5, w, h, plane_fraction=0.2)
    moved_out = moved_in = 0
    for i, ref_index in enumerate(refs):
        _, d = S.depth_frame(300 + i % 4, w, h, K=K, scene=300)
        add, rem = r.keyframe(g, d, mem, poses[i], ref_index)
        moved_out, moved_in = moved_out + len(rem), moved_in + len(add)
        if len(rem) or len(add):
            local = mo.move_add(local, rem, add)                 # moveAddSurfels :194-304
        new = o.fuse(ref_index, g, d, mem, poses[i], local)      # fuseMap :353-392 = fuseInitializeMap + the tail
        local = B.surfel_compact(local, new)
        assert _same(local, r.local()), i
        assert _same(mo.inactive(), r.inactive()), i
    assert moved_out > 20 and moved_in > 10 and len(local) > 500 and mo.inactive_size() > 1000


# ---------------------------------------------------------------------------------------------------------------------
# MapPoint: oracle/_ref/libmappoint_ref.so is the reference's own src/MapPoint.cc compiled unmodified on KeyFrame / Frame /
# Map stand-ins (oracle/ref_shim_mp/, oracle/ref_mappoint_wrap.cpp).

@pytest.fixture(scope="module")
def ref_mp(oracle):
    import ctypes as C
    so = oracle.build_ref(name="libmappoint_ref.so")
    if so is None:
        pytest.skip("oracle/_ref/libmappoint_ref.so not built and /root/reference absent")
    L = C.CDLL(so)
    L.ref_distinctive.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.ref_predict_scale.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_distinctive_descriptors_match_reference_source(ref_mp, oracle, seed):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263) of the reference itself -- observations in
    std::map<KeyFrame*, size_t> order, bad keyframes skipped, float distance matrix, sorted-row median, first least median --
    against the oracle's batched restatement"""
    r = np.random.default_rng(seed)
    n_kf, n_mp = 40, 300
    kf_rows = r.integers(5, 60, n_kf).astype(np.int32)
    kf_desc = r.integers(0, 256, (int(kf_rows.sum()), 32), dtype=np.uint8)
    # near-duplicate descriptors make medians tie: the first-least-median rule decides
    kf_desc[::3] = kf_desc[0] ^ (r.integers(0, 256, (len(kf_desc[::3]), 32), dtype=np.uint8) & r.integers(0, 2, (len(kf_desc[::3]), 32), dtype=np.uint8))
    kf_bad = (r.random(n_kf) < 0.15).astype(np.uint8)
    row0 = np.concatenate([[0], np.cumsum(kf_rows)])
    obs_off, obs_kf, obs_row, lists = [0], [], [], []
    for p in range(n_mp):
        k = int(r.choice([0, 1, 2, 3, 8, 20, 33]))
        ks = np.sort(r.choice(n_kf, min(k, n_kf), replace=False))
        rows = [int(r.integers(0, kf_rows[q])) for q in ks]
        obs_kf += ks.tolist()
        obs_row += rows
        obs_off.append(len(obs_kf))
        lists.append([kf_desc[row0[q] + rw] for q, rw in zip(ks, rows) if not kf_bad[q]])
    out = np.zeros((n_mp, 32), np.uint8)
    a = lambda x: np.ascontiguousarray(x, np.int32)
    oo, ok, orw = a(obs_off), a(obs_kf), a(obs_row)
    assert ref_mp.ref_distinctive(n_kf, kf_rows.ctypes.data, kf_desc.ctypes.data, kf_bad.ctypes.data, n_mp, oo.ctypes.data,
                                  ok.ctypes.data, orw.ctypes.data, out.ctypes.data) == 0
    bi, _ = oracle.distinctive_descriptors([np.stack(l) if l else np.zeros((0, 32), np.uint8) for l in lists])
    for p in range(n_mp):
        expect = lists[p][int(bi[p])] if lists[p] else np.zeros(32, np.uint8)
        assert np.array_equal(out[p], expect), p


def test_predict_scale_matches_reference_source(ref_mp):
    """MapPoint::PredictScale (src/MapPoint.cc:334-364): ceil(log(mfMaxDistance / dist) / mfLogScaleFactor) with float operands
    (std::log(float) = logf), clamped to [0, nLevels - 1] -- against the rule the matcher oracles and the stand-in MapPoint
    restate; plus the 0.8 / 1.2 distance-invariance getters (:324-332)"""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    libm.logf.restype, libm.logf.argtypes = C.c_float, [C.c_float]
    r = np.random.default_rng(4)
    lsf = np.float32(np.log(np.float64(np.float32(1.2))))
    for max_dist in (np.float32(3.7), np.float32(0.91), np.float32(12.25)):
        dist = np.concatenate([r.uniform(0.05, 40.0, 4000), max_dist / np.float32(1.2) ** np.arange(-2, 11)]).astype(np.float32)
        out, inv, top = np.zeros((len(dist), 2), np.int32), np.zeros(2, np.float32), np.array([3.5831808], np.float32)
        ref_mp.ref_predict_scale(float(max_dist), float(lsf), 8, top.ctypes.data, len(dist), dist.ctypes.data, out.ctypes.data, inv.ctypes.data)
        expect = np.zeros(len(dist), np.int32)
        for i, d in enumerate(dist):
            ratio = np.float32(max_dist / d)
            n = int(np.ceil(np.float32(libm.logf(float(ratio))) / lsf))
            expect[i] = min(max(n, 0), 7)
        assert np.array_equal(out[:, 0], expect) and np.array_equal(out[:, 1], expect)
        assert len(set(expect.tolist())) == 8
        assert inv[1] == np.float32(1.2) * max_dist and inv[0] == np.float32(0.8) * np.float32(max_dist / top[0])


# ---------------------------------------------------------------------------------------------------------------------
# Frame: oracle/_ref/libframe_ref.so holds six member functions of the reference's src/Frame.cc -- AssignFeaturesToGrid,
# GetFeaturesInArea, PosInGrid (the grid), UndistortKeyPoints, ComputeImageBounds,
# ComputeStereoFromRGBD (the frame glue, row f4) -- cut out of the file at build time and compiled unmodified inside a
# stand-in class (oracle/ref_frame_wrap.cpp); cv::undistortPoints is the oracle's cv2-pinned restatement.

@pytest.fixture(scope="module")
def ref_frame(oracle):
    import ctypes as C
    so = oracle.build_ref(name="libframe_ref.so")
    if so is None:
        pytest.skip("oracle/_ref/libframe_ref.so not built and /root/reference absent")
    L = C.CDLL(so)
    L.ref_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int]
    L.ref_frame_glue.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_image_bounds.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return L


def test_frame_grid_matches_reference_source(ref_frame, oracle):
    """PosInGrid rounds (so keypoints near the right / bottom edge never enter the grid), the query floors / ceils, candidates
    come back in (cell column, cell row, insertion) order -- which decides ties in every window search"""
    from manhattanslam_b200.matcher import frame_geom
    g = frame_geom()
    r = np.random.default_rng(8)
    for n in (0, 1, 300, 1000, 3000):
        xy = np.stack([r.uniform(-5, 645, n), r.uniform(-5, 485, n)], 1).astype(np.float32)
        octv = r.integers(0, 8, n).astype(np.int32)
        for _ in range(120):
            x, y = float(r.uniform(-30, 670)), float(r.uniform(-30, 510))
            rad = float(r.choice([0.5, 3.0, 7.0, 15.0, 40.0, 200.0, 900.0]))
            lo, hi = [(-1, -1), (0, 3), (2, 2), (-1, 4), (3, -1), (1, 0)][int(r.integers(0, 6))]
            got = oracle.features_in_area(g, xy, octv, x, y, rad, lo, hi)
            out = np.zeros(n + 1, np.int32)
            k = ref_frame.ref_features_in_area(g.ctypes.data, xy.ctypes.data, octv.ctypes.data, n, x, y, rad, lo, hi, out.ctypes.data, n + 1)
            assert k == len(got) and np.array_equal(out[:k], got), (n, x, y, rad, lo, hi)


@pytest.mark.parametrize("nd,distorted", [(5, True), (4, True), (5, False)])
def test_frame_glue_matches_reference_source(ref_frame, oracle, nd, distorted):
    """UndistortKeyPoints + ComputeStereoFromRGBD of the reference against the oracle's orc_undistort_keypoints /
    orc_stereo_from_rgbd (k1 == 0 copies the keypoints; 4 or 5 distortion coefficients)"""
    r = np.random.default_rng(6)
    n = 900
    kps = np.zeros(n, oracle.KP_DTYPE)
    kps["x"], kps["y"] = r.uniform(0, 639, n), r.uniform(0, 479, n)
    kps["size"], kps["angle"], kps["response"] = 31, r.uniform(0, 360, n), r.integers(7, 200, n)
    kps["octave"], kps["class_id"] = r.integers(0, 8, n), -1
    K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314][:nd] if distorted else [0.0] * nd, np.float32)
    D5 = np.zeros(5, np.float32)
    D5[:nd] = D
    depth = np.ascontiguousarray(S.depth_frame(3)[1], np.float32)
    un, ur, kd = np.zeros(n, oracle.KP_DTYPE), np.zeros(n, np.float32), np.zeros(n, np.float32)
    assert ref_frame.ref_frame_glue(n, kps.ctypes.data, K4.ctypes.data, D.ctypes.data, nd, depth.ctypes.data, 640, 480, 40.0,
                                    un.ctypes.data, ur.ctypes.data, kd.ctypes.data) == 0
    xy = np.stack([kps["x"], kps["y"]], 1)
    xy_un = oracle.undistort_keypoints(xy, K4, D5)
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), xy_un)
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(un[f], kps[f]), f
    ur_o, kd_o = oracle.stereo_from_rgbd(xy, xy_un, depth, 40.0)
    assert np.array_equal(kd, kd_o) and np.array_equal(ur, ur_o)
    assert (np.abs(xy_un - xy).max() > 0.5) == distorted


def test_image_bounds_of_reference_source(ref_frame):
    """ComputeImageBounds: the geometry every grid query is relative to -- (0, cols, 0, rows) without distortion; with it,
    the undistorted image corners (cv2-pinned undistortPoints)"""
    K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
    out = np.zeros(4, np.float32)
    Z = np.zeros(5, np.float32)
    ref_frame.ref_image_bounds(640, 480, K4.ctypes.data, Z.ctypes.data, 5, out.ctypes.data)
    assert out.tolist() == [0.0, 640.0, 0.0, 480.0]
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
    ref_frame.ref_image_bounds(640, 480, K4.ctypes.data, D.ctypes.data, 5, out.ctypes.data)
    import cv2
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    c = cv2.undistortPoints(np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32).reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    expect = [min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])]
    assert np.array_equal(out, np.array(expect, np.float32))


def test_keyframe_grid_matches_reference_source(ref_frame, oracle):
    """KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:469-504, no level filter) on the grid a KeyFrame copies from its Frame, and
    IsInImage (:540-542) -- what ORBmatcher::Fuse searches through"""
    import ctypes as C
    from manhattanslam_b200.matcher import frame_geom
    ref_frame.ref_keyframe_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                                        C.c_int, C.c_void_p]
    g = frame_geom()
    gg = {k: g[k][0] for k in g.dtype.names}
    r = np.random.default_rng(9)
    for n in (0, 500, 2000):
        xy = np.stack([r.uniform(-5, 645, n), r.uniform(-5, 485, n)], 1).astype(np.float32)
        octv = np.zeros(n, np.int32)
        for _ in range(150):
            x, y = np.float32(r.uniform(-30, 670)), np.float32(r.uniform(-30, 510))
            rad = float(r.choice([0.5, 3.0, 7.0, 15.0, 40.0, 200.0]))
            got = oracle.features_in_area(g, xy, octv, float(x), float(y), rad, -1, -1)
            out, inimg = np.zeros(n + 1, np.int32), np.zeros(1, np.int32)
            k = ref_frame.ref_keyframe_features_in_area(g.ctypes.data, xy.ctypes.data, n, float(x), float(y), rad, out.ctypes.data, n + 1,
                                                        inimg.ctypes.data)
            assert k == len(got) and np.array_equal(out[:k], got), (n, x, y, rad)
            expect = (x >= int(gg["mnMinX"])) and (x < int(gg["mnMaxX"])) and (y >= int(gg["mnMinY"])) and (y < int(gg["mnMaxY"]))
            assert bool(inimg[0]) == bool(expect)
