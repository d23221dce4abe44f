"""The oracle's SurfelFusion restatement against the REFERENCE'S OWN SOURCE.

oracle/_ref/libsurfel_ref.so is /root/reference/src/SurfelFusion.cpp compiled where it lies, unmodified, against stand-in
headers for OpenCV / Eigen / <thread> (oracle/ref_shim/, see oracle/ref_wrap.cpp): the reference's control flow and scalar
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
