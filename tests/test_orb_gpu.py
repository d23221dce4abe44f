"""GPU parity: CUDA ORB extractor (through the C ABI) vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): bit-exact for integer work (pyramid bytes, FAST scores/indices,
octree selection, octave, descriptors); <= 1e-4 relative for the float angle (observed: bit-exact).
"""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _compare_frame(oracle_orb, gpu, img, frame, kps_g, desc_g, check_stages=True):
    kps_o, desc_o = oracle_orb(img)
    if check_stages:
        for l in range(oracle_orb.nlevels):
            assert np.array_equal(oracle_orb.level_image(l), gpu.debug_level(frame, l)), "pyramid level %d" % l
            cand_o = oracle_orb.level_candidates(l)
            cand_g = gpu.debug_candidates(frame, l)
            assert cand_o.shape == cand_g.shape and np.array_equal(cand_o, cand_g), "FAST candidates level %d" % l
            bl = oracle_orb.level_image(l, blurred=True)
            if bl is not None:
                assert np.array_equal(bl, gpu.debug_level(frame, l, blurred=True)), "blurred level %d" % l
    assert len(kps_o) == len(kps_g), (len(kps_o), len(kps_g))
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps_o[f], kps_g[f]), f
    # float orientation: tolerance from the north_star (1e-4 relative); in practice bit-exact
    assert np.allclose(kps_o["angle"], kps_g["angle"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(desc_o, desc_g)
    return float(np.abs(kps_o["angle"] - kps_g["angle"]).max()) if len(kps_o) else 0.0


def test_orb_single_frame_stages(oracle, msl):
    img = S.gray_frame(1)
    gpu = msl.ORBextractor(width=640, height=480, max_batch=1)
    kps, desc = gpu(img)
    o = oracle.OrbOracle()
    err = _compare_frame(o, gpu, img, 0, kps, desc)
    assert err == 0.0  # angle is bit-exact as well
    assert len(kps) > 900
    # getters (include/ORBextractor.h:58-82)
    assert np.array_equal(gpu.GetScaleFactors(), o.scale_factors()[0])
    assert np.array_equal(gpu.GetInverseScaleSigmaSquares(), o.scale_factors()[3])


def test_orb_batch_matches_oracle(oracle, msl):
    B = 8
    imgs = S.gray_batch(10, B)
    gpu = msl.ORBextractor(width=640, height=480, max_batch=B)
    res = gpu.extract_batch(imgs)
    o = oracle.OrbOracle()
    for b in range(B):
        _compare_frame(o, gpu, imgs[b], b, res[b][0], res[b][1], check_stages=(b < 2))


@pytest.mark.parametrize("kind", ["flat", "noise", "sparse", "gradient"])
def test_orb_edge_images(oracle, msl, kind):
    r = np.random.default_rng(5)
    if kind == "flat":  # no corners at all -> 0 keypoints, empty descriptors
        img = np.full((480, 640), 77, np.uint8)
    elif kind == "noise":  # dense corners everywhere, threshold-7 fallback never used
        img = r.integers(0, 256, (480, 640), dtype=np.uint8)
    elif kind == "sparse":  # a few weak blobs: exercises the minThFAST fallback and tiny octrees
        img = np.full((480, 640), 100, np.uint8)
        for _ in range(12):
            x, y = r.integers(40, 600), r.integers(40, 440)
            img[y:y + 5, x:x + 5] += np.uint8(r.integers(9, 18))
    else:
        img = (np.add.outer(np.arange(480), np.arange(640)) % 256).astype(np.uint8)
    gpu = msl.ORBextractor(width=640, height=480, max_batch=1)
    kps, desc = gpu(img)
    _compare_frame(oracle.OrbOracle(), gpu, img, 0, kps, desc)
    if kind == "flat":
        assert len(kps) == 0 and desc.shape == (0, 32)


@pytest.mark.parametrize("w,h,nf,nl,sf", [(752, 480, 1200, 8, 1.2), (320, 240, 500, 6, 1.2), (1280, 960, 2000, 8, 1.2),
                                           (640, 480, 1000, 4, 1.5),
                                           # k_resize4's source window: four columns span exactly 12 bytes at 2.3 (the vector
                                           # form's limit), 13 at 2.6 (the one-pixel form takes over)
                                           (640, 480, 500, 3, 2.3), (640, 480, 400, 3, 2.6)])
def test_orb_other_sizes(oracle, msl, w, h, nf, nl, sf):
    img = S.gray_frame(3, w, h)
    gpu = msl.ORBextractor(nfeatures=nf, scaleFactor=sf, nlevels=nl, width=w, height=h, max_batch=1)
    kps, desc = gpu(img)
    _compare_frame(oracle.OrbOracle(nf, sf, nl), gpu, img, 0, kps, desc)


def test_orb_strided_input_and_empty(msl, oracle):
    import ctypes as C
    from manhattanslam_b200._lib import check, ptr
    big = np.zeros((480, 704), np.uint8)
    img = S.gray_frame(4)
    big[:, :640] = img
    gpu = msl.ORBextractor(width=640, height=480, max_batch=1)
    kps = np.empty(gpu.capacity, msl.KP_DTYPE)
    desc = np.empty((gpu.capacity, 32), np.uint8)
    cnt = np.zeros(1, np.int32)
    check(gpu._L.msl_orb_extract(gpu._h, ptr(big), C.c_int(704), C.c_size_t(704 * 480), C.c_int(1), ptr(kps),
                                 ptr(desc), ptr(cnt)))
    ko, do = oracle.OrbOracle()(img)
    assert cnt[0] == len(ko) and np.array_equal(desc[:cnt[0]], do)
    # _image.empty() => silent return (src/ORBextractor.cc:815-816)
    cnt[0] = 123
    check(gpu._L.msl_orb_extract(gpu._h, None, C.c_int(640), C.c_size_t(640 * 480), C.c_int(1), ptr(kps), ptr(desc),
                                 ptr(cnt)))
    assert cnt[0] == 0
