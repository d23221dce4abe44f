"""GPU parity: frame glue (cvtColor, depth conversion, UndistortKeyPoints, ComputeStereoFromRGBD) through the C ABI
vs the CPU oracle (which tests/test_oracle_primitives.py pins against cv2)."""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu

TUM1 = ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314))
TUM3 = ((535.4, 539.2, 320.1, 247.6), (0.0, 0.0, 0.0, 0.0, 0.0))


@pytest.mark.parametrize("w,h,ch", [(640, 480, 3), (640, 480, 4), (131, 97, 3), (66, 33, 4)])
def test_cvt_gray(oracle, msl, w, h, ch):
    r = np.random.default_rng(w + ch)
    g = msl.FrameGlue(w, h, max_batch=3)
    img = r.integers(0, 256, (3, h, w, ch), dtype=np.uint8)
    for rgb in (True, False):
        out = g.cvtColor(img, rgb)
        for b in range(3):
            assert np.array_equal(out[b], oracle.cvt_gray(img[b], rgb))
    assert np.array_equal(g.cvtColor(img[1]), oracle.cvt_gray(img[1], True))  # single frame


def test_depth_to_float(oracle, msl):
    g = msl.FrameGlue(640, 480, max_batch=2)
    d16 = np.stack([S.depth_frame(1)[0], S.depth_frame(2)[0]])
    f = np.float32(1.0 / 5000.0)
    assert np.array_equal(g.depthToFloat(d16, f), oracle.depth_to_float(d16, f))
    assert np.array_equal(g.depthToFloat(d16[0], f), S.depth_frame(1)[1])
    g2 = msl.FrameGlue(37, 11)  # ragged tail, unaligned
    small = np.random.default_rng(3).integers(0, 65536, (11, 37)).astype(np.uint16)
    assert np.array_equal(g2.depthToFloat(small, 0.001), oracle.depth_to_float(small, np.float32(0.001)))


@pytest.mark.parametrize("cam", [TUM1, TUM3])
def test_keypoint_glue_on_orb_output(oracle, msl, cam):
    """ORB keypoints -> UndistortKeyPoints -> ComputeStereoFromRGBD: bit-exact (fp64 iteration, same operation order)."""
    K4, D5 = cam
    img = S.gray_frame(11)
    _, depth = S.depth_frame(11)
    kps, _ = msl.ORBextractor(width=640, height=480, max_batch=1)(img)
    g = msl.FrameGlue(640, 480)
    xy, ur, kd = g.keypoints(kps, K4, D5, depth, mbf=40.0)
    kxy = np.stack([kps["x"], kps["y"]], 1)
    xy_o = oracle.undistort_keypoints(kxy, K4, D5)
    ur_o, kd_o = oracle.stereo_from_rgbd(kxy, xy_o, depth, 40.0)
    assert np.array_equal(xy, xy_o) and np.array_equal(ur, ur_o) and np.array_equal(kd, kd_o)
    assert (kd > 0).sum() > 500 and ((kd == -1) == (ur == -1)).all()
    if D5[0] == 0:
        assert np.array_equal(xy, kxy)
    # undistortion only (no depth)
    xy2, _, _ = g.keypoints(kps, K4, D5)
    assert np.array_equal(xy2, xy_o)
    assert g.keypoints(kps[:0], K4, D5)[0].shape == (0, 2)
