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


def test_frame_set_upload_feeds_every_stage(oracle, msl):
    """msl_glue_upload_frames: gray + CV_16U depth uploaded once; the frame set's gray feeds the ORB extractor, its device-made
    CV_32F depth (src/Tracking.cc:205-207) equals the host conversion bit for bit and feeds the keypoint glue, its CV_16U
    depth feeds the plane pre-stage; both slots, strided host rows, the aux plane."""
    import ctypes as C
    import torch
    B, W, H = 3, 640, 480
    dev = torch.device("cuda", 0)
    gray = np.stack([S.gray_frame(40 + b) for b in range(B)])
    dd = [S.depth_frame(40 + b) for b in range(B)]
    d16 = np.stack([d[0] for d in dd])
    depth = np.stack([d[1] for d in dd])
    aux = np.arange(B * 77, dtype=np.int32)
    g = msl.FrameGlue(W, H, max_batch=B)
    h_gray, h_d16, h_aux = torch.from_numpy(gray).pin_memory(), torch.from_numpy(d16.view(np.int16)).pin_memory(), torch.from_numpy(aux).pin_memory()
    cudart = None
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            cudart = C.CDLL(name)
            break
        except OSError:
            continue
    assert cudart is not None

    def fetch(ptr_, nbytes, dtype):
        out = np.zeros(nbytes // np.dtype(dtype).itemsize, dtype)
        assert cudart.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr_), C.c_size_t(nbytes), 2) == 0
        return out

    for slot in (0, 1, 0):
        pg, p16, pd, pa = g.upload_frames(slot, h_gray.data_ptr(), h_d16.data_ptr(), B, 1.0 / 5000.0, h_aux.data_ptr(), aux.size)
        g.frames_wait(slot)
        assert np.array_equal(fetch(pg, gray.size, np.uint8).reshape(gray.shape), gray)
        assert np.array_equal(fetch(p16, d16.size * 2, np.uint16).reshape(d16.shape), d16)
        assert np.array_equal(fetch(pd, depth.size * 4, np.float32).reshape(depth.shape), depth)
        assert np.array_equal(fetch(pa, aux.size * 4, np.int32), aux)
    # consumers on their own streams, ordered by frames_wait
    orb = msl.ORBextractor(width=W, height=H, max_batch=B)
    cap = orb.capacity
    d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
    pg, p16, pd, _ = g.upload_frames(1, h_gray.data_ptr(), h_d16.data_ptr(), B, 1.0 / 5000.0)
    g.frames_wait(1, orb.stream)
    orb.extract_dev(pg, W, W * H, B, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr())
    orb.sync()
    counts = d_counts.cpu().numpy()
    for b in range(B):
        ko, do = oracle.OrbOracle()(gray[b])
        assert counts[b] == len(ko)
        assert np.array_equal(d_desc[b, :counts[b]].cpu().numpy(), do)
    # strided host rows (a cv::Mat ROI): 8 bytes / 4 pixels of padding per row
    gpad = np.zeros((B, H, W + 8), np.uint8)
    gpad[:, :, :W] = gray
    dpad = np.zeros((B, H, W + 4), np.uint16)
    dpad[:, :, :W] = d16
    pg, p16, pd, _ = g.upload_frames(0, gpad.ctypes.data, dpad.ctypes.data, B, 1.0 / 5000.0, gray_stride=W + 8, depth_stride_px=W + 4)
    g.frames_wait(0)
    assert np.array_equal(fetch(pg, gray.size, np.uint8).reshape(gray.shape), gray)
    assert np.array_equal(fetch(pd, depth.size * 4, np.float32).reshape(depth.shape), depth)
