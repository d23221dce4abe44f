"""GPU parity of the batched, device-resident SearchByProjection(CurrentFrame, LastFrame, th) (BASELINE.json config 2: "ORB
extract + Hamming match ... masked SearchByProjection semantics"): extractor output -> frame glue -> Tracking::UpdateLastFrame
(src/Tracking.cc:1052-1104) + Frame::UnprojectStereo (src/Frame.cc:515-526) on the device -> ORBmatcher::SearchByProjection
(src/ORBmatcher.cc:548-678) for every consecutive pair of the batch, against the oracle fed with a numpy restatement of the
Last-frame side built from the oracle's own keypoints.  Match tables and counts bit-exact."""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def last_frame_side(oracle, kps, desc, xy_un, kdepth, Tcw, K, th_depth):
    """Tracking::UpdateLastFrame's selection + Frame::UnprojectStereo, restated: (z, index) pairs sorted, every point up to
    mThDepth and at least the 100 closest get a MapPoint at mRwc * x3Dc + mOw (cv::Mat arithmetic, pinned against cv2)."""
    fx, fy, cx, cy = (np.float32(v) for v in K)
    invfx, invfy = np.float32(1.0) / fx, np.float32(1.0) / fy
    Rcw, tcw = Tcw[:3, :3].astype(np.float32), Tcw[:3, 3].astype(np.float32)
    Rwc = np.ascontiguousarray(Rcw.T)
    Ow = oracle.cv_neg_rt_times_t(Rcw, tcw)
    n = len(kps)
    order = sorted((float(kdepth[i]), i) for i in range(n) if kdepth[i] > 0)
    has = np.zeros(n, np.uint8)
    world = np.zeros((n, 3), np.float32)
    taken = 0
    for z, i in order:
        z = np.float32(z)
        u, v = np.float32(xy_un[i, 0]), np.float32(xy_un[i, 1])
        x3 = np.array([(u - cx) * z * invfx, (v - cy) * z * invfy, z], np.float32)
        world[i] = oracle.cv_rx_plus_t(Rwc, x3, Ow)
        has[i] = 1
        taken += 1
        if z > th_depth and taken > 100:
            break
    return {"has_mp": has, "outlier": np.zeros(n, np.uint8), "mp_obs": np.zeros(n, np.uint8), "mp_world": world,
            "mp_desc": desc, "octave": kps["octave"].astype(np.int32), "angle": kps["angle"].astype(np.float32)}


@pytest.mark.parametrize("th_depth", [3.0, 0.5])
def test_search_by_projection_batch_matches_oracle(oracle, msl, th_depth):
    import torch
    B, W, H = 6, 640, 480
    K = S.K_DEFAULT
    mbf = 40.0
    # one scene under a small pose walk: consecutive frames really match
    gray = np.stack([S.gray_frame(300) for _ in range(B)])
    shift = [0, 3, 5, 9, 12, 14]
    gray = np.stack([np.roll(gray[b], shift[b], axis=1) for b in range(B)])
    depth = np.stack([S.depth_frame(300 + b, scene=300)[1] for b in range(B)])
    Twc = S.pose_walk(300, B).astype(np.float64)
    Tcw = np.stack([np.linalg.inv(Twc[b]) for b in range(B)]).astype(np.float32)
    dev = torch.device("cuda", 0)
    orb = msl.ORBextractor(width=W, height=H, max_batch=B)
    cap = orb.capacity
    d_gray, d_depth = torch.from_numpy(gray).to(dev), torch.from_numpy(depth).to(dev)
    d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
    orb.extract_dev(d_gray.data_ptr(), W, W * H, B, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr())
    glue = msl.FrameGlue(W, H, max_batch=B)
    d_xy = torch.zeros((B, cap, 2), dtype=torch.float32, device=dev)
    d_ur = torch.zeros((B, cap), dtype=torch.float32, device=dev)
    d_kd = torch.zeros((B, cap), dtype=torch.float32, device=dev)
    glue.keypoints_dev(d_kps.data_ptr(), cap, d_counts.data_ptr(), B, K, None, d_depth.data_ptr(), mbf, d_xy.data_ptr(),
                       d_ur.data_ptr(), d_kd.data_ptr(), stream=orb.stream)
    m = msl.ORBmatcher(nnratio=0.9, checkOri=True)
    g = msl.frame_geom(W, H, *K, bf=mbf)
    d_cm = torch.full((B - 1, cap), -7, dtype=torch.int32, device=dev)
    d_nm = torch.zeros(B - 1, dtype=torch.int32, device=dev)
    m.SearchByProjectionFrames_dev(g, 15.0, th_depth, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_counts.data_ptr(), B,
                                   d_xy.data_ptr(), d_ur.data_ptr(), d_kd.data_ptr(), Tcw, d_cm.data_ptr(), d_nm.data_ptr(),
                                   stream=orb.stream)
    orb.sync()
    torch.cuda.synchronize()
    counts = d_counts.cpu().numpy()
    kps_all = d_kps.cpu().numpy().view(msl.KP_DTYPE).reshape(B, cap)
    desc_all = d_desc.cpu().numpy()
    cm, nm = d_cm.cpu().numpy(), d_nm.cpu().numpy()
    total = 0
    for p in range(B - 1):
        nl, nc = int(counts[p]), int(counts[p + 1])
        kl, kc = kps_all[p, :nl], kps_all[p + 1, :nc]
        xy_l = np.stack([kl["x"], kl["y"]], 1)
        xy_c = np.stack([kc["x"], kc["y"]], 1)
        ur_l, kd_l = oracle.stereo_from_rgbd(xy_l, xy_l, depth[p], mbf)
        ur_c, kd_c = oracle.stereo_from_rgbd(xy_c, xy_c, depth[p + 1], mbf)
        last = last_frame_side(oracle, kl, desc_all[p, :nl], xy_l, kd_l, Tcw[p], K, th_depth)
        cur = {"xy": xy_c, "octave": kc["octave"].astype(np.int32), "angle": kc["angle"].astype(np.float32), "uright": ur_c,
               "desc": desc_all[p + 1, :nc], "occupied": np.zeros(nc, np.uint8)}
        n_o, cm_o = oracle.search_by_projection_frame(g, Tcw[p + 1], Tcw[p], 15.0, True, last, cur)
        assert n_o == nm[p], (p, n_o, nm[p])
        assert np.array_equal(cm_o, cm[p, :nc]), "pair %d" % p
        total += n_o
        if th_depth < 1.0:
            assert 100 < int(last["has_mp"].sum()) <= nl  # the "at least 100 closest" rule decided the selection
    assert total > 200
