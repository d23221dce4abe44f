#!/usr/bin/env python3
"""Generates the committed golden vectors.  Primitive outputs come from cv2 (4.13.0 here) -- the
third-party library the reference calls (cv::resize / GaussianBlur / FAST / fastAtan2); the
full-frame vector comes from the oracle itself once the primitives are pinned."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manhattanslam_b200 import synthetic as S  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fast(img, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FastFeatureDetector_TYPE_9_16)
    return np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)], np.int32).reshape(-1, 3)


def main():
    seed = 42
    img = S.gray_frame(seed)
    r = np.random.default_rng(9)
    atan_in = r.integers(-200000, 200000, (512, 2)).astype(np.float32)
    atan_out = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in atan_in], np.float32)
    np.savez_compressed(os.path.join(HERE, "orb_primitives.npz"), seed=seed, cv2_version=cv2.__version__,
                        resized=cv2.resize(img, (533, 400), interpolation=cv2.INTER_LINEAR),
                        blurred=cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101),
                        fast20=fast(img, 20), fast7_roi=fast(np.ascontiguousarray(img[100:137, 200:237]), 7),
                        atan_in=atan_in, atan_out=atan_out)
    kps, desc = ob.OrbOracle()(img)
    np.savez_compressed(os.path.join(HERE, "orb_frame.npz"), seed=seed, kps=kps.view(np.uint8).reshape(len(kps), -1),
                        desc=desc)
    print("golden written; cv2", cv2.__version__, "kps", len(kps))
    # frame glue: cvtColor + undistortPoints (TUM1.yaml camera)
    rgb = r.integers(0, 256, (33, 64, 3), dtype=np.uint8)
    K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
    D5 = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    pts = np.stack([r.uniform(-20, 660, 1000), r.uniform(-20, 500, 1000)], 1).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "glue.npz"), cv2_version=cv2.__version__, rgb=rgb,
                        gray_rgb=cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), gray_bgr=cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY),
                        K4=K4, D5=D5, pts=pts, undist=cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D5, None, K).reshape(-1, 2))


if __name__ == "__main__":
    main()


def stages():
    """Stage-level golden vectors produced by the oracle (self-consistency across compilers/builds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_stages import _surfel_case
    from manhattanslam_b200.matcher import frame_geom
    g, d, m, T, local = _surfel_case()
    lo = local.copy()
    o = ob.SurfelOracle()
    new = o.fuse(20, g, d, m, T, lo)
    d16, _ = S.depth_frame(3)
    cloud, blocks, seed, edges = ob.plane_prestage(d16)
    cur, last, mps, Tc, Tl = S.match_scene(3)
    n, cm = ob.search_by_projection_frame(frame_geom(), Tc, Tl, 15.0, True, last, cur)
    np.savez_compressed(os.path.join(HERE, "stages.npz"), sp_index=o.index(), sp_seeds=o.seeds().view(np.uint8),
                        local_after=lo.view(np.uint8), new=new.view(np.uint8), plane_seed=seed, plane_edges=edges,
                        plane_normals=blocks["normal"][seed == 1], match_n=n, match_cm=cm)
    print("stage golden written")
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    cur, kf, Tc = S.reloc_scene(3)
    n, cm = ob.search_by_projection_keyframe(frame_geom(), Tc, 15.0, 100, True, lsf, kf, cur)
    np.savez_compressed(os.path.join(HERE, "reloc.npz"), match_n=n, match_cm=cm)
    print("reloc golden written", n)


if __name__ == "__main__":
    stages()


def node_searches():
    """SearchByBoW / SearchForTriangulation / Fuse golden vectors (oracle output, after the oracle has been cross-checked
    against the independent Python restatements of tests/test_oracle_stages.py)."""
    from manhattanslam_b200.matcher import frame_geom
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    kf, f = S.bow_scene(3)
    nb, fm = ob.search_by_bow(0.7, True, kf, f)
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(3)
    nt, m12 = ob.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, kf1, kf2)
    mps, kfs, Tcw, ils = S.fuse_scene(3)
    nf, bi, bd = ob.fuse_search(frame_geom(), Tcw, 3.0, lsf, ils, mps, kfs)
    ddi, ddm = ob.distinctive_descriptors(S.observation_sets(3))
    np.savez_compressed(os.path.join(HERE, "node_searches.npz"), bow_n=nb, bow_match=fm, tri_n=nt, tri_match=m12, fuse_n=nf,
                        fuse_idx=bi, fuse_dist=bd, dd_idx=ddi, dd_median=ddm)
    print("node search golden written", nb, nt, nf)


if __name__ == "__main__":
    node_searches()


def cv_gemm():
    """cv::gemm / cv::norm outputs of cv2 for the 3x3 / 3x1 products of the matcher (see test_cv_gemm_semantics)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_primitives import _cv_gemm_cases
    Ts, xs, a, b, c, n = [], [], [], [], [], []
    for T, x in _cv_gemm_cases(512, seed=6):
        Rcw, tcw = T[:3, :3], T[:3, 3:4]
        Ts.append(T), xs.append(x)
        a.append(cv2.gemm(Rcw, x.reshape(3, 1), 1.0, tcw, 1.0).ravel())
        b.append(cv2.gemm(Rcw, tcw, -1.0, None, 0.0, flags=cv2.GEMM_1_T).ravel())
        c.append(cv2.gemm(np.ascontiguousarray(Rcw.T), tcw, -1.0, None, 0.0).ravel())
        n.append(cv2.norm(x.reshape(3, 1), cv2.NORM_L2))
    np.savez_compressed(os.path.join(HERE, "cv_gemm.npz"), cv2_version=cv2.__version__, T=np.stack(Ts), x=np.stack(xs),
                        rx_t=np.stack(a), neg_rt_t=np.stack(b), neg_rwc_t=np.stack(c), norm=np.asarray(n, np.float64))
    print("cv_gemm golden written")


if __name__ == "__main__":
    cv_gemm()
