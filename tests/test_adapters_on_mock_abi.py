"""The reference-side bindings of adapters/ EXECUTED on a CPU-only machine.

adapters/ORBextractor_msl.cc, PlaneExtractor_msl.cpp and SurfelFusion_msl.cpp are compiled against the reference's OWN
headers (include/ORBextractor.h, include/PlaneExtractor.h + the peac fitter, include/SurfelFusion.h) on the stand-in OpenCV /
Eigen of oracle/ref_shim_cv/, and linked with tests/host_emul/mock_abi.cpp -- the subset of include/msl_frontend.h they
call, implemented on the CPU oracle.  tests/host_emul/adapter_wrap.cpp then drives the class surfaces with the call
sequences of Frame::ExtractORB, Frame::ExtractPlanes and SurfelMapping::fuseMap, and the results are compared with the
REFERENCE'S OWN classes (oracle/_ref): what Tracking.cc / Frame.cc / SurfelMapping.cpp read from the bound classes is
what they read from the original ones.  This exercises the marshalling between the reference's types and the C ABI, not
the kernels.  Needs /root/reference; skipped elsewhere."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


@pytest.fixture(scope="module")
def adp(oracle):
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("/root/reference absent")
    for name in ("liborb_ref.so", "libplane_ref.so", "libsurfel_ref.so"):
        assert oracle.build_ref(name=name)
    he, out = os.path.join(HERE, "host_emul"), os.path.join(HERE, "host_emul", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libadapters_mock.so")
    flags = ["-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "oracle", "ref_shim_cv"),
             "-I" + os.path.join(ROOT, "oracle"), "-I" + os.path.join(REF, "include"), "-I" + REF, "-I" + os.path.join(ROOT, "include")]
    srcs = [os.path.join(he, "adapter_wrap.cpp"), os.path.join(he, "mock_abi.cpp"), os.path.join(ROOT, "adapters", "ORBextractor_msl.cc"),
            os.path.join(ROOT, "adapters", "PlaneExtractor_msl.cpp"), os.path.join(ROOT, "adapters", "SurfelFusion_msl.cpp"),
            os.path.join(ROOT, "oracle", "orb_oracle.cpp"), os.path.join(ROOT, "oracle", "plane_oracle.cpp"),
            os.path.join(ROOT, "oracle", "surfel_oracle.cpp")]
    deps = srcs + [os.path.join(ROOT, "oracle", "peac_oracle.inc"), os.path.join(ROOT, "include", "msl_frontend.h"),
                   os.path.join(ROOT, "oracle", "ref_shim_cv", "cvshim.hpp"), os.path.join(ROOT, "oracle", "ref_shim_cv", "eigenshim.hpp")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        # PlaneDetection's constructor / destructor / readColorImage come from the reference's own src/PlaneExtractor.cpp; its
        # readDepthImage / runPlaneDetection are renamed out of the way -- what the `#ifndef MSL_FRONTEND` of INTEGRATION.md does
        obj = os.path.join(out, "adp_plane_ref.o")
        subprocess.check_call(["g++"] + flags + ["-DreadDepthImage=readDepthImage_reference", "-DrunPlaneDetection=runPlaneDetection_reference",
                                                 "-c", "-o", obj, os.path.join(REF, "src", "PlaneExtractor.cpp")])
        subprocess.check_call(["g++"] + flags + ["-shared", "-o", so] + srcs + [obj])
    L = C.CDLL(so)
    L.adp_orb_extract.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p]
    L.adp_plane_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_void_p] * 6 + [C.c_int]
    L.adp_surfel_fuse.argtypes = ([C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int])
    return L


@pytest.mark.parametrize("seed,w,h,nf,nl", [(1, 640, 480, 1000, 8), (2, 752, 480, 1200, 8), (3, 320, 240, 500, 6)])
def test_orbextractor_binding_equals_reference_class(adp, oracle, seed, w, h, nf, nl):
    img = S.gray_frame(seed, w, h)
    cap = nf + 8 * nl + 64
    kps, desc, sf = np.zeros(cap, oracle.KP_DTYPE), np.zeros((cap, 32), np.uint8), np.zeros(nl, np.float32)
    n = adp.adp_orb_extract(nf, 1.2, nl, 20, 7, img.ctypes.data, w, h, w, kps.ctypes.data, desc.ctypes.data, cap, sf.ctypes.data)
    ref = oracle.RefOrbExtractor(nf, 1.2, nl, 20, 7)
    kr, dr = ref(img)
    assert n == len(kr) and kps[:n].tobytes() == kr.tobytes() and np.array_equal(desc[:n], dr)
    assert sf.tobytes() == ref.tables()[0].tobytes()  # GetScaleFactors()


def test_orbextractor_binding_empty_result(adp, oracle):
    img = np.full((480, 640), 50, np.uint8)
    kps, desc, sf = np.zeros(1200, oracle.KP_DTYPE), np.zeros((1200, 32), np.uint8), np.zeros(8, np.float32)
    assert adp.adp_orb_extract(1000, 1.2, 8, 20, 7, img.ctypes.data, 640, 480, 640, kps.ctypes.data, desc.ctypes.data, 1200,
                               sf.ctypes.data) == 0  # and the descriptor Mat was released (:831-832)


@pytest.mark.parametrize("seed,fac", [(2, 1.0), (3, 1.0), (4, 1.0), (5, 1.0 / 5000.0)])
def test_planedetection_binding_equals_reference_class(adp, oracle, seed, fac):
    d16, _ = S.depth_frame(seed)
    K = S.K_DEFAULT
    mem, cloud = np.zeros((240, 320), np.int32), np.zeros((240 * 320, 3))
    nrm, cen = np.zeros((64, 3)), np.zeros((64, 3))
    off, vert = np.zeros(65, np.int32), np.zeros(240 * 320, np.int32)
    n = adp.adp_plane_run(d16.ctypes.data, 640, 480, 640, K[0], K[1], K[2], K[3], fac, mem.ctypes.data, cloud.ctypes.data,
                          nrm.ctypes.data, cen.ctypes.data, off.ctypes.data, vert.ctypes.data, 64)
    mr, pr = oracle.ref_plane_run(d16, depth_map_factor=fac)
    cr = oracle.ref_plane_prestage(d16, depth_map_factor=fac)[0]
    assert n == len(pr["N"]) and np.array_equal(mem, mr)                      # plane_num_, plane_filter.membershipImg
    assert cloud.tobytes() == cr.reshape(-1, 3).tobytes()                     # cloud.vertices
    assert nrm[:n].tobytes() == pr["normal"].tobytes() and cen[:n].tobytes() == pr["center"].tobytes()  # extractedPlanes[i]
    for i in range(n):                                                        # plane_vertices_[i], in the reference's order
        assert np.array_equal(vert[off[i]:off[i + 1]], np.flatnonzero(mr.ravel() == i))
        assert off[i + 1] - off[i] == pr["vertices"][i]


@pytest.mark.parametrize("seed,pf,n", [(3, 0.0, 30000), (4, 0.4, 5000), (7, 0.0, 0)])
def test_surfelfusion_binding_equals_reference_class(adp, oracle, seed, pf, n):
    g = S.gray_frame(seed)
    _, d = S.depth_frame(seed)
    m = S.membership(seed, plane_fraction=pf)
    T = np.ascontiguousarray(S.pose_walk(seed, 1)[0], np.float32)
    local = S.surfel_map(seed, n, d, T, ref_index=20) if n else np.zeros(0, oracle.SURFEL_DTYPE)
    la, lr = local.copy(), local.copy()
    buf = np.zeros(480 * 640 + 3 * 640 + 16, np.uint8)  # the reference reads cv::Vec3b on the gray image
    buf[:480 * 640] = g.ravel()
    new = np.zeros(4800, oracle.SURFEL_DTYPE)
    k = adp.adp_surfel_fuse(640, 480, 525.0, 525.0, 319.5, 239.5, 30.0, 0.5, 20, buf.ctypes.data, 640, d.ctypes.data, m.ctypes.data,
                            T.ctypes.data, la.ctypes.data, len(la), new.ctypes.data, len(new))
    new_r = oracle.RefSurfelFusion().fuse(20, g, d, m, T, lr)
    assert k == len(new_r)
    for a, b in ((la, lr), (new[:k], new_r)):
        for f in a.dtype.names:
            x, y = a[f], b[f]
            if x.dtype.kind == "f":
                assert ((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))).all(), f
            else:
                assert np.array_equal(x, y), f


# ---- adapters/ORBmatcher_msl.cc: the binding's seven methods and the reference's own seven (src/ORBmatcher.cc, renamed out of
# the way by -D macros: what INTEGRATION.md's `#ifndef MSL_FRONTEND` does) in ONE library, called on the same stand-in
# Frame / KeyFrame / MapPoint object graphs (oracle/ref_shim_match/slam_standins.hpp) through the same wrapper
# (oracle/ref_match_wrap.cpp compiled twice); the C ABI underneath the binding = tests/host_emul/mock_abi_matcher.cpp.
@pytest.fixture(scope="module")
def madp(oracle):
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("/root/reference absent")
    orc, out = os.path.join(ROOT, "oracle"), os.path.join(HERE, "host_emul", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libmatch_adapter_mock.so")
    deps = [os.path.join(orc, f) for f in ("ref_match_wrap.cpp", "ref_shim_cv/cvshim.hpp", "ref_shim_match/slam_standins.hpp",
                                            "ref_shim_match/slam_standins.cpp", "match_oracle.cpp", "orb_oracle.cpp", "msl_oracle.h")]
    deps += [os.path.join(ROOT, "adapters", "ORBmatcher_msl.cc"), os.path.join(HERE, "host_emul", "mock_abi_matcher.cpp"),
             os.path.join(ROOT, "include", "msl_frontend.h")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        fl = ["-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + os.path.join(orc, "ref_shim_cv"),
              "-I" + os.path.join(orc, "ref_shim_match"), "-I" + orc, "-I" + os.path.join(REF, "include"), "-I" + REF,
              "-I" + os.path.join(ROOT, "include")]
        inc = ["-DMAPPOINT_H", "-DKEYFRAME_H", "-DFRAME_H", "-include", os.path.join(orc, "ref_shim_match", "slam_standins.hpp")]
        methods = ["SearchByProjection", "SearchByBoW", "SearchForTriangulation", "Fuse", "DescriptorDistance"]
        ren = ["-D%s=%s_reference" % (m, m) for m in methods]
        entries = ["descriptor_distance", "search_by_projection_frame", "search_by_projection_points", "search_by_projection_keyframe",
                   "search_by_bow", "search_for_triangulation", "fuse"]
        adp = ["-Dref_%s=adp_%s" % (e, e) for e in entries]
        objs = []

        def cc(name, src, extra):
            o = os.path.join(out, name)
            subprocess.check_call(["g++"] + fl + extra + ["-c", "-o", o, src])
            objs.append(o)
        cc("m_ref.o", os.path.join(REF, "src", "ORBmatcher.cc"), inc + ren)
        cc("m_wrap_ref.o", os.path.join(orc, "ref_match_wrap.cpp"), inc + ren)
        cc("m_wrap_adp.o", os.path.join(orc, "ref_match_wrap.cpp"), inc + adp)
        cc("m_adp.o", os.path.join(ROOT, "adapters", "ORBmatcher_msl.cc"), inc)
        cc("m_fv.o", os.path.join(REF, "Thirdparty", "DBoW2", "DBoW2", "FeatureVector.cpp"), [])
        subprocess.check_call(["g++"] + fl + inc + ["-shared", "-o", so] + objs +
                              [os.path.join(orc, "ref_shim_match", "slam_standins.cpp"), os.path.join(HERE, "host_emul", "mock_abi_matcher.cpp"),
                               os.path.join(orc, "match_oracle.cpp"), os.path.join(orc, "orb_oracle.cpp")])
        for o in objs:
            os.remove(o)
    return C.CDLL(so)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_orbmatcher_binding_equals_reference_methods(madp, oracle, seed):
    from manhattanslam_b200.matcher import frame_geom
    B, g = oracle, frame_geom()
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))

    def pair(fn, *a):
        with B.reference_matcher(madp, "ref_"):
            r = fn(*a)
        with B.reference_matcher(madp, "adp_"):
            d = fn(*a)
        assert r[0] == d[0] and np.array_equal(r[1], d[1]), fn.__name__
        return r[0]

    cur, last, mps, Tc, Tl = S.match_scene(seed)
    cur2, kf, Tc2 = S.reloc_scene(seed)
    kfb, f = S.bow_scene(seed, shuffle=bool(seed & 1))
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(seed)
    mpf, kfs, Tcw, ils = S.fuse_scene(seed)
    for chk in (False, True):
        assert pair(B.search_by_projection_frame, g, Tc, Tl, 15.0, chk, last, cur) > 100
        assert pair(B.search_by_projection_keyframe, g, Tc2, 15.0, 100, chk, lsf, kf, cur2) > 100
        assert pair(B.search_by_bow, 0.7, chk, kfb, f) > 100
        for only_stereo in (False, True):
            assert pair(B.search_for_triangulation, F12, Cw1, Tcw2, K2, only_stereo, chk, sf, ls, kf1, kf2) > 30
    for th in (1.0, 3.0):
        assert pair(B.search_by_projection_points, g, th, 0.8, mps, cur) > 100
    for th in (3.0, 5.0):
        r = B.ref_fuse(g, Tcw, th, lsf, ils, mpf, kfs, library=madp, name="ref_fuse")
        d = B.ref_fuse(g, Tcw, th, lsf, ils, mpf, kfs, library=madp, name="adp_fuse")
        assert r[0] == d[0] > 100 and np.array_equal(r[1], d[1])
    madp.adp_descriptor_distance.argtypes = madp.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    d = np.random.default_rng(seed).integers(0, 256, (2, 32), dtype=np.uint8)
    assert madp.adp_descriptor_distance(d[0].ctypes.data, d[1].ctypes.data) == madp.ref_descriptor_distance(d[0].ctypes.data, d[1].ctypes.data)


# ---- adapters/MapPoint_msl.cc: ComputeDistinctiveDescriptorsBatch on stand-in MapPoint / KeyFrame objects
@pytest.fixture(scope="module")
def mpadp(oracle):
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("/root/reference absent")
    orc, out = os.path.join(ROOT, "oracle"), os.path.join(HERE, "host_emul", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libmappoint_adapter_mock.so")
    srcs = [os.path.join(HERE, "host_emul", "mappoint_wrap.cpp"), os.path.join(ROOT, "adapters", "MapPoint_msl.cc"),
            os.path.join(HERE, "host_emul", "mock_abi_matcher.cpp"), os.path.join(orc, "ref_shim_match", "slam_standins.cpp"),
            os.path.join(REF, "Thirdparty", "DBoW2", "DBoW2", "FeatureVector.cpp"), os.path.join(orc, "match_oracle.cpp"),
            os.path.join(orc, "orb_oracle.cpp")]
    deps = srcs + [os.path.join(orc, "ref_shim_match", "slam_standins.hpp"), os.path.join(orc, "ref_shim_cv", "cvshim.hpp")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-w", "-I" + os.path.join(orc, "ref_shim_cv"),
                               "-I" + os.path.join(orc, "ref_shim_match"), "-I" + orc, "-I" + os.path.join(REF, "include"), "-I" + REF,
                               "-I" + os.path.join(ROOT, "include"), "-DMAPPOINT_H", "-DKEYFRAME_H", "-DFRAME_H", "-include",
                               os.path.join(orc, "ref_shim_match", "slam_standins.hpp"), "-shared", "-o", so] + srcs)
    L = C.CDLL(so)
    L.adp_distinctive.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    return L


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mappoint_batch_binding_picks_the_reference_descriptor(mpadp, oracle, seed):
    """per map point: the observed rows of the non-bad keyframes in mObservations order (keyframe address = index here), the
    row with the least median distance (src/MapPoint.cc:210-263) copied into mDescriptor; bad points, points without
    observations and points whose keyframes are all bad keep their descriptor"""
    r = np.random.default_rng(seed)
    n_kf, n_mp = 40, 300
    kf_rows = r.integers(5, 60, n_kf).astype(np.int32)
    kf_desc = r.integers(0, 256, (int(kf_rows.sum()), 32), dtype=np.uint8)
    kf_bad = (r.random(n_kf) < 0.15).astype(np.uint8)
    row0 = np.concatenate([[0], np.cumsum(kf_rows)])
    mp_bad = (r.random(n_mp) < 0.1).astype(np.uint8)
    obs_off, obs_kf, obs_row, expect = [0], [], [], []
    for p in range(n_mp):
        k = int(r.choice([0, 1, 2, 3, 8, 20, 33]))
        ks = np.sort(r.choice(n_kf, min(k, n_kf), replace=False))
        rows = [int(r.integers(0, kf_rows[q])) for q in ks]
        obs_kf += ks.tolist()
        obs_row += rows
        obs_off.append(len(obs_kf))
        good = [kf_desc[row0[q] + rw] for q, rw in zip(ks, rows) if not kf_bad[q]]
        if mp_bad[p] or not good:
            expect.append(np.zeros(32, np.uint8))
        else:
            bi, _ = oracle.distinctive_descriptors([np.stack(good)])
            expect.append(good[int(bi[0])])
    out = np.zeros((n_mp, 32), np.uint8)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    oo, ok, orw = a(obs_off, np.int32), a(obs_kf, np.int32), a(obs_row, np.int32)
    assert mpadp.adp_distinctive(n_kf, kf_rows.ctypes.data, kf_desc.ctypes.data, kf_bad.ctypes.data, n_mp, oo.ctypes.data,
                                 ok.ctypes.data, orw.ctypes.data, mp_bad.ctypes.data, out.ctypes.data) == 0
    assert np.array_equal(out, np.stack(expect))
    assert (out.any(axis=1)).sum() > n_mp // 2


# ---- adapters/FrameGlue_msl.cc: Frame::UndistortKeyPoints + Frame::ComputeStereoFromRGBD on the stand-in Frame
@pytest.fixture(scope="module")
def gadp(oracle):
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("/root/reference absent")
    orc, out = os.path.join(ROOT, "oracle"), os.path.join(HERE, "host_emul", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libglue_adapter_mock.so")
    srcs = [os.path.join(HERE, "host_emul", "glue_wrap.cpp"), os.path.join(ROOT, "adapters", "FrameGlue_msl.cc"),
            os.path.join(orc, "ref_shim_match", "slam_standins.cpp"), os.path.join(REF, "Thirdparty", "DBoW2", "DBoW2", "FeatureVector.cpp"),
            os.path.join(orc, "glue_oracle.cpp"), os.path.join(orc, "match_oracle.cpp"), os.path.join(orc, "orb_oracle.cpp")]
    deps = srcs + [os.path.join(orc, "ref_shim_match", "slam_standins.hpp"), os.path.join(orc, "ref_shim_cv", "cvshim.hpp")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-w", "-I" + os.path.join(orc, "ref_shim_cv"),
                               "-I" + os.path.join(orc, "ref_shim_match"), "-I" + orc, "-I" + os.path.join(REF, "include"), "-I" + REF,
                               "-I" + os.path.join(ROOT, "include"), "-DMAPPOINT_H", "-DKEYFRAME_H", "-DFRAME_H", "-include",
                               os.path.join(orc, "ref_shim_match", "slam_standins.hpp"), "-shared", "-o", so] + srcs)
    L = C.CDLL(so)
    L.adp_frame_glue.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("distorted", [True, False])
def test_frame_glue_binding(gadp, oracle, distorted):
    """the RGB-D Frame constructor's two calls (src/Frame.cc:105-112): mvKeysUn (all KeyPoint fields kept, pt undistorted by
    the cv2-pinned cv::undistortPoints restatement -- or copied when k1 == 0), then mvDepth / mvuRight from the depth at the
    DISTORTED pixel and the UNDISTORTED x (src/Frame.cc:495-513)"""
    r = np.random.default_rng(5)
    n = 800
    kps = np.zeros(n, oracle.KP_DTYPE)
    kps["x"], kps["y"] = r.uniform(0, 639, n), r.uniform(0, 479, n)
    kps["size"], kps["angle"], kps["response"] = 31, r.uniform(0, 360, n), r.integers(7, 200, n)
    kps["octave"], kps["class_id"] = r.integers(0, 8, n), -1
    K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)  # Example/TUM1.yaml
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314] if distorted else [0, 0, 0, 0, 0], np.float32)
    _, depth = S.depth_frame(3)
    depth = np.ascontiguousarray(depth, np.float32)
    mbf = 40.0
    un, ur, kd = np.zeros(n, oracle.KP_DTYPE), np.zeros(n, np.float32), np.zeros(n, np.float32)
    assert gadp.adp_frame_glue(n, kps.ctypes.data, K4.ctypes.data, D.ctypes.data, 5, depth.ctypes.data, 640, 480, mbf, un.ctypes.data,
                               ur.ctypes.data, kd.ctypes.data) == 0
    xy = np.stack([kps["x"], kps["y"]], 1)
    xy_un = oracle.undistort_keypoints(xy, K4, D)
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(un[f], kps[f]), f
    assert np.array_equal(np.stack([un["x"], un["y"]], 1), xy_un)
    assert (np.abs(xy_un - xy).max() > 0.5) == distorted
    ur_o, kd_o = oracle.stereo_from_rgbd(xy, xy_un, depth, mbf)
    assert np.array_equal(kd, kd_o) and np.array_equal(ur, ur_o)
    assert (kd > 0).sum() > n // 2 and (kd == -1).sum() > 0


# ---- adapters/SurfelMapping_msl.cpp (device-resident mode) inside the reference's own SurfelMapping class
@pytest.fixture(scope="module")
def smadp(oracle):
    if not os.path.isdir(os.path.join(REF, "include")):
        pytest.skip("/root/reference absent")
    assert oracle.build_ref(name="libmapping_ref.so")
    orc, out = os.path.join(ROOT, "oracle"), os.path.join(HERE, "host_emul", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libmapping_adapter_mock.so")
    srcs = [os.path.join(HERE, "host_emul", "mapping_adapter_wrap.cpp"), os.path.join(ROOT, "adapters", "SurfelMapping_msl.cpp"),
            os.path.join(ROOT, "adapters", "SurfelFusion_msl.cpp"), os.path.join(HERE, "host_emul", "mock_abi.cpp"),
            os.path.join(orc, "surfel_oracle.cpp"), os.path.join(orc, "orb_oracle.cpp"), os.path.join(orc, "plane_oracle.cpp")]
    deps = srcs + [os.path.join(orc, "ref_shim_map", "map_standins.hpp"), os.path.join(orc, "ref_shim_cv", "cvshim.hpp"),
                   os.path.join(orc, "ref_shim_cv", "eigenshim.hpp"), os.path.join(orc, "peac_oracle.inc")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        fl = ["-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-w", "-DMSL_SURFEL_RESIDENT", "-I" + os.path.join(orc, "ref_shim_cv"),
              "-I" + os.path.join(orc, "ref_shim_map"), "-I" + orc, "-I" + os.path.join(REF, "include"), "-I" + os.path.join(ROOT, "include"),
              "-DSYSTEM_H", "-DMAP_H", "-include", os.path.join(orc, "ref_shim_map", "map_standins.hpp")]
        obj = os.path.join(out, "sm_ref.o")
        # the reference's two definitions renamed out of the way: what `#ifndef MSL_SURFEL_RESIDENT` does in a checkout
        subprocess.check_call(["g++"] + fl + ["-DmoveAddSurfels=moveAddSurfels_reference", "-DfuseMap=fuseMap_reference", "-c", "-o", obj,
                                              os.path.join(REF, "src", "SurfelMapping.cpp")])
        subprocess.check_call(["g++"] + fl + ["-shared", "-o", so, obj] + srcs)
        os.remove(obj)
    L = C.CDLL(so)
    L.adp_mapping_create.restype = C.c_void_p
    L.adp_mapping_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6
    L.adp_mapping_keyframe.restype = C.c_int64
    L.adp_mapping_keyframe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    for f in (L.adp_mapping_local, L.adp_mapping_inactive):
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    return L


def test_surfelmapping_binding_equals_reference_class(smadp, oracle):
    """moveAddSurfels + fuseMap of the binding (surfel maps behind the C ABI, host vectors as mirrors) against the reference's
    own SurfelMapping over a 34-keyframe pose graph: Map::mvLocalSurfels and Map::mvInactiveSurfels after every keyframe"""
    w, h = 320, 240
    K = tuple(k * 0.5 for k in S.K_DEFAULT)
    r = oracle.RefSurfelMapping(w, h, *K)
    a = smadp.adp_mapping_create(w, h, *K, 30.0, 0.5)

    def get(fn):
        n = fn(a, None, 0)
        out = np.zeros(n, oracle.SURFEL_DTYPE)
        fn(a, out.ctypes.data, n)
        return out

    refs = [0] + [i - 1 for i in range(1, 26)] + [3, 26, 4, 28, 27, 2, 30, 31]
    poses, g, mem = S.pose_walk(5, len(refs)), S.gray_frame(5, w, h), np.ascontiguousarray(S.membership(5, w, h, plane_fraction=0.2), np.int32)
    moved = 0
    for i, ri in enumerate(refs):
        _, d = S.depth_frame(300 + i % 4, w, h, K=K, scene=300)
        add, rem = r.keyframe(g, d, mem, poses[i], ri)
        moved += len(add) + len(rem)
        buf = np.zeros(h * w + 3 * w + 16, np.uint8)
        buf[:h * w] = g.ravel()
        dd, T = np.ascontiguousarray(d, np.float32), np.ascontiguousarray(poses[i], np.float32)
        smadp.adp_mapping_keyframe(a, buf.ctypes.data, w, h, dd.ctypes.data, mem.ctypes.data, T.ctypes.data, ri)
        for x, y in ((get(smadp.adp_mapping_local), r.local()), (get(smadp.adp_mapping_inactive), r.inactive())):
            assert x.shape == y.shape, i
            for f in x.dtype.names:
                if x[f].dtype.kind == "f":
                    assert ((x[f].view(np.uint32) == y[f].view(np.uint32)) | (np.isnan(x[f]) & np.isnan(y[f]))).all(), (i, f)
                else:
                    assert np.array_equal(x[f], y[f]), (i, f)
    assert moved > 40 and len(r.inactive()) > 1000
