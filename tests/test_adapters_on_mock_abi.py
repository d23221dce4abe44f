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
