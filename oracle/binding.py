"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/msl_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (manhattanslam_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/libmsl_oracle.so with the committed Makefile (gcc only, no GPU needed)."""
    import hashlib
    so = os.path.join(_HERE, "libmsl_oracle.so")
    stamp = so + ".stamp"
    h = hashlib.sha256()
    for f in sorted(os.listdir(_HERE)):
        if f.endswith((".cpp", ".h", ".inc")) or f == "Makefile":
            h.update(f.encode())
            with open(os.path.join(_HERE, f), "rb") as fh:
                h.update(fh.read())
    digest = h.hexdigest()
    # content hash, not mtimes: mtimes do not survive the gpurun snapshot
    if force or not os.path.exists(so) or not os.path.exists(stamp) or open(stamp).read().strip() != digest:
        subprocess.check_call(["make", "-B", "-C", _HERE, "libmsl_oracle.so"], stdout=subprocess.DEVNULL)
        with open(stamp, "w") as fh:
            fh.write(digest)
    return so


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

SURFEL_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                         ("size", "<f4"), ("color", "<f4"), ("r", "<i4"), ("g", "<i4"), ("b", "<i4"),
                         ("weight", "<f4"), ("updateTimes", "<i4"), ("lastUpdate", "<i4")])

SEED_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"),
                       ("normX", "<f4"), ("normY", "<f4"), ("normZ", "<f4"),
                       ("posX", "<f4"), ("posY", "<f4"), ("posZ", "<f4"),
                       ("viewCos", "<f4"), ("meanDepth", "<f4"), ("meanIntensity", "<f4"),
                       ("r", "<i4"), ("g", "<i4"), ("b", "<i4"),
                       ("fused", "<i4"), ("stable", "<i4"), ("use", "<i4")])

BLOCK_DTYPE = np.dtype([("center", "<f8", 3), ("normal", "<f8", 3), ("mse", "<f8"), ("curvature", "<f8"),
                        ("N", "<i4"), ("nouse", "<i4")])

GEOM_DTYPE = np.dtype([("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4"),
                       ("mnMinX", "<f4"), ("mnMinY", "<f4"), ("mnMaxX", "<f4"), ("mnMaxY", "<f4"),
                       ("gridWInv", "<f4"), ("gridHInv", "<f4"), ("mb", "<f4"), ("mbf", "<f4"),
                       ("nlevels", "<i4"), ("scaleFactors", "<f4", 16)])


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


_LIB_OVERRIDE = None


def lib():
    global _LIB
    if _LIB_OVERRIDE is not None:
        return _LIB_OVERRIDE
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_orb_create.restype = C.c_void_p
        L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_orb_destroy.argtypes = [C.c_void_p]
        L.orc_orb_levels.argtypes = [C.c_void_p]
        L.orc_orb_scale_factors.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_orb_features_per_level.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_orb_umax.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int]
        L.orc_orb_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_orb_level_image.restype = C.c_void_p
        L.orc_orb_level_image.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_level_blurred.restype = C.c_void_p
        L.orc_orb_level_blurred.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_orb_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int]
        L.orc_gaussian_blur_7x7_s2_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_9_16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_score_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_cv_round_f.argtypes = [C.c_float]
        for name, setup in _LATE.items():
            if hasattr(L, name):
                setup(getattr(L, name))
        _LIB = L
    return _LIB


_LATE = {}


def _late(name):
    def deco(fn):
        _LATE[name] = fn
        return fn
    return deco


# ------------------------------------------------------------------------ ORB

class OrbOracle:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:42-104) on the CPU oracle."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.L = lib()
        self.h = self.L.orc_orb_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        if not self.h:
            raise ValueError("bad ORB parameters")
        self.nlevels = nlevels
        self.nfeatures = nfeatures

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_orb_destroy(self.h)
            self.h = None

    def scale_factors(self):
        out = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.L.orc_orb_scale_factors(self.h, *[_p(o) for o in out])
        return out

    def features_per_level(self):
        n = np.zeros(self.nlevels, np.int32)
        self.L.orc_orb_features_per_level(self.h, _p(n))
        return n

    def umax(self):
        u = np.zeros(16, np.int32)
        self.L.orc_orb_umax(self.h, _p(u))
        return u

    def __call__(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_orb_extract(self.h, _p(gray), w, h, gray.strides[0], _p(kps), _p(desc), cap)
        if n == -2:
            raise ValueError("the reference is undefined for this geometry (a pyramid level more than twice as tall as wide)")
        if n < 0:
            raise RuntimeError("oracle keypoint capacity exceeded")
        return kps[:n].copy(), desc[:n].copy()

    def level_image(self, level, blurred=False):
        w, h = C.c_int(), C.c_int()
        self.L.orc_orb_level_size(self.h, level, C.byref(w), C.byref(h))
        ptr = (self.L.orc_orb_level_blurred if blurred else self.L.orc_orb_level_image)(self.h, level)
        if not ptr:
            return None
        buf = (C.c_uint8 * (w.value * h.value)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(h.value, w.value).copy()

    def level_candidates(self, level):
        n = self.L.orc_orb_level_candidates(self.h, level, None, 0)
        out = np.zeros((n, 3), np.int32)
        self.L.orc_orb_level_candidates(self.h, level, _p(out), n)
        return out

    def level_keypoints(self, level):
        n = self.L.orc_orb_level_keypoints(self.h, level, None, 0)
        out = np.zeros(n, KP_DTYPE)
        self.L.orc_orb_level_keypoints(self.h, level, _p(out), n)
        return out


def resize_linear_u8(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gaussian_blur_7x7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orc_gaussian_blur_7x7_s2_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_9_16(img, threshold, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.zeros((cap, 3), np.int32)
    n = lib().orc_fast_9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), cap)
    return out[:n].copy()


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), out.strides[0])
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


# -------------------------------------------------------------------- surfels

@_late("orc_surfel_create")
def _s1(f):
    f.restype = C.c_void_p
    f.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6


@_late("orc_surfel_fuse")
def _s2(f):
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_int64, C.c_void_p, C.c_int, C.c_int]


@_late("orc_surfel_compact")
def _s3(f):
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]


class SurfelOracle:
    """Mirror of SurfelFusion (include/SurfelFusion.h:44-139) on the CPU oracle."""

    def __init__(self, w=640, h=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, fuseFar=30.0, fuseNear=0.5):
        self.L = lib()
        self.w, self.h = w, h
        self.nseeds = (w // 8) * (h // 8)
        self.hd = self.L.orc_surfel_create(w, h, fx, fy, cx, cy, fuseFar, fuseNear)
        for name in ("orc_surfel_index", "orc_surfel_seeds", "orc_surfel_normmap"):
            fn = getattr(self.L, name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p]
        for name in ("orc_surfel_seeds_iter", "orc_surfel_index_iter"):
            fn = getattr(self.L, name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p, C.c_int]
        self.L.orc_surfel_destroy.argtypes = [C.c_void_p]

    def __del__(self):
        if getattr(self, "hd", None):
            self.L.orc_surfel_destroy(self.hd)
            self.hd = None

    def fuse(self, ref, gray, depth, membership, Twc, local, threads=1):
        """fuseInitializeMap: local (SURFEL_DTYPE array) is updated in place; returns newSurfels."""
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        membership = np.ascontiguousarray(membership, np.int32)
        Twc = np.ascontiguousarray(Twc, np.float32)
        assert local.dtype == SURFEL_DTYPE and local.flags.c_contiguous
        new = np.zeros(self.nseeds, SURFEL_DTYPE)
        n = self.L.orc_surfel_fuse(self.hd, ref, _p(gray), gray.strides[0], _p(depth), _p(membership), _p(Twc),
                                   _p(local), len(local), _p(new), len(new), threads)
        return new[:n].copy()

    def _arr(self, ptr, dtype, count):
        buf = (C.c_uint8 * (np.dtype(dtype).itemsize * count)).from_address(ptr)
        return np.frombuffer(buf, dtype).copy()

    def seeds(self, it=None):
        ptr = self.L.orc_surfel_seeds(self.hd) if it is None else self.L.orc_surfel_seeds_iter(self.hd, it)
        return self._arr(ptr, SEED_DTYPE, self.nseeds)

    def index(self, it=None):
        ptr = self.L.orc_surfel_index(self.hd) if it is None else self.L.orc_surfel_index_iter(self.hd, it)
        return self._arr(ptr, np.int32, self.w * self.h).reshape(self.h, self.w)

    def normmap(self):
        return self._arr(self.L.orc_surfel_normmap(self.hd), np.float32, self.w * self.h * 3).reshape(self.h, self.w, 3)


def surfel_compact(local, new):
    """SurfelMapping::fuseMap tail (src/SurfelMapping.cpp:366-391); returns the new local array."""
    buf = np.zeros(len(local) + len(new), SURFEL_DTYPE)
    buf[:len(local)] = local
    new = np.ascontiguousarray(new)
    n = lib().orc_surfel_compact(_p(buf), len(local), _p(new), len(new))
    return buf[:n].copy()


# --------------------------------------------------------------- plane pre-stage

@_late("orc_plane_prestage")
def _p1(f):
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_void_p] * 4


def plane_prestage(depth_u16, K=(525.0, 525.0, 319.5, 239.5), depth_map_factor=1.0 / 5000.0):
    """-> (cloud[h2,w2,3] f64, blocks[BLOCK_DTYPE], seed[u8], edges[u8])"""
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth_u16.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    nb = (h2 // 10) * (w2 // 10)
    cloud = np.zeros((h2, w2, 3), np.float64)
    blocks = np.zeros(nb, BLOCK_DTYPE)
    seed = np.zeros(nb, np.uint8)
    edges = np.zeros(nb, np.uint8)
    lib().orc_plane_prestage(_p(depth_u16), w, h, depth_u16.strides[0] // 2, K[0], K[1], K[2], K[3], depth_map_factor,
                             _p(cloud), _p(blocks), _p(seed), _p(edges))
    return cloud, blocks, seed, edges


def plane_detect(depth_u16, K=(525.0, 525.0, 319.5, 239.5), depth_map_factor=1.0 / 5000.0, cap=64):
    """The whole of ahc::PlaneFitter::run (pre-stage + ahCluster + refineDetails) ->
    (membershipImg[h2,w2] i32, dict(normal, center, N, rid, vertices) of the extracted planes)"""
    L = lib()
    L.orc_plane_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_void_p] * 6 + [C.c_int]
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth_u16.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    mem = np.zeros((h2, w2), np.int32)
    nrm, cen = np.zeros((cap, 3)), np.zeros((cap, 3))
    N, rid, nv = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = L.orc_plane_detect(_p(depth_u16), w, h, depth_u16.strides[0] // 2, K[0], K[1], K[2], K[3], depth_map_factor,
                           _p(mem), _p(nrm), _p(cen), _p(N), _p(rid), _p(nv), cap)
    n = min(n, cap)
    return mem, dict(normal=nrm[:n], center=cen[:n], N=N[:n], rid=rid[:n], vertices=nv[:n])


def eig33sym(K):
    lib().orc_eig33sym.argtypes = [C.c_void_p] * 3
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    s = np.zeros(3)
    V = np.zeros(9)
    lib().orc_eig33sym(_p(K), _p(s), _p(V))
    return s, V.reshape(3, 3)


# ------------------------------------------------------------------- matcher

def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    lib().orc_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    return lib().orc_descriptor_distance(_p(a), _p(b))


def cv_rx_plus_t(R, x, t):
    """cv::Mat R * x + t for 3x3 / 3x1 CV_32F as the oracle evaluates it"""
    L = lib()
    out = np.zeros(3, np.float32)
    L.orc_cv_rx_plus_t(_p(np.ascontiguousarray(R, np.float32)), _p(np.ascontiguousarray(x, np.float32)),
                       _p(np.ascontiguousarray(t, np.float32)), _p(out))
    return out


def cv_neg_rt_times_t(R, t):
    """-R.t() * t (transposed operand)"""
    L = lib()
    out = np.zeros(3, np.float32)
    L.orc_cv_neg_rt_times_t(_p(np.ascontiguousarray(R, np.float32)), _p(np.ascontiguousarray(t, np.float32)), _p(out))
    return out


def cv_neg_rwc_times_t(Rcw, t):
    """Rwc = Rcw.t(); -Rwc * t (KeyFrame::SetPose)"""
    L = lib()
    out = np.zeros(3, np.float32)
    L.orc_cv_neg_rwc_times_t(_p(np.ascontiguousarray(Rcw, np.float32)), _p(np.ascontiguousarray(t, np.float32)), _p(out))
    return out


def cv_norm3(v):
    L = lib()
    L.orc_cv_norm3.restype = C.c_double
    return float(L.orc_cv_norm3(_p(np.ascontiguousarray(v, np.float32))))


def hamming_best2(q, t):
    L = lib()
    L.orc_hamming_best2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    bi, bd, sd = (np.zeros(len(q), np.int32) for _ in range(3))
    L.orc_hamming_best2(_p(q), len(q), _p(t), len(t), _p(bi), _p(bd), _p(sd))
    return bi, bd, sd


def features_in_area(geom, kp_xy, kp_octave, x, y, r, minLevel=-1, maxLevel=-1):
    L = lib()
    L.orc_features_in_area.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_int, C.c_int, C.c_void_p, C.c_int]
    kp_xy = np.ascontiguousarray(kp_xy, np.float32)
    kp_octave = np.ascontiguousarray(kp_octave, np.int32)
    out = np.zeros(len(kp_octave), np.int32)
    n = L.orc_features_in_area(_p(geom), _p(kp_xy), _p(kp_octave), len(kp_octave), x, y, r, minLevel, maxLevel, _p(out),
                               len(out))
    return out[:n].copy()


def search_by_projection_frame(geom, Tcw_cur, Tcw_last, th, check_ori, last, cur):
    L = lib()
    L.orc_search_by_projection_frame.argtypes = ([C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 7 +
                                                 [C.c_int] + [C.c_void_p] * 7)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    largs = [a(last["has_mp"], np.uint8), a(last["outlier"], np.uint8), a(last["mp_obs"], np.uint8),
             a(last["mp_world"], np.float32), a(last["mp_desc"], np.uint8), a(last["octave"], np.int32),
             a(last["angle"], np.float32)]
    cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["angle"], np.float32),
             a(cur["uright"], np.float32), a(cur["desc"], np.uint8), a(cur["occupied"], np.uint8)]
    cm = np.zeros(len(cur["octave"]), np.int32)
    n = L.orc_search_by_projection_frame(_p(geom), _p(a(Tcw_cur, np.float32)), _p(a(Tcw_last, np.float32)), th,
                                         int(check_ori), len(last["octave"]), *[_p(x) for x in largs],
                                         len(cur["octave"]), *[_p(x) for x in cargs], _p(cm))
    return n, cm


def search_by_projection_points(geom, th, nnratio, mps, cur):
    L = lib()
    L.orc_search_by_projection_points.argtypes = ([C.c_void_p, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 6 +
                                                  [C.c_int] + [C.c_void_p] * 6)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    margs = [a(mps["valid"], np.uint8), a(mps["obs"], np.uint8), a(mps["proj_xyr"], np.float32),
             a(mps["level"], np.int32), a(mps["viewcos"], np.float32), a(mps["desc"], np.uint8)]
    cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["uright"], np.float32),
             a(cur["desc"], np.uint8), a(cur["occupied"], np.uint8)]
    cm = np.zeros(len(cur["octave"]), np.int32)
    n = L.orc_search_by_projection_points(_p(geom), th, nnratio, len(mps["level"]), *[_p(x) for x in margs],
                                          len(cur["octave"]), *[_p(x) for x in cargs], _p(cm))
    return n, cm


def search_by_projection_keyframe(geom, Tcw_cur, th, orb_dist, check_ori, log_scale_factor, kf, cur):
    L = lib()
    L.orc_search_by_projection_keyframe.argtypes = ([C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int] +
                                                    [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    kargs = [a(kf["valid"], np.uint8), a(kf["mp_world"], np.float32), a(kf["mp_desc"], np.uint8),
             a(kf["mp_dist"], np.float32), a(kf["angle"], np.float32)]
    cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["angle"], np.float32), a(cur["desc"], np.uint8),
             a(cur["occupied"], np.uint8)]
    cm = np.zeros(len(cur["octave"]), np.int32)
    n = L.orc_search_by_projection_keyframe(_p(geom), _p(a(Tcw_cur, np.float32)), th, int(orb_dist), int(check_ori),
                                            log_scale_factor, len(kf["angle"]), *[_p(x) for x in kargs],
                                            len(cur["octave"]), *[_p(x) for x in cargs], _p(cm))
    return n, cm


def _csr(fv):
    """FeatureVector dict/list of (node id, [feature indices]) -> (ids u32, offsets i32, features i32), ascending ids."""
    if isinstance(fv, tuple) and len(fv) == 3 and isinstance(fv[0], np.ndarray):
        return fv  # already packed (bench.py packs once so that neither arm is timed on Python list handling)
    items = sorted(fv.items()) if isinstance(fv, dict) else sorted(fv)
    ids = np.asarray([k for k, _ in items], np.uint32)
    off = np.zeros(len(items) + 1, np.int32)
    for i, (_, v) in enumerate(items):
        off[i + 1] = off[i] + len(v)
    feat = np.asarray([x for _, v in items for x in v], np.int32)
    return ids, off, feat


def search_by_bow(nnratio, check_ori, kf, f):
    """kf: dict(featvec, valid, desc, angle); f: dict(featvec, desc, angle) -> (nmatches, f_match)"""
    L = lib()
    L.orc_search_by_bow.argtypes = ([C.c_float, C.c_int] + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 +
                                    [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    kid, koff, kfeat = _csr(kf["featvec"])
    fid, foff, ffeat = _csr(f["featvec"])
    kv, kd, ka = a(kf["valid"], np.uint8), a(kf["desc"], np.uint8), a(kf["angle"], np.float32)
    fd, fa = a(f["desc"], np.uint8), a(f["angle"], np.float32)
    fm = np.zeros(len(fa), np.int32)
    n = L.orc_search_by_bow(nnratio, int(check_ori), len(kid), _p(kid), _p(koff), _p(kfeat), len(fid), _p(fid), _p(foff),
                            _p(ffeat), len(ka), _p(kv), _p(kd), _p(ka), len(fa), _p(fd), _p(fa), _p(fm))
    return n, fm


def search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, check_ori, scale_factors2, level_sigma2_2, kf1, kf2):
    """kf1: dict(featvec, has_mp, uright, xy, angle, desc); kf2: the same + octave -> (nmatches, matches12)"""
    L = lib()
    L.orc_search_for_triangulation.argtypes = ([C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 3 +
                                               [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int] +
                                               [C.c_void_p] * 6 + [C.c_void_p])
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    id1, off1, ft1 = _csr(kf1["featvec"])
    id2, off2, ft2 = _csr(kf2["featvec"])
    sf, ls = a(scale_factors2, np.float32), a(level_sigma2_2, np.float32)
    a1 = [a(kf1["has_mp"], np.uint8), a(kf1["uright"], np.float32), a(kf1["xy"], np.float32), a(kf1["angle"], np.float32),
          a(kf1["desc"], np.uint8)]
    a2 = [a(kf2["has_mp"], np.uint8), a(kf2["uright"], np.float32), a(kf2["xy"], np.float32), a(kf2["octave"], np.int32),
          a(kf2["angle"], np.float32), a(kf2["desc"], np.uint8)]
    m12 = np.zeros(len(a1[0]), np.int32)
    n = L.orc_search_for_triangulation(_p(a(F12, np.float32)), _p(a(Cw1, np.float32)), _p(a(Tcw2, np.float32)),
                                       _p(a(K2, np.float32)), int(only_stereo), int(check_ori), len(sf), _p(sf), _p(ls),
                                       len(id1), _p(id1), _p(off1), _p(ft1), len(id2), _p(id2), _p(off2), _p(ft2),
                                       len(a1[0]), *[_p(x) for x in a1], len(a2[0]), *[_p(x) for x in a2], _p(m12))
    return n, m12


def fuse_search(geom, Tcw, th, log_scale_factor, inv_level_sigma2, mps, kf):
    """mps: dict(valid, world, normal, dist, desc); kf: dict(xy, octave, uright, desc) -> (nfused, best_idx, best_dist)"""
    L = lib()
    L.orc_fuse_search.argtypes = ([C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int] + [C.c_void_p] * 5 +
                                  [C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 2)
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    ils = a(inv_level_sigma2, np.float32)
    margs = [a(mps["valid"], np.uint8), a(mps["world"], np.float32), a(mps["normal"], np.float32), a(mps["dist"], np.float32),
             a(mps["desc"], np.uint8)]
    kargs = [a(kf["xy"], np.float32), a(kf["octave"], np.int32), a(kf["uright"], np.float32), a(kf["desc"], np.uint8)]
    bi = np.zeros(len(margs[0]), np.int32)
    bd = np.zeros(len(margs[0]), np.int32)
    n = L.orc_fuse_search(_p(geom), _p(a(Tcw, np.float32)), th, log_scale_factor, _p(ils), len(margs[0]),
                          *[_p(x) for x in margs], len(kargs[1]), *[_p(x) for x in kargs], _p(bi), _p(bd))
    return n, bi, bd


def distinctive_descriptors(desc_lists):
    """desc_lists: list of (N_k, 32) uint8 arrays -> (best_idx, best_median) per map point"""
    L = lib()
    off = np.zeros(len(desc_lists) + 1, np.int32)
    for k, d in enumerate(desc_lists):
        off[k + 1] = off[k] + len(d)
    desc = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_lists] +
                                               [np.zeros((0, 32), np.uint8)]))
    bi = np.zeros(len(desc_lists), np.int32)
    bm = np.zeros(len(desc_lists), np.int32)
    L.orc_distinctive_descriptors(len(desc_lists), _p(off), _p(desc), _p(bi), _p(bm))
    return bi, bm


REFERENCE_ROOT = "/root/reference"


_REF_DEPS = {
    "libsurfel_ref.so": (["src/SurfelFusion.cpp", "include/SurfelFusion.h"],
                         ["ref_wrap.cpp", "ref_shim_cv/eigenshim.hpp", "ref_shim_cv/cvshim.hpp", "ref_shim_cv/seq_thread/thread"]),
    "libsurfel_ref_threads.so": (["src/SurfelFusion.cpp", "include/SurfelFusion.h"],
                                 ["ref_wrap.cpp", "ref_shim_cv/eigenshim.hpp", "ref_shim_cv/cvshim.hpp"]),
    "libmapping_ref.so": (["src/SurfelMapping.cpp", "src/SurfelFusion.cpp", "include/SurfelMapping.h", "include/SurfelFusion.h"],
                          ["ref_mapping_wrap.cpp", "ref_shim_map/map_standins.hpp", "ref_shim_map/pcl/point_types.h",
                           "ref_shim_cv/cvshim.hpp", "ref_shim_cv/eigenshim.hpp", "ref_shim_cv/seq_thread/thread"]),
    "libmappoint_ref.so": (["src/MapPoint.cc", "include/MapPoint.h"],
                           ["ref_mappoint_wrap.cpp", "ref_shim_mp/mp_standins.hpp", "ref_shim_cv/cvshim.hpp", "match_oracle.cpp",
                            "orb_oracle.cpp", "msl_oracle.h"]),
    "libframe_ref.so": (["src/Frame.cc", "src/KeyFrame.cc"],
                        ["ref_frame_wrap.cpp", "ref_shim_cv/cvshim.hpp", "glue_oracle.cpp", "match_oracle.cpp", "orb_oracle.cpp",
                         "msl_oracle.h"]),
    "liborb_ref.so": (["src/ORBextractor.cc", "include/ORBextractor.h"],
                      ["ref_orb_wrap.cpp", "ref_arena.hpp", "ref_shim_cv/cvshim.hpp", "orb_oracle.cpp", "msl_oracle.h"]),
    "libplane_ref.so": (["src/PlaneExtractor.cpp", "include/PlaneExtractor.h", "include/peac/AHCPlaneFitter.hpp",
                         "include/peac/AHCPlaneSeg.hpp", "include/peac/AHCParamSet.hpp", "include/peac/eig33sym.hpp"],
                        ["ref_plane_wrap.cpp", "ref_arena.hpp", "ref_shim_cv/cvshim.hpp", "ref_shim_cv/eigenshim.hpp",
                         "plane_oracle.cpp", "orb_oracle.cpp", "msl_oracle.h"]),
    "libmatch_ref.so": (["src/ORBmatcher.cc", "include/ORBmatcher.h", "Thirdparty/DBoW2/DBoW2/FeatureVector.cpp",
                         "Thirdparty/DBoW2/DBoW2/FeatureVector.h"],
                        ["ref_match_wrap.cpp", "ref_shim_cv/cvshim.hpp", "ref_shim_match/slam_standins.hpp",
                         "ref_shim_match/slam_standins.cpp", "match_oracle.cpp",
                         "orb_oracle.cpp", "msl_oracle.h"]),
}


def build_ref(force=False, name="libsurfel_ref.so"):
    """oracle/_ref/<name>: one of the reference's own source files compiled where it lies, unmodified, against stand-in
    headers (libsurfel_ref.so: src/SurfelFusion.cpp, see oracle/ref_wrap.cpp; liborb_ref.so: src/ORBextractor.cc, see
    oracle/ref_orb_wrap.cpp; libplane_ref.so: src/PlaneExtractor.cpp + include/peac/, see oracle/ref_plane_wrap.cpp).  Built only where /root/reference exists; returns the path or None."""
    so = os.path.join(_HERE, "_ref", name)
    ref_deps, own_deps = _REF_DEPS[name]
    if os.path.exists(os.path.join(REFERENCE_ROOT, ref_deps[0])):
        deps = [os.path.join(REFERENCE_ROOT, d) for d in ref_deps] + [os.path.join(_HERE, d) for d in own_deps]
        if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["make", "-B", "-C", _HERE, "_ref/" + name], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


class RefOrbExtractor:
    """The reference's ORBextractor class itself (oracle/_ref/liborb_ref.so, see build_ref and oracle/ref_orb_wrap.cpp):
    one extractor built, called and destroyed per frame.  `arena=True` (default) serves every allocation of the call
    from a bump arena so that the heap-address tie-break of DistributeOctTree's sort (src/ORBextractor.cc:654) is the
    node creation order; `arena=False` leaves it to glibc.  Used by tests/test_oracle_ref.py and
    tools/ref_octree_tiebreak.py, nowhere else."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, arena=True):
        so = build_ref(name="liborb_ref.so")
        if so is None:
            raise RuntimeError("oracle/_ref/liborb_ref.so is not built and /root/reference is absent")
        self.L = C.CDLL(so)
        sig = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        for f in (self.L.ref_orb_extract, self.L.ref_orb_extract_malloc):
            f.argtypes = sig + [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.L.ref_orb_tables.argtypes = sig + [C.c_void_p] * 6
        self.params = (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.nfeatures, self.nlevels, self.arena = nfeatures, nlevels, arena
        self._levels = None

    def tables(self):
        """(scale, inv_scale, sigma2, inv_sigma2, features_per_level, umax) of the reference's constructor"""
        f = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        n, u = np.zeros(self.nlevels, np.int32), np.zeros(16, np.int32)
        self.L.ref_orb_tables(*self.params, *[_p(a) for a in f], _p(n), _p(u))
        return f + [n, u]

    def __call__(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
        wh = np.zeros((self.nlevels, 2), np.int32)
        pyr = np.zeros(w * h * self.nlevels + 64, np.uint8)  # every level is at most as large as the input
        fn = self.L.ref_orb_extract if self.arena else self.L.ref_orb_extract_malloc
        n = fn(*self.params, _p(gray), w, h, gray.strides[0], _p(kps), _p(desc), cap, _p(wh), _p(pyr))
        if n < 0 or n > cap:
            raise RuntimeError("reference ORB extraction failed (%d)" % n)
        self._levels, off = [], 0
        for lw, lh in wh:
            self._levels.append(pyr[off:off + lw * lh].reshape(lh, lw).copy())
            off += lw * lh
        return kps[:n].copy(), desc[:n].copy()

    def level_image(self, level):
        return self._levels[level]


class _RefMatchProxy:
    """orc_search_* -> ref_search_* of oracle/_ref/libmatch_ref.so (same flat signatures, see oracle/ref_match_wrap.cpp)"""

    def __init__(self, L, prefix="ref_"):
        self._L, self._prefix = L, prefix

    def __getattr__(self, name):
        if not name.startswith("orc_search_"):  # everything but the searches stays with the oracle library
            global _LIB_OVERRIDE
            saved, _LIB_OVERRIDE = _LIB_OVERRIDE, None
            try:
                return getattr(lib(), name)
            finally:
                _LIB_OVERRIDE = saved
        return getattr(self._L, self._prefix + name[4:])


_MATCH_REF = None


def _match_ref():
    global _MATCH_REF
    if _MATCH_REF is None:
        so = build_ref(name="libmatch_ref.so")
        if so is None:
            raise RuntimeError("oracle/_ref/libmatch_ref.so is not built and /root/reference is absent")
        _MATCH_REF = C.CDLL(so)
    return _MATCH_REF


class reference_matcher:
    """context manager: inside it search_by_projection_frame / _points / _keyframe, search_by_bow and
    search_for_triangulation of this module run the REFERENCE's own src/ORBmatcher.cc (oracle/_ref/libmatch_ref.so)
    instead of the oracle restatement.  Slots the oracle reports as -3 (assigned, then reset by the rotation check) come
    back as -1 -- the reference stores NULL for both.  Used by tests/test_oracle_ref.py, nowhere else."""

    def __init__(self, library=None, prefix="ref_"):
        """library / prefix: another ctypes library exporting <prefix>search_* with the same signatures (the adapter harness
        of tests/test_adapters_on_mock_abi.py)"""
        self._lib, self._prefix = library, prefix

    def __enter__(self):
        global _LIB_OVERRIDE
        _LIB_OVERRIDE = _RefMatchProxy(self._lib if self._lib is not None else _match_ref(), self._prefix)
        return self

    def __exit__(self, *exc):
        global _LIB_OVERRIDE
        _LIB_OVERRIDE = None
        return False


def ref_descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return _match_ref().ref_descriptor_distance(_p(a), _p(b))


def ref_fuse(geom, Tcw, th, log_scale_factor, inv_level_sigma2, mps, kf, library=None, name="ref_fuse"):
    """ORBmatcher::Fuse of the reference's own source on the arguments of fuse_search -> (nFused, fused_idx per map point:
    the KeyFrame keypoint the map point was added at, -1 if it was not fused)"""
    L = library if library is not None else _match_ref()
    fn = getattr(L, name)
    fn.argtypes = ([C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int] + [C.c_void_p] * 5 +
                           [C.c_int] + [C.c_void_p] * 4 + [C.c_void_p])
    a = lambda x, dt: np.ascontiguousarray(x, dt)
    ils = a(inv_level_sigma2, np.float32)
    margs = [a(mps["valid"], np.uint8), a(mps["world"], np.float32), a(mps["normal"], np.float32), a(mps["dist"], np.float32),
             a(mps["desc"], np.uint8)]
    kargs = [a(kf["xy"], np.float32), a(kf["octave"], np.int32), a(kf["uright"], np.float32), a(kf["desc"], np.uint8)]
    fi = np.zeros(len(margs[0]), np.int32)
    n = fn(_p(geom), _p(a(Tcw, np.float32)), th, log_scale_factor, _p(ils), len(margs[0]), *[_p(x) for x in margs],
           len(kargs[1]), *[_p(x) for x in kargs], _p(fi))
    return n, fi


_PLANE_REF = None


def _plane_ref():
    global _PLANE_REF
    if _PLANE_REF is None:
        so = build_ref(name="libplane_ref.so")
        if so is None:
            raise RuntimeError("oracle/_ref/libplane_ref.so is not built and /root/reference is absent")
        L = C.CDLL(so)
        head = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5
        L.ref_plane_prestage.argtypes = head + [C.c_void_p] * 4
        L.ref_plane_run.argtypes = head + [C.c_void_p] * 5 + [C.c_int]
        _PLANE_REF = L
    return _PLANE_REF


def ref_plane_prestage(depth_u16, K=(525.0, 525.0, 319.5, 239.5), depth_map_factor=1.0 / 5000.0):
    """plane_prestage() computed by the reference's own src/PlaneExtractor.cpp + include/peac/ (oracle/_ref/libplane_ref.so,
    see oracle/ref_plane_wrap.cpp): same outputs, center / normal of blocks with N < 4 returned as 0."""
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth_u16.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    nb = (h2 // 10) * (w2 // 10)
    cloud, blocks = np.zeros((h2, w2, 3), np.float64), np.zeros(nb, BLOCK_DTYPE)
    seed, edges = np.zeros(nb, np.uint8), np.zeros(nb, np.uint8)
    rc = _plane_ref().ref_plane_prestage(_p(depth_u16), w, h, depth_u16.strides[0] // 2, K[0], K[1], K[2], K[3], depth_map_factor,
                                         _p(cloud), _p(blocks), _p(seed), _p(edges))
    if rc != 0:
        raise RuntimeError("reference plane pre-stage failed (%d)" % rc)
    return cloud, blocks, seed, edges


def ref_plane_run(depth_u16, K=(525.0, 525.0, 319.5, 239.5), depth_map_factor=1.0 / 5000.0, cap=64):
    """Frame::ExtractPlanes' readColorImage / readDepthImage / runPlaneDetection (src/Frame.cc:607-609) by the reference's own
    code -> (membershipImg[h2,w2] i32, dict(normal, center, N, vertices) of the extracted planes)"""
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth_u16.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    mem = np.zeros((h2, w2), np.int32)
    nrm, cen = np.zeros((cap, 3)), np.zeros((cap, 3))
    N, nv = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = _plane_ref().ref_plane_run(_p(depth_u16), w, h, depth_u16.strides[0] // 2, K[0], K[1], K[2], K[3], depth_map_factor,
                                   _p(mem), _p(nrm), _p(cen), _p(N), _p(nv), cap)
    if n < 0:
        raise RuntimeError("reference plane detection failed (%d)" % n)
    n = min(n, cap)
    return mem, dict(normal=nrm[:n], center=cen[:n], N=N[:n], vertices=nv[:n])


def ref_plane_timed(depth_u16, K=(525.0, 525.0, 319.5, 239.5), depth_map_factor=1.0 / 5000.0, full=False, membership=None):
    """timing form of the reference's own plane code (bench.py's reference arm; thread-safe, no outputs): the pre-stage
    (readDepthImage + PlaneSeg constructors + initGraph) or, full=True, runPlaneDetection -> graph nodes / plane_num_"""
    depth_u16 = np.ascontiguousarray(depth_u16, np.uint16)
    h, w = depth_u16.shape
    L = _plane_ref()
    L.ref_plane_timed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 5 + [C.c_int, C.c_void_p]
    n = L.ref_plane_timed(_p(depth_u16), w, h, depth_u16.strides[0] // 2, K[0], K[1], K[2], K[3], depth_map_factor, int(full),
                          _p(membership))
    if n < 0:
        raise RuntimeError("reference plane code failed (%d)" % n)
    return n


class RefSurfelFusion:
    """The reference's SurfelFusion class itself (oracle/_ref, see build_ref): fuseInitializeMap + read-back of the
    private superpixel buffers.  Used by tests/test_oracle_ref.py to check the oracle restatement, nowhere else."""

    def __init__(self, w=640, h=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, fuseFar=30.0, fuseNear=0.5, real_threads=False):
        """real_threads: the build with the real <thread> (the reference's ten racing slices) -- for timing only"""
        so = build_ref(name="libsurfel_ref_threads.so" if real_threads else "libsurfel_ref.so")
        if so is None:
            raise RuntimeError("oracle/_ref/libsurfel_ref*.so is not built and /root/reference is absent")
        self.L = C.CDLL(so)
        self.L.ref_surfel_create.restype = C.c_void_p
        self.L.ref_surfel_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6
        self.L.ref_surfel_destroy.argtypes = [C.c_void_p]
        self.L.ref_surfel_fuse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        self.L.ref_surfel_index.argtypes = [C.c_void_p, C.c_void_p]
        self.L.ref_surfel_seeds.argtypes = [C.c_void_p, C.c_void_p]
        self.w, self.h = w, h
        self.hd = self.L.ref_surfel_create(w, h, fx, fy, cx, cy, fuseFar, fuseNear)

    def __del__(self):
        if getattr(self, "hd", None):
            self.L.ref_surfel_destroy(self.hd)
            self.hd = None

    def fuse(self, ref, gray, depth, membership, Twc, local):
        """fuseInitializeMap: mutates `local` in place, returns the new surfels."""
        h, w = self.h, self.w
        buf = np.zeros(h * w + 3 * w + 16, np.uint8)  # zero bytes behind the image: the reference reads cv::Vec3b on it
        buf[:h * w] = np.ascontiguousarray(gray, np.uint8).ravel()
        d = np.ascontiguousarray(depth, np.float32)
        m = np.ascontiguousarray(membership, np.int32)
        T = np.ascontiguousarray(Twc, np.float32)
        new = np.zeros((w // 8) * (h // 8), SURFEL_DTYPE)
        n = self.L.ref_surfel_fuse(self.hd, int(ref), _p(buf), w, _p(d), _p(m), _p(T), _p(local), len(local), _p(new), len(new))
        if n < 0:
            raise RuntimeError("the reference resized localSurfels")
        return new[:n].copy()

    def set_map(self, local):
        """timing form: the local map in a std::vector inside the library (as Map::mvLocalSurfels in the reference)"""
        self.L.ref_surfel_set_map.argtypes = [C.c_void_p, C.c_int64]
        local = np.ascontiguousarray(local)
        self.L.ref_surfel_set_map(_p(local), len(local))

    def fuse_resident(self, ref, gray, depth, membership, Twc):
        """fuseInitializeMap on the library-owned map (no copy of the map in or out) -> number of new surfels"""
        self.L.ref_surfel_fuse_resident.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        h, w = self.h, self.w
        buf = np.zeros(h * w + 3 * w + 16, np.uint8)
        buf[:h * w] = np.ascontiguousarray(gray, np.uint8).ravel()
        d = np.ascontiguousarray(depth, np.float32)
        m = np.ascontiguousarray(membership, np.int32)
        T = np.ascontiguousarray(Twc, np.float32)
        return self.L.ref_surfel_fuse_resident(self.hd, int(ref), _p(buf), w, _p(d), _p(m), _p(T))

    def compact_resident(self):
        """the tail of SurfelMapping::fuseMap on the library-owned map -> new map size"""
        self.L.ref_surfel_compact_resident.restype = C.c_int64
        return self.L.ref_surfel_compact_resident()

    def index(self):
        out = np.zeros((self.h, self.w), np.int32)
        self.L.ref_surfel_index(self.hd, _p(out))
        return out

    def seeds(self):
        out = np.zeros((self.w // 8) * (self.h // 8), SEED_DTYPE)
        self.L.ref_surfel_seeds(self.hd, _p(out))
        return out


class RefSurfelMapping:
    """The reference's SurfelMapping class itself (oracle/_ref/libmapping_ref.so: src/SurfelMapping.cpp + src/SurfelFusion.cpp
    compiled unmodified, see oracle/ref_mapping_wrap.cpp): keyframes go through InsertKeyFrame + ProcessNewKeyFrame."""

    def __init__(self, w=640, h=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, far=30.0, near=0.5):
        so = build_ref(name="libmapping_ref.so")
        if so is None:
            raise RuntimeError("oracle/_ref/libmapping_ref.so is not built and /root/reference is absent")
        L = self.L = C.CDLL(so)
        L.ref_mapping_create.restype = C.c_void_p
        L.ref_mapping_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6
        L.ref_mapping_destroy.argtypes = [C.c_void_p]
        L.ref_mapping_keyframe.restype = C.c_int64
        L.ref_mapping_keyframe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_mapping_last_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        for f in (L.ref_mapping_local, L.ref_mapping_inactive):
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        self.w, self.h = w, h
        self.hd = L.ref_mapping_create(w, h, fx, fy, cx, cy, far, near)

    def __del__(self):
        if getattr(self, "hd", None):
            self.L.ref_mapping_destroy(self.hd)
            self.hd = None

    def keyframe(self, gray, depth, membership, Twc, reference_index):
        """-> (posesToAdd, posesToRemove) that getAddRemovePoses handed to moveAddSurfels for this keyframe"""
        h, w = self.h, self.w
        buf = np.zeros(h * w + 3 * w + 16, np.uint8)
        buf[:h * w] = np.ascontiguousarray(gray, np.uint8).ravel()
        d = np.ascontiguousarray(depth, np.float32)
        m = np.ascontiguousarray(membership, np.int32)
        T = np.ascontiguousarray(Twc, np.float32)
        self.L.ref_mapping_keyframe(self.hd, _p(buf), w, h, _p(d), _p(m), _p(T), int(reference_index))
        add, rem, nr = np.zeros(4096, np.int32), np.zeros(4096, np.int32), np.zeros(1, np.int32)
        na = self.L.ref_mapping_last_lists(self.hd, _p(add), _p(rem), 4096, _p(nr))
        return add[:na].copy(), rem[:int(nr[0])].copy()

    def _get(self, fn):
        n = fn(self.hd, None, 0)
        out = np.zeros(n, SURFEL_DTYPE)
        fn(self.hd, _p(out), n)
        return out

    def local(self):
        return self._get(self.L.ref_mapping_local)

    def inactive(self):
        return self._get(self.L.ref_mapping_inactive)


class SurfelMappingOracle:
    """The part of SurfelMapping that moveAddSurfels touches (src/SurfelMapping.cpp:194-304)."""

    def __init__(self):
        self.L = lib()
        self.L.orc_mapping_create.restype = C.c_void_p
        self.L.orc_mapping_destroy.argtypes = [C.c_void_p]
        self.L.orc_move_add_surfels.restype = C.c_int64
        self.L.orc_move_add_surfels.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.L.orc_mapping_inactive.restype = C.c_int64
        self.L.orc_mapping_inactive.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        self.hd = self.L.orc_mapping_create()

    def __del__(self):
        if getattr(self, "hd", None):
            self.L.orc_mapping_destroy(self.hd)
            self.hd = None

    def move_add(self, local, poses_to_remove, poses_to_add, extra=0):
        """Returns the new local array (moved-out surfels stay as updateTimes == 0 slots)."""
        rem = np.ascontiguousarray(poses_to_remove, np.int32)
        add = np.ascontiguousarray(poses_to_add, np.int32)
        cap = len(local) + self.inactive_size() + extra
        buf = np.zeros(cap, SURFEL_DTYPE)
        buf[:len(local)] = local
        n = self.L.orc_move_add_surfels(self.hd, _p(buf), len(local), cap, _p(rem), len(rem), _p(add), len(add))
        if n < 0:
            raise RuntimeError("orc_move_add_surfels failed: %d" % n)
        return buf[:n].copy()

    def inactive_size(self):
        return int(self.L.orc_mapping_inactive(self.hd, None, 0))

    def inactive(self):
        out = np.zeros(self.inactive_size(), SURFEL_DTYPE)
        self.L.orc_mapping_inactive(self.hd, _p(out), len(out))
        return out


# ------------------------------------------------------------------- frame glue

def cvt_gray(img, rgb_order=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.zeros((h, w), np.uint8)
    L = lib()
    L.orc_cvt_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_cvt_gray(_p(img), w, h, img.strides[0], ch, int(rgb_order), _p(out), w)
    return out


def depth_to_float(d16, factor):
    d16 = np.ascontiguousarray(d16, np.uint16)
    out = np.zeros(d16.shape, np.float32)
    L = lib()
    L.orc_depth_to_float.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
    L.orc_depth_to_float(_p(d16), d16.size, factor, _p(out))
    return out


def undistort_keypoints(xy, K4, D5, force=False):
    xy = np.ascontiguousarray(xy, np.float32)
    K4 = np.ascontiguousarray(K4, np.float32)
    D5 = np.ascontiguousarray(D5, np.float32)
    out = np.zeros_like(xy)
    L = lib()
    fn = L.orc_undistort_points if force else L.orc_undistort_keypoints
    fn.argtypes = [C.c_int] + [C.c_void_p] * 4
    fn(len(xy), _p(xy), _p(K4), _p(D5), _p(out))
    return out


def stereo_from_rgbd(kp_xy, kpun_xy, depth, mbf):
    kp_xy = np.ascontiguousarray(kp_xy, np.float32)
    kpun_xy = np.ascontiguousarray(kpun_xy, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    n = len(kp_xy)
    ur, kd = np.zeros(n, np.float32), np.zeros(n, np.float32)
    L = lib()
    L.orc_stereo_from_rgbd.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_float] + [C.c_void_p] * 2
    L.orc_stereo_from_rgbd(n, _p(kp_xy), _p(kpun_xy), _p(depth), depth.shape[1], mbf, _p(ur), _p(kd))
    return ur, kd
