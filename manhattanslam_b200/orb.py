"""ORBextractor mirror (include/ORBextractor.h:42-104 of the reference) over the CUDA C ABI."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class ORBextractor:
    """ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST).

    The CUDA handle is sized for one image size; `width`/`height`/`max_batch` must be given up front
    (the reference allocates lazily per call)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 width=640, height=480, max_batch=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        prm = _Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        check(self._L.msl_orb_create(C.byref(prm), width, height, max_batch, device, C.byref(self._h)))
        self.width, self.height, self.max_batch = width, height, max_batch
        self.nlevels = nlevels
        self._L.msl_orb_capacity.argtypes = [C.c_void_p]
        self.capacity = self._L.msl_orb_capacity(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.msl_orb_destroy.argtypes = [C.c_void_p]
            self._L.msl_orb_destroy(self._h)
            self._h = None

    __del__ = close

    # getters, include/ORBextractor.h:58-82
    def GetLevels(self):
        return self.nlevels

    def _factors(self):
        out = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        check(self._L.msl_orb_scale_factors(self._h, *[ptr(o) for o in out]))
        return out

    def GetScaleFactors(self):
        return self._factors()[0]

    def GetInverseScaleFactors(self):
        return self._factors()[1]

    def GetScaleSigmaSquares(self):
        return self._factors()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._factors()[3]

    def GetScaleFactor(self):
        return float(self._factors()[0][1]) if self.nlevels > 1 else 1.0

    def extract_batch(self, gray):
        """gray: (B, H, W) uint8 host array -> list of (keypoints[KP_DTYPE], descriptors[n,32])."""
        gray = np.ascontiguousarray(gray, np.uint8)
        if gray.ndim == 2:
            gray = gray[None]
        B, H, W = gray.shape
        assert (H, W) == (self.height, self.width) and B <= self.max_batch
        kps = np.empty((B, self.capacity), KP_DTYPE)
        desc = np.empty((B, self.capacity, 32), np.uint8)
        counts = np.zeros(B, np.int32)
        check(self._L.msl_orb_extract(self._h, ptr(gray), C.c_int(W), C.c_size_t(H * W), C.c_int(B),
                                      ptr(kps), ptr(desc), ptr(counts)))
        return [(kps[b, :counts[b]].copy(), desc[b, :counts[b]].copy()) for b in range(B)]

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors); mask is ignored as in the reference."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2  # assert(image.type()==CV_8UC1), :819
        return self.extract_batch(image)[0]

    def extract_dev(self, d_gray_ptr, stride, frame_stride, batch, d_kps_ptr, d_desc_ptr, d_counts_ptr):
        check(self._L.msl_orb_extract_dev(self._h, ptr(d_gray_ptr), C.c_int(stride), C.c_size_t(frame_stride),
                                          C.c_int(batch), ptr(d_kps_ptr), ptr(d_desc_ptr), ptr(d_counts_ptr)))

    def sync(self):
        check(self._L.msl_orb_sync(self._h))

    @property
    def stream(self):
        self._L.msl_orb_stream.restype = C.c_void_p
        return self._L.msl_orb_stream(self._h)

    # stage read-back for parity tests
    def debug_level(self, frame, level, blurred=False):
        w, h = C.c_int(), C.c_int()
        check(self._L.msl_orb_debug_level_size(self._h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(self._L.msl_orb_debug_level(self._h, frame, level, int(blurred), ptr(out)))
        return out

    def debug_candidates(self, frame, level):
        cap = 1 << 18
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        check(self._L.msl_orb_debug_candidates(self._h, frame, level, ptr(out), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()
