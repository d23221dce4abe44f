"""PlaneDetection mirror (src/PlaneExtractor.cpp: readDepthImage :44-76, runPlaneDetection :78-82 = the peac fitter of
include/peac/) over the CUDA C ABI."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

BLOCK_DTYPE = np.dtype([("center", "<f8", 3), ("normal", "<f8", 3), ("mse", "<f8"), ("curvature", "<f8"),
                        ("N", "<i4"), ("nouse", "<i4")])
PLANE_DTYPE = np.dtype([("normal", "<f8", 3), ("center", "<f8", 3), ("N", "<i4"), ("rid", "<i4"), ("vertices", "<i4"),
                        ("pad", "<i4")])  # msl_plane_rec


class PlaneDetection:
    """readDepthImage(depthImg, K, depthMapFactor) + the data-parallel part of runPlaneDetection():
    organized half-resolution cloud, per-10x10-block PCA statistics, graph seeds and edges."""

    def __init__(self, width=640, height=480, max_batch=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.msl_plane_create(width, height, max_batch, device, C.byref(self._h)))
        self.width, self.height, self.max_batch = width, height, max_batch
        self.w2, self.h2 = (width + 1) // 2, (height + 1) // 2
        self.nblocks = (self.w2 // 10) * (self.h2 // 10)
        self._L.msl_plane_stream.restype = C.c_void_p

    def close(self):
        if getattr(self, "_h", None):
            self._L.msl_plane_destroy.argtypes = [C.c_void_p]
            self._L.msl_plane_destroy(self._h)
            self._h = None

    __del__ = close

    def prestage(self, depth_u16, K=(525.0, 525.0, 319.5, 239.5), depthMapFactor=1.0 / 5000.0, want_cloud=True):
        d = np.ascontiguousarray(depth_u16, np.uint16)
        if d.ndim == 2:
            d = d[None]
        B = d.shape[0]
        assert d.shape[1:] == (self.height, self.width) and B <= self.max_batch
        cloud = np.zeros((B, self.h2, self.w2, 3), np.float64) if want_cloud else None
        blocks = np.zeros((B, self.nblocks), BLOCK_DTYPE)
        seed = np.zeros((B, self.nblocks), np.uint8)
        edges = np.zeros((B, self.nblocks), np.uint8)
        Kf = np.asarray(K, np.float32)
        check(self._L.msl_plane_prestage(self._h, ptr(d), C.c_int(self.width), C.c_size_t(self.width * self.height),
                                         C.c_int(B), ptr(Kf), C.c_float(depthMapFactor), ptr(cloud), ptr(blocks),
                                         ptr(seed), ptr(edges)))
        return cloud, blocks, seed, edges

    def detect(self, depth_u16, K=(525.0, 525.0, 319.5, 239.5), depthMapFactor=1.0 / 5000.0, plane_cap=32):
        """readDepthImage + runPlaneDetection (src/Frame.cc:607-609) for a batch of depth images ->
        (membershipImg[B, h2, w2] int32, planes: list of PLANE_DTYPE arrays in extractedPlanes order).
        plane_vertices_[i] of frame b = np.flatnonzero(membership[b] == i)."""
        d = np.ascontiguousarray(depth_u16, np.uint16)
        if d.ndim == 2:
            d = d[None]
        B = d.shape[0]
        assert d.shape[1:] == (self.height, self.width) and B <= self.max_batch
        mem = np.zeros((B, self.h2, self.w2), np.int32)
        cnt = np.zeros(B, np.int32)
        rec = np.zeros((B, plane_cap), PLANE_DTYPE)
        Kf = np.asarray(K, np.float32)
        check(self._L.msl_plane_detect(self._h, ptr(d), C.c_int(self.width), C.c_size_t(self.width * self.height), C.c_int(B),
                                       ptr(Kf), C.c_float(depthMapFactor), ptr(mem), ptr(cnt), ptr(rec), C.c_int(plane_cap)))
        return mem, [rec[b, :min(int(cnt[b]), plane_cap)].copy() for b in range(B)]

    def prestage_dev(self, d_depth, batch, K=(525.0, 525.0, 319.5, 239.5), depthMapFactor=1.0 / 5000.0,
                     d_cloud=None, d_blocks=None, d_seed=None, d_edges=None):
        Kf = np.asarray(K, np.float32)
        check(self._L.msl_plane_prestage_dev(self._h, ptr(d_depth), C.c_int(self.width),
                                             C.c_size_t(self.width * self.height), C.c_int(batch), ptr(Kf),
                                             C.c_float(depthMapFactor), ptr(d_cloud), ptr(d_blocks), ptr(d_seed),
                                             ptr(d_edges)))

    def detect_dev(self, d_depth, batch, K, depthMapFactor, d_membership, d_plane_count, d_planes, plane_cap):
        """device buffers in, device buffers out, asynchronous on the handle's stream (msl_plane_detect_dev)"""
        Kf = np.asarray(K, np.float32)
        check(self._L.msl_plane_detect_dev(self._h, ptr(d_depth), C.c_int(self.width), C.c_size_t(self.width * self.height),
                                           C.c_int(batch), ptr(Kf), C.c_float(depthMapFactor), ptr(d_membership),
                                           ptr(d_plane_count), ptr(d_planes), C.c_int(plane_cap)))

    def debug_profile(self, frames):
        """profile of the last detect call (msl_plane_debug_profile): (frames, 16) int64 -- columns 0..6 phase stamps in ns,
        7 merge steps, 8..13 cycles of ahCluster's sub-phases"""
        out = np.zeros((frames, 16), np.int64)
        check(self._L.msl_plane_debug_profile(self._h, ptr(out), C.c_int(frames)))
        return out

    def sync(self):
        check(self._L.msl_plane_sync(self._h))

    @property
    def stream(self):
        return self._L.msl_plane_stream(self._h)
