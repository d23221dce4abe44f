"""SurfelFusion mirror (include/SurfelFusion.h:44-139 of the reference) over the CUDA C ABI."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

SURFEL_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                         ("size", "<f4"), ("color", "<f4"), ("r", "<i4"), ("g", "<i4"), ("b", "<i4"),
                         ("weight", "<f4"), ("updateTimes", "<i4"), ("lastUpdate", "<i4")])

SEED_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"),
                       ("normX", "<f4"), ("normY", "<f4"), ("normZ", "<f4"),
                       ("posX", "<f4"), ("posY", "<f4"), ("posZ", "<f4"),
                       ("viewCos", "<f4"), ("meanDepth", "<f4"), ("meanIntensity", "<f4"),
                       ("r", "<i4"), ("g", "<i4"), ("b", "<i4"),
                       ("fused", "<i4"), ("stable", "<i4"), ("use", "<i4")])


class SurfelFusion:
    """SurfelFusion(width, height, fx, fy, cx, cy, fuseFar, fuseNear) with a device-resident local map.

    fuseInitializeMap(referenceFrameIndex, image, depth, planeMembershipImg, pose, localSurfels, newSurfels)
    of the reference mutates a host vector; here the map lives on the GPU between calls and is moved with
    upload_map / download_map at the explicit sync points."""

    def __init__(self, width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, fuseFar=30.0, fuseNear=0.5,
                 max_surfels=1 << 20, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.msl_surfel_create(width, height, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                                        C.c_float(fuseFar), C.c_float(fuseNear), C.c_int64(max_surfels), device,
                                        C.byref(self._h)))
        self.width, self.height = width, height
        self.nseeds = (width // 8) * (height // 8)
        self._L.msl_surfel_map_size.restype = C.c_int64
        self._L.msl_surfel_map_size.argtypes = [C.c_void_p]
        self._L.msl_surfel_stream.restype = C.c_void_p

    def close(self):
        if getattr(self, "_h", None):
            self._L.msl_surfel_destroy.argtypes = [C.c_void_p]
            self._L.msl_surfel_destroy(self._h)
            self._h = None

    __del__ = close

    def upload_map(self, local):
        local = np.ascontiguousarray(local)
        assert local.dtype == SURFEL_DTYPE
        check(self._L.msl_surfel_upload_map(self._h, ptr(local), C.c_int64(len(local))))

    def map_size(self):
        return int(self._L.msl_surfel_map_size(self._h))

    def download_map(self):
        n = C.c_int64()
        check(self._L.msl_surfel_download_map(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.zeros(n.value, SURFEL_DTYPE)
        check(self._L.msl_surfel_download_map(self._h, ptr(out), C.c_int64(len(out)), C.byref(n)))
        return out

    def download_changed(self, ref):
        """(indices, records) of the surfels the last non-compacting fuse call with reference index `ref` updated or deleted"""
        n = C.c_int64()
        check(self._L.msl_surfel_download_changed(self._h, int(ref), None, None, C.c_int64(0), C.byref(n)))
        idx, rec = np.zeros(n.value, np.int32), np.zeros(n.value, SURFEL_DTYPE)
        if n.value:
            check(self._L.msl_surfel_download_changed(self._h, int(ref), ptr(idx), ptr(rec), C.c_int64(n.value), C.byref(n)))
        return idx, rec

    def moveAddSurfels(self, poses_to_remove, poses_to_add):
        """SurfelMapping::moveAddSurfels (src/SurfelMapping.cpp:194-304) on the device-resident maps; the two lists
        are getAddRemovePoses' outputs.  Returns (moved_out, moved_in, local_size)."""
        rem = np.ascontiguousarray(poses_to_remove, np.int32)
        add = np.ascontiguousarray(poses_to_add, np.int32)
        stats = np.zeros(3, np.int64)
        check(self._L.msl_surfel_move_add(self._h, ptr(rem), C.c_int(len(rem)), ptr(add), C.c_int(len(add)), ptr(stats)))
        return tuple(int(v) for v in stats)

    def download_inactive(self):
        """Map::mvInactiveSurfels in the reference's order."""
        n = C.c_int64()
        check(self._L.msl_surfel_download_inactive(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.zeros(n.value, SURFEL_DTYPE)
        check(self._L.msl_surfel_download_inactive(self._h, ptr(out), C.c_int64(len(out)), C.byref(n)))
        return out

    def fuseInitializeMap(self, referenceFrameIndex, image, depth, planeMembershipImg, pose, compact=False):
        """Returns (newSurfels, stats) with stats = (n_new, n_updated, n_deleted, map_size)."""
        image = np.ascontiguousarray(image, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        mem = np.ascontiguousarray(planeMembershipImg, np.int32)
        pose = np.ascontiguousarray(pose, np.float32)
        new = np.zeros(self.nseeds, SURFEL_DTYPE)
        stats = np.zeros(4, np.int64)
        check(self._L.msl_surfel_fuse(self._h, int(referenceFrameIndex), ptr(image), C.c_int(image.strides[0]),
                                      ptr(depth), ptr(mem), ptr(pose), ptr(new), C.c_int(len(new)), int(compact),
                                      ptr(stats)))
        return new[:stats[0]].copy(), tuple(int(v) for v in stats)

    def fuse_batch(self, ref0, images, depths, memberships, poses, compact=True):
        images = np.ascontiguousarray(images, np.uint8)
        depths = np.ascontiguousarray(depths, np.float32)
        mems = np.ascontiguousarray(memberships, np.int32)
        poses = np.ascontiguousarray(poses, np.float32)
        B = images.shape[0]
        stats = np.zeros(4, np.int64)
        check(self._L.msl_surfel_fuse_batch(self._h, int(ref0), ptr(images), C.c_int(images.strides[1]), ptr(depths),
                                            ptr(mems), ptr(poses), C.c_int(B), int(compact), ptr(stats)))
        return tuple(int(v) for v in stats)

    def fuse_batch_dev(self, ref0, d_gray, gray_stride, gray_frame_stride, d_depth, d_mem, poses, batch, compact=True):
        poses = np.ascontiguousarray(poses, np.float32)
        check(self._L.msl_surfel_fuse_batch_dev(self._h, int(ref0), ptr(d_gray), C.c_int(gray_stride),
                                                C.c_size_t(gray_frame_stride), ptr(d_depth), ptr(d_mem), ptr(poses),
                                                C.c_int(batch), int(compact)))

    def read_stats(self):
        stats = np.zeros(4, np.int64)
        check(self._L.msl_surfel_read_stats(self._h, ptr(stats)))
        return tuple(int(v) for v in stats)

    def sync(self):
        check(self._L.msl_surfel_sync(self._h))

    def set_timing(self, mode=1):
        """0 off, 1 scan + apply on every 8th frame (light: for use inside a timed region), 2 every mark of every frame"""
        check(self._L.msl_surfel_set_timing(self._h, int(mode)))

    def chain_times(self):
        """per-frame-chain kernel times in ms: dict(scan, apply, post, list, cmp_apply), frames"""
        out = (C.c_double * 5)()
        n = C.c_int()
        check(self._L.msl_surfel_chain_times(self._h, out, C.byref(n)))
        return dict(zip(("scan", "apply", "post", "list", "cmp_apply"), [float(v) for v in out])), n.value

    def fuse_kernels(self):
        """1: fuseSurfelsKernel runs as k_fuse_one; 2: as k_fuse_scan + k_fuse_apply (MSL_FUSE_ONE=0)"""
        return int(self._L.msl_surfel_fuse_kernels(self._h))

    def set_count_table(self, d_table):
        """device pointer of a (batch, 2) int32 table filled by every following fuse_batch_dev: {new, updated} per frame"""
        check(self._L.msl_surfel_set_count_table(self._h, ptr(d_table) if d_table else None))

    def set_fuse_ctas_per_sm(self, batch_ctas=0, single_ctas=0):
        """CTAs per SM of the persistent fuse kernel inside batches of >= 8 frames / for single frames (0: keep)"""
        check(self._L.msl_surfel_set_fuse_ctas_per_sm(self._h, int(batch_ctas), int(single_ctas)))

    def launch_info(self):
        """launch geometry of the last fuseSurfelsKernel launch (msl_surfel_launch_info)"""
        out = np.zeros(6, np.int32)
        check(self._L.msl_surfel_launch_info(self._h, ptr(out)))
        return dict(zip(("kernels", "form", "persistent", "grid", "warps_per_cta", "segments"), [int(v) for v in out]))

    def fuse_kernel_time(self):
        """(total milliseconds, launches) of the projective fuse scan since the last query."""
        ms, n = C.c_double(), C.c_int()
        check(self._L.msl_surfel_fuse_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def stream(self):
        return self._L.msl_surfel_stream(self._h)

    @property
    def input_stream(self):
        """the stream on which the batched API READS its frame inputs (gray, depth, membership): producers of device-resident
        inputs make it wait for their event, and record on it to learn when a buffer may be overwritten"""
        self._L.msl_surfel_input_stream.restype = C.c_void_p
        return self._L.msl_surfel_input_stream(self._h)

    def superpixels(self, images, depths, memberships, want_index=True):
        """generateSuperPixels for a batch of independent frames -> (seeds[B, nseeds], index[B, H, W])."""
        images = np.ascontiguousarray(images, np.uint8)
        depths = np.ascontiguousarray(depths, np.float32)
        mems = np.ascontiguousarray(memberships, np.int32)
        if images.ndim == 2:
            images, depths, mems = images[None], depths[None], mems[None]
        B = images.shape[0]
        seeds = np.zeros((B, self.nseeds), SEED_DTYPE)
        index = np.zeros((B, self.height, self.width), np.int32) if want_index else None
        check(self._L.msl_surfel_superpixels(self._h, ptr(images), C.c_int(images.strides[1]), ptr(depths), ptr(mems),
                                             C.c_int(B), ptr(seeds), ptr(index)))
        return seeds, index

    def debug_seeds(self):
        out = np.zeros(self.nseeds, SEED_DTYPE)
        check(self._L.msl_surfel_debug_seeds(self._h, ptr(out)))
        return out

    def debug_index(self):
        out = np.zeros((self.height, self.width), np.int32)
        check(self._L.msl_surfel_debug_index(self._h, ptr(out)))
        return out
