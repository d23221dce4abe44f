"""Frame glue mirror (Tracking::GrabImage conversions, Frame::UndistortKeyPoints, Frame::ComputeStereoFromRGBD --
src/Tracking.cc:184-211, src/Frame.cc:437-463, 495-513) over the CUDA C ABI."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr


class FrameGlue:
    def __init__(self, width=640, height=480, max_batch=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.msl_glue_create(width, height, max_batch, device, C.byref(self._h)))
        self.width, self.height = width, height
        self._L.msl_glue_stream.restype = C.c_void_p

    def close(self):
        if getattr(self, "_h", None):
            self._L.msl_glue_destroy.argtypes = [C.c_void_p]
            self._L.msl_glue_destroy(self._h)
            self._h = None

    __del__ = close

    def cvtColor(self, img, rgb=True):
        """(B,) H x W x {3,4} uint8 -> (B,) H x W gray (cv::cvtColor *2GRAY, src/Tracking.cc:189-200)."""
        img = np.ascontiguousarray(img, np.uint8)
        single = img.ndim == 3
        if single:
            img = img[None]
        B, H, W, ch = img.shape
        out = np.zeros((B, H, W), np.uint8)
        check(self._L.msl_glue_cvt_gray(self._h, ptr(img), C.c_int(img.strides[1]), C.c_int(ch), int(rgb), C.c_int(B), ptr(out)))
        return out[0] if single else out

    def depthToFloat(self, d16, factor):
        """imDepth.convertTo(CV_32F, mDepthMapFactor) (src/Tracking.cc:205-207)."""
        d16 = np.ascontiguousarray(d16, np.uint16)
        single = d16.ndim == 2
        if single:
            d16 = d16[None]
        out = np.zeros(d16.shape, np.float32)
        check(self._L.msl_glue_depth_to_float(self._h, ptr(d16), C.c_int(d16.shape[0]), C.c_float(factor), ptr(out)))
        return out[0] if single else out

    def keypoints(self, kps, K4, D5=None, depth=None, mbf=40.0):
        """UndistortKeyPoints + ComputeStereoFromRGBD -> (xy_un, uright, depth); kps: KP_DTYPE array."""
        kps = np.ascontiguousarray(kps)
        n = len(kps)
        K4 = np.ascontiguousarray(K4, np.float32)
        D5 = None if D5 is None else np.ascontiguousarray(D5, np.float32)
        xy = np.zeros((n, 2), np.float32)
        ur, kd = np.zeros(n, np.float32), np.zeros(n, np.float32)
        dep = None if depth is None else np.ascontiguousarray(depth, np.float32)
        check(self._L.msl_glue_keypoints(self._h, ptr(kps), C.c_int(n), ptr(K4), ptr(D5), ptr(dep), C.c_float(mbf), ptr(xy),
                                         ptr(ur), ptr(kd)))
        return xy, ur, kd

    def keypoints_dev(self, d_kps, rows, d_counts, batch, K4, D5, d_depth, mbf, d_xy_un, d_uright, d_kdepth, stream=None):
        """batched device form on msl_orb_extract_dev's ragged output (msl_glue_keypoints_dev)"""
        K4 = np.ascontiguousarray(K4, np.float32)
        D5 = None if D5 is None else np.ascontiguousarray(D5, np.float32)
        check(self._L.msl_glue_keypoints_dev(self._h, ptr(d_kps), C.c_int(rows), ptr(d_counts), C.c_int(batch), ptr(K4), ptr(D5),
                                             ptr(d_depth), C.c_float(mbf), ptr(d_xy_un), ptr(d_uright), ptr(d_kdepth),
                                             C.c_void_p(stream or 0)))

    def upload_frames(self, slot, gray_ptr, d16_ptr, batch, factor, aux_ptr=None, aux_ints=0, gray_stride=None, depth_stride_px=None):
        """msl_glue_upload_frames: the sensor frames (host pointers: gray u8, depth u16) of a batch uploaded once for every
        stage, CV_32F depth produced on the device (src/Tracking.cc:205-207).  Returns the device pointers
        (gray, depth16, depth_f32, aux) as ints; asynchronous -- order consumers with frames_wait."""
        dg, d16, dd, da = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._L.msl_glue_upload_frames(self._h, C.c_int(slot), C.c_void_p(gray_ptr), C.c_int(gray_stride or self.width),
                                             C.c_void_p(d16_ptr), C.c_int(depth_stride_px or self.width), C.c_int(batch),
                                             C.c_float(factor), C.c_void_p(aux_ptr or 0), C.c_size_t(aux_ints), C.byref(dg),
                                             C.byref(d16), C.byref(dd), C.byref(da)))
        return dg.value, d16.value, dd.value, da.value

    def frames_wait(self, slot, stream=None):
        """makes `stream` (a cudaStream_t as int; None blocks the host) wait for frame set `slot`"""
        check(self._L.msl_glue_frames_wait(self._h, C.c_int(slot), C.c_void_p(stream or 0)))

    @property
    def stream(self):
        return self._L.msl_glue_stream(self._h)

