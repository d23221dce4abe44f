"""ORBmatcher mirror (include/ORBmatcher.h:39-79: DescriptorDistance, the two tracking-time SearchByProjection
overloads, the relocalisation overload, SearchByBoW, SearchForTriangulation and the search part of Fuse) over the CUDA
C ABI.  The Frame / KeyFrame / MapPoint graph is passed as flat arrays, DBoW2 feature vectors in CSR form."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

GEOM_DTYPE = np.dtype([("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4"),
                       ("mnMinX", "<f4"), ("mnMinY", "<f4"), ("mnMaxX", "<f4"), ("mnMaxY", "<f4"),
                       ("gridWInv", "<f4"), ("gridHInv", "<f4"), ("mb", "<f4"), ("mbf", "<f4"),
                       ("nlevels", "<i4"), ("scaleFactors", "<f4", 16)])


def frame_geom(width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, bf=40.0, scale_factors=None):
    """Frame constants for an undistorted camera (src/Frame.cc:127-147, 466-493)."""
    g = np.zeros(1, GEOM_DTYPE)
    g["fx"], g["fy"], g["cx"], g["cy"] = fx, fy, cx, cy
    g["mnMinX"], g["mnMinY"], g["mnMaxX"], g["mnMaxY"] = 0.0, 0.0, width, height
    g["gridWInv"] = np.float32(64) / np.float32(width)
    g["gridHInv"] = np.float32(48) / np.float32(height)
    g["mbf"] = bf
    g["mb"] = np.float32(bf) / np.float32(fx)
    sf = np.asarray(scale_factors if scale_factors is not None else 1.2 ** np.arange(8), np.float32)
    g["nlevels"] = len(sf)
    g["scaleFactors"][0, :len(sf)] = sf
    return g


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

    def __init__(self, nnratio=0.6, checkOri=True, max_queries=4096, max_train=4096, max_batch=1, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        check(self._L.msl_matcher_create(max_queries, max_train, max_batch, device, C.byref(self._h)))
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self._L.msl_matcher_stream.restype = C.c_void_p

    def close(self):
        if getattr(self, "_h", None):
            self._L.msl_matcher_destroy.argtypes = [C.c_void_p]
            self._L.msl_matcher_destroy(self._h)
            self._h = None

    __del__ = close

    def hamming_best2(self, q, t):
        """q: (B, nq, 32), t: (B, nt, 32) uint8 -> best_idx, best_dist, second_dist (B, nq) int32."""
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        if q.ndim == 2:
            q, t = q[None], t[None]
        B, nq, nt = q.shape[0], q.shape[1], t.shape[1]
        bi, bd, sd = (np.zeros((B, nq), np.int32) for _ in range(3))
        check(self._L.msl_hamming_best2(self._h, ptr(q), nq, ptr(t), nt, B, ptr(bi), ptr(bd), ptr(sd)))
        return bi, bd, sd

    def hamming_best2_dev(self, d_q, nq, d_t, nt, batch, d_bi, d_bd, d_sd):
        check(self._L.msl_hamming_best2_dev(self._h, ptr(d_q), nq, ptr(d_t), nt, batch, ptr(d_bi), ptr(d_bd), ptr(d_sd)))

    def hamming_best2_counts_dev(self, d_q, d_t, rows, d_qcounts, d_tcounts, batch, d_bi, d_bd, d_sd, stream=None):
        check(self._L.msl_hamming_best2_counts_dev(self._h, ptr(d_q), ptr(d_t), rows, ptr(d_qcounts), ptr(d_tcounts), batch,
                                                   ptr(d_bi), ptr(d_bd), ptr(d_sd), C.c_void_p(stream or 0)))

    def hamming_all_pairs(self, q, t):
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        if q.ndim == 2:
            q, t = q[None], t[None]
        B, nq, nt = q.shape[0], q.shape[1], t.shape[1]
        d = np.zeros((B, nq, nt), np.uint16)
        check(self._L.msl_hamming_all_pairs(self._h, ptr(q), nq, ptr(t), nt, B, ptr(d)))
        return d

    @staticmethod
    def DescriptorDistance(a, b):
        return int(np.unpackbits(np.bitwise_xor(np.asarray(a, np.uint8), np.asarray(b, np.uint8))).sum())

    def SearchByProjectionFrame(self, geom, Tcw_cur, Tcw_last, th, last, cur):
        """SearchByProjection(Frame &Cur, const Frame &Last, th).  last/cur: dicts of flat arrays
        (see include/msl_frontend.h).  Returns (nmatches, cur_match)."""
        n_last, n_cur = len(last["octave"]), len(cur["octave"])
        cm = np.zeros(n_cur, np.int32)
        nm = C.c_int32()
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        args = [a(last["has_mp"], np.uint8), a(last["outlier"], np.uint8), a(last["mp_obs"], np.uint8),
                a(last["mp_world"], np.float32), a(last["mp_desc"], np.uint8), a(last["octave"], np.int32),
                a(last["angle"], np.float32)]
        cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["angle"], np.float32),
                 a(cur["uright"], np.float32), a(cur["desc"], np.uint8), a(cur["occupied"], np.uint8)]
        check(self._L.msl_search_by_projection_frame(
            self._h, ptr(geom), ptr(a(Tcw_cur, np.float32)), ptr(a(Tcw_last, np.float32)), C.c_float(th),
            int(self.mbCheckOrientation), n_last, *[ptr(x) for x in args], n_cur, *[ptr(x) for x in cargs], ptr(cm),
            C.byref(nm)))
        return self._ret(lambda: (nm.value, cm))

    def SearchByProjectionFrames_dev(self, geom, th, th_depth, d_kps, d_desc, rows, d_counts, n_frames, d_xy_un, d_uright, d_kdepth,
                                     Tcw, d_cur_match, d_nmatches, stream=None):
        """SearchByProjection(Cur, Last, th) for the n_frames - 1 consecutive pairs of a device-resident batch
        (msl_search_by_projection_frames_dev): Tracking::UpdateLastFrame + Frame::UnprojectStereo on the device."""
        Tcw = np.ascontiguousarray(Tcw, np.float32)
        assert Tcw.shape == (n_frames, 4, 4) or Tcw.shape == (n_frames, 16)
        check(self._L.msl_search_by_projection_frames_dev(
            self._h, ptr(geom), C.c_float(th), int(self.mbCheckOrientation), C.c_float(th_depth), ptr(d_kps), ptr(d_desc),
            C.c_int(rows), ptr(d_counts), C.c_int(n_frames), ptr(d_xy_un), ptr(d_uright), ptr(d_kdepth), ptr(Tcw),
            ptr(d_cur_match), ptr(d_nmatches), C.c_void_p(stream or 0)))

    def SearchByProjectionPoints(self, geom, th, mps, cur):
        """SearchByProjection(Frame &F, const vector<MapPoint*> &, th).  Returns (nmatches, cur_match)."""
        n_mp, n_cur = len(mps["level"]), len(cur["octave"])
        cm = np.zeros(n_cur, np.int32)
        nm = C.c_int32()
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        args = [a(mps["valid"], np.uint8), a(mps["obs"], np.uint8), a(mps["proj_xyr"], np.float32),
                a(mps["level"], np.int32), a(mps["viewcos"], np.float32), a(mps["desc"], np.uint8)]
        cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["uright"], np.float32),
                 a(cur["desc"], np.uint8), a(cur["occupied"], np.uint8)]
        check(self._L.msl_search_by_projection_points(
            self._h, ptr(geom), C.c_float(th), C.c_float(self.mfNNratio), n_mp, *[ptr(x) for x in args], n_cur,
            *[ptr(x) for x in cargs], ptr(cm), C.byref(nm)))
        return self._ret(lambda: (nm.value, cm))

    def SearchByProjectionKeyFrame(self, geom, Tcw_cur, th, ORBdist, kf, cur, log_scale_factor=None):
        """SearchByProjection(Frame &Cur, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
        (src/ORBmatcher.cc:680-797).  kf/cur: dicts of flat arrays (include/msl_frontend.h).
        Returns (nmatches, cur_match)."""
        n_kf, n_cur = len(kf["angle"]), len(cur["octave"])
        if log_scale_factor is None:  # Frame::mfLogScaleFactor = log(mfScaleFactor), src/Frame.cc:82
            log_scale_factor = float(np.float32(np.log(np.float64(np.float32(geom["scaleFactors"][0, 1])))))
        cm = np.zeros(n_cur, np.int32)
        nm = C.c_int32()
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        args = [a(kf["valid"], np.uint8), a(kf["mp_world"], np.float32), a(kf["mp_desc"], np.uint8),
                a(kf["mp_dist"], np.float32), a(kf["angle"], np.float32)]
        cargs = [a(cur["xy"], np.float32), a(cur["octave"], np.int32), a(cur["angle"], np.float32), a(cur["desc"], np.uint8),
                 a(cur["occupied"], np.uint8)]
        check(self._L.msl_search_by_projection_keyframe(
            self._h, ptr(geom), ptr(a(Tcw_cur, np.float32)), C.c_float(th), int(ORBdist), int(self.mbCheckOrientation),
            C.c_float(log_scale_factor), n_kf, *[ptr(x) for x in args], n_cur, *[ptr(x) for x in cargs], ptr(cm),
            C.byref(nm)))
        return self._ret(lambda: (nm.value, cm))

    # ---- deferred batches (msl_matcher_batch_begin / _end): inside `with matcher.batch():` the six searches return a
    # Deferred whose .get() yields the usual tuple once the block has ended (one upload, one CTA per call, one sync)
    class Deferred:
        def __init__(self, thunk):
            self._thunk, self._done = thunk, False

        def get(self):
            if not self._done:
                raise RuntimeError("the batch is still open")
            return self._thunk()

    class _Batch:
        def __init__(self, m):
            self.m = m

        def __enter__(self):
            check(self.m._L.msl_matcher_batch_begin(self.m._h))
            self.m._open = []
            return self.m

        def __exit__(self, *exc):
            pend, self.m._open = self.m._open, None
            check(self.m._L.msl_matcher_batch_end(self.m._h))
            for d in pend:
                d._done = True
            return False

    def batch(self):
        return ORBmatcher._Batch(self)

    def set_timing(self, on=True):
        check(self._L.msl_matcher_set_timing(self._h, int(on)))

    def last_execution(self):
        """(device ms of the last execution: upload + kernels + download, calls it ran)"""
        ms, n = C.c_double(), C.c_int()
        check(self._L.msl_matcher_last_execution(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def _ret(self, thunk):
        if getattr(self, "_open", None) is None:
            return thunk()
        d = ORBmatcher.Deferred(thunk)
        self._open.append(d)
        return d

    @staticmethod
    def _csr(fv):
        """DBoW2::FeatureVector (dict node id -> feature indices) -> (ids u32 ascending, offsets i32, features i32)"""
        if isinstance(fv, tuple) and len(fv) == 3 and isinstance(fv[0], np.ndarray):
            return fv  # already packed
        items = sorted(fv.items())
        ids = np.asarray([k for k, _ in items], np.uint32)
        off = np.zeros(len(items) + 1, np.int32)
        for i, (_, v) in enumerate(items):
            off[i + 1] = off[i] + len(v)
        feat = np.asarray([x for _, v in items for x in v], np.int32)
        return ids, off, feat

    def SearchByBoW(self, kf, f):
        """SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (src/ORBmatcher.cc:146-255).
        kf: dict(featvec, valid, desc, angle); f: dict(featvec, desc, angle).  Returns (nmatches, f_match)."""
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        kid, koff, kfeat = self._csr(kf["featvec"])
        fid, foff, ffeat = self._csr(f["featvec"])
        kv, kd, ka = a(kf["valid"], np.uint8), a(kf["desc"], np.uint8), a(kf["angle"], np.float32)
        fd, fa = a(f["desc"], np.uint8), a(f["angle"], np.float32)
        fm = np.zeros(len(fa), np.int32)
        nm = C.c_int32()
        check(self._L.msl_search_by_bow(self._h, C.c_float(self.mfNNratio), int(self.mbCheckOrientation), len(kid), ptr(kid),
                                        ptr(koff), ptr(kfeat), len(fid), ptr(fid), ptr(foff), ptr(ffeat), len(ka), ptr(kv),
                                        ptr(kd), ptr(ka), len(fa), ptr(fd), ptr(fa), ptr(fm), C.byref(nm)))
        return self._ret(lambda: (nm.value, fm))

    def SearchForTriangulation(self, kf1, kf2, F12, Cw1, Tcw2, K2, scale_factors2, level_sigma2_2, bOnlyStereo=False):
        """SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vMatchedPairs, bOnlyStereo)
        (src/ORBmatcher.cc:257-406).  kf1: dict(featvec, has_mp, uright, xy, angle, desc); kf2: the same + octave.
        Returns (nmatches, matches12); vMatchedPairs = [(i, matches12[i]) for i where matches12[i] >= 0]."""
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        id1, off1, ft1 = self._csr(kf1["featvec"])
        id2, off2, ft2 = self._csr(kf2["featvec"])
        sf, ls = a(scale_factors2, np.float32), a(level_sigma2_2, np.float32)
        a1 = [a(kf1["has_mp"], np.uint8), a(kf1["uright"], np.float32), a(kf1["xy"], np.float32), a(kf1["angle"], np.float32),
              a(kf1["desc"], np.uint8)]
        a2 = [a(kf2["has_mp"], np.uint8), a(kf2["uright"], np.float32), a(kf2["xy"], np.float32), a(kf2["octave"], np.int32),
              a(kf2["angle"], np.float32), a(kf2["desc"], np.uint8)]
        m12 = np.zeros(len(a1[0]), np.int32)
        nm = C.c_int32()
        check(self._L.msl_search_for_triangulation(
            self._h, ptr(a(F12, np.float32)), ptr(a(Cw1, np.float32)), ptr(a(Tcw2, np.float32)), ptr(a(K2, np.float32)),
            int(bOnlyStereo), int(self.mbCheckOrientation), len(sf), ptr(sf), ptr(ls), len(id1), ptr(id1), ptr(off1), ptr(ft1),
            len(id2), ptr(id2), ptr(off2), ptr(ft2), len(a1[0]), *[ptr(x) for x in a1], len(a2[0]), *[ptr(x) for x in a2],
            ptr(m12), C.byref(nm)))
        return self._ret(lambda: (nm.value, m12))

    def Fuse(self, geom, Tcw, mps, kf, inv_level_sigma2, th=3.0, log_scale_factor=None):
        """The search part of Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th) (src/ORBmatcher.cc:408-519).
        mps: dict(valid, world, normal, dist, desc); kf: dict(xy, octave, uright, desc).
        Returns (nFused, best_idx, best_dist); the Replace / AddObservation bookkeeping (:521-541) is the caller's."""
        if log_scale_factor is None:
            log_scale_factor = float(np.float32(np.log(np.float64(np.float32(geom["scaleFactors"][0, 1])))))
        a = lambda x, dt: np.ascontiguousarray(x, dt)
        ils = a(inv_level_sigma2, np.float32)
        margs = [a(mps["valid"], np.uint8), a(mps["world"], np.float32), a(mps["normal"], np.float32),
                 a(mps["dist"], np.float32), a(mps["desc"], np.uint8)]
        kargs = [a(kf["xy"], np.float32), a(kf["octave"], np.int32), a(kf["uright"], np.float32), a(kf["desc"], np.uint8)]
        bi = np.zeros(len(margs[0]), np.int32)
        bd = np.zeros(len(margs[0]), np.int32)
        nf = C.c_int32()
        check(self._L.msl_fuse_search(self._h, ptr(geom), ptr(a(Tcw, np.float32)), C.c_float(th), C.c_float(log_scale_factor),
                                      ptr(ils), len(margs[0]), *[ptr(x) for x in margs], len(kargs[1]), *[ptr(x) for x in kargs],
                                      ptr(bi), ptr(bd), C.byref(nf)))
        return self._ret(lambda: (nf.value, bi, bd))

    def ComputeDistinctiveDescriptors(self, desc_lists):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263) for a batch of map points.
        desc_lists: list of (N_k, 32) uint8 arrays (a point's observed descriptors).  Returns (best_idx, best_median)."""
        n = len(desc_lists)
        off = np.zeros(n + 1, np.int32)
        for k, d in enumerate(desc_lists):
            off[k + 1] = off[k] + len(d)
        desc = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_lists] +
                                                   [np.zeros((0, 32), np.uint8)]))
        bi = np.zeros(n, np.int32)
        bm = np.zeros(n, np.int32)
        check(self._L.msl_distinctive_descriptors(self._h, n, ptr(off), ptr(desc), ptr(bi), ptr(bm)))
        return bi, bm
