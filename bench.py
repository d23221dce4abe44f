#!/usr/bin/env python3
"""bench.py -- RGB-D front-end frames/s on synthetic RGB-D (BASELINE.json metric).

One step = one pass of the front-end over one batch of synthetic RGB-D frames per GPU.  The default workload is the
configuration the metric is quoted on (BASELINE.json configs 2-4 in one step, 640x480, batch 64, 5 M-surfel map):
ORB extraction, brute-force Hamming match of every frame's descriptors against the next frame's, ORBmatcher::SearchByProjection
of every frame against its predecessor (the frame glue + Tracking::UpdateLastFrame on the device), the plane pre-stage on the
u16 depth and the projective surfel fusion of the 64-frame stream into a device-resident map (superpixels batched,
fuse / initialise / compact frame by frame in order).  --workload selects BASELINE.json's other configurations
(orb_match_640x480_b64, plane_640x480_b256, surfel_640x480_b64_map5M, frontend_1280x960_b64_map5M).
Frames are independent across ranks (one chunk and one map replica per rank, weak scaling); the only collective is one NCCL
all-gather of the per-frame count table {keypoints, new surfels, updated surfels}, on its own stream.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]     our CUDA path
  python bench.py --impl reference ...                                      the reference's CPU path on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rgbd_frontend_frames_per_s"
UNIT = "frames/s"
MAP_STEADY_FRACTION = 0.89  # measured with the oracle on this generator (see make_inputs)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

ALL_STAGES = ["orb", "hamming_match", "search_by_projection", "plane_prestage", "surfel_fuse"]
MBF, TH_DEPTH_FACTOR, TH_PROJ = 40.0, 40.0, 15.0  # Camera.bf / ThDepth of Example/TUM3.yaml; th of TrackWithMotionModel (src/Tracking.cc:1249)
WORKLOADS = {
    # name: (width, height, frames per GPU per step, surfels per GPU, stages)
    "frontend_640x480_b64_map5M": (640, 480, 64, 5_000_000, ALL_STAGES),          # the headline (configs 2 + 3 + 4 in one step)
    "orb_match_640x480_b64": (640, 480, 64, 0, ["orb", "hamming_match", "search_by_projection"]),  # BASELINE.json config 2
    "plane_640x480_b256": (640, 480, 256, 0, ["plane_prestage"]),                   # config 3
    "surfel_640x480_b64_map5M": (640, 480, 64, 5_000_000, ["surfel_fuse"]),         # config 4
    "frontend_1280x960_b64_map5M": (1280, 960, 64, 5_000_000, ALL_STAGES),          # config 5: one rank's 64 of the 512 frames
    # SURVEY.md section 8(f) row f2 widened into the step: PlaneDetection::runPlaneDetection (ahCluster + refineDetails) on the
    # device produces the membership image the surfel stage consumes, as src/Tracking.cc:227-229 hands it over.  (With real
    # membership most of a piecewise-planar synthetic scene is "in a plane", where ManhattanSLAM builds no surfels.)
    "frontend_peac_640x480_b64_map5M": (640, 480, 64, 5_000_000, ["orb", "hamming_match", "search_by_projection", "plane_detect", "surfel_fuse"]),
}
DEFAULT_WORKLOAD = "frontend_640x480_b64_map5M"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's frames per GPU per step (diagnostic)")
    ap.add_argument("--surfels", type=int, default=-1, help="override the workload's surfels per GPU (diagnostic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-kernel pass, the parity replay and the widened diagnostics")
    ap.add_argument("--widened-only", nargs="?", const="matcher", default="", choices=["matcher", "peac"],
                    help="internal: print one part of the `widened` object alone (run_ours calls this in child processes so "
                         "that a fault in a diagnostic can never take the bench line down)")
    ap.add_argument("--cpu-frames", type=int, default=64, help="upper bound of the frames in the bounded CPU sample")
    ap.add_argument("--only", default="", help="diagnostic: comma list of stages (orb,match,plane,surfel) the device-resident "
                                               "step runs; anything but the workload's own stages is not a bench value")
    a = ap.parse_args()
    w, h, batch, surfels, stages = WORKLOADS[a.workload]
    a.W, a.H, a.stages = w, h, list(stages)
    a.diagnostic = bool(a.only) or a.batch > 0 or a.surfels >= 0
    a.batch = a.batch if a.batch > 0 else batch
    a.surfels = a.surfels if a.surfels >= 0 else surfels
    if a.only:
        alias = {"match": "hamming_match", "plane": "plane_prestage", "surfel": "surfel_fuse", "track": "search_by_projection"}
        want = {alias.get(x, x) for x in a.only.split(",") if x}
        a.stages = [s for s in a.stages if s in want]
    if "surfel_fuse" not in a.stages:
        a.surfels = 0
    return a


NO_GATHER = os.environ.get("MSL_BENCH_NO_GATHER") == "1"  # diagnostic (N > 1): ranks run free, no count-table collective; not a bench value


def config_of(a):
    """the SAME dict in both arms (the driver compares them); descriptive extras go to `config_detail`"""
    name = a.workload + ("_DIAGNOSTIC" if a.diagnostic or NO_GATHER else "")
    return {"workload": name, "frame": "%dx%d" % (a.W, a.H), "batch_per_gpu": a.batch, "surfels_per_gpu": a.surfels,
            "stages": list(a.stages)}


def camera(a):
    s = a.W / 640.0  # BASELINE.json config 5: "K scaled x2"
    return (525.0 * s, 525.0 * s, 319.5 * s + (s - 1) * 0.5, 239.5 * s + (s - 1) * 0.5)


def make_inputs(rank, batch, n_surfels, W=640, H=480, K=(525.0, 525.0, 319.5, 239.5)):
    """Seeded synthetic RGB-D batch + pose walk + surfel map."""
    from manhattanslam_b200 import synthetic as S
    seed0 = 1000 * (rank + 1)
    # 16 distinct frames cycled to the batch size keeps start-up short; every frame is still processed.
    # The depth stream is ONE scene (planes from seed0) seen with per-frame sensor noise and holes under the pose
    # walk: a keyframe stream into a local map.  (Independent scenes per frame would kill every in-view surfel in
    # the first pass and leave the fuse step with nothing to update.)
    uniq = min(batch, 16)
    gray_u = [S.gray_frame(seed0 + i, W, H) for i in range(uniq)]
    dd = [S.depth_frame(seed0 + i, W, H, K, scene=seed0) for i in range(uniq)]
    depth_u = [d[1] for d in dd]
    gray = np.stack([gray_u[i % uniq] for i in range(batch)])
    depth = np.stack([depth_u[i % uniq] for i in range(batch)])
    make_inputs.depth16 = np.stack([dd[i % uniq][0] for i in range(batch)])
    mem = np.stack([S.membership(seed0 + i, W, H) for i in range(batch)])
    poses = S.pose_walk(seed0, batch)
    # ~11 % of a fresh synthetic map leaves in the first few frames (unstable-drop rule :181-184 + the 5 % depth
    # outliers); the map is generated that much larger so that the steady state the timed region sees is n_surfels
    if n_surfels > 0:
        surfels = S.surfel_map(seed0, int(round(n_surfels / MAP_STEADY_FRACTION)), depth[0], poses[0], K, ref_index=100, w=W, h=H)
    else:
        from manhattanslam_b200.surfel import SURFEL_DTYPE
        surfels = np.zeros(0, SURFEL_DTYPE)
    return gray, depth, mem, poses, surfels


# ------------------------------------------------------------------------------------ CPU arm
_REF = {}


def _ref_lib(name):
    """True if oracle/_ref/<name> (the reference's own source compiled unmodified where /root/reference exists; the prebuilt
    library travels to the GPU box) is present AND loads"""
    if name not in _REF:
        try:
            import ctypes
            from oracle import binding as ob
            so = ob.build_ref(name=name)
            _REF[name] = so is not None and ctypes.CDLL(so) is not None
        except Exception:  # noqa: BLE001 -- fall back to the oracle port
            _REF[name] = False
    return _REF[name]


def ref_parts(stages):
    """which legs of the CPU arm run the reference's own source (oracle/_ref) and which the oracle port"""
    parts = {}
    if "orb" in stages:
        parts["orb"] = "reference src/ORBextractor.cc" if _ref_lib("liborb_ref.so") else "oracle port"
    if "hamming_match" in stages:
        # the reference has no brute-force matcher (DescriptorDistance, src/ORBmatcher.cc:835-849, is called from the window
        # searches); the all-pairs best-2 loop around that bit trick is the oracle's
        parts["hamming_match"] = "oracle port (all-pairs loop over the reference's DescriptorDistance bit trick)"
    if "search_by_projection" in stages:
        parts["search_by_projection"] = ("reference src/ORBmatcher.cc (SearchByProjection) on Last-frame map points built as "
                                         "Tracking::UpdateLastFrame / Frame::UnprojectStereo do (numpy restatement)"
                                         if _ref_lib("libmatch_ref.so") else "oracle port")
    if "plane_prestage" in stages:
        parts["plane_prestage"] = ("reference src/PlaneExtractor.cpp + include/peac (readDepthImage, PlaneSeg, initGraph)"
                                   if _ref_lib("libplane_ref.so") else "oracle port")
    if "plane_detect" in stages:
        parts["plane_detect"] = ("reference src/PlaneExtractor.cpp + include/peac (readDepthImage + runPlaneDetection), its "
                                 "membershipImg feeds SurfelFusion" if _ref_lib("libplane_ref.so") else "oracle port")
    if "surfel_fuse" in stages:
        parts["surfel_fuse"] = ("reference src/SurfelFusion.cpp (its ten std::threads, map resident in the library)"
                                if _ref_lib("libsurfel_ref_threads.so") else "oracle port (10 scan threads)")
    return parts


def baseline_kind(stages):
    """"reference" when the stage that dominates the CPU time of the workload is the reference's own source from
    oracle/_ref; "port" when it is the oracle port"""
    parts = ref_parts(stages)
    for s in ("surfel_fuse", "orb", "plane_detect", "search_by_projection", "plane_prestage", "hamming_match"):  # by CPU cost
        if s in parts:
            return "reference" if parts[s].startswith("reference") else "port"
    return "port"


def last_frame_side(kps, desc, xy_un, kdepth, Tcw, K, th_depth):
    """The Last-frame side of Tracking::TrackWithMotionModel as the reference builds it: Tracking::UpdateLastFrame
    (src/Tracking.cc:1052-1104) gives the keypoints with depth > 0 -- in (depth, index) order, all up to mThDepth, at least the
    100 closest -- "visual odometry" MapPoints at Frame::UnprojectStereo (src/Frame.cc:515-526): x3Dw = mRwc * x3Dc + mOw in
    cv::Mat arithmetic (CV_32F row sums a0*b0 + a1*b1 + a2*b2 in float, then (float)(t0 + c) in double; mOw = -mRcw.t() * mtcw
    with double accumulation).  Vectorised numpy; tests/test_track_batch_gpu.py checks it against the oracle's cv2-pinned
    primitives point by point."""
    f32 = np.float32
    fx, fy, cx, cy = (f32(v) for v in K)
    invfx, invfy = f32(1.0) / fx, f32(1.0) / fy
    Rcw, tcw = Tcw[:3, :3].astype(f32), Tcw[:3, 3].astype(f32)
    Ow = np.array([f32(sum((-np.float64(Rcw[k, r])) * np.float64(tcw[k]) for k in range(3))) for r in range(3)], f32)
    n = len(kps)
    z = np.asarray(kdepth, f32)
    pos = np.flatnonzero(z > 0)
    order = pos[np.lexsort((pos, z[pos]))]  # std::sort of (z, index) pairs
    zs = z[order]
    stop = np.flatnonzero((zs > f32(th_depth)) & (np.arange(1, len(order) + 1) > 100))
    sel = order[:(stop[0] + 1) if len(stop) else len(order)]
    has = np.zeros(n, np.uint8)
    has[sel] = 1
    u, v, zz = xy_un[sel, 0].astype(f32), xy_un[sel, 1].astype(f32), z[sel]
    x3 = np.stack([(u - cx) * zz * invfx, (v - cy) * zz * invfy, zz], 1).astype(f32)
    Rwc = np.ascontiguousarray(Rcw.T)
    world = np.zeros((n, 3), f32)
    for r in range(3):
        t0 = (Rwc[r, 0] * x3[:, 0] + Rwc[r, 1] * x3[:, 1]) + Rwc[r, 2] * x3[:, 2]  # float32 at every step
        world[sel, r] = (t0.astype(np.float64) + np.float64(Ow[r])).astype(f32)
    return {"has_mp": has, "outlier": np.zeros(n, np.uint8), "mp_obs": np.zeros(n, np.uint8), "mp_world": world,
            "mp_desc": desc, "octave": kps["octave"].astype(np.int32), "angle": kps["angle"].astype(np.float32)}


class CpuFrontend:
    """The reference's CPU path on the host cores.  ORB, the plane pre-stage and the Hamming match run frame-parallel over all
    host threads (the reference runs Frame::ExtractORB and Frame::ExtractPlanes on per-frame std::threads, src/Frame.cc:100-104);
    SurfelFusion runs frame by frame on a map that stays inside the library between steps like Map::mvLocalSurfels (created
    and filled ONCE, outside any timed region, exactly like the GPU arm's upload)."""

    def __init__(self, a, inputs, threads):
        from oracle import binding as ob
        self.ob, self.a, self.threads = ob, a, threads
        self.gray, self.depth, self.mem, self.poses, surfels = inputs
        self.depth16 = make_inputs.depth16
        self.K = camera(a)
        self.parts = ref_parts(a.stages)
        from manhattanslam_b200.matcher import frame_geom
        self.geom = frame_geom(a.W, a.H, *self.K, bf=MBF)
        self.Tcw = np.stack([np.linalg.inv(p.astype(np.float64)) for p in self.poses]).astype(np.float32)
        self.tl = threading.local()
        self.mem_det = np.full_like(self.mem, -1) if "plane_detect" in a.stages else None
        self.ref = 100
        self.rs = self.so = self.local = None
        if "surfel_fuse" in a.stages:
            if self.parts["surfel_fuse"].startswith("reference"):
                self.rs = ob.RefSurfelFusion(a.W, a.H, *self.K, real_threads=True)
                self.rs.set_map(surfels)
            else:
                self.so = ob.SurfelOracle(a.W, a.H, *self.K)
                self.local = surfels.copy()

    def _orb(self, i):
        ob = self.ob
        if not hasattr(self.tl, "o"):
            self.tl.o = ob.RefOrbExtractor(arena=False) if self.parts.get("orb", "").startswith("reference") else ob.OrbOracle()
        k, d = self.tl.o(self.gray[i])
        self.descs[i], self.kps[i] = d, k
        return len(k)

    def _plane(self, i):
        ob = self.ob
        if self.parts["plane_prestage"].startswith("reference"):
            return ob.ref_plane_timed(self.depth16[i], self.K, 1.0 / 5000.0, full=False)
        ob.plane_prestage(self.depth16[i], self.K)
        return 1

    def _detect(self, i):  # readDepthImage + runPlaneDetection; the membership image goes to the surfel stage
        ob = self.ob
        if self.parts["plane_detect"].startswith("reference"):
            return ob.ref_plane_timed(self.depth16[i], self.K, 1.0 / 5000.0, full=True, membership=self.mem_det[i])
        m, pl = ob.plane_detect(self.depth16[i], self.K, depth_map_factor=1.0 / 5000.0)
        self.mem_det[i][...] = m
        return len(pl["N"])

    def _match(self, i):  # brute-force Hamming best-2 of frame i against frame i+1
        return int(self.ob.hamming_best2(self.descs[i], self.descs[i + 1])[1].sum())

    def _track(self, i):
        """ORBmatcher::SearchByProjection(frame i+1, frame i, 15) as Tracking::TrackWithMotionModel calls it"""
        ob = self.ob
        kl, kc = self.kps[i], self.kps[i + 1]
        xy_l, xy_c = np.stack([kl["x"], kl["y"]], 1), np.stack([kc["x"], kc["y"]], 1)
        _, kd_l = ob.stereo_from_rgbd(xy_l, xy_l, self.depth[i], MBF)
        ur_c, _ = ob.stereo_from_rgbd(xy_c, xy_c, self.depth[i + 1], MBF)
        last = last_frame_side(kl, self.descs[i], xy_l, kd_l, self.Tcw[i], self.K, MBF * TH_DEPTH_FACTOR / self.K[0])
        cur = {"xy": xy_c, "octave": kc["octave"].astype(np.int32), "angle": kc["angle"].astype(np.float32), "uright": ur_c,
               "desc": self.descs[i + 1], "occupied": np.zeros(len(kc), np.uint8)}
        return ob.search_by_projection_frame(self.geom, self.Tcw[i + 1], self.Tcw[i], TH_PROJ, True, last, cur)[0]

    def step(self, frames):
        """one bounded step over the first `frames` frames of the batch -> seconds"""
        from concurrent.futures import ThreadPoolExecutor
        st = self.a.stages
        frames = min(frames, len(self.gray))
        self.descs, self.kps = [None] * frames, [None] * frames
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.threads) as ex:
            if "orb" in st:
                assert sum(ex.map(self._orb, range(frames))) > 0
            if "plane_prestage" in st:
                list(ex.map(self._plane, range(frames)))
            if "plane_detect" in st:
                list(ex.map(self._detect, range(frames)))
            if "hamming_match" in st and "orb" in st:
                list(ex.map(self._match, range(frames - 1)))
            if "search_by_projection" in st and "orb" in st:
                if self.parts["search_by_projection"].startswith("reference"):
                    with self.ob.reference_matcher():
                        list(ex.map(self._track, range(frames - 1)))
                else:
                    list(ex.map(self._track, range(frames - 1)))
        if "surfel_fuse" in st:
            mem = self.mem_det if self.mem_det is not None else self.mem
            for i in range(frames):
                if self.rs is not None:
                    self.rs.fuse_resident(self.ref + i, self.gray[i], self.depth[i], mem[i], self.poses[i])
                    self.rs.compact_resident()
                else:
                    new = self.so.fuse(self.ref + i, self.gray[i], self.depth[i], mem[i], self.poses[i], self.local,
                                       threads=min(10, self.threads))
                    self.local = self.ob.surfel_compact(self.local, new)
            self.ref += frames
        return time.perf_counter() - t0

    def sample_text(self, frames, seconds=None):
        legs = "; ".join("%s = %s" % (k, v) for k, v in self.parts.items())
        t = "" if seconds is None else " (%.1f s)" % seconds
        return "%d of %d frames per step%s on %d host threads, frame-parallel except SurfelFusion (frame by frame): %s" % (
            frames, self.a.batch, t, self.threads, legs)


def cpu_stage_breakdown(a, gray, depth, mem, poses, surfels, frames=2):
    """SURVEY.md section 8(d) 'CPU baseline timing': where the CPU path spends its time -- each stage of the oracle port on ONE
    thread (ms per frame; the reference runs ORB on one thread per frame), SurfelFusion with its 10 scan threads, and the
    OpenCV primitives ORB is made of timed through cv2 (SIMD, the 'optimised OpenCV' lower bound for the oracle's scalar
    FAST / resize / blur).  A few frames only; reported beside the baseline, never part of it."""
    from oracle import binding as ob
    out = {}
    K = camera(a)
    W, H = a.W, a.H
    try:
        def ms(fn, n=frames):
            t0 = time.perf_counter()
            for i in range(n):
                fn(i)
            return 1e3 * (time.perf_counter() - t0) / n
        o = ob.OrbOracle()
        descs = {}

        def orb(i):
            descs[i] = o(gray[i])[1]
        out["orb_1_thread_ms_per_frame"] = ms(orb)
        out["plane_prestage_1_thread_ms_per_frame"] = ms(lambda i: ob.plane_prestage(make_inputs.depth16[i], K))
        out["plane_detect_1_thread_ms_per_frame"] = ms(lambda i: ob.plane_detect(make_inputs.depth16[i], K, depth_map_factor=1.0))
        out["hamming_1000x1000_1_thread_ms"] = ms(lambda i: ob.hamming_best2(descs[0], descs[1]), 1)
        if len(surfels):
            so, local = ob.SurfelOracle(W, H, *K), surfels.copy()
            out["surfel_fuse_10_threads_ms_per_frame"] = ms(lambda i: so.fuse(100 + i, gray[i], depth[i], mem[i], poses[i], local,
                                                                              threads=min(10, os.cpu_count() or 1)))
            try:  # the reference's OWN src/SurfelFusion.cpp (oracle/_ref, compiled unmodified) with its own ten std::threads
                rs = ob.RefSurfelFusion(W, H, *K, real_threads=True)
                rs.set_map(surfels)  # the map stays inside the library, like Map::mvLocalSurfels: the fuse alone is timed
                out["surfel_fuse_reference_source_10_threads_ms_per_frame"] = ms(
                    lambda i: rs.fuse_resident(100 + i, gray[i], depth[i], mem[i], poses[i]))
            except Exception as e:  # noqa: BLE001 -- the prebuilt library did not travel
                out["surfel_fuse_reference_source_10_threads_ms_per_frame"] = "unavailable: %s" % e
    except Exception as e:  # noqa: BLE001 -- diagnostics only
        out["error"] = "%s: %s" % (type(e).__name__, e)
    try:
        import cv2
        cv2.setNumThreads(1)
        g = gray[0]
        det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FastFeatureDetector_TYPE_9_16)

        def best(fn, n=5):
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return 1e3 * min(ts)

        def pyramid():
            lv = g
            for l in range(1, 8):
                s = 1.0 / (1.2 ** l)
                lv = cv2.resize(lv, (int(round(W * s)), int(round(H * s))), interpolation=cv2.INTER_LINEAR)
        out["cv2_1_thread_ms"] = {"version": cv2.__version__, "pyramid_7_resizes": best(pyramid),
                                  "fast20_nms_level0_whole_image": best(lambda: det.detect(g)),
                                  "gaussian_blur_7x7_level0": best(lambda: cv2.GaussianBlur(g, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))}
    except Exception as e:  # noqa: BLE001
        out["cv2_1_thread_ms"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return out


def run_reference(a, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    inputs = make_inputs(0, a.batch, a.surfels, a.W, a.H, camera(a))
    cpu = CpuFrontend(a, inputs, threads)  # map created and resident before anything is timed
    # bounded sample: one short untimed pass estimates the host's speed, then the frames per step are chosen so that the
    # K timed steps together take about two and a half minutes (never more than --cpu-frames / the batch, never fewer than 4)
    probe = min(4, a.batch)
    fps_est = probe / cpu.step(probe)
    frames = max(min(4, a.batch), min(a.cpu_frames, a.batch, int(fps_est * 150.0 / max(a.steps, 1))))
    for _ in range(min(a.warmup, 1)):
        cpu.step(frames)
    ts = [cpu.step(frames) for _ in range(a.steps)]
    ms = 1e3 * sum(ts) / len(ts)
    value = frames / (ms / 1e3)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8+f32", "data": "synthetic", "config": config_of(a),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": baseline_kind(a.stages),
                            "sample": cpu.sample_text(frames), "legs": cpu.parts},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; started before the warm-up (nvidia-smi needs a few
    hundred ms to come up) and filtered to the wall-clock window of the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        self.p = None
        self.t0 = self.t1 = None
        if index is None:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", str(period_ms)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        import datetime
        sm, allsm, mx, reasons = [], [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk = float(f[1])
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            allsm.append(clk)
            if self.t0 is not None and (ts < self.t0 - 0.02 or ts > self.t1 + 0.02):
                continue
            sm.append(clk)
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        use = sm if sm else allsm[-5:]
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def widened_ops(msl, reps=20):
    """SURVEY.md section 8(f) rows built beyond the step (the vocabulary-node searches and the Fuse search of ORBmatcher):
    wall time of one call through the host C ABI (H2D of the flat arrays, kernel, D2H, sync) next to the CPU oracle on the
    same inputs, and whether the two agree.  Reported beside the headline numbers, never part of them; any failure is
    reported as text instead of numbers."""
    try:
        from manhattanslam_b200 import synthetic as S
        from oracle import binding as ob
        lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
        m = msl.ORBmatcher(nnratio=0.7)

        def timed(fn, n):
            fn()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                out = fn()
                ts.append(time.perf_counter() - t0)
            return 1e6 * float(np.median(ts)), out

        res = {}

        def packed(d):  # DBoW2 feature vector -> CSR arrays once, outside the timed calls (both arms take the packed form)
            d = dict(d)
            d["featvec"] = msl.ORBmatcher._csr(d["featvec"])
            return d

        kf, f = (packed(x) for x in S.bow_scene(1))
        g_us, (n_g, fm_g) = timed(lambda: m.SearchByBoW(kf, f), reps)
        c_us, (n_c, fm_c) = timed(lambda: ob.search_by_bow(0.7, True, kf, f), 3)
        res["SearchByBoW_1000x1000"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                        "equal": bool(n_g == n_c and np.array_equal(fm_g, fm_c))}
        kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(1)
        kf1, kf2 = packed(kf1), packed(kf2)
        g_us, (n_g, m_g) = timed(lambda: m.SearchForTriangulation(kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls), reps)
        c_us, (n_c, m_c) = timed(lambda: ob.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, kf1, kf2), 3)
        res["SearchForTriangulation_900x900"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                                 "equal": bool(n_g == n_c and np.array_equal(m_g, m_c))}
        mps, kfs, Tcw, ils = S.fuse_scene(1)
        geom = msl.frame_geom()
        g_us, (n_g, bi_g, bd_g) = timed(lambda: m.Fuse(geom, Tcw, mps, kfs, ils, th=3.0, log_scale_factor=lsf), reps)
        c_us, (n_c, bi_c, bd_c) = timed(lambda: ob.fuse_search(geom, Tcw, 3.0, lsf, ils, mps, kfs), 3)
        res["Fuse_1200x1000"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nfused": int(n_g),
                                 "equal": bool(n_g == n_c and np.array_equal(bi_g, bi_c) and np.array_equal(bd_g, bd_c))}
        sets = S.observation_sets(1)
        g_us, (bi_g, bm_g) = timed(lambda: m.ComputeDistinctiveDescriptors(sets), reps)
        c_us, (bi_c, bm_c) = timed(lambda: ob.distinctive_descriptors(sets), 3)
        res["ComputeDistinctiveDescriptors_400_points"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us,
                                                           "equal": bool(np.array_equal(bi_g, bi_c) and np.array_equal(bm_g, bm_c))}
        cur, last, mps2, Tc, Tl = S.match_scene(1)
        m2 = msl.ORBmatcher(nnratio=0.8)
        g_us, (n_g, cm_g) = timed(lambda: m2.SearchByProjectionFrame(geom, Tc, Tl, 7.0, last, cur), reps)
        c_us, (n_c, cm_c) = timed(lambda: ob.search_by_projection_frame(geom, Tc, Tl, 7.0, True, last, cur), 3)
        res["SearchByProjection_frame_900x1000"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                                    "equal": bool(n_g == n_c and np.array_equal(cm_g, cm_c))}
        g_us, (n_g, cm_g) = timed(lambda: m2.SearchByProjectionPoints(geom, 3.0, mps2, cur), reps)
        c_us, (n_c, cm_c) = timed(lambda: ob.search_by_projection_points(geom, 3.0, 0.8, mps2, cur), 3)
        res["SearchByProjection_points_900x1000"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                                     "equal": bool(n_g == n_c and np.array_equal(cm_g, cm_c))}
        cur3, kf3, Tc3 = S.reloc_scene(1)
        g_us, (n_g, cm_g) = timed(lambda: m2.SearchByProjectionKeyFrame(geom, Tc3, 15.0, 100, kf3, cur3, lsf), reps)
        c_us, (n_c, cm_c) = timed(lambda: ob.search_by_projection_keyframe(geom, Tc3, 15.0, 100, True, lsf, kf3, cur3), 3)
        res["SearchByProjection_keyframe_reloc"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                                    "equal": bool(n_g == n_c and np.array_equal(cm_g, cm_c))}
        # deferred batches (msl_matcher_batch_begin / _end): eight calls recorded, one upload, one CTA per call -- the form the
        # reference's loops over candidate / neighbour keyframes take (src/Tracking.cc:1930-1950, src/LocalMapping.cc:330-351,
        # :540-570); per-item time next to the single-thread oracle on the same eight inputs
        m.set_timing(True)
        for NB in (8, 32):
            bows = [tuple(packed(x) for x in S.bow_scene(1 + k)) for k in range(NB)]
            tris = [S.triangulation_scene(1 + k) for k in range(NB)]
            tris = [(packed(t_[0]), packed(t_[1])) + tuple(t_[2:]) for t_ in tris]
            fus = [S.fuse_scene(1 + k) for k in range(NB)]

            def batch_bow():
                with m.batch():
                    r = [m.SearchByBoW(a_, b_) for a_, b_ in bows]
                dev_log.append(m.last_execution())
                return [x.get() for x in r]

            def batch_tri():
                with m.batch():
                    r = [m.SearchForTriangulation(*t_) for t_ in tris]
                dev_log.append(m.last_execution())
                return [x.get() for x in r]

            def batch_fuse():
                with m.batch():
                    r = [m.Fuse(geom, t_[2], t_[0], t_[1], t_[3], th=3.0, log_scale_factor=lsf) for t_ in fus]
                dev_log.append(m.last_execution())
                return [x.get() for x in r]

            dev_log = []

            def entry(g_us, c_us, equal):
                dev_us = 1e3 * float(np.median([d_[0] / max(d_[1], 1) for d_ in dev_log]))  # per item, median over the repetitions
                del dev_log[:]
                return {"gpu_us_per_item": g_us / NB, "gpu_device_us_per_item": dev_us,
                        "cpu_oracle_us_per_item": c_us / NB, "speedup": c_us / g_us,
                        "speedup_device": c_us / NB / dev_us, "equal": bool(equal)}

            g_us, out_g = timed(batch_bow, reps if NB == 8 else 5)
            c_us, out_c = timed(lambda: [ob.search_by_bow(0.7, True, a_, b_) for a_, b_ in bows], 2)
            res["SearchByBoW_batch%d" % NB] = entry(g_us, c_us, all(x[0] == y[0] and np.array_equal(x[1], y[1]) for x, y in zip(out_g, out_c)))
            g_us, out_g = timed(batch_tri, reps if NB == 8 else 5)
            c_us, out_c = timed(lambda: [ob.search_for_triangulation(t_[2], t_[3], t_[4], t_[5], False, True, t_[6], t_[7], t_[0], t_[1]) for t_ in tris], 2)
            res["SearchForTriangulation_batch%d" % NB] = entry(g_us, c_us, all(x[0] == y[0] and np.array_equal(x[1], y[1]) for x, y in zip(out_g, out_c)))
            g_us, out_g = timed(batch_fuse, reps if NB == 8 else 5)
            c_us, out_c = timed(lambda: [ob.fuse_search(geom, t_[2], 3.0, lsf, t_[3], t_[0], t_[1]) for t_ in fus], 2)
            res["Fuse_batch%d" % NB] = entry(g_us, c_us, all(x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) for x, y in zip(out_g, out_c)))
        res["note"] = ("one call through the host C ABI incl. the Python mirror's array packing, H2D, kernel, D2H and sync; "
                       "*_batchN: N calls recorded between msl_matcher_batch_begin / _end (one upload, one CTA per call), per item -- "
                       "gpu_us = wall time incl. N ctypes calls of 20-30 arguments each (the CPU arm pays the same per item), "
                       "gpu_device_us = upload + kernels + download on the stream (CUDA events); cpu_oracle = the oracle "
                       "restatement, single thread; not part of the step")
        m.close()
        return res
    except Exception as e:  # noqa: BLE001 -- diagnostics only: never take the bench line down
        return {"error": "%s: %s" % (type(e).__name__, e)}


def widened_peac(msl, frames=64):
    """readDepthImage + the whole peac fitter (msl_plane_detect: pre-stage, ahCluster,
    refineDetails) for a batch of depth frames through the host C ABI, next to the CPU oracle (single thread) on the same
    frames, and whether membership images and planes agree."""
    try:
        from manhattanslam_b200 import synthetic as S
        from oracle import binding as ob
        d = np.stack([S.depth_frame(100 + b)[0] for b in range(frames)])
        pd = msl.PlaneDetection(max_batch=frames)
        pd.detect(d, depthMapFactor=1.0)
        t0 = time.perf_counter()
        mem, planes = pd.detect(d, depthMapFactor=1.0)
        g_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        nchk = min(frames, 8)  # the single-thread oracle on a sample of the frames, scaled to the batch
        ref = [ob.plane_detect(d[b], depth_map_factor=1.0) for b in range(nchk)]
        c_s = (time.perf_counter() - t0) * frames / nchk
        equal = all(np.array_equal(mem[b], ref[b][0]) and np.array_equal(planes[b]["N"], ref[b][1]["N"]) and
                    planes[b]["normal"].tobytes() == ref[b][1]["normal"].tobytes() for b in range(nchk))
        return {"plane_detect_640x480": {"frames": frames, "gpu_call_ms_per_batch": 1e3 * g_s, "cpu_oracle_ms_per_batch": 1e3 * c_s,
                                         "speedup_vs_one_cpu_thread": c_s / g_s, "frames_checked": nchk,
                                         "planes_per_frame": [len(p) for p in planes[:8]], "equal": bool(equal),
                                         "flood_serial": os.environ.get("MSL_PEAC_FLOOD_SERIAL", "0"),
                                         "note": "host API incl. H2D / D2H; one CTA per frame; the region grow runs level by level "
                                                 "(MSL_PEAC_FLOOD_SERIAL=1: as a FIFO on one thread)"}}
    except Exception as e:  # noqa: BLE001 -- diagnostics only
        return {"plane_detect_640x480": {"error": "%s: %s" % (type(e).__name__, e)}}


def widened_in_child(device, timeout_s=180):
    """widened_ops / widened_peac in child processes with a deadline: the diagnostics exercise kernels outside the timed step,
    and neither a device fault nor a hang there may cost the bench line."""
    res = _widened_child(device, "matcher", timeout_s)
    res.update(_widened_child(device, "peac", min(timeout_s, 90)))
    return res


def _widened_child(device, which, timeout_s):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    if not env["CUDA_VISIBLE_DEVICES"]:
        env["CUDA_VISIBLE_DEVICES"] = str(device)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--widened-only", which], env=env, capture_output=True,
                           text=True, timeout=timeout_s)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {which + "_error": "child exited %d: %s" % (r.returncode, (r.stderr or "").strip()[-300:])}
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return {which + "_error": "timed out after %d s" % timeout_s}
    except Exception as e:  # noqa: BLE001
        return {which + "_error": "%s: %s" % (type(e).__name__, e)}


def kernel_alg_bytes(W, H, B, n_map, upd, killed, kp_rows):
    """Algorithmic bytes per LAUNCH of every kernel of the step (DESIGN.md section 5 / SURVEY.md section 8d: what the kernel
    must read + write once; B frames per batched launch).  Levels of the 8-level, 1.2x pyramid are ceil-free approximations
    (round(W / 1.2^l)); the per-frame figures reproduce SURVEY's 640x480 numbers (926,546 R + 643,332 W for the pyramid ...)."""
    lv = [(int(round(W / 1.2 ** l)), int(round(H / 1.2 ** l))) for l in range(8)]
    px = [w * h for w, h in lv]
    pyr = sum(px)
    npx, nseeds = W * H, (W // 8) * (H // 8)
    w2, h2 = (W + 1) // 2, (H + 1) // 2
    nblk = (w2 // 10) * (h2 // 10)
    t = {
        "k_load_level0": B * 2.0 * npx,
        "k_resize": None,  # per level, below
        "k_fast_cells": B * (pyr + 0.1e6 * npx / 307200.0),
        "k_octree": B * 0.1e6 * npx / 307200.0,
        "k_blur": B * 2.0 * pyr,
        "k_describe": B * (kp_rows * (749 + 31 * 31) + kp_rows * 60.0),
        "k_hamming_best2": (B - 1) * (2 * kp_rows * 32.0 + kp_rows * 12.0),
        "k_keypoint_glue": B * kp_rows * (28.0 + 4.0 + 16.0),
        "k_track_last": B * kp_rows * (28.0 + 8.0 + 4.0 + 1.0 + 12.0 + 8.0),
        "k_search_batch": (B - 1) * kp_rows * (2 * 32.0 + 13.0 + 8.0 + 12.0 + 4.0),
        "k_plane_blocks": B * (npx * 1.0 + nblk * 72.0),  # u16 depth at even rows / columns, sector-granular = W*H bytes
        "k_plane_edges": B * nblk * (72.0 + 2.0),
        "k_peac_frame": B * (npx * 1.0 + w2 * h2 * (4.0 + 4.0)),  # depth in, membership + distance map out; the rest stays on chip
        "k_sp_init": B * (nseeds * (72.0 + 40.0) + nseeds * 8.0),
        "k_sp_pixels": B * (npx * (1 + 4 + 1) + npx * 8.0),           # gray + depth + membership(1/4 res, 4 B) R; target + index W
        "k_sp_pixels4": B * (npx * (1 + 4 + 1) + npx * 8.0),          # the same, four pixels per thread
        "k_sp_fix": B * npx * 8.0 * 0.25,                                # pending pixels only (a quarter, typically)
        "k_sp_seeds": B * (npx * (4 + 1 + 4) + nseeds * 112.0 * 2),      # index + gray + depth R once; seeds R + W
        "k_sp_seeds2": B * (npx * (4 + 1 + 4) + nseeds * 112.0 * 2),
        "k_sp_commit": B * nseeds * (72.0 * 2 + 40.0),
        "k_sp_norms": B * (npx * 4.0 + npx * 12.0),
        "k_sp_fit": B * (npx * (4 + 4 + 12) + nseeds * 72.0 * 2),
        "k_sp_fit2": B * (npx * (4 + 4 + 12) + nseeds * 72.0 * 2),
        "k_sp_records": B * nseeds * (72.0 + 80.0 + 4.0),
        # the fuse kernels: 24 B per surfel streamed + 4 B per killed surfel + per fused surfel 16 B read (q1) + 56 B written
        "k_fuse_one": n_map * 24.0 + killed * 4.0 + upd * 72.0,
        "k_fuse_stream": n_map * 24.0 + killed * 4.0 + upd * 72.0,
        "k_fuse_stream2": n_map * 24.0 + killed * 4.0 + upd * 72.0,
        "k_fuse_pipe": n_map * 24.0 + killed * 4.0 + upd * 72.0,
        "k_fuse_scan": n_map * 24.0 + killed * 4.0,
        "k_fuse_apply": upd * 100.0,
        "k_cmp_list": 4096.0, "k_cmp_apply": 4096.0, "k_frame_counts": B * 16.0,
    }
    for l in range(1, 8):
        t["k_resize_L%d" % l] = B * (px[l - 1] + px[l]) * 1.0
    t["k_resize"] = sum(t["k_resize_L%d" % l] for l in range(1, 8)) / 7.0  # average launch (7 launches per batch)
    t["k_resize4"] = t["k_resize"]  # the same bytes, four output pixels per thread (default)
    return t


BOUND_NOTE = {  # what each kernel is bound by in practice; the reported fraction is always of HBM peak
    "k_fast_cells": "integer ALU (FAST-9 segment test, ~200 ops per pixel); tiles staged by TMA, data from L2",
    "k_octree": "latency (a few thousand keys per level, sequential rounds)", "k_describe": "L2 gather + shuffle",
    "k_hamming_best2": "integer ALU (xor + popc)", "k_plane_blocks": "fp64 latency (per-block Jacobi)",
    "k_search_batch": "latency (one CTA per frame pair: sorted grid in shared memory, fixed point of the slot blocking)",
    "k_track_last": "latency (depth ranks by counting, one CTA per frame)",
    "k_peac_frame": "latency (one CTA per frame: ~600 dependent merge steps with an fp64 eigen-solve each, then a level-synchronous region grow)",
    "k_plane_edges": "latency", "k_sp_pixels": "fp64 issue (the reference's float/double cost)", "k_sp_fix": "latency",
    "k_sp_pixels4": "float<->double conversion rate (30 per pixel; F2F issues at 15.5 per clock per SM, profiles/r03e)",
    "k_sp_seeds2": "shared-memory latency (sequential float sums per seed)", "k_sp_fit2": "fp64 latency (sequential sums per seed)",
    "k_resize": "L2", "k_resize4": "latency (seven dependent launches per batch, two round trips per CTA)", "k_blur": "L2", "k_load_level0": "hbm", "k_sp_norms": "hbm", "k_sp_records": "hbm",
}


def short_kernel_name(name):
    n = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if n.startswith("void "):
        n = n[5:]
    return n.split("<")[0].split("(")[0].strip()


def per_kernel_pass(torch, step_dev, barrier, steps, alg, peak):
    """second, untimed-headline pass of `steps` steps under CUPTI activity records (torch.profiler: kernel name + GPU duration of
    every launch in the process, the library's own kernels included) -> roofline.kernels[]"""
    try:
        from torch.profiler import ProfilerActivity, profile
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                step_dev()
            barrier()
        if os.environ.get("MSL_TIMELINE"):  # diagnostic: the launches' start / end times per stream, for tools/timeline.py
            prof.export_chrome_trace(os.environ["MSL_TIMELINE"])
        agg = {}
        for ev in prof.key_averages():
            us = getattr(ev, "self_device_time_total", None)
            if us is None:
                us = getattr(ev, "self_cuda_time_total", 0.0)
            if not us or "memcpy" in ev.key.lower() or "memset" in ev.key.lower():
                continue
            k = short_kernel_name(ev.key)
            c = agg.setdefault(k, [0, 0.0])
            c[0] += int(ev.count)
            c[1] += float(us)
        rows = []
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            avg = us / n
            b = alg.get(k)
            gbs = b / (avg * 1e-6) / 1e9 if b else None
            rows.append({"kernel": k, "launches": n, "avg_us": avg, "total_us_per_step": us / steps, "alg_bytes_per_launch": b,
                         "achieved_gbs": gbs, "frac": gbs / peak if gbs else None, "bound": "hbm",
                         "in_practice": BOUND_NOTE.get(k)})
        return {"source": "CUPTI activity records (torch.profiler) over %d further steps, streams overlapping as in the timed "
                          "region; not the headline pass" % steps, "kernels": rows}
    except Exception as e:  # noqa: BLE001 -- diagnostics only
        return {"error": "%s: %s" % (type(e).__name__, e)}


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import manhattanslam_b200 as msl

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the single JSON line only: NCCL's version banner / debug output (printed whenever NCCL_DEBUG is
        # set, e.g. by the launcher's environment) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B, W, H, st = a.batch, a.W, a.H, a.stages
    K4 = camera(a)
    # Weak scaling: every rank streams the SAME synthetic scene, so that the work per GPU is exactly the N = 1 work (a scene
    # of its own per rank makes the ranks' steps differ by up to 4 % -- keypoints, superpixel iterations, surfels in view --
    # and the job, which advances in lock step through the count-table gather, runs at the slowest scene's pace:
    # profiles/r03l-r03n).  MSL_BENCH_RANK_SCENES=1 gives every rank its own scene again; MSL_BENCH_SCENE=<r> picks one.
    scene = rank if os.environ.get("MSL_BENCH_RANK_SCENES") == "1" else int(os.environ.get("MSL_BENCH_SCENE", "0"))
    gray, depth, mem, poses, surfels = make_inputs(scene, B, a.surfels, W, H, K4)
    depth16 = make_inputs.depth16

    do_orb, do_match, do_plane, do_surfel = "orb" in st, "hamming_match" in st and "orb" in st, "plane_prestage" in st, "surfel_fuse" in st
    do_track = "search_by_projection" in st and "orb" in st
    do_detect = "plane_detect" in st
    orb = msl.ORBextractor(width=W, height=H, max_batch=B, device=local_rank) if do_orb else None
    cap = orb.capacity if orb else 0
    sf = None
    if do_surfel:
        sf = msl.SurfelFusion(W, H, *K4, max_surfels=len(surfels) + 4 * B * (W // 8) * (H // 8), device=local_rank)
        sf.upload_map(surfels)
    matcher = msl.ORBmatcher(nnratio=0.9, max_queries=cap, max_train=cap, max_batch=B, device=local_rank) if (do_match or do_track) else None
    glue = msl.FrameGlue(W, H, max_batch=B, device=local_rank) if do_track else None
    geom = msl.frame_geom(W, H, *K4, bf=MBF)
    th_depth = MBF * TH_DEPTH_FACTOR / K4[0]  # Tracking::mThDepth (src/Tracking.cc:130)
    Tcw = np.stack([np.linalg.inv(p.astype(np.float64)) for p in poses]).astype(np.float32)
    plane = msl.PlaneDetection(W, H, max_batch=B, device=local_rank) if (do_plane or do_detect) else None
    nblk = plane.nblocks if plane else 0

    # pinned host staging (e2e leg) and device-resident inputs (kernel leg)
    h_gray = torch.from_numpy(gray).pin_memory()
    h_depth = torch.from_numpy(depth).pin_memory()
    h_mem = torch.from_numpy(mem).pin_memory()
    h_d16 = torch.from_numpy(depth16.view(np.int16)).pin_memory()
    d_gray, d_depth, d_mem, d_d16 = h_gray.to(dev), h_depth.to(dev), h_mem.to(dev), h_d16.to(dev)
    d_bi = torch.zeros((B, max(cap, 1)), dtype=torch.int32, device=dev)
    d_bd, d_sd = torch.zeros_like(d_bi), torch.zeros_like(d_bi)
    d_blocks = torch.zeros((B, max(nblk, 1), 72), dtype=torch.uint8, device=dev)
    d_seedm = torch.zeros((B, max(nblk, 1)), dtype=torch.uint8, device=dev)
    d_edges = torch.zeros((B, max(nblk, 1)), dtype=torch.uint8, device=dev)
    h_match = torch.zeros((3, B, max(cap, 1)), dtype=torch.int32).pin_memory()
    h_blocks = torch.zeros((B, max(nblk, 1), 72), dtype=torch.uint8).pin_memory()
    h_seedm = torch.zeros((2, B, max(nblk, 1)), dtype=torch.uint8).pin_memory()
    d_kps = torch.empty((B, max(cap, 1), 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, max(cap, 1), 32), dtype=torch.uint8, device=dev)
    h_kps = torch.empty((B, max(cap, 1), 28), dtype=torch.uint8).pin_memory()
    h_desc = torch.empty((B, max(cap, 1), 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros(B, dtype=torch.int32).pin_memory()
    d_xy = torch.zeros((B, max(cap, 1), 2), dtype=torch.float32, device=dev)
    d_ur, d_kd = torch.zeros((B, max(cap, 1)), dtype=torch.float32, device=dev), torch.zeros((B, max(cap, 1)), dtype=torch.float32, device=dev)
    d_cm = torch.zeros((B, max(cap, 1)), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(B, dtype=torch.int32, device=dev)
    h_cm = torch.zeros((B, max(cap, 1)), dtype=torch.int32).pin_memory()
    h_nm = torch.zeros(B, dtype=torch.int32).pin_memory()
    # the count table of SURVEY.md section 8(e), double-buffered so that the all-gather of step k (own stream) never holds up
    # the kernels of step k+1: per frame {keypoints, new surfels, updated surfels}
    d_kpc = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(2)]
    d_sfc = [torch.zeros((B, 2), dtype=torch.int32, device=dev) for _ in range(2)]
    d_table = [torch.zeros((B, 3), dtype=torch.int32, device=dev) for _ in range(2)]
    gathered = [torch.zeros((world * B, 3), dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
    # plane detection on the device: membership image (double-buffered: the detection of step k+1 may run while the superpixel
    # stage of step k still reads step k's), plane count and records per frame
    PCAP = 32
    d_memdet = [torch.full(((B, (H + 1) // 2, (W + 1) // 2)), -1, dtype=torch.int32, device=dev) for _ in range(2)] if do_detect else None
    d_pcount = torch.zeros(B, dtype=torch.int32, device=dev)
    d_precs = torch.zeros((B, PCAP, 64), dtype=torch.uint8, device=dev)
    ev_mem_free = [None, None]
    torch.cuda.synchronize()

    s_orb = torch.cuda.ExternalStream(orb.stream, device=dev) if orb else None
    s_sf = torch.cuda.ExternalStream(sf.stream, device=dev) if sf else None
    s_pl = torch.cuda.ExternalStream(plane.stream, device=dev) if plane else None
    s_sfin = torch.cuda.ExternalStream(sf.input_stream, device=dev) if (sf and do_detect) else None
    lib_streams = [s for s in (s_orb, s_sf, s_pl) if s is not None]
    s_comm = torch.cuda.Stream(device=dev)
    ev_gather = [None, None]
    state = {"ref": 100, "k": 0}

    resident = {"gray": d_gray.data_ptr(), "depth": d_depth.data_ptr(), "d16": d_d16.data_ptr(), "mem": d_mem.data_ptr()}

    def step_dev(src=None):
        """inputs resident in HBM (src: device pointers of another copy of the batch, the e2e leg's frame sets); ORB
        (+ matching), the plane pre-stage and the surfel stream run on their own CUDA streams and overlap; the count table
        goes to the other ranks on a fourth stream"""
        src = src or resident
        k = state["k"] & 1
        state["k"] += 1
        if ev_gather[k] is not None:  # the gather of two steps ago read this buffer set (long done)
            for s in lib_streams:
                s.wait_event(ev_gather[k])
        if do_orb:
            orb.extract_dev(src["gray"], W, W * H, B, d_kps.data_ptr(), d_desc.data_ptr(), d_kpc[k].data_ptr())
        if do_match:  # frame b vs frame b+1, chained on the ORB stream (no host sync between extraction and matching)
            matcher.hamming_best2_counts_dev(d_desc.data_ptr(), d_desc.data_ptr() + cap * 32, cap, d_kpc[k].data_ptr(),
                                             d_kpc[k].data_ptr() + 4, B - 1, d_bi.data_ptr(), d_bd.data_ptr(), d_sd.data_ptr(),
                                             stream=orb.stream)
        if do_track:  # frame glue + UpdateLastFrame + SearchByProjection(frame b+1, frame b), still on the ORB stream
            glue.keypoints_dev(d_kps.data_ptr(), cap, d_kpc[k].data_ptr(), B, K4, None, src["depth"], MBF, d_xy.data_ptr(),
                               d_ur.data_ptr(), d_kd.data_ptr(), stream=orb.stream)
            matcher.SearchByProjectionFrames_dev(geom, TH_PROJ, th_depth, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_kpc[k].data_ptr(),
                                                 B, d_xy.data_ptr(), d_ur.data_ptr(), d_kd.data_ptr(), Tcw, d_cm.data_ptr(),
                                                 d_nm.data_ptr(), stream=orb.stream)
        if do_plane:
            plane.prestage_dev(src["d16"], B, K4, 1.0 / 5000.0, None, d_blocks.data_ptr(), d_seedm.data_ptr(),
                               d_edges.data_ptr())
        mem_ptr = src["mem"]
        if do_detect:  # readDepthImage + runPlaneDetection for the batch; its membership image is the surfel stage's input
            if ev_mem_free[k] is not None:
                s_pl.wait_event(ev_mem_free[k])  # the superpixel stage of two steps ago has read this buffer
            plane.detect_dev(src["d16"], B, K4, 1.0 / 5000.0, d_memdet[k].data_ptr(), d_pcount.data_ptr(), d_precs.data_ptr(), PCAP)
            mem_ptr = d_memdet[k].data_ptr()
            if do_surfel:
                ev = torch.cuda.Event()
                ev.record(s_pl)
                s_sfin.wait_event(ev)
        if do_surfel:
            sf.set_count_table(d_sfc[k].data_ptr())
            sf.fuse_batch_dev(state["ref"], src["gray"], W, W * H, src["depth"], mem_ptr, poses, B, True)
            if do_detect:
                ev = torch.cuda.Event()
                ev.record(s_sfin)
                ev_mem_free[k] = ev
        state["ref"] += B
        if world > 1 and not NO_GATHER:  # the path's single collective: the per-frame count table to every rank, off the compute streams
            for s in lib_streams:
                s_comm.wait_stream(s)
            with torch.cuda.stream(s_comm):
                d_table[k][:, 0].copy_(d_kpc[k])
                d_table[k][:, 1:].copy_(d_sfc[k])
                dist.all_gather_into_tensor(gathered[k], d_table[k])
                ev = torch.cuda.Event()
                ev.record(s_comm)
                ev_gather[k] = ev

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(n):
        """n steps bracketed by events on the current stream that every library stream is fenced against -> ms"""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        ev0.record(cur)
        for s in lib_streams + [s_comm]:
            s.wait_event(ev0)
        th0 = time.perf_counter()
        for _ in range(n):
            step_dev()
        state["host_enqueue_ms"] = 1e3 * (time.perf_counter() - th0) / max(n, 1)  # host time to enqueue one step (no sync inside)
        for s in lib_streams + [s_comm]:
            e = torch.cuda.Event()
            e.record(s)
            cur.wait_event(e)
        ev1.record(cur)
        barrier()
        return ev0.elapsed_time(ev1)

    # MSL_BENCH_CLOCKS (diagnostic): "off" = no sampler, "rank0" = only rank 0 samples (its own GPU); default: every rank its GPU
    clk_mode = os.environ.get("MSL_BENCH_CLOCKS", "")
    clk = ClockSampler(None if clk_mode == "off" or (clk_mode == "rank0" and rank != 0) else local_rank,
                       int(os.environ.get("MSL_BENCH_CLOCKS_MS", "20")))
    for _ in range(a.warmup):
        step_dev()
    barrier()
    launches0 = msl.lib().msl_kernel_launch_count()
    # ---- the headline: K steps, no timing aid of any kind inside the region
    clk.begin()
    ms_total = timed_steps(a.steps)
    host_enqueue_ms = state.get("host_enqueue_ms")
    clk.end()
    # the same with an empty launch queue (one step after a full synchronisation): what the host really spends per step --
    # inside the timed loop the figure above includes waiting for room in the queue once the host is a few steps ahead
    barrier()
    th0 = time.perf_counter()
    step_dev()
    host_unloaded_ms = 1e3 * (time.perf_counter() - th0)
    barrier()
    clocks = clk.stop()
    launches = int(msl.lib().msl_kernel_launch_count() - launches0)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    mine_ms = ms_total / a.steps
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = world * B / (ms_step / 1e3)

    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass

    roofline, per_rank, n_map = None, None, 0
    if do_surfel:
        # ---- second pass, same K steps, with the library's CUDA events around the fuse kernel of every 8th frame (an event
        # record costs ~2.7 us of stream time, which is why the headline pass runs without): the kernel's time INSIDE a step
        sf.set_timing(1)
        timed_steps(a.steps)
        chain, chain_frames = sf.chain_times()
        sf.set_timing(0)
        # ---- the same kernel timed alone (one extra stream call after a full sync: its superpixel stage precedes its
        # chain and no other stream is busy), to separate the kernel's own efficiency from SM sharing in the timed region
        barrier()
        sf.set_timing(2)
        sf.fuse_batch_dev(state["ref"], d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), (d_memdet[0] if do_detect else d_mem).data_ptr(), poses, B, True)
        state["ref"] += B
        iso_chain, iso_frames = sf.chain_times()
        st1 = sf.read_stats()
        info = sf.launch_info()
        # ---- and alone with a full wave of CTAs (three per SM): inside a batch the library launches two per SM so that the
        # next batch's superpixel kernels find room beside the chain (msl_surfel_set_fuse_ctas_per_sm) -- better for the
        # step, 6 % worse for the kernel on its own
        barrier()
        sf.set_fuse_ctas_per_sm(3, 0)
        sf.fuse_batch_dev(state["ref"], d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), (d_memdet[0] if do_detect else d_mem).data_ptr(), poses, B, True)
        state["ref"] += B
        full_chain, full_frames = sf.chain_times()
        info_full = sf.launch_info()
        sf.set_fuse_ctas_per_sm(int(os.environ.get("MSL_STREAM_WAVE_BATCH", os.environ.get("MSL_STREAM_WAVE", "2"))), 0)
        sf.set_timing(0)
        n_map = st1[3]
        # stats accumulate per fuse_batch call: st1 holds the last call's totals over its B launches
        upd_per_launch = st1[1] / B
        del_per_launch = st1[2] / B
        try:
            with open(os.path.join(ROOT, "profiles", "fuse_traffic.json")) as f:
                ncu_traffic = json.load(f)
        except Exception:
            ncu_traffic = {}
        kname = {0: "k_fuse_scan+k_fuse_apply", 1: "k_fuse_one", 2: "k_fuse_stream", 3: "k_fuse_stream2", 4: "k_fuse_pipe"}[info["form"]]
        # algorithmic bytes per launch: 24 B per surfel streamed (the {px, py, pz, size} quad,
        # updateTimes, lastUpdate) + 4 B per killed surfel + per fused surfel 16 B read (the normal / weight quad) + 56 B
        # written = 72 B.  The two-kernel chain (MSL_FUSE_ONE=0) additionally moves 8 B of queue + 20 B re-read per fused surfel.
        alg_one = n_map * 24.0 + del_per_launch * 4.0 + upd_per_launch * (72.0 if info["kernels"] == 1 else 100.0)

        def entry(times, frames, iso_times, iso_n):
            key_ms = (times["scan"] + (times["apply"] if info["kernels"] == 2 else 0.0)) / max(frames, 1)
            iso = (iso_times["scan"] + (iso_times["apply"] if info["kernels"] == 2 else 0.0)) / max(iso_n, 1)
            ach = alg_one / (key_ms * 1e-3) / 1e9 if key_ms > 0 else None
            ach_iso = alg_one / (iso * 1e-3) / 1e9 if iso > 0 else None
            tr = ncu_traffic.get(kname, {}).get("dram_bytes_per_surfel")
            return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak if ach else None, "traffic": tr * n_map if tr else None, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg_one, "avg_launch_ms": key_ms, "launches_timed": frames,
                    "launches_per_step": B, "share_of_step": key_ms * B / mine_ms if mine_ms else None,
                    "timed_in": "a second pass of the same %d steps with the library's CUDA events on every 8th launch "
                                "(the headline pass carries no timing aid)" % a.steps,
                    "isolated": {"avg_launch_ms": iso, "achieved": ach_iso, "frac": ach_iso / peak if ach_iso else None,
                                 "note": "same kernel, same map, no other stream active"}}

        roofline = entry(chain, chain_frames, iso_chain, iso_frames)
        roofline["launch"] = info
        roofline["chain_us_per_frame"] = {k: 1e3 * v / max(chain_frames, 1) for k, v in chain.items()}
        roofline["isolated"]["chain_us_per_frame"] = {k: 1e3 * v / max(iso_frames, 1) for k, v in iso_chain.items()}
        roofline["isolated"]["grid"] = info["grid"]
        full_ms = (full_chain["scan"] + (full_chain["apply"] if info["kernels"] == 2 else 0.0)) / max(full_frames, 1)
        ach_full = alg_one / (full_ms * 1e-3) / 1e9 if full_ms > 0 else None
        roofline["isolated_full_wave"] = {"avg_launch_ms": full_ms, "achieved": ach_full, "frac": ach_full / peak if ach_full else None,
                                          "grid": info_full["grid"], "note": "same kernel alone with three CTAs per SM (the geometry "
                                          "of single-frame calls); the step's chain launches two per SM to leave room for the next "
                                          "batch's superpixel kernels"}
        roofline["fused_per_launch"] = upd_per_launch
        roofline["killed_per_launch"] = del_per_launch
        roofline["note"] = "timed-region launches share the SMs with the next batch's superpixel kernels (stream overlap)"
        if world > 1:  # per-rank step and kernel times: is a slow rank or a slow kernel behind the max-over-ranks time?
            mine = torch.tensor([mine_ms, roofline["avg_launch_ms"] * 1e3, host_enqueue_ms or 0.0], dtype=torch.float64, device=dev)
            allr = torch.zeros((world, 3), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allr, mine)
            per_rank = {"ms_per_step": [float(x) for x in allr[:, 0]], "fuse_kernel_us": [float(x) for x in allr[:, 1]],
                        "host_enqueue_ms_per_step": [float(x) for x in allr[:, 2]]}
    elif world > 1:
        mine = torch.tensor([mine_ms], dtype=torch.float64, device=dev)
        allr = torch.zeros((world, 1), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {"ms_per_step": [float(x) for x in allr[:, 0]]}

    extras = rank == 0 and not a.no_extras
    # ---- per-kernel table (third pass, CUPTI): every kernel of the step against the HBM roofline
    if extras:
        kp_rows = 1000.0
        alg = kernel_alg_bytes(W, H, B, n_map, roofline["fused_per_launch"] if roofline else 0.0,
                               roofline["killed_per_launch"] if roofline else 0.0, kp_rows)
        kt = per_kernel_pass(torch, step_dev, barrier, min(a.steps, 5), alg, peak)
        if roofline is None:  # a workload without the surfel stage: the kernel with the largest share is the headline entry
            rows = (kt.get("kernels") or [])
            top = rows[0] if rows else None
            roofline = {"kernel": top["kernel"] if top else None, "bound": "hbm", "achieved": top["achieved_gbs"] if top else None,
                        "peak": peak, "unit": "GB/s", "frac": top["frac"] if top else None, "traffic": None,
                        "peak_source": peak_src, "avg_launch_ms": top["avg_us"] / 1e3 if top else None,
                        "alg_bytes_per_launch": top["alg_bytes_per_launch"] if top else None,
                        "in_practice": top["in_practice"] if top else None}
        roofline["kernels"] = kt.get("kernels")
        roofline["kernels_source"] = kt.get("source", kt.get("error"))
        # the whole step against the HBM roofline: the algorithmic bytes of every launch of a step over the step's time.  The
        # dominant kernel's own `frac` divides by its launch duration INSIDE the step, where -- by design -- it shares every
        # SM with the other streams' kernels (two of its CTAs per SM, the rest of the SM theirs)
        steps_k = min(a.steps, 5)
        step_bytes = sum((k_["alg_bytes_per_launch"] or 0.0) * k_["launches"] / steps_k for k_ in (kt.get("kernels") or []))
        if step_bytes and ms_step:
            ach_step = step_bytes / (ms_step * 1e-3) / 1e9
            roofline["step"] = {"alg_bytes_per_step": step_bytes, "achieved": ach_step, "unit": "GB/s", "frac": ach_step / peak,
                                "note": "sum over the step's launches of their algorithmic bytes / ms_per_step"}

    # ---- parity replay: the last frames of the run again on the CPU oracle from the downloaded pre-state
    parity = None
    if extras and do_surfel and world == 1:
        try:
            from oracle import binding as ob
            NCHK = 4
            barrier()
            pre = sf.download_map()
            r0 = state["ref"]
            mem_chk = d_memdet[0].cpu().numpy() if do_detect else mem
            sf.fuse_batch_dev(r0, d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), (d_memdet[0] if do_detect else d_mem).data_ptr(), poses, NCHK, True)
            state["ref"] += NCHK
            post = sf.download_map()
            so = ob.SurfelOracle(W, H, *K4)
            buf = np.zeros(len(pre) + NCHK * (W // 8) * (H // 8), pre.dtype)
            buf[:len(pre)] = pre
            n = len(pre)
            L = ob.lib()
            for i in range(NCHK):
                new = np.ascontiguousarray(so.fuse(r0 + i, gray[i], depth[i], mem_chk[i], poses[i], buf[:n], threads=min(16, os.cpu_count() or 1)))
                n = L.orc_surfel_compact(ob._p(buf), n, ob._p(new), len(new))
            ok = n == len(post) and np.array_equal(buf[:n].view(np.uint8), post.view(np.uint8))
            parity = {"check": "ok" if ok else "MISMATCH", "frames": NCHK, "map_surfels": int(len(pre)),
                      "what": "the map after %d more frames of the stream, downloaded, against the CPU oracle replaying the same "
                              "frames from the downloaded pre-state: bit-identical records" % NCHK}
        except Exception as e:  # noqa: BLE001
            parity = {"check": "error: %s: %s" % (type(e).__name__, e)}

    # ---- e2e_dropin: SurfelFusion::fuseInitializeMap exactly as adapters/SurfelFusion_msl.cpp performs it when nothing else
    # of the reference is touched: Map::mvLocalSurfels stays authoritative on the host, so EVERY keyframe uploads the whole
    # map (pageable memory, like a std::vector), fuses, and downloads the surfels the call changed (dirty download); the
    # unchanged fuseMap tail then compacts the host vector (not part of the replaced method, not timed)
    dropin = None
    if extras and do_surfel and world == 1:
        try:
            from oracle import binding as ob
            ND = 4
            barrier()
            host = sf.download_map()
            buf = np.zeros(len(host) + ND * (W // 8) * (H // 8), host.dtype)
            buf[:len(host)] = host
            n = len(host)
            sf2 = msl.SurfelFusion(W, H, *K4, max_surfels=len(buf) + 4096, device=local_rank)
            L = ob.lib()
            t_sum, up_b, down_b = 0.0, 0, 0
            r0 = state["ref"]
            for i in range(ND + 1):
                view = buf[:n]
                t0 = time.perf_counter()
                sf2.upload_map(view)
                new, _ = sf2.fuseInitializeMap(r0 + i, gray[i], depth[i], mem[i], poses[i], compact=False)
                idx, rec = sf2.download_changed(r0 + i)
                view[idx] = rec
                dt = time.perf_counter() - t0
                if i > 0:  # the first call carries the allocations
                    t_sum += dt
                    up_b += view.nbytes + gray[i].nbytes + depth[i].nbytes + mem[i].nbytes
                    down_b += idx.nbytes + rec.nbytes + new.nbytes
                new = np.ascontiguousarray(new)
                n = L.orc_surfel_compact(ob._p(buf), n, ob._p(new), len(new))
            state["ref"] += ND + 1
            dropin = {"value": ND / t_sum, "unit": UNIT, "frames": ND, "ms_per_frame": 1e3 * t_sum / ND,
                      "h2d_bytes_per_frame": up_b // ND, "d2h_bytes_per_frame": down_b // ND, "map_surfels": int(n),
                      "what": "SurfelFusion::fuseInitializeMap as adapters/SurfelFusion_msl.cpp runs it without any other change to "
                              "the reference: whole map uploaded from pageable host memory every keyframe, dirty download of the "
                              "changed surfels, patch into the host vector; surfel stage only (the device-resident mode is `e2e`)"}
            sf2.close()
        except Exception as e:  # noqa: BLE001
            dropin = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- e2e: the public host API with pinned host buffers, H2D of the inputs + D2H of the results every step
    import ctypes as C
    from manhattanslam_b200._lib import check, ptr

    # The three host-API call sequences are issued from three host threads, as the reference does (Frame::ExtractORB
    # and Frame::ExtractPlanes run on per-frame std::threads, src/Frame.cc:100-104; SurfelMapping has its own thread,
    # src/System.cc:98-99).  Every handle owns its stream(s); ctypes releases the GIL for the duration of a call.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=3)
    Kf = np.asarray(K4, np.float32)
    if sf:
        sf.set_count_table(None)

    def e2e_orb_match():
        torch.cuda.set_device(local_rank)
        check(orb._L.msl_orb_extract(orb._h, C.c_void_p(h_gray.data_ptr()), C.c_int(W), C.c_size_t(W * H), C.c_int(B),
                                     C.c_void_p(h_kps.data_ptr()), C.c_void_p(h_desc.data_ptr()),
                                     C.c_void_p(h_counts.data_ptr())))
        # matching on the descriptors just produced (device-resident copy inside the ORB handle is not exposed, so the
        # host API path re-uploads them: that is what a host-side caller of the C ABI pays)
        if do_match:
            check(matcher._L.msl_hamming_best2(matcher._h, C.c_void_p(h_desc.data_ptr()), C.c_int(cap),
                                               C.c_void_p(h_desc.data_ptr() + cap * 32), C.c_int(cap), C.c_int(B - 1),
                                               C.c_void_p(h_match[0].data_ptr()), C.c_void_p(h_match[1].data_ptr()),
                                               C.c_void_p(h_match[2].data_ptr())))
        if do_track:
            check(matcher._L.msl_search_by_projection_frames(
                matcher._h, glue._h, ptr(geom), C.c_float(TH_PROJ), 1, C.c_float(th_depth), C.c_void_p(h_kps.data_ptr()),
                C.c_void_p(h_desc.data_ptr()), C.c_int(cap), C.c_void_p(h_counts.data_ptr()), C.c_int(B),
                C.c_void_p(h_depth.data_ptr()), C.c_int(W), C.c_int(H), ptr(Tcw), C.c_void_p(h_cm.data_ptr()),
                C.c_void_p(h_nm.data_ptr())))

    h_memdet = torch.zeros((B, (H + 1) // 2, (W + 1) // 2), dtype=torch.int32).pin_memory() if do_detect else None
    h_pcount = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_precs = torch.zeros((B, PCAP, 64), dtype=torch.uint8).pin_memory()

    def e2e_detect():
        torch.cuda.set_device(local_rank)
        check(plane._L.msl_plane_detect(plane._h, C.c_void_p(h_d16.data_ptr()), C.c_int(W), C.c_size_t(W * H), C.c_int(B), ptr(Kf),
                                        C.c_float(1.0 / 5000.0), C.c_void_p(h_memdet.data_ptr()), C.c_void_p(h_pcount.data_ptr()),
                                        C.c_void_p(h_precs.data_ptr()), C.c_int(PCAP)))

    def e2e_plane():
        torch.cuda.set_device(local_rank)
        check(plane._L.msl_plane_prestage(plane._h, C.c_void_p(h_d16.data_ptr()), C.c_int(W), C.c_size_t(W * H), C.c_int(B),
                                          ptr(Kf), C.c_float(1.0 / 5000.0), None, C.c_void_p(h_blocks.data_ptr()),
                                          C.c_void_p(h_seedm[0].data_ptr()), C.c_void_p(h_seedm[1].data_ptr())))

    def e2e_surfel(ref):
        torch.cuda.set_device(local_rank)
        stats = np.zeros(4, np.int64)
        if do_detect:
            e2e_detect()  # the membership image comes back to the host and goes in again with the frame, as in the reference
        check(sf._L.msl_surfel_fuse_batch(sf._h, ref, C.c_void_p(h_gray.data_ptr()), C.c_int(W),
                                          C.c_void_p(h_depth.data_ptr()), C.c_void_p((h_memdet if do_detect else h_mem).data_ptr()), ptr(poses),
                                          C.c_int(B), 1, ptr(stats)))
        return stats

    def step_e2e():
        futs = []
        if do_orb:
            futs.append(pool.submit(e2e_orb_match))
        if do_plane:
            futs.append(pool.submit(e2e_plane))
        if do_surfel:
            futs.append(pool.submit(e2e_surfel, state["ref"]))
        elif do_detect:
            futs.append(pool.submit(e2e_detect))
        state["ref"] += B
        for f in futs:
            f.result()

    for _ in range(min(a.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    pool.shutdown()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_calls_value = world * B * a.steps / float(t.item())

    # ---- e2e (headline): the sensor frames of a step -- gray (CV_8U) and depth (CV_16U), what System::TrackRGBD receives -- go up
    # ONCE from pinned host memory into a frame set every stage reads (msl_glue_upload_frames; the CV_32F depth of
    # Tracking::GrabImageRGBD is produced on the device), the stages run through their *_dev entry points on their own
    # streams, and every result comes back to pinned host memory.  Steps are pipelined the way a streaming front-end runs:
    # the upload of step k+2 and the launches of step k+1 are issued before the host waits for the results of step k (three
    # frame sets, two sets of result buffers); all K uploads and all K downloads lie inside the timed region, which ends when the last
    # step's results are on the host.
    glue_up = glue if glue is not None else msl.FrameGlue(W, H, max_batch=B, device=local_rank)
    wait_streams = [x for x in (orb.stream if orb else None, plane.stream if plane else None,
                                sf.input_stream if sf else None) if x]
    aux = h_mem if (do_surfel and not do_detect) else None

    def upload(slot):
        g_, d16_, dep_, aux_ = glue_up.upload_frames(slot, h_gray.data_ptr(), h_d16.data_ptr(), B, 1.0 / 5000.0,
                                                     aux.data_ptr() if aux is not None else None,
                                                     aux.numel() if aux is not None else 0)
        return {"gray": g_, "depth": dep_, "d16": d16_, "mem": aux_ or resident["mem"]}

    # results land in two sets of pinned host buffers: step i's launches (and its downloads, enqueued behind them on the
    # producing streams) are issued BEFORE step i-1's results are awaited, as a streaming consumer does -- otherwise the
    # host's wait at the end of a step would drain the device and the next step's superpixel stage could not overlap this
    # step's fuse chain, which is where the device-resident number comes from
    def pin_like(t_):
        return torch.empty_like(t_).pin_memory()

    hres = [{"kps": h_kps, "desc": h_desc, "counts": h_counts, "match": h_match, "cm": h_cm, "nm": h_nm, "blocks": h_blocks,
             "seedm": h_seedm, "memdet": h_memdet, "pcount": h_pcount, "precs": h_precs,
             "sfc": torch.zeros((B, 2), dtype=torch.int32).pin_memory()}]
    hres.append({k_: (pin_like(v_) if v_ is not None else None) for k_, v_ in hres[0].items()})
    done_ev = [None, None]

    NSETS = 3  # MSL_GLUE_FRAME_SETS

    def launch_frames(i, sets):
        slot = i % NSETS
        H_ = hres[i & 1]
        for st_ in wait_streams:
            glue_up.frames_wait(slot, st_)
        k = state["k"] & 1
        step_dev(sets[slot])
        evs = []
        if do_orb:
            with torch.cuda.stream(s_orb):
                H_["kps"].copy_(d_kps, non_blocking=True)
                H_["desc"].copy_(d_desc, non_blocking=True)
                H_["counts"].copy_(d_kpc[k], non_blocking=True)
                if do_match:
                    H_["match"][0].copy_(d_bi, non_blocking=True)
                    H_["match"][1].copy_(d_bd, non_blocking=True)
                    H_["match"][2].copy_(d_sd, non_blocking=True)
                if do_track:
                    H_["cm"].copy_(d_cm, non_blocking=True)
                    H_["nm"].copy_(d_nm, non_blocking=True)
                e_ = torch.cuda.Event()
                e_.record(s_orb)
                evs.append(e_)
        if plane is not None:
            with torch.cuda.stream(s_pl):
                if do_plane:
                    H_["blocks"].copy_(d_blocks, non_blocking=True)
                    H_["seedm"][0].copy_(d_seedm, non_blocking=True)
                    H_["seedm"][1].copy_(d_edges, non_blocking=True)
                if do_detect:
                    H_["memdet"].copy_(d_memdet[k], non_blocking=True)
                    H_["pcount"].copy_(d_pcount, non_blocking=True)
                    H_["precs"].copy_(d_precs, non_blocking=True)
                e_ = torch.cuda.Event()
                e_.record(s_pl)
                evs.append(e_)
        if do_surfel:  # the surfel stage's per-frame result: {new surfels, updated surfels} of every frame
            with torch.cuda.stream(s_sf):
                H_["sfc"].copy_(d_sfc[k], non_blocking=True)
                e_ = torch.cuda.Event()
                e_.record(s_sf)
                evs.append(e_)
        done_ev[i & 1] = evs

    def collect_frames(i):
        for e_ in done_ev[i & 1] or []:
            e_.synchronize()
        done_ev[i & 1] = None

    def run_frames(n_steps):
        # three frame sets: the frames of step i+2 go up while step i runs, so that step i+1's superpixel stage finds its
        # input on the device when step i's fuse chain starts and runs beside it, as in the device-resident loop (with two
        # sets the upload of step i+1 could only start once step i-1 was collected, 1.7 ms into step i's chain)
        sets = [None] * NSETS
        sets[0] = upload(0)
        if n_steps > 1:
            sets[1] = upload(1)
        for i in range(n_steps):
            launch_frames(i, sets)
            if i > 0:
                collect_frames(i - 1)  # step i-1 is complete on the host: its frame set may be overwritten
            if i + 2 < n_steps:
                sets[(i + 2) % NSETS] = upload((i + 2) % NSETS)  # the set of step i-1
        collect_frames(n_steps - 1)

    barrier()
    run_frames(min(a.warmup, 2) + 1)
    barrier()
    t0 = time.perf_counter()
    run_frames(a.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * a.steps / float(t.item())
    fh2d = h_gray.numel() + h_d16.numel() * 2 + (aux.numel() * 4 if aux is not None else 0) + (64 * B if do_surfel else 0) + (64 * B if do_track else 0)
    fd2h = 0
    if do_orb:
        fd2h += h_kps.numel() + h_desc.numel() + h_counts.numel() * 4
    if do_match:
        fd2h += h_match.numel() * 4
    if do_track:
        fd2h += h_cm.numel() * 4 + h_nm.numel() * 4
    if do_plane:
        fd2h += h_blocks.numel() + h_seedm.numel()
    if do_detect:
        fd2h += h_memdet.numel() * 4 + B * 4 + h_precs.numel()
    if do_surfel:
        fd2h += B * 8
    h2d = d2h = 0
    if do_orb:
        h2d += h_gray.numel()
        d2h += h_kps.numel() + h_desc.numel() + h_counts.numel() * 4
    if do_match:
        h2d += 2 * (B - 1) * cap * 32
        d2h += 3 * (B - 1) * cap * 4
    if do_track:
        h2d += B * cap * (28 + 32) + B * 4 + h_depth.numel() * 4 + B * 64
        d2h += (B - 1) * cap * 4 + (B - 1) * 4
    if do_plane:
        h2d += h_d16.numel() * 2
        d2h += h_blocks.numel() + 2 * B * nblk
    if do_detect:
        h2d += h_d16.numel() * 2
        d2h += h_memdet.numel() * 4 + B * 4 + h_precs.numel()
    if do_surfel:
        h2d += h_gray.numel() + h_depth.numel() * 4 + h_mem.numel() * 4 + 64 * B
        d2h += 32

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cf = CpuFrontend(a, (gray, depth, mem, poses, surfels), threads)
        probe = min(4, B)
        fps_est = probe / cf.step(probe)
        frames = max(min(4, B), min(a.cpu_frames, B, int(fps_est * 20.0)))  # about 20 s of CPU work
        dt = cf.step(frames)
        cpu = {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": baseline_kind(st),
               "sample": cf.sample_text(frames, dt), "legs": cf.parts,
               "stages": cpu_stage_breakdown(a, gray, depth, mem, poses, surfels)}
        del cf

    widened = widened_in_child(local_rank) if (extras and world == 1 and not a.no_cpu_baseline and a.workload == DEFAULT_WORKLOAD) else None

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8+f32", "data": "synthetic", "config": config_of(a),
               "config_detail": {"map_size_end": n_map,
                                 "l2": "working set per step (%d MB surfel planes + %d MB pyramids / frames) exceeds the 126 MB L2" % (
                                     n_map * 56 // 1000000, B * W * H * 5 // 1000000),
                                 "collective": "nccl all_gather_into_tensor of the per-frame {keypoints, new, updated} table on "
                                               "its own stream" if world > 1 else "none (1 GPU)",
                                 "timing": "CUDA events around K steps, all library streams fenced; no timing aid inside the region",
                                 "scenes": ("one synthetic scene per rank (MSL_BENCH_RANK_SCENES=1)" if os.environ.get("MSL_BENCH_RANK_SCENES") == "1"
                                            else "every rank streams the same synthetic scene: per-GPU work is exactly the N = 1 work")},
               "roofline": roofline, "cpu_baseline": cpu,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(fh2d), "d2h_bytes_per_step": int(fd2h),
                       "what": "host C ABI, pinned host buffers: the sensor frames (gray CV_8U + depth CV_16U%s) uploaded once per step "
                               "into a frame set every stage reads (msl_glue_upload_frames, CV_32F depth produced on the device), "
                               "stages through their *_dev entry points, every result copied back to the host inside the step; "
                               "the upload of step k+2 and the launches of step k+1 are issued before the results of step k are awaited (three frame sets, two "
                               "sets of host result buffers), every copy inside the timed region" % (
                                   " + the host-computed membership image" if aux is not None else "")},
               "e2e_host_calls": {"value": e2e_calls_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                                  "what": "the same step through the per-class host entry points (msl_orb_extract, msl_hamming_best2, "
                                          "msl_search_by_projection_frames, msl_plane_prestage / msl_plane_detect, msl_surfel_fuse_batch) "
                                          "from three host threads: every call uploads its own copy of the frame (gray twice, depth "
                                          "three times), which is what a binding that keeps the reference's per-class cv::Mat "
                                          "arguments pays"},
               "e2e_dropin": dropin,
               "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "host_enqueue_ms_unloaded": host_unloaded_ms, "clocks": clocks, "parity_check": parity, "per_rank": per_rank, "widened": widened}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.widened_only:
        import manhattanslam_b200 as msl
        print(json.dumps(widened_ops(msl) if a.widened_only == "matcher" else widened_peac(msl)))
        return
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
