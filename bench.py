#!/usr/bin/env python3
"""bench.py -- RGB-D front-end frames/s on synthetic 640x480 RGB-D (BASELINE.json metric).

One step = one pass of the front-end over one batch of 64 synthetic RGB-D frames per GPU:
ORB extraction, brute-force Hamming match of every frame's descriptors against the next frame's (BASELINE.json
config 2), the plane pre-stage on the u16 depth (config 3) and the projective surfel fusion of the 64-frame
stream into a device-resident 5M-surfel map (config 4: superpixels batched, fuse/initialise/compact frame by
frame in order).  Frames are independent across ranks
(one chunk and one map replica per rank, weak scaling); the only collective is one NCCL all-gather
of the per-frame keypoint counts + surfel statistics.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                           the reference's CPU path (oracle port)

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

W, H = 640, 480
METRIC = "rgbd_frontend_frames_per_s"
UNIT = "frames/s"
MAP_STEADY_FRACTION = 0.89  # measured with the oracle on this generator (see make_inputs)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--surfels", type=int, default=5_000_000, help="surfels in the local map per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--widened-only", nargs="?", const="matcher", default="", choices=["matcher", "peac"],
                    help="internal: print one part of the `widened` object alone (run_ours calls this in child processes so "
                         "that a fault in a diagnostic can never take the bench line down)")
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames in the bounded CPU sample")
    ap.add_argument("--only", default="", help="diagnostic: comma list of stages (orb,match,plane,surfel) the device-resident "
                                               "step runs; the default (empty) is the full front-end -- anything else is not a bench value")
    return ap.parse_args()


def workload_name(a):
    return "frontend_%dx%d_b%d_map%s" % (W, H, a.batch, ("%dM" % (a.surfels // 1_000_000)) if a.surfels >= 1_000_000 else str(a.surfels))


def make_inputs(rank, batch, n_surfels):
    """Seeded synthetic RGB-D batch + pose walk + surfel map."""
    from manhattanslam_b200 import synthetic as S
    seed0 = 1000 * (rank + 1)
    # 16 distinct frames cycled to the batch size keeps start-up short; every frame is still processed.
    # The depth stream is ONE scene (planes from seed0) seen with per-frame sensor noise and holes under the pose
    # walk: a keyframe stream into a local map.  (Independent scenes per frame would kill every in-view surfel in
    # the first pass and leave the fuse step with nothing to update.)
    uniq = min(batch, 16)
    gray_u = [S.gray_frame(seed0 + i) for i in range(uniq)]
    dd = [S.depth_frame(seed0 + i, scene=seed0) for i in range(uniq)]
    depth_u = [d[1] for d in dd]
    gray = np.stack([gray_u[i % uniq] for i in range(batch)])
    depth = np.stack([depth_u[i % uniq] for i in range(batch)])
    make_inputs.depth16 = np.stack([dd[i % uniq][0] for i in range(batch)])
    mem = np.stack([S.membership(seed0 + i) for i in range(batch)])
    poses = S.pose_walk(seed0, batch)
    # ~11 % of a fresh synthetic map leaves in the first few frames (unstable-drop rule :181-184 + the 5 % depth
    # outliers); the map is generated that much larger so that the steady state the timed region sees is n_surfels
    surfels = S.surfel_map(seed0, int(round(n_surfels / MAP_STEADY_FRACTION)), depth[0], poses[0], ref_index=100)
    return gray, depth, mem, poses, surfels


# ------------------------------------------------------------------------------------ CPU arm
_REF_SURFEL = None


def ref_surfel_available():
    """the reference's own src/SurfelFusion.cpp compiled unmodified with the real <thread> (oracle/_ref/libsurfel_ref_threads.so:
    built where /root/reference exists, travels to the GPU box prebuilt) -- present AND loadable"""
    global _REF_SURFEL
    if _REF_SURFEL is None:
        try:
            from oracle import binding as ob
            _REF_SURFEL = ob.build_ref(name="libsurfel_ref_threads.so") is not None
            if _REF_SURFEL:
                ob.RefSurfelFusion(W, H, real_threads=True)  # loads the library, builds and destroys one SurfelFusion
        except Exception:  # noqa: BLE001 -- fall back to the oracle port
            _REF_SURFEL = False
    return _REF_SURFEL


def baseline_kind():
    """"reference" when the dominant CPU stage (SurfelFusion: >90 % of the CPU time of a frame at a 5 M-surfel map) is the
    reference's own source from oracle/_ref; "port" when every stage is the oracle port"""
    return "reference" if ref_surfel_available() else "port"


def surfel_baseline_text():
    if ref_surfel_available():
        return ("SurfelFusion = the reference's own src/SurfelFusion.cpp (oracle/_ref, compiled unmodified against stand-in "
                "OpenCV / Eigen headers, its ten std::threads) frame by frame")
    return "oracle SurfelFusion with 10 scan threads frame by frame"


def cpu_frontend(gray, depth, mem, poses, surfels, frames, threads):
    """The reference's CPU path: ORB + plane pre-stage + Hamming match by the oracle port, frame-parallel over all host threads
    (the reference runs one ORB thread per frame); SurfelFusion frame by frame by the REFERENCE'S OWN src/SurfelFusion.cpp
    (oracle/_ref, compiled unmodified, its ten std::threads, include/SurfelFusion.h:34) on a map that stays inside the
    library like Map::mvLocalSurfels -- or, where that library is absent, by the oracle port with 10 scan threads.
    Returns (frames/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import binding as ob
    frames = min(frames, len(gray))
    local = surfels.copy()
    t0 = time.perf_counter()
    tl = threading.local()

    depth16 = make_inputs.depth16
    descs = [None] * frames

    def orb(i):  # ORB + plane pre-stage of frame i (independent per frame)
        if not hasattr(tl, "o"):
            tl.o = ob.OrbOracle()
        k, d = tl.o(gray[i])
        descs[i] = d
        ob.plane_prestage(depth16[i])
        return len(k)

    def match(i):  # brute-force Hamming best-2 of frame i against frame i+1
        return int(ob.hamming_best2(descs[i], descs[i + 1])[1].sum())

    with ThreadPoolExecutor(max_workers=threads) as ex:
        counts = list(ex.map(orb, range(frames)))
        list(ex.map(match, range(frames - 1)))
    if ref_surfel_available():
        rs = ob.RefSurfelFusion(W, H, real_threads=True)
        rs.set_map(local)
        for i in range(frames):
            rs.fuse_resident(100 + i, gray[i], depth[i], mem[i], poses[i])
            rs.compact_resident()
    else:
        so = ob.SurfelOracle(W, H)
        for i in range(frames):
            new = so.fuse(100 + i, gray[i], depth[i], mem[i], poses[i], local, threads=min(10, threads))
            local = ob.surfel_compact(local, new)
    dt = time.perf_counter() - t0
    assert sum(counts) > 0
    return frames / dt, dt


def cpu_stage_breakdown(gray, depth, mem, poses, surfels, frames=2):
    """SURVEY.md section 8(d) 'CPU baseline timing': where the CPU path spends its time -- each stage of the oracle port on ONE
    thread (ms per frame; the reference runs ORB on one thread per frame), SurfelFusion with its 10 scan threads, and the
    OpenCV primitives ORB is made of timed through cv2 (SIMD, the 'optimised OpenCV' lower bound for the oracle's scalar
    FAST / resize / blur).  A few frames only; reported beside the baseline, never part of it."""
    from oracle import binding as ob
    out = {}
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
        out["plane_prestage_1_thread_ms_per_frame"] = ms(lambda i: ob.plane_prestage(make_inputs.depth16[i]))
        out["plane_detect_1_thread_ms_per_frame"] = ms(lambda i: ob.plane_detect(make_inputs.depth16[i], depth_map_factor=1.0))
        out["hamming_1000x1000_1_thread_ms"] = ms(lambda i: ob.hamming_best2(descs[0], descs[1]), 1)
        so, local = ob.SurfelOracle(W, H), surfels.copy()
        out["surfel_fuse_10_threads_ms_per_frame"] = ms(lambda i: so.fuse(100 + i, gray[i], depth[i], mem[i], poses[i], local,
                                                                          threads=min(10, os.cpu_count() or 1)))
        try:  # the reference's OWN src/SurfelFusion.cpp (oracle/_ref, compiled unmodified) with its own ten std::threads
            rs = ob.RefSurfelFusion(W, H, real_threads=True)
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
    gray, depth, mem, poses, surfels = make_inputs(0, a.batch, a.surfels)
    frames = min(a.cpu_frames, a.batch)
    # bounded sample: one short untimed pass estimates the host's speed, then the frames per step are chosen so that the
    # K timed steps together take about two and a half minutes (never more than --cpu-frames, never fewer than 4)
    fps_est, _ = cpu_frontend(gray, depth, mem, poses, surfels, min(frames, 4), threads)
    frames = max(4, min(frames, int(fps_est * 150.0 / max(a.steps, 1))))
    ts = []
    for _ in range(a.steps):
        fps, dt = cpu_frontend(gray, depth, mem, poses, surfels, frames, threads)
        ts.append(dt)
    ms = 1e3 * sum(ts) / len(ts)
    value = frames / (ms / 1e3)
    sample = "%d of %d frames per step: ORB + plane pre-stage + Hamming match (oracle port) frame-parallel on %d threads, %s into a %d-surfel map" % (
        frames, a.batch, threads, surfel_baseline_text(), a.surfels)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8+f32", "data": "synthetic",
           "config": {"workload": workload_name(a), "batch_per_gpu": a.batch, "surfels_per_gpu": a.surfels,
                      "stages": ["orb", "hamming_match", "plane_prestage", "surfel_fuse"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": baseline_kind(), "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; started before the warm-up (nvidia-smi needs a few
    hundred ms to come up) and filtered to the wall-clock window of the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
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
        kf, f = S.bow_scene(1)
        g_us, (n_g, fm_g) = timed(lambda: m.SearchByBoW(kf, f), reps)
        c_us, (n_c, fm_c) = timed(lambda: ob.search_by_bow(0.7, True, kf, f), 3)
        res["SearchByBoW_1000x1000"] = {"gpu_call_us": g_us, "cpu_oracle_us": c_us, "nmatches": int(n_g),
                                        "equal": bool(n_g == n_c and np.array_equal(fm_g, fm_c))}
        kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(1)
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
        res["note"] = ("one call through the host C ABI incl. the Python mirror's array packing, H2D, kernel, D2H and sync; "
                       "cpu_oracle = the oracle restatement, single thread; not part of the step")
        m.close()
        return res
    except Exception as e:  # noqa: BLE001 -- diagnostics only: never take the bench line down
        return {"error": "%s: %s" % (type(e).__name__, e)}


def widened_peac(msl, frames=8):
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
        ref = [ob.plane_detect(d[b], depth_map_factor=1.0) for b in range(frames)]
        c_s = time.perf_counter() - t0
        equal = all(np.array_equal(mem[b], ref[b][0]) and np.array_equal(planes[b]["N"], ref[b][1]["N"]) and
                    planes[b]["normal"].tobytes() == ref[b][1]["normal"].tobytes() for b in range(frames))
        return {"plane_detect_640x480": {"frames": frames, "gpu_call_ms_per_batch": 1e3 * g_s, "cpu_oracle_ms_per_batch": 1e3 * c_s,
                                         "planes_per_frame": [len(p) for p in planes], "equal": bool(equal),
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
    import subprocess
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
    B = a.batch
    gray, depth, mem, poses, surfels = make_inputs(rank, B, a.surfels)

    orb = msl.ORBextractor(width=W, height=H, max_batch=B, device=local_rank)
    sf = msl.SurfelFusion(W, H, max_surfels=len(surfels) + 4 * B * 4800, device=local_rank)
    sf.upload_map(surfels)
    cap = orb.capacity
    matcher = msl.ORBmatcher(max_queries=cap, max_train=cap, max_batch=B, device=local_rank)
    plane = msl.PlaneDetection(W, H, max_batch=B, device=local_rank)
    depth16 = make_inputs.depth16
    nblk = plane.nblocks

    # pinned host staging (e2e leg) and device-resident inputs (kernel leg)
    h_gray = torch.from_numpy(gray).pin_memory()
    h_depth = torch.from_numpy(depth).pin_memory()
    h_mem = torch.from_numpy(mem).pin_memory()
    h_d16 = torch.from_numpy(depth16.view(np.int16)).pin_memory()
    d_gray, d_depth, d_mem, d_d16 = h_gray.to(dev), h_depth.to(dev), h_mem.to(dev), h_d16.to(dev)
    d_bi = torch.zeros((B, cap), dtype=torch.int32, device=dev)
    d_bd, d_sd = torch.zeros_like(d_bi), torch.zeros_like(d_bi)
    d_blocks = torch.zeros((B, nblk, 72), dtype=torch.uint8, device=dev)
    d_seedm = torch.zeros((B, nblk), dtype=torch.uint8, device=dev)
    d_edges = torch.zeros((B, nblk), dtype=torch.uint8, device=dev)
    h_match = torch.zeros((3, B, cap), dtype=torch.int32).pin_memory()
    h_blocks = torch.zeros((B, nblk, 72), dtype=torch.uint8).pin_memory()
    h_seedm = torch.zeros((2, B, nblk), dtype=torch.uint8).pin_memory()
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
    h_kps = torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory()
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros(B, dtype=torch.int32).pin_memory()
    gathered = torch.zeros((world, B), dtype=torch.int32, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    s_orb = torch.cuda.ExternalStream(orb.stream, device=dev)
    s_sf = torch.cuda.ExternalStream(sf.stream, device=dev)
    s_pl = torch.cuda.ExternalStream(plane.stream, device=dev)
    K4 = (525.0, 525.0, 319.5, 239.5)
    state = {"ref": 100}

    only = set(x for x in a.only.split(",") if x) or {"orb", "match", "plane", "surfel"}

    def step_dev():
        """inputs resident in HBM; ORB and the surfel stream run on their own CUDA streams and overlap"""
        if world > 1:
            s_orb.wait_stream(torch.cuda.current_stream())  # previous all-gather still reads d_counts
        if "orb" in only:
            orb.extract_dev(d_gray.data_ptr(), W, W * H, B, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr())
        # frame b vs frame b+1, chained on the ORB stream (no host sync between extraction and matching)
        if "match" in only:
            matcher.hamming_best2_counts_dev(d_desc.data_ptr(), d_desc.data_ptr() + cap * 32, cap, d_counts.data_ptr(),
                                             d_counts.data_ptr() + 4, B - 1, d_bi.data_ptr(), d_bd.data_ptr(), d_sd.data_ptr(),
                                             stream=orb.stream)
        if "plane" in only:
            plane.prestage_dev(d_d16.data_ptr(), B, K4, 1.0 / 5000.0, None, d_blocks.data_ptr(), d_seedm.data_ptr(),
                               d_edges.data_ptr())
        if "surfel" in only:
            sf.fuse_batch_dev(state["ref"], d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), d_mem.data_ptr(), poses, B, True)
        state["ref"] += B
        if world > 1:  # the path's single collective: per-frame counts to every rank
            torch.cuda.current_stream().wait_stream(s_orb)
            dist.all_gather_into_tensor(gathered.view(-1), d_counts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local_rank)
    for _ in range(a.warmup):
        step_dev()
    barrier()
    launches0 = msl.lib().msl_kernel_launch_count()
    sf.set_timing(1)  # light: scan + apply marks on every 8th frame (an event record costs ~2.7 us of stream time)
    st0 = sf.read_stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_o, ev_s, ev_p = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    cur = torch.cuda.current_stream()
    clk.begin()
    ev0.record(cur)
    s_orb.wait_event(ev0)
    s_sf.wait_event(ev0)
    s_pl.wait_event(ev0)
    for _ in range(a.steps):
        step_dev()
    ev_o.record(s_orb)
    ev_s.record(s_sf)
    ev_p.record(s_pl)
    cur.wait_event(ev_o)
    cur.wait_event(ev_s)
    cur.wait_event(ev_p)
    ev1.record(cur)
    barrier()
    clk.end()
    ms_total = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    launches = int(msl.lib().msl_kernel_launch_count() - launches0)
    fuse_ms, fuse_launches = sf.fuse_kernel_time()
    chain, chain_frames = sf.chain_times()
    sf.set_timing(0)
    st1 = sf.read_stats()
    orb.sync()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = world * B / (ms_step / 1e3)

    # ---- the same kernel timed alone (one extra stream call after a full sync: its superpixel stage precedes its
    # chain and no other stream is busy), to separate the kernel's own efficiency from SM sharing in the timed region
    barrier()
    orb.sync()
    plane.sync()
    sf.set_timing(2)
    sf.fuse_batch_dev(state["ref"], d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), d_mem.data_ptr(), poses, B, True)
    state["ref"] += B
    iso_ms, iso_launches = sf.fuse_kernel_time()
    iso_chain, iso_frames = sf.chain_times()
    sf.set_timing(0)
    st1 = sf.read_stats()

    # ---- roofline of the two kernels of the per-frame fuse chain, measured live with CUDA events recorded inside the
    # library on the launch stream.  The one that takes the larger share of the step is the headline entry.
    n_map = st1[3]
    # stats accumulate per fuse_batch call: st1 holds the last call's totals over its B launches
    upd_per_launch = st1[1] / B
    del_per_launch = st1[2] / B
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "fuse_traffic.json")) as f:
            ncu_traffic = json.load(f)
    except Exception:
        ncu_traffic = {}
    # algorithmic bytes per launch:
    #   k_fuse_scan : 24 B per surfel streamed (the {px, py, pz, size} quad, updateTimes, lastUpdate) + 4 B per surfel
    #                 it kills; the survivor queue (8 B per in-view surfel) is not counted
    #   k_fuse_apply: per fused surfel 8 B queue entry + 36 B read (two quads + updateTimes) + 56 B written (three
    #                 quads + updateTimes + lastUpdate) = 100 B; the 80-byte seed records stay in L1/L2, not counted
    #   k_fuse_one  : (default) both in one kernel: 24 B per surfel streamed + 4 B per killed surfel + per fused surfel
    #                 16 B read (the normal/weight quad) + 56 B written = 72 B; no queue, no second read of q0 / updateTimes
    alg = {"k_fuse_scan": n_map * 24.0 + del_per_launch * 4.0, "k_fuse_apply": upd_per_launch * 100.0,
           "k_fuse_one": n_map * 24.0 + del_per_launch * 4.0 + upd_per_launch * 72.0}
    unit_of = {"k_fuse_scan": ("dram_bytes_per_surfel", n_map), "k_fuse_apply": ("dram_bytes_per_fused", upd_per_launch),
               "k_fuse_one": ("dram_bytes_per_surfel", n_map)}

    def entry(kernel, key, times, frames, iso_times, iso_n):
        ms = times[key] / max(frames, 1)
        iso = iso_times[key] / max(iso_n, 1)
        ach = alg[kernel] / (ms * 1e-3) / 1e9 if ms > 0 else None
        ach_iso = alg[kernel] / (iso * 1e-3) / 1e9 if iso > 0 else None
        tkey, units = unit_of[kernel]
        t = ncu_traffic.get(kernel, {}).get(tkey)
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak if ach else None, "traffic": t * units if t else None, "peak_source": peak_src,
                "alg_bytes_per_launch": alg[kernel], "avg_launch_ms": ms, "launches_timed": frames,
                "launches_in_region": B * a.steps,  # every 8th is bracketed by events (an event costs ~2.7 us of stream time)
                "share_of_step": ms * B / ms_step if ms_step else None,
                "isolated": {"avg_launch_ms": iso, "achieved": ach_iso, "frac": ach_iso / peak if ach_iso else None,
                             "note": "same kernel, same map, no other stream active"}}

    if sf.fuse_kernels() == 1:
        # the "scan" interval of the timing aid brackets k_fuse_one; "apply" is an empty interval (one event record)
        roofline = entry("k_fuse_one", "scan", chain, chain_frames, iso_chain, iso_frames)
        tot = chain["scan"]
        roofline["share_of_fuse_chain"] = {"k_fuse_one": 1.0}
    else:
        r_scan = entry("k_fuse_scan", "scan", chain, chain_frames, iso_chain, iso_frames)
        r_apply = entry("k_fuse_apply", "apply", chain, chain_frames, iso_chain, iso_frames)
        roofline, other = (r_apply, r_scan) if chain["apply"] >= chain["scan"] else (r_scan, r_apply)
        roofline["other_kernel"] = other
        # share_of_step is against the wall time of a step in which three streams overlap (the shares of all kernels sum to
        # more than 1); the ncu launch list serialises every stream.  The split of the fuse chain itself is comparable:
        tot = chain["scan"] + chain["apply"]
        roofline["share_of_fuse_chain"] = {"k_fuse_scan": chain["scan"] / tot, "k_fuse_apply": chain["apply"] / tot} if tot else None
    roofline["chain_us_per_frame"] = {k: 1e3 * v / max(chain_frames, 1) for k, v in chain.items()}
    roofline["isolated"]["chain_us_per_frame"] = {k: 1e3 * v / max(iso_frames, 1) for k, v in iso_chain.items()}
    roofline["fused_per_launch"] = upd_per_launch
    roofline["killed_per_launch"] = del_per_launch
    roofline["note"] = "timed-region launches share the SMs with the next batch's superpixel kernels (stream overlap)"

    # ---- e2e: the public host API with pinned host buffers, H2D of the inputs + D2H of the results every step
    import ctypes as C
    from manhattanslam_b200._lib import check, ptr

    # The three host-API call sequences are issued from three host threads, as the reference does (Frame::ExtractORB
    # and Frame::ExtractPlanes run on per-frame std::threads, src/Frame.cc:100-104; SurfelMapping has its own thread,
    # src/System.cc:98-99).  Every handle owns its stream(s); ctypes releases the GIL for the duration of a call.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=3)
    Kf = np.asarray(K4, np.float32)

    def e2e_orb_match():
        torch.cuda.set_device(local_rank)
        check(orb._L.msl_orb_extract(orb._h, C.c_void_p(h_gray.data_ptr()), C.c_int(W), C.c_size_t(W * H), C.c_int(B),
                                     C.c_void_p(h_kps.data_ptr()), C.c_void_p(h_desc.data_ptr()),
                                     C.c_void_p(h_counts.data_ptr())))
        # matching on the descriptors just produced (device-resident copy inside the ORB handle is not exposed, so the
        # host API path re-uploads them: that is what a host-side caller of the C ABI pays)
        check(matcher._L.msl_hamming_best2(matcher._h, C.c_void_p(h_desc.data_ptr()), C.c_int(cap),
                                           C.c_void_p(h_desc.data_ptr() + cap * 32), C.c_int(cap), C.c_int(B - 1),
                                           C.c_void_p(h_match[0].data_ptr()), C.c_void_p(h_match[1].data_ptr()),
                                           C.c_void_p(h_match[2].data_ptr())))

    def e2e_plane():
        torch.cuda.set_device(local_rank)
        check(plane._L.msl_plane_prestage(plane._h, C.c_void_p(h_d16.data_ptr()), C.c_int(W), C.c_size_t(W * H), C.c_int(B),
                                          ptr(Kf), C.c_float(1.0 / 5000.0), None, C.c_void_p(h_blocks.data_ptr()),
                                          C.c_void_p(h_seedm[0].data_ptr()), C.c_void_p(h_seedm[1].data_ptr())))

    def e2e_surfel(ref):
        torch.cuda.set_device(local_rank)
        stats = np.zeros(4, np.int64)
        check(sf._L.msl_surfel_fuse_batch(sf._h, ref, C.c_void_p(h_gray.data_ptr()), C.c_int(W),
                                          C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_mem.data_ptr()), ptr(poses),
                                          C.c_int(B), 1, ptr(stats)))
        return stats

    def step_e2e():
        futs = [pool.submit(e2e_orb_match), pool.submit(e2e_plane), pool.submit(e2e_surfel, state["ref"])]
        state["ref"] += B
        return [f.result() for f in futs][2]

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
    e2e_value = world * B * a.steps / float(t.item())
    h2d = int(h_gray.numel() + h_depth.numel() * 4 + h_mem.numel() * 4 + h_d16.numel() * 2 + 2 * (B - 1) * cap * 32 + 64 * B)
    d2h = int(h_kps.numel() + h_desc.numel() + h_counts.numel() * 4 + 3 * (B - 1) * cap * 4 + h_blocks.numel() +
              2 * B * nblk + 32)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        frames = min(a.cpu_frames, B)
        fps, dt = cpu_frontend(gray, depth, mem, poses, surfels, frames, threads)
        cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": baseline_kind(),
               "sample": "%d frames (%.1f s): oracle ORB + plane pre-stage + Hamming match frame-parallel on %d threads, %s into the "
                         "%d-surfel map" % (frames, dt, threads, surfel_baseline_text(), a.surfels),
               "stages": cpu_stage_breakdown(gray, depth, mem, poses, surfels)}

    widened = widened_in_child(local_rank) if (rank == 0 and world == 1 and not a.no_cpu_baseline) else None

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8+f32", "data": "synthetic",
               "config": {"workload": workload_name(a) + ("" if not a.only else "_DIAGNOSTIC_only_" + a.only), "batch_per_gpu": B, "surfels_per_gpu": a.surfels,
                          "stages": ["orb", "hamming_match", "plane_prestage", "surfel_fuse"], "map_size_end": n_map,
                          "l2": "working set per step (280 MB surfel planes + 190 MB pyramids) exceeds the 126 MB L2",
                          "collective": "nccl all_gather of per-frame counts" if world > 1 else "none (1 GPU)"},
               "roofline": roofline, "cpu_baseline": cpu,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": launches, "clocks": clocks, "widened": widened}
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
