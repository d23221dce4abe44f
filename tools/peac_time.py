#!/usr/bin/env python3
"""Time msl_plane_detect (f2: pre-stage + ahCluster + refineDetails, one CTA per frame) through the host C ABI for a batch of
synthetic 640x480 depth frames, check it against the oracle on a few frames, print one JSON line.
    python tools/peac_time.py [batch] [reps]          (MSL_PEAC_FLOOD_SERIAL=1 for the FIFO region grow)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import manhattanslam_b200 as msl  # noqa: E402
from manhattanslam_b200 import synthetic as S  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    d = np.stack([S.depth_frame(200 + b)[0] for b in range(batch)])
    pd = msl.PlaneDetection(max_batch=batch)
    mem, planes = pd.detect(d, depthMapFactor=1.0)  # warm-up (allocations, shared-memory opt-in)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        mem, planes = pd.detect(d, depthMapFactor=1.0)
        ts.append(time.perf_counter() - t0)
    prof = None
    try:
        pr = pd.debug_profile(batch).astype(np.float64)
        names = ["graph", "ahCluster", "membership+seeds", "region_grow", "final_merge", "remap"]
        dd = np.diff(pr[:, :7], axis=1) / 1e3
        prof = {"phase_us_median_over_frames": {n: float(np.median(dd[:, i])) for i, n in enumerate(names)},
                "frame_us_median": float(np.median((pr[:, 6] - pr[:, 0]) / 1e3)), "frame_us_max": float(np.max((pr[:, 6] - pr[:, 0]) / 1e3)),
                "batch_span_us": float((pr[:, 6].max() - pr[:, 0].min()) / 1e3), "merge_steps_median": float(np.median(pr[:, 7])),
                "ahCluster_cycles_per_step_median": {n: float(np.median(pr[:, 8 + i] / np.maximum(pr[:, 7], 1))) for i, n in enumerate(
                    ["pop", "candidate_fits", "select", "publish_decide", "adjacency", "copy_push"])}}
    except Exception as e:  # noqa: BLE001
        prof = "unavailable: %s" % e
    ok = None
    try:
        from oracle import binding as ob
        ok = all(np.array_equal(mem[b], ob.plane_detect(d[b], depth_map_factor=1.0)[0]) for b in range(min(batch, 4)))
    except Exception as e:  # noqa: BLE001
        ok = "oracle unavailable: %s" % e
    print(json.dumps({"batch": batch, "flood_serial": os.environ.get("MSL_PEAC_FLOOD_SERIAL", "0"), "ms_per_batch_min": 1e3 * min(ts),
                      "ms_per_batch_median": 1e3 * float(np.median(ts)), "frames_per_s": batch / float(np.median(ts)),
                      "threads": os.environ.get("MSL_PEAC_THREADS", "256"), "profile": prof, "planes_first_frames": [len(p) for p in planes[:8]], "equals_oracle_first_frames": ok,
                      "note": "host API: H2D of the depth frames, pre-stage, k_peac_frame, D2H of membership + planes, sync"}))


if __name__ == "__main__":
    main()
