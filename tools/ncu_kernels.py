#!/usr/bin/env python3
"""Driver for per-kernel ncu captures (north_star: every kernel committed with an ncu capture): runs every stage of the
front-end once on a batch of synthetic 640x480 frames with device-resident inputs -- ORB extraction, brute-force Hamming,
the three SearchByProjection window searches, the plane pre-stage, plane detection (peac) on a few frames, the superpixel
stage and a few frames of the surfel fuse chain into a 1 M-surfel map -- so that one
  ncu --set full --clock-control none --import-source on -k regex:'^k_' -o OUT python tools/ncu_kernels.py
holds at least one launch of every kernel of the step.  Nothing printed here is a bench value."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import manhattanslam_b200 as msl
    from manhattanslam_b200 import synthetic as S
    import bench
    B = int(os.environ.get("NCU_BATCH", "64"))
    W, H = 640, 480
    dev = torch.device("cuda", 0)
    gray, depth, mem, poses, surfels = bench.make_inputs(0, B, int(os.environ.get("NCU_SURFELS", "1000000")))
    d16 = bench.make_inputs.depth16
    orb = msl.ORBextractor(width=W, height=H, max_batch=B)
    cap = orb.capacity
    d_gray = torch.from_numpy(gray).to(dev)
    d_depth = torch.from_numpy(depth).to(dev)
    d_mem = torch.from_numpy(mem).to(dev)
    d_d16 = torch.from_numpy(d16.view(np.int16)).to(dev)
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros(B, dtype=torch.int32, device=dev)
    orb.extract_dev(d_gray.data_ptr(), W, W * H, B, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr())
    orb.sync()
    m = msl.ORBmatcher(max_queries=cap, max_train=cap, max_batch=B)
    d_bi = torch.zeros((B, cap), dtype=torch.int32, device=dev)
    d_bd, d_sd = torch.zeros_like(d_bi), torch.zeros_like(d_bi)
    m.hamming_best2_counts_dev(d_desc.data_ptr(), d_desc.data_ptr() + cap * 32, cap, d_counts.data_ptr(), d_counts.data_ptr() + 4,
                               B - 1, d_bi.data_ptr(), d_bd.data_ptr(), d_sd.data_ptr())
    torch.cuda.synchronize()
    # the three window searches (k_search) on synthetic tracking scenes, host API
    geom = msl.frame_geom()
    cur, last, mps, Tc, Tl = S.match_scene(1)
    m.SearchByProjectionFrame(geom, Tc, Tl, 7.0, last, cur)
    # the batched SearchByProjection(frame b+1, frame b): frame glue, UpdateLastFrame on the device, k_search_batch
    glue = msl.FrameGlue(W, H, max_batch=B)
    d_xy = torch.zeros((B, cap, 2), dtype=torch.float32, device=dev)
    d_ur, d_kd = torch.zeros((B, cap), dtype=torch.float32, device=dev), torch.zeros((B, cap), dtype=torch.float32, device=dev)
    d_cm, d_nm = torch.zeros((B, cap), dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)
    K4 = (525.0, 525.0, 319.5, 239.5)
    glue.keypoints_dev(d_kps.data_ptr(), cap, d_counts.data_ptr(), B, K4, None, d_depth.data_ptr(), 40.0, d_xy.data_ptr(),
                       d_ur.data_ptr(), d_kd.data_ptr())
    torch.cuda.synchronize()
    Tcw = np.stack([np.linalg.inv(p.astype(np.float64)) for p in poses]).astype(np.float32)
    m.SearchByProjectionFrames_dev(msl.frame_geom(W, H, *K4, bf=40.0), 15.0, 3.05, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                   d_counts.data_ptr(), B, d_xy.data_ptr(), d_ur.data_ptr(), d_kd.data_ptr(), Tcw, d_cm.data_ptr(),
                                   d_nm.data_ptr())
    torch.cuda.synchronize()
    pl = msl.PlaneDetection(W, H, max_batch=B)
    nblk = pl.nblocks
    d_blocks = torch.zeros((B, nblk, 72), dtype=torch.uint8, device=dev)
    d_seedm = torch.zeros((B, nblk), dtype=torch.uint8, device=dev)
    d_edges = torch.zeros((B, nblk), dtype=torch.uint8, device=dev)
    pl.prestage_dev(d_d16.data_ptr(), B, (525.0, 525.0, 319.5, 239.5), 1.0 / 5000.0, None, d_blocks.data_ptr(), d_seedm.data_ptr(),
                    d_edges.data_ptr())
    pl.sync()
    pl.detect(d16[:min(B, 8)], depthMapFactor=1.0 / 5000.0)
    sf = msl.SurfelFusion(W, H, max_surfels=len(surfels) + 4 * B * 4800)
    sf.upload_map(surfels)
    nb = min(B, int(os.environ.get("NCU_FUSE_FRAMES", "8")))
    # superpixel kernels on the whole batch (one launch each per iteration), then a few frames of the fuse chain
    sf.fuse_batch_dev(100, d_gray.data_ptr(), W, W * H, d_depth.data_ptr(), d_mem.data_ptr(), poses, B if nb == B else nb, True)
    sf.sync()
    if nb != B:
        sf.superpixels(gray, depth, mem, want_index=False)
    print("ncu_kernels: done, %d launches" % msl.lib().msl_kernel_launch_count())


if __name__ == "__main__":
    main()
