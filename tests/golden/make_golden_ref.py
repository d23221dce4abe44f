#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/reference_source.npz).

oracle/_ref/*.so are /root/reference/src/{ORBextractor.cc, ORBmatcher.cc, PlaneExtractor.cpp, SurfelFusion.cpp} compiled
unmodified against stand-in headers (oracle/Makefile).  This script runs them on seeded synthetic
inputs and stores their outputs, so that machines without /root/reference (the GPU box) can still compare against the
reference's code: tests/test_oracle_stages.py::test_oracle_equals_reference_source_golden (oracle, CPU) and
tests/test_v_reference_golden_gpu.py (CUDA path, GPU).  Run it here, where /root/reference exists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manhattanslam_b200 import synthetic as S  # noqa: E402
from manhattanslam_b200.matcher import frame_geom  # noqa: E402
from oracle import binding as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LSF = float(np.float32(np.log(np.float64(np.float32(1.2)))))
ORB_SEED, PLANE_SEED, SURFEL_SEED, MATCH_SEED = 61, 62, 63, 64
N_SURFELS = 6000


def surfel_inputs():
    g = S.gray_frame(SURFEL_SEED)
    _, d = S.depth_frame(SURFEL_SEED)
    m = S.membership(SURFEL_SEED, plane_fraction=0.25)
    T = S.pose_walk(SURFEL_SEED, 1)[0]
    local = S.surfel_map(SURFEL_SEED, N_SURFELS, d, T, ref_index=20)
    return g, d, m, T, local


MAP_W, MAP_H = 320, 240
MAP_K = tuple(k * 0.5 for k in S.K_DEFAULT)
MAP_REFS = [0] + [i - 1 for i in range(1, 26)] + [3, 26, 4, 28, 27, 2, 30, 31]


def mapping_inputs():
    """the keyframe stream of the SurfelMapping golden: (gray, membership, poses, depth frames by keyframe)"""
    poses = S.pose_walk(5, len(MAP_REFS))
    g = S.gray_frame(5, MAP_W, MAP_H)
    mem = np.ascontiguousarray(S.membership(5, MAP_W, MAP_H, plane_fraction=0.2), np.int32)
    depths = [S.depth_frame(300 + i % 4, MAP_W, MAP_H, K=MAP_K, scene=300)[1] for i in range(len(MAP_REFS))]
    return g, mem, poses, depths


def main():
    out = {}
    # ORBextractor::operator() (src/ORBextractor.cc) on one frame
    k, dsc = B.RefOrbExtractor()(S.gray_frame(ORB_SEED))
    out["orb_kps"], out["orb_desc"] = k.view(np.uint8).reshape(len(k), -1), dsc
    # PlaneDetection::readDepthImage + PlaneSeg + initGraph, and the full peac run (src/PlaneExtractor.cpp, include/peac/)
    d16, _ = S.depth_frame(PLANE_SEED)
    _, blocks, seed, edges = B.ref_plane_prestage(d16, depth_map_factor=1.0)
    out["plane_blocks"], out["plane_seed"], out["plane_edges"] = blocks.view(np.uint8).reshape(len(blocks), -1), seed, edges
    mem, planes = B.ref_plane_run(d16, depth_map_factor=1.0)
    out["peac_membership"] = mem.astype(np.int8)  # values -5 .. a few planes
    out["peac_plane_N"] = planes["N"]
    # SurfelFusion::fuseInitializeMap (src/SurfelFusion.cpp) with the membership image peac produced above?  No: the
    # synthetic membership keeps the case independent; a second case uses the real peac image.
    g, d, m, T, local = surfel_inputs()
    r = B.RefSurfelFusion()
    lo = local.copy()
    new = r.fuse(21, g, d, m, T, lo)
    out["surfel_index"] = r.index().astype(np.int16)
    out["surfel_seeds"] = r.seeds().view(np.uint8).reshape(-1, r.seeds().dtype.itemsize)
    out["surfel_local_after"] = lo.view(np.uint8).reshape(len(lo), -1)
    out["surfel_new"] = new.view(np.uint8).reshape(len(new), -1)
    lo2 = local.copy()
    _, d62 = S.depth_frame(PLANE_SEED)
    new2 = r.fuse(21, g, d62, mem, T, lo2)  # the reference's own membership image (trail counters <= -2 included)
    out["surfel_peac_index"] = r.index().astype(np.int16)
    out["surfel_peac_new"] = new2.view(np.uint8).reshape(len(new2), -1)
    # ORBmatcher (src/ORBmatcher.cc): match tables with NULL for slots reset by the rotation check
    geom = frame_geom()
    cur, last, mps, Tc, Tl = S.match_scene(MATCH_SEED)
    cur2, kf, Tc2 = S.reloc_scene(MATCH_SEED)
    kfb, f = S.bow_scene(MATCH_SEED)
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(MATCH_SEED)
    mpf, kfs, Tcw, ils = S.fuse_scene(MATCH_SEED)
    with B.reference_matcher():
        out["m_frame_n"], out["m_frame"] = B.search_by_projection_frame(geom, Tc, Tl, 15.0, True, last, cur)
        out["m_points_n"], out["m_points"] = B.search_by_projection_points(geom, 3.0, 0.8, mps, cur)
        out["m_reloc_n"], out["m_reloc"] = B.search_by_projection_keyframe(geom, Tc2, 15.0, 100, True, LSF, kf, cur2)
        out["m_bow_n"], out["m_bow"] = B.search_by_bow(0.7, True, kfb, f)
        out["m_tri_n"], out["m_tri"] = B.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, kf1, kf2)
    out["m_fuse_n"], out["m_fuse"] = B.ref_fuse(geom, Tcw, 3.0, LSF, ils, mpf, kfs)
    # SurfelMapping (src/SurfelMapping.cpp): a 34-keyframe pose graph through InsertKeyFrame + ProcessNewKeyFrame; the pose lists
    # getAddRemovePoses handed to moveAddSurfels per keyframe, and both surfel vectors at the end
    g, mem, poses, depths = mapping_inputs()
    rm = B.RefSurfelMapping(MAP_W, MAP_H, *MAP_K)
    adds, rems = [], []
    for i, ri in enumerate(MAP_REFS):
        a, r_ = rm.keyframe(g, depths[i], mem, poses[i], ri)
        adds.append(a), rems.append(r_)
    out["map_add_off"] = np.cumsum([0] + [len(a) for a in adds]).astype(np.int32)
    out["map_add"] = np.concatenate(adds + [np.zeros(0, np.int32)]).astype(np.int32)
    out["map_rem_off"] = np.cumsum([0] + [len(a) for a in rems]).astype(np.int32)
    out["map_rem"] = np.concatenate(rems + [np.zeros(0, np.int32)]).astype(np.int32)
    out["map_local"] = rm.local().view(np.uint8).reshape(-1, 56)
    out["map_inactive"] = rm.inactive().view(np.uint8).reshape(-1, 56)
    np.savez_compressed(os.path.join(HERE, "reference_source.npz"), **out)
    print("written:", {k: (np.asarray(v).shape if np.ndim(v) else int(v)) for k, v in out.items()})
    print("bytes:", os.path.getsize(os.path.join(HERE, "reference_source.npz")))


if __name__ == "__main__":
    main()
