"""Dry run of GPU test SCRIPTS without a GPU (tools/dryrun_gpu_tests.sh): `msl` is replaced by a fake whose classes answer
from the CPU oracle, so that the Python of a test file written while no GPU was available -- dtype views, shapes, argument
order, fixtures, golden-file keys -- is exercised before its first real run.  It proves nothing about the kernels."""
import os, sys
import numpy as np
import pytest
sys.path.insert(0, os.environ.get("MSL_REPO", "/root/repo"))
from oracle import binding as B
from manhattanslam_b200.plane import PLANE_DTYPE, BLOCK_DTYPE
from manhattanslam_b200.matcher import frame_geom

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: x")

class FakeORB:
    def __init__(self, width=640, height=480, max_batch=1): self.o = B.OrbOracle()
    def __call__(self, img): return self.o(img)

class FakePlane:
    def __init__(self, width=640, height=480, max_batch=1):
        self.width, self.height = width, height
        if (((width+1)//2)//10) * (((height+1)//2)//10) > 3072: self.big = True
        else: self.big = False
    def prestage(self, d16, K=(525.0,525.0,319.5,239.5), depthMapFactor=1/5000., want_cloud=True):
        c,b,s,e = B.plane_prestage(d16, K=K, depth_map_factor=depthMapFactor)
        return (c[None] if want_cloud else None), b[None], s[None], e[None]
    def detect(self, d, K=(525.0,525.0,319.5,239.5), depthMapFactor=1/5000., plane_cap=32):
        if self.big: raise RuntimeError("unsupported")
        d = np.asarray(d)
        if d.ndim == 2: d = d[None]
        mems, pls = [], []
        for b in range(len(d)):
            m, p = B.plane_detect(d[b], K=K, depth_map_factor=depthMapFactor, cap=plane_cap)
            rec = np.zeros(len(p["N"]), PLANE_DTYPE)
            for f in ("normal","center","N","rid","vertices"): rec[f] = p[f]
            mems.append(m); pls.append(rec)
        return np.stack(mems), pls

class FakeSurfel:
    def __init__(self, width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, fuseFar=30.0, fuseNear=0.5, max_surfels=0):
        self.o = B.SurfelOracle(width, height, fx, fy, cx, cy, fuseFar, fuseNear)
        self.m = B.SurfelMappingOracle()
        self.local = np.zeros(0, B.SURFEL_DTYPE)
    def upload_map(self, local): self.local = local.copy()
    def moveAddSurfels(self, rem, add):
        self.local = self.m.move_add(self.local, rem, add)
        return (0, 0, len(self.local))
    def fuseInitializeMap(self, ref, g, d, m, T, compact=False):
        new = self.o.fuse(ref, g, d, m, T, self.local)
        if compact: self.local = B.surfel_compact(self.local, new)
        return new, (len(new), 0, 0, len(self.local))
    def download_map(self): return self.local.copy()
    def download_inactive(self): return self.m.inactive()
    def debug_index(self): return self.o.index()

class FakeMatcher:
    def __init__(self, nnratio=0.6): self.nn = nnratio; self.mbCheckOrientation = True
    def SearchByProjectionFrame(self, g, Tc, Tl, th, last, cur): return B.search_by_projection_frame(g, Tc, Tl, th, self.mbCheckOrientation, last, cur)
    def SearchByProjectionKeyFrame(self, g, Tc, th, od, kf, cur, lsf): return B.search_by_projection_keyframe(g, Tc, th, od, self.mbCheckOrientation, lsf, kf, cur)
    def SearchForTriangulation(self, kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls, bOnlyStereo=False): return B.search_for_triangulation(F12, Cw1, Tcw2, K2, bOnlyStereo, self.mbCheckOrientation, sf, ls, kf1, kf2)
    def Fuse(self, g, Tcw, mps, kf, ils, th=3.0, log_scale_factor=0.0): return B.fuse_search(g, Tcw, th, log_scale_factor, ils, mps, kf)
    def SearchByProjectionPoints(self, g, th, mps, cur): return B.search_by_projection_points(g, th, self.nn, mps, cur)
    def SearchByBoW(self, kf, f): return B.search_by_bow(self.nn, self.mbCheckOrientation, kf, f)
    def ComputeDistinctiveDescriptors(self, sets):
        if len(sets) == 0: return np.zeros(0, np.int32), np.zeros(0, np.int32)
        return B.distinctive_descriptors(sets)

class FakeMsl:
    ORBextractor = FakeORB; PlaneDetection = FakePlane; SurfelFusion = FakeSurfel; ORBmatcher = FakeMatcher
    SURFEL_DTYPE = B.SURFEL_DTYPE
    frame_geom = staticmethod(frame_geom)

@pytest.fixture(scope="session")
def oracle():
    B.build(); return B
@pytest.fixture(scope="session")
def msl(): return FakeMsl
