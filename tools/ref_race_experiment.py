#!/usr/bin/env python3
"""How deterministic is the reference's own SurfelFusion?  Builds /root/reference/src/SurfelFusion.cpp twice against the
stand-in OpenCV / Eigen headers of oracle/ref_shim_cv/: once with the sequential <thread> stand-in (the serialisation the
oracle fixes) and once with the REAL std::thread (ten slices racing on the `stable` flag of updatePixelsKernel,
src/SurfelFusion.cpp:357-415), runs the threaded build several times per input and counts the pixels whose superpixel
index differs from the sequential result.  Test infrastructure; needs /root/reference.

Observed here (8 host cores): 13 of 20 threaded runs differ from the sequential serialisation, by 1-11 of 307,200
superpixel indices; runs on the same input differ from each other."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manhattanslam_b200 import synthetic as S  # noqa: E402
from oracle import binding as ob  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="msl_race_")
    try:
        # the include path WITHOUT oracle/ref_shim_cv/seq_thread: <thread> falls through to the system header
        so = os.path.join(tmp, "libsurfel_ref_threads.so")
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-w", "-pthread",
                               "-I" + os.path.join(ROOT, "oracle", "ref_shim_cv"), "-I" + os.path.join(ROOT, "oracle"),
                               "-I/root/reference/include", "-shared", "-o", so, "/root/reference/src/SurfelFusion.cpp",
                               os.path.join(ROOT, "oracle", "ref_wrap.cpp")])
        ob.build()
        ob.build_ref = lambda force=False, name="libsurfel_ref.so": so
        differ = runs = 0
        for seed, pf, n in [(3, 0.0, 30000), (4, 0.4, 50000), (5, 0.2, 10000), (8, 0.1, 20000)]:
            g = S.gray_frame(seed)
            _, d = S.depth_frame(seed)
            m = S.membership(seed, plane_fraction=pf)
            T = S.pose_walk(seed, 1)[0]
            local = S.surfel_map(seed, n, d, T, ref_index=20)
            lo = local.copy()
            o = ob.SurfelOracle()
            o.fuse(20, g, d, m, T, lo)
            for rep in range(5):
                lr = local.copy()
                r = ob.RefSurfelFusion()
                r.fuse(20, g, d, m, T, lr)
                npx = int((o.index() != r.index()).sum())
                runs += 1
                differ += npx > 0
                print("seed %d run %d: %d superpixel indices differ from the sequential serialisation" % (seed, rep, npx))
        print("%d of %d threaded runs differ" % (differ, runs))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
