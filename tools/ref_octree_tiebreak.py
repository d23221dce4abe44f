#!/usr/bin/env python3
"""How much of ORBextractor's output depends on the heap allocator?  (TEST INFRASTRUCTURE / evidence for DESIGN.md section 2)

DistributeOctTree sorts (key count, ExtractorNode*) pairs (src/ORBextractor.cc:654): nodes with equal counts are expanded
in HEAP-ADDRESS order.  oracle/_ref/liborb_ref.so is the reference's own src/ORBextractor.cc; this script runs it on the
same frames (a) inside a bump arena, where address order = creation order (what the oracle restates), and (b) on glibc's
allocator, and reports how many keypoints differ.  Needs /root/reference (or a prebuilt oracle/_ref/liborb_ref.so)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manhattanslam_b200 import synthetic as S  # noqa: E402
from oracle import binding as B  # noqa: E402


def main(n=12):
    o, arena, heap = B.OrbOracle(), B.RefOrbExtractor(arena=True), B.RefOrbExtractor(arena=False)
    differing = 0
    for seed in range(n):
        g = S.gray_frame(seed)
        ko, do = o(g)
        ka, da = arena(g)
        kh, dh = heap(g)
        assert ko.tobytes() == ka.tobytes() and np.array_equal(do, da), "oracle != reference in the arena"
        sa, sh = set(map(bytes, ka)), set(map(bytes, kh))
        differing += sa != sh
        print("frame %2d: arena %4d keypoints (== oracle), glibc %4d; only in arena %2d, only with glibc %2d"
              % (seed, len(ka), len(kh), len(sa - sh), len(sh - sa)))
    print("%d of %d frames differ between the two allocators" % (differing, n))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)
