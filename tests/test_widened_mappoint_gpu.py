"""GPU parity: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263) batched over map points, through the
C ABI, vs the CPU oracle.  (Its own file, sorted after the other GPU tests: the kernel was written after the round's GPU
budget was spent, so the round-end run is its first -- see DESIGN.md section 10.)"""
import os

import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S

# never run on hardware before the round end: a hang must end the process instead of holding the GPU box (pytest-timeout's
# thread method exits the interpreter, which tears the CUDA context down)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("seed", [1, 2])
def test_distinctive_descriptors(oracle, msl, seed):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263), batched over map points: BestIdx and BestMedian."""
    sets = S.observation_sets(seed)
    m = msl.ORBmatcher()
    bi_o, bm_o = oracle.distinctive_descriptors(sets)
    bi_g, bm_g = m.ComputeDistinctiveDescriptors(sets)
    assert np.array_equal(bi_o, bi_g) and np.array_equal(bm_o, bm_g)
    assert (bi_g == -1).sum() == sum(len(d) == 0 for d in sets) > 0
    # the ragged edges: 1, 2, 3 observations, exactly one warp, one more than a warp, and nothing at all
    r = np.random.default_rng(seed)
    edge = [r.integers(0, 256, (n, 32), dtype=np.uint8) for n in (1, 2, 3, 4, 5, 31, 32, 33, 64, 129)]
    bi_o, bm_o = oracle.distinctive_descriptors(edge)
    bi_g, bm_g = m.ComputeDistinctiveDescriptors(edge)
    assert np.array_equal(bi_o, bi_g) and np.array_equal(bm_o, bm_g)
    bi_g, bm_g = m.ComputeDistinctiveDescriptors([])
    assert len(bi_g) == 0
    gold = np.load(os.path.join(GOLD, "node_searches.npz"))
    if seed == 1:
        bi_g, bm_g = m.ComputeDistinctiveDescriptors(S.observation_sets(3))
        assert np.array_equal(bi_g, gold["dd_idx"]) and np.array_equal(bm_g, gold["dd_median"])
