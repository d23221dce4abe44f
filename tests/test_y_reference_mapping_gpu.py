"""GPU parity of f1 against golden vectors produced by the REFERENCE'S OWN SurfelMapping (tests/golden/reference_source.npz,
tests/golden/make_golden_ref.py; src/SurfelMapping.cpp + src/SurfelFusion.cpp compiled unmodified).  Its own file, sorted
last: the stream runs at 320x240, a frame size the CUDA SurfelFusion has not been run at before, and was written after the
round's GPU budget was spent."""
import os
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_source.npz"))


@pytest.fixture(scope="module")
def G():
    import make_golden_ref
    return make_golden_ref


def test_surfel_mapping_stream_equals_reference_source(msl, gold, G):
    """f1 on the device-resident maps against the reference's own SurfelMapping (src/SurfelMapping.cpp compiled unmodified):
    34 keyframes of moveAddSurfels (with the pose lists the reference's getAddRemovePoses produced) + fuseMap; both surfel
    vectors at the end, record for record"""
    g, mem, poses, depths = G.mapping_inputs()
    sf = msl.SurfelFusion(G.MAP_W, G.MAP_H, *G.MAP_K, max_surfels=20000)
    sf.upload_map(np.zeros(0, msl.SURFEL_DTYPE))
    for i, ri in enumerate(G.MAP_REFS):
        add = gold["map_add"][gold["map_add_off"][i]:gold["map_add_off"][i + 1]]
        rem = gold["map_rem"][gold["map_rem_off"][i]:gold["map_rem_off"][i + 1]]
        if len(add) or len(rem):
            sf.moveAddSurfels(rem, add)
        sf.fuseInitializeMap(ri, g, depths[i], mem, poses[i], compact=True)
    local, inactive = sf.download_map(), sf.download_inactive()
    ref_local = gold["map_local"].copy().view(local.dtype).reshape(-1)
    ref_inactive = gold["map_inactive"].copy().view(local.dtype).reshape(-1)
    for name, a, r in (("local", local, ref_local), ("inactive", inactive, ref_inactive)):
        assert len(a) == len(r), name
        for f in a.dtype.names:
            if a[f].dtype.kind == "f":
                assert np.allclose(a[f], r[f], rtol=1e-4, atol=1e-6, equal_nan=True), (name, f)
            else:
                assert np.array_equal(a[f], r[f]), (name, f)
    assert len(ref_inactive) > 1000 and gold["map_add_off"][-1] > 10
