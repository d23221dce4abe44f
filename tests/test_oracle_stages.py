"""CPU tests of the non-ORB oracle stages: independent cross-checks (numpy/LAPACK brute force) of the parts
that restate third-party arithmetic, invariants of the restated reference logic, and golden vectors."""
import os
import sys

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
        # cv::gemm small-matrix path: float row sum, then (float)(t0 + c) in double (test_cv_gemm_semantics)
        row = lambda r: f32(f64(f32(f32(f32(R[r, 0] * x[0]) + f32(R[r, 1] * x[1])) + f32(R[r, 2] * x[2]))) + f64(t[r]))
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


# ---- independent Python restatements of the vocabulary-node searches and of Fuse (small cases pin the C++ oracle's control flow)
def _pop(a, b):
    return int(np.unpackbits(np.asarray(a, np.uint8) ^ np.asarray(b, np.uint8)).sum())


def _rot_prune(hist, clear):
    """ComputeThreeMaxima (src/ORBmatcher.cc:799-830) + the pruning loops; returns the number of removed matches"""
    sizes = [len(h) for h in hist]
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < np.float32(0.1) * np.float32(m1):
        i2 = i3 = -1
    elif m3 < np.float32(0.1) * np.float32(m1):
        i3 = -1
    removed = 0
    for i in range(30):
        if i not in (i1, i2, i3):
            for j in hist[i]:
                clear(j)
                removed += 1
    return removed


def _bin(a, b):
    f32, f64 = np.float32, np.float64
    rot = f32(f32(a) - f32(b))
    if rot < 0:
        rot = f32(rot + f32(360.0))
    k = int(np.floor(f64(f32(rot * f32(1.0 / 30))) + 0.5))
    return 0 if k == 30 else k


def _py_bow(nnratio, check, kf, f):
    fm = np.full(len(f["angle"]), -1, np.int32)
    hist = [[] for _ in range(30)]
    n = 0
    for node in sorted(set(kf["featvec"]) & set(f["featvec"])):
        for ik in kf["featvec"][node]:
            if not kf["valid"][ik]:
                continue
            b1, b2, bi = 256, 256, -1
            for jf in f["featvec"][node]:
                if fm[jf] >= 0:
                    continue
                d = _pop(kf["desc"][ik], f["desc"][jf])
                if d < b1:
                    b2, b1, bi = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= 50 and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                fm[bi] = ik
                n += 1
                if check:
                    hist[_bin(kf["angle"][ik], f["angle"][bi])].append(bi)
    if check:
        def clear(j):
            fm[j] = -3
        n -= _rot_prune(hist, clear)
    return n, fm


def _py_triangulation(F12, Cw1, Tcw2, K2, only_stereo, check, sf, ls, kf1, kf2):
    f32, f64 = np.float32, np.float64
    T = np.asarray(Tcw2, f32).reshape(4, 4)
    Cw = np.asarray(Cw1, f32)
    C2 = [f32(f64(f32(f32(f32(T[r, 0] * Cw[0]) + f32(T[r, 1] * Cw[1])) + f32(T[r, 2] * Cw[2]))) + f64(T[r, 3])) for r in range(3)]
    invz = f32(f32(1.0) / C2[2])
    ex = f32(f32(f32(K2[0] * C2[0]) * invz) + K2[2])
    ey = f32(f32(f32(K2[1] * C2[1]) * invz) + K2[3])
    F = np.asarray(F12, f32).reshape(3, 3)
    m12 = np.full(len(kf1["angle"]), -1, np.int32)
    hist = [[] for _ in range(30)]
    n = 0
    for node in sorted(set(kf1["featvec"]) & set(kf2["featvec"])):
        for i1 in kf1["featvec"][node]:
            if kf1["has_mp"][i1]:
                continue
            st1 = kf1["uright"][i1] >= 0
            if only_stereo and not st1:
                continue
            x1, y1 = kf1["xy"][i1]
            a = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
            b = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
            c = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
            best, bi = 50, -1
            for i2 in kf2["featvec"][node]:
                if kf2["has_mp"][i2]:
                    continue
                st2 = kf2["uright"][i2] >= 0
                if only_stereo and not st2:
                    continue
                d = _pop(kf1["desc"][i1], kf2["desc"][i2])
                if d > 50 or d > best:
                    continue
                x2, y2 = kf2["xy"][i2]
                o2 = kf2["octave"][i2]
                if not st1 and not st2:
                    dx, dy = f32(ex - x2), f32(ey - y2)
                    if f32(f32(dx * dx) + f32(dy * dy)) < f32(f32(100) * sf[o2]):
                        continue
                num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                den = f32(f32(a * a) + f32(b * b))
                if den == 0:
                    continue
                dsqr = f32(f32(num * num) / den)
                if f64(dsqr) < 3.84 * f64(ls[o2]):
                    bi, best = i2, d
            if bi >= 0:
                m12[i1] = bi
                n += 1
                if check:
                    hist[_bin(kf1["angle"][i1], kf2["angle"][bi])].append(i1)
    if check:
        def clear(j):
            m12[j] = -3
        n -= _rot_prune(hist, clear)
    return n, m12


def _py_fuse(oracle, g, Tcw, th, lsf, ils, mps, kf):
    f32, f64 = np.float32, np.float64
    T = np.asarray(Tcw, f32).reshape(4, 4)
    R, t = T[:3, :3], T[:3, 3]
    # KeyFrame::SetPose: Rwc = Rcw.t(); Ow = -Rwc * tcw (small-matrix path, alpha = -1)
    Ow = np.array([-f32(f32(f32(R[0, r] * t[0]) + f32(R[1, r] * t[1])) + f32(R[2, r] * t[2])) for r in range(3)], f32)
    gg = {k: g[k][0] for k in g.dtype.names}
    n_mp = len(mps["valid"])
    bi, bd = np.full(n_mp, -1, np.int32), np.full(n_mp, 256, np.int32)
    nf = 0
    for i in range(n_mp):
        if not mps["valid"][i]:
            continue
        x = mps["world"][i].astype(f32)
        # cv::gemm small-matrix path: float row sum, then (float)(t0 + c) in double (test_cv_gemm_semantics)
        row = lambda r: f32(f64(f32(f32(f32(R[r, 0] * x[0]) + f32(R[r, 1] * x[1])) + f32(R[r, 2] * x[2]))) + f64(t[r]))
        p0, p1, p2 = row(0), row(1), row(2)
        if p2 < 0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = f32(f32(1) / p2)
            u = f32(f32(gg["fx"] * f32(p0 * invz)) + gg["cx"])
            v = f32(f32(gg["fy"] * f32(p1 * invz)) + gg["cy"])
        if not (u >= gg["mnMinX"] and u < gg["mnMaxX"] and v >= gg["mnMinY"] and v < gg["mnMaxY"]):
            continue
        ur = f32(u - f32(gg["mbf"] * invz))
        mind, maxd = mps["dist"][i].astype(f32)
        po = (x - Ow).astype(f32)
        d3 = f32(np.sqrt(sum(f64(p) * f64(p) for p in po)))
        if d3 < f32(0.8) * mind or d3 > f32(1.2) * maxd:
            continue
        if sum(f64(po[k]) * f64(mps["normal"][i][k]) for k in range(3)) < 0.5 * f64(d3):
            continue
        lvl = int(np.ceil(f32(np.log(f32(maxd / d3))) / f32(lsf)))
        lvl = min(max(lvl, 0), int(gg["nlevels"]) - 1)
        radius = f32(f32(th) * gg["scaleFactors"][lvl])
        best, b = 256, -1
        for j in oracle.features_in_area(g, kf["xy"], kf["octave"], float(u), float(v), float(radius)):
            kl = int(kf["octave"][j])
            if kl < lvl - 1 or kl > lvl:
                continue
            ex, ey = f32(u - kf["xy"][j][0]), f32(v - kf["xy"][j][1])
            if kf["uright"][j] >= 0:
                er = f32(ur - kf["uright"][j])
                e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                if f64(f32(e2 * ils[kl])) > 7.8:
                    continue
            else:
                e2 = f32(f32(ex * ex) + f32(ey * ey))
                if f64(f32(e2 * ils[kl])) > 5.99:
                    continue
            d = _pop(mps["desc"][i], kf["desc"][j])
            if d < best:
                best, b = d, int(j)
        bi[i], bd[i] = b, best
        if best <= 50:
            nf += 1
    return nf, bi, bd


def test_search_by_bow_vs_python(oracle):
    for seed, shuffle in ((31, False), (32, True)):
        kf, f = S.bow_scene(seed, n_kf=260, n_f=240, n_nodes=30, shuffle=shuffle)
        for check in (False, True):
            n, fm = oracle.search_by_bow(0.7, check, kf, f)
            n_py, fm_py = _py_bow(0.7, check, kf, f)
            assert n == n_py and np.array_equal(fm, fm_py)
            assert n > 40 and n == (fm >= 0).sum()
        # a Frame slot is matched at most once, and only by a valid KeyFrame keypoint of the same node
        node_of_f = {j: nd for nd, v in f["featvec"].items() for j in v}
        node_of_k = {i: nd for nd, v in kf["featvec"].items() for i in v}
        for j in np.nonzero(fm >= 0)[0]:
            assert kf["valid"][fm[j]] and node_of_f[j] == node_of_k[fm[j]]


def test_search_for_triangulation_vs_python(oracle):
    for seed in (41, 42):
        kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(seed, n=280, n_nodes=24)
        for only_stereo in (False, True):
            for check in (False, True):
                n, m = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, only_stereo, check, sf, ls, kf1, kf2)
                n_py, m_py = _py_triangulation(F12, Cw1, Tcw2, K2, only_stereo, check, sf, ls, kf1, kf2)
                assert n == n_py and np.array_equal(m, m_py)
                assert n == (m >= 0).sum()
        n, m = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, False, False, sf, ls, kf1, kf2)
        ok = np.nonzero(m >= 0)[0]
        assert len(ok) > 40 and (kf2["truth"][ok] == m[ok]).mean() > 0.9     # mostly the true correspondences
        assert not kf1["has_mp"][ok].any() and not kf2["has_mp"][m[ok]].any()


def test_fuse_search_vs_python(oracle):
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    for seed, th in ((51, 3.0), (52, 5.0)):
        mps, kf, Tcw, ils = S.fuse_scene(seed, n_mp=300, n_kf=260)
        g = frame_geom()
        n, bi, bd = oracle.fuse_search(g, Tcw, th, lsf, ils, mps, kf)
        n_py, bi_py, bd_py = _py_fuse(oracle, g, Tcw, th, lsf, ils, mps, kf)
        assert n == n_py and np.array_equal(bi, bi_py) and np.array_equal(bd, bd_py)
        assert n > 60 and n == (bd <= 50).sum() and ((bi >= 0) == (bd < 256)).all()
        assert (bi[mps["valid"] == 0] == -1).all()


def test_distinctive_descriptors_vs_numpy(oracle):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263): the oracle against a numpy restatement
    (full distance matrix, row sort, element int(0.5 (N-1)), first row with the least median)"""
    sets = S.observation_sets(61, n_points=150)
    bi, bm = oracle.distinctive_descriptors(sets)
    for k, d in enumerate(sets):
        if len(d) == 0:
            assert bi[k] == -1
            continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (len(d) - 1))]
        assert bi[k] == int(np.argmin(med)) and bm[k] == med.min(), k
    identical = [k for k, d in enumerate(sets) if len(d) > 1 and (d == d[0]).all()]
    assert identical and all(bi[k] == 0 and bm[k] == 0 for k in identical)


def test_golden_node_searches(oracle):
    gold = np.load(os.path.join(GOLD, "node_searches.npz"))
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))
    kf, f = S.bow_scene(3)
    n, fm = oracle.search_by_bow(0.7, True, kf, f)
    assert n == int(gold["bow_n"]) and np.array_equal(fm, gold["bow_match"])
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(3)
    n, m = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, kf1, kf2)
    assert n == int(gold["tri_n"]) and np.array_equal(m, gold["tri_match"])
    mps, kfs, Tcw, ils = S.fuse_scene(3)
    n, bi, bd = oracle.fuse_search(frame_geom(), Tcw, 3.0, lsf, ils, mps, kfs)
    assert n == int(gold["fuse_n"]) and np.array_equal(bi, gold["fuse_idx"]) and np.array_equal(bd, gold["fuse_dist"])
    ddi, ddm = oracle.distinctive_descriptors(S.observation_sets(3))
    assert np.array_equal(ddi, gold["dd_idx"]) and np.array_equal(ddm, gold["dd_median"])


# ---- golden vectors produced by the reference's own source (tests/golden/make_golden_ref.py, oracle/_ref/*.so)
def _ref_gold():
    return np.load(os.path.join(GOLD, "reference_source.npz"))


def _null3(a):
    a = a.copy()
    a[a == -3] = -1  # the reference stores NULL for "assigned, then reset by the rotation check"
    return a


def test_oracle_equals_reference_source_golden(oracle):
    """The oracle against outputs of /root/reference/src/{ORBextractor.cc, PlaneExtractor.cpp + peac, SurfelFusion.cpp,
    ORBmatcher.cc} compiled unmodified, committed as a fixture so that the comparison also runs
    where /root/reference does not exist."""
    sys.path.insert(0, GOLD)
    import make_golden_ref as G
    gold = _ref_gold()
    k, d = oracle.OrbOracle()(S.gray_frame(G.ORB_SEED))
    assert np.array_equal(k.view(np.uint8).reshape(len(k), -1), gold["orb_kps"]) and np.array_equal(d, gold["orb_desc"])
    d16, _ = S.depth_frame(G.PLANE_SEED)
    _, blocks, seed, edges = oracle.plane_prestage(d16, depth_map_factor=1.0)
    gb = gold["plane_blocks"].copy().view(oracle.BLOCK_DTYPE).reshape(-1)
    v = gb["N"] >= 4
    assert np.array_equal(blocks["N"], gb["N"]) and np.array_equal(blocks["nouse"], gb["nouse"])
    for f in ("center", "normal", "mse", "curvature"):
        assert np.array_equal(blocks[f][v].view(np.uint64), gb[f][v].view(np.uint64)), f
    assert np.array_equal(seed, gold["plane_seed"]) and np.array_equal(edges, gold["plane_edges"])
    g, dd, m, T, local = G.surfel_inputs()
    o = oracle.SurfelOracle()
    lo = local.copy()
    new = o.fuse(21, g, dd, m, T, lo)
    assert np.array_equal(o.index(), gold["surfel_index"].astype(np.int32))
    assert np.array_equal(o.seeds().view(np.uint8).reshape(-1, 72), gold["surfel_seeds"])
    assert np.array_equal(lo.view(np.uint8).reshape(len(lo), -1), gold["surfel_local_after"])
    assert np.array_equal(new.view(np.uint8).reshape(len(new), -1), gold["surfel_new"])
    # the reference's own peac membership image (trail counters <= -2 included) as SurfelFusion's input
    mem = gold["peac_membership"].astype(np.int32)
    assert mem.min() <= -2 and mem.max() >= 1
    lo2 = local.copy()
    _, d62 = S.depth_frame(G.PLANE_SEED)
    new2 = o.fuse(21, g, d62, mem, T, lo2)
    assert np.array_equal(o.index(), gold["surfel_peac_index"].astype(np.int32))
    assert np.array_equal(new2.view(np.uint8).reshape(len(new2), -1), gold["surfel_peac_new"])
    geom = frame_geom()
    cur, last, mps, Tc, Tl = S.match_scene(G.MATCH_SEED)
    n, cm = oracle.search_by_projection_frame(geom, Tc, Tl, 15.0, True, last, cur)
    assert n == int(gold["m_frame_n"]) and np.array_equal(_null3(cm), gold["m_frame"])
    n, cm = oracle.search_by_projection_points(geom, 3.0, 0.8, mps, cur)
    assert n == int(gold["m_points_n"]) and np.array_equal(_null3(cm), gold["m_points"])
    cur2, kf, Tc2 = S.reloc_scene(G.MATCH_SEED)
    n, cm = oracle.search_by_projection_keyframe(geom, Tc2, 15.0, 100, True, G.LSF, kf, cur2)
    assert n == int(gold["m_reloc_n"]) and np.array_equal(_null3(cm), gold["m_reloc"])
    kfb, f = S.bow_scene(G.MATCH_SEED)
    n, fm = oracle.search_by_bow(0.7, True, kfb, f)
    assert n == int(gold["m_bow_n"]) and np.array_equal(_null3(fm), gold["m_bow"])
    kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(G.MATCH_SEED)
    n, mm = oracle.search_for_triangulation(F12, Cw1, Tcw2, K2, False, True, sf, ls, kf1, kf2)
    assert n == int(gold["m_tri_n"]) and np.array_equal(_null3(mm), gold["m_tri"])
    mpf, kfs, Tcw, ils = S.fuse_scene(G.MATCH_SEED)
    n, bi, bd = oracle.fuse_search(geom, Tcw, 3.0, G.LSF, ils, mpf, kfs)
    assert n == int(gold["m_fuse_n"]) and np.array_equal(np.where(bd <= 50, bi, -1), gold["m_fuse"])


def test_distinctive_kernel_algorithm_replay(oracle):
    """k_distinctive (manhattanslam_b200/csrc/match.cu) replayed step by step in numpy -- four warps take rows i = w, w + 4, ...;
    the row median of rank (int)(0.5 * (N - 1)) found by bisection on the value with one warp-wide count per step; strict <
    inside a warp, (median, row) order across warps -- against the oracle's sort-based restatement of
    MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263).  The kernel itself has not run on a GPU yet."""
    def replay(desc):
        n = len(desc)
        if n == 0:
            return -1, 0x7fffffff
        bits = np.unpackbits(desc, axis=1).astype(np.int32)
        dist = (bits[:, None, :] != bits[None, :, :]).sum(2)  # what hamming256 recomputes per (row, j)
        r = int(0.5 * (n - 1))
        per_warp = []
        for w in range(4):
            my_med, my_row = 0x7fffffff, -1
            for i in range(w, n, 4):
                lo, hi = 0, 256
                while lo < hi:
                    mid = (lo + hi) >> 1
                    c = int((dist[i] <= mid).sum())  # sum over the 32 lanes' strided counts
                    if c > r:
                        hi = mid
                    else:
                        lo = mid + 1
                if lo < my_med:
                    my_med, my_row = lo, i
            per_warp.append((my_med, my_row))
        bm, br = 0x7fffffff, -1
        for med, row in per_warp:
            if row >= 0 and (br < 0 or med < bm or (med == bm and row < br)):
                bm, br = med, row
        return br, bm

    r = np.random.default_rng(3)
    sets = S.observation_sets(5, n_points=120, max_obs=40)
    sets += [r.integers(0, 256, (n, 32), dtype=np.uint8) for n in (1, 2, 3, 4, 5, 31, 32, 33, 64, 129)]
    base = r.integers(0, 256, 32, dtype=np.uint8)
    sets += [np.tile(base, (7, 1)), np.tile(base, (1, 1))]  # identical descriptors: every median 0, the first row wins
    bi, bm = oracle.distinctive_descriptors(sets)
    for k, d in enumerate(sets):
        assert replay(np.asarray(d, np.uint8).reshape(-1, 32)) == (int(bi[k]), int(bm[k])), k
