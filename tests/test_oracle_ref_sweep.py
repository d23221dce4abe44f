"""Randomised sweeps of the oracle against the REFERENCE'S OWN SOURCE (oracle/_ref, see tests/test_oracle_ref.py): random
image sizes, extractor / camera / matcher parameters and scene shapes drawn from fixed seeds -- the corners the
hand-picked cases of test_oracle_ref.py do not visit.  Everything bit for bit.  Two inputs on which the reference itself is
undefined came out of these sweeps and are excluded (and refused by the product): a pyramid level more than twice as tall
as wide (zero octree roots: division by zero and an index into an empty vector, src/ORBextractor.cc:535-560), and
ORBdist >= 256 in the relocalisation search (a query without candidates then writes mvpMapPoints[-1], src/ORBmatcher.cc:
725-760)."""
import numpy as np
import pytest

from manhattanslam_b200 import synthetic as S


def _need(oracle, name):
    if oracle.build_ref(name=name) is None:
        pytest.skip("oracle/_ref/%s not built and /root/reference absent" % name)
    return oracle


def test_orb_sweep(oracle):
    B = _need(oracle, "liborb_ref.so")
    r = np.random.default_rng(123)
    done = 0
    for _ in range(60):
        w, h = int(r.integers(120, 900)), int(r.integers(100, 700))
        nf, sf = int(r.choice([50, 200, 500, 1000, 2000, 4000])), float(r.choice([1.1, 1.2, 1.3, 1.5, 2.0]))
        nl, ini, mn = int(r.integers(1, 10)), int(r.choice([5, 10, 20, 40, 80])), int(r.choice([2, 5, 7, 20]))
        mn = min(mn, ini)
        kind, seed = int(r.integers(0, 4)), int(r.integers(0, 1000))
        if min(w, h) / sf ** (nl - 1) < 60 or w < 0.8 * h:
            continue
        if kind == 0:
            img = S.gray_frame(seed, w, h)
        elif kind == 1:
            img = r.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 2:
            img = np.full((h, w), 100, np.uint8)
            for _k in range(20):
                x, y = r.integers(20, w - 20), r.integers(20, h - 20)
                img[y:y + 4, x:x + 4] += np.uint8(r.integers(5, 40))
        else:
            img = (S.gray_frame(seed, w, h) // 8 * 8).astype(np.uint8)  # plateaus: many equal FAST scores
        ko, do = B.OrbOracle(nf, sf, nl, ini, mn)(img)
        kr, dr = B.RefOrbExtractor(nf, sf, nl, ini, mn)(img)
        assert ko.tobytes() == kr.tobytes() and np.array_equal(do, dr), (w, h, nf, sf, nl, ini, mn, kind)
        done += 1
    assert done >= 30


def test_orb_undefined_geometry_is_reported(oracle):
    with pytest.raises(ValueError):
        oracle.OrbOracle(1000, 1.1, 9, 10, 5)(S.gray_frame(1, 132, 509))


def test_plane_and_peac_sweep(oracle):
    B = _need(oracle, "libplane_ref.so")
    r = np.random.default_rng(7)
    for _ in range(30):
        w = int(r.integers(10, 33)) * 20 + int(r.choice([0, 2, 6, 14]))
        h = int(r.integers(8, 25)) * 20 + int(r.choice([0, 2, 10]))
        sc = w / 640.0
        K = (525.0 * sc * float(r.uniform(0.8, 1.2)), 525.0 * sc * float(r.uniform(0.8, 1.2)), (w - 1) / 2 + float(r.uniform(-5, 5)),
             (h - 1) / 2 + float(r.uniform(-5, 5)))
        d16, _d = S.depth_frame(int(r.integers(0, 10000)), w, h, K=K, holes=bool(r.integers(0, 2)))
        if r.random() < 0.3:
            d16 = (d16 // 4 * 4).astype(np.uint16)
        fac = float(r.choice([1.0, 1 / 5000., 0.2, 1 / 1000.]))
        co, bo, so, eo = B.plane_prestage(d16, K=K, depth_map_factor=fac)
        cr, br, sr, er = B.ref_plane_prestage(d16, K=K, depth_map_factor=fac)
        v = br["N"] >= 4
        assert co.tobytes() == cr.tobytes() and np.array_equal(bo["N"], br["N"]) and np.array_equal(so, sr) and np.array_equal(eo, er)
        for f in ("center", "normal", "mse", "curvature"):
            assert np.array_equal(bo[f][v].view(np.uint64), br[f][v].view(np.uint64)), (f, w, h, fac)
        mo, po = B.plane_detect(d16, K=K, depth_map_factor=fac)
        mr, pr = B.ref_plane_run(d16, K=K, depth_map_factor=fac)
        assert np.array_equal(mo, mr) and np.array_equal(po["N"], pr["N"]) and np.array_equal(po["vertices"], pr["vertices"]), (w, h, fac)
        assert po["normal"].tobytes() == pr["normal"].tobytes() and po["center"].tobytes() == pr["center"].tobytes()


def _same_rec(a, b):
    if a.shape != b.shape:
        return False
    for f in a.dtype.names:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            if not ((x.view(np.uint32) == y.view(np.uint32)) | (np.isnan(x) & np.isnan(y))).all():
                return False
        elif not np.array_equal(x, y):
            return False
    return True


def test_surfel_fusion_sweep(oracle):
    B = _need(oracle, "libsurfel_ref.so")
    r = np.random.default_rng(17)
    for _ in range(10):
        w, h = [(640, 480), (320, 240), (328, 248), (480, 360)][int(r.integers(0, 4))]
        sc = w / 640.0
        K = (525.0 * sc * float(r.uniform(0.8, 1.2)), 525.0 * sc * float(r.uniform(0.8, 1.2)), (w - 1) / 2 + float(r.uniform(-5, 5)),
             (h - 1) / 2 + float(r.uniform(-5, 5)))
        far, near = float(r.choice([3.0, 5.0, 30.0])), float(r.choice([0.1, 0.5, 1.0]))
        seed = int(r.integers(0, 10000))
        g = S.gray_frame(seed, w, h)
        _d16, d = S.depth_frame(seed, w, h, K=K)
        m = S.membership(seed, w, h, plane_fraction=float(r.choice([0, 0.2, 0.6])))
        T = S.pose_walk(seed, 3)[int(r.integers(0, 3))]
        n = int(r.choice([0, 1, 500, 20000]))
        local = S.surfel_map(seed, n, d, T, K=K, ref_index=9, w=w, h=h) if n else np.zeros(0, B.SURFEL_DTYPE)
        lo, lr = local.copy(), local.copy()
        o = B.SurfelOracle(w, h, K[0], K[1], K[2], K[3], far, near)
        rr = B.RefSurfelFusion(w, h, K[0], K[1], K[2], K[3], far, near)
        ref_idx = int(r.integers(9, 20))
        no, nr = o.fuse(ref_idx, g, d, m, T, lo), rr.fuse(ref_idx, g, d, m, T, lr)
        assert np.array_equal(o.index(), rr.index()) and _same_rec(o.seeds(), rr.seeds()), (w, h, far, near, n)
        assert _same_rec(lo, lr) and _same_rec(no, nr), (w, h, far, near, n)


def test_matcher_sweep(oracle):
    from manhattanslam_b200.matcher import frame_geom
    B = _need(oracle, "libmatch_ref.so")
    r = np.random.default_rng(11)
    g = frame_geom()
    lsf = float(np.float32(np.log(np.float64(np.float32(1.2)))))

    def null(a):
        a = a.copy()
        a[a == -3] = -1
        return a

    def both(fn, *args):
        a = fn(*args)
        with B.reference_matcher():
            b = fn(*args)
        return a[0] == b[0] and np.array_equal(null(a[1]), b[1])

    for _ in range(25):
        seed, nc, nl = int(r.integers(0, 100000)), int(r.choice([60, 200, 400, 1000])), int(r.choice([60, 150, 300, 900]))
        col, th, chk = float(r.choice([0, 0.3, 0.5])), float(r.choice([1.0, 3.0, 7.0, 15.0, 40.0])), bool(r.integers(0, 2))
        ratio, od, nn = float(r.choice([0.5, 0.6, 0.8, 0.95])), int(r.choice([30, 50, 64, 100, 255])), int(r.choice([5, 30, 120]))
        tag = (seed, nc, nl, col, th, chk, ratio, od, nn)
        cur, last, mps, Tc, Tl = S.match_scene(seed, nc, nl, collide=col)
        assert both(B.search_by_projection_frame, g, Tc, Tl, th, chk, last, cur), tag
        assert both(B.search_by_projection_points, g, th, ratio, mps, cur), tag
        cur2, kf, Tc2 = S.reloc_scene(seed, nc, nl, collide=col)
        assert both(B.search_by_projection_keyframe, g, Tc2, th, od, chk, lsf, kf, cur2), tag
        kfb, f = S.bow_scene(seed, nc, nl, n_nodes=nn, collide=col, shuffle=bool(r.integers(0, 2)))
        assert both(B.search_by_bow, ratio, chk, kfb, f), tag
        kf1, kf2, F12, Cw1, Tcw2, K2, sf, ls = S.triangulation_scene(seed, n=nl, n_nodes=nn)
        assert both(B.search_for_triangulation, F12, Cw1, Tcw2, K2, bool(r.integers(0, 2)), chk, sf, ls, kf1, kf2), tag
        mpf, kfs, Tcw, ils = S.fuse_scene(seed, n_mp=nl, n_kf=nc)
        n, bi, bd = B.fuse_search(g, Tcw, th, lsf, ils, mpf, kfs)
        nr, fi = B.ref_fuse(g, Tcw, th, lsf, ils, mpf, kfs)
        assert n == nr and np.array_equal(np.where(bd <= 50, bi, -1), fi), tag
