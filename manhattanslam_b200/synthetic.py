"""Seeded synthetic RGB-D inputs: identical bytes feed the CPU oracle and the
CUDA path in tests and bench.  numpy only."""
import numpy as np

K_DEFAULT = (525.0, 525.0, 319.5, 239.5)  # fx, fy, cx, cy (Example/TUM3.yaml-style, no distortion)


def gray_frame(seed, w=640, h=480):
    """128 base + 300 random axis-aligned rectangles (8-80 px, delta U(-60,60)) + N(0,2) noise, u8."""
    r = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.float32)
    for _ in range(300):
        rw, rh = r.integers(8, 81, 2)
        x = r.integers(0, w)
        y = r.integers(0, h)
        img[y:y + rh, x:x + rw] += r.uniform(-60, 60)
    img += r.normal(0, 2, (h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def gray_batch(seed0, batch, w=640, h=480):
    return np.stack([gray_frame(seed0 + i, w, h) for i in range(batch)])


def depth_frame(seed, w=640, h=480, K=K_DEFAULT, holes=True, scene=None):
    """Piecewise-planar scene: 3-6 random planes (normals within 60 deg of -z, 0.8-4 m) z-buffered
    through K, + N(0,(1.5e-3 z^2)) noise, 2 % zero holes in 8x8 patches, quantised to u16 (factor 5000).
    scene=None: planes, noise and holes all come from `seed` (independent frames).  scene=s: the planes come
    from seed s and only the sensor noise / holes from `seed` -- consecutive frames of ONE scene, which is what
    a keyframe stream into a local surfel map looks like.
    Returns (depth_u16, depth_f32_metres)."""
    r = np.random.default_rng((seed if scene is None else scene) + 100003)
    fx, fy, cx, cy = [k * (w / 640.0) for k in K]
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    rx, ry = (u - cx) / fx, (v - cy) / fy
    z = np.full((h, w), np.inf)
    for _ in range(int(r.integers(3, 7))):
        ang = np.deg2rad(r.uniform(0, 60))
        az = r.uniform(0, 2 * np.pi)
        n = np.array([np.sin(ang) * np.cos(az), np.sin(ang) * np.sin(az), -np.cos(ang)])
        d0 = r.uniform(0.8, 4.0)
        # plane through (x0, y0, d0) on a random ray
        px, py = r.uniform(-0.5, 0.5), r.uniform(-0.4, 0.4)
        p0 = np.array([px * d0, py * d0, d0])
        denom = n[0] * rx + n[1] * ry + n[2]
        with np.errstate(divide="ignore", invalid="ignore"):
            zz = (n @ p0) / denom
        zz[(zz < 0.4) | ~np.isfinite(zz)] = np.inf
        # each plane covers a random half-plane/rectangle so that several stay visible
        x0, x1 = sorted(r.integers(0, w, 2))
        y0, y1 = sorted(r.integers(0, h, 2))
        if x1 - x0 < w // 3:
            x0, x1 = 0, w
        if y1 - y0 < h // 3:
            y0, y1 = 0, h
        m = np.zeros((h, w), bool)
        m[y0:y1, x0:x1] = True
        zz[~m] = np.inf
        z = np.minimum(z, zz)
    z[~np.isfinite(z)] = 4.5
    if scene is not None:
        r = np.random.default_rng(seed + 200003)
    z = z + r.normal(0, 1.0, (h, w)) * (1.5e-3 * z * z)
    if holes:
        nh = int(0.02 * (w // 8) * (h // 8))
        for _ in range(nh):
            x = int(r.integers(0, w // 8)) * 8
            y = int(r.integers(0, h // 8)) * 8
            z[y:y + 8, x:x + 8] = 0
    d16 = np.clip(np.rint(z * 5000.0), 0, 65535).astype(np.uint16)
    # Tracking::GrabImage: imDepth.convertTo(CV_32F, mDepthMapFactor) with factor 1/5000 (src/Tracking.cc:133-137,205-207)
    dep = (d16.astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
    return d16, dep


def membership(seed, w=640, h=480, plane_fraction=0.0):
    """Half-resolution CV_32SC1 plane membership image: -1 = no plane.  plane_fraction of the pixels
    (as 20x20 half-res patches) are assigned plane id 0."""
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    m = np.full((h2, w2), -1, np.int32)
    if plane_fraction > 0:
        r = np.random.default_rng(seed + 7)
        n = int(plane_fraction * (w2 // 20) * (h2 // 20))
        for _ in range(n):
            x = int(r.integers(0, w2 // 20)) * 20
            y = int(r.integers(0, h2 // 20)) * 20
            m[y:y + 20, x:x + 20] = 0
    return m


def pose_walk(seed, n):
    """Twc random walk (<= 2 cm, <= 1 deg per frame) from identity, float32 row-major 4x4."""
    r = np.random.default_rng(seed + 31)
    T = np.eye(4)
    out = []
    for _ in range(n):
        ax = r.normal(size=3)
        ax /= np.linalg.norm(ax)
        a = np.deg2rad(r.uniform(0, 1.0))
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx
        dT = np.eye(4)
        dT[:3, :3] = R
        dT[:3, 3] = r.uniform(-0.02, 0.02, 3) / np.sqrt(3)
        T = T @ dT
        out.append(T.astype(np.float32))
    return np.stack(out)


SURFEL_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                         ("size", "<f4"), ("color", "<f4"), ("r", "<i4"), ("g", "<i4"), ("b", "<i4"),
                         ("weight", "<f4"), ("updateTimes", "<i4"), ("lastUpdate", "<i4")])


def surfel_map(seed, n, depth_f32, Twc, K=K_DEFAULT, ref_index=100, w=640, h=480):
    """n surfels laid out the way a real local map is: the back-projected superpixel seeds
    (80x60 raster, ~4800 each) of ceil(n/4800) synthetic keyframes whose views overlap the current one, keyframe
    after keyframe -- so neighbours in memory are neighbours in space.  ~35 % project into the frame and agree
    with the depth, ~10 % hit the unstable-drop rule (old lastUpdate, few updates), the rest are out of view."""
    r = np.random.default_rng(seed + 1234)
    fx, fy, cx, cy = K
    s = np.zeros(n, SURFEL_DTYPE)
    spw, sph = w // 8, h // 8
    per = spw * sph
    nkf = (n + per - 1) // per
    col = np.tile(np.arange(spw), sph).astype(np.float64)
    row = np.repeat(np.arange(sph), spw).astype(np.float64)
    ou = r.uniform(-0.9 * w, 0.9 * w, nkf)
    ov = r.uniform(-0.9 * h, 0.9 * h, nkf)
    sc = r.uniform(0.8, 1.25, nkf)
    u = (ou[:, None] + sc[:, None] * (8 * col[None, :] + 4)).reshape(-1)[:n] + r.uniform(-2, 2, n)
    v = (ov[:, None] + sc[:, None] * (8 * row[None, :] + 4)).reshape(-1)[:n] + r.uniform(-2, 2, n)
    ui = np.clip(np.rint(u), 1, w - 2).astype(np.int64)
    vi = np.clip(np.rint(v), 1, h - 2).astype(np.int64)
    z = depth_f32[vi, ui].astype(np.float64)
    z = np.where(z > 0.1, z, r.uniform(0.8, 4.0, n))
    z = z + r.normal(0, 0.01, n) + (r.random(n) < 0.05) * r.uniform(-1.5, 1.5, n)
    z = np.maximum(z, 0.3)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z, np.ones(n)], 1)
    pw = pc @ Twc.astype(np.float64).T
    s["px"], s["py"], s["pz"] = pw[:, 0], pw[:, 1], pw[:, 2]
    nc = np.stack([r.normal(0, 0.25, n), r.normal(0, 0.25, n), -np.ones(n)], 1)
    nc /= np.linalg.norm(nc, axis=1, keepdims=True)
    nw = nc @ Twc[:3, :3].astype(np.float64).T
    s["nx"], s["ny"], s["nz"] = nw[:, 0], nw[:, 1], nw[:, 2]
    s["size"] = r.uniform(0.005, 0.05, n)
    s["color"] = r.uniform(0, 255, n)
    s["r"] = r.integers(0, 256, n)
    s["g"] = r.integers(0, 256, n)
    s["b"] = r.integers(0, 256, n)
    s["weight"] = r.uniform(1, 20, n)
    s["updateTimes"] = r.integers(1, 31, n)
    s["lastUpdate"] = ref_index - r.integers(0, 9, n)
    return s


def match_scene(seed, n_cur=1000, n_last=900, w=640, h=480, K=K_DEFAULT, collide=0.3):
    """Synthetic tracking situation for ORBmatcher::SearchByProjection: current-frame keypoints with
    random descriptors, and last-frame map points that project (through Tcw_cur) near a chosen current
    keypoint with a noisy copy of its descriptor.  `collide` = fraction of queries that target a
    keypoint also targeted by another query (exercises the Observations()>0 slot blocking)."""
    r = np.random.default_rng(seed + 555)
    fx, fy, cx, cy = K
    cur = {"xy": np.stack([r.uniform(2, w - 2, n_cur), r.uniform(2, h - 2, n_cur)], 1).astype(np.float32),
           "octave": r.integers(0, 8, n_cur).astype(np.int32),
           "angle": r.uniform(0, 360, n_cur).astype(np.float32),
           "uright": None,
           "desc": r.integers(0, 256, (n_cur, 32), dtype=np.uint8),
           "occupied": (r.random(n_cur) < 0.05).astype(np.uint8)}
    # poses: small motion between last and current
    T = pose_walk(seed, 2)
    Twc_last, Twc_cur = T[0].astype(np.float64), T[1].astype(np.float64)
    Twc_cur[2, 3] += r.choice([-0.3, 0.0, 0.3])  # exercises bForward / bBackward / neither
    Tcw_cur = np.linalg.inv(Twc_cur).astype(np.float32)
    Tcw_last = np.linalg.inv(Twc_last).astype(np.float32)
    tgt = r.integers(0, n_cur, n_last)
    ncol = int(collide * n_last)
    tgt[:ncol] = tgt[ncol:2 * ncol]
    r.shuffle(tgt)
    cur_z = r.uniform(1.0, 5.0, n_cur)
    cur["uright"] = np.where(r.random(n_cur) < 0.7, cur["xy"][:, 0] - 40.0 / cur_z, -1).astype(np.float32)
    z = cur_z[tgt] * (1.0 + r.normal(0, 0.01, n_last))
    uv = cur["xy"][tgt].astype(np.float64) + r.normal(0, 2.0, (n_last, 2))
    pc = np.stack([(uv[:, 0] - cx) / fx * z, (uv[:, 1] - cy) / fy * z, z, np.ones(n_last)], 1)
    pw = (np.linalg.inv(Tcw_cur.astype(np.float64)) @ pc.T).T[:, :3]
    desc = cur["desc"][tgt].copy()
    flips = r.integers(0, 70, n_last)
    for i in range(n_last):
        bits = r.choice(256, flips[i], replace=False)
        for b in bits:
            desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    last = {"has_mp": (r.random(n_last) < 0.9).astype(np.uint8), "outlier": (r.random(n_last) < 0.05).astype(np.uint8),
            "mp_obs": (r.random(n_last) < 0.85).astype(np.uint8), "mp_world": pw.astype(np.float32), "mp_desc": desc,
            "octave": np.clip(cur["octave"][tgt] + r.integers(-1, 2, n_last), 0, 7).astype(np.int32),
            "angle": ((cur["angle"][tgt] + r.normal(0, 8, n_last) + (r.random(n_last) < 0.1) * 90) % 360).astype(np.float32)}
    # map-point flavour (SearchLocalPoints): projections already computed by Frame::isInFrustum
    invz = 1.0 / z
    mps = {"valid": (r.random(n_last) < 0.9).astype(np.uint8), "obs": last["mp_obs"],
           "proj_xyr": np.stack([uv[:, 0], uv[:, 1], uv[:, 0] - 40.0 * invz], 1).astype(np.float32),
           "level": last["octave"], "viewcos": r.uniform(0.9, 1.0, n_last).astype(np.float32), "desc": desc}
    return cur, last, mps, Tcw_cur, Tcw_last


def reloc_scene(seed, n_cur=1000, n_kf=900, collide=0.3, negative_depth=0.02):
    """Relocalisation flavour of match_scene for ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist)
    (src/ORBmatcher.cc:680-797): KeyFrame map points with scale-invariance distances (mfMinDistance, mfMaxDistance)
    chosen so that MapPoint::PredictScale lands near the target keypoint's octave; a few points lie outside their
    distance range or behind the camera (that overload has no invzc<0 test)."""
    cur, last, _, Tcw_cur, _ = match_scene(seed, n_cur, n_kf, collide=collide)
    r = np.random.default_rng(seed + 777)
    Twc = np.linalg.inv(Tcw_cur.astype(np.float64))
    Ow = Twc[:3, 3]
    pw = last["mp_world"].astype(np.float64).copy()
    nneg = int(negative_depth * n_kf)
    if nneg:  # mirror a few points through the camera centre: z < 0 but the projection stays in the image
        pw[:nneg] = 2 * Ow[None, :] - pw[:nneg]
    dist = np.linalg.norm(pw - Ow[None, :], axis=1)
    lvl = np.clip(last["octave"] + r.integers(-1, 2, n_kf), 0, 7)
    maxd = dist * 1.2 ** (lvl - r.uniform(0.05, 0.95, n_kf))       # ceil(log(max/dist)/log 1.2) == lvl
    mind = maxd / 1.2 ** 7
    far = r.random(n_kf) < 0.05
    maxd = np.where(far, dist / 1.3, maxd)                           # outside [0.8 min, 1.2 max]
    kf = {"valid": (r.random(n_kf) < 0.9).astype(np.uint8), "mp_world": pw.astype(np.float32), "mp_desc": last["mp_desc"],
          "mp_dist": np.stack([mind, maxd], 1).astype(np.float32), "angle": last["angle"]}
    cur = dict(cur)
    cur["occupied"] = (r.random(n_cur) < 0.1).astype(np.uint8)      # mvpMapPoints[j] != NULL on entry
    return cur, kf, Tcw_cur


def _noisy_desc(r, desc, max_flips):
    """copies of 32-byte descriptors with 0..max_flips-1 random bits flipped per row"""
    out = desc.copy()
    flips = r.integers(0, max_flips, len(desc))
    for i in range(len(desc)):
        for b in r.choice(256, flips[i], replace=False):
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def _feature_vector(node_of_feature, order_rng=None):
    """DBoW2::FeatureVector as a dict node id -> list of feature indices (ascending, as FeatureVector::addFeature builds
    it when features are transformed in index order; shuffled inside a node if order_rng is given)."""
    fv = {}
    for i, nd in enumerate(node_of_feature):
        fv.setdefault(int(nd), []).append(i)
    if order_rng is not None:
        for v in fv.values():
            order_rng.shuffle(v)
    return fv


def bow_scene(seed, n_kf=1000, n_f=1000, n_nodes=120, collide=0.3, shuffle=False):
    """Synthetic ORBmatcher::SearchByBoW situation (src/ORBmatcher.cc:146-255): a Frame with random descriptors spread
    over vocabulary nodes (sparse 32-bit node ids), and a KeyFrame most of whose keypoints are noisy copies of a Frame
    keypoint in the same node.  `collide` = fraction of KeyFrame keypoints that target a Frame keypoint another one
    targets too (exercises the "slot already matched" skip); some nodes exist on one side only."""
    r = np.random.default_rng(seed + 4242)
    ids = np.sort(r.choice(100000, n_nodes + 20, replace=False)).astype(np.uint32)
    f_nodes = ids[r.integers(0, n_nodes, n_f)]                       # the last 20 ids never occur in the Frame
    f = {"desc": r.integers(0, 256, (n_f, 32), dtype=np.uint8), "angle": r.uniform(0, 360, n_f).astype(np.float32)}
    tgt = r.integers(0, n_f, n_kf)
    ncol = int(collide * n_kf) // 2
    tgt[:ncol] = tgt[ncol:2 * ncol]
    r.shuffle(tgt)
    linked = r.random(n_kf) < 0.8
    kf_nodes = np.where(linked, f_nodes[tgt], ids[r.integers(10, n_nodes + 20, n_kf)])
    desc = _noisy_desc(r, f["desc"][tgt], 70)
    desc[~linked] = r.integers(0, 256, (int((~linked).sum()), 32), dtype=np.uint8)
    # exact duplicates of a Frame descriptor in the same node: best == second best, the ratio test must reject
    dup = r.choice(n_f, 10, replace=False)
    for d in dup:
        same = np.nonzero(f_nodes == f_nodes[d])[0]
        if len(same) > 1:
            f["desc"][same[0]] = f["desc"][d]
    ang = (f["angle"][tgt] + r.normal(0, 6, n_kf) + (r.random(n_kf) < 0.1) * 90) % 360
    kf = {"valid": (r.random(n_kf) < 0.9).astype(np.uint8), "desc": desc, "angle": ang.astype(np.float32),
          "featvec": _feature_vector(kf_nodes, r if shuffle else None)}
    f["featvec"] = _feature_vector(f_nodes, r if shuffle else None)
    return kf, f


def triangulation_scene(seed, n=900, n_nodes=100, K=K_DEFAULT, w=640, h=480):
    """Two KeyFrames for ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:257-406): 3-D points seen by both, the
    fundamental matrix F12 = K1^-T [t12]x R12 K2^-1 (LocalMapping::ComputeF12, src/LocalMapping.cc), keypoints =
    projections + noise, descriptors = noisy copies, a few exact duplicates in KeyFrame 2 (equal distances: the later
    candidate wins, :336), mixed stereo / monocular keypoints, some keypoints with MapPoints already.
    Returns (kf1, kf2, F12, Cw1, Tcw2, K2, scale_factors, level_sigma2)."""
    r = np.random.default_rng(seed + 9090)
    fx, fy, cx, cy = K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    T = pose_walk(seed + 3, 2)
    Twc1 = T[0].astype(np.float64)
    Twc2 = T[1].astype(np.float64)
    Twc2[:3, 3] += np.array([0.25, 0.02, 0.05])  # a baseline, as between neighbouring keyframes
    Tcw1, Tcw2 = np.linalg.inv(Twc1), np.linalg.inv(Twc2)
    uv1 = np.stack([r.uniform(20, w - 20, n), r.uniform(20, h - 20, n)], 1)
    z = r.uniform(1.0, 6.0, n)
    pc1 = np.stack([(uv1[:, 0] - cx) / fx * z, (uv1[:, 1] - cy) / fy * z, z, np.ones(n)], 1)
    pw = (Twc1 @ pc1.T).T
    pc2 = (Tcw2 @ pw.T).T
    uv2 = np.stack([fx * pc2[:, 0] / pc2[:, 2] + cx, fy * pc2[:, 1] / pc2[:, 2] + cy], 1) + r.normal(0, 0.7, (n, 2))
    bad = r.random(n) < 0.15                      # off the epipolar line
    uv2[bad] += r.normal(0, 25, (int(bad.sum()), 2))
    R12 = Tcw1[:3, :3] @ Tcw2[:3, :3].T
    t12 = -R12 @ Tcw2[:3, 3] + Tcw1[:3, 3]
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(Kmat).T @ tx @ R12 @ np.linalg.inv(Kmat)).astype(np.float32)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    ids = np.sort(r.choice(50000, n_nodes, replace=False)).astype(np.uint32)
    nodes1 = ids[r.integers(0, n_nodes, n)]
    nodes2 = np.where(r.random(n) < 0.85, nodes1, ids[r.integers(0, n_nodes, n)])
    desc1 = r.integers(0, 256, (n, 32), dtype=np.uint8)
    desc2 = _noisy_desc(r, desc1, 60)
    oct2 = r.integers(0, 8, n).astype(np.int32)
    ang1 = r.uniform(0, 360, n)
    ang2 = (ang1 + r.normal(0, 6, n) + (r.random(n) < 0.1) * 120) % 360
    # duplicates: keypoint j+1 of KeyFrame 2 repeats keypoint j (same node, same descriptor, almost the same place)
    for j in r.choice(n - 1, 25, replace=False):
        nodes2[j + 1], desc2[j + 1], uv2[j + 1], oct2[j + 1] = nodes2[j], desc2[j], uv2[j] + 0.1, oct2[j]
    kf1 = {"has_mp": (r.random(n) < 0.3).astype(np.uint8), "uright": np.where(r.random(n) < 0.6, uv1[:, 0] - 40.0 / z, -1).astype(np.float32),
           "xy": uv1.astype(np.float32), "angle": ang1.astype(np.float32), "desc": desc1, "featvec": _feature_vector(nodes1)}
    perm = r.permutation(n)                      # KeyFrame 2 lists its keypoints in another order
    inv = np.argsort(perm)
    kf2 = {"has_mp": (r.random(n) < 0.3).astype(np.uint8), "uright": np.where(r.random(n) < 0.6, uv2[perm, 0] - 40.0 / pc2[perm, 2], -1).astype(np.float32),
           "xy": uv2[perm].astype(np.float32), "octave": oct2[perm], "angle": ang2[perm].astype(np.float32), "desc": desc2[perm],
           "featvec": _feature_vector(nodes2[perm]), "truth": inv}
    Cw1 = Twc1[:3, 3].astype(np.float32)
    return kf1, kf2, F12, Cw1, Tcw2.astype(np.float32), np.asarray(K, np.float32), sf, (sf * sf).astype(np.float32)


def fuse_scene(seed, n_mp=1200, n_kf=1000, K=K_DEFAULT, w=640, h=480):
    """ORBmatcher::Fuse (src/ORBmatcher.cc:408-546): a KeyFrame's keypoints and map points of its neighbours that
    project near them.  Covers: points behind the camera / outside the image / outside their scale-invariance range /
    seen at more than 60 degrees, stereo and monocular keypoints (two chi-square gates), level window.
    Returns (mps, kf, Tcw, inv_level_sigma2)."""
    r = np.random.default_rng(seed + 6060)
    fx, fy, cx, cy = K
    sf = 1.2 ** np.arange(8)
    Twc = pose_walk(seed + 11, 1)[0].astype(np.float64)
    Tcw = np.linalg.inv(Twc)
    xy = np.stack([r.uniform(2, w - 2, n_kf), r.uniform(2, h - 2, n_kf)], 1)
    octv = r.integers(0, 8, n_kf)
    zk = r.uniform(1.0, 6.0, n_kf)
    kf = {"xy": xy.astype(np.float32), "octave": octv.astype(np.int32),
          "uright": np.where(r.random(n_kf) < 0.6, xy[:, 0] - 40.0 / zk, -1).astype(np.float32),
          "desc": r.integers(0, 256, (n_kf, 32), dtype=np.uint8)}
    tgt = r.integers(0, n_kf, n_mp)
    z = zk[tgt] * (1.0 + r.normal(0, 0.004, n_mp))
    uv = xy[tgt] + r.normal(0, 0.8, (n_mp, 2)) * sf[octv[tgt]][:, None]
    far_off = r.random(n_mp) < 0.08
    uv[far_off] += r.normal(0, 60, (int(far_off.sum()), 2))       # some leave the image
    pc = np.stack([(uv[:, 0] - cx) / fx * z, (uv[:, 1] - cy) / fy * z, z, np.ones(n_mp)], 1)
    behind = r.random(n_mp) < 0.03
    pc[behind, :3] *= -1
    pw = (Twc @ pc.T).T[:, :3]
    Ow = Twc[:3, 3]
    dist = np.linalg.norm(pw - Ow, axis=1)
    lvl = np.clip(octv[tgt] + r.integers(0, 2, n_mp), 0, 7)       # predicted level = keypoint level or one above
    maxd = dist * 1.2 ** (lvl - r.uniform(0.05, 0.95, n_mp))
    mind = maxd / 1.2 ** 7
    out_of_range = r.random(n_mp) < 0.05
    maxd = np.where(out_of_range, dist / 1.3, maxd)
    normal = (pw - Ow) / dist[:, None] + r.normal(0, 0.25, (n_mp, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    oblique = r.random(n_mp) < 0.08
    normal[oblique] = -normal[oblique]                              # PO . Pn < 0.5 dist
    mps = {"valid": (r.random(n_mp) < 0.9).astype(np.uint8), "world": pw.astype(np.float32), "normal": normal.astype(np.float32),
           "dist": np.stack([mind, maxd], 1).astype(np.float32), "desc": _noisy_desc(r, kf["desc"][tgt], 80)}
    inv_sigma2 = (1.0 / (sf * sf)).astype(np.float32)
    return mps, kf, Tcw.astype(np.float32), inv_sigma2


def observation_sets(seed, n_points=400, max_obs=40):
    """Descriptor sets of map points for MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:210-263): every point has
    1..max_obs observations that are noisy copies of one descriptor (a few outliers), some points have none, a few have
    many (more than a warp / more than a tile), some consist of identical descriptors (all medians tie: the first wins)."""
    r = np.random.default_rng(seed + 7171)
    out = []
    for k in range(n_points):
        n = int(r.integers(1, max_obs + 1))
        if k % 37 == 5:
            n = 0
        elif k % 53 == 7:
            n = int(r.integers(100, 260))
        base = r.integers(0, 256, (1, 32), dtype=np.uint8)
        d = _noisy_desc(r, np.repeat(base, n, 0), 60) if n else np.zeros((0, 32), np.uint8)
        if n and k % 11 == 3:
            d[:] = base                      # every row identical
        if n > 3 and k % 5 == 0:
            d[r.integers(0, n)] = r.integers(0, 256, 32, dtype=np.uint8)   # an outlier observation
        out.append(d)
    return out
