"""Seeded synthetic correspondence sets for the RANSAC hot path.

No datasets are available offline, so every workload in BASELINE.json is
generated here: 2D-3D sets for PnPsolver / MLPnPsolver (reference
src/PnPsolver.cpp:11-55 and src/MLPnPsolver.cpp:5-53 define the fields a solver
snapshots from a Frame) and 3D-3D sets for Sim3Solver (src/Sim3Solver.cpp:6-85).

Conventions
-----------
* camera: EuRoC intrinsics (reference Examples/Stereo/EuRoC.yaml), 752x480
* seed of problem `i` of config `cfg`: 1000*cfg + i  (numpy default_rng)
* octave of a match: 0..7 with probability ~ 1.2**(-2*level);
  sigma2 = (1.2**level)**2 evaluated as f32 products like
  reference src/ORBextractor.cpp:352-359
* inlier pixel noise N(0, sigma2 px^2); outliers: pixel uniform in the image
  (PnP/MLPnP) or the second 3-D point replaced by an unrelated frustum sample (Sim3)
"""
from __future__ import annotations

import numpy as np

EUROC = dict(fx=435.2046959714599, fy=435.2046959714599, cx=367.4517211914062, cy=252.2008514404297,
             width=752, height=480)
KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, width=1241, height=376)

N_LEVELS = 8
SCALE_FACTOR = np.float32(1.2)


def level_sigma2() -> np.ndarray:
    """mvLevelSigma2 as ORBextractor builds it: f32 running products (ORBextractor.cpp:352-359)."""
    sf = np.empty(N_LEVELS, np.float32)
    s2 = np.empty(N_LEVELS, np.float32)
    sf[0] = 1.0
    s2[0] = 1.0
    for i in range(1, N_LEVELS):
        sf[i] = np.float32(sf[i - 1] * SCALE_FACTOR)
        s2[i] = np.float32(sf[i] * sf[i])
    return s2


_SIGMA2 = level_sigma2()
_LEVEL_P = (1.2 ** (-2.0 * np.arange(N_LEVELS)))
_LEVEL_P = _LEVEL_P / _LEVEL_P.sum()


def rodrigues(w: np.ndarray) -> np.ndarray:
    th = float(np.linalg.norm(w))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], float)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def random_pose(rng: np.random.Generator, max_angle: float = 0.5, max_t: float = 1.0):
    """rotation vector uniform in the ball |w| <= max_angle, translation uniform in [-max_t, max_t]^3."""
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    w = d * max_angle * rng.uniform() ** (1.0 / 3.0)
    return rodrigues(w), rng.uniform(-max_t, max_t, size=3)


def frustum_points(rng: np.random.Generator, n: int, cam=EUROC, zmin=2.0, zmax=20.0) -> np.ndarray:
    """points in the camera frame, uniform pixel x uniform depth in [zmin, zmax]."""
    u = rng.uniform(0, cam["width"], size=n)
    v = rng.uniform(0, cam["height"], size=n)
    z = rng.uniform(zmin, zmax, size=n)
    return np.stack([(u - cam["cx"]) / cam["fx"] * z, (v - cam["cy"]) / cam["fy"] * z, z], axis=1)


def project(Xc: np.ndarray, cam=EUROC) -> np.ndarray:
    return np.stack([cam["fx"] * Xc[:, 0] / Xc[:, 2] + cam["cx"], cam["fy"] * Xc[:, 1] / Xc[:, 2] + cam["cy"]], axis=1)


def pnp_problem(seed: int, n: int = 500, outlier_ratio: float = 0.5, cam=EUROC, noise: bool = True):
    """One 2D-3D correspondence set with ground truth.

    Returns dict(p3d f32 [n,3], p2d f32 [n,2], sigma2 f32 [n], octave, inlier bool [n],
                 R [3,3], t [3], K=(fx,fy,cx,cy))."""
    rng = np.random.default_rng(seed)
    R, t = random_pose(rng)
    Xc = frustum_points(rng, n, cam)
    Xw = (Xc - t) @ R          # R^T (Xc - t)
    octave = rng.choice(N_LEVELS, size=n, p=_LEVEL_P)
    sigma2 = _SIGMA2[octave]
    uv = project(Xc, cam)
    if noise:
        uv = uv + rng.normal(size=(n, 2)) * np.sqrt(sigma2.astype(float))[:, None]
    n_out = int(round(n * outlier_ratio))
    inlier = np.ones(n, bool)
    out_idx = rng.permutation(n)[:n_out]
    inlier[out_idx] = False
    uv[out_idx, 0] = rng.uniform(0, cam["width"], size=n_out)
    uv[out_idx, 1] = rng.uniform(0, cam["height"], size=n_out)
    return dict(p3d=np.ascontiguousarray(Xw, np.float32), p2d=np.ascontiguousarray(uv, np.float32),
                sigma2=np.ascontiguousarray(sigma2, np.float32), octave=octave, inlier=inlier, R=R, t=t,
                K=(cam["fx"], cam["fy"], cam["cx"], cam["cy"]))


def pnp_batch(cfg: int, C: int, n: int = 500, outlier_ratio: float = 0.5, cam=EUROC, first: int = 0):
    """C independent problems, seeds 1000*cfg + first + i, stacked: p3d [C,n,3], p2d [C,n,2], sigma2 [C,n]."""
    ps = [pnp_problem(1000 * cfg + first + i, n, outlier_ratio, cam) for i in range(C)]
    return dict(p3d=np.stack([p["p3d"] for p in ps]), p2d=np.stack([p["p2d"] for p in ps]),
                sigma2=np.stack([p["sigma2"] for p in ps]), inlier=np.stack([p["inlier"] for p in ps]),
                R=np.stack([p["R"] for p in ps]), t=np.stack([p["t"] for p in ps]), K=ps[0]["K"],
                seeds=np.array([1000 * cfg + first + i for i in range(C)], np.uint32))


def bearing_covariances(p: dict) -> np.ndarray:
    """cfg2: Sigma_i = diag(sigma_i^2/fx^2, sigma_i^2/fy^2, 0) for the MLPnP use_cov branch
    (reference src/MLPnPsolver.cpp:375-388)."""
    fx, fy = p["K"][0], p["K"][1]
    s2 = p["sigma2"].astype(np.float64)
    cov = np.zeros(s2.shape + (3, 3))
    cov[..., 0, 0] = s2 / fx ** 2
    cov[..., 1, 1] = s2 / fy ** 2
    return cov


def sim3_problem(seed: int, n: int = 200, outlier_ratio: float = 0.4, scale: float = 1.0, cam=EUROC):
    """One 3D-3D correspondence set: camera-frame points of two keyframes observing the same
    landmarks (reference src/Sim3Solver.cpp:57-63), X1c = s*R12*X2c + t12 up to noise.

    Returns dict(x1c, x2c f32 [n,3], sigma2_1, sigma2_2 f32 [n], inlier, R12, t12, s, K)."""
    rng = np.random.default_rng(seed)
    R1, t1 = random_pose(rng, 0.3, 0.5)
    Rrel, trel = random_pose(rng, 0.3, 0.5)
    X1 = frustum_points(rng, n, cam, 3.0, 15.0)              # camera-1 frame
    # camera 2 = camera 1 moved by (Rrel, trel):  X2 = Rrel X1 + trel ; then X1 = R12 (s X2') + t12
    X2 = X1 @ Rrel.T + trel
    # keep only geometry in front of camera 2
    X2[:, 2] = np.maximum(X2[:, 2], 1.0)
    X1 = (X2 - trel) @ Rrel
    R12 = Rrel.T
    t12 = -Rrel.T @ trel
    oct1 = rng.choice(N_LEVELS, size=n, p=_LEVEL_P)
    oct2 = rng.choice(N_LEVELS, size=n, p=_LEVEL_P)
    s1, s2 = _SIGMA2[oct1], _SIGMA2[oct2]
    # lateral noise giving ~sigma px of reprojection noise in each image
    X1n = X1.copy()
    X2n = X2.copy()
    X1n[:, :2] += rng.normal(size=(n, 2)) * (0.5 * np.sqrt(s1.astype(float)) * X1[:, 2] / cam["fx"])[:, None]
    X2n[:, :2] += rng.normal(size=(n, 2)) * (0.5 * np.sqrt(s2.astype(float)) * X2[:, 2] / cam["fx"])[:, None]
    n_out = int(round(n * outlier_ratio))
    inlier = np.ones(n, bool)
    out_idx = rng.permutation(n)[:n_out]
    inlier[out_idx] = False
    X2n[out_idx] = frustum_points(rng, n_out, cam, 3.0, 15.0)
    X2n = X2n / scale                                       # monocular map 2 lives at another scale
    K = (np.float32(cam["fx"]), np.float32(cam["fy"]), np.float32(cam["cx"]), np.float32(cam["cy"]))
    return dict(x1c=np.ascontiguousarray(X1n, np.float32), x2c=np.ascontiguousarray(X2n, np.float32),
                sigma2_1=np.ascontiguousarray(s1, np.float32), sigma2_2=np.ascontiguousarray(s2, np.float32),
                inlier=inlier, R12=R12, t12=t12, s=scale, K=K)


def scoring_stress(seed: int = 5000, H: int = 4096, n: int = 10000, cam=EUROC):
    """cfg5: one 50%-outlier correspondence set of size n and H poses = ground truth perturbed by
    N(0, 0.02) in rotation vector and translation."""
    p = pnp_problem(seed, n, 0.5, cam)
    rng = np.random.default_rng(seed + 1)
    poses = np.empty((H, 12), np.float32)
    for h in range(H):
        dR = rodrigues(rng.normal(size=3) * 0.02)
        R = dR @ p["R"]
        t = p["t"] + rng.normal(size=3) * 0.02
        poses[h, :9] = R.reshape(-1)
        poses[h, 9:] = t
    p["poses"] = poses
    return p


def poseopt_problem(seed: int, n: int = 250, outlier_ratio: float = 0.2, stereo_ratio: float = 0.0, cam=EUROC,
                    bf: float = 47.9, pose_noise=(0.02, 0.05), noise_scale: float = 1.0):
    """One Optimizer::PoseOptimization input (Optimizer.cpp:244-323): matched map points, observations
    (u, v, uR; uR < 0 for monocular keypoints), 1/sigma^2 per keypoint level, and an initial pose = ground truth
    perturbed by a rotation of `pose_noise[0]` rad and a translation of `pose_noise[1]` m (what RANSAC + Refine
    hands over); inlier pixel noise is `noise_scale` x the level's sigma.  Returns dict(p3d [n,3], obs [n,3], inv_sigma2 [n], K (fx,fy,cx,cy,bf) f32, Rcw [3,3] f32,
    tcw [3] f32, R, t ground truth, inlier [n])."""
    p = pnp_problem(seed, n, outlier_ratio, cam, noise=False)
    rng = np.random.default_rng(seed + 77_000_000)
    sig = np.sqrt(p["sigma2"].astype(float)) * noise_scale
    p["p2d"] = (p["p2d"] + np.where(p["inlier"][:, None], rng.normal(size=(n, 2)) * sig[:, None], 0.0)).astype(np.float32)
    Xc = p["p3d"].astype(float) @ p["R"].T + p["t"]
    ur = np.full(n, -1.0)
    is_st = rng.uniform(size=n) < stereo_ratio
    ur_true = p["p2d"][:, 0].astype(float) - bf / Xc[:, 2]
    ur[is_st] = ur_true[is_st] + rng.normal(size=int(is_st.sum())) * sig[is_st]
    bad = is_st & ~p["inlier"]
    ur[bad] = rng.uniform(0, cam["width"], size=int(bad.sum()))
    ur[is_st & (ur < 0)] = 0.0
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    dR = rodrigues(d * pose_noise[0])
    dt = rng.normal(size=3) * pose_noise[1]
    R0 = dR @ p["R"]
    t0 = dR @ p["t"] + dt
    inv_sigma2 = (np.float32(1.0) / p["sigma2"]).astype(np.float32)      # mvInvLevelSigma2 (ORBextractor.cpp:361-365)
    obs = np.concatenate([p["p2d"], ur[:, None].astype(np.float32)], axis=1).astype(np.float32)
    return dict(p3d=p["p3d"], obs=np.ascontiguousarray(obs), inv_sigma2=inv_sigma2,
                K=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"], bf], np.float32),
                Rcw=np.ascontiguousarray(R0, np.float32), tcw=np.ascontiguousarray(t0, np.float32),
                R=p["R"], t=p["t"], inlier=p["inlier"])


def sim3opt_problem(seed: int, n: int = 100, outlier_ratio: float = 0.15, cam=EUROC, pose_noise=(0.01, 0.03), scale: float = 1.0):
    """One Optimizer::OptimizeSim3 input (Optimizer.cpp:1054-1160): matched map points of two keyframes in their own
    camera frames, their keypoints (projection + sigma px of noise per level), 1/sigma^2, and an initial S12 = ground
    truth perturbed by `pose_noise` (rad, m) -- what Sim3Solver hands over.  Returns dict(x1c, x2c [n,3], obs1, obs2
    [n,2], inv_sigma2_1/2 [n], K (fx,fy,cx,cy) f32, S12 [13] f32 (R, t, s), R12, t12, s, inlier)."""
    q = sim3_problem(seed, n, outlier_ratio, scale, cam)
    rng = np.random.default_rng(seed + 55_000_000)
    sg1 = np.sqrt(q["sigma2_1"].astype(float))
    sg2 = np.sqrt(q["sigma2_2"].astype(float))
    obs1 = project(q["x1c"].astype(float), cam) + rng.normal(size=(n, 2)) * 0.5 * sg1[:, None]
    obs2 = project(q["x2c"].astype(float) * scale, cam) + rng.normal(size=(n, 2)) * 0.5 * sg2[:, None]
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    dR = rodrigues(d * pose_noise[0])
    R0 = dR @ q["R12"]
    t0 = dR @ q["t12"] + rng.normal(size=3) * pose_noise[1]
    S12 = np.concatenate([R0.ravel(), t0, [q["s"]]]).astype(np.float32)
    return dict(x1c=q["x1c"], x2c=q["x2c"], obs1=np.ascontiguousarray(obs1, np.float32), obs2=np.ascontiguousarray(obs2, np.float32),
                inv_sigma2_1=(np.float32(1) / q["sigma2_1"]).astype(np.float32), inv_sigma2_2=(np.float32(1) / q["sigma2_2"]).astype(np.float32),
                K=np.array(q["K"], np.float32), S12=S12, R12=q["R12"], t12=q["t12"], s=q["s"], inlier=q["inlier"])


# ---------------------------------------------------------------- ORB features + bag-of-words feature vectors (SearchByBoW)
def _feature_vector(nodes: np.ndarray):
    """DBoW2::FeatureVector as CSR: node ids ascending, per node the feature indices in insertion (= index) order"""
    order = np.argsort(nodes, kind="stable")
    ids, start = np.unique(nodes[order], return_index=True)
    off = np.concatenate([start, [len(nodes)]]).astype(np.int32)
    return ids.astype(np.uint32), off, order.astype(np.uint32)


def bow_frame(seed: int, n_feat: int = 1500, n_nodes: int = 100):
    """a frame: random 256-bit ORB descriptors, keypoint angles, every feature in one of n_nodes vocabulary nodes"""
    rng = np.random.default_rng(seed)
    desc = rng.integers(0, 2 ** 32, size=(n_feat, 8), dtype=np.uint64).astype(np.uint32)
    angle = rng.uniform(0.0, 360.0, n_feat).astype(np.float32)
    node_pool = np.sort(rng.choice(10 ** 6, n_nodes, replace=False)).astype(np.int64)
    nodes = node_pool[rng.integers(0, n_nodes, n_feat)]
    ids, off, feat = _feature_vector(nodes)
    return dict(desc=desc, angle=angle, valid=None, node_ids=ids, node_off=off, node_feat=feat, nodes=nodes, node_pool=node_pool)


def bow_keyframe(seed: int, frame: dict, n_feat: int = 1500, shared: float = 0.4, flip_bits: int = 25, rot: float = 33.0,
                 valid_ratio: float = 0.7, wrong_node: float = 0.05, wrong_rot: float = 0.1):
    """a keyframe that saw `shared` of the frame's features: their descriptors with up to flip_bits bits flipped, the same
    vocabulary node (a few land in another node), angle = frame angle + rot (+ noise; some with an unrelated rotation);
    the rest are unrelated features.  valid = the feature has a usable MapPoint."""
    rng = np.random.default_rng(seed)
    nF = frame["desc"].shape[0]
    desc = rng.integers(0, 2 ** 32, size=(n_feat, 8), dtype=np.uint64).astype(np.uint32)
    angle = rng.uniform(0.0, 360.0, n_feat).astype(np.float32)
    pool = frame["node_pool"]
    nodes = pool[rng.integers(0, len(pool), n_feat)]
    ns = int(shared * min(n_feat, nF))
    src = rng.choice(nF, ns, replace=False)
    dst = rng.choice(n_feat, ns, replace=False)
    d = frame["desc"][src].copy()
    for i in range(ns):
        nb = int(rng.integers(0, flip_bits + 1))
        bits = rng.choice(256, nb, replace=False)
        for b in bits:
            d[i, b >> 5] ^= np.uint32(1) << np.uint32(b & 31)
    desc[dst] = d
    nodes[dst] = frame["nodes"][src]
    moved = rng.random(ns) < wrong_node
    nodes[dst[moved]] = pool[rng.integers(0, len(pool), int(moved.sum()))]
    a = frame["angle"][src].astype(np.float64) + rot + rng.normal(0.0, 3.0, ns)
    odd = rng.random(ns) < wrong_rot
    a[odd] += rng.uniform(40.0, 320.0, int(odd.sum()))
    angle[dst] = np.mod(a, 360.0).astype(np.float32)
    valid = (rng.random(n_feat) < valid_ratio).astype(np.uint8)
    ids, off, feat = _feature_vector(nodes)
    return dict(desc=desc, angle=angle, valid=valid, node_ids=ids, node_off=off, node_feat=feat, nodes=nodes, node_pool=pool,
                truth=(dst, src))


# ---------------------------------------------------------------- one frame, many candidate keyframes (indexed wire format)
def reloc_frame(seed: int, C: int, n_kp: int = 2000, n_match: int = 500, outlier_ratio: float = 0.5, n_map: int = 200000, cam=EUROC):
    """A relocalisation as Tracking::Relocalization sees it: ONE frame (n_kp undistorted keypoints with octaves, one true
    pose) and C candidate keyframes, each contributing n_match (keypoint, map point) pairs -- the true map point of the
    keypoint (inliers) or an unrelated one (outliers).  Returns the tables, the index pairs and the equivalent flat arrays."""
    rng = np.random.default_rng(seed)
    R, t = random_pose(rng)
    octave = rng.choice(N_LEVELS, size=n_kp, p=_LEVEL_P)
    sigma2 = np.ascontiguousarray(_SIGMA2[octave], np.float32)
    Xc = frustum_points(rng, n_kp, cam)
    uv = (project(Xc, cam) + rng.normal(size=(n_kp, 2)) * np.sqrt(sigma2)[:, None]).astype(np.float32)
    mp = np.empty((n_map, 3), np.float32)
    true_id = rng.choice(n_map, n_kp, replace=False)
    mp[:] = ((frustum_points(rng, n_map, cam) - t) @ R).astype(np.float32)
    mp[true_id] = ((Xc - t) @ R).astype(np.float32)
    kp_idx = np.empty((C, n_match), np.uint16)
    mp_idx = np.empty((C, n_match), np.uint32)
    for c in range(C):
        ks = rng.choice(n_kp, n_match, replace=False)
        ms = true_id[ks].copy()
        out = rng.random(n_match) < outlier_ratio
        # wrong matches: distinct map points, none of them one of this candidate's true points -- SearchByBoW hands every
        # keyframe map point to at most one frame keypoint (ORBmatcher.cpp:162,219), so a candidate's set never holds a map
        # point twice (a duplicate would make the minimal sets that draw both copies degenerate)
        n_out = int(out.sum())
        pool = np.setdiff1d(rng.choice(n_map, 2 * n_out + 16, replace=False), ms, assume_unique=False)
        ms[out] = rng.permutation(pool)[:n_out]
        kp_idx[c], mp_idx[c] = ks, ms
    flat = dict(p3d=mp[mp_idx], p2d=uv[kp_idx], sigma2=sigma2[kp_idx])
    return dict(K=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float64), R=R, t=t, kp_uv=uv, kp_sigma2=sigma2, mp_xyz=mp, kp_idx=kp_idx, mp_idx=mp_idx,
                seeds=(np.arange(C) + 1000 * (seed % 1000) + 7).astype(np.uint32), **flat)


# ---------------------------------------------------------------- keyframe database (candidate retrieval)
def _bow_vector(rng, words):
    """an L1-normalised TF-IDF style BowVector over the given word ids (ascending, unique)"""
    words = np.unique(np.asarray(words, np.uint32))
    w = rng.gamma(2.0, 1.0, len(words)) + 0.05
    return words, (w / w.sum()).astype(np.float64)


def kf_database(seed: int, K: int = 400, n_places: int = 40, vocab: int = 100000, words_per_kf: int = 900, pool: int = 2500):
    """A keyframe database as KeyFrameDatabase sees it: K keyframes along a trajectory through n_places places.  A place owns a
    pool of vocabulary words; a keyframe draws most of its words from the pools of the (one or two) places it sees and a few
    from anywhere (words shared by unrelated keyframes).  Covisibility: the ten nearest keyframes of the same stretch of the
    trajectory, nearest first.  Returns the CSR arrays (bow_off, bow_word, bow_val), covis [K,10], and the place of every keyframe."""
    rng = np.random.default_rng(seed)
    pools = [rng.choice(vocab, pool, replace=False) for _ in range(n_places)]
    place = np.minimum((np.arange(K) * n_places) // max(K, 1), n_places - 1)
    # revisits: the last tenth of the trajectory goes back to the first places (what loop closing looks for)
    nrev = K // 10
    if nrev:
        place[K - nrev:] = place[:nrev]
    offs, ws, vs = [0], [], []
    for k in range(K):
        p = place[k]
        q = min(p + 1, n_places - 1)
        mix = rng.random()
        n_local = int(words_per_kf * 0.85)
        a = rng.choice(pools[p], int(n_local * (0.5 + 0.5 * mix)), replace=False)
        b = rng.choice(pools[q], max(1, n_local - len(a)), replace=False)
        c = rng.integers(0, vocab, words_per_kf - n_local)
        w, v = _bow_vector(rng, np.concatenate([a, b, c]))
        ws.append(w); vs.append(v); offs.append(offs[-1] + len(w))
    covis = np.full((K, 10), -1, np.int32)
    for k in range(K):
        near = [j for d in range(1, 9) for j in (k - d, k + d) if 0 <= j < K and abs(int(place[j]) - int(place[k])) <= 1]
        if rng.random() < 0.3:
            near = near[:rng.integers(0, 6)]                  # young keyframes have few connections
        covis[k, :min(10, len(near))] = near[:10]
    return dict(K=K, bow_off=np.array(offs, np.int64), bow_word=np.concatenate(ws), bow_val=np.concatenate(vs), covis=covis,
                place=place, pools=pools, vocab=vocab)


def kf_query(seed: int, db: dict, place: int, n_words: int = 1000):
    """the BowVector of a frame that looks at `place`"""
    rng = np.random.default_rng(seed)
    p = int(place)
    q = min(p + 1, len(db["pools"]) - 1)
    a = rng.choice(db["pools"][p], int(n_words * 0.6), replace=False)
    b = rng.choice(db["pools"][q], int(n_words * 0.25), replace=False)
    c = rng.integers(0, db["vocab"], n_words - len(a) - len(b))
    return _bow_vector(rng, np.concatenate([a, b, c]))


# ---------------------------------------------------------------- keyframe views for guided matching
GRID_COLS, GRID_ROWS = 64, 48      # FRAME_GRID_COLS / FRAME_GRID_ROWS (include/Frame.hpp:20-21)


def _grid_csr(xy: np.ndarray, cam=EUROC):
    """Frame::AssignFeaturesToGrid (src/Frame.cpp:293-318) + PosInGrid (:449-459): cell (ix, iy) at ix*ROWS + iy, features in index order"""
    w_inv = np.float32(GRID_COLS) / np.float32(cam["width"] - 0.0)
    h_inv = np.float32(GRID_ROWS) / np.float32(cam["height"] - 0.0)
    px = np.round((xy[:, 0] - np.float32(0)) * w_inv).astype(np.int64)
    py = np.round((xy[:, 1] - np.float32(0)) * h_inv).astype(np.int64)
    ok = (px >= 0) & (px < GRID_COLS) & (py >= 0) & (py < GRID_ROWS)
    cell = np.where(ok, px * GRID_ROWS + py, -1)
    off = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    idx = []
    for c in range(GRID_COLS * GRID_ROWS):
        members = np.flatnonzero(cell == c)
        idx.append(members)
        off[c + 1] = off[c] + len(members)
    return off, (np.concatenate(idx) if idx else np.zeros(0)).astype(np.int32), float(w_inv), float(h_inv)


def kf_view(rng, R, t, mp_xyz, mp_desc, mp_maxdist, mp_mindist, visible, n_extra=400, flip_bits=20, cam=EUROC, pix_noise=0.7,
            mp_angle=None, roll=0.0):
    """What ORBmatcher reads from one keyframe: features = the visible map points (projected, noisy pixel, octave predicted from
    the distance, descriptor = the map point's with a few bits flipped) followed by n_extra features without a map point."""
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    Xc = mp_xyz[visible].astype(np.float64) @ R.T + t
    uv = project(Xc, cam) + rng.normal(size=(len(visible), 2)) * pix_noise
    dist = np.linalg.norm(Xc, axis=1)
    lvl = np.clip(np.ceil(np.log(mp_maxdist[visible] / dist) / np.log(1.2)) - rng.integers(0, 2, len(visible)), 0, 7).astype(np.int32)
    inside = (uv[:, 0] > 1) & (uv[:, 0] < cam["width"] - 1) & (uv[:, 1] > 1) & (uv[:, 1] < cam["height"] - 1) & (Xc[:, 2] > 0.1)
    visible = visible[inside]
    uv, lvl = uv[inside], lvl[inside]
    n_mp = len(visible)
    d = mp_desc[visible].copy()
    for i in range(n_mp):
        bits = rng.choice(256, rng.integers(0, flip_bits + 1), replace=False)
        for b in bits:
            d[i, b >> 5] ^= np.uint32(1) << np.uint32(b & 31)
    ex_uv = np.stack([rng.uniform(1, cam["width"] - 1, n_extra), rng.uniform(1, cam["height"] - 1, n_extra)], 1)
    ex_d = rng.integers(0, 2 ** 32, (n_extra, 8), dtype=np.uint64).astype(np.uint32)
    n = n_mp + n_extra
    perm = rng.permutation(n)                                  # features in detection order, not map-point order
    kp_xy = np.concatenate([uv, ex_uv]).astype(np.float32)[perm]
    octave = np.concatenate([lvl, rng.integers(0, 8, n_extra).astype(np.int32)])[perm]
    # keypoint orientation: the map point's canonical angle seen under this keyframe's in-plane rotation (+ noise); a few wrong
    ang_mp = (mp_angle[visible] + roll + rng.normal(0.0, 3.0, n_mp)) if mp_angle is not None else rng.uniform(0, 360, n_mp)
    odd = rng.random(n_mp) < 0.08
    ang_mp = np.where(odd, ang_mp + rng.uniform(40, 320, n_mp), ang_mp)
    angle = np.mod(np.concatenate([ang_mp, rng.uniform(0, 360, n_extra)]), 360.0).astype(np.float32)[perm]
    desc = np.concatenate([d, ex_d])[perm]
    mp_id = np.concatenate([visible, np.full(n_extra, -1)]).astype(np.int64)[perm]
    valid = (mp_id >= 0).astype(np.uint8)
    bad = (rng.random(n) < 0.03) & (mp_id >= 0)                # a few bad map points (isBad)
    valid[bad] = 0
    safe = np.maximum(mp_id, 0)
    off, idx, w_inv, h_inv = _grid_csr(kp_xy, cam)
    return dict(n_feat=n, kp_xy=kp_xy, kp_octave=octave.astype(np.int32), kp_angle=angle, desc=np.ascontiguousarray(desc), mp_valid=valid, mp_id=mp_id,
                mp_xyz=np.ascontiguousarray(mp_xyz[safe], np.float32), mp_desc=np.ascontiguousarray(mp_desc[safe]),
                mp_maxdist=np.ascontiguousarray(mp_maxdist[safe], np.float32), mp_mindist=np.ascontiguousarray(mp_mindist[safe], np.float32),
                Rcw=R.astype(np.float32), tcw=t.astype(np.float32), bounds=np.array([0.0, cam["width"], 0.0, cam["height"]], np.float32),
                grid_cols=GRID_COLS, grid_rows=GRID_ROWS, grid_w_inv=w_inv, grid_h_inv=h_inv, grid_off=off, grid_idx=idx,
                n_levels=8, scale_factors=sf, log_scale_factor=float(np.float32(np.log(np.float32(1.2)))))


def kf_view_pair(seed: int, n_points: int = 1200, n_extra: int = 400, pose_noise: float = 0.01, prematched: float = 0.3, cam=EUROC):
    """Two keyframes of a loop closure looking at the same n_points map points from nearby poses, the Sim3 (R12, t12, s = 1)
    that Sim3Solver would hand over (ground truth + a small perturbation), and the matches SearchByBoW already found
    (matched12_in: KF2 feature index, -2 = a map point KF2 does not observe, -1 = none)."""
    rng = np.random.default_rng(seed)
    R1, t1 = random_pose(rng, 0.3, 0.5)
    dR, dt = random_pose(rng, 0.15, 0.6)
    R2, t2 = dR @ R1, dR @ t1 + dt
    Xc1 = frustum_points(rng, n_points, cam, 2.0, 15.0)
    Xw = (Xc1 - t1) @ R1
    mp_desc = rng.integers(0, 2 ** 32, (n_points, 8), dtype=np.uint64).astype(np.uint32)
    ref_dist = np.linalg.norm(Xc1, axis=1)
    ref_lvl = rng.integers(0, 5, n_points)
    mp_maxdist = (ref_dist * (1.2 ** ref_lvl)).astype(np.float32)          # MapPoint::UpdateNormalAndDepth: dist * levelScaleFactor
    mp_mindist = (mp_maxdist / np.float32(1.2 ** 7)).astype(np.float32)
    vis = np.arange(n_points)
    mp_angle = rng.uniform(0, 360, n_points)
    kf1 = kf_view(rng, R1, t1, Xw, mp_desc, mp_maxdist, mp_mindist, vis[rng.random(n_points) < 0.9], n_extra, cam=cam, mp_angle=mp_angle, roll=0.0)
    kf2 = kf_view(rng, R2, t2, Xw, mp_desc, mp_maxdist, mp_mindist, vis[rng.random(n_points) < 0.9], n_extra, cam=cam, mp_angle=mp_angle, roll=25.0)
    # T12 = T1w * T2w^-1 (+ noise)
    nR, nt = random_pose(rng, pose_noise, pose_noise * 3)
    R12 = nR @ R1 @ R2.T
    t12 = nR @ (t1 - R1 @ R2.T @ t2) + nt
    idx2_of = {int(m): i for i, m in enumerate(kf2["mp_id"]) if m >= 0}
    matched = np.full(kf1["n_feat"], -1, np.int32)
    for i, m in enumerate(kf1["mp_id"]):
        if m >= 0 and rng.random() < prematched:
            matched[i] = idx2_of.get(int(m), -2)
    K = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32)
    return dict(kf1=kf1, kf2=kf2, K=K, R12=R12.astype(np.float32), t12=t12.astype(np.float32), matched12_in=matched)


def _clone_kf_features(rng, kf: dict, frac: float):
    """append near-duplicates of a fraction of the keyframe's map-point features (same 3D point up to millimetres, descriptor
    with a few flipped bits): they compete for the SAME frame keypoint, which is what makes SearchByProjection's greedy,
    order-dependent assignment (ORBmatcher.cpp:1389,1403) observable"""
    src = np.flatnonzero(kf["mp_valid"] > 0)
    src = rng.choice(src, int(len(src) * frac), replace=False) if len(src) else src
    if len(src) == 0:
        return kf
    out = dict(kf)
    d = kf["mp_desc"][src].copy()
    for i in range(len(src)):
        for b in rng.choice(256, rng.integers(0, 6), replace=False):
            d[i, b >> 5] ^= np.uint32(1) << np.uint32(b & 31)
    app = dict(kp_xy=kf["kp_xy"][src], kp_octave=kf["kp_octave"][src], kp_angle=kf["kp_angle"][src], desc=kf["desc"][src],
               mp_valid=kf["mp_valid"][src], mp_id=-2 - np.arange(len(src)), mp_desc=d,
               mp_xyz=(kf["mp_xyz"][src] + rng.normal(0, 0.002, (len(src), 3))).astype(np.float32),
               mp_maxdist=kf["mp_maxdist"][src], mp_mindist=kf["mp_mindist"][src])
    for k_, v in app.items():
        out[k_] = np.ascontiguousarray(np.concatenate([kf[k_], v]))
    out["n_feat"] = kf["n_feat"] + len(src)
    # interleave: clones must not all come last (the greedy order matters)
    perm = rng.permutation(out["n_feat"])
    for k_ in app:
        out[k_] = np.ascontiguousarray(out[k_][perm])
    off, idx, w_inv, h_inv = _grid_csr(out["kp_xy"])
    out["grid_off"], out["grid_idx"] = off, idx
    return out


def proj_search_case(seed: int, n_points: int = 1000, n_extra: int = 400, found: float = 0.25, pose_noise: float = 0.004, clones: float = 0.15,
                     cam=EUROC):
    """Tracking::Relocalization after the first PoseOptimization of a candidate (Tracking.cpp:1285-1296): the current frame with a
    pose estimate, some of its keypoints already holding map points (`occupied`, the inliers so far = sFound), and the candidate
    keyframe whose remaining map points SearchByProjection tries to add."""
    rng = np.random.default_rng(seed)
    p = kf_view_pair(seed, n_points=n_points, n_extra=n_extra, prematched=0.0, cam=cam)
    frame, kf = p["kf2"], _clone_kf_features(rng, p["kf1"], clones)
    nR, nt = random_pose(rng, pose_noise, pose_noise * 2)
    Rcw = (nR @ frame["Rcw"].astype(np.float64)).astype(np.float32)
    tcw = (nR @ frame["tcw"].astype(np.float64) + nt).astype(np.float32)
    common = np.intersect1d(frame["mp_id"][frame["mp_id"] >= 0], kf["mp_id"][kf["mp_id"] >= 0])
    sfound = set(rng.choice(common, int(len(common) * found), replace=False).tolist()) if len(common) else set()
    occupied = np.array([1 if int(m) in sfound else 0 for m in frame["mp_id"]], np.uint8)
    occupied |= (rng.random(frame["n_feat"]) < 0.02).astype(np.uint8)          # keypoints holding map points of other keyframes
    already = np.array([1 if int(m) in sfound else 0 for m in kf["mp_id"]], np.uint8)
    return dict(frame=frame, kf=kf, K=p["K"], Rcw=Rcw, tcw=tcw, occupied=occupied, already_found=already)


# ---------------------------------------------------------------- a relocalisation with descriptors AND geometry
def reloc_world(seed: int, C: int = 24, n_kp: int = 1500, n_kf_feat: int = 1200, n_map: int = 60000, n_nodes: int = 100, cam=EUROC):
    """One lost frame and C candidate keyframes, consistent across the stages of Tracking::Relocalization: the frame's keypoints
    carry descriptors, vocabulary nodes and angles (SearchByBoW's input) AND are projections of map points under the true pose
    (the PnP solver's input).  Keyframe c re-observes a share of the frame's features (descriptor with flipped bits, same node)
    and its features point at map-point table slots through mp_index: a re-observed feature at the frame keypoint's true map
    point, the others at unrelated points.  Candidates 0, 5, 10, ... share next to nothing with the frame (discarded: < 15 matches)."""
    rng = np.random.default_rng(seed)
    R, t = random_pose(rng)
    frame = bow_frame(seed + 1, n_kp, n_nodes)
    Xc = frustum_points(rng, n_kp, cam)
    octave = rng.choice(N_LEVELS, size=n_kp, p=_LEVEL_P)
    sigma2 = np.ascontiguousarray(_SIGMA2[octave], np.float32)
    uv = (project(Xc, cam) + rng.normal(size=(n_kp, 2)) * np.sqrt(sigma2)[:, None]).astype(np.float32)
    mp = ((frustum_points(rng, n_map, cam) - t) @ R).astype(np.float32)
    true_id = rng.choice(n_map, n_kp, replace=False)                  # map-point table slot of every frame keypoint's true point
    mp[true_id] = ((Xc - t) @ R).astype(np.float32)
    kfs = []
    for c in range(C):
        shared = 0.01 if c % 5 == 0 else float(rng.uniform(0.15, 0.4))
        kf = bow_keyframe(seed * 1000 + c, frame, n_kf_feat, shared=shared, rot=float(rng.uniform(0, 360)), valid_ratio=0.85)
        dst, src = kf["truth"]
        mpi = rng.choice(n_map, n_kf_feat, replace=False).astype(np.uint32)       # unrelated points (distinct per keyframe)
        good = rng.random(len(dst)) < 0.8                                          # 20 % of the re-observations are wrong associations
        mpi[dst[good]] = true_id[src[good]].astype(np.uint32)
        kf["mp_index"] = mpi
        kfs.append(kf)
    return dict(frame=frame, kfs=kfs, kp_uv=uv, kp_sigma2=sigma2, mp_xyz=mp,
                K=np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float64), R=R, t=t,
                seeds=(np.arange(C) + 31 * seed + 5).astype(np.uint32))
