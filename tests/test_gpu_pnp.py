"""GPU parity: batched EPnP RANSAC (CUDA, through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): identical minimal sets => inlier counts and masks
bit-exact (no near-threshold allowance is needed: the scoring kernel reproduces the
reference's inlier bit exactly), poses of 4-point hypotheses bit-identical (shared
arithmetic contract), final R/t within 1e-4 relative.
"""
import numpy as np
import pytest

from ransac_b200 import capi, synth

pytestmark = pytest.mark.gpu

PRM = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.2, th2=5.991)   # cfg1: H = 300, minInl = 100


def _batch(cfg, C, n, outl=0.5, first=0):
    b = synth.pnp_batch(cfg, C, n, outl, first=first)
    offsets = np.arange(C + 1, dtype=np.int32) * n
    return b, offsets


def _oracle_all(oracle, b, C, prm, tables, flags=None, per_hyp=True):
    """flags=None: the engine's default mode (QR null space for the 4-point sets)"""
    if flags is None:
        flags = oracle.FLAG_EPNP_QR_NULLSPACE
    out = []
    for c in range(C):
        pb = oracle.pnp_problem(b["p3d"][c], b["p2d"][c], b["sigma2"][c], b["K"])
        out.append(oracle.pnp_ransac(pb, oracle.params(**prm), tables[c], flags | oracle.FLAG_EXHAUSTIVE, per_hyp=per_hyp))
    return out


def _check_results(res, masks_list, orc):
    for c, o in enumerate(orc):
        r = res[c]
        assert r["ok"] == o["ok"], c
        assert r["no_more"] == o["no_more"], c
        assert r["n_inliers"] == o["n_inliers"], c
        assert r["best_hyp"] == o["best_hyp"], c
        assert r["refined"] == o["refined"], c
        assert r["best_count"] == o["best_count"], c
        assert (masks_list[c] == o["mask"]).all(), c
        if o["ok"]:
            T = o["T"]
            # north_star tolerance: 1e-4 relative on R, t (observed: bit-identical)
            assert np.allclose(r["R"].reshape(3, 3), T[:3, :3], rtol=1e-4, atol=1e-6), c
            assert np.allclose(r["t"], T[:3, 3], rtol=1e-4, atol=1e-6), c


@pytest.mark.parametrize("mode", ["qr", "eigen"])
def test_pnp_cfg1_per_hypothesis_bit_exact(engine, oracle, mode):
    """cfg1: N=500, 50% outliers, H=300; 8 problems; explicit index tables.  Both null-space modes:
    the default Householder QR of the 4-point M^T and the 12x12 M^T M eigen-solve (RSAC_FLAG_EPNP_EIGEN)."""
    C, n = 8, 500
    b, offsets = _batch(1, C, n)
    tables = [oracle.index_table(int(s), n, 4, 300) for s in b["seeds"]]
    toff = np.arange(C + 1, dtype=np.int64) * 1200
    dev_flags = capi.FLAG_EPNP_EIGEN if mode == "eigen" else 0
    orc_flags = 0 if mode == "eigen" else oracle.FLAG_EPNP_QR_NULLSPACE
    res, masks = engine.pnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**PRM),
                                  tables=np.concatenate(tables), table_offsets=toff, flags=dev_flags)
    poses, counts = engine.pnp_hypotheses()
    orc = _oracle_all(oracle, b, C, PRM, tables, flags=orc_flags)
    for c in range(C):
        hp = poses[c * 300:(c + 1) * 300]
        op = orc[c]["hyp_pose"]
        same = (hp.view(np.uint32) == op.view(np.uint32)) | (np.isnan(hp) & np.isnan(op))
        assert same.all(), f"problem {c}: {np.argwhere(~same.all(axis=1)).ravel()[:5]} differ"
        assert (counts[c * 300:(c + 1) * 300] == orc[c]["hyp_counts"]).all()
    _check_results(res, engine.split_masks(masks, offsets), orc)
    assert sum(o["ok"] for o in orc) >= C - 1          # the synthetic problems are solvable


def test_pnp_duplicate_map_points_nan_hypotheses(engine, oracle):
    """Two keypoints matched to the SAME (wrong) map point: a minimal set that draws both is degenerate and EPnP returns a
    NaN translation.  The reference scores such a pose as "no inliers" (NaN compares false, PnPsolver.cpp:258-262); the
    scoring kernel takes a shortcut for NaN poses instead of the exact tier -- hypothesis poses, counts, records and masks
    must still equal the oracle's, in the exhaustive and in the staged run."""
    C, n = 6, 300
    b, offsets = _batch(13, C, n)
    p3d = b["p3d"].copy()
    rng = np.random.default_rng(5)
    tables = [np.asarray(oracle.index_table(int(s), n, 4, 300)).reshape(300, 4).copy() for s in b["seeds"]]
    for c in range(C):
        for h in rng.choice(300, 12, replace=False):          # 12 hypotheses per problem draw one map point twice
            i, j = tables[c][h][0], tables[c][h][2]
            p3d[c][j] = p3d[c][i]
    tabs = [t.reshape(-1).astype(np.uint32) for t in tables]
    toff = np.arange(C + 1, dtype=np.int64) * 1200
    orc = _oracle_all(oracle, dict(b, p3d=p3d), C, PRM, tabs)
    assert sum(int(np.isnan(o["hyp_pose"]).any(axis=1).sum()) for o in orc) >= C      # the case is really exercised
    for flags in (0, capi.FLAG_EARLY_EXIT):
        res, masks = engine.pnp_solve(offsets, p3d, b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**PRM),
                                      tables=np.concatenate(tabs), table_offsets=toff, flags=flags)
        _check_results(res, engine.split_masks(masks, offsets), orc)
        if flags == 0:
            poses, counts = engine.pnp_hypotheses()
            for c in range(C):
                hp, op = poses[c * 300:(c + 1) * 300], orc[c]["hyp_pose"]
                assert ((hp.view(np.uint32) == op.view(np.uint32)) | (np.isnan(hp) & np.isnan(op))).all(), c
                assert (counts[c * 300:(c + 1) * 300] == orc[c]["hyp_counts"]).all(), c


def test_pnp_device_generated_tables_match_libc(engine, oracle):
    """seeds only: the device restatement of glibc rand() must give the oracle's tables"""
    C, n = 5, 300
    b, offsets = _batch(11, C, n)
    res, masks = engine.pnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**PRM), seeds=b["seeds"])
    minInl, H = capi.pnp_ransac_setup(n, capi.ransac_params(**PRM))
    tables = [oracle.index_table(int(s), n, 4, H) for s in b["seeds"]]
    orc = _oracle_all(oracle, b, C, PRM, tables, per_hyp=False)
    _check_results(res, engine.split_masks(masks, offsets), orc)


def test_pnp_ragged_and_edge_cases(engine, oracle):
    """ragged batch incl. an empty problem, n < minInliers, n == minSet, all-outlier and noise-free sets"""
    sizes = [0, 3, 4, 9, 37, 64, 65, 500, 257]
    prm = dict(PRM)
    parts = []
    for i, n in enumerate(sizes):
        outl = 1.0 if i == 4 else (0.0 if i == 5 else 0.4)
        parts.append(synth.pnp_problem(7000 + i, max(n, 1), outl, noise=(i != 6)))
    p3d = np.concatenate([p["p3d"][:n] for p, n in zip(parts, sizes)])
    p2d = np.concatenate([p["p2d"][:n] for p, n in zip(parts, sizes)])
    s2 = np.concatenate([p["sigma2"][:n] for p, n in zip(parts, sizes)])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    seeds = np.arange(len(sizes), dtype=np.uint32) + 77
    res, masks = engine.pnp_solve(offsets, p3d, p2d, s2, [parts[0]["K"]], capi.ransac_params(**prm), seeds=seeds)
    ml = engine.split_masks(masks, offsets)
    for c, n in enumerate(sizes):
        minInl, H = (capi.pnp_ransac_setup(n, capi.ransac_params(**prm)) if n > 0 else (10, 0))
        if n < max(minInl, 4):
            assert res[c]["ok"] == 0 and res[c]["no_more"] == 1 and res[c]["n_inliers"] == 0
            assert not ml[c].any()
            continue
        pb = oracle.pnp_problem(parts[c]["p3d"][:n], parts[c]["p2d"][:n], parts[c]["sigma2"][:n], parts[c]["K"])
        o = oracle.pnp_ransac(pb, oracle.params(**prm), oracle.index_table(int(seeds[c]), n, 4, H), oracle.FLAG_EPNP_QR_NULLSPACE)
        _check_results(res[c:c + 1], ml[c:c + 1], [o])


def test_pnp_tracking_parameters_failed_refines(engine, oracle):
    """Tracking's own call (src/Tracking.cpp:1226): eps=0.5 => H=35, minInl=N/2: refines do fail here;
    the engine must follow the clean (n-bounded) semantics through them."""
    prm = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.5, th2=5.991)
    C, n = 48, 200
    b, offsets = _batch(12, C, n, outl=0.45)
    res, masks = engine.pnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**prm), seeds=b["seeds"])
    minInl, H = capi.pnp_ransac_setup(n, capi.ransac_params(**prm))
    assert (minInl, H) == (100, 35)
    tables = [oracle.index_table(int(s), n, 4, H) for s in b["seeds"]]
    orc = _oracle_all(oracle, b, C, prm, tables, per_hyp=False)
    _check_results(res, engine.split_masks(masks, offsets), orc)
    for c in range(C):
        assert res[c]["n_refines"] == orc[c]["n_refines"]


def test_score_pnp_bit_exact_small(engine, oracle):
    """CheckInliers kernel vs oracle, masks and counts, incl. ragged N (not a multiple of 32), NaN / zero /
    behind-camera poses and points on the camera plane."""
    rng = np.random.default_rng(3)
    for (H, n) in ((1, 1), (33, 31), (70, 1000), (300, 4097)):
        p = synth.scoring_stress(5000 + n, H, n)
        poses = p["poses"].copy()
        if H > 8:
            poses[1] = np.nan
            poses[2] = 0.0
            poses[3, 9:] = [0, 0, -30]          # everything behind the camera
            poses[4, :9] *= 1e-3                # degenerate scale
            poses[5, 11] = -p["poses"][5, 6:9] @ p["p3d"][0]   # point 0 exactly on the camera plane (z ~ 0)
            poses[6] = rng.normal(size=12) * 100
        max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
        counts, masks = engine.score_pnp(poses, p["p3d"], p["p2d"], max_err, p["K"])
        pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
        oc, om = oracle.pnp_score(pb, max_err, poses)
        assert (counts == oc).all(), (H, n)
        assert (capi.unpack_mask(masks, n) == om.astype(bool)).all(), (H, n)
        ex = engine.score_exact_evals()
        assert 0 <= ex <= H * n


def test_score_pnp_adversarial_camera_plane_and_threshold(engine, oracle):
    """The fast tier has no separate guard for points near the camera plane: its certainty test is
    |D| > band|z| + eps2.  Stress exactly that regime -- camera centres placed on / next to world points
    (x, y, z all ~ 0), points a hair in front of or behind the camera plane, poses scaled up and down by
    2^+-20, and thresholds nudged so that many evaluations sit within an ulp of the decision boundary."""
    rng = np.random.default_rng(77)
    H, n = 192, 640
    p = synth.scoring_stress(5900, H, n)
    P3 = p["p3d"].astype(np.float32)
    poses = p["poses"].copy()
    for h in range(H):
        R = poses[h, :9].reshape(3, 3).astype(np.float64)
        k = rng.integers(0, n)
        mode = h % 6
        if mode == 0:      # camera centre at world point k, offset by 10^-e metres
            d = rng.normal(size=3) * 10.0 ** (-rng.integers(0, 9))
            poses[h, 9:] = (-(R @ P3[k].astype(np.float64)) + d).astype(np.float32)
        elif mode == 1:    # point k on the camera plane up to rounding, others wherever they fall
            t = poses[h, 9:].astype(np.float64)
            t[2] = -(R[2] @ P3[k].astype(np.float64)) * (1.0 + rng.normal() * 1e-7)
            poses[h, 9:] = t.astype(np.float32)
        elif mode == 2:    # whole pose scaled: K[R|t]/B must be scale-free
            poses[h] *= np.float32(2.0 ** rng.integers(-20, 21))
        elif mode == 3:    # huge sideways translation: z stays small against B
            poses[h, 9] += np.float32(10.0 ** rng.integers(2, 7))
        elif mode == 4:    # tiny rotation rows, large t_z
            poses[h, :9] *= np.float32(1e-4)
            poses[h, 11] = np.float32(rng.uniform(1, 50))
        # mode 5: unchanged (near ground truth)
    # thresholds set to the reference's own error for one pose => evaluations exactly at / one ulp off the boundary
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
    ref = poses[5].astype(np.float64)
    Xc = (P3.astype(np.float32) @ poses[5, :9].reshape(3, 3).T.astype(np.float32)) + poses[5, 9:]
    with np.errstate(all="ignore"):
        inv = np.float32(1) / Xc[:, 2]
        ue = (p["K"][2] + p["K"][0] * Xc[:, 0].astype(np.float64) * inv.astype(np.float64)).astype(np.float32)
        ve = (p["K"][3] + p["K"][1] * Xc[:, 1].astype(np.float64) * inv.astype(np.float64)).astype(np.float32)
        e = ((ue - p["p2d"][:, 0]) ** 2 + (ve - p["p2d"][:, 1]) ** 2).astype(np.float32)
    sel = np.isfinite(e) & (e > 0) & (e < 1e4)
    nudged = max_err.copy()
    nudged[sel] = np.nextafter(e[sel], np.where(rng.random(sel.sum()) < 0.5, np.float32(0), np.float32(np.inf))).astype(np.float32)
    nudged[::7] = e[::7] if np.isfinite(e[::7]).all() else nudged[::7]
    for thr in (max_err, nudged):
        counts, masks = engine.score_pnp(poses, p["p3d"], p["p2d"], thr, p["K"])
        oc, om = oracle.pnp_score(pb, thr, poses)
        assert (counts == oc).all()
        assert (capi.unpack_mask(masks, n) == om.astype(bool)).all()


def test_pnp_many_hypotheses_multiple_tiles_per_problem(engine, oracle):
    """maxIterations = 2500 with eps = 0.08 keeps H > 1024, so every problem spans several hypothesis tiles of the
    scoring kernel, and 160 problems give more (problem x tile) groups than there are CTAs: the round-robin
    work lists, tile tails (H not a multiple of the tile) and ragged n are all exercised; per-hypothesis counts
    must equal the oracle's."""
    prm = dict(prob=0.99, min_inliers=10, max_its=2500, min_set=4, eps=0.08, th2=5.991)
    sizes = [64 + 7 * (i % 9) for i in range(160)]
    parts = [synth.pnp_problem(9100 + i, n, 0.5) for i, n in enumerate(sizes)]
    p3d = np.concatenate([p["p3d"] for p in parts])
    p2d = np.concatenate([p["p2d"] for p in parts])
    s2 = np.concatenate([p["sigma2"] for p in parts])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    seeds = np.arange(len(sizes), dtype=np.uint32) + 500
    res, masks = engine.pnp_solve(offsets, p3d, p2d, s2, [parts[0]["K"]], capi.ransac_params(**prm), seeds=seeds)
    poses, counts = engine.pnp_hypotheses()
    ml = engine.split_masks(masks, offsets)
    h0 = 0
    checked = 0
    for c, n in enumerate(sizes):
        minInl, H = capi.pnp_ransac_setup(n, capi.ransac_params(**prm))
        assert H > 1024
        if c % 8 == 0:            # the oracle's exhaustive pass over 1000+ hypotheses is slow: every 8th problem
            pb = oracle.pnp_problem(parts[c]["p3d"], parts[c]["p2d"], parts[c]["sigma2"], parts[c]["K"])
            o = oracle.pnp_ransac(pb, oracle.params(**prm), oracle.index_table(int(seeds[c]), n, 4, H),
                                  oracle.FLAG_EXHAUSTIVE | oracle.FLAG_EPNP_QR_NULLSPACE, per_hyp=True)
            assert (counts[h0:h0 + H] == o["hyp_counts"]).all(), c
            _check_results(res[c:c + 1], ml[c:c + 1], [o])
            checked += 1
        h0 += H
    assert checked == 20 and h0 == len(counts)


# ---------------------------------------------------------------- early exit in phases (RSAC_FLAG_EARLY_EXIT)
_REC_FIELDS = ("ok", "no_more", "n_inliers", "best_hyp", "refined", "n_refines", "best_count", "n_hyp")


def _same_records(a, b):
    for f in _REC_FIELDS:
        assert (a[f] == b[f]).all(), (f, np.argwhere(a[f] != b[f]).ravel()[:8])
    assert (a["R"].view(np.uint32) == b["R"].view(np.uint32)).all()
    assert (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all()


@pytest.mark.parametrize("first_phase", [8, 32, 56, 299])
def test_pnp_early_exit_equals_exhaustive_and_oracle(engine, oracle, first_phase):
    """The phased run (first `first_phase` hypotheses of every problem, the rest only where still needed) must
    return the exhaustive run's records and masks bit for bit, and the sequential oracle's (early exit is what
    the reference does: PnPsolver.cpp:225-236)."""
    C, n = 96, 500
    b, offsets = _batch(4, C, n)
    args = (offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**PRM))
    res0, masks0 = engine.pnp_solve(*args, seeds=b["seeds"])
    engine.set_first_phase(first_phase)
    try:
        res1, masks1 = engine.pnp_solve(*args, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
        ha, nB, nC, solved = engine.pnp_phase_stats()
    finally:
        engine.set_first_phase(0)
    assert ha == first_phase
    _same_records(res0, res1)
    assert (masks0 == masks1).all()
    assert 0 <= nB <= C and solved <= C * 300
    if first_phase <= 56:
        assert solved < C * 300          # work was actually skipped
    tables = [oracle.index_table(int(s), n, 4, 300) for s in b["seeds"][:16]]
    orc = [oracle.pnp_ransac(oracle.pnp_problem(b["p3d"][c], b["p2d"][c], b["sigma2"][c], b["K"]), oracle.params(**PRM),
                             tables[c], oracle.FLAG_EPNP_QR_NULLSPACE) for c in range(16)]
    _check_results(res1[:16], engine.split_masks(masks1, offsets)[:16], orc)
    for c in range(16):
        assert res1[c]["n_refines"] == orc[c]["n_refines"] and res1[c]["n_hyp"] == orc[c]["n_hyp"]


def test_pnp_early_exit_failed_refines_reach_the_cleanup_phase(engine, oracle):
    """Tracking's parameters (H = 35, minInl = N/2): refines fail, so problems predicted to finish inside the
    first phase do not; the replay hands them to the clean-up phase, which must continue the scan exactly where
    the sequential reference would be (same refine count, same result)."""
    prm = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.5, th2=5.991)
    C, n = 192, 200
    b, offsets = _batch(12, C, n, outl=0.45)
    args = (offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**prm))
    res0, masks0 = engine.pnp_solve(*args, seeds=b["seeds"])
    seen_c = 0
    for first_phase in (4, 9, 20):
        engine.set_first_phase(first_phase)
        try:
            res1, masks1 = engine.pnp_solve(*args, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
            ha, nB, nC, solved = engine.pnp_phase_stats()
        finally:
            engine.set_first_phase(0)
        _same_records(res0, res1)
        assert (masks0 == masks1).all()
        assert (res1["reserved"] == 0).all()      # nothing left undecided
        seen_c += nC
    assert seen_c > 0, "the clean-up phase was never exercised"
    minInl, H = capi.pnp_ransac_setup(n, capi.ransac_params(**prm))
    orc = [oracle.pnp_ransac(oracle.pnp_problem(b["p3d"][c], b["p2d"][c], b["sigma2"][c], b["K"]), oracle.params(**prm),
                             oracle.index_table(int(b["seeds"][c]), n, 4, H), oracle.FLAG_EPNP_QR_NULLSPACE) for c in range(24)]
    _check_results(res1[:24], engine.split_masks(masks1, offsets)[:24], orc)
    for c in range(24):
        assert res1[c]["n_refines"] == orc[c]["n_refines"]


@pytest.mark.parametrize("phases", [(6, 13), (20, 21), (40, 120), (55, 300)])
def test_pnp_early_exit_three_stages(engine, phases):
    """explicit stage boundaries [0, first), [first, second), [second, H): both device-side lists (second and third
    stage) are exercised on a hard batch (65 % outliers: many candidates go on), and with Tracking's parameters
    (failing refines after the second stage as well)"""
    first, second = phases
    for (cfgno, C, n, outl, prm) in ((15, 128, 300, 0.65, PRM),
                                     (12, 96, 200, 0.45, dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.5, th2=5.991))):
        b, offsets = _batch(cfgno, C, n, outl=outl)
        args = (offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**prm))
        res0, masks0 = engine.pnp_solve(*args, seeds=b["seeds"])
        engine.set_phases(first, second)
        try:
            res1, masks1 = engine.pnp_solve(*args, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
            ha, nB, nC, solved = engine.pnp_phase_stats()
        finally:
            engine.set_phases(0, 0)
        _same_records(res0, res1)
        assert (masks0 == masks1).all()
        assert (res1["reserved"] == 0).all()
        H = capi.pnp_ransac_setup(n, capi.ransac_params(**prm))[1]
        assert solved <= C * H


def test_pnp_early_exit_ragged_batch(engine):
    """ragged problems (empty, n < minInliers, H = 0, H < first_phase) through the phased run"""
    sizes = [0, 3, 4, 9, 37, 64, 65, 500, 257, 130]
    parts = [synth.pnp_problem(7100 + i, max(n, 1), 0.4 if i != 4 else 1.0) for i, n in enumerate(sizes)]
    p3d = np.concatenate([p["p3d"][:n] for p, n in zip(parts, sizes)])
    p2d = np.concatenate([p["p2d"][:n] for p, n in zip(parts, sizes)])
    s2 = np.concatenate([p["sigma2"][:n] for p, n in zip(parts, sizes)])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    seeds = np.arange(len(sizes), dtype=np.uint32) + 177
    prms = [capi.ransac_params(**PRM), capi.ransac_params(prob=0.99, min_inliers=10, max_its=20, min_set=4, eps=0.5, th2=5.991)]
    for prm in prms:
        args = (offsets, p3d, p2d, s2, [parts[0]["K"]], prm)
        res0, masks0 = engine.pnp_solve(*args, seeds=seeds)
        for first_phase in (1, 7, 40):
            engine.set_first_phase(first_phase)
            try:
                res1, masks1 = engine.pnp_solve(*args, seeds=seeds, flags=capi.FLAG_EARLY_EXIT)
            finally:
                engine.set_first_phase(0)
            _same_records(res0, res1)
            assert (masks0 == masks1).all()


def test_pnp_early_exit_then_later_iterate_calls(engine):
    """rsac_pnp_rerun after a phased run: the hypotheses that were skipped are computed on demand, and the scan
    continues like after an exhaustive run (Tracking.cpp:1239-1334 keeps calling iterate on a candidate whose
    pose failed PoseOptimization)."""
    C, n = 64, 400
    b, offsets = _batch(14, C, n)
    args = (offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**PRM))
    res0, _ = engine.pnp_solve(*args, seeds=b["seeds"])
    resume = res0["n_hyp"].astype(np.int32)
    engine.pnp_rerun(resume)
    ref, refm = engine.pnp_download()
    engine.set_first_phase(24)
    try:
        res1, _ = engine.pnp_solve(*args, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
        _same_records(res0, res1)
        engine.pnp_rerun(resume)
        got, gotm = engine.pnp_download()
        _same_records(ref, got)
        assert (refm == gotm).all()
        engine.pnp_rerun(got["n_hyp"].astype(np.int32))       # and once more, now with everything computed
        engine.pnp_download()
    finally:
        engine.set_first_phase(0)
    poses1, counts1 = engine.pnp_hypotheses()
    engine.pnp_solve(*args, seeds=b["seeds"])
    poses0, counts0 = engine.pnp_hypotheses()
    assert (counts0 == counts1).all()                          # after the rerun every hypothesis exists
    assert (poses0.view(np.uint32) == poses1.view(np.uint32)).all()


def test_pnp_early_exit_many_hypotheses_several_tiles(engine):
    """H > 1024 per problem: the second phase spans several hypothesis tiles per problem (list-driven scoring with
    more than one record per problem), ragged n; a hard batch (70 % outliers) so that many problems need it."""
    prm = capi.ransac_params(prob=0.99, min_inliers=10, max_its=2500, min_set=4, eps=0.08, th2=5.991)
    sizes = [64 + 7 * (i % 9) for i in range(96)]
    parts = [synth.pnp_problem(9300 + i, n, 0.7) for i, n in enumerate(sizes)]
    p3d = np.concatenate([p["p3d"] for p in parts])
    p2d = np.concatenate([p["p2d"] for p in parts])
    s2 = np.concatenate([p["sigma2"] for p in parts])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    seeds = np.arange(len(sizes), dtype=np.uint32) + 900
    args = (offsets, p3d, p2d, s2, [parts[0]["K"]], prm)
    res0, masks0 = engine.pnp_solve(*args, seeds=seeds)
    seen_b = 0
    for first_phase in (16, 100, 1500):
        engine.set_first_phase(first_phase)
        try:
            res1, masks1 = engine.pnp_solve(*args, seeds=seeds, flags=capi.FLAG_EARLY_EXIT)
            ha, nB, nC, solved = engine.pnp_phase_stats()
        finally:
            engine.set_first_phase(0)
        assert ha == first_phase
        _same_records(res0, res1)
        assert (masks0 == masks1).all()
        seen_b += nB
    assert seen_b > 0


@pytest.mark.parametrize("bounds", [[3, 6, 12, 24, 48, 96, 192], [10, 11, 12, 250], [64, 128]])
def test_pnp_early_exit_many_stages(engine, bounds):
    """up to eight stages: the candidate lists alternate between two device buffers, boundaries beyond H are clipped
    (Tracking's parameters give H = 35); records must not depend on the staging"""
    for (cfgno, C, n, outl, prm) in ((16, 160, 250, 0.6, PRM),
                                     (12, 64, 200, 0.45, dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.5, th2=5.991))):
        b, offsets = _batch(cfgno, C, n, outl=outl)
        args = (offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], capi.ransac_params(**prm))
        res0, masks0 = engine.pnp_solve(*args, seeds=b["seeds"])
        engine.set_stages(bounds)
        try:
            res1, masks1 = engine.pnp_solve(*args, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
            ha, nB, nC, solved = engine.pnp_phase_stats()
            engine.pnp_rerun(res1["n_hyp"].astype(np.int32))
            after, _ = engine.pnp_download()
        finally:
            engine.set_stages([])
        H = capi.pnp_ransac_setup(n, capi.ransac_params(**prm))[1]
        assert ha == (bounds[0] if bounds[0] < H else 0)      # a first stage that covers H is the exhaustive run
        _same_records(res0, res1)
        assert (masks0 == masks1).all()
        engine.pnp_solve(*args, seeds=b["seeds"])
        engine.pnp_rerun(res0["n_hyp"].astype(np.int32))
        ref, _ = engine.pnp_download()
        _same_records(ref, after)


def test_pnp_indexed_wire_format_equals_flat_upload(engine):
    """rsac_pnp_upload_indexed (6 B per correspondence over resident keypoint / map-point tables) must give the flat upload's
    records and masks bit for bit; a second sweep re-uses the resident tables; an out-of-range index becomes an outlier."""
    C, n = 48, 500
    f = synth.reloc_frame(3, C, n_kp=2000, n_match=n, n_map=50000)
    offsets = (np.arange(C + 1) * n).astype(np.int32)
    prm = capi.ransac_params(**PRM)
    res0, m0 = engine.pnp_solve(offsets, f["p3d"], f["p2d"], f["sigma2"], [f["K"]], prm, seeds=f["seeds"], flags=capi.FLAG_EARLY_EXIT)
    engine.pnp_upload_indexed(offsets, f["kp_idx"], f["mp_idx"], f["K"], prm, seeds=f["seeds"], kp_uv=f["kp_uv"], kp_sigma2=f["kp_sigma2"],
                              mp_xyz=f["mp_xyz"])
    engine.pnp_run(capi.FLAG_EARLY_EXIT)
    res1, m1 = engine.pnp_download()
    _same_records(res0, res1)
    assert (m0 == m1).all()
    assert int(res1["ok"].sum()) >= C - 1 and np.abs(res1[0]["R"].reshape(3, 3) - f["R"]).max() < 0.02
    # the second sweep of the same frame: tables stay resident, only the pairs travel (other candidates: a permutation)
    perm = np.random.default_rng(0).permutation(C)
    engine.pnp_upload_indexed(offsets, f["kp_idx"][perm], f["mp_idx"][perm], f["K"], prm, seeds=f["seeds"][perm])
    engine.pnp_run(capi.FLAG_EARLY_EXIT)
    res2, m2 = engine.pnp_download()
    for fld in _REC_FIELDS:
        assert (res2[fld] == res0[fld][perm]).all(), fld
    # an index outside the tables is never dereferenced: that correspondence is a certain outlier
    bad = f["mp_idx"].copy()
    bad[0, :5] = 50000 + 17
    engine.pnp_upload_indexed(offsets[:2], f["kp_idx"][:1], bad[:1], f["K"], prm, seeds=f["seeds"][:1])
    engine.pnp_run(0)
    res3, m3 = engine.pnp_download()
    assert not capi.unpack_mask(m3[:16], n)[:5].any()


def test_pnp_indexed_fused_pack_equals_flat_pack(engine, oracle):
    """The indexed batch is packed by one kernel straight from the index pairs (per-keypoint table of thresholds and bound
    factors + two f64 products per correspondence).  Its packed records must be the flat path's bit for bit: same records
    and masks, same per-hypothesis counts, and the same NUMBER of evaluations sent to the exact tier (the bounds decide that
    and nothing else).  Also: the chained PoseOptimization still finds the flat arrays."""
    C, n = 24, 411                      # ragged tail in every 32-correspondence word
    f = synth.reloc_frame(5, C, n_kp=1500, n_match=n, n_map=30000)
    offsets = (np.arange(C + 1) * n).astype(np.int32)
    prm = capi.ransac_params(**PRM)
    res0, m0 = engine.pnp_solve(offsets, f["p3d"], f["p2d"], f["sigma2"], [f["K"]], prm, seeds=f["seeds"], flags=0)
    ex0 = engine.score_exact_evals()
    _, cnt0 = engine.pnp_hypotheses()
    engine.poseopt_from_pnp()
    engine.poseopt_run()
    po0, fl0 = engine.poseopt_download()
    engine.pnp_upload_indexed(offsets, f["kp_idx"], f["mp_idx"], f["K"], prm, seeds=f["seeds"], kp_uv=f["kp_uv"], kp_sigma2=f["kp_sigma2"],
                              mp_xyz=f["mp_xyz"])
    engine.pnp_run(0)
    res1, m1 = engine.pnp_download()
    ex1 = engine.score_exact_evals()
    _, cnt1 = engine.pnp_hypotheses()
    _same_records(res0, res1)
    assert (m0 == m1).all() and (cnt0 == cnt1).all()
    assert ex0 == ex1 and ex0 > 0, (ex0, ex1)
    engine.poseopt_from_pnp()
    engine.poseopt_run()
    po1, fl1 = engine.poseopt_download()
    assert po0.tobytes() == po1.tobytes() and (fl0 == fl1).all()
