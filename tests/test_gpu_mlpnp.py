"""GPU parity: batched MLPnPsolver (CUDA, through the C ABI) vs the CPU oracle.

MLPnP uses sin/cos/acos/pow; the CUDA math library and glibc differ by <= 2 ulp there, so
hypothesis poses agree to ~1e-12 relative rather than bit for bit (measured: max 7e-12).
Tolerances (BASELINE.json north_star): per-hypothesis pose 1e-9 relative (well inside the 1e-4
allowed for R/t), inlier counts / masks exact except for correspondences whose squared error
lies within 1e-6 (relative) of the chi-square threshold."""
import numpy as np
import pytest

from ransac_b200 import capi, synth

pytestmark = pytest.mark.gpu

PRM = dict(prob=0.99, min_inliers=10, max_its=300, min_set=6, eps=0.2, th2=5.991)   # cfg2: H = 300, minInl = N/5


def _near_threshold(oracle, pb, thr, T):
    """correspondences whose oracle error is within 1e-6 relative of the threshold for pose T"""
    _, _, e2 = oracle.mlpnp_check_inliers(pb, thr, T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64))
    return np.abs(e2 - thr) <= 1e-6 * thr


def _run(engine, oracle, C, n, outl, use_cov, cfg=2, flags=0, oflags=0):
    b = synth.pnp_batch(cfg, C, n, outl)
    cov = np.stack([synth.bearing_covariances(dict(K=b["K"], sigma2=b["sigma2"][c])) for c in range(C)]) if use_cov else None
    Kf = np.array([b["K"]], np.float32)
    offsets = (np.arange(C + 1) * n).astype(np.int32)
    res, masks = engine.mlpnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], Kf, capi.ransac_params(**PRM), cov=cov,
                                    seeds=b["seeds"], flags=flags)
    poses, counts = engine.mlpnp_hypotheses()
    ml = engine.split_masks(masks, offsets)
    _, H = capi.pnp_ransac_setup(n, capi.ransac_params(**PRM))
    for c in range(C):
        pb = oracle.mlpnp_problem(b["p3d"][c], b["p2d"][c], b["sigma2"][c], tuple(Kf[0]), None if cov is None else cov[c])
        tab = oracle.index_table(int(b["seeds"][c]), n, 6, H)
        o = oracle.mlpnp_ransac(pb, oracle.params(**PRM), tab, oracle.FLAG_EXHAUSTIVE | oflags, per_hyp=True)
        gp, op = poses[c * H:(c + 1) * H], o["hyp_pose"]
        rel = np.abs(gp - op) / np.maximum(1.0, np.abs(op))
        finite = np.isfinite(op).all(axis=1)
        assert (np.isfinite(gp).all(axis=1) == finite).all()
        assert np.nanmax(rel[finite]) < 1e-9, f"frame {c}: hypothesis poses differ by {np.nanmax(rel[finite])}"
        # per-hypothesis counts.  (i) Given the SAME pose the scoring kernel is bit-exact: the oracle's CheckInliers on the
        # GPU's own hypothesis pose must return the GPU's count, for every hypothesis.  (ii) Where the GPU's count
        # differs from the oracle's (its pose differs in the last bits: libm vs CUDA sin/cos/acos), every evaluation
        # whose inlier bit flips between the two poses must lie within 1e-6 (relative) of the chi-square threshold --
        # north_star's only allowance.
        thr_c = (b["sigma2"][c] * np.float32(PRM["th2"])).astype(np.float32)
        gc = counts[c * H:(c + 1) * H]
        cd = gc - o["hyp_counts"]
        for h in range(H):
            if not finite[h]:
                assert gc[h] == o["hyp_counts"][h] == 0, f"frame {c} hyp {h}: non-finite pose must score zero"
                continue
            cg, mg, _ = oracle.mlpnp_check_inliers(pb, thr_c, gp[h, :9].reshape(3, 3), gp[h, 9:])
            assert cg == gc[h], f"frame {c} hyp {h}: scoring differs on the GPU's own pose ({gc[h]} vs {cg})"
            if cd[h] != 0:
                co, mo, e2 = oracle.mlpnp_check_inliers(pb, thr_c, op[h, :9].reshape(3, 3), op[h, 9:])
                assert co == o["hyp_counts"][h]
                flipped = mg != mo
                assert flipped.sum() >= abs(int(cd[h]))
                assert (np.abs(e2[flipped] - thr_c[flipped]) <= 1e-6 * thr_c[flipped]).all(), \
                    f"frame {c} hyp {h}: count differs by {cd[h]} away from the threshold"
        assert (cd != 0).mean() <= 0.01, f"frame {c}: {(cd != 0).sum()} of {H} per-hypothesis counts differ"
        r = res[c]
        assert r["ok"] == o["ok"] and r["no_more"] == o["no_more"] and r["best_hyp"] == o["best_hyp"] and r["refined"] == o["refined"]
        if o["ok"]:
            assert np.allclose(r["R"].reshape(3, 3), o["T"][:3, :3], rtol=1e-4, atol=1e-6)
            assert np.allclose(r["t"], o["T"][:3, 3], rtol=1e-4, atol=1e-6)
            thr = (b["sigma2"][c] * np.float32(PRM["th2"])).astype(np.float32)
            diff = ml[c] != o["mask"]
            if diff.any():
                assert _near_threshold(oracle, pb, thr, o["T"])[diff].all(), f"frame {c}: mask differs away from the threshold"
            assert abs(int(r["n_inliers"]) - int(o["n_inliers"])) <= int(diff.sum())
    return res, b


def test_mlpnp_cfg2_with_covariances(engine, oracle):
    """cfg2 shape: N=1000 matches, 50% outliers, bearing covariances (use_cov branch), 6 frames"""
    res, b = _run(engine, oracle, 6, 1000, 0.5, True)
    assert res["ok"].all()
    for c in range(6):
        assert np.abs(res[c]["R"].reshape(3, 3) - b["R"][c]).max() < 0.02


def test_mlpnp_without_covariances(engine, oracle):
    """the reference's own call: covs(1) => use_cov = false (MLPnPsolver.cpp:99)"""
    _run(engine, oracle, 4, 400, 0.4, False, cfg=21)


def test_mlpnp_refine_discard_quirk(engine, oracle):
    """Q6: MLPnPsolver::Refine as shipped never stores its pose; the flag reproduces that"""
    _run(engine, oracle, 3, 300, 0.5, False, cfg=22, flags=capi.FLAG_MLPNP_DISCARD_REFINE,
         oflags=4)


def test_mlpnp_planar_scene(engine, oracle):
    """planar branch (MLPnPsolver.cpp:354-364, 497-558): world points exactly on a plane through the origin"""
    n, C = 200, 2
    b = synth.pnp_batch(23, C, n, 0.0)
    K = b["K"]
    p3d, p2d = b["p3d"].copy(), b["p2d"].copy()
    for c in range(C):
        R, t = b["R"][c], b["t"][c]
        X = p3d[c].astype(np.float64)
        X[:, 2] = 0.0                      # z = 0 plane: rank(P P^T) == 2 exactly
        Xc = X @ R.T + t
        Xc[:, 2] = np.abs(Xc[:, 2]) + 3.0  # keep everything in front of the camera
        X = (Xc - t) @ R
        X[:, 2] = 0.0
        Xc = X @ R.T + t
        p3d[c] = X.astype(np.float32)
        Xc = p3d[c].astype(np.float64) @ R.T + t
        p2d[c] = np.stack([K[0] * Xc[:, 0] / Xc[:, 2] + K[2], K[1] * Xc[:, 1] / Xc[:, 2] + K[3]], 1).astype(np.float32)
    Kf = np.array([K], np.float32)
    offsets = (np.arange(C + 1) * n).astype(np.int32)
    res, masks = engine.mlpnp_solve(offsets, p3d, p2d, b["sigma2"], Kf, capi.ransac_params(**PRM), seeds=b["seeds"])
    poses, counts = engine.mlpnp_hypotheses()
    _, H = capi.pnp_ransac_setup(n, capi.ransac_params(**PRM))
    for c in range(C):
        pb = oracle.mlpnp_problem(p3d[c], p2d[c], b["sigma2"][c], tuple(Kf[0]))
        o = oracle.mlpnp_ransac(pb, oracle.params(**PRM), oracle.index_table(int(b["seeds"][c]), n, 6, H),
                                oracle.FLAG_EXHAUSTIVE, per_hyp=True)
        gp, op = poses[c * H:(c + 1) * H], o["hyp_pose"]
        finite = np.isfinite(op).all(axis=1) & np.isfinite(gp).all(axis=1)
        rel = np.abs(gp - op) / np.maximum(1.0, np.abs(op))
        assert finite.mean() > 0.5 and np.nanmax(rel[finite]) < 1e-7
        assert res[c]["ok"] == o["ok"] and res[c]["best_hyp"] == o["best_hyp"]


_REC_FIELDS = ("ok", "no_more", "n_inliers", "best_hyp", "refined", "n_refines", "best_count", "n_hyp")


def _same_records(a, b):
    for f in _REC_FIELDS:
        assert (a[f] == b[f]).all(), (f, np.argwhere(a[f] != b[f]).ravel()[:8])
    assert (a["R"].view(np.uint32) == b["R"].view(np.uint32)).all()
    assert (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all()


@pytest.mark.parametrize("stages", [(8,), (16, 48), (5, 10, 20, 40, 80, 160), (299,)])
@pytest.mark.parametrize("use_cov", [False, True])
def test_mlpnp_early_exit_equals_exhaustive(engine, stages, use_cov):
    """MLPnPsolver::iterate returns at the first successful Refine (MLPnPsolver.cpp:144-160).  The staged run (hypotheses
    [0, b0) of every frame, later stages only for the frames still without an acceptable hypothesis, replay, clean-up) must
    give the exhaustive run's records and masks bit for bit -- same device arithmetic, fewer hypotheses computed.  The batch
    holds easy frames, frames with 70 % outliers (late or no success: the clean-up is exercised) and a frame too small to run."""
    sizes = [300, 300, 257, 300, 64, 5, 300, 300]
    outl = [0.3, 0.5, 0.7, 0.72, 0.4, 0.0, 0.75, 0.5]
    parts = [synth.pnp_problem(8100 + i, max(n, 6), o) for i, (n, o) in enumerate(zip(sizes, outl))]
    p3d = np.concatenate([p["p3d"][:n] for p, n in zip(parts, sizes)])
    p2d = np.concatenate([p["p2d"][:n] for p, n in zip(parts, sizes)])
    s2 = np.concatenate([p["sigma2"][:n] for p, n in zip(parts, sizes)])
    cov = np.concatenate([synth.bearing_covariances(dict(K=p["K"], sigma2=p["sigma2"][:n])) for p, n in zip(parts, sizes)]) if use_cov else None
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    Kf = np.array([parts[0]["K"]], np.float32)
    seeds = np.arange(len(sizes), dtype=np.uint32) + 901
    prm = capi.ransac_params(**PRM)
    engine.set_stages([])
    ref, mref = engine.mlpnp_solve(offsets, p3d, p2d, s2, Kf, prm, cov=cov, seeds=seeds, flags=0)
    try:
        engine.set_stages(list(stages))
        for rep in range(3):                      # eager, captured, replayed
            res, m = engine.mlpnp_solve(offsets, p3d, p2d, s2, Kf, prm, cov=cov, seeds=seeds, flags=capi.FLAG_EARLY_EXIT) if rep == 0 else (None, None)
            if rep > 0:
                engine.mlpnp_run(capi.FLAG_EARLY_EXIT)
                res, m = engine.mlpnp_download()
            _same_records(ref, res)
            assert (mref == m).all()
        st = engine.mlpnp_phase_stats()
        assert st[0] == stages[0]
        H = sum(capi.pnp_ransac_setup(n, prm)[1] if n >= 60 else 0 for n in sizes)
        if stages[0] < 100:
            assert st[3] < H, "the staged run must skip hypotheses"
            assert st[1] >= 1, "the hard frames must go on to the second stage"
        # a later iterate() call resumes anywhere: first the skipped hypotheses are computed, then the replay resumes
        resume = np.minimum(ref["n_hyp"], 7).astype(np.int32)
        engine.mlpnp_rerun(resume, flags=capi.FLAG_EARLY_EXIT)
        r_ee, m_ee = engine.mlpnp_download()
    finally:
        engine.set_stages([])
    engine.mlpnp_solve(offsets, p3d, p2d, s2, Kf, prm, cov=cov, seeds=seeds, flags=0)
    engine.mlpnp_rerun(resume, flags=0)
    r_ex, m_ex = engine.mlpnp_download()
    _same_records(r_ex, r_ee)
    assert (m_ex == m_ee).all()
