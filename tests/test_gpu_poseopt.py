"""GPU parity: batched Optimizer::PoseOptimization (CUDA, through the C ABI) vs the CPU oracle.

The kernel sums the edges of a frame over 32 lanes and a shuffle butterfly, the oracle in edge order
(as g2o does), and sin/cos come from two maths libraries, so sums differ in the last bits: poses are compared
to 1e-7 absolute (north_star tolerance: 1e-4 relative), the step-control trajectory (LM iterations, trials) and
the outlier flags must be identical except for edges whose chi2 lies within 1e-6 (relative) of the threshold."""
import numpy as np
import pytest

from ransac_b200 import capi, synth

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-7
CHI_TOL = 1e-6
STATS = {"iter_diff": 0, "trial_diff": 0, "pose_diff": 0.0}


def _chi2(p, R, t):
    """chi2 of every edge at the pose (R, t), float64, for the near-threshold exemption"""
    Xc = p["p3d"].astype(np.float64) @ R.T + t
    fx, fy, cx, cy, bf = [float(k) for k in p["K"]]
    u = fx * Xc[:, 0] / Xc[:, 2] + cx
    v = fy * Xc[:, 1] / Xc[:, 2] + cy
    e = (p["obs"][:, 0] - u) ** 2 + (p["obs"][:, 1] - v) ** 2
    st = p["obs"][:, 2] >= 0
    e = e + np.where(st, (p["obs"][:, 2] - (u - bf / Xc[:, 2])) ** 2, 0.0)
    return e * p["inv_sigma2"].astype(np.float64), np.where(st, 7.815, 5.991)


def _run(engine, oracle, ps):
    sizes = [p["p3d"].shape[0] for p in ps]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    cat = lambda k: np.concatenate([p[k] for p in ps]) if sum(sizes) else np.zeros((0,) + ps[0][k].shape[1:], np.float32)
    K = np.stack([p["K"] for p in ps])
    Tcw = np.stack([np.concatenate([p["Rcw"].ravel(), p["tcw"]]) for p in ps])
    res, outlier = engine.poseopt_solve(offsets, cat("p3d"), cat("obs"), cat("inv_sigma2"), K, Tcw)
    near_total = 0
    for c, p in enumerate(ps):
        pb = oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"])
        o, oout = oracle.pose_optimization(pb)
        r = res[c]
        g = outlier[offsets[c]:offsets[c + 1]]
        assert r["rounds"] == o["rounds"], (c, r["rounds"], o["rounds"])
        # at a converged pose chi2(trial) - chi2(current) is rounding noise, so whether g2o accepts the last, tiny
        # steps (and hence the number of trials) is decided by the summation order: informational only
        STATS["iter_diff"] = max(STATS["iter_diff"], abs(int(r["iterations"]) - int(o["iterations"])))
        STATS["trial_diff"] = max(STATS["trial_diff"], abs(int(r["trials"]) - int(o["trials"])))
        STATS["pose_diff"] = max(STATS["pose_diff"], float(np.abs(r["R"].reshape(3, 3) - o["R"]).max()), float(np.abs(r["t"] - o["t"]).max()))
        assert np.abs(r["R"].reshape(3, 3) - o["R"]).max() < POSE_TOL, c
        assert np.abs(r["t"] - o["t"]).max() < POSE_TOL * max(1.0, np.abs(o["t"]).max()), c
        assert np.abs(r["Rf"].reshape(3, 3) - o["Rf"]).max() <= 1.2e-7 and np.abs(r["tf"] - o["tf"]).max() <= 1e-6, c
        diff = np.flatnonzero(g != oout)
        if sizes[c] >= 3 and diff.size:
            chi, thr = _chi2(p, o["R"], o["t"])
            near = np.abs(chi[diff] - thr[diff]) <= CHI_TOL * thr[diff]
            assert near.all(), (c, diff[~near][:5], chi[diff][~near][:5])
            near_total += diff.size
        assert abs(int(r["n_inliers"]) - int(o["n_inliers"])) <= diff.size, c
        assert int(r["n_inliers"]) == (sizes[c] - int(g.sum()) if sizes[c] >= 3 else 0), c
    return res, near_total


def test_poseopt_monocular_batch(engine, oracle):
    ps = [synth.poseopt_problem(9000 + i, 250, 0.2, 0.0) for i in range(48)]
    res, near = _run(engine, oracle, ps)
    assert (res["rounds"] == 4).all() and (res["n_inliers"] > 150).all()
    for r, p in zip(res, ps):       # converged to the ground truth from a 0.02 rad / 5 cm perturbation
        assert np.abs(r["R"].reshape(3, 3) - p["R"]).max() < 5e-3


def test_poseopt_stereo_and_mixed(engine, oracle):
    ps = [synth.poseopt_problem(9100 + i, 300, 0.3, sr) for i, sr in enumerate([1.0, 1.0, 0.5, 0.5, 0.2, 0.8] * 4)]
    _run(engine, oracle, ps)


def test_poseopt_ragged_and_degenerate(engine, oracle):
    sizes = [0, 1, 2, 3, 5, 9, 10, 11, 31, 32, 33, 64, 1200, 2000]
    ps = [synth.poseopt_problem(9200 + i, n, 0.25, 0.3 * (i % 3)) for i, n in enumerate(sizes)]
    res, _ = _run(engine, oracle, ps)
    assert res["rounds"][0] == 0 and res["n_inliers"][2] == 0          # fewer than 3 correspondences: return 0
    assert res["rounds"][5] == 1 and res["rounds"][6] == 4             # edges().size() < 10 stops after one round


def test_poseopt_bad_initial_pose_and_heavy_outliers(engine, oracle):
    """far initial poses exercise rejected LM trials (lambda growth, roll-back, stale errors at classification)"""
    ps = [synth.poseopt_problem(9300 + i, 200, 0.5, 0.0, pose_noise=(0.3, 1.0)) for i in range(24)]
    res, _ = _run(engine, oracle, ps)
    assert (res["trials"] > res["iterations"]).any()


def test_poseopt_points_behind_the_camera_and_non_finite_input(engine, oracle):
    """no positive-depth test in the reference (EdgeSE3ProjectXYZOnlyPose::isDepthPositive is never called): a point
    behind the camera or at the camera centre is an ordinary (huge) residual; a NaN observation or an infinite point
    makes every chi2 sum NaN, every LM step is rejected (rho > 0 is false), the pose stays where it started and the NaN
    edge counts as an inlier (`chi2 > th` is false) -- the kernel must do exactly the same."""
    ps = []
    for k in range(4):
        p = synth.poseopt_problem(9400 + k, 120, 0.2, 0.5 * (k % 2))
        cam_centre = -(p["R"].T @ p["t"])
        p["p3d"][3] = (cam_centre + p["R"].T @ np.array([0.1, 0.2, -3.0])).astype(np.float32)
        p["p3d"][7] = cam_centre.astype(np.float32)
        if k >= 2:
            p["obs"][11, 0] = np.nan
            p["p3d"][15] = np.float32(np.inf)
        ps.append(p)
    sizes = [120] * 4
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    cat = lambda k: np.concatenate([p[k] for p in ps])
    res, outlier = engine.poseopt_solve(offsets, cat("p3d"), cat("obs"), cat("inv_sigma2"), np.stack([p["K"] for p in ps]),
                                        np.stack([np.concatenate([p["Rcw"].ravel(), p["tcw"]]) for p in ps]))
    for c, p in enumerate(ps):
        o, oout = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        g = outlier[offsets[c]:offsets[c + 1]]
        assert res[c]["rounds"] == o["rounds"]
        assert np.abs(res[c]["R"].reshape(3, 3) - o["R"]).max() < POSE_TOL and np.abs(res[c]["t"] - o["t"]).max() < POSE_TOL, c
        assert (g != oout).sum() <= 1 and abs(int(res[c]["n_inliers"]) - o["n_inliers"]) <= 1, c
        assert g[3] == oout[3] and g[7] == oout[7] and g[3] == 1      # the mirrored projection of a point behind the camera is far off
        if c >= 2:
            assert res[c]["iterations"] == 40 and res[c]["trials"] == 40 and g[11] == 0


def test_poseopt_large_batch_takes_the_one_warp_per_frame_kernel(engine, oracle):
    """batches of at least 4 x SM-count frames run one warp per frame (four frames per CTA), smaller ones four warps
    per frame: both shapes must agree with the oracle (every other test here is a small batch)"""
    ps = [synth.poseopt_problem(9900 + i, 40 + (i % 7), 0.2, 0.5 * (i % 2)) for i in range(640)]
    _run(engine, oracle, ps)


def test_poseopt_empty_batch(engine):
    res, out = engine.poseopt_solve(np.zeros(1, np.int32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32),
                                    np.zeros(0, np.float32), np.zeros((0, 5), np.float32), np.zeros((0, 12), np.float32))
    assert res.shape[0] == 0 and out.shape[0] == 0
    print("poseopt GPU vs oracle:", STATS)


def test_poseopt_chained_behind_a_pnp_sweep_on_the_device(engine, oracle):
    """rsac_poseopt_from_pnp: the frames are built on the device from the PnP engine's resident correspondences, final
    inlier masks and poses (Tracking.cpp:1258-1284); result = PoseOptimization of the compacted inlier sets from the
    RANSAC poses, as the oracle computes it from the downloaded records."""
    C, n = 12, 300
    b = synth.pnp_batch(11, C, n, 0.5)
    b["p2d"][5] += 400.0                                     # one candidate that RANSAC cannot verify: empty frame
    offsets = (np.arange(C + 1) * n).astype(np.int32)
    prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
    res, masks = engine.pnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"], flags=capi.FLAG_EARLY_EXIT)
    ml = engine.split_masks(masks, offsets)
    engine.poseopt_from_pnp(bf=0.0)
    engine.poseopt_run()
    pres, flags = engine.poseopt_download()
    assert flags.shape[0] == C * n
    n_ok = 0
    for c in range(C):
        f = flags[c * n:(c + 1) * n]
        if not res[c]["ok"]:
            assert pres[c]["rounds"] == 0 and pres[c]["n_inliers"] == 0
            continue
        n_ok += 1
        m = ml[c].astype(bool)
        assert (f[~m] == 2).all() and (f[m] < 2).all()
        p = dict(p3d=b["p3d"][c][m], obs=np.concatenate([b["p2d"][c][m], np.full((int(m.sum()), 1), -1, np.float32)], axis=1),
                 isig=(np.float32(1) / b["sigma2"][c][m]).astype(np.float32), K=np.array(list(b["K"]) + [0.0], np.float32))
        o, oout = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["isig"], p["K"], res[c]["R"].reshape(3, 3), res[c]["t"]))
        assert pres[c]["rounds"] == o["rounds"]
        assert np.abs(pres[c]["R"].reshape(3, 3) - o["R"]).max() < POSE_TOL and np.abs(pres[c]["t"] - o["t"]).max() < POSE_TOL, c
        assert (f[m] != oout).sum() <= 1 and abs(int(pres[c]["n_inliers"]) - o["n_inliers"]) <= 1, c
    assert n_ok >= C - 2 and not res[5]["ok"]
