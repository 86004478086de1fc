"""CPU suite, part 1: the oracle (oracle/, a plain-C restatement of the reference's solvers) against
its pins: libc rand() known answers, an independent EPnP (cv2, n >= 6), numpy
linear algebra on the same matrices, a numpy emulation of the mixed-precision scoring expressions,
ground-truth recovery on noise-free data and the frozen regression vectors in tests/golden/."""
import json
import os

import numpy as np
import pytest

from ransac_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ------------------------------------------------------------------ RNG (R01)
def test_rng_known_answers_from_survey(oracle):
    # verified against glibc 2.39 srand(1) by the survey
    assert oracle.index_table(1, 500, 4, 3).tolist() == [[420, 196, 389, 396], [455, 98, 166, 381], [138, 276, 237, 312]]
    assert oracle.index_table(1, 200, 3, 3).tolist() == [[168, 78, 155], [159, 181, 39], [67, 152, 54]]
    assert oracle.index_table(1, 1000, 6, 2).tolist() == [[840, 393, 781, 796, 908, 196], [335, 767, 277, 552, 475, 625]]
    # RandomInt(0, 999) = int(rand()/(RAND_MAX+1.0)*1000) of the first three rand() values 1804289383 846930886 1681692777
    assert oracle.random_ints(1, 0, 999, 3) == [840, 394, 783]


def test_rng_golden_file(oracle):
    g = json.load(open(os.path.join(GOLD, "rng_known_answers.json")))
    for key, tab in g["tables"].items():
        s, n, k = (int(x) for x in key.split("_"))
        assert oracle.index_table(s, n, k, 8).tolist() == tab
    assert g["rand"]["0"] == g["rand"]["1"]          # glibc: srand(0) == srand(1)


def test_minimal_sets_are_distinct_and_in_range(oracle):
    for (n, k) in ((4, 4), (7, 6), (500, 4), (200, 3)):
        t = oracle.index_table(99, n, k, 200)
        assert t.max() < n
        assert all(len(set(row)) == k for row in t.tolist())


# ------------------------------------------------------ small dense solves vs numpy
def test_jacobi_eig_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    for n in (3, 4, 12):
        for _ in range(20):
            A = rng.normal(size=(n, n))
            A = A @ A.T * 10 ** rng.uniform(-3, 6)
            w, v = oracle.jacobi_eig(A)
            w2 = np.linalg.eigvalsh(A)
            assert np.allclose(w, w2, rtol=1e-10, atol=1e-12 * abs(w2).max())
            assert np.abs(A @ v - v * w).max() < 1e-10 * abs(w2).max()
            assert np.abs(v.T @ v - np.eye(n)).max() < 1e-12
    for _ in range(20):
        A = rng.normal(size=(4, 4)).astype(np.float32)
        A = (A @ A.T).astype(np.float32)
        w, v = oracle.jacobi_eig(A, np.float32)
        assert np.allclose(w, np.linalg.eigvalsh(A.astype(np.float64)), rtol=2e-5, atol=2e-5)


def test_jacobi_lowest_null_space(oracle):
    """EPnP n=4: M is 8x12, MtM has an exactly 4-dimensional null space: the four returned
    vectors must be orthonormal and annihilated by M"""
    rng = np.random.default_rng(1)
    for _ in range(30):
        M = rng.normal(size=(8, 12)) * 100
        w, v = oracle.jacobi_lowest(M.T @ M, 4)
        assert np.abs(v.T @ v - np.eye(4)).max() < 1e-12
        assert np.abs(M @ v).max() < 1e-9 * np.abs(M).max() * 12
    for _ in range(30):       # generic SPD: the 4 smallest eigenpairs
        A = rng.normal(size=(12, 12))
        A = A @ A.T
        w, v = oracle.jacobi_lowest(A, 4)
        assert np.allclose(w, np.linalg.eigvalsh(A)[:4], rtol=1e-9, atol=1e-12)
        assert np.abs(A @ v - v * w).max() < 1e-10 * np.linalg.norm(A)


def test_svd_lstsq_matches_numpy_including_rank_deficient(oracle):
    rng = np.random.default_rng(2)
    for k in (3, 4, 5):
        for _ in range(20):
            L = rng.normal(size=(6, k))
            b = rng.normal(size=6)
            assert np.allclose(oracle.svd_lstsq(L, b), np.linalg.lstsq(L, b, rcond=None)[0], rtol=1e-9, atol=1e-11)
        L = rng.normal(size=(6, k))
        L[:, -1] = L[:, 0] * 2.0          # rank deficient: minimum-norm solution (Eigen bdcSvd().solve semantics)
        b = rng.normal(size=6)
        assert np.allclose(oracle.svd_lstsq(L, b), np.linalg.lstsq(L, b, rcond=None)[0], rtol=1e-8, atol=1e-10)


def test_inv3_polar3_rank3_ldlt6(oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = rng.normal(size=(3, 3))
        assert np.allclose(oracle.inv3(m), np.linalg.inv(m), rtol=1e-9, atol=1e-11)
        u, _, vt = np.linalg.svd(m)
        assert np.allclose(oracle.polar3(m), u @ vt, atol=1e-11)
        a = rng.normal(size=(6, 6))
        a = a @ a.T + np.eye(6)
        g = rng.normal(size=6)
        assert np.allclose(oracle.ldlt6_solve(a, g), np.linalg.solve(a, g), rtol=1e-9, atol=1e-11)
    P = rng.normal(size=(3, 10))
    assert oracle.rank3(P @ P.T) == 3
    P[2] = 0.0
    assert oracle.rank3(P @ P.T) == 2      # the planarity test of MLPnPsolver.cpp:346-354
    assert oracle.rank3(np.zeros((3, 3))) == 0


# ------------------------------------------------------------------- solvers
def test_epnp_recovers_ground_truth_and_matches_cv2(oracle):
    g = np.load(os.path.join(GOLD, "cv2_epnp.npz"))
    for n in (6, 50):
        for seed in (7, 8, 9):
            p = synth.pnp_problem(seed, n, 0.0, noise=False)
            pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
            R, t, err = oracle.epnp_pose(pb, np.arange(n))
            assert err < 1e-3                                   # inputs are f32: ~1e-5 px
            assert np.abs(R - p["R"]).max() < 1e-5 and np.abs(t - p["t"]).max() < 1e-4
            assert np.abs(R - g[f"R_{n}_{seed}"]).max() < 1e-5 and np.abs(t - g[f"t_{n}_{seed}"]).max() < 1e-4


def test_horn_recovers_ground_truth_with_and_without_scale(oracle):
    rng = np.random.default_rng(4)
    for s in (1.0, 0.5, 1.7):
        R, t = synth.random_pose(rng)
        P2 = rng.normal(size=(3, 3)) * 3 + np.array([0, 0, 8.0])
        P1 = (s * (R @ P2.T)).T + t
        Rr, tr, sr = oracle.sim3_compute(P1, P2, fix_scale=(s == 1.0))
        assert np.abs(Rr - R).max() < 3e-5 and np.abs(tr - t).max() < 3e-4 and abs(sr - s) < 1e-4   # f32 Horn on 3 points


def test_mlpnp_recovers_ground_truth_all_branches(oracle):
    for n in (6, 50):
        p = synth.pnp_problem(11, n, 0.0, noise=False)
        K = tuple(np.float32(k) for k in p["K"])
        for cov in (None, synth.bearing_covariances(p)):
            pb = oracle.mlpnp_problem(p["p3d"], p["p2d"], p["sigma2"], K, cov)
            R, t = oracle.mlpnp_pose(pb, np.arange(n))
            assert np.abs(R - p["R"]).max() < 1e-5 and np.abs(t - p["t"]).max() < 1e-4
    # planar branch
    p = synth.pnp_problem(12, 40, 0.0, noise=False)
    X = p["p3d"].astype(np.float64)
    X[:, 2] = 0.0
    Xc = X @ p["R"].T + p["t"]
    keep = Xc[:, 2] > 0.5
    X, Xc = X[keep], Xc[keep]
    uv = np.stack([p["K"][0] * Xc[:, 0] / Xc[:, 2] + p["K"][2], p["K"][1] * Xc[:, 1] / Xc[:, 2] + p["K"][3]], 1)
    pb = oracle.mlpnp_problem(X.astype(np.float32), uv.astype(np.float32), p["sigma2"][keep], tuple(np.float32(k) for k in p["K"]))
    R, t = oracle.mlpnp_pose(pb, np.arange(len(X)))
    assert np.abs(R - p["R"]).max() < 1e-4 and np.abs(t - p["t"]).max() < 1e-3


def test_mlpnp_jacobian_is_the_derivative_of_the_residual(oracle):
    rng = np.random.default_rng(5)
    for _ in range(10):
        pt = rng.normal(size=3) + [0, 0, 5]
        w = rng.normal(size=3) * 0.4
        t = rng.normal(size=3)
        f = np.array([rng.normal() * 0.3, rng.normal() * 0.3, 1.0])
        nr = np.cross(f, [1, 0, 0]); nr /= np.linalg.norm(nr)
        ns = np.cross(f, nr); ns /= np.linalg.norm(ns)
        _, J = oracle.mlpnp_res_jac(pt, nr, ns, w, t)
        x = np.concatenate([w, t])
        Jn = np.zeros((2, 6))
        for j in range(6):
            d = np.zeros(6); d[j] = 1e-6
            rp, _ = oracle.mlpnp_res_jac(pt, nr, ns, (x + d)[:3], (x + d)[3:])
            rm, _ = oracle.mlpnp_res_jac(pt, nr, ns, (x - d)[:3], (x - d)[3:])
            Jn[:, j] = (rp - rm) / 2e-6
        assert np.abs(J - Jn).max() < 1e-8


def test_rodrigues_round_trip(oracle):
    rng = np.random.default_rng(6)
    for _ in range(10):
        w = rng.normal(size=3)
        w *= rng.uniform(0.01, 3.0) / np.linalg.norm(w)
        assert np.allclose(oracle.rot2rodrigues(oracle.rodrigues2rot(w)), w, atol=1e-9)
        assert np.allclose(oracle.rodrigues2rot(w), synth.rodrigues(w), atol=1e-12)


# ------------------------------------------------------------------- scoring
def test_scoring_matches_numpy_mixed_precision_emulation(oracle):
    g = np.load(os.path.join(GOLD, "scoring_numpy.npz"))
    p = synth.scoring_stress(int(g["seed"]), int(g["H"]), int(g["n"]))
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    mb = oracle.mlpnp_problem(p["p3d"], p["p2d"], p["sigma2"], tuple(np.float32(k) for k in p["K"]))
    thr = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
    for h in range(int(g["H"])):
        R, t = p["poses"][h, :9].reshape(3, 3), p["poses"][h, 9:]
        cnt, mask, e2 = oracle.pnp_check_inliers(pb, thr, R, t)
        assert (e2.view(np.uint32) == g["pnp_err2"][h].view(np.uint32)).all()       # bit-exact
        assert (mask == (g["pnp_err2"][h] < thr)).all() and cnt == mask.sum()
        cnt, mask, e2 = oracle.mlpnp_check_inliers(mb, thr, R.astype(np.float64), t.astype(np.float64))
        assert (e2.view(np.uint32) == g["mlpnp_err2"][h].view(np.uint32)).all()
        assert (mask == (g["mlpnp_err2"][h] < thr)).all()


def test_scoring_nan_and_behind_camera(oracle):
    p = synth.pnp_problem(77, 64, 0.0)
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    thr = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
    cnt, _, _ = oracle.pnp_check_inliers(pb, thr, np.full((3, 3), np.nan), np.zeros(3))
    assert cnt == 0                                                    # NaN pose scores zero (no exception)
    cnt, mask, _ = oracle.pnp_check_inliers(pb, thr, p["R"], p["t"])
    assert cnt >= 55      # chi-square 95% gate on 64 noisy inliers
    # no positive-depth test in the reference (PnPsolver.cpp:241-268): mirrored geometry behind the camera counts
    cnt2, _, _ = oracle.pnp_check_inliers(pb, thr, -p["R"], -p["t"])
    assert cnt2 == cnt


def test_sim3_thresholds_are_truncated_to_integers(oracle):
    # Q4: 9.210*sigma2 stored in vector<size_t>: levels 0..7 -> 9,13,19,27,39,57,82,118
    s2 = synth.level_sigma2()
    assert [int(9.210 * float(x)) for x in s2] == [9, 13, 19, 27, 39, 57, 82, 118]
    q = synth.sim3_problem(31, 50, 0.0)
    pb = oracle.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"])
    cnt, mask, err = oracle.sim3_check_inliers(pb, q["R12"], q["t12"])
    thr1 = np.floor(9.210 * q["sigma2_1"].astype(np.float64)).astype(np.float32)
    thr2 = np.floor(9.210 * q["sigma2_2"].astype(np.float64)).astype(np.float32)
    assert (mask == ((err[:, 0] < thr1) & (err[:, 1] < thr2))).all() and cnt > 40


# ------------------------------------------------------------ RANSAC semantics
def test_ransac_parameter_arithmetic(oracle):
    # Tracking.cpp:1226: (0.99,10,300,4,0.5,5.991) -> H = 35, minInl = N/2; cfg1 parameters -> H = 300, minInl = 100
    assert oracle.ransac_setup_pnp(500, oracle.params(0.99, 10, 300, 4, 0.5, 5.991)) == (250, 35)
    assert oracle.ransac_setup_pnp(500, oracle.params(0.99, 10, 300, 4, 0.2, 5.991)) == (100, 300)
    assert oracle.ransac_setup_pnp(1000, oracle.params(0.99, 10, 300, 6, 0.2, 5.991)) == (200, 300)   # eps^3 for minSet 6 too (Q10)
    assert oracle.ransac_setup_pnp(10, oracle.params(0.99, 10, 300, 4, 0.5, 5.991)) == (10, 1)        # minInl == N
    assert oracle.ransac_setup_sim3(200, 0.99, 20, 300) == 300                                         # LoopClosing.cpp:261
    assert oracle.ransac_setup_sim3(20, 0.99, 20, 300) == 1


def test_pnp_sequential_semantics(oracle):
    """early exit vs exhaustive give the same outcome; strict '>' keeps the first maximum; n < minInl => bNoMore"""
    p = synth.pnp_problem(1000, 500, 0.5)
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    prm = oracle.params(0.99, 10, 300, 4, 0.2, 5.991)
    tab = oracle.index_table(1000, 500, 4, 300)
    a = oracle.pnp_ransac(pb, prm, tab)
    b = oracle.pnp_ransac(pb, prm, tab, oracle.FLAG_EXHAUSTIVE, per_hyp=True)
    for k in ("ok", "n_inliers", "best_hyp", "refined"):
        assert a[k] == b[k]
    assert (a["mask"] == b["mask"]).all() and a["n_hyp"] <= 300 and b["n_hyp"] == 300
    c = b["hyp_counts"]
    first = int(np.argmax(c >= 100))
    assert a["ok"] and a["refined"] and a["n_hyp"] == first + 1 and a["best_hyp"] == first
    small = oracle.pnp_problem(p["p3d"][:9], p["p2d"][:9], p["sigma2"][:9], p["K"])
    r = oracle.pnp_ransac(small, prm, np.zeros((1, 4), np.uint32))
    assert r["ok"] == 0 and r["no_more"] == 1 and r["n_inliers"] == 0


def test_pnp_stale_rows_quirk_only_matters_after_a_refine(oracle):
    """Q1: as shipped, the EPnP sums run over all ALLOCATED rows (PnPsolver.cpp:301,356,435-436).  The buffers
    only grow in Refine(), so hypotheses up to and including the first refine are identical to the clean
    (n-bounded) semantics; every later 4-point hypothesis is corrupted by stale rows.  The reference only
    reaches those when a Refine() fails and RANSAC continues; exhaustive mode exposes them."""
    prm = oracle.params(0.99, 10, 300, 4, 0.5, 5.991)     # Tracking's own parameters (Tracking.cpp:1226)
    corrupted = 0
    for seed in range(12):
        p = synth.pnp_problem(12000 + seed, 200, 0.3)
        pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
        tab = oracle.index_table(seed, 200, 4, 35)
        a = oracle.pnp_ransac(pb, prm, tab, oracle.FLAG_EXHAUSTIVE, per_hyp=True)
        b = oracle.pnp_ransac(pb, prm, tab, oracle.FLAG_EXHAUSTIVE | oracle.FLAG_STALE_ROWS, per_hyp=True)
        ca, cb = a["hyp_counts"], b["hyp_counts"]
        qualifying = np.flatnonzero(ca >= 100)
        if len(qualifying) == 0:
            assert (ca == cb).all()
            continue
        first = int(qualifying[0])
        assert (ca[:first + 1] == cb[:first + 1]).all()                    # identical until the first Refine()
        same_pose = np.array_equal(a["hyp_pose"][first + 1:].view(np.uint32), b["hyp_pose"][first + 1:].view(np.uint32))
        corrupted += int(first + 1 < 35 and not same_pose)
        # outcome of the reference-semantics run is unaffected when that first refine succeeds
        ra = oracle.pnp_ransac(pb, prm, tab)
        rb = oracle.pnp_ransac(pb, prm, tab, oracle.FLAG_STALE_ROWS)
        if ra["n_failed_refines"] == 0:
            assert ra["ok"] == rb["ok"] and ra["n_inliers"] == rb["n_inliers"] and (ra["mask"] == rb["mask"]).all()
    assert corrupted > 0


def test_sim3_sequential_semantics(oracle):
    q = synth.sim3_problem(3000, 200, 0.4)
    pb = oracle.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"])
    tab = oracle.index_table(3000, 200, 3, 300)
    a = oracle.sim3_ransac(pb, 0.99, 20, 300, tab)
    b = oracle.sim3_ransac(pb, 0.99, 20, 300, tab, oracle.FLAG_EXHAUSTIVE, per_hyp=True)
    c = b["hyp_counts"]
    first = int(np.argmax(c > 20))
    assert a["ok"] and a["best_hyp"] == first and a["n_inliers"] == c[first] and a["n_hyp"] == first + 1
    # all-outlier set: budget exhausted, best = LAST arg-max ('>=', Sim3Solver.cpp:155)
    q = synth.sim3_problem(3204, 33, 1.0)
    pb = oracle.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"])
    H = oracle.ransac_setup_sim3(33, 0.99, 20, 300)
    r = oracle.sim3_ransac(pb, 0.99, 20, 300, oracle.index_table(5, 33, 3, H), 0, per_hyp=True)
    c = r["hyp_counts"]
    assert r["ok"] == 0 and r["no_more"] == 1 and r["best_hyp"] == int(np.flatnonzero(c == c.max())[-1])


# ------------------------------------------------------------ frozen regression
def test_oracle_frozen_vectors(oracle):
    g = np.load(os.path.join(GOLD, "oracle_frozen.npz"))
    p = synth.pnp_problem(1000, 500, 0.5)
    pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    r = oracle.pnp_ransac(pb, oracle.params(0.99, 10, 300, 4, 0.2, 5.991), oracle.index_table(1000, 500, 4, 300),
                          oracle.FLAG_EXHAUSTIVE, per_hyp=True)
    assert (r["hyp_counts"] == g["cfg1_counts"]).all()
    assert np.array_equal(r["hyp_pose"].view(np.uint32), g["cfg1_pose"].view(np.uint32))
    assert [r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]] == g["cfg1_meta"].tolist()
    assert (r["mask"] == g["cfg1_mask"]).all()
    r = oracle.pnp_ransac(pb, oracle.params(0.99, 10, 300, 4, 0.2, 5.991), oracle.index_table(1000, 500, 4, 300),
                          oracle.FLAG_EXHAUSTIVE | oracle.FLAG_EPNP_QR_NULLSPACE, per_hyp=True)
    assert (r["hyp_counts"] == g["cfg1q_counts"]).all()
    assert np.array_equal(r["hyp_pose"].view(np.uint32), g["cfg1q_pose"].view(np.uint32))
    assert [r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]] == g["cfg1q_meta"].tolist()
    assert (r["mask"] == g["cfg1q_mask"]).all()
    for tag, sc in (("cfg3", 1.0), ("cfg3s", 1.6)):
        q = synth.sim3_problem(3000, 200, 0.4, sc)
        sb = oracle.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"], fix_scale=(sc == 1.0))
        r = oracle.sim3_ransac(sb, 0.99, 20, 300, oracle.index_table(3000, 200, 3, 300), oracle.FLAG_EXHAUSTIVE, per_hyp=True)
        assert (r["hyp_counts"] == g[tag + "_counts"]).all()
        assert np.array_equal(r["hyp_pose"].view(np.uint32), g[tag + "_pose"].view(np.uint32))
        assert [r["ok"], r["n_inliers"], r["best_hyp"], r["best_count"]] == g[tag + "_meta"].tolist()
    p = synth.pnp_problem(2000, 1000, 0.5)
    mb = oracle.mlpnp_problem(p["p3d"], p["p2d"], p["sigma2"], tuple(np.float32(k) for k in p["K"]), synth.bearing_covariances(p))
    r = oracle.mlpnp_ransac(mb, oracle.params(0.99, 10, 300, 6, 0.2, 5.991), oracle.index_table(2000, 1000, 6, 300),
                            oracle.FLAG_EXHAUSTIVE, per_hyp=True)
    assert (r["hyp_counts"] == g["cfg2_counts"]).all()
    assert np.allclose(r["hyp_pose"], g["cfg2_pose"], rtol=1e-12, atol=1e-14, equal_nan=True)
    assert [r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]] == g["cfg2_meta"].tolist()


# ------------------------------------------------ Optimizer::PoseOptimization
def _po(oracle, p):
    return oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))


def test_poseopt_matches_scipy_least_squares_on_outlier_free_frames(oracle):
    """every edge stays an inlier, so the last (kernel-free) round minimises the plain weighted reprojection cost:
    scipy's Levenberg-Marquardt optimum frozen in tests/golden/poseopt_scipy.npz (made by make_golden.py) is an
    independent pin.  Stereo edges use a float 1/z in the reference (types_six_dof_expmap.cpp:300), scipy a double
    one, hence the looser bound there."""
    cases = [dict(seed=9500 + i, n=n, outl=0.0, stereo=sr, noise_scale=0.3) for i, (n, sr) in
             enumerate([(60, 0.0), (200, 0.0), (200, 1.0), (400, 0.5)])]     # = make_golden._poseopt_cases()
    g = np.load(os.path.join(GOLD, "poseopt_scipy.npz"))
    for k, c in enumerate(cases):
        p = synth.poseopt_problem(c["seed"], c["n"], c["outl"], c["stereo"], noise_scale=c["noise_scale"])
        d, out = _po(oracle, p)
        assert d["n_bad"] == 0 and d["rounds"] == 4 and not out.any()
        tol = 1e-8 if c["stereo"] == 0.0 else 5e-6
        assert np.abs(d["R"] - g[f"case{k}_R"]).max() < tol, k
        assert np.abs(d["t"] - g[f"case{k}_t"]).max() < tol, k
        assert np.abs(d["R"] - p["R"]).max() < 1e-3          # and that optimum is the ground truth up to the noise


def test_poseopt_rejects_gross_outliers_and_recovers_the_pose(oracle):
    for seed, sr in ((9700, 0.0), (9701, 1.0), (9702, 0.4)):
        p = synth.poseopt_problem(seed, 300, 0.3, sr)
        d, out = _po(oracle, p)
        gross = ~p["inlier"]
        assert out[gross].mean() > 0.97                       # a random pixel is almost never within the chi2 gate
        assert out[~gross].mean() < 0.12                      # ~5 % (mono) of true inliers exceed a 95 % gate
        assert d["n_inliers"] == 300 - int(out.sum()) and d["n_bad"] == int(out.sum())
        assert np.abs(d["R"] - p["R"]).max() < 5e-3 and np.abs(d["t"] - p["t"]).max() < 5e-2
        assert np.abs(d["R"] @ d["R"].T - np.eye(3)).max() < 1e-12
        assert np.array_equal(d["Rf"], d["R"].astype(np.float32)) and np.array_equal(d["tf"], d["t"].astype(np.float32))


def test_poseopt_control_flow_edge_cases(oracle):
    # fewer than 3 correspondences: return 0, pose untouched (Optimizer.cpp:326-327)
    p = synth.poseopt_problem(9710, 2, 0.0)
    d, out = _po(oracle, p)
    assert d["n_inliers"] == 0 and d["rounds"] == 0 and d["iterations"] == 0
    assert np.abs(d["R"] - p["Rcw"].astype(np.float64)).max() < 1e-6
    # fewer than 10 edges: one round only (Optimizer.cpp:409-410)
    d, _ = _po(oracle, synth.poseopt_problem(9711, 9, 0.0))
    assert d["rounds"] == 1
    d, _ = _po(oracle, synth.poseopt_problem(9712, 10, 0.0))
    assert d["rounds"] == 4
    # every round restarts from the initial pose: at most 10 LM iterations per round, rejected trials on top
    d, _ = _po(oracle, synth.poseopt_problem(9713, 200, 0.5, pose_noise=(0.3, 1.0)))
    assert d["iterations"] <= 40 and d["trials"] >= d["iterations"]


def test_poseopt_frozen_vectors(oracle):
    g = np.load(os.path.join(GOLD, "poseopt_scipy.npz"))
    for k, (seed, n, outl, sr) in enumerate([(9600, 250, 0.2, 0.0), (9601, 250, 0.3, 1.0), (9602, 500, 0.4, 0.5), (9603, 9, 0.2, 0.0)]):
        d, out = _po(oracle, synth.poseopt_problem(seed, n, outl, sr))
        assert np.allclose(d["R"], g[f"frozen{k}_R"], rtol=0, atol=1e-9) and np.allclose(d["t"], g[f"frozen{k}_t"], rtol=0, atol=1e-9)
        assert (out != g[f"frozen{k}_outlier"]).sum() <= 1
        assert [d["n_inliers"], d["n_bad"], d["rounds"]] == g[f"frozen{k}_meta"][:3].tolist() or (out != g[f"frozen{k}_outlier"]).sum() == 1


# ------------------------------------------------ Optimizer::OptimizeSim3
def _so(oracle, p, **kw):
    return oracle.optimize_sim3(oracle.sim3opt_problem(p["x1c"], p["x2c"], p["obs1"], p["obs2"], p["inv_sigma2_1"], p["inv_sigma2_2"],
                                                       p["K"], p["K"], p["S12"], **kw))


def test_sim3opt_improves_the_estimate_and_removes_outliers(oracle):
    for seed in range(8600, 8606):
        p = synth.sim3opt_problem(seed, 150, 0.2)
        d, rem = _so(oracle, p)
        assert d["optimized"] == 1 and d["s"] == 1.0
        assert np.abs(d["R"] - p["R12"]).max() < 0.5 * np.abs(p["S12"][:9].reshape(3, 3) - p["R12"]).max()
        assert rem[~p["inlier"]].mean() > 0.9 and rem[p["inlier"]].mean() < 0.1
        assert d["n_inliers"] == 150 - int(rem.sum())
        assert abs(np.linalg.norm(d["q"]) - 1.0) < 1e-6      # never normalised by g2o, stays near 1 over 15 steps


def test_sim3opt_free_scale_and_early_return(oracle):
    p = synth.sim3opt_problem(8610, 150, 0.1, scale=1.5)
    d, _ = _so(oracle, p, fix_scale=False)
    assert abs(d["s"] - 1.5) < 0.05
    d, rem = _so(oracle, synth.sim3opt_problem(8611, 9, 0.0))       # fewer than 10 matches: return 0, estimate untouched
    assert d["optimized"] == 0 and d["n_inliers"] == 0
    d, rem = _so(oracle, synth.sim3opt_problem(8612, 0, 0.0))
    assert d["optimized"] == 0 and d["iterations"] == 0


def test_sim3opt_matches_scipy_on_outlier_free_pairs(oracle):
    """independent pin: with no outliers and noise well inside the Huber band the cost is plain weighted least squares
    over both edge families; scipy's optimum (computed here, seconds) against the oracle after its 5 + 5 iterations"""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    p = synth.sim3opt_problem(8620, 80, 0.0)
    K = [float(k) for k in p["K"]]
    X1, X2 = p["x1c"].astype(np.float64), p["x2c"].astype(np.float64)
    w1, w2 = np.sqrt(p["inv_sigma2_1"].astype(np.float64)), np.sqrt(p["inv_sigma2_2"].astype(np.float64))

    def proj(X):
        return np.stack([K[0] * X[:, 0] / X[:, 2] + K[2], K[1] * X[:, 1] / X[:, 2] + K[3]], axis=1)

    def resid(x):
        R = Rotation.from_rotvec(x[:3]).as_matrix()
        e12 = (p["obs1"] - proj(X2 @ R.T + x[3:])) * w1[:, None]
        e21 = (p["obs2"] - proj((X1 - x[3:]) @ R)) * w2[:, None]
        return np.concatenate([e12.ravel(), e21.ravel()])

    R0 = p["S12"][:9].reshape(3, 3).astype(np.float64)
    sol = least_squares(resid, np.concatenate([Rotation.from_matrix(R0).as_rotvec(), p["S12"][9:12].astype(np.float64)]), method="lm",
                        xtol=1e-14, ftol=1e-14, gtol=1e-14)
    d, rem = _so(oracle, p)
    if rem.sum() == 0:
        assert np.abs(d["R"] - Rotation.from_rotvec(sol.x[:3]).as_matrix()).max() < 1e-5
        assert np.abs(d["t"] - sol.x[3:]).max() < 1e-4
