"""Generates the golden fixtures under tests/golden/ (run here, commit the outputs).

Sources of truth, in order of independence:
  * rng_known_answers.json  -- libc rand() stream and minimal-set tables (the numbers
                               are also hard-coded in tests/test_cpu_oracle.py)
  * cv2_epnp.npz            -- cv2.solvePnP(SOLVEPNP_EPNP) poses for noise-free n = 6 / 50 sets: an
                               independent EPnP implementation (valid for n >= 6 only)
  * scoring_numpy.npz       -- numpy float32/float64 emulation of the three mixed-precision
                               CheckInliers expressions (PnPsolver.cpp:250-258, MLPnPsolver.cpp:231-245,
                               Sim3Solver.cpp:269-293) on seeded inputs
  * oracle_frozen.npz       -- outputs of the oracle itself on the BASELINE configs (regression pin:
                               RANSAC outcomes, per-hypothesis counts, poses)
  * poseopt_scipy.npz       -- Optimizer::PoseOptimization: scipy.optimize.least_squares optimum of the same
                               weighted reprojection cost on outlier-free frames (every edge stays an inlier, so
                               the reference's last, kernel-free round minimises exactly that cost), and the
                               oracle's own outputs on noisy frames with outliers (regression pin)
The reference cannot be run to generate vectors: it needs Eigen/OpenCV C++ headers.
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
import oracle_api as O  # noqa: E402
from ransac_b200 import synth  # noqa: E402


def rng():
    libc = ctypes.CDLL("libc.so.6")
    out = {}
    for seed in (0, 1, 7, 4000):
        libc.srand(ctypes.c_uint(seed))
        out[str(seed)] = [libc.rand() for _ in range(16)]
    tabs = {f"{s}_{n}_{k}": O.index_table(s, n, k, 8).tolist() for (s, n, k) in ((1, 500, 4), (1, 200, 3), (1, 1000, 6), (4000, 37, 4))}
    json.dump({"rand": out, "tables": tabs}, open(os.path.join(HERE, "rng_known_answers.json"), "w"), indent=1)


def cv2_epnp():
    import cv2
    recs = {}
    for n in (6, 50):
        for seed in (7, 8, 9):
            p = synth.pnp_problem(seed, n, 0.0, noise=False)
            K = np.array([[p["K"][0], 0, p["K"][2]], [0, p["K"][1], p["K"][3]], [0, 0, 1]])
            ok, rv, tv = cv2.solvePnP(p["p3d"].astype(np.float64), p["p2d"].astype(np.float64), K, None, flags=cv2.SOLVEPNP_EPNP)
            R, _ = cv2.Rodrigues(rv)
            recs[f"R_{n}_{seed}"] = R
            recs[f"t_{n}_{seed}"] = tv.ravel()
    np.savez(os.path.join(HERE, "cv2_epnp.npz"), **recs)


def scoring_numpy():
    f32, f64 = np.float32, np.float64
    p = synth.scoring_stress(5100, 16, 257)
    X = p["p3d"]
    u, v = p["p2d"][:, 0], p["p2d"][:, 1]
    fx, fy, cx, cy = [f64(k) for k in p["K"]]
    fxf, fyf, cxf, cyf = [f32(k) for k in p["K"]]
    pnp_e, ml_e = [], []
    for h in range(16):
        R = p["poses"][h, :9].reshape(3, 3)
        t = p["poses"][h, 9:]
        # PnPsolver.cpp:250-258: f32 transform, f32 reciprocal, f64 projection narrowed to f32
        xc = f32(f32(f32(R[0, 0] * X[:, 0]) + f32(R[0, 1] * X[:, 1])) + f32(R[0, 2] * X[:, 2])) + t[0]
        yc = f32(f32(f32(R[1, 0] * X[:, 0]) + f32(R[1, 1] * X[:, 1])) + f32(R[1, 2] * X[:, 2])) + t[1]
        zc = f32(f32(f32(R[2, 0] * X[:, 0]) + f32(R[2, 1] * X[:, 1])) + f32(R[2, 2] * X[:, 2])) + t[2]
        iz = f32(1) / zc
        ue = (cx + fx * xc.astype(f64) * iz.astype(f64)).astype(f32)
        ve = (cy + fy * yc.astype(f64) * iz.astype(f64)).astype(f32)
        du, dv = ue - u, ve - v
        pnp_e.append(f32(du * du) + f32(dv * dv))
        # MLPnPsolver.cpp:231-245: f64 transform narrowed to f32, two f32 divisions
        Rd, td = R.astype(f64), t.astype(f64)
        Xd = X.astype(f64)
        xm = (Rd[0, 0] * Xd[:, 0] + Rd[0, 1] * Xd[:, 1] + Rd[0, 2] * Xd[:, 2] + td[0]).astype(f32)
        ym = (Rd[1, 0] * Xd[:, 0] + Rd[1, 1] * Xd[:, 1] + Rd[1, 2] * Xd[:, 2] + td[1]).astype(f32)
        zm = (Rd[2, 0] * Xd[:, 0] + Rd[2, 1] * Xd[:, 1] + Rd[2, 2] * Xd[:, 2] + td[2]).astype(f32)
        um = f32(f32(fxf * xm) / zm) + cxf
        vm = f32(f32(fyf * ym) / zm) + cyf
        dx, dy = u - um, v - vm
        ml_e.append(f32(dx * dx) + f32(dy * dy))
    np.savez(os.path.join(HERE, "scoring_numpy.npz"), seed=5100, H=16, n=257, pnp_err2=np.array(pnp_e), mlpnp_err2=np.array(ml_e))


def oracle_frozen():
    out = {}
    # cfg1
    p = synth.pnp_problem(1000, 500, 0.5)
    pb = O.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
    prm = O.params(0.99, 10, 300, 4, 0.2, 5.991)
    r = O.pnp_ransac(pb, prm, O.index_table(1000, 500, 4, 300), O.FLAG_EXHAUSTIVE, per_hyp=True)
    out["cfg1_counts"] = r["hyp_counts"]; out["cfg1_pose"] = r["hyp_pose"]; out["cfg1_T"] = r["T"]
    out["cfg1_meta"] = np.array([r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]])
    out["cfg1_mask"] = r["mask"]
    # cfg1 in the engine's default mode (4-point null space by Householder QR)
    r = O.pnp_ransac(pb, prm, O.index_table(1000, 500, 4, 300), O.FLAG_EXHAUSTIVE | O.FLAG_EPNP_QR_NULLSPACE, per_hyp=True)
    out["cfg1q_counts"] = r["hyp_counts"]; out["cfg1q_pose"] = r["hyp_pose"]; out["cfg1q_T"] = r["T"]
    out["cfg1q_meta"] = np.array([r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]])
    out["cfg1q_mask"] = r["mask"]
    # cfg3 (fixed scale and free scale)
    for tag, sc in (("cfg3", 1.0), ("cfg3s", 1.6)):
        q = synth.sim3_problem(3000, 200, 0.4, sc)
        sb = O.sim3_problem(q["x1c"], q["x2c"], q["sigma2_1"], q["sigma2_2"], q["K"], q["K"], fix_scale=(sc == 1.0))
        r = O.sim3_ransac(sb, 0.99, 20, 300, O.index_table(3000, 200, 3, 300), O.FLAG_EXHAUSTIVE, per_hyp=True)
        out[tag + "_counts"] = r["hyp_counts"]; out[tag + "_pose"] = r["hyp_pose"]; out[tag + "_T"] = r["T"]
        out[tag + "_meta"] = np.array([r["ok"], r["n_inliers"], r["best_hyp"], r["best_count"]]); out[tag + "_scale"] = np.float32(r["scale"])
        out[tag + "_mask"] = r["mask"]
    # cfg2 (one frame, with covariances)
    p = synth.pnp_problem(2000, 1000, 0.5)
    Kf = tuple(np.float32(k) for k in p["K"])
    mb = O.mlpnp_problem(p["p3d"], p["p2d"], p["sigma2"], Kf, synth.bearing_covariances(p))
    prm = O.params(0.99, 10, 300, 6, 0.2, 5.991)
    r = O.mlpnp_ransac(mb, prm, O.index_table(2000, 1000, 6, 300), O.FLAG_EXHAUSTIVE, per_hyp=True)
    out["cfg2_counts"] = r["hyp_counts"]; out["cfg2_pose"] = r["hyp_pose"]; out["cfg2_T"] = r["T"]
    out["cfg2_meta"] = np.array([r["ok"], r["n_inliers"], r["best_hyp"], r["refined"], r["n_refines"]])
    out["cfg2_mask"] = r["mask"]
    np.savez_compressed(os.path.join(HERE, "oracle_frozen.npz"), **out)


def _poseopt_cases():
    return [dict(seed=9500 + i, n=n, outl=0.0, stereo=sr, noise_scale=0.3) for i, (n, sr) in
            enumerate([(60, 0.0), (200, 0.0), (200, 1.0), (400, 0.5)])]


def poseopt_scipy():
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    out = {}
    for k, c in enumerate(_poseopt_cases()):
        p = synth.poseopt_problem(c["seed"], c["n"], c["outl"], c["stereo"], noise_scale=c["noise_scale"])
        X = p["p3d"].astype(np.float64)
        obs = p["obs"].astype(np.float64)
        w = np.sqrt(p["inv_sigma2"].astype(np.float64))
        fx, fy, cx, cy, bf = [float(v) for v in p["K"]]
        st = obs[:, 2] >= 0

        def resid(x):
            R = Rotation.from_rotvec(x[:3]).as_matrix()
            Xc = X @ R.T + x[3:]
            u = fx * Xc[:, 0] / Xc[:, 2] + cx
            v = fy * Xc[:, 1] / Xc[:, 2] + cy
            r = [w * (obs[:, 0] - u), w * (obs[:, 1] - v), np.where(st, w * (obs[:, 2] - (u - bf / Xc[:, 2])), 0.0)]
            return np.concatenate(r)

        x0 = np.concatenate([Rotation.from_matrix(p["Rcw"].astype(np.float64)).as_rotvec(), p["tcw"].astype(np.float64)])
        sol = least_squares(resid, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
        out[f"case{k}_R"] = Rotation.from_rotvec(sol.x[:3]).as_matrix()
        out[f"case{k}_t"] = sol.x[3:]
        out[f"case{k}_cost"] = 2.0 * sol.cost
    # regression pin: the oracle on noisy frames with outliers, monocular / stereo / mixed
    for k, (seed, n, outl, sr) in enumerate([(9600, 250, 0.2, 0.0), (9601, 250, 0.3, 1.0), (9602, 500, 0.4, 0.5), (9603, 9, 0.2, 0.0)]):
        p = synth.poseopt_problem(seed, n, outl, sr)
        d, o = O.pose_optimization(O.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        out[f"frozen{k}_R"] = d["R"]; out[f"frozen{k}_t"] = d["t"]; out[f"frozen{k}_outlier"] = o
        out[f"frozen{k}_meta"] = np.array([d["n_inliers"], d["n_bad"], d["rounds"], d["iterations"], d["trials"]])
    np.savez_compressed(os.path.join(HERE, "poseopt_scipy.npz"), **out)


if __name__ == "__main__":
    O.build()
    if "poseopt" in sys.argv:
        poseopt_scipy()
    elif "frozen" in sys.argv:      # after a deliberate change of the arithmetic contract: only the regression pin
        oracle_frozen()
    else:
        rng(); cv2_epnp(); scoring_numpy(); oracle_frozen(); poseopt_scipy()
    print(sorted(os.listdir(HERE)))
