"""Outcome-level agreement of 4-point EPnP RANSAC across null-space bases.

With four correspondences M is 8 x 12, so M^T M has an exact 4-dimensional null space and "the four smallest
eigenvectors" (PnPsolver.cpp:380-382) are an arbitrary orthonormal basis of it.  Hypothesis-level parity with an
Eigen-built binary is therefore impossible in principle; what CAN be measured is whether the *outcome* of the
whole RANSAC (accepted or not, final refined pose, final inlier set) depends on the basis.  Three bases:

  qr      the engine's default: Householder QR of M^T                         (oracle FLAG_EPNP_QR_NULLSPACE)
  jacobi  the oracle's 12 x 12 cyclic-Jacobi eigen-solve of M^T M             (the reference's structure)
  lapack  numpy.linalg.eigh of the same M^T M -- Householder tridiagonalisation + implicit QL/QR, the algorithm
          class of Eigen's SelfAdjointEigenSolver -- fed to the oracle through orc_epnp_pose_basis

Each runs the reference's sequential control flow (iterate + Refine, PnPsolver.cpp:102-238) on the same index
tables.  Thresholds below are stated and asserted; the measured figures are printed (pytest -s) and recorded in
DESIGN.md section 2.
"""
import numpy as np

from ransac_b200 import synth

PRM = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.2, th2=5.991)   # cfg1 / cfg4
N_PROBLEMS = 48
N = 500


def _ransac(oracle, pb, thr, table, minInl, pose_fn):
    """PnPsolver::iterate (first call) + Refine with a pluggable minimal solver"""
    best, bestmask = 0, None
    for h in range(len(table)):
        R, t = pose_fn(table[h])
        cnt, mask, _ = oracle.pnp_check_inliers(pb, thr, R, t)
        if cnt >= minInl:
            if cnt > best:
                best, bestmask = cnt, mask
            Rr, tr, _ = oracle.epnp_pose(pb, np.flatnonzero(bestmask))       # Refine: n > 4, well-posed eigen-solve
            cr, mr, _ = oracle.pnp_check_inliers(pb, thr, Rr, tr)
            if cr > minInl:
                return dict(ok=1, n_hyp=h + 1, n=cr, R=Rr, t=tr, mask=mr)
    return dict(ok=0, n_hyp=len(table), n=0, R=None, t=None, mask=None)


def _lapack_basis(oracle, pb, idx):
    w, v = np.linalg.eigh(oracle.epnp_mtm(pb, idx))
    return v[:, :4]


def test_outcome_agreement_across_null_space_bases(oracle):
    b = synth.pnp_batch(4, N_PROBLEMS, N, 0.5)
    prm = oracle.params(**PRM)
    minInl, H = oracle.ransac_setup_pnp(N, prm)
    out = {k: [] for k in ("qr", "jacobi", "lapack")}
    for c in range(N_PROBLEMS):
        pb = oracle.pnp_problem(b["p3d"][c], b["p2d"][c], b["sigma2"][c], b["K"])
        thr = (b["sigma2"][c] * np.float32(PRM["th2"])).astype(np.float32)
        table = oracle.index_table(int(b["seeds"][c]), N, 4, H)
        fns = {
            "qr": lambda idx: oracle.epnp_pose(pb, idx, oracle.FLAG_EPNP_QR_NULLSPACE)[:2],
            "jacobi": lambda idx: oracle.epnp_pose(pb, idx, 0)[:2],
            "lapack": lambda idx: oracle.epnp_pose_basis(pb, idx, _lapack_basis(oracle, pb, idx))[:2],
        }
        for k, fn in fns.items():
            o = _ransac(oracle, pb, thr, table, minInl, fn)
            # cross-check the harness against the oracle's own RANSAC for the two in-oracle modes
            if k != "lapack":
                ref = oracle.pnp_ransac(pb, prm, table, oracle.FLAG_EPNP_QR_NULLSPACE if k == "qr" else 0)
                assert (o["ok"], o["n_hyp"], o["n"]) == (ref["ok"], ref["n_hyp"], ref["n_inliers"])
            o["gt_R"], o["gt_t"] = b["R"][c], b["t"][c]
            out[k].append(o)

    def pair(a, b_):
        acc = [(x["ok"], y["ok"]) for x, y in zip(out[a], out[b_])]
        same_acc = np.mean([x == y for x, y in acc])
        dR, dt, jac, dn = [], [], [], []
        for x, y in zip(out[a], out[b_]):
            if x["ok"] and y["ok"]:
                dR.append(np.abs(x["R"] - y["R"]).max())
                dt.append(np.abs(x["t"] - y["t"]).max() / max(1.0, np.abs(y["t"]).max()))
                jac.append((x["mask"] & y["mask"]).sum() / max(1, (x["mask"] | y["mask"]).sum()))
                dn.append(abs(x["n"] - y["n"]))
        return dict(same_accept=same_acc, dR_max=max(dR), dR_med=float(np.median(dR)), dt_max=max(dt), dt_med=float(np.median(dt)),
                    jaccard_min=min(jac), jaccard_med=float(np.median(jac)), dn_max=max(dn))

    rep = {f"{a}~{b_}": pair(a, b_) for a, b_ in (("qr", "jacobi"), ("qr", "lapack"), ("jacobi", "lapack"))}
    hyps = {k: float(np.mean([o["n_hyp"] for o in v])) for k, v in out.items()}
    inl = {k: float(np.mean([o["n"] for o in v])) for k, v in out.items()}
    gt = {k: float(np.max([np.abs(o["R"] - o["gt_R"]).max() for o in v if o["ok"]])) for k, v in out.items()}
    print("\nbasis agreement on %d cfg4 problems: mean hypotheses to success %s, mean final inliers %s, max |R - R_gt| %s" % (N_PROBLEMS, hyps, inl, gt))
    for k, v in rep.items():
        print(" ", k, {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()})

    # Measured (48 cfg4 problems, recorded in DESIGN.md section 2): every basis accepts the same candidates, but the
    # sequential scan stops at a different hypothesis (mean 38-45 of 300), so the refined poses differ at the level
    # of the pixel noise carried by the inlier sets they were refined on -- |dR| <= 7e-3, |dt| <= 7e-2 relative,
    # mask Jaccard median 0.89-0.93 -- and jacobi~lapack (two eigen-solvers of the reference's own structure) differ
    # by as much as qr~jacobi.  The 1e-4 bar of north_star is met only on IDENTICAL hypotheses (GPU vs oracle).
    for k, v in rep.items():
        assert v["same_accept"] == 1.0, (k, v)            # every basis accepts exactly the same candidates
        assert v["dR_max"] < 1.5e-2 and v["dt_max"] < 0.15, (k, v)
        assert v["jaccard_med"] > 0.8 and v["jaccard_min"] > 0.5, (k, v)
    # the QR basis is inside the ambiguity the reference's own eigen-solve has: not further from the Jacobi
    # eigen-solve than LAPACK's eigen-solve is
    ref = rep["jacobi~lapack"]
    for k in ("qr~jacobi", "qr~lapack"):
        assert rep[k]["dR_med"] <= 2.0 * ref["dR_med"] and rep[k]["dt_med"] <= 2.0 * ref["dt_med"], (k, rep[k], ref)
        assert rep[k]["jaccard_med"] >= ref["jaccard_med"] - 0.1, (k, rep[k], ref)
    for k in out:
        assert gt[k] < 1e-2                               # and every basis recovers the ground-truth pose
        assert all(o["ok"] for o in out[k])
