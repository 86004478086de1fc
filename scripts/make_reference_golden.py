#!/usr/bin/env python
"""Generates tests/golden/reference_build.npz: outputs of THE REFERENCE'S OWN PnPsolver.cpp / Sim3Solver.cpp, compiled
unmodified by `make -C oracle ref` (stand-in Eigen / OpenCV headers under oracle/shim/, see oracle/shim/Eigen/Dense),
on seeded synthetic inputs -- the inputs are stored with them, so the file is self-contained.

Runs only where /root/reference exists (this container).  The oracle is checked against the file by
tests/test_cpu_reference_build.py::test_oracle_equals_reference_golden (no reference needed) and the CUDA engine by
tests/test_gpu_reference_golden.py on the GPU box.

    python scripts/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
import ref_api  # noqa: E402
from ransac_b200 import synth  # noqa: E402

CFG4 = dict(prob=0.99, min_inliers=10, max_its=300, min_set=4, eps=0.2, th2=5.991)
SIM3 = dict(prob=0.99, min_inliers=20, max_its=300)


def main():
    assert ref_api.available(), "oracle/_ref/libref_solvers.so could not be built (needs /root/reference)"
    ls2 = synth.level_sigma2()
    out = {"level_sigma2": ls2}

    # ---- PnPsolver: 16 cfg4 problems (500 matches, 50 % outliers), Relocalization's first iterate() call
    Cn, n = 16, 500
    seeds = np.arange(4000, 4000 + Cn, dtype=np.uint32)
    ps = [synth.pnp_problem(int(s), n, 0.5) for s in seeds]
    K = np.array([np.float32(k) for k in ps[0]["K"]], np.float32)
    rec = dict(ok=[], no_more=[], n_inliers=[], iterations=[], best_inliers=[], n_refines=[], T=[], inliers=[], min_inliers=[], max_its=[])
    for s, p in zip(seeds, ps):
        sol = ref_api.PnP(p["p2d"], p["octave"], ls2, p["p3d"], K)
        sol.set_params(**CFG4)
        rp = sol.params()
        ref_api.seed(int(s))
        ref_api.eig_record(False)
        r = sol.iterate(rp["max_its"])
        st = sol.state(n)
        rec["ok"].append(r["ok"]); rec["no_more"].append(r["no_more"]); rec["n_inliers"].append(r["n_inliers"])
        rec["iterations"].append(st["iterations"]); rec["best_inliers"].append(st["best_inliers"])
        rec["n_refines"].append(ref_api.eig_calls() - st["iterations"])
        rec["T"].append(r["T"]); rec["inliers"].append(r["inliers"])
        rec["min_inliers"].append(rp["min_inliers"]); rec["max_its"].append(rp["max_its"])
    assert all(rec["ok"]) and all(k == 1 for k in rec["n_refines"]), "golden PnP runs must succeed at their first Refine (no stale rows, SURVEY Q1)"
    out.update(pnp_seeds=seeds, pnp_K=K, pnp_params=np.array([CFG4[k] for k in ("prob", "min_inliers", "max_its", "min_set", "eps", "th2")], np.float64),
               pnp_p3d=np.stack([p["p3d"] for p in ps]), pnp_p2d=np.stack([p["p2d"] for p in ps]),
               pnp_octave=np.stack([p["octave"] for p in ps]).astype(np.int32))
    for k, v in rec.items():
        out["pnp_" + k] = np.array(v)

    # per-call EPnP poses (PnPsolver::compute_pose) on subsets of problem 0: m = 4 .. 250
    rng = np.random.default_rng(1)
    inl = np.flatnonzero(ps[0]["inlier"])
    sets, Rs, ts = [], [], []
    for m in (4, 4, 4, 4, 5, 6, 8, 20, 100, 250):
        idx = (rng.permutation(n)[:4] if m == 4 else rng.permutation(inl)[:m]).astype(np.int32)
        sol = ref_api.PnP(ps[0]["p2d"], ps[0]["octave"], ls2, ps[0]["p3d"], K)
        sol.set_params(**CFG4)
        R, t, _ = sol.compute_pose(idx)
        sets.append(np.pad(idx, (0, 250 - m), constant_values=-1)); Rs.append(R); ts.append(t)
    out.update(pose_sets=np.stack(sets), pose_R=np.stack(Rs), pose_t=np.stack(ts))

    # ---- Sim3Solver: 12 loop candidates (200 matches, 40 - 80 % outliers), iterate(5) until done (LoopClosing.cpp:275)
    Cs, ns = 12, 200
    sseeds = np.arange(5200, 5200 + Cs, dtype=np.uint32)
    srec = dict(ok=[], n_inliers=[], iterations=[], best_inliers=[], R=[], t=[], inliers=[], max_its=[])
    sp = []
    for c, s in enumerate(sseeds):
        p = synth.sim3_problem(int(s), ns, (0.4, 0.6, 0.8)[c % 3])
        o1 = np.searchsorted(ls2, p["sigma2_1"]).astype(np.int32)
        o2 = np.searchsorted(ls2, p["sigma2_2"]).astype(np.int32)
        p["o1"], p["o2"] = o1, o2
        sp.append(p)
        sol = ref_api.Sim3(p["x1c"], p["x2c"], o1, o2, ls2, p["K"], p["K"])
        sol.set_params(**SIM3)
        ref_api.seed(int(s))
        while True:
            r = sol.iterate(5)
            if r["ok"] or r["no_more"]:
                break
        st = sol.state(ns)
        srec["ok"].append(r["ok"]); srec["n_inliers"].append(r["n_inliers"]); srec["iterations"].append(st["iterations"])
        srec["best_inliers"].append(st["best_inliers"]); srec["R"].append(st["R"]); srec["t"].append(st["t"])
        srec["inliers"].append(r["inliers"]); srec["max_its"].append(sol.params()["max_its"])
    out.update(sim3_seeds=sseeds, sim3_K=np.array(sp[0]["K"], np.float32), sim3_params=np.array([SIM3["prob"], SIM3["min_inliers"], SIM3["max_its"]], np.float64),
               sim3_x1c=np.stack([p["x1c"] for p in sp]), sim3_x2c=np.stack([p["x2c"] for p in sp]),
               sim3_oct1=np.stack([p["o1"] for p in sp]), sim3_oct2=np.stack([p["o2"] for p in sp]))
    for k, v in srec.items():
        out["sim3_" + k] = np.array(v)

    # ---- MLPnPsolver: 6 cfg2-shaped frames (1000 matches, 50 % outliers, no covariances: the path iterate() takes)
    MLP = dict(prob=0.99, min_inliers=10, max_its=300, min_set=6, eps=0.2, th2=5.991)
    Cm, nm = 6, 1000
    mseeds = np.arange(2000, 2000 + Cm, dtype=np.uint32)
    mps = [synth.pnp_problem(int(s), nm, 0.5) for s in mseeds]
    mrec = dict(ok=[], no_more=[], n_inliers=[], iterations=[], best_inliers=[], T=[], inliers=[], max_its=[])
    for s, p in zip(mseeds, mps):
        sol = ref_api.MLPnP(p["p2d"], p["octave"], ls2, p["p3d"], K)
        sol.set_params(**MLP)
        H = sol.params()["max_its"]
        ref_api.seed(int(s))
        r = sol.iterate(H)
        st = sol.state()
        mrec["ok"].append(r["ok"]); mrec["no_more"].append(r["no_more"]); mrec["n_inliers"].append(r["n_inliers"])
        mrec["iterations"].append(st["iterations"]); mrec["best_inliers"].append(st["best_inliers"]); mrec["T"].append(r["T"])
        mrec["inliers"].append(r["inliers"]); mrec["max_its"].append(H)
    out.update(mlpnp_seeds=mseeds, mlpnp_params=np.array([MLP[k] for k in ("prob", "min_inliers", "max_its", "min_set", "eps", "th2")], np.float64),
               mlpnp_p3d=np.stack([p["p3d"] for p in mps]), mlpnp_p2d=np.stack([p["p2d"] for p in mps]),
               mlpnp_octave=np.stack([p["octave"] for p in mps]).astype(np.int32))
    for k, v in mrec.items():
        out["mlpnp_" + k] = np.array(v)

    path = os.path.join(ROOT, "tests", "golden", "reference_build.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "PnP iterations", out["pnp_iterations"].tolist(), "inliers", out["pnp_n_inliers"].tolist(),
          "| Sim3 ok", out["sim3_ok"].astype(int).tolist(), "iterations", out["sim3_iterations"].tolist(),
          "| MLPnP iterations", out["mlpnp_iterations"].tolist(), "inliers", out["mlpnp_n_inliers"].tolist())




# ---------------------------------------------------------------------------------------------------------------------------
# second file: retrieval and matching.  The inputs are large (descriptors, grids, bag-of-words vectors)
# and come from seeded generators (ransac_b200.synth), so only the generator arguments, a SHA-1 of the generated arrays and the
# compiled reference's OUTPUTS are stored: tests/golden/reference_build_matching.npz
def matching_cases():
    """the cases, as (name, builder) -- shared with tests/test_gpu_reference_golden.py"""
    def bow0():
        F = synth.bow_frame(11, 600, 40)
        kfs = [synth.bow_keyframe(1000 + i, F, 500 + 37 * (i % 5), shared=0.1 + 0.05 * (i % 7), rot=10.0 * i) for i in range(6)]
        return F, kfs

    def bow1():
        cur = synth.bow_frame(21, 500, 40)
        cur["valid"] = (np.random.default_rng(5).random(500) < 0.75).astype(np.uint8)
        kfs = [synth.bow_keyframe(2000 + i, cur, 450, shared=0.3, rot=5.0 * i) for i in range(4)]
        return cur, kfs

    def pairs():
        return [synth.kf_view_pair(10 + i, n_points=300 + 250 * i, n_extra=100 + 60 * i, prematched=0.1 * i) for i in range(3)]

    def proj():
        return [synth.proj_search_case(s, n_points=450, n_extra=150) for s in (1, 2, 3)]

    def kfdb():
        db = synth.kf_database(1, K=300, n_places=30)
        return db, [synth.kf_query(100 + q, db, place=(q * 7) % 30) for q in range(12)]

    return dict(bow0=bow0, bow1=bow1, pairs=pairs, proj=proj, kfdb=kfdb)


def digest(*arrays):
    import hashlib

    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def matching_digests(cases):
    F, kfs = cases["bow0"]()
    cur, kfs1 = cases["bow1"]()
    prs = cases["pairs"]()
    pj = cases["proj"]()
    db, qs = cases["kfdb"]()
    return dict(
        bow0=digest(F["desc"], F["angle"], F["node_feat"], *[k["desc"] for k in kfs], *[k["node_feat"] for k in kfs]),
        bow1=digest(cur["desc"], cur["valid"], *[k["desc"] for k in kfs1]),
        pairs=digest(*[p["kf1"]["desc"] for p in prs], *[p["kf2"]["mp_xyz"] for p in prs], *[p["R12"] for p in prs], *[p["matched12_in"] for p in prs]),
        proj=digest(*[c["frame"]["desc"] for c in pj], *[c["kf"]["mp_xyz"] for c in pj], *[c["Rcw"] for c in pj], *[c["occupied"] for c in pj]),
        kfdb=digest(db["bow_word"], db["bow_val"], db["covis"], *[q[0] for q in qs], *[q[1] for q in qs]))


def main_matching():
    import oracle_api as O

    cases = matching_cases()
    out = {("sha1_" + k): np.array(v) for k, v in matching_digests(cases).items()}
    F, kfs = cases["bow0"]()
    tF = O.bow_features(F)
    res = [ref_api.search_by_bow(O.bow_features(k), tF, 0.75, True, 0) for k in kfs]
    out["bow0_match"] = np.stack([r[0] for r in res]); out["bow0_n"] = np.array([r[1] for r in res])
    cur, kfs1 = cases["bow1"]()
    qC = O.bow_features(cur)
    res = [ref_api.search_by_bow(qC, O.bow_features(k), 0.75, True, 1) for k in kfs1]
    out["bow1_match"] = np.stack([r[0] for r in res]); out["bow1_n"] = np.array([r[1] for r in res])
    for i, p in enumerate(cases["pairs"]()):
        k1, k2 = O.kf_view(p["kf1"]), O.kf_view(p["kf2"])
        m, n = ref_api.search_by_sim3(k1, k2, p["K"], p["R12"], p["t12"], 7.5, p["matched12_in"])
        m = m.copy(); m[np.asarray(p["matched12_in"]) >= 0] = -1          # the new matches only (the reference leaves the old ones in place)
        out["sim3s_match_%d" % i] = m; out["sim3s_n_%d" % i] = np.array(n)
    for i, c in enumerate(cases["proj"]()):
        m, n = ref_api.search_by_projection(O.kf_view(c["frame"]), O.kf_view(c["kf"]), c["K"], c["Rcw"], c["tcw"], 10.0, 100, True, c["occupied"], c["already_found"])
        out["proj_match_%d" % i] = m; out["proj_n_%d" % i] = np.array(n)
    db, qs = cases["kfdb"]()
    rdb = ref_api.KfDb(db)
    for q, (qw, qv) in enumerate(qs):
        out["kfdb_cand_%d" % q] = np.array(rdb.reloc(qw, qv, frame_id=7000 + q), np.int32)
    out["kfdb_state"] = rdb.reloc_scores()
    path = os.path.join(ROOT, "tests", "golden", "reference_build_matching.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; SearchByBoW matches", out["bow0_n"].tolist(), out["bow1_n"].tolist(),
          "| SearchBySim3", [int(out["sim3s_n_%d" % i]) for i in range(3)], "| SearchByProjection", [int(out["proj_n_%d" % i]) for i in range(3)],
          "| candidates per query", [len(out["kfdb_cand_%d" % q]) for q in range(12)])


if __name__ == "__main__":
    main()
    main_matching()
