"""Short program for ncu: PnP sweeps (C candidates, early exit in phases unless RSAC_PROF_EXHAUSTIVE=1) and a few cfg5
scoring launches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = capi.Engine(0)
n = 500
b = synth.pnp_batch(4, C, n, 0.5)
offsets = (np.arange(C + 1) * n).astype(np.int32)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
FLAGS = 0 if os.environ.get("RSAC_PROF_EXHAUSTIVE") == "1" else capi.FLAG_EARLY_EXIT
for _ in range(reps):
    eng.pnp_run(FLAGS)
res, _ = eng.pnp_download()
p = synth.scoring_stress(5000, 4096, 10000)
max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
eng.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
for _ in range(reps + 1):
    eng.score_pnp_run(True)
eng.sync()
print("ok", int(res["ok"].sum()))

# ---- round 2: the other kernels north_star asks a capture for (RSAC_PROF_ALL=1)
if os.environ.get("RSAC_PROF_ALL") == "1":
    C2, N2 = 64, 1000
    b2 = synth.pnp_batch(2, C2, N2, 0.5)
    cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
    eng.mlpnp_upload((np.arange(C2 + 1) * N2).astype(np.int32), b2["p3d"], b2["p2d"], b2["sigma2"], np.array([b2["K"]], np.float32),
                     capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991), cov=cov, seeds=b2["seeds"])
    for _ in range(reps):
        eng.mlpnp_run()
    ps = [synth.sim3_problem(3000 + i, 200, 0.4, 1.0) for i in range(64)]
    cat = lambda k: np.concatenate([q[k] for q in ps])
    K3 = np.array([ps[0]["K"]], np.float32)
    for C3 in (1, 64):
        n3 = 200 * C3
        eng.sim3_upload((np.arange(C3 + 1) * 200).astype(np.int32), cat("x1c")[:n3], cat("x2c")[:n3], cat("sigma2_1")[:n3], cat("sigma2_2")[:n3],
                        K3, K3, capi.Sim3Params(0.99, 20, 300, 1), seeds=np.arange(C3, dtype=np.uint32) + 3000)
        for _ in range(reps):
            eng.sim3_run()
    Fb = synth.bow_frame(11, 1500, 100)
    kfs = [synth.bow_keyframe(1000 + i, Fb, 1200, shared=0.25, rot=7.0 * i) for i in range(8)]
    eng.bow_upload([Fb] + kfs, [1 + (i % 8) for i in range(1024)], [0] * 1024, 0.75, True, 0)
    for _ in range(reps):
        eng.bow_run()
    eng.sync()
    print("all ok")

# ---- round 2, second session: retrieval and guided matching (RSAC_PROF_NEW=1)
if os.environ.get("RSAC_PROF_NEW") == "1":
    dbk = synth.kf_database(21, K=2048, n_places=128)
    eng.kfdb_upload(dbk)
    eng.kfdb_query_upload([synth.kf_query(500 + i, dbk, (37 * i) % 128) for i in range(16)], mode=0)
    for _ in range(reps):
        eng.kfdb_run()
    prs = [synth.kf_view_pair(700 + i, n_points=1200, n_extra=400, prematched=0.3) for i in range(4)]
    vws = [v for p_ in prs for v in (p_["kf1"], p_["kf2"])]
    k1i, k2i = [2 * (i % 4) for i in range(32)], [2 * (i % 4) + 1 for i in range(32)]
    eng.sim3_search_upload(vws, k1i, k2i, [prs[i % 4]["K"] for i in range(32)], [prs[i % 4]["R12"] for i in range(32)],
                           [prs[i % 4]["t12"] for i in range(32)], 7.5, [prs[i % 4]["matched12_in"] for i in range(32)])
    for _ in range(reps):
        eng.sim3_search_run()
    cs = [synth.proj_search_case(800 + i, n_points=1200, n_extra=400, clones=0.1) for i in range(4)]
    vwp = [v for c_ in cs for v in (c_["frame"], c_["kf"])]
    eng.proj_search_upload(vwp, k1i, k2i, [cs[i % 4]["K"] for i in range(32)], [cs[i % 4]["Rcw"] for i in range(32)],
                           [cs[i % 4]["tcw"] for i in range(32)], 10.0, 100, True, [cs[i % 4]["occupied"] for i in range(32)],
                           [cs[i % 4]["already_found"] for i in range(32)])
    for _ in range(reps):
        eng.proj_search_run()
    # staged MLPnP
    C2, N2 = 64, 1000
    b2 = synth.pnp_batch(2, C2, N2, 0.5)
    cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
    eng.set_stages([128])
    eng.mlpnp_upload((np.arange(C2 + 1) * N2).astype(np.int32), b2["p3d"], b2["p2d"], b2["sigma2"], np.array([b2["K"]], np.float32),
                     capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991), cov=cov, seeds=b2["seeds"])
    eng.set_graphs(False)
    for _ in range(reps):
        eng.mlpnp_run(capi.FLAG_EARLY_EXIT)
    eng.set_stages([])
    eng.sync()
    print("new ok")
