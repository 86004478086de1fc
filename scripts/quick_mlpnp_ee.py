"""Developer probe: cfg2 MLPnP batches (64 frames x 1000 matches, covariances), six in flight, over early-exit stage shapes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
C2, N2 = 64, 1000
b2 = synth.pnp_batch(2, C2, N2, 0.5)
cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
Kf = np.array([b2["K"]], np.float32)
off2 = (np.arange(C2 + 1) * N2).astype(np.int32)
prm2 = capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991)
NF = int(os.environ.get("NF", "6"))
pool = [capi.Engine(0) for _ in range(NF)]
for q in pool:
    q.mlpnp_upload(off2, b2["p3d"], b2["p2d"], b2["sigma2"], Kf, prm2, cov=cov, seeds=b2["seeds"])
    q.mlpnp_run()
ref, _ = pool[0].mlpnp_download()
print("n_hyp of the reference semantics: mean %.1f median %.1f max %d" % (ref["n_hyp"].mean(), np.median(ref["n_hyp"]), ref["n_hyp"].max()))
shapes = [None] + [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for stg in shapes:
    flags = 0 if stg is None else capi.FLAG_EARLY_EXIT
    for q in pool:
        q.set_stages([] if stg is None else list(stg))
        for _ in range(3):
            q.mlpnp_run(flags)
    for q in pool:
        q.sync()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        for q in pool:
            q.mlpnp_run(flags)
    for q in pool:
        q.sync()
    dt = (time.perf_counter() - t0) / (reps * len(pool))
    r, _ = pool[0].mlpnp_download()
    same = all((r[f] == ref[f]).all() for f in ("ok", "n_inliers", "best_hyp", "n_refines", "n_hyp"))
    st = pool[0].mlpnp_phase_stats()
    e = pool[0]
    e.timer_begin()
    for _ in range(5):
        e.mlpnp_run(flags)
    one = e.timer_end() / 5
    print("stages %-22s in flight %.3f ms/batch  %.0f frames/s   one batch alone %.3f ms   done %.1f %%  stage1 %d cleanup %d  same=%s" % (
        stg, dt * 1e3, C2 / dt, one, 100.0 * st[3] / (C2 * 300), st[1], st[2], same))
