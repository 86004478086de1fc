"""Developer probe: cfg5 scoring launches with k independent jobs in flight; host enqueue time vs device time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
H, N = 4096, 10000
p = synth.scoring_stress(5000, H, N)
max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
fp32, _ = capi.Engine(0).measure_peaks()
for k in (1, 2, 3, 4, 6):
    pool = [capi.Engine(0) for _ in range(k)]
    for q in pool:
        q.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
        for _ in range(3):
            q.score_pnp_run(True)
    for q in pool:
        q.sync()
    reps = 600 // k
    t0 = time.perf_counter()
    for _ in range(reps):
        for q in pool:
            q.score_pnp_run(True)
    th = time.perf_counter() - t0
    for q in pool:
        q.sync()
    tt = time.perf_counter() - t0
    ms = tt * 1e3 / (reps * k)
    print("in flight %d: %.2f us per launch (host enqueue %.2f us)  %.1f %% of FP32 peak" % (k, ms * 1e3, th * 1e6 / (reps * k), 100 * H * N * 31 / (ms * 1e-3) / 1e12 / fp32))
    for q in pool:
        q.close()
