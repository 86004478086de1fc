"""Developer timing probe: cfg3 Sim3 RANSAC (200 matches, 300 iterations) for 1, 8, 64 candidates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

eng = capi.Engine(0)
ps = [synth.sim3_problem(3000 + i, 200, 0.4, 1.0) for i in range(64)]
cat = lambda k: np.concatenate([q[k] for q in ps])
off3 = (np.arange(len(ps) + 1) * 200).astype(np.int32)
K3 = np.array([ps[0]["K"]], np.float32)
prm3 = capi.Sim3Params(0.99, 20, 300, 1)
seeds3 = np.arange(len(ps), dtype=np.uint32) + 3000
for C3 in (1, 8, 64):
    o3 = off3[:C3 + 1]
    n3 = int(o3[-1])
    eng.sim3_upload(o3, cat("x1c")[:n3], cat("x2c")[:n3], cat("sigma2_1")[:n3], cat("sigma2_2")[:n3], K3, K3, prm3, seeds=seeds3[:C3])
    for _ in range(3):
        eng.sim3_run()
    eng.sync()
    eng.timer_begin()
    for _ in range(20):
        eng.sim3_run()
    ms3 = eng.timer_end() / 20
    res, _ = eng.sim3_download()
    print("sim3 C=%d: %.4f ms per batch, %.2f G evals/s, ok %d" % (C3, ms3, C3 * 300 * 200 / ms3 * 1e-6, int(res["ok"].sum())))
