"""One OptimizeSim3 batch (256 keyframe pairs x 100 matches) for ncu / timing.  usage: python scripts/sim3opt_prof.py [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100
pp = [synth.sim3opt_problem(8000 + i, N, 0.15) for i in range(C)]
off = (np.arange(C + 1) * N).astype(np.int32)
cat = lambda k: np.concatenate([q[k] for q in pp])
K = np.stack([q["K"] for q in pp])
eng = capi.Engine(0)
res, rem = eng.sim3opt_solve(off, cat("x1c"), cat("x2c"), cat("obs1"), cat("obs2"), cat("inv_sigma2_1"), cat("inv_sigma2_2"), K, K,
                             np.stack([q["S12"] for q in pp]), 10.0)
eng.sync()
eng.timer_begin()
for _ in range(reps):
    eng.sim3opt_run()
ms = eng.timer_end() / reps
print(f"sim3opt {C} x {N}: {ms:.4f} ms per batch, {C / ms * 1e3:.0f} pairs/s, iterations {res['iterations'].mean():.1f}, "
      f"trials {res['trials'].mean():.1f}, inliers {res['n_inliers'].mean():.1f}")
