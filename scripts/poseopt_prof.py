"""One cfg4-sized PoseOptimization batch (1024 frames x 250 edges) for ncu / timing.  usage: python scripts/poseopt_prof.py [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
N = int(sys.argv[3]) if len(sys.argv) > 3 else 250
pp = [synth.poseopt_problem(4000 + i, N, 0.2, 0.0) for i in range(C)]
off = (np.arange(C + 1) * N).astype(np.int32)
cat = lambda k: np.concatenate([q[k] for q in pp])
eng = capi.Engine(0)
eng.poseopt_upload(off, cat("p3d"), cat("obs"), cat("inv_sigma2"), np.stack([q["K"] for q in pp]),
                   np.stack([np.concatenate([q["Rcw"].ravel(), q["tcw"]]) for q in pp]))
eng.poseopt_run()
eng.sync()
eng.timer_begin()
for _ in range(reps):
    eng.poseopt_run()
ms = eng.timer_end() / reps
res, out = eng.poseopt_download()
print(f"poseopt {C} x {N}: {ms:.4f} ms per batch, {C / ms * 1e3:.0f} frames/s, iterations {res['iterations'].mean():.1f}, "
      f"trials {res['trials'].mean():.1f}, inliers {res['n_inliers'].mean():.1f}")
