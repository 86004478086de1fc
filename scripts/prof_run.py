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
