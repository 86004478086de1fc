"""Tuning sweep of the scoring kernel on cfg5 (developer tool): planner overrides via RSAC_SCORE_* env vars."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

eng = capi.Engine(0)
H, N = 4096, 10000
p = synth.scoring_stress(5000, H, N)
max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
ref = None
rows = []
for hpl, warps, ctas, cw in itertools.product((1, 2, 4), (4, 8), (0, 2, 3, 4), (1, 2)):
    os.environ["RSAC_SCORE_HPL"] = str(hpl)
    os.environ["RSAC_SCORE_WARPS"] = str(warps)
    os.environ["RSAC_SCORE_CTAS"] = str(ctas)
    os.environ["RSAC_SCORE_CW"] = str(cw)
    eng.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
    for _ in range(5):
        eng.score_pnp_run(True)
    eng.sync()
    eng.timer_begin()
    K = 100
    for _ in range(K):
        eng.score_pnp_run(True)
    ms = eng.timer_end() / K
    counts, _ = eng.score_pnp_download(False)
    if ref is None:
        ref = counts
    ok = bool((counts == ref).all())
    tf = H * N * 31 / (ms * 1e-3) / 1e12
    rows.append((ms, hpl, warps, ctas, cw, tf, ok))
    print("hpl=%d warps=%d ctas=%d cw=%d: %.4f ms  %.2f TFLOP/s  same=%s" % (hpl, warps, ctas, cw, ms, tf, ok), flush=True)
rows.sort()
print("best:", rows[:5])
