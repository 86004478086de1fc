"""Developer timing probe: stage times of the MLPnP batch (cfg2) and the Sim3 batch (cfg3)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

eng = capi.Engine(0)
C2, N2 = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 1000
b2 = synth.pnp_batch(2, C2, N2, 0.5)
cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
Kf = np.array([b2["K"]], np.float32)
off2 = (np.arange(C2 + 1) * N2).astype(np.int32)
prm2 = capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991)
eng.mlpnp_upload(off2, b2["p3d"], b2["p2d"], b2["sigma2"], Kf, prm2, cov=cov, seeds=b2["seeds"])
for _ in range(3):
    eng.mlpnp_run()
eng.sync()
eng.profile_enable(True)
eng.profile_reset()
eng.timer_begin()
for _ in range(5):
    eng.mlpnp_run()
ms = eng.timer_end() / 5
print("cfg2 C=%d: %.3f ms/batch" % (C2, ms))
for k, (tms, nl) in eng.profile().items():
    if nl:
        print("   %-7s %8.3f ms/launch (%d launches)" % (k, tms / nl, nl))
res, _ = eng.mlpnp_download()
print("   ok:", int(res["ok"].sum()), "mean n_hyp", res["n_hyp"].mean(), "n_refines", res["n_refines"].mean(), "exact", eng.score_exact_evals())
