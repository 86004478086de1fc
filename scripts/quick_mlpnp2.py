"""Developer probe: cfg2 (MLPnP, 64 frames x 1000 matches, bearing covariances) stage times."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
eng = capi.Engine(0)
C2, N2 = 64, 1000
b2 = synth.pnp_batch(2, C2, N2, 0.5)
cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
Kf = np.array([b2["K"]], np.float32)
off2 = (np.arange(C2 + 1) * N2).astype(np.int32)
prm2 = capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991)
eng.mlpnp_upload(off2, b2["p3d"], b2["p2d"], b2["sigma2"], Kf, prm2, cov=cov, seeds=b2["seeds"])
for _ in range(3):
    eng.mlpnp_run()
eng.sync()
eng.profile_enable(True); eng.profile_reset()
for _ in range(5):
    eng.mlpnp_run()
eng.sync()
print({k: "%.3f" % (t / max(n, 1)) for k, (t, n) in eng.profile().items() if n})
