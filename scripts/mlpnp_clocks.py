import sys, ctypes
sys.path.insert(0, "orb-slam2-optimized_b200")
import numpy as np
from ransac_b200 import capi, synth
eng = capi.Engine(0)
C2, N2 = 64, 1000
b2 = synth.pnp_batch(2, C2, N2, 0.5)
cov = np.stack([synth.bearing_covariances(dict(K=b2["K"], sigma2=b2["sigma2"][c])) for c in range(C2)])
eng.mlpnp_upload((np.arange(C2 + 1) * N2).astype(np.int32), b2["p3d"], b2["p2d"], b2["sigma2"], np.array([b2["K"]], np.float32),
                 capi.ransac_params(0.99, 10, 300, 6, 0.2, 5.991), cov=cov, seeds=b2["seeds"])
for _ in range(3):
    eng.mlpnp_run()
eng.sync()
clk = (ctypes.c_longlong * 8)()
eng.L.rsac_debug_mlpnp_clocks(eng.h, clk)
c = list(clk)
names = ["nullspaces+weights", "AtPA", "eigen-solve", "recover", "gauss-newton", "exit"]
print([(names[i - 1], c[i] - c[i - 1]) for i in range(1, 7)], "total", c[6] - c[0], "GN iterations", c[7])
eng.profile_enable(True); eng.profile_reset(); eng.mlpnp_run(); eng.sync(); print(eng.profile_trace())
