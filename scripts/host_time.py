"""Developer probe: host-side (CPU) time of the upload / run calls of a cfg4 sweep."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
import torch
C, n = 1024, 500
eng = capi.Engine(0)
b = synth.pnp_batch(4, C, n, 0.5)
offsets = np.arange(C + 1, dtype=np.int32) * n
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_p3d, h_p2d, h_s2 = pin(b["p3d"].reshape(-1, 3)), pin(b["p2d"].reshape(-1, 2)), pin(b["sigma2"].reshape(-1))
for flags in (0, capi.FLAG_EARLY_EXIT):
    for it in range(3):
        eng.pnp_upload(offsets, h_p3d.numpy(), h_p2d.numpy(), h_s2.numpy(), [b["K"]], prm, seeds=b["seeds"]); eng.pnp_run(flags)
    eng.sync()
    tu = tr = 0.0
    K = 20
    for it in range(K):
        t0 = time.perf_counter()
        eng.pnp_upload(offsets, h_p3d.numpy(), h_p2d.numpy(), h_s2.numpy(), [b["K"]], prm, seeds=b["seeds"])
        t1 = time.perf_counter()
        eng.pnp_run(flags)
        t2 = time.perf_counter()
        eng.sync()
        tu += t1 - t0; tr += t2 - t1
    print("flags=%d host ms: upload %.3f run %.3f" % (flags, tu / K * 1e3, tr / K * 1e3))
