"""Developer probe: resident cfg4 sweeps, NF engines in flight (CUDA graphs), ms per sweep.  RSAC_DBG_SKIP_SELECT=1 (only honoured by
a library built with -DRSAC_TIMING_EXPERIMENTS) drops the replay kernel: the results are then wrong and only the timing means
something -- what the replay costs when sweeps overlap (round 2: 0.212 ms per sweep without it, 0.383 ms with it)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
C, n = 1024, 500
NF = int(os.environ.get("NF", "6"))
offsets = (np.arange(C + 1) * n).astype(np.int32)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
pool = []
for i in range(NF):
    b = synth.pnp_batch(4, C, n, 0.5, first=i * C)
    q = capi.Engine(0)
    q.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
    for _ in range(3):
        q.pnp_run(capi.FLAG_EARLY_EXIT)
    pool.append(q)
for q in pool:
    q.sync()
reps = 40
t0 = time.perf_counter()
for _ in range(reps):
    for q in pool:
        q.pnp_run(capi.FLAG_EARLY_EXIT)
for q in pool:
    q.sync()
dt = (time.perf_counter() - t0) / (reps * NF)
print("NF=%d skip_solve=%s skip_score=%s skip_select=%s: %.4f ms per sweep" % (NF, os.environ.get("RSAC_DBG_SKIP_SOLVE"), os.environ.get("RSAC_DBG_SKIP_SCORE"),
                                                                   os.environ.get("RSAC_DBG_SKIP_SELECT"), dt * 1e3))
