"""Developer probe: where the latency of ONE candidate (cfg1: 500 matches, H = 300) goes -- host wall clock of rsac_pnp_solve,
resident run + download, and the per-launch trace of one run."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
n = 500
b = synth.pnp_batch(1, 1, n, 0.5)
off = np.array([0, n], np.int32)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
eng = capi.Engine(0)
def wall(fn, reps=50):
    for _ in range(5): fn()
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    eng.sync()
    return (time.perf_counter() - t0) / reps * 1e3
for fl, nm in ((0, "exhaustive"), (capi.FLAG_EARLY_EXIT, "early-exit")):
    print(nm, "pnp_solve wall ms", wall(lambda: eng.pnp_solve(off, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"], flags=fl)))
    eng.pnp_upload(off, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
    print(nm, "upload wall ms", wall(lambda: eng.pnp_upload(off, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])))
    print(nm, "run (resident, async issue + sync each) ms", wall(lambda: (eng.pnp_run(fl), eng.sync())))
    print(nm, "run+download ms", wall(lambda: (eng.pnp_run(fl), eng.pnp_download())))
    eng.set_graphs(False)
    eng.profile_enable(True)
    eng.pnp_run(fl); eng.sync()
    eng.profile_reset()
    eng.pnp_run(fl); eng.sync()
    tr = eng.profile_trace()
    print(nm, "trace", [(s, round(m, 4)) for s, m in tr], "sum", round(sum(m for _, m in tr), 4))
    eng.profile_enable(False)
    eng.set_graphs(True)
