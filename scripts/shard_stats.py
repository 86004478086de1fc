"""Developer probe: early-exit phase statistics and sweep time of the eight cfg4 shards a weak-scaling run uses."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
C, n = 1024, 500
eng = capi.Engine(0)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
offsets = np.arange(C + 1, dtype=np.int32) * n
for shard in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    b = synth.pnp_batch(4, C, n, 0.5, first=shard * C)
    eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
    for it in range(3):
        eng.pnp_run(capi.FLAG_EARLY_EXIT)
    eng.sync()
    eng.timer_begin()
    for it in range(10):
        eng.pnp_run(capi.FLAG_EARLY_EXIT)
    ms = eng.timer_end() / 10
    res, _ = eng.pnp_download()
    st = eng.pnp_phase_stats()
    eng.profile_enable(True); eng.profile_reset()
    eng.pnp_run(capi.FLAG_EARLY_EXIT); eng.sync()
    tr = eng.profile_trace(); eng.profile_enable(False)
    poses, counts = eng.pnp_hypotheses()
    cnt = counts.reshape(C, 300)[:, :st[0]]
    best = cnt.max(axis=1)
    multi = res["n_refines"] > 1
    print("shard %d: %.3f ms  B=%d C=%d  ok=%d  n_refines>1: %d  max n_refines %d | best-in-A of multi-refine candidates: %s" % (
        shard, ms, st[1], st[2], int(res["ok"].sum()), int(multi.sum()), int(res["n_refines"].max()), sorted(best[multi].tolist())[:12]))
    print("     ", " ".join("%s %.3f" % (k, m) for k, m in tr))
