"""Developer probe: exact-tier evaluations and per-launch times of a cfg4 sweep on the two synthetic generators."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth
C, n = 1024, 500
offsets = (np.arange(C + 1) * n).astype(np.int32)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
A = synth.pnp_batch(4, C, n, 0.5)
B = synth.reloc_frame(4000, C, 2000, n, 0.5, 200000)
B2 = dict(B); B2["seeds"] = A["seeds"]
A2 = dict(A); A2["seeds"] = B["seeds"]
for name, b in (("reloc_frame", B), ("pnp_batch", A), ("reloc_frame seeds of pnp_batch", B2), ("pnp_batch seeds of reloc", A2), ("reloc_frame", B)):
    eng = capi.Engine(0)
    eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
    for flags in (0, capi.FLAG_EARLY_EXIT):
        for _ in range(3):
            eng.pnp_run(flags)
        eng.sync()
        ex = eng.score_exact_evals()
        eng.profile_enable(True); eng.profile_reset()
        eng.pnp_run(flags); eng.sync()
        tr = eng.profile_trace()
        eng.profile_enable(False)
        res, _ = eng.pnp_download()
        print(name, "flags", flags, "exact evals (last scoring launch)", ex, "ok", int(res["ok"].sum()), "mean inl", float(res["n_inliers"].mean()),
              "n_hyp mean", float(res["n_hyp"].mean()))
        print("   ", " ".join("%s %.3f" % (k, m) for k, m in tr))
    X = b["p3d"].reshape(-1, 3)
    print("   |X|inf max", float(np.abs(X).max()), "uv range", float(b["p2d"].min()), float(b["p2d"].max()))
    eng.close()
