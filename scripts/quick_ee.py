"""Developer timing probe: cfg4 sweep, exhaustive vs early exit in phases, over first-phase sizes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
HAS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 32, 40, 48, 56, 64, 96, 128]
eng = capi.Engine(0)
n = 500
b = synth.pnp_batch(4, C, n, 0.5)
offsets = np.arange(C + 1, dtype=np.int32) * n
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])


def run(flags, K=8):
    for it in range(3):
        eng.pnp_run(flags)
    eng.sync()
    eng.profile_enable(False)
    eng.timer_begin()
    for it in range(K):
        eng.pnp_run(flags)
    ms = eng.timer_end() / K
    eng.profile_enable(True)
    eng.profile_reset()
    for it in range(4):
        eng.pnp_run(flags)
    eng.sync()
    prof = {k: (t / 4, nl // 4) for k, (t, nl) in eng.profile().items() if nl}
    tr = eng.profile_trace()
    per = len(tr) // 4
    print("      last sweep:", " ".join("%s %.3f" % (k, m) for k, m in tr[-per:]))
    eng.profile_enable(False)
    return ms, prof


ms, prof = run(0)
res0, m0 = eng.pnp_download()
print("exhaustive: %.3f ms/sweep  %.0f cand/s  %s" % (ms, C / ms * 1e3, {k: "%.3f/%d" % v for k, v in prof.items()}))
for ha in HAS:
    eng.set_first_phase(ha)
    ms, prof = run(capi.FLAG_EARLY_EXIT)
    res1, m1 = eng.pnp_download()
    st = eng.pnp_phase_stats()
    same = all((res0[f] == res1[f]).all() for f in ("ok", "n_inliers", "best_hyp", "n_refines", "n_hyp")) and (m0 == m1).all()
    print("HA=%3d (used %3d): %.3f ms/sweep  %.0f cand/s  B=%d C=%d solved=%.1f%%  same=%s  %s" % (
        ha, st[0], ms, C / ms * 1e3, st[1], st[2], 100.0 * st[3] / (C * 300), same, {k: "%.3f/%d" % v for k, v in prof.items()}))

import ctypes
clk = (ctypes.c_longlong * 16)()
eng.set_first_phase(0)
eng.pnp_run()
eng.sync()
eng.L.rsac_debug_select_clocks(eng.h, clk)
c = list(clk)
names = ["start", "found", "refine:begin", "presums", "MtM", "jacobi", "betas", "sums2", "horn+reproj", "score", "end"]
print("select phases (block 0, cycles):", [(names[i], c[i] - c[i - 1]) for i in range(1, 11)], "total", c[10] - c[0])
print("12x12 eigen-solve: forward sweeps", c[11] - c[4], "cycles over", c[12], "sweeps; selection + back-application", c[5] - c[11])
