"""Developer timing probe (not the bench contract): stage times of the PnP sweep (cfg4) and the
scoring stress (cfg5) with CUDA events, plus measured FP32/FP64 CUDA-core peaks."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = capi.Engine(0)
print(eng.device_info())
print("peaks TFLOP/s fp32, fp64:", eng.measure_peaks())

n = 500
t0 = time.time()
b = synth.pnp_batch(4, C, n, 0.5)
print("synth %.2fs" % (time.time() - t0))
offsets = np.arange(C + 1, dtype=np.int32) * n
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
for it in range(3):
    eng.pnp_run()
eng.sync()
eng.profile_enable(True)
eng.profile_reset()
eng.timer_begin()
K = 5
for it in range(K):
    eng.pnp_run()
ms = eng.timer_end()
print("cfg4 C=%d: %.3f ms/sweep, %.3g evals/s, %.3g cand/s" % (C, ms / K, C * 300 * n / (ms / K * 1e-3), C / (ms / K * 1e-3)))
for k, (tms, nl) in eng.profile().items():
    if nl:
        print("   %-7s %8.3f ms/launch (%d launches)" % (k, tms / nl, nl))
res, masks = eng.pnp_download()
print("   ok:", int(res["ok"].sum()), "refined:", int(res["refined"].sum()), "mean n_hyp", res["n_hyp"].mean(), "exact evals", eng.score_exact_evals())
# e2e through host buffers
t0 = time.time()
for it in range(3):
    eng.pnp_solve(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
dt = (time.time() - t0) / 3
print("   e2e (pageable host buffers, wall): %.3f ms" % (dt * 1e3))

# cfg5
H, N = 4096, 10000
p = synth.scoring_stress(5000, H, N)
max_err = (p["sigma2"] * np.float32(5.991)).astype(np.float32)
eng.profile_enable(False)
eng.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
for want in (True, False):
    for it in range(5):
        eng.score_pnp_run(want)
    eng.sync()
    eng.timer_begin()
    K = 50
    for it in range(K):
        eng.score_pnp_run(want)
    ms = eng.timer_end() / K
    ev = H * N / (ms * 1e-3)
    print("cfg5 masks=%s: %.4f ms, %.3g evals/s, %.2f TFLOP/s (31 FLOP/eval), exact evals %d (%.2e)" %
          (want, ms, ev, ev * 31 / 1e12, eng.score_exact_evals(), eng.score_exact_evals() / (H * N)))

# phase clocks of the replay kernel (block 0)
import ctypes
clk = (ctypes.c_longlong * 16)()
eng.profile_enable(False)
eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
eng.pnp_run()
eng.sync()
eng.L.rsac_debug_select_clocks(eng.h, clk)
c = list(clk)
names = ["start", "found", "refine:begin", "presums", "MtM", "jacobi", "betas", "sums2", "horn+reproj", "score", "end"]
print("select phases (block 0, cycles):", [(names[i], c[i] - c[i - 1]) for i in range(1, 11)], "total", c[10] - c[0])

# cfg5 kernel-only duration (events bracket the kernel launch, not the memset)
eng.score_pnp_upload(p["poses"], p["p3d"], p["p2d"], max_err, p["K"])
for want in (True, False):
    for it in range(5):
        eng.score_pnp_run(want)
    eng.sync()
    eng.profile_enable(True)
    eng.profile_reset()
    for it in range(50):
        eng.score_pnp_run(want)
    eng.sync()
    tms, nl = eng.profile()["score"]
    eng.profile_enable(False)
    ms = tms / nl
    print("cfg5 masks=%s kernel only: %.4f ms -> %.2f TFLOP/s (%.1f%% of 71.7)" % (want, ms, H * N * 31 / (ms * 1e-3) / 1e12, H * N * 31 / (ms * 1e-3) / 1e12 / 71.7 * 100))

eng.score_pnp_run(True)
eng.sync()
sc = (ctypes.c_ulonglong * 64)()
eng.L.rsac_debug_score_clocks(eng.h, sc)
sc = np.array(list(sc), dtype=np.int64).reshape(8, 8)
t0 = sc[:, 0].min()
print("score CTA timelines (ns from first entry): entry, first chunk, folded, last chunk done, exit, chunks")
for r in sc:
    print("   ", [int(x - t0) for x in r[:5]], int(r[5]))
al = (ctypes.c_ulonglong * 4096)()
eng.L.rsac_debug_score_all(eng.h, al)
al = np.array(list(al), dtype=np.int64).reshape(1024, 4)[:296]
t0 = al[:, 0].min()
dur = al[:, 1] - al[:, 0]
print("all CTAs: entry spread %d ns, exit min/median/max %d/%d/%d ns, chunks min/max %d/%d" %
      (al[:, 0].max() - t0, (al[:, 1] - t0).min(), np.median(al[:, 1] - t0), (al[:, 1] - t0).max(), al[:, 2].min(), al[:, 2].max()))
per = dur / np.maximum(al[:, 2], 1)
order = np.argsort(per)
print("ns per chunk: fastest", per[order[:5]].astype(int), "slowest", per[order[-5:]].astype(int))
print("slowest CTAs (blockIdx, smid, chunks):", [(int(i), int(al[i, 3]), int(al[i, 2])) for i in order[-8:]])
print("fastest CTAs (blockIdx, smid, chunks):", [(int(i), int(al[i, 3]), int(al[i, 2])) for i in order[:8]])
sm_counts = np.bincount(al[:, 3].astype(int), minlength=148)
print("CTAs per SM histogram:", np.bincount(sm_counts))
# per-SM: sum of chunks and max exit
import collections
bysm = collections.defaultdict(list)
for i in range(296):
    bysm[int(al[i, 3])].append((int(al[i, 1] - t0), int(al[i, 2]), i))
late = sorted(bysm.items(), key=lambda kv: -max(x[0] for x in kv[1]))[:5]
print("latest SMs:", late)

eng.pnp_upload(offsets, b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"])
eng.pnp_run()
eng.sync()
eng.L.rsac_debug_solve_clocks(eng.h, clk)
c = list(clk)
names = ["ctrl pts+alphas", "QR nullspace", "L,rho", "approx1", "GN1", "approx2+GN2", "approx3+GN3", "ccs/pcs/M (k=0)", "horn (k=0)", "rest (reproj k=0, k=1,2)"]
print("solve phases (block 0 warp 0, cycles):", [(names[i - 1], c[i] - c[i - 1]) for i in range(1, 11)], "total", c[10] - c[0])
