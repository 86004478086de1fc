"""Developer probe: host-side (CPU) time of the calls of one cfg4 batch -- indexed upload / flat upload / run / async download --
with the device idle (a synchronise between iterations, outside the timed calls)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth, shard
import torch
C, n = 1024, 500
eng = capi.Engine(0)
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
f = synth.reloc_frame(4000, C, 2000, n, 0.5, 200000)
offsets = (np.arange(C + 1) * n).astype(np.int32)
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
kp_idx, mp_idx = pin(f["kp_idx"].reshape(-1)), pin(f["mp_idx"].reshape(-1))
kp_uv, kp_s2 = pin(f["kp_uv"]), pin(f["kp_sigma2"])
p3d, p2d, s2 = pin(f["p3d"].reshape(-1, 3)), pin(f["p2d"].reshape(-1, 2)), pin(f["sigma2"].reshape(-1))
h_res = torch.empty((C, shard.REC_WORDS), dtype=torch.int32).pin_memory()
h_msk = torch.empty((C * 16,), dtype=torch.int32).pin_memory()
d_out = torch.empty((C, shard.REC_WORDS), dtype=torch.int32, device="cuda")
ids = np.arange(C, dtype=np.int32)
eng.pnp_upload_indexed(offsets, kp_idx.numpy(), mp_idx.numpy(), f["K"], prm, seeds=f["seeds"], kp_uv=kp_uv.numpy(), kp_sigma2=kp_s2.numpy(), mp_xyz=f["mp_xyz"])
for mode in ("indexed", "flat"):
    acc = np.zeros(4)
    K = 30
    for it in range(K + 5):
        t0 = time.perf_counter()
        eng.set_problem_ids(ids)
        if mode == "indexed":
            eng.pnp_upload_indexed(offsets, kp_idx.numpy(), mp_idx.numpy(), f["K"], prm, seeds=f["seeds"], kp_uv=kp_uv.numpy(), kp_sigma2=kp_s2.numpy())
        else:
            eng.pnp_upload(offsets, p3d.numpy(), p2d.numpy(), s2.numpy(), [f["K"]], prm, seeds=f["seeds"])
        t1 = time.perf_counter()
        eng.pnp_run(capi.FLAG_EARLY_EXIT, d_out.data_ptr())
        t2 = time.perf_counter()
        eng.pnp_download_async(h_res.data_ptr(), h_msk.data_ptr())
        t3 = time.perf_counter()
        eng.sync()
        t4 = time.perf_counter()
        if it >= 5:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    print("%s: host ms per batch: upload %.3f  run %.3f  download_async %.3f  (then sync %.3f)" % ((mode,) + tuple(acc / K * 1e3)))
