"""Developer probe: where the end-to-end (host buffers) sweep loses time against the resident one."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, synth, shard
import torch
C, n = 1024, 500
dev = torch.device("cuda", 0)
b = synth.pnp_batch(4, C, n, 0.5)
offsets = np.arange(C + 1, dtype=np.int32) * n
prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_p3d, h_p2d, h_s2 = pin(b["p3d"].reshape(-1, 3)), pin(b["p2d"].reshape(-1, 2)), pin(b["sigma2"].reshape(-1))
NS = 2
engines, streams = [], []
for i in range(NS):
    e = capi.Engine(0); s = torch.cuda.Stream(device=dev); e.set_stream(s.cuda_stream); engines.append(e); streams.append(s)
words = int(((np.diff(offsets) + 31) // 32).sum())
h_res = [torch.empty((C, 24), dtype=torch.int32).pin_memory() for _ in range(NS)]
h_msk = [torch.empty((words,), dtype=torch.int32).pin_memory() for _ in range(NS)]
FL = capi.FLAG_EARLY_EXIT
cd = [None]

def step(i, up=True, chain=True, down=True, prof=False):
    with torch.cuda.stream(streams[i]):
        if up:
            engines[i].pnp_upload(offsets, h_p3d.numpy(), h_p2d.numpy(), h_s2.numpy(), [b["K"]], prm, seeds=b["seeds"])
        if chain and cd[0] is not None:
            streams[i].wait_event(cd[0])
        engines[i].pnp_run(FL)
        ev = torch.cuda.Event(); ev.record(streams[i]); cd[0] = ev
        if down:
            engines[i].pnp_download_async(h_res[i].data_ptr(), h_msk[i].data_ptr())

def timed(K, **kw):
    cd[0] = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        step(k % NS, **kw)
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3, th / K * 1e3

for i in range(NS):
    step(i)
torch.cuda.synchronize()
for kw in (dict(up=False, down=False), dict(up=False, down=True), dict(up=True, down=False), dict(up=True, down=True), dict(up=True, down=True, chain=False)):
    timed(10, **kw)
    ms, host = timed(100, **kw)
    print(kw, "ms/step %.3f (host enqueue %.3f)" % (ms, host))
# one engine, serial, with per-launch trace
e = engines[0]
e.profile_enable(True); e.profile_reset()
with torch.cuda.stream(streams[0]):
    for k in range(3):
        e.pnp_upload(offsets, h_p3d.numpy(), h_p2d.numpy(), h_s2.numpy(), [b["K"]], prm, seeds=b["seeds"])
        e.pnp_run(FL)
torch.cuda.synchronize()
tr = e.profile_trace()
print("serial upload+run trace:", " ".join("%s %.3f" % (k, m) for k, m in tr[-(len(tr) // 3):]))
