"""torchrun worker for tests/test_gpu_multirank.py: a cfg4-style sweep sharded over WORLD_SIZE GPUs must gather, through
the library's own NCCL path (rsac_nccl_allgather_results) and through torch.distributed, records that are byte-identical
to a one-rank run of the same sweep."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
from ransac_b200 import capi, shard, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    C, n = 96, 500                              # not a multiple of every world size: the last shard is shorter
    prm = capi.ransac_params(0.99, 10, 300, 4, 0.2, 5.991)
    b = synth.pnp_batch(4, C, n, 0.5)
    first, count = shard.block_range(C, rank, world)
    cap = shard.per_rank_capacity(C, world)
    eng = capi.Engine(torch.cuda.current_device())
    eng.set_problem_base(first)
    d_local = torch.full((cap, shard.REC_WORDS), -1, dtype=torch.int32, device=dev)
    off = (np.arange(count + 1) * n).astype(np.int32)
    eng.pnp_upload(off, b["p3d"][first:first + count], b["p2d"][first:first + count], b["sigma2"][first:first + count], [b["K"]], prm,
                   seeds=b["seeds"][first:first + count])
    eng.pnp_run(capi.FLAG_EARLY_EXIT, d_local.data_ptr())
    eng.sync()
    g_torch = torch.empty((world * cap, shard.REC_WORDS), dtype=torch.int32, device=dev)
    g_native = torch.zeros_like(g_torch)
    dist.all_gather_into_tensor(g_torch, d_local)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.Engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    eng.nccl_init(uid.cpu().numpy().tobytes(), rank, world)
    eng.nccl_allgather_results(d_local.data_ptr(), cap, g_native.data_ptr())
    eng.sync()
    eng.nccl_destroy()
    rec_t, rec_n = shard.records_from_tensor(g_torch), shard.records_from_tensor(g_native)
    assert rec_t.tobytes() == rec_n.tobytes(), "native all-gather differs from torch.distributed's"
    assert len(rec_t) == C and (rec_t["problem"] == np.arange(C)).all()
    if rank == 0:
        one = capi.Engine(torch.cuda.current_device())
        r1, _ = one.pnp_solve((np.arange(C + 1) * n).astype(np.int32), b["p3d"], b["p2d"], b["sigma2"], [b["K"]], prm, seeds=b["seeds"],
                              flags=capi.FLAG_EARLY_EXIT)
        assert r1.tobytes() == rec_t.tobytes(), "N-rank records differ from the 1-rank run"
        assert shard.first_verified(rec_t) == shard.first_verified(r1)
        one.close()
        print("multirank ok: world %d, %d candidates, %d verified" % (world, C, int(rec_t["ok"].sum())))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
