"""CPU suite, part 3: the N > 1 path on gloo, world_size 2.  Candidates shard in contiguous blocks,
each rank fills its per-candidate records, one all-gather exchanges them, and every rank rebuilds the
same index-ordered result list.  The records here come from the oracle (the checker,
allowed in tests/); on the GPU box the same layer gathers the engine's device records over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, C, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
    import oracle_api as O
    from ransac_b200 import capi, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.block_range(C, rank, world)
    cap = shard.per_rank_capacity(C, world)
    rec = np.zeros(cap, capi.RESULT_DTYPE)
    rec["problem"] = -1
    prm = O.params(0.99, 10, 300, 4, 0.2, 5.991)
    for i in range(count):
        g = first + i
        p = synth.pnp_problem(4000 + g, 120, 0.5)
        pb = O.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
        _, H = O.ransac_setup_pnp(120, prm)
        r = O.pnp_ransac(pb, prm, O.index_table(4000 + g, 120, 4, H))
        rec[i]["ok"], rec[i]["no_more"], rec[i]["n_inliers"], rec[i]["best_hyp"] = r["ok"], r["no_more"], r["n_inliers"], r["best_hyp"]
        rec[i]["R"] = r["T"][:3, :3].reshape(-1)
        rec[i]["t"] = r["T"][:3, 3]
        rec[i]["s"] = 1.0
        rec[i]["problem"] = g
    local = torch.from_numpy(rec.view(np.int32).reshape(cap, shard.REC_WORDS).copy())
    gathered = shard.gather_records(local, C, world)
    out = shard.records_from_tensor(gathered)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("C", [7, 8])
def test_gloo_world2_gather_equals_single_rank(tmp_path, C, oracle):
    from ransac_b200 import capi, shard, synth
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, C, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert r0.tobytes() == r1.tobytes()                      # every rank holds the same gathered list
    assert (r0["problem"] == np.arange(C)).all()             # padding dropped, index order restored
    # byte-identical to the single-rank computation
    prm = oracle.params(0.99, 10, 300, 4, 0.2, 5.991)
    for g in range(C):
        p = synth.pnp_problem(4000 + g, 120, 0.5)
        pb = oracle.pnp_problem(p["p3d"], p["p2d"], p["sigma2"], p["K"])
        _, H = oracle.ransac_setup_pnp(120, prm)
        r = oracle.pnp_ransac(pb, prm, oracle.index_table(4000 + g, 120, 4, H))
        assert r0[g]["ok"] == r["ok"] and r0[g]["n_inliers"] == r["n_inliers"]
        assert np.array_equal(r0[g]["R"], r["T"][:3, :3].reshape(-1)) and np.array_equal(r0[g]["t"], r["T"][:3, 3])
    first = shard.first_verified(r0, 30)
    assert first == next((int(x["problem"]) for x in r0 if x["ok"] and x["n_inliers"] >= 30), -1)


def _worker_poseopt(rank, world, port, C, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "orb-slam2-optimized_b200"))
    import oracle_api as O
    from ransac_b200 import capi, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard.block_range(C, rank, world)
    cap = shard.per_rank_capacity(C, world)
    rec = np.zeros(cap, capi.POSEOPT_DTYPE)
    rec["problem"] = -1
    for i in range(count):
        g = first + i
        p = synth.poseopt_problem(4100 + g, 80, 0.2, 0.5 * (g % 2))
        d, _ = O.pose_optimization(O.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        for k in ("n_inliers", "n_bad", "rounds", "iterations", "trials"):
            rec[i][k] = d[k]
        rec[i]["R"], rec[i]["t"], rec[i]["Rf"], rec[i]["tf"] = d["R"].ravel(), d["t"], d["Rf"].ravel(), d["tf"]
        rec[i]["problem"] = g
    words = capi.POSEOPT_DTYPE.itemsize // 4
    local = torch.from_numpy(rec.view(np.int32).reshape(cap, words).copy())
    out = shard.records_from_tensor(shard.gather_records(local, C, world), capi.POSEOPT_DTYPE)
    np.save(os.path.join(out_dir, f"po_rank{rank}.npy"), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_pose_optimization_records(tmp_path, oracle):
    """frames of a PoseOptimization batch shard like candidates: contiguous blocks, one all-gather of the 168-byte
    records, index order restored on every rank"""
    from ransac_b200 import synth
    C, world = 5, 2
    mp.spawn(_worker_poseopt, args=(world, _free_port(), C, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "po_rank0.npy"), np.load(tmp_path / "po_rank1.npy")
    assert r0.tobytes() == r1.tobytes() and (r0["problem"] == np.arange(C)).all()
    for g in range(C):
        p = synth.poseopt_problem(4100 + g, 80, 0.2, 0.5 * (g % 2))
        d, _ = oracle.pose_optimization(oracle.poseopt_problem(p["p3d"], p["obs"], p["inv_sigma2"], p["K"], p["Rcw"], p["tcw"]))
        assert r0[g]["n_inliers"] == d["n_inliers"] and np.array_equal(r0[g]["R"], d["R"].ravel()) and np.array_equal(r0[g]["t"], d["t"])
