"""Candidate sharding across ranks.

Each relocalisation / loop-closure candidate is an independent problem with its own
correspondences, RANSAC state and index table (reference src/Tracking.cpp:1207-1232,
src/LoopClosing.cpp:238-265), so the batch shards with no data-path collective: rank r owns a
contiguous block of ceil(C/world) candidates.  The only exchange is one all-gather of the
fixed-size per-candidate record (rsac_result, 96 B); the host then scans candidates IN INDEX
ORDER so that the reference's "first candidate that verifies" rule (Tracking.cpp:1241-1331)
is preserved.
"""
from __future__ import annotations

import numpy as np

from . import capi

REC_WORDS = capi.RESULT_DTYPE.itemsize // 4   # 24 x 4-byte words


def block_range(C: int, rank: int, world: int):
    """[first, first+count) owned by `rank`; identical to rsac_shard_range in the C ABI."""
    per = (C + world - 1) // world
    first = min(C, rank * per)
    last = min(C, first + per)
    return first, last - first


def per_rank_capacity(C: int, world: int) -> int:
    return (C + world - 1) // world


def gather_records(local: "torch.Tensor", C: int, world: int, group=None):
    """all-gather of per-candidate records.

    local: int32 tensor [cap, 24] (cap = per_rank_capacity; unused rows have problem = -1) on the
    device of the backend (CUDA for nccl, CPU for gloo).  Returns int32 [world*cap, 24]."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def records_from_tensor(t, dtype=None) -> np.ndarray:
    """int32 [n, words] tensor/array -> structured record array (rsac_result by default; capi.POSEOPT_DTYPE /
    capi.SIM3OPT_DTYPE for the two optimisers, whose frames / keyframe pairs shard the same way), padding rows
    (problem < 0) dropped, ordered by global problem index."""
    a = t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
    rec = np.ascontiguousarray(a, np.int32).view(dtype or capi.RESULT_DTYPE).reshape(-1)
    rec = rec[rec["problem"] >= 0]
    return rec[np.argsort(rec["problem"], kind="stable")]


def first_verified(rec: np.ndarray, min_inliers: int = 0):
    """index-order scan: the first candidate whose RANSAC returned a pose (and enough inliers)."""
    for r in rec:
        if r["ok"] and r["n_inliers"] >= min_inliers:
            return int(r["problem"])
    return -1
