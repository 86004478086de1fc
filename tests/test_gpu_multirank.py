"""On-hardware rank identity: N ranks' gathered records == the 1-rank run, byte for byte, through
torch.distributed and through rsac_nccl_allgather_results.  Needs two GPUs on the box; bench.py asserts the same at
every N > 1 (its `checks` object)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_one_rank(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (bench.py --gpus N runs the same check on the multi-GPU box)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multirank ok" in r.stdout
