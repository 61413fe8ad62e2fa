"""Multi-GPU parity: launches tests/multi_gpu_check.py under torchrun -- one rank per GPU when the box has
at least two, and ALWAYS two ranks on cuda:0 (gloo for the small collectives, CUDA IPC for the pairs), so that
the fused partition + peer exchange and the shard sort are exercised on a single-GPU box too."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_two_rank_sharded_sort_matches_oracle():
    if _n_gpus() < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU parity ok" in out.stdout


def test_two_ranks_on_one_gpu_sharded_sort_matches_oracle():
    if _n_gpus() < 1:
        pytest.skip("needs a GPU")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "multi_gpu_check.py"),
           "--same-gpu"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU parity ok" in out.stdout
    assert "exchange=peer" in out.stdout, "the peer exchange was not exercised:\n" + out.stdout[-2000:]
