"""GPU, needs >= 2 devices (skipped on a single-GPU box): the row-sharded gallery over NCCL ranks -- fused
peer-memory candidate exchange == NCCL all-gather path, bit for bit, every rank holds the global answer, and the
sharded answer equals the single-GPU answer."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs on one box")
def test_two_rank_exchange_matches_nccl_and_single_gpu():
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "test_exchange.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "EXCHANGE_OK" in out.stdout, out.stdout[-2000:]
    assert "SINGLE_GPU_EQUAL" in out.stdout, out.stdout[-2000:]
