"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): runs tools/dist_check.py under torchrun,
one rank per GPU, NCCL halo exchange / allreduce / coarse-level gather, against the CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_parity():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "DIST CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
