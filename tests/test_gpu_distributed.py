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


@pytest.mark.gpu
def test_two_gpu_parity_fused_ghost_push():
    """The same check with the plane-per-step kernel on every level (PMG_TILE_VARIANT=6), so that the smoother's applies
    exchange their ghost planes by the fused push (csrc/pmg_apply_plane_launch.h) also on the check's small meshes; and once
    with the push switched off (PMG_FUSED_HALO=0): same results, an exchange before every apply."""
    import re
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    counts = {}
    for fused in ("1", "0"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29534", os.path.join(ROOT, "tools", "dist_check.py")]
        env = dict(os.environ, PMG_TILE_VARIANT="6", PMG_FUSED_HALO=fused)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        assert "DIST CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
        counts[fused] = int(re.search(r"fused ghost push \(no exchange of their own\): (\d+)", out.stdout).group(1))
    assert counts["0"] == 0
    if counts["1"] == 0:
        pytest.skip("peer mapping (CUDA IPC) not available on this box: the fused push stayed off")
