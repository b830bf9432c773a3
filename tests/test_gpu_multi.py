"""Multi-GPU parity (one process per GPU over NCCL + CUDA IPC): runs tools/multi_gpu_check.py under torchrun when the box has
>= 2 GPUs; skipped on a single-GPU box (the world_size-2 host logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity_and_halo():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "multi-gpu check passed" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
