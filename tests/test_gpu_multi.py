"""Row-sharded evaluation over NCCL on real GPUs: world_size 2, one
process per GPU under torchrun.  Skips on a single-GPU box; the same logic runs on
CPU over gloo in test_distributed_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_sharded_eval_and_fit():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nccl_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), worker]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "NCCL_SHARDED_OK world=2" in out.stdout
