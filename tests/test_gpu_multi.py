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


def test_two_devices_in_one_process():
    """Handles on different GPUs of one process: per-device kernel attributes, upload pools
    and streams; the same shard evaluated on cuda:0 and cuda:1 gives the same bits."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from anml_b200 import _native as N

    rng = np.random.default_rng(2)
    n, p = 300_007, 64
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    obs = (rng.uniform(size=n) < 0.5).astype(float)
    x = rng.normal(size=p) * 0.1
    outs = []
    for dev in (0, 1, 0):
        d = N.Design(n, p, device=dev)
        d.set_columns(0, X)
        m = N.NativeModel(n, "binomial", 1, device=dev)
        m.set_param(0, d, N.LINK_CODES["Expit"])
        m.set_obs(obs)
        outs.append(m.eval(x))
        outs.append(m.eval(x, N.WANT_OBJ | N.WANT_GRAD))
    for o in outs[2::2]:
        assert o[0] == outs[0][0] and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
    wide = rng.normal(size=(2000, 130))
    for dev in (1, 0):
        d = N.Design(2000, 130, device=dev)
        d.set_columns(0, wide)
        m = N.NativeModel(2000, "binomial", 1, device=dev)
        m.set_param(0, d, N.LINK_CODES["Expit"])
        m.set_obs(obs[:2000])
        o = m.eval(np.zeros(130))
        assert np.isfinite(o[0]) and np.array_equal(o[2], o[2].T)
