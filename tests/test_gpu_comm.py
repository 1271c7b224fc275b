"""The peer-memory all-reduce kernel (csrc/comm.cuh) on one GPU: a one-rank window (the launch, the
publish / wait / sum phases and the alternating data buffers all run; the sum over one rank is the
identity).  The multi-rank path runs in tests/test_gpu_multi.py (2 GPUs) and, on the host side, in
tests/test_distributed_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_rank_window_is_the_identity_and_epochs_alternate():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from anml_b200 import _native as N

    n = 4226  # the packed record of P = 64
    comm = N.PeerComm(0, 0, 1, n)
    assert len(comm.export()) == N.PeerComm.handle_bytes() == 64
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    for k in range(7):
        buf = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        want = buf.clone()
        comm.allreduce_async(buf.data_ptr(), n, 1)  # cudaStreamLegacy
        torch.cuda.synchronize()
        assert torch.equal(buf, want)
    # a large record spreads over many CTAs (the last one to finish publishes the flag)
    big = N.PeerComm(0, 0, 1, 1 << 20)
    buf = torch.randn(1 << 20, dtype=torch.float64, device="cuda", generator=g)
    want = buf.clone()
    for _ in range(3):
        big.allreduce_async(buf.data_ptr(), 1 << 20, 1)
    torch.cuda.synchronize()
    assert torch.equal(buf, want)
    assert comm.failed_epoch() == 0 and big.failed_epoch() == 0
    with pytest.raises(ValueError):
        comm.allreduce_async(buf.data_ptr(), n + 1, 1)  # beyond the window's capacity
    comm.free()
    big.free()


def test_comm_argument_errors():
    from anml_b200 import _native as N

    with pytest.raises(ValueError):
        N.PeerComm(0, 2, 2, 16)  # rank outside the world
    with pytest.raises(ValueError):
        N.PeerComm(0, 0, 1, 0)
    c = N.PeerComm(0, 0, 2, 16)
    with pytest.raises(ValueError):
        c.connect(b"x")  # one handle per rank
    with pytest.raises(ValueError, match="not connected"):
        c.allreduce_async(1, 16, 1)
    c.free()
