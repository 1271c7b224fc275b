"""Row-sharded evaluation across the GPUs of one node.

Rows are independent units and every quantity is a sum over rows, so rank ``r``
of ``W`` owns the contiguous block ``[r*n//W, (r+1)*n//W)`` of X / obs / offset
and evaluates it with the same fused kernel.  The only exchange is one sum of
the packed record ``[obj | grad(P) | hess(P*P) | flag]`` per evaluation (33 KB
at P = 64); the Gaussian prior is added on rank 0 only, before the reduce, so it
is counted once.

One process per GPU (``torchrun``); ``torch.distributed`` is the plumbing
(rendezvous, the handle exchange, barriers).  On the evaluation path itself:

* the sum runs over **peer memory**: one kernel per evaluation pushes the
  rank's record, value and flag in the same 16-byte store, into every peer's
  window over NVLink and sums what arrives in its own window in rank order
  (``csrc/comm.cuh``, :class:`PeerAllReduce`): no fence, no second round trip,
  and every rank gets the same bits.  Records above 512 KB (P > 250),
  ``ANML_B200_COLLECTIVE=nccl`` or a failed peer mapping select
  ``dist.all_reduce`` instead;
* rank 0 hands the coefficient vector to the serving ranks through a host
  shared-memory mailbox (:class:`ShmMailbox`; one node, so ``/dev/shm`` reaches
  everybody) instead of an NCCL broadcast plus a device-to-host copy per rank.

The solver runs on rank 0 (``fit``); the other ranks sit in ``serve()`` and
evaluate whatever coefficient vector rank 0 posts.

The local evaluator is injected, so the collective logic is exercised on CPU
(gloo, world_size 2) in ``tests/test_distributed_cpu.py`` with a checker
evaluator, while the product path uses :class:`NativeLocalEvaluator`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

OP_STOP, OP_EVAL_HESS, OP_EVAL_GRAD = 0.0, 1.0, 2.0
PEER_MAX_ELEMS = 1 << 16  # packed records up to 512 KB go through the peer-memory kernel, larger ones through NCCL


def shard_rows(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank`` (balanced to within one row)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    return (rank * n) // world, ((rank + 1) * n) // world


class PeerAllReduce:
    """Sum of a CUDA tensor over the ranks through mapped peer memory (``anml_comm``).

    Construction is collective: every rank exports its window, the handles travel through
    ``torch.distributed`` once, every rank maps its peers, and a barrier makes sure nobody
    launches before everybody is mapped."""

    def __init__(self, device_index: int, max_elems: int, group=None):
        import torch
        import torch.distributed as dist

        from . import _native

        self.group = group
        self.torch_float64 = torch.float64
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.comm = _native.PeerComm(device_index, rank, world, max_elems)
        dev = torch.device("cuda", device_index)
        mine = torch.frombuffer(bytearray(self.comm.export()), dtype=torch.uint8).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        handles = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            self.comm.connect(handles)
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also the barrier: everybody is mapped
        if int(ok.item()) == 0:
            self.comm.free()
            raise RuntimeError("peer-memory mapping failed on at least one rank")
        self.device = dev

    def all_reduce(self, tensor, stream_ptr: int):
        if tensor.dtype != self.torch_float64 or not tensor.is_contiguous():
            raise TypeError("peer all-reduce takes a contiguous float64 CUDA tensor")
        self.comm.allreduce_async(tensor.data_ptr(), tensor.numel(), stream_ptr)

    def check(self):
        e = self.comm.failed_epoch()
        if e:
            raise RuntimeError(f"peer all-reduce: a rank did not arrive for evaluation {e} (timeout)")

    def close(self):
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier(group=self.group)  # nobody still reads this rank's window
        self.comm.free()


class ShmMailbox:
    """[seq | op | x(P)] in host shared memory: rank 0 posts, the other ranks of the node poll.

    Construction is collective (the segment name travels through ``torch.distributed``).  The
    sequence number is written last and read first; x86's store order makes that sufficient."""

    def __init__(self, P: int, group=None):
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.rank = dist.get_rank(group)
        self.P = P
        nbytes = 8 * (2 + P)
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # the creator unlinks; keep Python's resource tracker of this process out of it
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.buf = np.ndarray((2 + P,), dtype=np.float64, buffer=self.shm.buf)
        if self.rank == 0:
            self.buf[:] = 0.0
        dist.barrier(group=group)
        self.seq = 0.0

    def post(self, op: float, x: Optional[np.ndarray]):
        self.buf[1] = op
        if x is not None:
            self.buf[2:] = x
        self.seq += 1.0
        self.buf[0] = self.seq

    def wait(self):
        import time

        want = self.seq + 1.0
        buf = self.buf
        t0 = None
        spins = 0
        while buf[0] < want:
            spins += 1
            if (spins & 1023) == 0:
                # an evaluation takes 0.1-500 ms and the next post follows it at once, so keep polling hot for a
                # long while (a sleeping poller adds its sleep quantum to EVERY rank's step: the all-reduce waits
                # for the slowest); only a solver that thinks for more than 2 s gets the core back
                now = time.perf_counter()
                if t0 is None:
                    t0 = now
                elif now - t0 > 2.0:
                    time.sleep(0.0005)
        self.seq = want
        return float(self.buf[1]), self.buf[2:].copy()

    def close(self):
        self.buf = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


class NativeLocalEvaluator:
    """Local shard evaluated by libanml_b200 into a torch CUDA tensor, on torch's
    current stream, without host synchronisation."""

    def __init__(self, model, torch_device=None):
        import torch

        self.torch = torch
        self.model = model
        self.num_coefs = model.size
        self.packed_size = model.native.packed_size()
        self.device = torch.device("cuda", model.device) if torch_device is None else torch_device
        self.packed = torch.zeros(self.packed_size, dtype=torch.float64, device=self.device)

    def eval_packed(self, x: np.ndarray, want_hessian: bool):
        from . import _native

        want = _native.WANT_OBJ | _native.WANT_GRAD | (_native.WANT_HESS if want_hessian else 0)
        # torch's default stream has handle 0, which the C-ABI reads as "the model's own
        # stream"; cudaStreamLegacy (0x1) names the legacy default stream explicitly
        self.model.native.eval_async(x, want, self.packed.data_ptr(), self.stream_ptr())
        return self.packed

    def stream_ptr(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream or 1

    def rescan_ranges(self):
        """Forget the column ranges of the INT8 Hessian path (taken again at the next evaluation)."""
        self.model.native.set_option("hess_i8_rescan", 1)


class ShardedModel:
    """objective / gradient / hessian of a row-sharded model.

    ``local`` provides ``num_coefs``, ``packed_size`` and
    ``eval_packed(x, want_hessian) -> tensor`` (length ``2 + P + P*P``).  Every
    rank must construct it with its own shard; rank 0's shard carries the prior.
    """

    def __init__(self, local, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.local = local
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.P = int(local.num_coefs)
        if local.packed_size != 2 + self.P + self.P * self.P:
            raise ValueError("packed record has the wrong length")
        self._cache_key = None
        self._cache = None
        self.num_evaluations = 0
        self._cmd = None
        self._cmd_host = None
        self._packed_host = None
        self._on_gpu = hasattr(local, "packed") and local.packed.is_cuda
        self._peer = None
        self._mailbox = None
        if self.world > 1:
            self._check_prior_placement()
            self._setup_fast_paths()

    def _check_prior_placement(self):
        """A Gaussian prior must be enabled on exactly one rank, or it is counted ``world`` times --
        silently: the fit still converges.  Checked once, collectively (models without a prior pass)."""
        model = getattr(self.local, "model", None)
        if model is None or not hasattr(model, "prior_enabled"):
            return
        torch, dist = self.torch, self.dist
        dev = self.local.packed.device if self._on_gpu else torch.device("cpu")
        has = bool(getattr(model, "has_gaussian_prior", True))
        flag = torch.tensor([1 if has else 0, 1 if (has and model.prior_enabled) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, group=self.group)
        carriers, enabled = int(flag[0].item()), int(flag[1].item())
        if carriers > 0 and enabled != 1:
            raise ValueError(f"the prior is enabled on {enabled} ranks; call model.set_prior_enabled(rank == 0) "
                             "so that it is counted exactly once")

    def _setup_fast_paths(self):
        """Peer-memory all-reduce and the shared-memory mailbox (both collective to set up).  Each falls
        back -- on every rank together -- to its torch.distributed equivalent."""
        import os

        torch, dist = self.torch, self.dist
        # the peer window serves latency-bound records; from ~0.5 MB up (P > 250) NCCL's bandwidth-optimal
        # algorithms win (measured at 8.4 MB on 2 GPUs: NCCL 44 us, one-shot peer reads 81 us)
        small = self.local.packed.numel() <= PEER_MAX_ELEMS if self._on_gpu else False
        if self._on_gpu and small and os.environ.get("ANML_B200_COLLECTIVE", "peer").lower() != "nccl":
            try:
                self._peer = PeerAllReduce(self.local.packed.device.index, self.local.packed.numel(), self.group)
            except Exception as exc:  # every rank raises together (the constructor agrees on the outcome)
                self._peer = None
                if self.rank == 0:
                    import warnings

                    warnings.warn(f"peer-memory all-reduce unavailable ({exc}); using dist.all_reduce")
        if os.environ.get("ANML_B200_MAILBOX", "shm").lower() == "shm":
            dev = self.local.packed.device if self._on_gpu else torch.device("cpu")
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            box = None
            try:
                box = ShmMailbox(self.P, self.group)
            except Exception:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 1:
                self._mailbox = box
            elif box is not None:
                box.close()

    def close(self):
        """Collective: release the peer windows and the mailbox."""
        if self._peer is not None:
            self._peer.close()
            self._peer = None
        if self._mailbox is not None:
            self._mailbox.close()
            self._mailbox = None

    # ---- collective evaluation (every rank calls this with the same x) -------------
    def evaluate_collective(self, x: np.ndarray, hessian: bool = True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.P:
            raise ValueError(f"x has {x.size} entries, the model has {self.P} coefficients")
        packed = self.local.eval_packed(x, hessian)
        if self.world > 1:
            if self._peer is not None:
                self._peer.all_reduce(packed, self.local.stream_ptr())
            else:
                self.dist.all_reduce(packed, op=self.dist.ReduceOp.SUM, group=self.group)
        return packed

    def unpack(self, packed, hessian: bool = True):
        P = self.P
        if hasattr(packed, "detach"):
            if packed.is_cuda:
                # pinned staging buffer + one stream sync instead of a pageable blocking copy
                if self._packed_host is None:
                    self._packed_host = self.torch.empty(packed.numel(), dtype=packed.dtype, pin_memory=True)
                self._packed_host.copy_(packed.detach(), non_blocking=True)
                self.torch.cuda.current_stream(packed.device).synchronize()
                host = self._packed_host.numpy()
            else:
                host = packed.detach().numpy()
        else:
            host = np.asarray(packed)
        if self._peer is not None and host[0] != host[0]:
            self._peer.check()  # NaN objective: a peer may have missed the all-reduce (timeout)
        if host[1 + P + P * P] >= 1024.0:  # (summed over ranks: the INT8 Hessian path's range flag, csrc/fused_hess_i8.cuh)
            if hasattr(self.local, "rescan_ranges"):
                self.local.rescan_ranges()
            raise RuntimeError("design values outside the column ranges recorded for the INT8 Hessian path (was a wrapped "
                               "matrix rewritten in place?); the ranges are taken again at the next evaluation")
        if host[1 + P + P * P] != 0.0:
            raise ValueError("linear predictor outside the domain of a Log/Logit mapping")
        obj = float(host[0])
        grad = host[1: 1 + P].copy()
        hess = host[1 + P: 1 + P + P * P].reshape(P, P).copy() if hessian else None
        if hess is not None and self.world > 1 and self._peer is None and self._on_gpu:
            # NCCL may add the ranks' contributions to (i, j) and (j, i) in different orders (different
            # chunks of its ring / tree): mirror the lower triangle so that the Hessian stays exactly
            # symmetric, as the peer-memory kernel (fixed rank order) and the single-GPU path deliver it
            il = np.tril_indices(P, -1)
            hess.T[il] = hess[il]
        return obj, grad, hess

    # ---- rank-0 driver / worker loop ---------------------------------------------------
    def _command(self, op: float, x: Optional[np.ndarray]):
        """Rank 0 broadcasts [op | x]; every rank returns (op, x).  Rank 0 does not read the
        broadcast back, the others do one pinned device-to-host copy."""
        torch = self.torch
        if self.world == 1:
            return op, x
        if self._mailbox is not None:
            if self.rank == 0:
                self._mailbox.post(op, x)
                return op, (None if x is None else np.array(x, dtype=np.float64, copy=True))
            return self._mailbox.wait()
        dev = self.local.packed.device if hasattr(self.local, "packed") else torch.device("cpu")
        if self._cmd is None:
            self._cmd = torch.zeros(1 + self.P, dtype=torch.float64, device=dev)
            self._cmd_host = torch.zeros(1 + self.P, dtype=torch.float64, pin_memory=self._on_gpu)
        if self.rank == 0:
            buf = self._cmd_host.numpy()
            buf[0] = op
            buf[1:] = 0.0 if x is None else x
            self._cmd.copy_(self._cmd_host, non_blocking=True)
            self.dist.broadcast(self._cmd, src=0, group=self.group)
            return op, (None if x is None else np.array(x, dtype=np.float64, copy=True))
        self.dist.broadcast(self._cmd, src=0, group=self.group)
        self._cmd_host.copy_(self._cmd, non_blocking=True)
        if self._on_gpu:
            torch.cuda.current_stream(dev).synchronize()
        host = self._cmd_host.numpy()
        return float(host[0]), host[1:].copy()

    def evaluate(self, x, hessian: bool = True):
        """Rank 0: broadcast ``x``, evaluate everywhere, return (obj, grad, hess)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        key = x.tobytes()
        if self._cache_key is not None and self._cache_key[0] == key and (self._cache_key[1] or not hessian):
            return self._cache
        _, xb = self._command(OP_EVAL_HESS if hessian else OP_EVAL_GRAD, x)
        out = self.unpack(self.evaluate_collective(xb, hessian), hessian)
        self._cache_key, self._cache = (key, hessian), out
        self.num_evaluations += 1
        return out

    # trust-constr asks for objective, gradient and Hessian at the same x back to back, so by default the
    # first call pays for the full evaluation and the other two hit the cache.  With eager_hessian = False
    # objective / gradient take the gradient-only (HBM-bound) pass unless a Hessian evaluation is cached.
    eager_hessian = True

    def objective(self, x) -> float:
        return self.evaluate(x, hessian=self.eager_hessian)[0]

    def gradient(self, x) -> np.ndarray:
        return self.evaluate(x, hessian=self.eager_hessian)[1].copy()

    def hessian(self, x) -> np.ndarray:
        return self.evaluate(x, hessian=True)[2].copy()

    def serve(self) -> int:
        """Ranks != 0: evaluate on request until rank 0 calls ``stop()``."""
        served = 0
        while True:
            op, xb = self._command(OP_STOP, None)
            if op == OP_STOP:
                return served
            self.evaluate_collective(xb, op == OP_EVAL_HESS)
            served += 1

    def stop(self) -> None:
        if self.rank == 0 and self.world > 1:
            self._command(OP_STOP, None)

    def fit(self, x0=None, options=None, bounds=None, constraints=()):
        """SciPy trust-constr on rank 0 while the other ranks serve; returns the
        result on rank 0 and ``None`` elsewhere."""
        if self.rank != 0:
            self.serve()
            return None
        from scipy.optimize import minimize

        x0 = np.zeros(self.P) if x0 is None else np.asarray(x0, dtype=np.float64)
        opts = {"gtol": 1e-10, "xtol": 1e-14, "maxiter": 200}
        opts.update(options or {})
        kwargs = {}
        if bounds is not None:
            kwargs["bounds"] = bounds
        if constraints:
            kwargs["constraints"] = constraints
        try:
            return minimize(self.objective, x0, jac=self.gradient, hess=self.hessian, method="trust-constr", options=opts, **kwargs)
        finally:
            self.stop()


# ---- data-dependent spline knots on sharded data -------------------------------------------
def resolve_sharded_spline_knots(parameters, df, group=None) -> int:
    """Give every rank the knots the *whole* data set would produce.

    A ``SplineGetter`` with ``knots_type='rel_domain'`` / ``'rel_freq'`` places its knots from
    the column it is attached to (getter/spline.py:92-99).  On row shards that would differ from
    rank to rank, and the per-rank Hessians would not be blocks of one model.  Call this on every
    rank *before* ``attach`` with the rank's own frame: ``rel_domain`` knots come from an
    all-reduce(MIN/MAX) of the column range, ``rel_freq`` knots from the quantiles of the column
    gathered on rank 0 (8 bytes per row, once) and broadcast; the getters are replaced in place by
    the resolved ``XSpline`` (as the first ``attach`` would do, variable/spline.py:103-107).
    Returns the number of variables resolved.  Without an initialised process group it resolves
    against the local frame only.
    """
    import torch
    import torch.distributed as dist

    from .data.frame import as_frame
    from .getter.spline import SplineGetter
    from .variable.spline import SplineVariable
    from .xspline import XSpline

    df = as_frame(df)
    live = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if live else 1
    rank = dist.get_rank(group) if live else 0
    on_gpu = live and dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")

    def quantiles(values: np.ndarray, q) -> np.ndarray:
        if on_gpu and values.size >= (1 << 16):
            from . import _native

            return _native.vector_quantile(q, values, device=dev.index)
        return np.quantile(values, q)

    resolved = 0
    for par in parameters:
        for var in par.variables:
            if not isinstance(var, SplineVariable) or not isinstance(var.spline, SplineGetter):
                continue
            getter = var.spline
            if getter.knots_type == "abs":
                continue
            var.component.attach(df)
            col = np.ascontiguousarray(var.component.value, dtype=np.float64)
            if getter.knots_type == "rel_domain":
                # empty shards must not win the reduction
                lo = col.min() if col.size else np.inf
                hi = col.max() if col.size else -np.inf
                ends = torch.tensor([lo, -hi], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(ends, op=dist.ReduceOp.MIN, group=group)
                lo, hi = np.float64(ends[0].item()), np.float64(-ends[1].item())
                knots = lo + np.asarray(getter.knots) * (hi - lo)
            else:  # rel_freq: exact quantiles of the union of the shards
                if world > 1:
                    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
                    sizes[rank] = col.size
                    dist.all_reduce(sizes, group=group)
                    longest = int(sizes.max().item())
                    mine = torch.full((longest,), float("nan"), dtype=torch.float64, device=dev)
                    mine[: col.size] = torch.from_numpy(col).to(dev)
                    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
                    dist.gather(mine, parts, dst=0, group=group)
                    kn = torch.empty(len(getter.knots), dtype=torch.float64, device=dev)
                    if rank == 0:
                        full = np.concatenate([parts[r][: int(sizes[r].item())].cpu().numpy() for r in range(world)])
                        kn.copy_(torch.from_numpy(np.asarray(quantiles(full, getter.knots), dtype=np.float64)))
                    dist.broadcast(kn, src=0, group=group)
                    knots = kn.cpu().numpy()
                else:
                    knots = quantiles(col, getter.knots)
            var.spline = XSpline(knots, getter.degree, ldegree=getter.ldegree, rdegree=getter.rdegree)
            resolved += 1
    return resolved
