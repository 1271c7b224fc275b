"""Row-sharding + all-reduce logic on CPU: world_size 2 over gloo, with the
oracle injected as the local evaluator (the product path uses the CUDA
evaluator; this exercises partitioning, prior-once, broadcast/serve and fit)."""
import os
import socket

import numpy as np
import pytest

from anml_b200.distributed import ShardedModel, shard_rows
from oracle import anml_oracle as O


def test_shard_rows_partition():
    for n, w in [(10, 3), (100_000_000, 8), (7, 8), (0, 2)]:
        blocks = [shard_rows(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


class OracleShard:
    """Checker evaluator: the NumPy oracle on one row block."""

    def __init__(self, X, obs, prior_sd, with_prior):
        import torch

        self.torch = torch
        self.X, self.obs = X, obs
        self.num_coefs = X.shape[1]
        self.packed_size = 2 + self.num_coefs + self.num_coefs**2
        self.priors = [{"direct": (np.zeros(self.num_coefs), np.full(self.num_coefs, prior_sd))}] if with_prior else None
        from types import SimpleNamespace

        self.model = SimpleNamespace(prior_enabled=with_prior)  # what ShardedModel's prior-placement check reads

    def eval_packed(self, x, want_hessian):
        o, g, h = O.model_eval_fused("binomial", self.obs, [self.X], ["expit"], x, priors=self.priors)
        return self.torch.from_numpy(np.concatenate([[o], g, h.ravel(), [0.0]]))


def _problem():
    rng = np.random.default_rng(11)
    n, p = 4001, 5
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    beta = rng.normal(size=p) * 0.3
    obs = (rng.uniform(size=n) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    return X, obs, beta


def _worker(rank, world, port, out, mailbox="shm"):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["ANML_B200_MAILBOX"] = mailbox
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, obs, beta = _problem()
        lo, hi = shard_rows(X.shape[0], world, rank)
        # a prior enabled on every rank would be counted world times: refused at construction
        try:
            ShardedModel(OracleShard(X[lo:hi], obs[lo:hi], 2.0, with_prior=True))
            raise AssertionError("prior on every rank was accepted")
        except ValueError as e:
            assert "exactly once" in str(e)
        model = ShardedModel(OracleShard(X[lo:hi], obs[lo:hi], 2.0, with_prior=(rank == 0)))
        assert (model._mailbox is not None) == (mailbox == "shm")  # host shared-memory command channel
        assert model._peer is None                                  # peer memory needs GPUs
        x = 0.5 * beta
        # collective call: every rank gets the full-sum result
        o, g, h = model.unpack(model.evaluate_collective(x, True))
        want = O.model_eval_fused("binomial", obs, [X], ["expit"], x, priors=[{"direct": (np.zeros(5), np.full(5, 2.0))}])
        assert abs(o - want[0]) <= 1e-12 * abs(want[0])
        assert np.abs(g - want[1]).max() <= 1e-12 * np.abs(want[1]).max()
        assert np.abs(h - want[2]).max() <= 1e-12 * np.abs(want[2]).max()
        # driver / worker protocol: rank 0 fits, the others serve
        res = model.fit(np.zeros(5), options={"gtol": 1e-9})
        if rank == 0:
            single = ShardedModel.__new__(ShardedModel)  # noqa: F841  (no process group needed for the check below)
            from scipy.optimize import minimize

            f = lambda z: O.model_eval_fused("binomial", obs, [X], ["expit"], z, priors=[{"direct": (np.zeros(5), np.full(5, 2.0))}])
            ref = minimize(lambda z: f(z)[0], np.zeros(5), jac=lambda z: f(z)[1], hess=lambda z: f(z)[2], method="trust-constr",
                           options={"gtol": 1e-9, "xtol": 1e-14})
            assert np.abs(res.x - ref.x).max() < 1e-8
            out.put(("ok", model.num_evaluations))
        else:
            assert res is None
            out.put(("ok", 0))
        model.close()
    except Exception as e:  # pragma: no cover
        out.put(("fail", repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mailbox", ["shm", "broadcast"])
def test_two_rank_gloo_sharded_fit(mailbox):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, mailbox)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in results), results
    assert max(r[1] for r in results) > 3


def _knots_worker(rank, world, port, out):
    import pandas as pd
    import torch.distributed as dist

    from anml_b200.distributed import resolve_sharded_spline_knots
    from anml_b200.getter.spline import SplineGetter
    from anml_b200.parameter import Parameter
    from anml_b200.variable import SplineVariable
    from anml_b200.xspline import XSpline

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(31)
        full = np.round(rng.gamma(2.0, 1.0, size=10_001), 3)
        lo, hi = shard_rows(full.size, world, rank)
        df = pd.DataFrame({"x": full[lo:hi], "z": -full[lo:hi]})
        fractions = np.linspace(0.0, 1.0, 7)
        va = SplineVariable("x", SplineGetter(fractions, degree=3, knots_type="rel_domain"))
        vb = SplineVariable("z", SplineGetter(fractions, degree=2, knots_type="rel_freq"))
        vc = SplineVariable("x", SplineGetter(np.array([0.0, 1.0, 20.0]), degree=1, knots_type="abs"))
        n = resolve_sharded_spline_knots([Parameter([va, vb]), Parameter([vc])], df)
        assert n == 2 and isinstance(va.spline, XSpline) and isinstance(vb.spline, XSpline)
        assert isinstance(vc.spline, SplineGetter)  # absolute knots need no data
        want_a = full.min() + fractions * (full.max() - full.min())
        want_b = np.quantile(-full, fractions)
        assert np.array_equal(va.spline.knots, want_a), (va.spline.knots, want_a)
        assert np.array_equal(vb.spline.knots, want_b), (vb.spline.knots, want_b)
        assert vb.spline.degree == 2
        out.put(("ok", rank))
    except Exception as e:  # pragma: no cover
        out.put(("fail", repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_spline_knots_are_those_of_the_whole_data_set():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_knots_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in results), results


def test_range_flag_of_the_int8_hessian_path_is_reported_and_triggers_a_rescan():
    """Flag values >= 1024 in the packed record come from csrc/fused_hess_i8.cuh (a value outside the recorded column
    ranges); the sharded model turns them into an error -- never a silently wrong Hessian -- and asks the local
    evaluator to take the ranges again.  Smaller non-zero flags stay the Log/Logit domain error."""
    import torch

    P = 3

    class FlagShard:
        def __init__(self, flag):
            self.num_coefs, self.packed_size, self.flag, self.rescans = P, 2 + P + P * P, flag, 0

        def eval_packed(self, x, want_hessian):
            rec = torch.zeros(self.packed_size, dtype=torch.float64)
            rec[0] = 1.0
            rec[-1] = self.flag
            return rec

        def rescan_ranges(self):
            self.rescans += 1

    ok = FlagShard(0.0)
    obj, grad, hess = ShardedModel(ok).evaluate(np.zeros(P), True)
    assert obj == 1.0 and hess.shape == (P, P) and ok.rescans == 0
    ranged = FlagShard(1024.0)
    with pytest.raises(RuntimeError, match="INT8 Hessian path"):
        ShardedModel(ranged).evaluate(np.zeros(P), True)
    assert ranged.rescans == 1
    with pytest.raises(ValueError, match="domain"):
        ShardedModel(FlagShard(1.0)).evaluate(np.zeros(P), True)
