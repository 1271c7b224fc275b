"""Worker for tests/test_gpu_multi.py: one rank of a row-sharded evaluation + fit over NCCL.

Launched by ``python -m torch.distributed.run``; every rank rebuilds the same seeded
host data set, keeps its own row block on its GPU and joins the ShardedModel.  Rank 0
checks the all-reduced result against the NumPy oracle on the full data set and the
fitted coefficients against a trust-constr run on the oracle callbacks.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import pandas as pd
    import torch
    import torch.distributed as dist

    from anml_b200.data import Component, DataExample
    from anml_b200.distributed import NativeLocalEvaluator, ShardedModel, shard_rows
    from anml_b200.model import Model
    from anml_b200.parameter import Expit, Parameter
    from anml_b200.prior import GaussianPrior
    from anml_b200.variable import Variable
    from oracle import anml_oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        rng = np.random.default_rng(20261019)
        n, p = 200_003, 24
        X = rng.normal(size=(n, p))
        X[:, 0] = 1.0
        beta = rng.normal(size=p) * 0.2
        obs = (rng.uniform(size=n) < 1 / (1 + np.exp(-X @ beta))).astype(float)
        lo, hi = shard_rows(n, world, rank)
        df = pd.DataFrame({f"c{j}": X[lo:hi, j] for j in range(1, p)})
        df["obs"] = obs[lo:hi]

        prior = [GaussianPrior(mean=0.0, sd=2.0)]
        variables = [Variable(Component("intercept", default_value=1.0), priors=prior)]
        variables += [Variable(f"c{j}", priors=prior) for j in range(1, p)]
        par = Parameter(variables, transform=Expit(), device=local_rank)
        model = Model("binomial", DataExample("obs", "obs_se"), [par], device=local_rank)
        model.attach(df)
        model.set_prior_enabled(rank == 0)  # the prior is counted once
        sharded = ShardedModel(NativeLocalEvaluator(model, torch.device("cuda", local_rank)))
        # the product path: one peer-memory kernel per evaluation (csrc/comm.cuh), x through the host mailbox
        assert sharded._peer is not None, "peer-memory all-reduce was not set up"
        assert sharded._mailbox is not None

        priors = [{"direct": (np.zeros(p), np.full(p, 2.0))}]
        x = 0.5 * beta
        o, g, h = sharded.unpack(sharded.evaluate_collective(x, True))
        # the same evaluation with NCCL doing the sum: same numbers to summation order
        os.environ["ANML_B200_COLLECTIVE"] = "nccl"
        via_nccl = ShardedModel(NativeLocalEvaluator(model, torch.device("cuda", local_rank)))
        os.environ["ANML_B200_COLLECTIVE"] = "peer"
        assert via_nccl._peer is None
        o2, g2, h2 = via_nccl.unpack(via_nccl.evaluate_collective(x, True))
        assert abs(o2 - o) <= 1e-13 * abs(o) and np.abs(g2 - g).max() <= 1e-13 * np.abs(g).max()
        assert np.abs(h2 - h).max() <= 1e-13 * np.abs(h).max()
        via_nccl.close()
        # many evaluations back to back (the two data buffers of the window alternate), gradient-only too
        for k in range(40):
            ok_, gk, hk = sharded.unpack(sharded.evaluate_collective(x * (1.0 + 1e-3 * k), k % 3 != 0), k % 3 != 0)
            wk = O.model_eval_fused("binomial", obs, [X], ["expit"], x * (1.0 + 1e-3 * k), priors=priors)
            assert abs(ok_ - wk[0]) <= 1e-10 * abs(wk[0]) and np.abs(gk - wk[1]).max() <= 1e-10 * np.abs(wk[1]).max()
            if hk is not None:
                assert np.abs(hk - wk[2]).max() <= 1e-9 * np.abs(wk[2]).max()
        want = O.model_eval_fused("binomial", obs, [X], ["expit"], x, priors=priors)
        assert abs(o - want[0]) <= 1e-10 * abs(want[0]), (o, want[0])
        assert np.abs(g - want[1]).max() <= 1e-10 * np.abs(want[1]).max()
        assert np.abs(h - want[2]).max() <= 1e-9 * np.abs(want[2]).max()
        assert np.array_equal(h, h.T)
        # every rank holds the same bits after the all-reduce
        mine = torch.tensor(np.concatenate([[o], g, h.ravel()]), device=f"cuda:{local_rank}")
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        assert torch.equal(mine, ref0)

        # ---- a sharded one-spline model: knots of the whole data set on every rank, banded evaluator per shard
        from anml_b200.distributed import resolve_sharded_spline_knots
        from anml_b200.getter.spline import SplineGetter
        from anml_b200.parameter import Identity
        from anml_b200.variable import SplineVariable

        xs_full = rng.gamma(2.0, 1.0, size=n)
        ys_full = np.sin(xs_full) + 0.1 * rng.normal(size=n)
        for kt in ("rel_domain", "rel_freq"):
            fractions = np.linspace(0.0, 1.0, 9)
            svar = SplineVariable("xs", SplineGetter(fractions, degree=3, knots_type=kt))
            spar = Parameter([svar], transform=Identity(), device=local_rank)
            sdf = pd.DataFrame({"xs": xs_full[lo:hi], "obs": ys_full[lo:hi]})
            assert resolve_sharded_spline_knots([spar], sdf) == 1
            want_knots = (xs_full.min() + fractions * (xs_full.max() - xs_full.min())) if kt == "rel_domain" else np.quantile(xs_full, fractions)
            assert np.array_equal(svar.spline.knots, want_knots)
            smodel = Model("gaussian", DataExample("obs", "obs_se"), [spar], device=local_rank)
            smodel.attach(sdf)
            smodel.set_prior_enabled(rank == 0)
            assert smodel.banded_spline is True
            assert not spar.device_design.materialized  # the dense design of a banded model is never built
            ssh = ShardedModel(NativeLocalEvaluator(smodel, torch.device("cuda", local_rank)))
            bx = np.linspace(-1.0, 1.0, spar.size)
            so, sg, sh = ssh.unpack(ssh.evaluate_collective(bx, True))
            B = O.spline_design(want_knots, 3, xs_full)
            sw = O.model_eval_fused("gaussian", ys_full, [B], ["identity"], bx, obs_se=np.ones(n))
            assert abs(so - sw[0]) <= 1e-10 * abs(sw[0])
            assert np.abs(sg - sw[1]).max() <= 1e-10 * np.abs(sw[1]).max()
            assert np.abs(sh - sw[2]).max() <= 1e-9 * np.abs(sw[2]).max()
            ssh.close()

        # ---- a wide design (P = 260: the packed record exceeds the peer window's 512 KB limit) goes through NCCL;
        # the Hessian must come back exactly symmetric all the same
        pw = 260
        Xw = rng.normal(size=(6000, pw)) / np.sqrt(pw / 8.0)
        Xw[:, 0] = 1.0
        bw = rng.normal(size=pw) * 0.2
        ow = (rng.uniform(size=6000) < 1 / (1 + np.exp(-Xw @ bw))).astype(float)
        wlo, whi = shard_rows(6000, world, rank)
        wdf = pd.DataFrame({f"c{j}": Xw[wlo:whi, j] for j in range(1, pw)})
        wdf["obs"] = ow[wlo:whi]
        wvars = [Variable(Component("intercept", default_value=1.0))] + [Variable(f"c{j}") for j in range(1, pw)]
        wmodel = Model("binomial", DataExample("obs", "obs_se"), [Parameter(wvars, transform=Expit(), device=local_rank)], device=local_rank)
        wmodel.attach(wdf)
        wsh = ShardedModel(NativeLocalEvaluator(wmodel, torch.device("cuda", local_rank)))
        assert wsh._peer is None  # 2 + P + P^2 = 67862 doubles > 65536
        wo, wg, wh = wsh.unpack(wsh.evaluate_collective(0.5 * bw, True))
        ww = O.model_eval_fused("binomial", ow, [Xw], ["expit"], 0.5 * bw)
        assert abs(wo - ww[0]) <= 1e-10 * abs(ww[0]) and np.abs(wg - ww[1]).max() <= 1e-10 * np.abs(ww[1]).max()
        assert np.abs(wh - ww[2]).max() <= 1e-9 * np.abs(ww[2]).max() and np.array_equal(wh, wh.T)
        wsh.close()

        res = sharded.fit(np.zeros(p), options={"gtol": 1e-10})
        if rank == 0:
            from scipy.optimize import minimize

            f = lambda z: O.model_eval_fused("binomial", obs, [X], ["expit"], z, priors=priors)
            ref = minimize(lambda z: f(z)[0], np.zeros(p), jac=lambda z: f(z)[1], hess=lambda z: f(z)[2], method="trust-constr",
                           options={"gtol": 1e-10, "xtol": 1e-14, "maxiter": 200})
            err = np.abs(res.x - ref.x).max()
            assert err < 1e-8, err
            print(f"NCCL_SHARDED_OK world={world} evals={sharded.num_evaluations} coef_err={err:.2e}", flush=True)
        else:
            assert res is None
        sharded.close()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
