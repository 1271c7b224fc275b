#!/usr/bin/env python
"""Benchmark of the likelihood-evaluation hot path (BASELINE.json north star and configs #2-#5).

    python bench.py --gpus N --steps K --warmup W                  # this repo (CUDA path), north star
    python bench.py --config c4 --gpus N ...                       # another BASELINE config
    python bench.py --impl reference --gpus N --steps K ...        # the reference's CPU route (rank 0 only)

Workloads (`--config`, all FP64, synthetic):
  northstar  binomial/expit, 1e8 rows x 64 covariates, GaussianPrior(0,1) per coefficient   [default]
  c2         the same at 1e7 rows (BASELINE config #2, one GPU)
  c3         one B-spline variable (20 knots, degree 3, 'rel_domain'), second-derivative smoothing
             prior, gaussian NLL, 5e7 rows, basis evaluated on the fly (banded evaluator)
  c4         mean (identity) / variance (exp) model over ONE 1e8 x 64 design, P = 128
  c5         binomial/expit, 2e7 rows x 1024 covariates (Hessian all-reduce of 8.4 MB)
Rows are sharded over the N GPUs (strong scaling: total rows fixed).  The data set is generated from
GLOBAL 65536-row chunks, each seeded by its chunk number, so it is the same for every N; the line
carries a `digest` (objective, max |gradient|, trace and Frobenius norm of the Hessian at a fixed
coefficient vector) that must agree across N, and a `parity_check` of the same sharded path on a small
problem against an independent torch FP64 evaluation.

One step = one full objective + gradient + Hessian evaluation at a fresh coefficient vector, including
the prior terms and (N > 1) the all-reduce of the packed result.

One JSON line on stdout (rank 0):
  value     evaluations/s, data resident in HBM, CUDA-event timed, max over ranks
  e2e       same through the public API (Model / ShardedModel.evaluate) with the coefficient vector in
            host memory and results returned to the host
  roofline  dominant kernel of the config: algorithmic flops (or bytes) per evaluation / CUDA-event kernel
            time, against cuBLAS DGEMM measured in-run (or the measured HBM copy peak)
  cpu_baseline   (N = 1) the reference's own classes (baseline/_ref) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import datetime as _dt
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261019
UNIT = "evals/s"
CHUNK = 1 << 16  # rows per global data chunk (seed = SEED + config offset + chunk number)

CONFIGS = {
    "northstar": dict(kind="logistic", n=100_000_000, p=64, prior=True, seed=1000,
                      metric="obj+grad+Hessian evals/sec at N=100M,p=64"),
    "c2": dict(kind="logistic", n=10_000_000, p=64, prior=True, seed=2000,
               metric="obj+grad+Hessian evals/sec at N=10M,p=64 (BASELINE config #2)"),
    "c3": dict(kind="spline", n=50_000_000, p=22, nk=20, degree=3, prior=True, seed=3000,
               metric="obj+grad+Hessian evals/sec, one B-spline (20 knots, degree 3), N=50M (BASELINE config #3)"),
    "c4": dict(kind="meanvar", n=100_000_000, p=64, prior=False, seed=4000,
               metric="obj+grad+Hessian evals/sec, mean/variance model, N=100M, 2 x 64 coefficients (BASELINE config #4)"),
    "c5": dict(kind="logistic", n=20_000_000, p=1024, prior=False, seed=5000,
               metric="obj+grad+Hessian evals/sec at N=20M,p=1024 (BASELINE config #5)"),
}


def workload_name(cfg):
    n, p = cfg["n"], cfg["p"]
    if cfg["kind"] == "logistic":
        return (f"binomial/expit logistic, n={n:.3g} rows x p={p} (FP64, row-major X resident in HBM)"
                + (", GaussianPrior(0,1) per coefficient" if cfg["prior"] else "") + ", one obj+grad+Hessian evaluation per step")
    if cfg["kind"] == "meanvar":
        return (f"gaussian mean (identity) / variance (exp) model over one design, n={n:.3g} rows x p={p} (P={2 * p} coefficients, "
                "FP64, X resident in HBM), one obj+grad+Hessian evaluation per step")
    return (f"gaussian NLL, one B-spline variable ({cfg['nk']} knots on the data range, degree {cfg['degree']}, {p} bases), second-derivative "
            f"GaussianPrior(0,1) at 100 points, n={n:.3g} rows, basis on the fly from x (banded evaluator), one obj+grad+Hessian evaluation per step")


def config_dict(cfg, name, world, rows_local=None):
    """The `config` object of the JSON line: the same for both arms at a given N."""
    n, p = cfg["n"], cfg["p"]
    per_gpu_bytes = {"logistic": 8.0 * (p + 1), "meanvar": 8.0 * (p + 1), "spline": 16.0}[cfg["kind"]] * n / world
    return {"workload": workload_name(cfg), "name": name, "n_rows_total": n, "rows_per_gpu": (n + world - 1) // world, "p": p,
            "coefficients": (2 * p if cfg["kind"] == "meanvar" else p),
            "parallelism": f"row-sharded x{world}, one packed all-reduce per evaluation" if world > 1 else "single GPU",
            "l2": f"inputs ({per_gpu_bytes / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush needed",
            "fresh_x_every_step": True,
            "data_generation": f"global {CHUNK}-row chunks seeded by chunk number (identical data set for every N)"}


# --------------------------------------------------------------------------------------
# host threads
# --------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        info = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if info:
            return int(max(i["num_threads"] for i in info)), info[0].get("internal_api", "blas")
    except Exception:
        pass
    return os.cpu_count() or 1, "blas"


def all_blas_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms are meant to use every host
    core, so lift the BLAS pool back to os.cpu_count() for their duration."""
    try:
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        import contextlib

        return contextlib.nullcontext()


# --------------------------------------------------------------------------------------
# CPU arm: the reference's own classes (baseline/_ref = pip install of /root/reference)
# --------------------------------------------------------------------------------------
def import_reference():
    """`anml` from baseline/_ref (the unmodified reference, installed by `pip install --target`; see
    DESIGN.md).  None when it is not there."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "anml")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import anml.parameter.main as apm  # noqa: F401  (imports variable/prior/data; never xspline)
        import anml.parameter.smoothmapping  # noqa: F401

        return sys.modules["anml"]
    except Exception:
        return None


def cpu_sample_data(cfg, n_sample, seed):
    rng = np.random.default_rng(seed)
    p = cfg["p"]
    if cfg["kind"] == "spline":
        x = rng.random(n_sample)
        obs = np.sin(2 * np.pi * x) + 0.1 * rng.standard_normal(n_sample)
        return {"x": x, "obs": obs}
    X = rng.standard_normal((n_sample, p))
    X[:, 0] = 1.0
    beta = rng.standard_normal(p) * (0.1 / np.sqrt(p / 64.0))
    if cfg["kind"] == "logistic":
        obs = (rng.random(n_sample) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
        return {"X": X, "obs": obs, "x0": 0.5 * beta}
    bs = rng.standard_normal(p) * 0.05
    obs = X @ beta + np.exp(0.5 * (X @ bs)) * rng.standard_normal(n_sample)
    return {"X": X, "obs": obs, "x0": np.concatenate([0.5 * beta, 0.5 * bs])}


class ReferenceRoute:
    """One evaluation = the reference's `Parameter.get_params(order=0)` (parameter/main.py:265-275), its
    SmoothMapping for f', f'' (smoothmapping.py:59-126), then the chunked X^T r and (X^T * w) X that
    BASELINE.md section 4.2 calls the best-effort NumPy route, and `Parameter.prior_*`
    (parameter/main.py:282-340).  (The literal order-2 route materialises an (n,p,p) tensor: timed
    separately on a few thousand rows.)"""

    kind = "reference"

    def __init__(self, cfg, data):
        import pandas as pd

        import anml.data.component as adc
        import anml.parameter.main as apm
        import anml.parameter.smoothmapping as asm
        import anml.prior.main as aprior
        import anml.variable.main as avm

        self.cfg, self.obs = cfg, data["obs"]
        X, p = data["X"], cfg["p"]
        df = pd.DataFrame({f"cov{j}": X[:, j] for j in range(1, p)}, copy=False)

        def variables():
            prior = [aprior.GaussianPrior(mean=0.0, sd=1.0)] if cfg["prior"] else None
            vs = [avm.Variable(adc.Component("intercept", default_value=1.0), priors=prior)]
            vs += [avm.Variable(f"cov{j}", priors=prior) for j in range(1, p)]
            return vs

        if cfg["kind"] == "logistic":
            self.pars = [apm.Parameter(variables(), transform=asm.Expit())]
        else:
            self.pars = [apm.Parameter(variables(), transform=asm.Identity()), apm.Parameter(variables(), transform=asm.Exp())]
        for par in self.pars:
            par.attach(df)  # design_mat = np.hstack([...]), parameter/main.py:160-162
        del df

    def evaluate(self, x, chunk=1 << 18):
        cfg, obs, p = self.cfg, self.obs, self.cfg["p"]
        K = len(self.pars)
        xs = [x[k * p:(k + 1) * p] for k in range(K)]
        theta = [self.pars[k].get_params(xs[k], order=0) for k in range(K)]
        ys = [self.pars[k].design_mat.dot(xs[k]) for k in range(K)]
        f1 = [self.pars[k].transform(ys[k], order=1) for k in range(K)]
        f2 = [self.pars[k].transform(ys[k], order=2) for k in range(K)]
        P = K * p
        grad, hess = np.zeros(P), np.zeros((P, P))
        if cfg["kind"] == "logistic":
            th = theta[0]
            obj = float(-np.sum(obs * np.log(th) + (1.0 - obs) * np.log(1.0 - th)))
            dl = -obs / th + (1.0 - obs) / (1.0 - th)
            d2l = obs / th**2 + (1.0 - obs) / (1.0 - th) ** 2
            r = [dl * f1[0]]
            w = {(0, 0): d2l * f1[0] ** 2 + dl * f2[0]}
        else:
            mu, v = theta
            res = obs - mu
            iv = 1.0 / v
            obj = float(np.sum(0.5 * np.log(v) + 0.5 * res * res * iv))
            d0 = -res * iv
            d1 = 0.5 * iv - 0.5 * res * res * iv * iv
            r = [d0 * f1[0], d1 * f1[1]]
            w = {(0, 0): iv * f1[0] ** 2 + d0 * f2[0], (0, 1): res * iv * iv * f1[0] * f1[1],
                 (1, 1): (-0.5 * iv * iv + res * res * iv**3) * f1[1] ** 2 + d1 * f2[1]}
        n = obs.shape[0]
        for k in range(K):
            Xk = self.pars[k].design_mat
            grad[k * p:(k + 1) * p] = Xk.T @ r[k]
            for m in range(k, K):
                Xm = self.pars[m].design_mat
                blk = np.zeros((p, p))
                for lo in range(0, n, chunk):
                    hi = min(n, lo + chunk)
                    blk += (Xk[lo:hi].T * w[(k, m)][lo:hi]) @ Xm[lo:hi]
                hess[k * p:(k + 1) * p, m * p:(m + 1) * p] = blk
                if m != k:
                    hess[m * p:(m + 1) * p, k * p:(k + 1) * p] = blk.T
        for k in range(K):
            sl = slice(k * p, (k + 1) * p)
            obj += float(self.pars[k].prior_objective(xs[k]))
            grad[sl] += self.pars[k].prior_gradient(xs[k])
            hess[sl, sl] += self.pars[k].prior_hessian(xs[k])
        return obj, grad, hess

    def literal_rows_per_s(self, x, n_lit):
        """The reference's literal route on n_lit rows: get_params orders 1 and 2 materialise J (n,p) and
        T (n,p,p) (parameter/main.py:276-280)."""
        import pandas as pd

        par = self.pars[0]
        p = self.cfg["p"]
        X = par.design_mat[:n_lit]
        df = pd.DataFrame({f"cov{j}": X[:, j] for j in range(1, p)})
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            par.attach(df)
            th = par.get_params(x[:p], order=0)
            J = par.get_params(x[:p], order=1)
            T = par.get_params(x[:p], order=2)
            ob = self.obs[:n_lit]
            dl = -ob / th + (1.0 - ob) / (1.0 - th)
            d2l = ob / th**2 + (1.0 - ob) / (1.0 - th) ** 2
            g = J.T @ dl
            H = (J.T * d2l) @ J + np.tensordot(dl, T, axes=1)
            best = min(best, time.perf_counter() - t0)
            del J, T, g, H
        return {"rows_per_s": n_lit / best, "rows": int(n_lit),
                "note": "reference get_params orders 1 and 2: J (n,p) and T (n,p,p) materialised (attach included); best of 3"}


class PortRoute:
    """Config #3 only: `xspline` (third-party, absent here) blocks the reference's spline classes, so the
    CPU route is the oracle port: basis by Cox-de Boor, then the chunked NumPy normal equations."""

    kind = "port"

    def __init__(self, cfg, data):
        from oracle import anml_oracle as O

        self.O, self.cfg, self.data = O, cfg, data
        x = data["x"]
        self.knots = x.min() + np.linspace(0.0, 1.0, cfg["nk"]) * (x.max() - x.min())
        pts = O.spline_prior_points(self.knots, 100, (0.0, 1.0), "rel")
        self.prior = {"linear": (np.zeros(100), np.ones(100), O.spline_design(self.knots, cfg["degree"], pts, order=2))}

    def evaluate(self, x, chunk=1 << 18):
        O, d = self.O, self.data
        B = O.spline_design(self.knots, self.cfg["degree"], d["x"])
        return O.model_eval_fused("gaussian", d["obs"], [B], ["identity"], x, obs_se=np.ones(d["obs"].size), priors=[self.prior], chunk=chunk)


def cpu_route(cfg, n_sample, seed):
    data = cpu_sample_data(cfg, n_sample, seed)
    if cfg["kind"] != "spline" and import_reference() is not None:
        route = ReferenceRoute(cfg, data)
    elif cfg["kind"] == "spline":
        route = PortRoute(cfg, data)
    else:  # baseline/_ref missing: the oracle's chunked NumPy route
        from oracle import anml_oracle as O

        class _Port:
            kind = "port"

            def evaluate(self, x, chunk=1 << 18):
                p = cfg["p"]
                pri = [{"direct": (np.zeros(p), np.ones(p))}] if cfg["prior"] else None
                if cfg["kind"] == "logistic":
                    return O.model_eval_fused("binomial", data["obs"], [data["X"]], ["expit"], x, priors=pri, chunk=chunk)
                return O.model_eval_fused("gaussian_meanvar", data["obs"], [data["X"], data["X"]], ["identity", "exp"], x, chunk=chunk)

        route = _Port()
    x0 = data.get("x0")
    if x0 is None:
        x0 = np.linspace(-1.0, 1.0, cfg["p"])
    return route, x0


def default_cpu_rows(cfg):
    # about 1 s of host work per evaluation on a 16-core box
    return {"logistic": 4_000_000 if cfg["p"] <= 64 else 60_000, "meanvar": 1_500_000, "spline": 400_000}[cfg["kind"]]


def run_reference(args, cfg):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores, all
    threads, on a bounded row sample of the same workload; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = cfg["n"]
    n_sample = min(n_total, args.cpu_rows or default_cpu_rows(cfg))
    with all_blas_threads():
        route, x0 = cpu_route(cfg, n_sample, SEED + 2)
        times = []
        for k in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            route.evaluate(x0 * (1.0 + 1e-3 * k))
            times.append(time.perf_counter() - t0)
        cores, api = cpu_threads()
    sec = float(np.mean(times[args.warmup:]))
    factor = n_total / n_sample
    value = 1.0 / (sec * factor)
    what = ("anml.parameter.main.Parameter.get_params(order=0) + SmoothMapping orders 1, 2 + chunked (X^T * w) X + Parameter.prior_* "
            "from baseline/_ref (the unmodified reference)") if route.kind == "reference" else "oracle port (NumPy restatement)"
    sample = (f"{n_sample} of {n_total} rows per step (ms_per_step is the MEASURED time of that sample; value = evals/s of the whole "
              f"workload, linear in rows, extrapolation_factor {factor:g}); {what}; NumPy {np.__version__} ({api}, {cores} threads of "
              f"{os.cpu_count()} cores); row chunks of 262144")
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "extrapolation_factor": factor, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(cfg, args.config, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": route.kind, "sample": sample,
                         "rows_per_s": n_sample / sec},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for ln in open(self.path):
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = _dt.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
            except Exception:
                continue
        try:
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or rows
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[4][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": inside[0][2],
                "power_w_max": max(r[3] for r in inside), "reasons": reasons, "samples": len(inside)}


# --------------------------------------------------------------------------------------
# synthetic data: global chunks, identical for every N
# --------------------------------------------------------------------------------------
def true_coefficients(torch, dev, cfg):
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + cfg["seed"])
    p = cfg["p"]
    beta = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * (0.1 / np.sqrt(p / 64.0))
    beta_s = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.05
    return beta, beta_s


def make_shard(torch, dev, cfg, lo, hi, seed_shift=0):
    """Rows [lo, hi) of the config's global data set.  Every global chunk of CHUNK rows is drawn, in full
    and in a fixed order, from a generator seeded with the chunk number; a shard slices the chunks it
    overlaps, so any partition of the rows sees the same numbers.
    x_ij ~ N(0,1) with an intercept column of ones, beta_true ~ N(0, 0.1^2)."""
    n, p, kind = hi - lo, cfg["p"], cfg["kind"]
    g = torch.Generator(device=dev)
    beta, beta_s = true_coefficients(torch, dev, cfg)
    out = {}
    if kind == "spline":
        xs = torch.empty(n, dtype=torch.float64, device=dev)
    else:
        X = torch.empty((n, p), dtype=torch.float64, device=dev)
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    base = SEED + cfg["seed"] + 17 + seed_shift
    for c in range(lo // CHUNK, (hi - 1) // CHUNK + 1 if hi > lo else 0):
        g.manual_seed(base + c)
        c0 = c * CHUNK
        a, b = max(lo, c0), min(hi, c0 + CHUNK)
        if kind == "spline":
            xc = torch.rand(CHUNK, dtype=torch.float64, device=dev, generator=g)
            oc = torch.sin(2 * np.pi * xc) + 0.1 * torch.randn(CHUNK, dtype=torch.float64, device=dev, generator=g)
            xs[a - lo:b - lo] = xc[a - c0:b - c0]
        else:
            Xc = torch.randn((CHUNK, p), dtype=torch.float64, device=dev, generator=g)
            Xc[:, 0] = 1.0
            if kind == "logistic":
                oc = (torch.rand(CHUNK, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(Xc @ beta)).to(torch.float64)
            else:
                oc = Xc @ beta + torch.exp(0.5 * (Xc @ beta_s)) * torch.randn(CHUNK, dtype=torch.float64, device=dev, generator=g)
            X[a - lo:b - lo] = Xc[a - c0:b - c0]
        obs[a - lo:b - lo] = oc[a - c0:b - c0]
    torch.cuda.synchronize(dev)
    out["obs"] = obs
    if kind == "spline":
        out["x"] = xs
        out["x0"] = np.linspace(-1.0, 1.0, p)
    else:
        out["X"] = X
        b0 = 0.5 * beta.cpu().numpy()
        out["x0"] = b0 if kind == "logistic" else np.concatenate([b0, 0.5 * beta_s.cpu().numpy()])
    return out


def build_model(cfg, data, local_rank, rank, world):
    """The public classes over the shard: Parameter / Variable / Prior objects as a user of anml writes
    them, design matrices adopted from device memory (the data were generated there)."""
    from anml_b200.data import Component, DataExample
    from anml_b200.model import Model
    from anml_b200.parameter import Exp, Expit, Identity, Parameter
    from anml_b200.prior import GaussianPrior
    from anml_b200.variable import Variable

    p, kind = cfg["p"], cfg["kind"]
    n = int(data["obs"].shape[0])

    def variables():
        prior = [GaussianPrior(mean=0.0, sd=1.0)] if cfg["prior"] else None
        vs = [Variable(Component("intercept", default_value=1.0), priors=prior)]
        vs += [Variable(f"cov{j}", priors=prior) for j in range(1, p)]
        return vs

    if kind == "spline":
        import pandas as pd

        from anml_b200.distributed import resolve_sharded_spline_knots
        from anml_b200.getter.prior import SplinePriorGetter
        from anml_b200.getter.spline import SplineGetter
        from anml_b200.variable import SplineVariable

        getter = SplineGetter(np.linspace(0.0, 1.0, cfg["nk"]), degree=cfg["degree"], knots_type="rel_domain")
        smooth = SplinePriorGetter(GaussianPrior(mean=0.0, sd=1.0), size=100, order=2)
        par = Parameter([SplineVariable("x", getter, priors=[smooth] if cfg["prior"] else None)], transform=Identity(), device=local_rank)
        df = pd.DataFrame({"x": data["x"].cpu().numpy(), "obs": data["obs"].cpu().numpy()})
        if world > 1:
            resolve_sharded_spline_knots([par], df)  # knots of the WHOLE data range on every rank
        model = Model("gaussian", DataExample("obs", "obs_se"), [par], device=local_rank)
        model.attach(df)
        del df
        if not model.banded_spline:
            raise SystemExit("config c3 expects the banded spline evaluator")
        pars = [par]
    else:
        X = data["X"]
        links = [Expit()] if kind == "logistic" else [Identity(), Exp()]
        pars = []
        for link in links:
            par = Parameter(variables(), transform=link, device=local_rank)
            par.attach_device(X.data_ptr(), n, ld=p, keepalive=X)
            pars.append(par)
        family = "binomial" if kind == "logistic" else "gaussian_meanvar"
        model = Model(family, DataExample("obs", "obs_se"), pars, device=local_rank)
        model.attach_device(n, data["obs"].data_ptr(), keepalive=data["obs"])
    model.set_prior_enabled(rank == 0)
    return model, pars


def torch_reference_eval(torch, cfg, data, x, pars):
    """Independent FP64 evaluation with plain torch ops (cuBLAS) of the same model on `data` (small n):
    the checker of `parity_check`.  Prior terms come from the host NumPy layer (Parameter.prior_*)."""
    xt = torch.tensor(x, dtype=torch.float64, device=data["obs"].device)
    obs, p = data["obs"], cfg["p"]
    if cfg["kind"] == "logistic":
        X = data["X"]
        y = X @ xt
        th = torch.sigmoid(y)
        obj = -(obs * torch.log(th) + (1 - obs) * torch.log1p(-th)).sum()
        grad = X.T @ (th - obs)
        hess = (X.T * (th * (1 - th))) @ X
    elif cfg["kind"] == "meanvar":
        X = data["X"]
        mu, lv = X @ xt[:p], X @ xt[p:]
        v = torch.exp(lv)
        res = obs - mu
        obj = (0.5 * lv + 0.5 * res * res / v).sum()
        r0, r1 = -res / v, 0.5 - 0.5 * res * res / v
        grad = torch.cat([X.T @ r0, X.T @ r1])
        h00, h01, h11 = (X.T / v) @ X, (X.T * (res / v)) @ X, (X.T * (0.5 * res * res / v)) @ X
        hess = torch.cat([torch.cat([h00, h01], 1), torch.cat([h01.T, h11], 1)], 0)
    else:
        B = torch.tensor(pars[0].design_mat, dtype=torch.float64, device=obs.device)  # dense basis of the same rows
        res = B @ xt - obs
        obj = 0.5 * (res * res).sum()
        grad = B.T @ res
        hess = B.T @ B
    obj, grad, hess = float(obj), grad.cpu().numpy(), hess.cpu().numpy()
    at = 0
    for par in pars:
        sl = slice(at, at + par.size)
        obj += float(par.prior_objective(x[sl]))
        grad[sl] += par.prior_gradient(x[sl])
        hess[sl, sl] += par.prior_hessian(x[sl])
        at += par.size
    return obj, grad, hess


def dgemm_peak_tflops(torch, dev, n=4096, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def ncu_traffic(kernels):
    """DRAM bytes per row of the kernels named (signature strings from the library), from the ncu --set
    full captures committed under profiles/ (profiles/r2_traffic.json, written by tools/ncu_traffic.py from
    the CSV extracts).  A kernel whose signature has no capture -- e.g. because its template arguments
    changed since -- yields None and a note: the number never goes stale silently."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception as exc:
        return None, f"profiles/r2_traffic.json unreadable: {exc!r}"
    total, src = 0.0, []
    for k in kernels:
        ent = table.get(k)
        if ent is None:
            return None, f"no ncu capture for kernel signature '{k}' in profiles/r2_traffic.json"
        total += float(ent["dram_bytes_per_row"])
        src.append(f"{k}: {ent['dram_bytes_per_row']:.2f} B/row ({ent['csv']}, n={ent['rows']:g})")
    return total, "; ".join(src)


def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist

    from anml_b200.distributed import NativeLocalEvaluator, ShardedModel, shard_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: anml_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    kind, p = cfg["kind"], cfg["p"]
    P = 2 * p if kind == "meanvar" else p

    # ---- parity check of the sharded path on a small problem (every rank takes part) ----
    n_small = 70_001  # (at least 8192 rows per rank at N = 8: the INT8 Hessian path takes part in the check)
    small_cfg = dict(cfg, n=n_small)
    slo, shi = shard_rows(n_small, world, rank)
    sdata = make_shard(torch, dev, small_cfg, slo, shi, seed_shift=777)
    smodel, spars = build_model(small_cfg, sdata, local_rank, rank, world)
    ssharded = ShardedModel(NativeLocalEvaluator(smodel, dev))
    so, sg, sh = ssharded.unpack(ssharded.evaluate_collective(sdata["x0"], True))
    parity = None
    if rank == 0:
        full = make_shard(torch, dev, small_cfg, 0, n_small, seed_shift=777)
        fpars = spars
        if kind == "spline":  # the checker needs the dense basis of ALL rows with the same knots
            import pandas as pd

            from anml_b200.parameter import Identity, Parameter
            from anml_b200.variable import SplineVariable

            fvar = SplineVariable("x", spars[0].variables[0].spline, priors=list(spars[0].variables[0].priors))
            fpar = Parameter([fvar], transform=Identity(), device=local_rank)
            fpar.attach(pd.DataFrame({"x": full["x"].cpu().numpy()}))
            fpars = [fpar]
        wo, wg, wh = torch_reference_eval(torch, small_cfg, full, sdata["x0"], fpars)
        rel = [abs(so - wo) / abs(wo), float(np.abs(sg - wg).max() / np.abs(wg).max()), float(np.abs(sh - wh).max() / np.abs(wh).max())]
        parity = {"ok": bool(rel[0] <= 1e-10 and rel[1] <= 1e-10 and rel[2] <= 1e-9 and np.array_equal(sh, sh.T)),
                  "max_rel": {"obj": rel[0], "grad": rel[1], "hess": rel[2]}, "rows": n_small, "ranks": world,
                  "against": "independent torch FP64 evaluation (cuBLAS) of the whole small data set on rank 0",
                  "tolerance": {"obj": 1e-10, "grad": 1e-10, "hess": 1e-9}}
        del full
    ssharded.close()
    del ssharded, smodel, spars, sdata
    barrier()

    # ---- the workload ----
    n_total = cfg["n"]
    lo, hi = shard_rows(n_total, world, rank)
    n = hi - lo
    data = make_shard(torch, dev, cfg, lo, hi)
    model, pars = build_model(cfg, data, local_rank, rank, world)
    local = NativeLocalEvaluator(model, dev)
    sharded = ShardedModel(local)
    x0 = data["x0"]
    if kind == "spline":
        del data["x"]  # the model keeps its own sorted copy

    steps, warmup = args.steps, max(args.warmup, 3)
    xs = [x0 * (1.0 + 1e-3 * k) for k in range(steps + warmup)]

    # digest at a fixed x: must agree across N (same data, different summation order: ~1e-13)
    do, dg, dh = sharded.unpack(sharded.evaluate_collective(x0, True))
    digest = {"obj": do, "grad_inf": float(np.abs(dg).max()), "hess_trace": float(np.trace(dh)), "hess_fro": float(np.linalg.norm(dh)),
              "at": "x0 = 0.5 * beta_true" if kind != "spline" else "x0 = linspace(-1, 1, 22)"}

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-timed steps: data resident, coefficient vector from host, result left on the device
    for k in range(warmup):
        sharded.evaluate_collective(xs[k], True)
    barrier()
    model.native.kernel_time(reset=True)
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        sharded.evaluate_collective(xs[warmup + k], True)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    k_ms, k_launches, all_launches = model.native.kernel_time(reset=True)
    main_kernel = model.native.main_kernel()
    step_kernels = model.native.last_kernels()  # the timed kernels of one evaluation, in launch order
    ms_step, kernel_ms = max_over_ranks([ms_total / steps, k_ms / steps])
    launches_per_step = k_launches / steps

    # ---- end to end through the public API: host x in, host (obj, grad, hess) out, every step.
    # Rank 0 drives (as a solver would), the other ranks serve; rank 0's wall clock spans every
    # rank's work because each evaluation ends with the all-reduce.
    barrier()
    e2e_ms, e2e_kernel_ms = 0.0, 0.0
    if rank == 0:
        for k in range(warmup):
            sharded.evaluate(xs[k] * (1.0 + 2e-6), True)
        torch.cuda.synchronize(dev)
        model.native.kernel_time(reset=True)
        t0 = time.perf_counter()
        for k in range(steps):
            sharded.evaluate(xs[warmup + k] * (1.0 + 1e-6), True)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        e2e_kernel_ms = model.native.kernel_time(reset=True)[0] / steps  # rank 0's dominant kernel(s) inside this loop
        sharded.stop()
    else:
        sharded.serve()
    barrier()
    t_wall1 = time.time()
    (e2e_ms,) = max_over_ranks([e2e_ms])

    # ---- objective + gradient only (no Hessian): the HBM-bound pass of the one-parameter fused path
    grad_kernel_ms, grad_kernel = None, None
    if kind == "logistic" and p <= 64:
        for k in range(3):
            sharded.evaluate_collective(xs[k], False)
        barrier()
        model.native.kernel_time(reset=True)
        gsteps = max(5, min(steps, 20))
        for k in range(gsteps):
            sharded.evaluate_collective(xs[warmup + (k % steps)], False)
        barrier()
        g_ms, g_launches, _ = model.native.kernel_time(reset=True)
        grad_kernel = model.native.main_kernel()
        (grad_kernel_ms,) = max_over_ranks([g_ms / max(g_launches, 1)])

    # ---- the same evaluation with the Hessian kept on the FP64 tensor cores (DMMA): the INT8 path's reference point
    dmma_kernel_ms, dmma_kernel = None, None
    if main_kernel.startswith("fused_hess_i8"):
        model.native.set_option("hess_i8", 0)
        for k in range(2):
            sharded.evaluate_collective(xs[k], True)
        barrier()
        model.native.kernel_time(reset=True)
        dsteps = max(3, min(steps, 10))
        for k in range(dsteps):
            sharded.evaluate_collective(xs[warmup + (k % steps)], True)
        barrier()
        d_ms, d_launches, _ = model.native.kernel_time(reset=True)
        dmma_kernel = model.native.main_kernel()
        (dmma_kernel_ms,) = max_over_ranks([d_ms / max(d_launches, 1)])
        model.native.set_option("hess_i8", 1)

    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- the collective alone (N > 1): the peer-memory kernel and, beside it, NCCL on the same record
    collective = None
    if world > 1:
        packed = local.packed
        collective = {"bytes": int(packed.numel() * 8), "kind": "peer-memory kernel (csrc/comm.cuh)" if sharded._peer is not None else "nccl"}

        def time_collective(fn, reps=50):
            for _ in range(5):
                fn()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(reps):
                fn()
            c1.record()
            barrier()
            return max_over_ranks([c0.elapsed_time(c1) * 1e3 / reps])[0]

        packed.zero_()
        if sharded._peer is not None:
            collective["peer_us"] = time_collective(lambda: sharded._peer.all_reduce(packed, local.stream_ptr()))
        collective["nccl_us"] = time_collective(lambda: dist.all_reduce(packed))

    # FP64 roofline denominator: cuBLAS DGEMM on this GPU, measured after the timed regions so that its
    # power draw does not leak into the clock / power samples above
    peak_tf = dgemm_peak_tflops(torch, dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dmma_peak = 37.0  # tools/microbench/fp64_bench.cu on this pool (profiles/r1_microbench_fp64.jsonl)
        n_blocks = 3 if kind == "meanvar" else 1
        flops = float(n) * p * (p + 1) * n_blocks
        if kind == "spline":
            bytes_alg = 16.0 * n
            achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
            traffic_row, traffic_src = ncu_traffic([main_kernel])
            roofline = {"kernel": main_kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": None if traffic_row is None else traffic_row * n, "kernel_ms": kernel_ms, "bytes_per_launch": bytes_alg,
                        "algorithmic": "8 B (x) + 8 B (obs) per row", "peak_source": hbm_src, "traffic_source": traffic_src}
        elif main_kernel.startswith("fused_hess_i8"):
            # Hessian on the INT8 tensor cores (csrc/fused_hess_i8.cuh): the pass is bounded by reading X once
            bytes_alg = 8.0 * n * p + 8.0 * n
            achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
            traffic_row, traffic_src = ncu_traffic([main_kernel])
            int8_macs = 65536.0 * n  # four tcgen05.mma per 32-row block: M=128 x N=128 x K=32 each
            roofline = {"kernel": main_kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": None if traffic_row is None else traffic_row * n, "kernel_ms": kernel_ms, "bytes_per_launch": bytes_alg,
                        "launches_per_step": launches_per_step,
                        "algorithmic": "8 B x (p + 1) per row: X read once, obs",
                        "peak_source": hbm_src, "traffic_source": traffic_src,
                        "hessian": "X^T diag(w) X rebuilt exactly from six int8 digit planes of sqrt(w) x (Ozaki splitting): tcgen05.mma kind::i8, "
                                   "int32 accumulators in TMEM, FP64 flush every 131040 rows",
                        "int8_tensor_tops": 2.0 * int8_macs / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0,
                        "fp64_equivalent_tflops": flops / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0}
            if dmma_kernel_ms:
                ach = flops / (dmma_kernel_ms * 1e-3) / 1e12
                roofline["dmma_path"] = {"kernel": dmma_kernel, "bound": "tensor", "kernel_ms": dmma_kernel_ms, "achieved": ach, "peak": peak_tf,
                                         "unit": "TFLOP/s", "frac": ach / peak_tf, "peak_dmma_microbench": dmma_peak,
                                         "frac_of_dmma_microbench": ach / dmma_peak,
                                         "note": "same evaluation with option hess_i8 = 0 (Hessian on the FP64 tensor cores, csrc/fused_hess.cuh), "
                                                 "SYRK convention n*p*(p+1); peak = cuBLAS DGEMM 4096^3 measured in this run"}
        else:
            achieved = flops / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
            names = step_kernels if kind == "meanvar" else [main_kernel]
            traffic_row, traffic_src = ncu_traffic(names)
            roofline = {"kernel": main_kernel if kind != "meanvar" else " + ".join(step_kernels),
                        "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                        "traffic": None if traffic_row is None else traffic_row * n, "kernel_ms": kernel_ms, "launches_per_step": launches_per_step,
                        "flops_per_launch": flops / max(launches_per_step, 1), "flops_per_step": flops,
                        "algorithmic": f"SYRK convention n*p*(p+1) per weighted block, {n_blocks} block(s)"
                                       + ("; blocks (0,0) and (1,1) on the INT8 tensor cores (exact digit planes, csrc/fused_hess_i8.cuh), the signed "
                                          "cross block on DMMA: FP64-equivalent rate" if any(k.startswith("fused_hess_i8") for k in step_kernels) else ""),
                        "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (FP64 is absent from MEASURED_PEAKS.json)",
                        "peak_dmma_microbench": dmma_peak, "frac_of_dmma_microbench": achieved / dmma_peak,
                        "traffic_source": traffic_src}
        sizes = [p, p] if kind == "meanvar" else [p]
        h2d = int(8 * (P + sum((q + 63) // 64 * 64 for q in sizes)))  # x plus the zero-padded copy per parameter
        line = {
            "metric": cfg["metric"], "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, args.config, world),
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "kernel_ms_in_this_loop": e2e_kernel_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(8 * (2 + P + P * P)),
                    "note": "Model/ShardedModel.evaluate(x): host x -> host (obj, grad, Hessian); data attached once, as the reference caches design_mat"},
            "gpu_launches": int(all_launches + (steps if (world > 1 and sharded._peer is not None) else 0)),
            "roofline": roofline, "digest": digest, "parity_check": parity, "clocks": clocks,
        }
        if collective is not None:
            line["collective"] = collective
        if grad_kernel_ms:
            bytes_alg = 8.0 * n * p + 8.0 * n
            achieved_gbs = bytes_alg / (grad_kernel_ms * 1e-3) / 1e9
            line["roofline_grad_pass"] = {"kernel": grad_kernel, "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                          "frac": achieved_gbs / hbm_peak, "kernel_ms": grad_kernel_ms, "bytes_per_launch": bytes_alg,
                                          "peak_source": hbm_src}
        if world == 1 and not args.no_cpu_baseline:
            if kind == "logistic" and p <= 64:
                # one-off ingest (Parameter.attach from a pandas frame: upload + device scatter + device
                # validators) on a bounded sample, for transparency about what `e2e` leaves outside
                try:
                    import pandas as pd

                    from anml_b200.data import Component
                    from anml_b200.parameter import Expit, Parameter
                    from anml_b200.variable import Variable

                    n_in = min(n_total, 1_000_000)
                    rng = np.random.default_rng(SEED + 9)
                    df = pd.DataFrame({f"cov{j}": rng.standard_normal(n_in) for j in range(p - 1)})
                    par_in = Parameter([Variable(Component("intercept", default_value=1.0))] + [Variable(f"cov{j}") for j in range(p - 1)],
                                       transform=Expit(), device=local_rank)
                    par_in.attach(df)  # warm-up
                    t_in = time.perf_counter()
                    par_in.attach(df)
                    t_in = time.perf_counter() - t_in
                    line["ingest"] = {"rows_per_s": n_in / t_in, "GBps": n_in * p * 8 / t_in / 1e9,
                                      "sample": f"Parameter.attach of a {n_in} x {p} pandas frame (host float64 columns)"}
                    del par_in, df
                except Exception as exc:  # never let the optional extra break the bench line
                    line["ingest"] = {"error": repr(exc)}
            n_sample = min(n_total, args.cpu_rows or default_cpu_rows(cfg))
            with all_blas_threads():
                route, cx0 = cpu_route(cfg, n_sample, SEED + 2)
                times = []
                for k in range(4):
                    t0 = time.perf_counter()
                    route.evaluate(cx0 * (1.0 + 1e-3 * k))
                    times.append(time.perf_counter() - t0)
                cores, api = cpu_threads()
                sec = float(min(times[1:]))
                line["cpu_baseline"] = {
                    "value": (n_sample / sec) / n_total, "unit": UNIT, "cores": cores, "kind": route.kind,
                    "sample": f"{n_sample} of {n_total} rows, best of 3 after a warm-up, evals/s extrapolated linearly in rows; "
                              + ("the reference's own Parameter.get_params(order=0) / SmoothMapping / prior_* (baseline/_ref) + chunked (X^T * w) X"
                                 if route.kind == "reference" else "oracle port (NumPy)") + f", {api} {cores} threads of {os.cpu_count()} cores",
                    "rows_per_s": n_sample / sec}
                if route.kind == "reference" and kind == "logistic":
                    # the reference's literal route (get_params order 2 materialises an (n,p,p) tensor,
                    # parameter/main.py:278-280) at the largest n whose tensor stays under ~1 GB
                    n_lit = max(16, min(n_sample, int(1e9 / (8.0 * p * p))))
                    line["cpu_baseline"]["literal_route"] = route.literal_rows_per_s(cx0, n_lit)
        print(json.dumps(line), flush=True)
    sharded.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="northstar", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=lambda s: int(float(s)), default=None, help="total rows over all GPUs (default: the config's)")
    ap.add_argument("--p", type=int, default=None, help="covariates (default: the config's)")
    ap.add_argument("--cpu-rows", type=lambda s: int(float(s)), default=None, help="rows of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.rows is not None:
        cfg["n"] = args.rows
    if args.p is not None and cfg["kind"] != "spline":
        cfg["p"] = args.p
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
