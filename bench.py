#!/usr/bin/env python
"""Benchmark of the likelihood-evaluation hot path (BASELINE.json north star).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (rank 0 only)

Workload (config.workload): binomial / expit logistic model, n = 1e8 rows x
p = 64 covariates in total (rows sharded over the N GPUs: strong scaling),
GaussianPrior(0, 1) on every coefficient.  One step = one full
objective + gradient + Hessian evaluation at a fresh coefficient vector,
including the prior and (N > 1) the NCCL all-reduce of the packed result.

One JSON line on stdout (rank 0):
  value     evaluations/s, X / obs resident in HBM, CUDA-event timed, max over ranks
  e2e       same through the public API (Model / ShardedModel.evaluate) with the
            coefficient vector in host memory and results returned to the host
  roofline  dominant kernel (fused_hess_kernel): useful SYRK flops n*P*(P+1) per
            launch / CUDA-event kernel time, against cuBLAS DGEMM measured in-run
  roofline_grad_pass  objective+gradient-only evaluation against the HBM peak
  cpu_baseline        the NumPy oracle route timed on this box's host cores
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
N_TOTAL = 100_000_000
P = 64
METRIC = "obj+grad+Hessian evals/sec at N=100M,p=64"
UNIT = "evals/s"
# dram bytes per row of the fused kernel from the ncu --set full capture in profiles/ (n = 1e7: 5.2136 GB, profiles/r1_fused_hess_kernel_ncu.csv)
NCU_DRAM_BYTES_PER_ROW = 521.20


def workload_name(n_total, p):
    return (f"binomial/expit logistic, n={n_total:.3g} rows x p={p} (FP64, row-major X resident in HBM), "
            f"GaussianPrior(0,1) per coefficient, one obj+grad+Hessian evaluation per step")


# --------------------------------------------------------------------------------------
# CPU arm: the oracle's fused NumPy route
# --------------------------------------------------------------------------------------
def cpu_sample(n_sample, p, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n_sample, p))
    X[:, 0] = 1.0
    beta_true = rng.standard_normal(p) * 0.1
    obs = (rng.random(n_sample) < 1.0 / (1.0 + np.exp(-(X @ beta_true)))).astype(np.float64)
    return X, obs, beta_true


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


def cpu_eval_seconds(X, obs, x, p, reps, warmup=1):
    from oracle import anml_oracle as O

    priors = [{"direct": (np.zeros(p), np.ones(p))}]
    times = []
    for k in range(warmup + reps):
        t0 = time.perf_counter()
        O.model_eval_fused("binomial", obs, [X], ["expit"], x * (1.0 + 1e-3 * k), priors=priors, chunk=1 << 18)
        times.append(time.perf_counter() - t0)
    return times[warmup:]


def cpu_literal_rows_per_s(X, obs, x, p):
    from oracle import anml_oracle as O

    priors = [{"direct": (np.zeros(p), np.ones(p))}]
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        O.model_eval_literal("binomial", obs, [X], ["expit"], x, priors=priors)
        best = min(best, time.perf_counter() - t0)
    return {"rows_per_s": X.shape[0] / best, "rows": int(X.shape[0]),
            "note": "J (n,p) and T (n,p,p) materialised as the reference's get_params orders 1 and 2 do; best of 3"}


def run_reference(args):
    """`--impl reference`: the reference's CPU route for this path, all host threads.
    anml is pure NumPy and has no model layer; the timed code is the oracle port of
    the get_params/transform/prior arithmetic in its chunked X^T diag(w) X form
    (the literal (n,p,p) route cannot hold even 1e6 rows)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total, p = args.rows, args.p
    n_sample = min(n_total, args.cpu_rows)
    X, obs, beta_true = cpu_sample(n_sample, p, SEED + 2)
    with all_blas_threads():
        times = cpu_eval_seconds(X, obs, 0.5 * beta_true, p, reps=args.steps, warmup=args.warmup)
        cores, api = cpu_threads()
    sec = float(np.mean(times))
    value = (n_sample / sec) / n_total
    sample = (f"{n_sample} of {n_total} rows per step, evals/s extrapolated linearly in rows; NumPy {np.__version__} "
              f"({api}, {cores} threads of {os.cpu_count()} cores); row chunks of 262144")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3 * (n_total / n_sample), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_total, p), "n_rows_total": n_total, "p": p},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
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
# GPU arm
# --------------------------------------------------------------------------------------
def make_shard(torch, dev, n, p, rank):
    """Synthetic shard on the device: x_ij ~ N(0,1), intercept column of ones,
    obs ~ Bernoulli(expit(X beta_true)), beta_true ~ N(0, 0.1^2)."""
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    beta_true = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.1  # same on every rank
    g.manual_seed(SEED + 1000 + rank)
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 1 << 22
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g)
        X[lo:hi, 0] = 1.0
        prob = torch.sigmoid(X[lo:hi] @ beta_true)
        obs[lo:hi] = (torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) < prob).to(torch.float64)
    torch.cuda.synchronize(dev)
    return X, obs, beta_true.cpu().numpy()


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


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from anml_b200 import _native
    from anml_b200.data import Component, DataExample
    from anml_b200.distributed import NativeLocalEvaluator, ShardedModel, shard_rows
    from anml_b200.model import Model
    from anml_b200.parameter import Expit, Parameter
    from anml_b200.prior import GaussianPrior
    from anml_b200.variable import Variable

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
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

    n_total, p = args.rows, args.p
    lo, hi = shard_rows(n_total, world, rank)
    n = hi - lo
    X, obs, beta_true = make_shard(torch, dev, n, p, rank)

    prior = [GaussianPrior(mean=0.0, sd=1.0)]
    variables = [Variable(Component("intercept", default_value=1.0), priors=prior)]
    variables += [Variable(f"cov{j}", priors=prior) for j in range(p - 1)]
    par = Parameter(variables, transform=Expit(), device=local_rank)
    par.attach_device(X.data_ptr(), n, ld=p, keepalive=X)
    model = Model("binomial", DataExample("obs", "obs_se"), [par], device=local_rank)
    model.attach_device(n, obs.data_ptr(), keepalive=obs)
    model.set_prior_enabled(rank == 0)
    local = NativeLocalEvaluator(model, dev)
    sharded = ShardedModel(local)

    steps, warmup = args.steps, max(args.warmup, 3)
    xs = [0.5 * beta_true * (1.0 + 1e-3 * k) for k in range(steps + warmup)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-timed steps: X resident, coefficient vector from host, result left on the device
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
    tms = torch.tensor([ms_total, k_ms / max(k_launches, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms[0]) / steps
    kernel_ms = float(tms[1])

    # ---- end to end through the public API: host x in, host (obj, grad, hess) out, every step.
    # Rank 0 drives (as a solver would), the other ranks serve; rank 0's wall clock spans every
    # rank's work because each evaluation ends with the all-reduce.
    barrier()
    e2e_ms = 0.0
    if rank == 0:
        for k in range(warmup):
            sharded.evaluate(xs[k] * (1.0 + 2e-6), True)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(steps):
            sharded.evaluate(xs[warmup + k] * (1.0 + 1e-6), True)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        sharded.stop()
    else:
        sharded.serve()
    barrier()
    t_wall1 = time.time()
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])

    # ---- gradient pass alone (objective + gradient, no Hessian): the HBM-bound kernel
    for k in range(3):
        sharded.evaluate_collective(xs[k], False)
    barrier()
    model.native.kernel_time(reset=True)
    gsteps = max(5, min(steps, 20))
    for k in range(gsteps):
        sharded.evaluate_collective(xs[warmup + (k % steps)], False)
    barrier()
    g_ms, g_launches, _ = model.native.kernel_time(reset=True)
    tg = torch.tensor([g_ms / max(g_launches, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    grad_kernel_ms = float(tg[0])

    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
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
        flops = float(n) * p * (p + 1)
        bytes_alg = 8.0 * n * p + 8.0 * n
        achieved_tf = flops / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        # (the general path, p > 64, has no single timed gradient kernel)
        achieved_gbs = bytes_alg / (grad_kernel_ms * 1e-3) / 1e9 if grad_kernel_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n_total, p), "n_rows_total": n_total, "rows_per_gpu": n, "p": p,
                       "parallelism": f"row-sharded x{world}, one packed all-reduce per evaluation" if world > 1 else "single GPU",
                       "l2": f"inputs ({bytes_alg / 1e9:.1f} GB per GPU) exceed the 126 MB L2; no flush needed",
                       "fresh_x_every_step": True},
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * (p + 64)), "d2h_bytes_per_step": int(8 * (2 + p + p * p)),
                    "note": "Model/ShardedModel.evaluate(x): host x -> host (obj, grad, Hessian); X attached once, as the reference caches design_mat"},
            "gpu_launches": int(all_launches),
            "roofline": {"kernel": "fused_hess_kernel" if p <= 64 else "wide_hessian_kernel", "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": NCU_DRAM_BYTES_PER_ROW * n if p == 64 else None,
                         "kernel_ms": kernel_ms, "flops_per_launch": flops,
                         "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (FP64 is absent from MEASURED_PEAKS.json)",
                         "traffic_source": "ncu --set full capture at n=1e7 (profiles/), dram bytes per row x rows"},
            "roofline_grad_pass": {"kernel": "fused_eval_kernel<HESS=false>", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak,
                                   "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "kernel_ms": grad_kernel_ms,
                                   "bytes_per_launch": bytes_alg, "peak_source": hbm_src},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            # one-off ingest (Parameter.attach from a pandas frame: upload + device scatter + device
            # validators) on a bounded sample, for transparency about what `e2e` leaves outside
            try:
                import pandas as pd

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
            n_sample = min(n_total, args.cpu_rows)
            Xc, oc, bt = cpu_sample(n_sample, p, SEED + 2)
            with all_blas_threads():
                times = cpu_eval_seconds(Xc, oc, 0.5 * bt, p, reps=3, warmup=1)
                cores, api = cpu_threads()
            sec = float(min(times))
            line["cpu_baseline"] = {
                "value": (n_sample / sec) / n_total, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_sample} of {n_total} rows, best of 3, evals/s extrapolated linearly in rows; NumPy oracle route "
                          f"(chunked X^T diag(w) X), {api} {cores} threads of {os.cpu_count()} cores",
                "rows_per_s": n_sample / sec}
            # the reference's literal route (get_params order 2 materialises an (n,p,p) tensor,
            # parameter/main.py:278-280) at the largest n whose tensor stays under ~1 GB
            n_lit = max(16, min(n_sample, int(1e9 / (8.0 * p * p))))
            with all_blas_threads():
                line["cpu_baseline"]["literal_route"] = cpu_literal_rows_per_s(Xc[:n_lit], oc[:n_lit], 0.5 * bt, p)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=lambda s: int(float(s)), default=N_TOTAL, help="total rows over all GPUs")
    ap.add_argument("--p", type=int, default=P)
    ap.add_argument("--cpu-rows", type=lambda s: int(float(s)), default=8_000_000, help="rows of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
