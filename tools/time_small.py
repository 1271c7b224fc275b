"""Dev tool: evaluation latency for small problems (launch-bound regime), next to the same sums in plain NumPy."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from anml_b200 import _native as N

def one(n, p, family, link):
    rng = np.random.default_rng(1)
    X = rng.normal(size=(n, p)); X[:, 0] = 1.0
    beta = rng.normal(size=p) * 0.1
    if family == "binomial":
        obs = (rng.uniform(size=n) < 1 / (1 + np.exp(-X @ beta))).astype(float); se = None
    else:
        obs = X @ beta + rng.normal(size=n); se = np.ones(n)
    d = N.Design(n, p); d.set_columns(0, X)
    m = N.NativeModel(n, family, 1); m.set_param(0, d, N.LINK_CODES[link]); m.set_obs(obs)
    if se is not None: m.set_obs_se(se)
    x = 0.5 * beta
    for _ in range(20): m.eval(x)
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps): m.eval(x)
    gpu_us = (time.perf_counter() - t0) / reps * 1e6
    t0 = time.perf_counter()
    for _ in range(reps): m.eval(x, N.WANT_OBJ | N.WANT_GRAD)
    gpu_g_us = (time.perf_counter() - t0) / reps * 1e6
    creps = max(3, min(200, int(2e7 / (n * p))))
    t0 = time.perf_counter()
    for _ in range(creps):
        y = X @ x
        if family == "binomial":
            q = 1.0 / (1.0 + np.exp(-y)); r = q - obs; w = q * (1.0 - q)
            _ = (np.log1p(np.exp(-y)).sum() + ((1.0 - obs) * y).sum(), X.T @ r, (X.T * w) @ X)
        else:
            r = (y - obs) / se**2
            _ = (0.5 * (((y - obs) / se) ** 2).sum(), X.T @ r, (X.T / se**2) @ X)
    cpu_us = (time.perf_counter() - t0) / creps * 1e6
    print(json.dumps({"n": n, "p": p, "family": family, "gpu_eval_us": round(gpu_us, 1), "gpu_objgrad_us": round(gpu_g_us, 1), "numpy_us": round(cpu_us, 1)}))

if __name__ == "__main__":
    one(10_000, 4, "gaussian", "Identity")
    one(10_000, 64, "binomial", "Expit")
    one(100_000, 64, "binomial", "Expit")
    one(1_000_000, 64, "binomial", "Expit")
