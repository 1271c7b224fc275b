"""Small evaluations of every hand-written kernel family, for compute-sanitizer (memcheck / racecheck):
fused_hess (one parameter, signed weights, dual-predictor MODE 3 + MODE 1), fused_eval (gradient pass),
finish_fused, wide_hessian, the banded spline evaluator and the one-rank comm kernel.

    compute-sanitizer --tool memcheck  python tools/sanitizer_case.py
    compute-sanitizer --tool racecheck python tools/sanitizer_case.py

No torch: only the C-ABI, so the sanitizer sees just this library's kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from anml_b200 import _native as N

rng = np.random.default_rng(0)


def logistic(n, p, link="Expit", family="binomial"):
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    obs = (rng.uniform(size=n) < 0.5).astype(float)
    d = N.Design(n, p)
    d.set_columns(0, X)
    m = N.NativeModel(n, family, 1)
    m.set_param(0, d, N.LINK_CODES[link])
    m.set_obs(obs)
    m.set_gaussian_prior(0, np.zeros(p), np.ones(p))
    x = rng.normal(size=p) * 0.1
    o, g, H = m.eval(x)
    o2, g2, _ = m.eval(x, N.WANT_OBJ | N.WANT_GRAD)
    assert np.isfinite(o) and np.array_equal(H, H.T) and abs(o - o2) <= 1e-9 * abs(o)
    return o


print("fused p=64", logistic(3000, 64))
print("fused p=22 (64-row blocks)", logistic(3000, 22))
print("fused p=40 signed", logistic(2000, 40, link="Exp", family="gaussian"))
print("wide p=200", logistic(1500, 200))

# two parameters over one design: MODE 3 + MODE 1 + MODE 1
n, p = 2500, 64
X = rng.normal(size=(n, p))
X[:, 0] = 1.0
obs = X @ (rng.normal(size=p) * 0.1) + rng.normal(size=n)
d = N.Design(n, p)
d.set_columns(0, X)
m = N.NativeModel(n, "gaussian_meanvar", 2)
m.set_param(0, d, N.LINK_CODES["Identity"])
m.set_param(1, d, N.LINK_CODES["Exp"])
m.set_obs(obs)
o, g, H = m.eval(rng.normal(size=2 * p) * 0.05)
assert np.isfinite(o) and np.array_equal(H, H.T)
print("meanvar p=64", o)

# banded spline evaluator
n, nk, deg = 20000, 12, 3
xs = rng.uniform(size=n)
knots = np.linspace(0.0, 1.0, nk)
nb = nk - 1 + deg
d = N.Design(n, nb)
d.set_spline(0, xs, knots, deg)
m = N.NativeModel(n, "gaussian", 1)
m.set_param(0, d, N.LINK_CODES["Identity"])
m.set_obs(np.sin(6 * xs) + 0.1 * rng.normal(size=n))
assert m.enable_banded_spline(knots, deg, x=xs)
o, g, H = m.eval(np.linspace(-1, 1, nb))
assert np.isfinite(o) and np.array_equal(H, H.T)
print("banded", o)

# one-rank peer window
comm = N.PeerComm(0, 0, 1, 5000)
buf = N.DeviceBuffer(0, rng.normal(size=5000))
for _ in range(3):
    comm.allreduce_async(buf.ptr.value, 5000, 1)
print("comm", float(buf.download()[0]))
