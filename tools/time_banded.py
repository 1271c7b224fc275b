"""Dev tool: time the banded spline evaluator alone (config #3 shape: x and obs only)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anml_b200 import _native as N
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
nk, deg = 20, 3
nb = nk - 1 + deg
g = torch.Generator(device="cuda"); g.manual_seed(3)
xs = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
obs = torch.sin(2 * np.pi * xs) + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
knots = np.linspace(0, 1, nk)
d = N.Design(n, nb, deferred=True)
d.set_spline(0, None, knots, deg, x_dev_ptr=xs.data_ptr(), keepalive=xs)
m = N.NativeModel(n, "gaussian", 1)
m.set_param(0, d, N.LINK_CODES["Identity"])
m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
assert m.enable_banded_spline(knots, deg, x_dev_ptr=xs.data_ptr())
for kv in sys.argv[2:]:  # e.g. banded_register_prefetch=1
    k, v = kv.split("=")
    m.set_option(k, int(v))
x = np.linspace(-1, 1, nb)
for _ in range(5): m.eval(x)
m.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(30): o, gr, H = m.eval(x)
e2e = (time.perf_counter() - t0) / 30 * 1e3
kms, _, _ = m.kernel_time(reset=True)
for _ in range(30): m.eval(x, N.WANT_OBJ | N.WANT_GRAD)
kg, _, _ = m.kernel_time(reset=True)
print(json.dumps({"n": n, "kernel_ms": kms / 30, "objgrad_kernel_ms": kg / 30, "eval_ms": e2e, "GBps": 16.0 * n / (kms / 30) / 1e6,
                  "obj": o, "materialized": d.materialized, "kernel": m.main_kernel()}))
