"""Dev tool: time the on-device spline design build (BASELINE config #3 shape)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anml_b200 import _native as N
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
nk, deg = 20, 3
nb = nk - 1 + deg
g = torch.Generator(device="cuda"); g.manual_seed(3)
xs = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
knots = np.linspace(0, 1, nk)
d = N.Design(n, nb)
for _ in range(2):
    d.set_spline(0, None, knots, deg, x_dev_ptr=xs.data_ptr())
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); d.set_spline(0, None, knots, deg, x_dev_ptr=xs.data_ptr()); ts.append(time.perf_counter() - t0)
t = min(ts)
print(json.dumps({"n": n, "nb": nb, "build_ms": t * 1e3, "write_GBps": n * nb * 8 / t / 1e9, "rows_per_s": n / t}))

# config #3 evaluation on the dense 22-column design built above: gaussian NLL, identity link
obs = torch.sin(2 * np.pi * xs) + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
se = torch.ones(n, dtype=torch.float64, device="cuda")
m = N.NativeModel(n, "gaussian", 1)
m.set_param(0, d, N.LINK_CODES["Identity"])
m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs); m.set_obs_se(dev_ptr=se.data_ptr(), keepalive=se)
x = np.linspace(-1, 1, nb)
for _ in range(3): m.eval(x)
m.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(10): o, gr, H = m.eval(x)
e2e = (time.perf_counter() - t0) / 10 * 1e3
kms, nl, al = m.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(10): m.eval(x, N.WANT_OBJ | N.WANT_GRAD)
e2e_g = (time.perf_counter() - t0) / 10 * 1e3
band = int(np.max(np.abs(np.subtract.outer(np.arange(nb), np.arange(nb)))[H != 0]))
print(json.dumps({"eval_ms": e2e, "main_kernel_ms": kms / 10, "objgrad_ms": e2e_g, "bytes_dense_GB": n * (nb + 2) * 8 / 1e9,
                  "dense_GBps": n * (nb + 2) * 8 / (kms / 10) / 1e6, "hessian_bandwidth": band}))

# the same model through the banded evaluator (basis on the fly from x, rows sorted once)
t0 = time.perf_counter(); ok = m.enable_banded_spline(knots, deg, x_dev_ptr=xs.data_ptr()); torch.cuda.synchronize(); t_en = time.perf_counter() - t0
for _ in range(3): m.eval(x)
m.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(10): ob, gb, Hb = m.eval(x)
e2e_b = (time.perf_counter() - t0) / 10 * 1e3
kms_b, _, _ = m.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(10): m.eval(x, N.WANT_OBJ | N.WANT_GRAD)
e2e_bg = (time.perf_counter() - t0) / 10 * 1e3
print(json.dumps({"banded_enabled": bool(ok), "enable_ms": t_en * 1e3, "banded_eval_ms": e2e_b, "banded_kernel_ms": kms_b / 10, "banded_objgrad_ms": e2e_bg,
                  "obj_rel_diff": abs(ob - o) / abs(o), "grad_rel_diff": float(np.abs(gb - gr).max() / np.abs(gr).max()),
                  "hess_rel_diff": float(np.abs(Hb - H).max() / np.abs(H).max()), "hess_sym": bool(np.array_equal(Hb, Hb.T)),
                  "algorithmic_GBps": n * 3 * 8 / (kms_b / 10) / 1e6}))

# BASELINE config #3 proper: default obs_se (not read), x and obs only -> 16 bytes per row
m2 = N.NativeModel(n, "gaussian", 1)
m2.set_param(0, d, N.LINK_CODES["Identity"])
m2.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
ok2 = m2.enable_banded_spline(knots, deg, x_dev_ptr=xs.data_ptr())
for _ in range(3): m2.eval(x)
m2.kernel_time(reset=True)
t0 = time.perf_counter()
for _ in range(20): o2, g2, H2 = m2.eval(x)
e2e_2 = (time.perf_counter() - t0) / 20 * 1e3
kms_2, _, _ = m2.kernel_time(reset=True)
m2.kernel_time(reset=True)
for _ in range(20): m2.eval(x, N.WANT_OBJ | N.WANT_GRAD)
kms_2g, _, _ = m2.kernel_time(reset=True)
print(json.dumps({"plain_banded_enabled": bool(ok2), "eval_ms": e2e_2, "kernel_ms": kms_2 / 20, "objgrad_kernel_ms": kms_2g / 20,
                  "algorithmic_GBps": n * 2 * 8 / (kms_2 / 20) / 1e6, "obj_rel_diff_vs_dense": abs(o2 - o) / abs(o),
                  "grad_rel_diff_vs_dense": float(np.abs(g2 - gr).max() / np.abs(gr).max()),
                  "hess_rel_diff_vs_dense": float(np.abs(H2 - H).max() / np.abs(H).max())}))
