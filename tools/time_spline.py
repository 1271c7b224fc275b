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
