"""Dev tool: time the fused evaluator on synthetic device-resident data."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from anml_b200 import _native as N

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    want_h = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(20261019)
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    chunk = 1 << 22
    beta_true = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.1
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g)
        X[lo:hi, 0] = 1.0
        prob = torch.sigmoid(X[lo:hi] @ beta_true)
        obs[lo:hi] = (torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) < prob).double()
    torch.cuda.synchronize()
    design = N.Design.wrap(X.data_ptr(), n, p, p, keepalive=X)
    model = N.NativeModel(n, "binomial", 1)
    model.set_param(0, design, N.LINK_CODES["Expit"])
    model.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
    for kv in sys.argv[4:]:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    x = (0.5 * beta_true).cpu().numpy()
    want = N.WANT_OBJ | N.WANT_GRAD | (N.WANT_HESS if want_h else 0)
    for _ in range(3):
        o, gr, H = model.eval(x, want)
    model.kernel_time(reset=True)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        o, gr, H = model.eval(x, want)
    t1 = time.perf_counter()
    ms, nl, al = model.kernel_time(reset=True)
    k_ms = ms / nl
    bytes_ = 8.0 * n * p + 8.0 * n
    flops = n * p * (p + 1.0)
    # spot check vs torch on a slice
    m = min(n, 1 << 20)
    y = X[:m] @ torch.tensor(x, device=dev)
    th = torch.sigmoid(y)
    g_ref = (X[:m].t() @ (th - obs[:m])).cpu().numpy()
    print(json.dumps({"n": n, "p": p, "hess": want_h, "kernel_ms": k_ms, "e2e_ms": (t1 - t0) / reps * 1e3,
                      "gbs": bytes_ / k_ms / 1e6, "tflops_syrk": flops / k_ms / 1e9 if want_h else 0, "obj": o,
                      "launches_per_eval": al / reps}))

if __name__ == "__main__":
    main()
