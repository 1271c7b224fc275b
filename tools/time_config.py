"""Dev tool: time BASELINE configs #4 (two-parameter, shared X) and #5 (wide P) on one GPU."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from anml_b200 import _native as N

def gen(n, p, dev, scale):
    g = torch.Generator(device=dev); g.manual_seed(20261019)
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    beta = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * scale
    chunk = max(1, (1 << 28) // p)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g); X[lo:hi, 0] = 1.0
    return X, beta, g

def main():
    mode = sys.argv[1]; n = int(float(sys.argv[2])); p = int(sys.argv[3])
    dev = torch.device("cuda:0")
    if mode == "c5":
        X, beta, g = gen(n, p, dev, 0.1 / np.sqrt(p / 64))
        obs = torch.empty(n, dtype=torch.float64, device=dev)
        chunk = max(1, (1 << 28) // p)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            obs[lo:hi] = (torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(X[lo:hi] @ beta)).double()
        d = N.Design.wrap(X.data_ptr(), n, p, p, keepalive=X)
        m = N.NativeModel(n, "binomial", 1); m.set_param(0, d, N.LINK_CODES["Expit"]); m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
        x = (0.5 * beta).cpu().numpy(); nblocks = 1
        flops = n * p * (p + 1.0)
    else:
        X, beta, g = gen(n, p, dev, 0.1)
        bs = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.05
        obs = torch.empty(n, dtype=torch.float64, device=dev)
        chunk = 1 << 22
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            obs[lo:hi] = X[lo:hi] @ beta + torch.exp(0.5 * (X[lo:hi] @ bs)) * torch.randn(hi - lo, dtype=torch.float64, device=dev, generator=g)
        d = N.Design.wrap(X.data_ptr(), n, p, p, keepalive=X)
        m = N.NativeModel(n, "gaussian_meanvar", 2)
        m.set_param(0, d, N.LINK_CODES["Identity"]); m.set_param(1, d, N.LINK_CODES["Exp"]); m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
        x = np.concatenate([(0.5 * beta).cpu().numpy(), (0.5 * bs).cpu().numpy()])
        flops = 3 * n * p * (p + 1.0)
    torch.cuda.synchronize()
    for _ in range(2):
        o, gr, H = m.eval(x)
    m.kernel_time(reset=True)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        o, gr, H = m.eval(x)
    e2e = (time.perf_counter() - t0) / reps * 1e3
    ms, nl, al = m.kernel_time(reset=True)
    print(json.dumps({"mode": mode, "n": n, "p": p, "e2e_ms": e2e, "main_kernels_ms_per_eval": ms / reps, "main_launches_per_eval": nl / reps,
                      "launches_per_eval": al / reps, "tflops_syrk_on_main_kernels": flops / (ms / reps) / 1e9, "tflops_syrk_e2e": flops / e2e / 1e9,
                      "obj": o, "hess_sym": bool(np.array_equal(H, H.T))}))

if __name__ == "__main__":
    main()
