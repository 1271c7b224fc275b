"""Measure the cuBLAS DGEMM rate on this box: the FP64 roofline denominator
that MEASURED_PEAKS.json does not carry."""
import json
import torch

def main():
    dev = torch.device("cuda:0")
    out = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"dgemm_{n}_tflops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    # skinny SYRK-like shape of the Hessian: (64 x K) @ (K x 64)
    k = 1 << 22
    x = torch.randn(k, 64, dtype=torch.float64, device=dev)
    for _ in range(2):
        x.t() @ x
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); x.t() @ x; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["xtx_64_k4M_ms"] = best
    out["xtx_64_k4M_tflops_gemm_equiv"] = 2.0 * k * 64 * 64 / (best * 1e-3) / 1e12
    print(json.dumps(out))

if __name__ == "__main__":
    main()
