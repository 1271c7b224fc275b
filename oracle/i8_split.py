"""NumPy restatement of the digit splitting of anml_b200/csrc/fused_hess_i8.cuh (TEST INFRASTRUCTURE ONLY: imported
by tests/, never by anml_b200/).  It states, in exact integer arithmetic, what the INT8-tensor-core Hessian kernel
computes, so the CPU suite can check the error bound the kernel's header claims without a GPU:

    t = x~ + (16 + B 2^-48),  B = 0x808080808080      (one rounding, to 48 fractional bits, |x~| < 1/4)
    digit a (1 = most significant) = signed byte (6 - a) of the mantissa of t, i.e. byte ^ 0x80
    x~ ~ sum_a d_a 2^(-8a);   H ~ sum over kept plane pairs 2^(-8(a+b)) D_a^T D_b,
    kept: ceil(a/2) + ceil(b/2) <= 4   (plane pairs {1,2},{3,4},{5,6}: pair products (0,0),(0,1),(1,0),(0,2),(1,1),(2,0)).
"""
import numpy as np

MAGIC = 16.0 + float(0x808080808080) / 2.0**48


def digits(xt):
    """xt: float64 array, |xt| < 1/4.  Returns int64 array (6, ...) of digits d_1 .. d_6 in [-128, 127]."""
    xt = np.asarray(xt, dtype=np.float64)
    t = xt + MAGIC  # NumPy adds in round-to-nearest-even, like the kernel's FMA (x * s is formed exactly there)
    bits = t.view(np.uint64)
    if np.any((bits >> np.uint64(48)) != np.uint64(0x4030)):
        raise ValueError("value outside the digit range")
    out = np.empty((6,) + xt.shape, dtype=np.int64)
    for a in range(1, 7):
        byte = ((bits >> np.uint64(8 * (6 - a))) & np.uint64(0xFF)).astype(np.int64)
        out[a - 1] = (byte ^ 0x80) - ((byte ^ 0x80) >> 7 << 8)  # two's complement byte -> signed
    return out


def reconstruct(d):
    return sum(d[a - 1].astype(np.float64) * 2.0 ** (-8 * a) for a in range(1, 7))


def kept(a, b):
    return (a + 1) // 2 + (b + 1) // 2 <= 4


def gram(Xt):
    """The kernel's Hessian of the scaled block Xt (n, p) in exact integer arithmetic (Python ints), as float64."""
    d = digits(Xt)  # (6, n, p)
    n, p = Xt.shape
    H = np.zeros((p, p), dtype=object)
    for a in range(1, 7):
        for b in range(1, 7):
            if not kept(a, b):
                continue
            S = d[a - 1].T.astype(object) @ d[b - 1].astype(object)  # exact
            H = H + S * (2 ** (8 * (12 - a - b)))  # common denominator 2^96
    return np.array([[float(v) for v in row] for row in H], dtype=np.float64) / 2.0**96
