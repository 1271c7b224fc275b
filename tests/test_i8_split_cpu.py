"""CPU check of the digit splitting behind the INT8-tensor-core Hessian (csrc/fused_hess_i8.cuh), on its NumPy
restatement oracle/i8_split.py: the digits rebuild the rounded value exactly, and the kept plane-pair products give the
Gram matrix to ~2^-47 per row -- the bound the kernel's header states."""
import numpy as np
import pytest

from oracle import i8_split as S


def test_digits_rebuild_the_value_rounded_to_48_bits():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-0.2499, 0.2499, 20000), [0.0, 2.0**-60, -2.0**-49, 0.2499999, -0.2499999, 2.0**-48]])
    d = S.digits(x)
    assert d.min() >= -128 and d.max() <= 127
    rebuilt = S.reconstruct(d)
    assert np.array_equal(rebuilt, np.round(x * 2.0**48) / 2.0**48)  # exactly the 48-bit rounding
    assert np.abs(rebuilt - x).max() <= 2.0**-49


def test_out_of_range_is_detected():
    with pytest.raises(ValueError):
        S.digits(np.array([0.6]))
    with pytest.raises(ValueError):
        S.digits(np.array([-0.51]))


def test_kept_pairs():
    kept = {(a, b) for a in range(1, 7) for b in range(1, 7) if S.kept(a, b)}
    assert all((b, a) in kept for (a, b) in kept)                      # symmetric set: the result is symmetric
    assert all((a, b) in kept for a in range(1, 7) for b in range(1, 7) if a + b <= 7)
    assert {(a, b) for (a, b) in kept if a + b > 7} == {(2, 6), (6, 2), (4, 4)}


def test_gram_matrix_error_bound():
    rng = np.random.default_rng(2)
    n, p = 400, 6
    Xt = rng.uniform(-0.25, 0.25, size=(n, p)) * rng.uniform(0.0, 1.0, size=(n, 1))  # sqrt(w) <= 1 folded in
    H = S.gram(Xt)
    want = Xt.T @ Xt
    assert np.array_equal(H, H.T)
    # per row: two roundings of 2^-49 against factors < 1/4, plus dropped products of weight 2^-64 * 2^14
    bound = n * (2 * 2.0**-49 * 0.25 + 2.0**-49 * 2.0**-49 + 12 * 2.0**-50)
    assert np.abs(H - want).max() <= bound
    assert np.abs(H - want).max() <= 1e-12 * np.abs(want).max()


def test_precision_falls_quadratically_with_a_loose_scale():
    """Why the digit scale must follow the data tightly (block (0,0) of the mean / variance model stays on
    DMMA, block (1,1) takes its scale from the exact max w): values that sit 2^-k below the scale lose 2^2k of relative
    accuracy, because the dropped digit-plane products assume a leading digit in plane 1."""
    rng = np.random.default_rng(3)
    n, p = 64, 4
    X = rng.uniform(-0.25, 0.25, size=(n, p))
    rel = []
    for k in (0, 5, 10):
        Xt = X * 2.0**-k
        H = S.gram(Xt)
        want = Xt.T @ Xt
        rel.append(np.abs(H - want).max() / np.abs(want).max())
    assert rel[0] <= 1e-13                              # 5e-15 measured
    assert 1e-13 < rel[1] <= 1e-10                      # 7e-12: ~2^10 worse
    assert rel[2] > 100 * rel[1] and rel[2] > 1e-9      # 5e-9: ~2^20 worse, outside BASELINE's Hessian tolerance
