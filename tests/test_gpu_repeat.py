"""Run-to-run repeatability of every kernel family, as a stand-in for compute-sanitizer racecheck (closed on
this pool: profiles/r2_sanitizer_note.txt).  Every result is a fixed-order sum, so a data race in the
helper -> tensor-warp hand-off of fused_hess_kernel (csrc/fused_hess.cuh), in the TMA / pair-barrier / in-place digit
write / tcgen05 commit chain of fused_hess_i8_kernel, in the full/empty barriers of wide_hessian_kernel or in the
stage ring of spline_banded_bulk_kernel shows up as a run-to-run difference of
some bit: each case is evaluated REPS times at a size that gives every CTA several blocks and must return
bit-identical records, at alternating coefficient vectors (so a stale buffer from the previous launch is
caught too)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPS = 30


def repeat_eval(model, xs, want=None):
    first = {}
    for k in range(REPS):
        x = xs[k % len(xs)]
        out = model.eval(x) if want is None else model.eval(x, want)
        key = k % len(xs)
        if key not in first:
            first[key] = out
            assert np.isfinite(out[0])
            continue
        ref = first[key]
        assert out[0] == ref[0], f"objective differs on repetition {k}"
        assert np.array_equal(out[1], ref[1]), f"gradient differs on repetition {k}"
        if out[2] is not None and ref[2] is not None:
            assert np.array_equal(out[2], ref[2]), f"Hessian differs on repetition {k}"
    return first[0]


def logistic_model(n, p, seed, link="Expit", family="binomial"):
    from anml_b200 import _native as N

    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    beta = rng.normal(size=p) * 0.1
    if family == "binomial":
        obs = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(float)
    else:
        obs = np.exp(X @ beta) + 0.1 * rng.normal(size=n)
    design = N.Design(n, p)
    design.set_columns(0, X)
    model = N.NativeModel(n, family, 1)
    model.set_param(0, design, N.LINK_CODES[link])
    model.set_obs(obs)
    model.set_gaussian_prior(0, np.zeros(p), np.ones(p))
    return model, [0.5 * beta, 0.4 * beta]


@pytest.mark.parametrize("n,p", [(600_011, 64), (400_003, 48), (700_001, 22), (900_007, 16), (300_000, 7)])
def test_repeatability_fused(n, p):
    model, xs = logistic_model(n, p, seed=n + p)
    o, g, H = repeat_eval(model, xs)
    assert np.array_equal(H, H.T)


def test_repeatability_fused_dmma_at_64_columns():
    # p = 64 logistic defaults to the INT8-tensor-core Hessian (csrc/fused_hess_i8.cuh, covered above); this is the DMMA
    # kernel at the same shape
    model, xs = logistic_model(600_011, 64, seed=8)
    model.set_option("hess_i8", 0)
    o, g, H = repeat_eval(model, xs)
    assert model.main_kernel().startswith("fused_hess_kernel"), model.main_kernel()
    assert np.array_equal(H, H.T)


def test_repeatability_fused_gradient_pass():
    from anml_b200 import _native as N

    model, xs = logistic_model(500_009, 64, seed=5)
    repeat_eval(model, xs, want=N.WANT_OBJ | N.WANT_GRAD)


def test_repeatability_fused_signed_weights():
    # gaussian / exp: w = f'^2 + (theta - obs) f'' changes sign from row to row (SIGNED instantiation)
    model, xs = logistic_model(300_017, 40, seed=11, link="Exp", family="gaussian")
    repeat_eval(model, xs)


def test_repeatability_meanvar():
    from anml_b200 import _native as N

    rng = np.random.default_rng(21)
    n, p = 400_009, 64
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    bm, bs = rng.normal(size=p) * 0.1, rng.normal(size=p) * 0.05
    obs = X @ bm + np.exp(0.5 * (X @ bs)) * rng.normal(size=n)
    design = N.Design(n, p)
    design.set_columns(0, X)
    model = N.NativeModel(n, "gaussian_meanvar", 2)
    model.set_param(0, design, N.LINK_CODES["Identity"])
    model.set_param(1, design, N.LINK_CODES["Exp"])
    model.set_obs(obs)
    x0 = np.concatenate([0.5 * bm, 0.5 * bs])
    o, g, H = repeat_eval(model, [x0, 0.9 * x0])
    assert np.array_equal(H, H.T)


def test_repeatability_wide():
    model, xs = logistic_model(60_013, 200, seed=31)
    o, g, H = repeat_eval(model, xs)
    assert np.array_equal(H, H.T)


def test_repeatability_banded_spline():
    from anml_b200 import _native as N

    rng = np.random.default_rng(41)
    n, nk, deg = 3_000_017, 20, 3
    xs = rng.uniform(size=n)
    knots = np.linspace(0.0, 1.0, nk)
    nb = nk - 1 + deg
    design = N.Design(n, nb)
    design.set_spline(0, xs, knots, deg)
    model = N.NativeModel(n, "gaussian", 1)
    model.set_param(0, design, N.LINK_CODES["Identity"])
    model.set_obs(np.sin(6 * xs) + 0.1 * rng.normal(size=n))
    assert model.enable_banded_spline(knots, deg, x=xs)
    o, g, H = repeat_eval(model, [np.linspace(-1, 1, nb), np.linspace(1, -1, nb)])
    assert np.array_equal(H, H.T)
