"""GPU parity of the evaluation kernels against the NumPy oracle, through the
C-ABI (ctypes).  Tolerances are BASELINE.json's: objective / gradient 1e-10,
Hessian 1e-9, relative (gradient and Hessian relative to their largest entry)."""
import numpy as np
import pytest

from oracle import anml_oracle as O

pytestmark = pytest.mark.gpu

LINK_NAME = {"identity": "Identity", "exp": "Exp", "log": "Log", "expit": "Expit", "logit": "Logit"}


def make_problem(family, n, p, seed, link=None, with_offset=False, with_se=False):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, p))
    if p > 0:
        X[:, 0] = 1.0
    beta_true = rng.normal(size=p) * 0.1
    y = X @ beta_true
    off = rng.normal(size=n) * 0.2 if with_offset else None
    if off is not None:
        y = y + off
    se = rng.uniform(0.5, 1.5, size=n) if with_se else None
    if family == "gaussian":
        link = link or "identity"
        obs = O.link_eval(link, y) + 0.1 * rng.normal(size=n)
    elif family == "binomial":
        link = link or "expit"
        obs = (rng.uniform(size=n) < 1 / (1 + np.exp(-y))).astype(float)
    else:
        link = link or "exp"
        obs = rng.poisson(np.exp(y)).astype(float)
    return X, obs, se, off, link, 0.5 * beta_true


def gpu_eval(family, X, obs, se, off, link, x, prior=None, options=()):
    from anml_b200 import _native as N

    n, p = X.shape
    design = N.Design(n, p)
    design.set_columns(0, X)
    off_buf = N.DeviceBuffer(0, off) if off is not None else None
    model = N.NativeModel(n, family, 1)
    model.set_param(0, design, N.LINK_CODES[LINK_NAME[link]], off_buf)
    model.set_obs(obs)
    if se is not None:
        model.set_obs_se(se)
    if prior is not None:
        model.set_gaussian_prior(0, *prior)
    for name, val in options:
        model.set_option(name, val)
    out = model.eval(x)
    out2 = model.eval(x)  # determinism: bitwise identical on repeat
    assert np.array_equal(out[0], out2[0], equal_nan=True) and np.array_equal(out[1], out2[1], equal_nan=True)
    assert np.array_equal(out[2], out2[2], equal_nan=True)
    return out


def check(got, want, og=1e-10, h=1e-9):
    o, g, H = got
    ow, gw, Hw = want
    assert abs(o - ow) <= og * abs(ow)
    assert np.abs(g - gw).max() <= og * max(np.abs(gw).max(), 1e-300)
    assert np.abs(H - Hw).max() <= h * np.abs(Hw).max()
    assert np.array_equal(H, H.T)


@pytest.mark.parametrize("p", [1, 4, 16, 17, 22, 33, 48, 63, 64])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 20011])
def test_binomial_expit_shapes(n, p):
    X, obs, se, off, link, x = make_problem("binomial", n, p, seed=n * 100 + p)
    want = O.model_eval_fused("binomial", obs, [X], [link], x)
    check(gpu_eval("binomial", X, obs, se, off, link, x), want)


@pytest.mark.parametrize("family,link", [("gaussian", "identity"), ("gaussian", "exp"), ("binomial", "expit"),
                                         ("poisson", "exp"), ("gaussian", "expit")])
@pytest.mark.parametrize("fast", [0, 1])
def test_families_links_offset_se(family, link, fast):
    X, obs, se, off, link, x = make_problem(family, 5003, 7, seed=7, link=link, with_offset=True, with_se=(family == "gaussian"))
    want = O.model_eval_fused(family, obs, [X], [link], x, offsets=[off], obs_se=se)
    got = gpu_eval(family, X, obs, se, off, link, x, options=(("no_fast_math_paths", 1 - fast),))
    check(got, want)


@pytest.mark.parametrize("p", [4, 22, 40, 64])
def test_symmetric_warp_kernel_variant(p):
    """fused_eval_kernel<.., HESS, MODE 0> (all warps alike) stays available behind an option."""
    X, obs, se, off, link, x = make_problem("binomial", 7001, p, seed=p, with_offset=True)
    want = O.model_eval_fused("binomial", obs, [X], [link], x, offsets=[off])
    check(gpu_eval("binomial", X, obs, se, off, link, x, options=(("no_warp_specialisation", 1),)), want)


@pytest.mark.parametrize("p", [4, 40, 64])
def test_signed_weight_path(p):
    """The tensor warp multiplies sqrt(|w|) x fragments; rows with w < 0 flip the B
    operand's sign bit.  Force that path on a canonical model (all w > 0) and run a
    non-canonical one (gaussian + exp link: w changes sign with the residual)."""
    X, obs, se, off, link, x = make_problem("binomial", 6007, p, seed=p + 7, with_offset=True)
    want = O.model_eval_fused("binomial", obs, [X], [link], x, offsets=[off])
    check(gpu_eval("binomial", X, obs, se, off, link, x, options=(("force_signed_weights", 1),)), want)
    X, obs, se, off, link, x = make_problem("gaussian", 6007, p, seed=p + 9, link="exp")
    obs = obs + 2.0 * np.random.default_rng(p).normal(size=obs.size)  # large residuals -> negative weights
    want = O.model_eval_fused("gaussian", obs, [X], ["exp"], x)
    w_neg = ((np.exp(X @ x) - obs) * np.exp(X @ x) + np.exp(2 * (X @ x)) < 0).mean()
    assert 0.02 < w_neg < 0.98
    check(gpu_eval("gaussian", X, obs, None, None, "exp", x), want)


def test_general_path_matches_fused():
    X, obs, se, off, link, x = make_problem("binomial", 4099, 37, seed=3, with_offset=True)
    want = O.model_eval_fused("binomial", obs, [X], [link], x, offsets=[off])
    got = gpu_eval("binomial", X, obs, se, off, link, x, options=(("force_generic", 1),))
    check(got, want)


def test_priors_direct_and_linear():
    X, obs, se, off, link, x = make_problem("binomial", 3000, 6, seed=5)
    rng = np.random.default_rng(1)
    mean, sd = rng.normal(size=6) * 0.1, np.array([1.0, 0.5, np.inf, 2.0, 1.0, 3.0])
    M, lm, ls = rng.normal(size=(9, 6)), rng.normal(size=9), rng.uniform(0.5, 2.0, size=9)
    want = O.model_eval_fused("binomial", obs, [X], [link], x, priors=[{"direct": (mean, sd), "linear": (lm, ls, M)}])
    check(gpu_eval("binomial", X, obs, se, off, link, x, prior=(mean, sd, M, lm, ls)), want)


def test_golden_model_cases(golden):
    for case, fam in [("gauss_p4", "gaussian"), ("binom_p6", "binomial"), ("binom_p64", "binomial"),
                      ("poisson_p5_off", "poisson"), ("gauss_exp_p3", "gaussian")]:
        X = golden[f"model/{case}/X"]
        link = str(golden[f"model/{case}/links"])
        key = f"model/{case}/offset"
        off = golden[key] if key in golden.files else None
        sd = float(golden[f"model/{case}/prior_sd"])
        p = X.shape[1]
        prior = (np.zeros(p), np.full(p, sd)) if sd > 0 else None
        got = gpu_eval(fam, X, golden[f"model/{case}/obs"], golden[f"model/{case}/se"], off, link, golden[f"model/{case}/x"], prior=prior)
        check(got, (float(golden[f"model/{case}/obj"]), golden[f"model/{case}/grad"], golden[f"model/{case}/hess"]))


def test_log_link_domain_error():
    X, obs, se, off, link, x = make_problem("gaussian", 100, 3, seed=2)
    with pytest.raises(ValueError):
        gpu_eval("gaussian", X, np.abs(obs) + 1, se, off, "log", -np.abs(x) - 1.0)


@pytest.mark.parametrize("n,p", [(3000, 65), (5001, 128), (2047, 200), (1500, 257), (700, 520)])
def test_wide_designs_use_the_blocked_hessian_kernel(n, p):
    """p > 64: general path (linpred -> rowterms -> xtr + wide_hessian_kernel)."""
    X, obs, se, off, link, x = make_problem("binomial", n, p, seed=p, with_offset=True)
    x = x * (4.0 / np.sqrt(p))  # keep |X beta| moderate for wide designs
    want = O.model_eval_fused("binomial", obs, [X], [link], x, offsets=[off])
    check(gpu_eval("binomial", X, obs, se, off, link, x), want)


def gpu_eval_meanvar(X0, X1, obs, x, links=("identity", "exp"), share=False, options=(), want_hess=True):
    from anml_b200 import _native as N

    n = X0.shape[0]
    d0 = N.Design(n, X0.shape[1]); d0.set_columns(0, X0)
    if share:
        d1 = d0
    else:
        d1 = N.Design(n, X1.shape[1]); d1.set_columns(0, X1)
    model = N.NativeModel(n, "gaussian_meanvar", 2)
    model.set_param(0, d0, N.LINK_CODES[LINK_NAME[links[0]]])
    model.set_param(1, d1, N.LINK_CODES[LINK_NAME[links[1]]])
    model.set_obs(obs)
    for name, val in options:
        model.set_option(name, val)
    if not want_hess:
        return model.eval(x, N.WANT_OBJ | N.WANT_GRAD)
    return model.eval(x)


@pytest.mark.parametrize("p0,p1,share", [(5, 5, True), (5, 7, False), (64, 64, True), (70, 70, True), (40, 130, False)])
def test_two_parameter_mean_variance_model(p0, p1, share):
    rng = np.random.default_rng(p0 * 1000 + p1)
    n = 2500
    X0 = rng.normal(size=(n, p0)); X0[:, 0] = 1.0
    X1 = X0 if share else rng.normal(size=(n, p1))
    X1[:, 0] = 1.0
    b0 = rng.normal(size=p0) * 0.2 / np.sqrt(p0 / 5)
    b1 = rng.normal(size=p1) * 0.1 / np.sqrt(p1 / 5)
    obs = X0 @ b0 + np.exp(0.5 * (X1 @ b1)) * rng.normal(size=n)
    x = np.concatenate([0.7 * b0, 0.7 * b1])
    want = O.model_eval_fused("gaussian_meanvar", obs, [X0, X1], ["identity", "exp"], x)
    check(gpu_eval_meanvar(X0, X1, obs, x, share=share), want)
    # the general row-term expressions (no identity/exp shortcut) and the gradient-only evaluation
    check(gpu_eval_meanvar(X0, X1, obs, x, share=share, options=(("no_fast_math_paths", 1),)), want)
    o, g, _ = gpu_eval_meanvar(X0, X1, obs, x, share=share, want_hess=False)
    assert abs(o - want[0]) <= 1e-10 * abs(want[0]) and np.abs(g - want[1]).max() <= 1e-10 * np.abs(want[1]).max()
    if share and p0 <= 64:  # without the warp-specialised kernel: dual-predictor pass + term-fed blocks
        check(gpu_eval_meanvar(X0, X1, obs, x, share=share, options=(("no_warp_specialisation", 1),)), want)


@pytest.mark.parametrize("share,p", [(True, 9), (False, 9), (True, 70)])
def test_two_parameter_model_with_offsets_and_links(share, p):
    """Offsets on both parameters, non-default links; shared design (dual-predictor pass),
    separate designs (linpred + rowterms) and a wide shared design (blocked Hessian)."""
    from anml_b200 import _native as N

    rng = np.random.default_rng(p + (1 if share else 0))
    n = 3001
    X0 = rng.normal(size=(n, p)); X0[:, 0] = 1.0
    X1 = X0 if share else np.column_stack([np.ones(n), rng.normal(size=(n, p - 1))])
    b0 = rng.normal(size=p) * 0.15 / np.sqrt(p / 9)
    b1 = rng.normal(size=p) * 0.1 / np.sqrt(p / 9)
    off0, off1 = rng.normal(size=n) * 0.2, rng.normal(size=n) * 0.1
    mu = np.exp(X0 @ b0 + off0)
    obs = mu + np.exp(0.5 * (X1 @ b1 + off1)) * rng.normal(size=n)
    x = np.concatenate([0.8 * b0, 0.8 * b1])
    want = O.model_eval_fused("gaussian_meanvar", obs, [X0, X1], ["exp", "exp"], x, offsets=[off0, off1])
    d0 = N.Design(n, p); d0.set_columns(0, X0)
    d1 = d0 if share else N.Design(n, p)
    if not share:
        d1.set_columns(0, X1)
    o0, o1 = N.DeviceBuffer(0, off0), N.DeviceBuffer(0, off1)
    m = N.NativeModel(n, "gaussian_meanvar", 2)
    m.set_param(0, d0, N.LINK_CODES["Exp"], o0)
    m.set_param(1, d1, N.LINK_CODES["Exp"], o1)
    m.set_obs(obs)
    check(m.eval(x), want)
    if share and p <= 64:
        m.set_option("force_generic", 1)  # same numbers through linpred + rowterms
        check(m.eval(x), want)


@pytest.mark.parametrize("family,link,p", [("binomial", "expit", 64), ("gaussian", "exp", 5), ("poisson", "exp", 20)])
def test_row_weights_scale_every_term(family, link, p):
    """Per-row likelihood weights (trimming): fused path, symmetric variant and general path."""
    from anml_b200 import _native as N

    X, obs, se, off, link, x = make_problem(family, 4003, p, seed=p, link=link, with_offset=True)
    wts = np.random.default_rng(p).uniform(0.0, 2.0, size=obs.size)
    wts[::7] = 0.0  # trimmed rows
    want = O.model_eval_fused(family, obs, [X], [link], x, offsets=[off], weights=wts)
    for opts in ((), (("no_warp_specialisation", 1),), (("force_generic", 1),)):
        d = N.Design(X.shape[0], p); d.set_columns(0, X)
        ob = N.DeviceBuffer(0, off)
        m = N.NativeModel(X.shape[0], family, 1)
        m.set_param(0, d, N.LINK_CODES[LINK_NAME[link]], ob)
        m.set_obs(obs)
        m.set_weights(wts)
        for k, v in opts:
            m.set_option(k, v)
        check(m.eval(x), want)
    with pytest.raises(ValueError):
        m.set_weights(-wts - 1.0)
    m.set_weights(None)  # back to unweighted
    check(m.eval(x), O.model_eval_fused(family, obs, [X], [link], x, offsets=[off]))


def test_logistic_objective_product_accumulator_extremes():
    """The Hessian kernel multiplies the factors 1 + e^-y of the logistic NLL and takes one
    log at the end (rowmath.cuh LogProduct).  Wide linear predictors exercise the exponent
    bookkeeping (|y| up to ~600: factors up to e^600), an overflowing row must give +inf and
    a NaN row NaN, as the per-row log did."""
    rng = np.random.default_rng(17)
    n, p = 40_003, 8
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    x = rng.normal(size=p)
    x[1] = 150.0  # y spreads over +-600
    y = X @ x
    keep = np.abs(y) < 600.0
    X, y = X[keep], y[keep]
    obs = (rng.uniform(size=X.shape[0]) < 0.5).astype(float)
    got = gpu_eval("binomial", X, obs, None, None, "expit", x)
    # stable reference for the objective: log(1 + e^-y) = logaddexp(0, -y)
    want_obj = np.logaddexp(0.0, -y).sum() + ((1.0 - obs) * y).sum()
    assert np.isfinite(got[0]) and abs(got[0] - want_obj) <= 1e-12 * abs(want_obj)
    q = 1.0 / (1.0 + np.exp(-y))
    want_g = X.T @ (q - obs)
    assert np.abs(got[1] - want_g).max() <= 1e-10 * np.abs(want_g).max()

    Xo = X[:1000].copy()
    Xo[5, 1] = -20.0  # y ~ -3000: e^-y overflows
    got = gpu_eval("binomial", Xo, obs[:1000], None, None, "expit", x)
    assert got[0] == np.inf
    Xn = X[:1000].copy()
    Xn[7, 2] = np.nan
    got = gpu_eval("binomial", Xn, obs[:1000], None, None, "expit", x)
    assert np.isnan(got[0])


# ---- Log / Logit mappings on the device: VALUES against the reference (SURVEY.md row a7,
# src/anml/parameter/smoothmapping.py:96-104, :146-156), not only the domain errors
@pytest.mark.parametrize("link", ["log", "logit"])
@pytest.mark.parametrize("general", [0, 1])
def test_log_and_logit_links_in_a_model(link, general):
    rng = np.random.default_rng(17)
    n, p = 4003, 6
    X = np.abs(rng.normal(size=(n, p))) + 0.05
    X[:, 0] = 1.0
    beta = np.abs(rng.normal(size=p)) * 0.2 + 0.05
    if link == "logit":  # keep the linear predictor strictly inside (0, 1)
        scale = 0.9 / (X @ beta).max()
        beta = beta * scale
    y = X @ beta
    assert (y > 0).all() and (link == "log" or (y < 1).all())
    obs = O.link_eval(link, y) + 0.1 * rng.normal(size=n)
    se = rng.uniform(0.5, 1.5, size=n)
    want = O.model_eval_fused("gaussian", obs, [X], [link], beta, obs_se=se)
    got = gpu_eval("gaussian", X, obs, se, None, link, beta, options=(("force_generic", general),))
    check(got, want)
    with pytest.raises(ValueError):  # outside the domain: the reference raises before computing (:98-99, :148-151)
        gpu_eval("gaussian", X, obs, se, None, link, -beta, options=(("force_generic", general),))


# ---------------------------------------------------------------------------------------------
# Hessian on the INT8 tensor cores (csrc/fused_hess_i8.cuh): 64-column logistic / unit-variance Gaussian models
# ---------------------------------------------------------------------------------------------
def i8_model(family, n, seed, scale_cols=False, offset=False, prior=True):
    from anml_b200 import _native as N

    X, obs, se, off, link, x = make_problem(family, n, 64, seed, with_offset=offset)
    if scale_cols:  # columns of very different magnitude: the per-column power-of-two scales matter
        X = X * np.exp2(np.arange(64) % 29 - 14)[None, :]
        X[:, 0] = 1.0
        x = x / np.exp2(np.arange(64) % 29 - 14)
        x[0] *= np.exp2(-14.0)
    design = N.Design(n, 64)
    design.set_columns(0, X)
    off_buf = N.DeviceBuffer(0, off) if off is not None else None
    model = N.NativeModel(n, family, 1)
    model.set_param(0, design, N.LINK_CODES[LINK_NAME[link]], off_buf)
    model.set_obs(obs)
    pr = None
    if prior:
        pr = {"direct": (np.zeros(64), np.full(64, 2.0))}
        model.set_gaussian_prior(0, *pr["direct"])
    model.set_option("hess_i8", 2)  # also below the 8192-row minimum of the default
    return model, (X, obs, off, link, x, pr), (design, off_buf)


@pytest.mark.parametrize("family", ["binomial", "gaussian"])
@pytest.mark.parametrize("n", [1, 31, 33, 1000, 20011, 300007])
def test_i8_hessian_against_oracle_and_dmma(family, n):
    model, (X, obs, off, link, x, pr), keep = i8_model(family, n, seed=n + 5, offset=(n == 20011))
    got = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_i8_kernel"), model.main_kernel()
    again = model.eval(x)
    assert got[0] == again[0] and np.array_equal(got[1], again[1]) and np.array_equal(got[2], again[2])
    want = O.model_eval_fused(family, obs, [X], [link], x, offsets=[off] if off is not None else None, priors=[pr])
    check(got, want)
    # against the DMMA path of the same library: same objective / gradient code, Hessian to ~1e-13
    model.set_option("hess_i8", 0)
    ref = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_kernel"), model.main_kernel()
    assert np.abs(got[2] - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
    assert abs(got[0] - ref[0]) <= 1e-13 * abs(ref[0])
    assert np.abs(got[1] - ref[1]).max() <= 1e-12 * np.abs(ref[1]).max()


def test_i8_hessian_badly_scaled_columns_entrywise():
    # with per-column scales every entry is accurate relative to sqrt(H_ii H_jj), not only to the largest entry
    model, (X, obs, off, link, x, pr), keep = i8_model("binomial", 50021, seed=77, scale_cols=True, prior=False)
    got = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_i8_kernel")
    want = O.model_eval_fused("binomial", obs, [X], [link], x)
    d = np.sqrt(np.diag(want[2]))
    assert np.abs((got[2] - want[2]) / np.outer(d, d)).max() <= 1e-11
    assert np.array_equal(got[2], got[2].T)


def test_i8_hessian_many_epochs_against_dmma():
    # more than 1024 blocks per CTA: the int32 accumulators are flushed several times
    import torch

    from anml_b200 import _native as N

    n = 24_000_017
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    X = torch.randn((n, 64), dtype=torch.float64, device="cuda", generator=g)
    X[:, 0] = 1.0
    beta = torch.randn(64, dtype=torch.float64, device="cuda", generator=g) * 0.1
    obs = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) < torch.sigmoid(X @ beta)).double()
    design = N.Design.wrap(X.data_ptr(), n, 64, 64, keepalive=X)
    model = N.NativeModel(n, "binomial", 1)
    model.set_param(0, design, N.LINK_CODES["Expit"])
    model.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
    x = (0.5 * beta).cpu().numpy()
    got = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_i8_kernel")
    model.set_option("hess_i8", 0)
    ref = model.eval(x)
    assert np.abs(got[2] - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
    assert np.array_equal(got[2], got[2].T)
    assert abs(got[0] - ref[0]) <= 1e-11 * abs(ref[0])
    # a wrapped matrix rewritten in place beyond the recorded column ranges is detected, not silently mis-evaluated
    # (coefficient 3 set to zero so that the row keeps an ordinary weight: a huge linear predictor would zero it)
    model.set_option("hess_i8", 1)
    x = x.copy()
    x[3] = 0.0
    X[5, 3] = 1.0e6
    torch.cuda.synchronize()
    with pytest.raises((RuntimeError, ValueError)):
        model.eval(x)
    got2 = model.eval(x)  # ranges taken again
    assert model.main_kernel().startswith("fused_hess_i8_kernel")
    model.set_option("hess_i8", 0)
    ref2 = model.eval(x)
    assert np.abs(got2[2] - ref2[2]).max() <= 1e-11 * np.abs(ref2[2]).max()


def test_i8_hessian_gaussian_with_obs_se():
    # sqrt(w) = 1 / obs_se: the bound comes from min(obs_se), taken when the model is first evaluated
    from anml_b200 import _native as N

    n = 40009
    X, obs, se, off, link, x = make_problem("gaussian", n, 64, 3, with_se=True)
    se = se * np.exp2(np.random.default_rng(4).integers(-6, 7, size=n))  # standard errors over four decades
    design = N.Design(n, 64)
    design.set_columns(0, X)
    model = N.NativeModel(n, "gaussian", 1)
    model.set_param(0, design, N.LINK_CODES["Identity"])
    model.set_obs(obs)
    model.set_obs_se(se)
    got = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_i8_kernel"), model.main_kernel()
    want = O.model_eval_fused("gaussian", obs, [X], [link], x, obs_se=se)
    check(got, want)
    # a new obs_se vector with a smaller minimum: the bound is taken again, not reused
    se2 = se * 2.0**-9
    model.set_obs_se(se2)
    got2 = model.eval(x)
    assert model.main_kernel().startswith("fused_hess_i8_kernel")
    check(got2, O.model_eval_fused("gaussian", obs, [X], [link], x, obs_se=se2))


def test_i8_default_row_minimum():
    # by default small problems stay on DMMA (the digit planes bound absolute, not relative, row errors: DESIGN.md 3.1a)
    from anml_b200 import _native as N

    for n, want_i8 in ((8191, False), (8192, True)):
        X, obs, se, off, link, x = make_problem("binomial", n, 64, 5)
        design = N.Design(n, 64)
        design.set_columns(0, X)
        model = N.NativeModel(n, "binomial", 1)
        model.set_param(0, design, N.LINK_CODES["Expit"])
        model.set_obs(obs)
        got = model.eval(x)
        assert model.main_kernel().startswith("fused_hess_i8_kernel") == want_i8, (n, model.main_kernel())
        check(got, O.model_eval_fused("binomial", obs, [X], [link], x))


def test_i8_not_used_outside_its_domain():
    # per-row weights, Poisson, p != 64: the DMMA kernel
    from anml_b200 import _native as N

    X, obs, se, off, link, x = make_problem("poisson", 5000, 64, 3)
    got = gpu_eval("poisson", X, obs, None, None, link, x)
    want = O.model_eval_fused("poisson", obs, [X], [link], x)
    check(got, want)
    model, prob, keep = i8_model("binomial", 4000, seed=9)
    model.set_weights(np.linspace(0.5, 1.5, 4000))
    model.eval(prob[4])
    assert model.main_kernel().startswith("fused_hess_kernel")


def meanvar64_model(n, seed, scale=1.0, offsets=False):
    from anml_b200 import _native as N

    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 64)); X[:, 0] = 1.0
    b0 = rng.normal(size=64) * 0.1
    b1 = rng.normal(size=64) * 0.05 * scale
    obs = X @ b0 + np.exp(0.5 * (X @ b1)) * rng.normal(size=n)
    off = [rng.normal(size=n) * 0.3, rng.normal(size=n) * 0.2] if offsets else [None, None]
    d = N.Design(n, 64); d.set_columns(0, X)
    bufs = [N.DeviceBuffer(0, o) if o is not None else None for o in off]
    model = N.NativeModel(n, "gaussian_meanvar", 2)
    model.set_param(0, d, N.LINK_CODES["Identity"], bufs[0])
    model.set_param(1, d, N.LINK_CODES["Exp"], bufs[1])
    model.set_obs(obs)
    model.set_gaussian_prior(0, np.zeros(64), np.full(64, 3.0))
    model.set_option("hess_i8", 2)
    x = np.concatenate([0.6 * b0, 0.6 * b1])
    pri = [{"direct": (np.zeros(64), np.full(64, 3.0))}, None]
    want = O.model_eval_fused("gaussian_meanvar", obs, [X, X], ["identity", "exp"], x, offsets=off if offsets else None, priors=pri)
    return model, x, want, (d, bufs)


@pytest.mark.parametrize("n,offsets", [(1, False), (33, False), (5000, True), (200_003, False)])
def test_i8_mean_variance_blocks(n, offsets):
    # BASELINE config #4 shape: block (1,1) (w11 >= 0, bound = its exact maximum) on the INT8 tensor cores, blocks (0,0) and
    # the signed cross block (0,1) on DMMA
    model, x, want, keep = meanvar64_model(n, seed=n + 3, offsets=offsets)
    got = model.eval(x)
    ks = model.last_kernels()
    assert [k.split("<")[0] for k in ks] == ["fused_hess_kernel", "fused_hess_i8_kernel", "fused_hess_kernel"], ks
    check(got, want)
    again = model.eval(x)
    assert got[0] == again[0] and np.array_equal(got[1], again[1]) and np.array_equal(got[2], again[2])
    model.set_option("hess_i8", 0)
    ref = model.eval(x)
    assert all(k.startswith("fused_hess_kernel") for k in model.last_kernels())
    assert np.abs(got[2] - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
    assert np.abs(got[1] - ref[1]).max() <= 1e-12 * np.abs(ref[1]).max()
    assert abs(got[0] - ref[0]) <= 1e-12 * abs(ref[0])


def test_i8_mean_variance_wide_range_of_weights():
    # variance coefficients 8 times larger: w11 = res^2 / (2 v) spans many decades; the digit scale follows its exact maximum
    model, x, want, keep = meanvar64_model(30011, seed=12, scale=8.0)
    got = model.eval(x)
    assert "fused_hess_i8_kernel<1>" in model.last_kernels(), model.last_kernels()
    check(got, want)
    model.set_option("hess_i8", 0)
    ref = model.eval(x)
    d = np.sqrt(np.abs(np.diag(ref[2])))
    assert np.abs((got[2] - ref[2]) / np.outer(d, d)).max() <= 1e-10
