"""Pin the oracle: replay every golden vector recorded from the real reference
(tests/golden/make_golden.py) through oracle/anml_oracle.py.  CPU only."""
import numpy as np
import pytest

from oracle import anml_oracle as O

LINKS = ["identity", "exp", "log", "expit", "logit"]
PARAM_CASES = ["p5_identity", "p5_exp_off", "p7_expit", "p17_expit_off", "p64_expit"]
MODEL_CASES = ["gauss_p4", "binom_p6", "binom_p64", "poisson_p5_off", "meanvar_p5", "gauss_exp_p3"]


@pytest.mark.parametrize("link", LINKS)
@pytest.mark.parametrize("order", [0, 1, 2])
def test_links_bit_exact(golden, link, order):
    x = golden[f"link/{link}/x"]
    with np.errstate(all="ignore"):
        got = O.link_eval(link, x.copy(), order)
    np.testing.assert_array_equal(got, golden[f"link/{link}/order{order}"])


@pytest.mark.parametrize("link", LINKS)
def test_link_order_error(link):
    with pytest.raises(ValueError):
        O.link_eval(link, np.array([0.5]), order=3)


@pytest.mark.parametrize("case", PARAM_CASES)
@pytest.mark.parametrize("order", [0, 1, 2])
def test_get_params_bit_exact(golden, case, order):
    X = golden[f"params/{case}/X"]
    beta = golden[f"params/{case}/beta"]
    link = str(golden[f"params/{case}/link"])
    key = f"params/{case}/offset"
    offset = golden[key] if key in golden.files else None
    got = O.get_params(X, beta, link, offset, order)
    np.testing.assert_array_equal(got, golden[f"params/{case}/order{order}"])


def test_prior_direct_and_linear(golden):
    x = golden["prior/direct/x"]
    o, g, h = O.gaussian_prior_terms(x, np.array([0.1, -0.2, 0.3]), np.array([1.0, 0.5, 2.0]))
    assert o == float(golden["prior/direct/obj"])
    np.testing.assert_array_equal(g, golden["prior/direct/grad"])
    np.testing.assert_array_equal(h, golden["prior/direct/hess"])
    o, g, h = O.gaussian_prior_terms(x, golden["prior/linear/mean"], golden["prior/linear/sd"], golden["prior/linear/mat"])
    assert o == float(golden["prior/linear/obj"])
    np.testing.assert_array_equal(g, golden["prior/linear/grad"])
    np.testing.assert_array_equal(h, golden["prior/linear/hess"])


def test_parameter_prior_sum(golden):
    x = golden["prior/param/x"]
    d = golden["prior/param/direct/GaussianPrior/params"]
    l = golden["prior/param/linear/GaussianPrior/params"]
    m = golden["prior/param/linear/GaussianPrior/mat"]
    o, g, h = O.parameter_prior_terms(x, direct=(d[0], d[1]), linear=(l[0], l[1], m))
    np.testing.assert_allclose(o, float(golden["prior/param/obj"]), rtol=1e-14)
    np.testing.assert_allclose(g, golden["prior/param/grad"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(h, golden["prior/param/hess"], rtol=1e-13, atol=1e-15)


def _model_inputs(golden, case):
    fam = str(golden[f"model/{case}/family"])
    links = str(golden[f"model/{case}/links"]).split(",")
    X = golden[f"model/{case}/X"]
    key = f"model/{case}/offset"
    off = golden[key] if key in golden.files else None
    sd = float(golden[f"model/{case}/prior_sd"])
    p = X.shape[1]
    priors = [{"direct": (np.zeros(p), np.full(p, sd))} if sd > 0 else None for _ in links]
    kw = dict(offsets=[off] * len(links), obs_se=golden[f"model/{case}/se"], priors=priors)
    return fam, golden[f"model/{case}/obs"], [X] * len(links), links, golden[f"model/{case}/x"], kw


@pytest.mark.parametrize("case", MODEL_CASES)
@pytest.mark.parametrize("route", ["literal", "fused"])
def test_model_layer_matches_reference_composition(golden, case, route):
    fam, obs, Xs, links, x, kw = _model_inputs(golden, case)
    fn = O.model_eval_literal if route == "literal" else O.model_eval_fused
    if route == "fused":
        kw["chunk"] = 64
    o, g, h = fn(fam, obs, Xs, links, x, **kw)
    np.testing.assert_allclose(o, float(golden[f"model/{case}/obj"]), rtol=1e-12)
    scale_g = np.abs(golden[f"model/{case}/grad"]).max()
    scale_h = np.abs(golden[f"model/{case}/hess"]).max()
    np.testing.assert_allclose(g, golden[f"model/{case}/grad"], rtol=0, atol=1e-12 * scale_g)
    np.testing.assert_allclose(h, golden[f"model/{case}/hess"], rtol=0, atol=1e-12 * scale_h)


def test_spline_against_scipy():
    from scipy.interpolate import BSpline

    knots = np.linspace(0.0, 1.0, 20)
    d = 3
    t = np.concatenate([np.repeat(knots[0], d), knots, np.repeat(knots[-1], d)])
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0, 1, 2000), knots])
    D, idx = O.spline_design(knots, d, x, return_index=True)
    assert D.shape == (x.size, 22)
    np.testing.assert_allclose(D, BSpline.design_matrix(x, t, d).toarray(), rtol=0, atol=1e-15)
    np.testing.assert_allclose(D.sum(axis=1), 1.0, atol=1e-14)
    assert (np.count_nonzero(D, axis=1) <= 4).all()
    # interval indices [0, 18, 5] for x = [k0, k_last, knots[5]]
    np.testing.assert_array_equal(O.spline_interval_index(knots, [knots[0], knots[-1], knots[5]]), [0, 18, 5])
    nb = D.shape[1]
    inner = x < 1.0
    for m in (1, 2):
        Dm = O.spline_design(knots, d, x, order=m)
        ref = np.column_stack([BSpline(t, np.eye(nb)[j], d).derivative(m)(x) for j in range(nb)])
        np.testing.assert_allclose(Dm[inner], ref[inner], rtol=0, atol=1e-12 * np.abs(ref).max())


def test_spline_extrapolation_is_taylor():
    knots = np.array([0.0, 0.3, 0.7, 1.5])
    d = 3
    edge = np.array([0.0, 1.5])
    xe = np.array([-0.4, 2.0])
    dx = (xe - edge)[:, None]
    D0, D1, D2 = (O.spline_design(knots, d, edge, order=m) for m in (0, 1, 2))
    np.testing.assert_allclose(O.spline_design(knots, d, xe, ldegree=0, rdegree=0), D0, atol=1e-15)
    np.testing.assert_allclose(O.spline_design(knots, d, xe, ldegree=1, rdegree=1), D0 + D1 * dx, atol=1e-13)
    np.testing.assert_allclose(O.spline_design(knots, d, xe, ldegree=2, rdegree=2), D0 + D1 * dx + 0.5 * D2 * dx**2, atol=1e-12)
    np.testing.assert_allclose(O.spline_design(knots, d, xe, order=1, ldegree=2, rdegree=2), D1 + D2 * dx, atol=1e-12)


# ---- spline rows against the REAL xspline: consumed when tests/golden/make_spline_golden.py has been run
# where xspline is importable (it is not in this image: the restatement stays "parity unpinned" until then)
def _xspline_golden():
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xspline_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/xspline_golden.npz not recorded (xspline is not importable here): spline parity unpinned")
    return np.load(path, allow_pickle=False)


def test_oracle_spline_basis_against_recorded_xspline():
    g = _xspline_golden()
    for name in [str(c) for c in g["cases"]]:
        knots, x = g[f"{name}/knots"], g[f"{name}/x"]
        degree, ldeg, rdeg = (int(v) for v in g[f"{name}/params"])
        for order in range(degree + 1):
            want = g[f"{name}/order{order}"]
            got = O.spline_design(knots, degree, x, order=order, ldegree=ldeg, rdegree=rdeg)
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(want).max()), err_msg=f"{name} order {order}")


def test_oracle_knot_rules_against_recorded_reference():
    g = _xspline_golden()
    data = g["getter/data"]
    for kt in ("abs", "rel_domain", "rel_freq"):
        np.testing.assert_array_equal(O.spline_knots(np.linspace(0.0, 1.0, 7), kt, data), g[f"getter/{kt}/knots"])
