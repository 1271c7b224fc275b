"""GPU tests of the anml-compatible Python API (Parameter / SplineVariable /
Model) against the golden vectors recorded from the reference and against the
NumPy oracle.  Tolerances: BASELINE.json (obj/grad 1e-10, Hessian 1e-9, fitted
coefficients 1e-8)."""
import numpy as np
import pandas as pd
import pytest

from oracle import anml_oracle as O

pytestmark = pytest.mark.gpu

LINKS = {"identity": "Identity", "exp": "Exp", "log": "Log", "expit": "Expit", "logit": "Logit"}


def frame_from(X, extra=None):
    cols = {f"cov{j}": X[:, j + 1] for j in range(X.shape[1] - 1)}
    cols.update(extra or {})
    return pd.DataFrame(cols)


def make_parameter(p, link="identity", offset=None, priors=None, var_priors=None):
    from anml_b200.data import Component
    from anml_b200 import parameter as P
    from anml_b200.variable import Variable

    variables = [Variable(Component("intercept", default_value=1.0), priors=var_priors)]
    variables += [Variable(f"cov{j}", priors=var_priors) for j in range(p - 1)]
    return P.Parameter(variables, transform=getattr(P, LINKS[link])(), offset=offset, priors=priors)


# ---- the reference's own known-answer tests for Parameter (tests/parameter/test_main.py:94-153)
def test_parameter_known_answers():
    from anml_b200.parameter import Parameter
    from anml_b200.variable import Variable

    df = pd.DataFrame({"cov1": np.random.randn(5), "cov2": np.random.randn(5)})
    p = Parameter([Variable("cov1"), Variable("cov2")])
    x = np.ones(2)
    assert np.allclose(p.get_params(x, df, order=0), df.sum(axis=1))
    assert np.allclose(p.get_params(x, df, order=1), df.values)
    assert np.allclose(p.get_params(x, df, order=2), np.zeros((5, 2, 2)))
    assert all(v.component.value is not None for v in p.variables)
    assert p.design_mat is not None and p.design_mat.flags.c_contiguous
    np.testing.assert_array_equal(p.design_mat, df.values)
    for cat in ("direct", "linear"):
        for typ in ("GaussianPrior", "UniformPrior"):
            assert typ in p.prior_dict[cat]
    assert p.prior_dict["direct"]["GaussianPrior"].shape == (2, 2)
    assert p.prior_dict["linear"]["UniformPrior"].mat.shape == (0, 2)
    assert p.prior_objective(x) == 0.0


@pytest.mark.parametrize("case", ["p5_identity", "p5_exp_off", "p7_expit", "p17_expit_off", "p64_expit"])
def test_get_params_matches_reference_golden(golden, case):
    X = golden[f"params/{case}/X"]
    beta = golden[f"params/{case}/beta"]
    link = str(golden[f"params/{case}/link"])
    key = f"params/{case}/offset"
    extra = {"off": golden[key]} if key in golden.files else None
    par = make_parameter(X.shape[1], link, offset="off" if extra else None)
    par.attach(frame_from(X, extra))
    np.testing.assert_array_equal(par.design_mat, X)
    for order in (0, 1, 2):
        want = golden[f"params/{case}/order{order}"]
        got = par.get_params(beta, order=order)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13 * np.abs(want).max())


def test_get_params_domain_and_order_errors():
    X = np.abs(np.random.default_rng(0).normal(size=(20, 3))) + 0.1
    par = make_parameter(3, "log")
    par.attach(frame_from(X))
    with pytest.raises(ValueError, match="positive"):
        par.get_params(-np.ones(3))
    with pytest.raises(ValueError, match="Order"):
        par.get_params(np.ones(3), order=3)
    assert np.all(np.isfinite(par.get_params(np.ones(3), order=2)))


@pytest.mark.parametrize("case,fam", [("gauss_p4", "gaussian"), ("binom_p6", "binomial"), ("binom_p64", "binomial"),
                                      ("poisson_p5_off", "poisson"), ("gauss_exp_p3", "gaussian"), ("meanvar_p5", "gaussian_meanvar")])
def test_model_matches_reference_composition(golden, case, fam):
    from anml_b200.data import DataExample
    from anml_b200.model import Model
    from anml_b200.prior import GaussianPrior

    X = golden[f"model/{case}/X"]
    links = str(golden[f"model/{case}/links"]).split(",")
    key = f"model/{case}/offset"
    extra = {"obs": golden[f"model/{case}/obs"], "se": golden[f"model/{case}/se"]}
    if key in golden.files:
        extra["off"] = golden[key]
    sd = float(golden[f"model/{case}/prior_sd"])
    vp = [GaussianPrior(mean=0.0, sd=sd)] if sd > 0 else None
    pars = [make_parameter(X.shape[1], l, offset="off" if "off" in extra else None, var_priors=vp) for l in links]
    model = Model(fam, DataExample("obs", "se"), pars)
    model.attach(frame_from(X, extra))
    x = golden[f"model/{case}/x"]
    obj, grad, hess = model.objective(x), model.gradient(x), model.hessian(x)
    assert model.num_evaluations == 1  # the three callbacks share one device evaluation
    ow, gw, hw = float(golden[f"model/{case}/obj"]), golden[f"model/{case}/grad"], golden[f"model/{case}/hess"]
    assert abs(obj - ow) <= 1e-10 * abs(ow)
    assert np.abs(grad - gw).max() <= 1e-10 * np.abs(gw).max()
    assert np.abs(hess - hw).max() <= 1e-9 * np.abs(hw).max()
    if fam == "gaussian_meanvar":
        assert pars[0].device_design is pars[1].device_design  # same columns -> one matrix in HBM


def test_config1_trust_constr_fit_matches_reference(golden):
    """BASELINE config #1: Gaussian, intercept + 3 covariates, 10k rows; fitted
    coefficients within 1e-8 of the fit driven by the reference's callbacks."""
    from anml_b200.data import DataExample
    from anml_b200.model import Model

    n, p = 10_000, 4
    r = np.random.default_rng(int(golden["c1/seed"]))
    Xc = r.normal(size=(n, p - 1))
    beta_true = r.normal(size=p) * 0.1
    np.testing.assert_array_equal(beta_true, golden["c1/beta_true"])
    obs = beta_true[0] + Xc @ beta_true[1:] + 0.1 * r.normal(size=n)
    df = pd.DataFrame({f"cov{j}": Xc[:, j] for j in range(p - 1)})
    df["obs"] = obs
    model = Model("gaussian", DataExample("obs", "obs_se"), [make_parameter(p)])
    model.attach(df)
    assert abs(model.objective(np.zeros(p)) - float(golden["c1/obj_at_zero"])) <= 1e-10 * float(golden["c1/obj_at_zero"])
    res = model.fit(np.zeros(p))
    assert np.abs(res.x - golden["c1/coef"]).max() <= 1e-8
    assert np.abs(res.x - beta_true).max() < 0.01


def test_model_bounds_and_constraints_reach_the_solver():
    from anml_b200.data import Component, DataExample
    from anml_b200.model import Model
    from anml_b200.parameter import Parameter
    from anml_b200.prior import UniformPrior
    from anml_b200.variable import Variable

    r = np.random.default_rng(4)
    n = 2000
    df = pd.DataFrame({"a": r.normal(size=n), "b": r.normal(size=n)})
    df["obs"] = 1.0 + 2.0 * df["a"] - 1.0 * df["b"] + 0.01 * r.normal(size=n)
    vs = [Variable(Component("intercept", default_value=1.0)), Variable("a", priors=[UniformPrior(lb=0.0, ub=1.5)]), Variable("b")]
    par = Parameter(vs, priors=[UniformPrior(lb=[-0.5], ub=[0.5], mat=np.array([[0.0, 1.0, 1.0]]))])
    model = Model("gaussian", DataExample("obs", "obs_se"), [par])
    model.attach(df)
    b = model.bounds()
    assert b.ub[1] == 1.5 and np.isinf(b.ub[0])
    assert len(model.constraints()) == 1
    res = model.fit(options={"gtol": 1e-8})
    assert res.x[1] <= 1.5 + 1e-6  # the bound binds (unconstrained optimum is 2.0)
    assert abs(res.x[1] + res.x[2]) <= 0.5 + 1e-6


def test_model_requires_attach_and_validates():
    from anml_b200.data import DataExample
    from anml_b200.model import Model

    with pytest.raises(ValueError):
        Model("binomial", DataExample("obs", "se"), [make_parameter(2), make_parameter(2)])
    with pytest.raises(ValueError):
        Model("nope", DataExample("obs", "se"), [make_parameter(2)])
    m = Model("binomial", DataExample("obs", "se"), [make_parameter(2)])
    with pytest.raises(ValueError, match="attach"):
        m.objective(np.zeros(2))


def test_spline_variable_in_a_parameter_and_smoothing_prior():
    """BASELINE config #3 in miniature: SplineGetter (rel_domain knots, degree 3)
    + SplinePriorGetter(GaussianPrior, order=2) smoothing, Gaussian likelihood."""
    from anml_b200.data import DataExample
    from anml_b200.getter import SplineGetter, SplinePriorGetter
    from anml_b200.model import Model
    from anml_b200.parameter import Parameter
    from anml_b200.prior import GaussianPrior, Prior
    from anml_b200.variable import SplineVariable
    from anml_b200.xspline import XSpline

    r = np.random.default_rng(20261019 + 3)
    n = 30_000
    xs = r.uniform(0, 1, size=n)
    df = pd.DataFrame({"x": xs, "obs": np.sin(2 * np.pi * xs) + 0.1 * r.normal(size=n)})
    sv = SplineVariable("x", SplineGetter(np.linspace(0, 1, 20), degree=3, knots_type="rel_domain"),
                        priors=[SplinePriorGetter(GaussianPrior(0.0, 50.0), size=100, order=2)])
    assert sv.size == 22
    par = Parameter([sv])
    model = Model("gaussian", DataExample("obs", "obs_se"), [par])
    model.attach(df)
    assert isinstance(sv.spline, XSpline) and isinstance(sv.priors[0], Prior) and sv.priors[0].mat.shape == (100, 22)
    knots = O.spline_knots(np.linspace(0, 1, 20), "rel_domain", xs)
    np.testing.assert_array_equal(sv.spline.knots, knots)
    D = par.design_mat
    assert D.shape == (n, 22)
    np.testing.assert_array_equal(D, O.spline_design(knots, 3, xs))  # bit-exact basis inside the knot span
    pts = O.spline_prior_points(knots, 100)
    M = O.spline_design(knots, 3, pts, order=2)
    np.testing.assert_allclose(sv.priors[0].mat, M, rtol=0, atol=1e-12 * np.abs(M).max())
    lin = par.prior_dict["linear"]["GaussianPrior"]
    x = r.normal(size=22) * 0.3
    want = O.model_eval_fused("gaussian", df["obs"].to_numpy(), [D], ["identity"], x,
                              priors=[{"linear": (lin.mean, lin.sd, lin.mat)}])
    got = model.evaluate(x)
    assert abs(got[0] - want[0]) <= 1e-10 * abs(want[0])
    assert np.abs(got[1] - want[1]).max() <= 1e-10 * np.abs(want[1]).max()
    assert np.abs(got[2] - want[2]).max() <= 1e-9 * np.abs(want[2]).max()
    # host prior path of the Parameter agrees with the device prior epilogue
    nop = O.model_eval_fused("gaussian", df["obs"].to_numpy(), [D], ["identity"], x)
    assert abs((got[0] - nop[0]) - par.prior_objective(x)) <= 1e-9 * abs(got[0])
    res = model.fit(options={"gtol": 1e-8, "maxiter": 200})
    assert np.abs(model.gradient(res.x)).max() < 1e-5  # a stationary point of the penalised likelihood
    fitted = D @ res.x
    assert np.sqrt(np.mean((fitted - np.sin(2 * np.pi * xs)) ** 2)) < 0.05


def test_sharded_model_single_rank_uses_the_async_device_path(golden):
    """ShardedModel + NativeLocalEvaluator on one rank: eval_async into a torch CUDA
    tensor on torch's current stream, pinned read-back, solver loop."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from anml_b200.data import DataExample
    from anml_b200.distributed import NativeLocalEvaluator, ShardedModel
    from anml_b200.model import Model
    from anml_b200.prior import GaussianPrior

    case = "binom_p64"
    X = golden[f"model/{case}/X"]
    extra = {"obs": golden[f"model/{case}/obs"], "se": golden[f"model/{case}/se"]}
    par = make_parameter(X.shape[1], "expit", var_priors=[GaussianPrior(mean=0.0, sd=1.0)])
    model = Model("binomial", DataExample("obs", "se"), [par])
    model.attach(frame_from(X, extra))
    sharded = ShardedModel(NativeLocalEvaluator(model))
    x = golden[f"model/{case}/x"]
    with torch.cuda.stream(torch.cuda.Stream()):  # a non-default torch stream
        o, g, H = sharded.evaluate(x)
    o2, g2, H2 = sharded.evaluate(x * 1.0)  # cache hit on equal bytes
    assert sharded.num_evaluations == 1 and o2 == o
    assert abs(o - float(golden[f"model/{case}/obj"])) <= 1e-10 * abs(o)
    assert np.abs(g - golden[f"model/{case}/grad"]).max() <= 1e-10 * np.abs(g).max()
    assert np.abs(H - golden[f"model/{case}/hess"]).max() <= 1e-9 * np.abs(H).max()
    direct = model.evaluate(x)
    assert direct[0] == o and np.array_equal(direct[1], g) and np.array_equal(direct[2], H)
    res = sharded.fit(np.zeros(x.size), options={"gtol": 1e-8, "maxiter": 50})
    assert np.abs(sharded.gradient(res.x)).max() < 1e-6


def test_attach_validators_run_on_the_device():
    """NoNans (default for string components) and Positive are checked on the uploaded
    columns; same ValueError text as the host validators of the reference."""
    from anml_b200.data import Component, NoNans, Positive
    from anml_b200.parameter import Parameter
    from anml_b200.variable import Variable

    df = pd.DataFrame({"a": [1.0, 2.0, np.nan, 4.0], "b": [1.0, -2.0, 3.0, 4.0], "c": [1.0, 2.0, 3.0, 4.0]})
    with pytest.raises(ValueError, match="Column 'a' contains nans."):
        Parameter([Variable("c"), Variable("a")]).attach(df)
    with pytest.raises(ValueError, match="Column 'b' contains nonpositive numbers."):
        Parameter([Variable(Component("b", [NoNans(), Positive()]))]).attach(df)
    p = Parameter([Variable("b"), Variable("c"), Variable(Component("a"))])  # no validator on 'a': NaN passes through
    p.attach(df)
    assert np.isnan(p.design_mat[2, 2]) and p.design_mat[1, 0] == -2.0


def test_model_picks_up_a_weights_component():
    from anml_b200.data import Component, DataPrototype, NoNans
    from anml_b200.model import Model

    rng = np.random.default_rng(8)
    n, p = 2000, 4
    X = rng.normal(size=(n, p)); X[:, 0] = 1.0
    obs = X @ (rng.normal(size=p) * 0.3) + rng.normal(size=n)
    wts = rng.uniform(0, 1, size=n)
    df = frame_from(X, {"obs": obs, "wt": wts})
    data = DataPrototype({"obs": Component("obs", [NoNans()]), "weights": Component("wt", [NoNans()])})
    model = Model("gaussian", data, [make_parameter(p)])
    model.attach(df)
    x = rng.normal(size=p)
    want = O.model_eval_fused("gaussian", obs, [X], ["identity"], x, weights=wts)
    got = model.evaluate(x)
    assert abs(got[0] - want[0]) <= 1e-10 * abs(want[0])
    assert np.abs(got[2] - want[2]).max() <= 1e-9 * np.abs(want[2]).max()
    res = model.fit()
    wls = np.linalg.solve((X.T * wts) @ X, (X.T * wts) @ obs)  # weighted least squares in closed form
    assert np.abs(res.x - wls).max() < 1e-8


def test_empty_frame_gives_prior_only_terms():
    """Zero rows: likelihood sums are empty, only the Gaussian prior remains (NumPy semantics of
    the reference: sums over empty arrays are 0)."""
    from anml_b200.data import DataExample
    from anml_b200.model import Model
    from anml_b200.prior import GaussianPrior

    p = 5
    df = pd.DataFrame({**{f"cov{j}": np.empty(0) for j in range(p - 1)}, "obs": np.empty(0)})
    par = make_parameter(p, "expit", var_priors=[GaussianPrior(mean=0.5, sd=2.0)])
    model = Model("binomial", DataExample("obs", "obs_se"), [par])
    model.attach(df)
    assert par.design_mat.shape == (0, p)
    assert par.get_params(np.ones(p)).shape == (0,)
    assert par.get_params(np.ones(p), order=2).shape == (0, p, p)
    x = np.linspace(-1, 1, p)
    o, g, H = model.evaluate(x)
    assert abs(o - par.prior_objective(x)) < 1e-14
    np.testing.assert_allclose(g, par.prior_gradient(x), atol=1e-15)
    np.testing.assert_allclose(H, par.prior_hessian(x), atol=1e-15)


def test_prediction_on_a_new_frame_reuses_spline_knots():
    """get_params(x, df_new): knots come from the first attached frame only
    (variable/spline.py:103-107); the new frame is evaluated with them."""
    from anml_b200.getter import SplineGetter
    from anml_b200.parameter import Exp, Parameter
    from anml_b200.variable import SplineVariable, Variable

    r = np.random.default_rng(3)
    df1 = pd.DataFrame({"x": r.uniform(0, 1, 500), "z": r.normal(size=500)})
    sv = SplineVariable("x", SplineGetter(np.linspace(0, 1, 6), degree=3, knots_type="rel_domain"))
    par = Parameter([sv, Variable("z")], transform=Exp())
    beta = r.normal(size=9) * 0.2
    par.get_params(beta, df1)
    knots = sv.spline.knots.copy()
    df2 = pd.DataFrame({"x": r.uniform(0.1, 0.9, 50), "z": r.normal(size=50)})
    pred = par.get_params(beta, df2, order=0)
    np.testing.assert_array_equal(sv.spline.knots, knots)
    X2 = np.column_stack([O.spline_design(knots, 3, df2["x"].to_numpy()), df2["z"].to_numpy()])
    np.testing.assert_allclose(pred, np.exp(X2 @ beta), rtol=1e-13)
    np.testing.assert_array_equal(par.design_mat, X2)


@pytest.mark.gpu
@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_upload_paths_round_trip_exactly(threads, monkeypatch):
    """Ingest (upload.cuh): dense block, contiguous columns, a strided column view and a
    device vector, over several chunks and host threads, all land bit-exactly."""
    from anml_b200 import _native as N

    monkeypatch.setenv("ANML_B200_UPLOAD_THREADS", threads)
    rng = np.random.default_rng(5)
    n, p = 1_500_007, 6
    A = rng.standard_normal((n, p))
    d = N.Design(n, p)
    d.set_columns(0, A)  # dense: DMA straight into place (ld == p because p is even)
    assert np.array_equal(d.download(), A)
    B = rng.standard_normal((n, p))
    F = np.asfortranarray(B)
    d.set_columns(1, F[:, 1])       # contiguous column -> scatter
    d.set_columns(4, B[:, 4:5])     # strided view of a C-ordered matrix
    d.set_columns(2, B[:, 2:4])     # two columns of a wider block (made contiguous by the wrapper)
    want = A.copy()
    want[:, 1] = B[:, 1]
    want[:, 4] = B[:, 4]
    want[:, 2:4] = B[:, 2:4]
    assert np.array_equal(d.download(), want)
    v = rng.standard_normal(2_000_003)
    assert np.array_equal(N.DeviceBuffer(0, v).download(), v)
    # odd width: ld is padded to even, so the dense block also takes the scatter route
    d3 = N.Design(100_001, 3)
    C3 = rng.standard_normal((100_001, 3))
    d3.set_columns(0, C3)
    assert np.array_equal(d3.download(), C3)


@pytest.mark.gpu
def test_attach_from_pyarrow_table_equals_attach_from_pandas():
    pa = pytest.importorskip("pyarrow")
    import pandas as pd

    from anml_b200.data import Component, DataExample
    from anml_b200.model import Model
    from anml_b200.parameter import Expit, Parameter
    from anml_b200.variable import Variable

    rng = np.random.default_rng(21)
    n = 30_011
    cols = {f"c{j}": rng.normal(size=n) for j in range(1, 6)}
    cols["obs"] = (rng.uniform(size=n) < 0.4).astype(float)
    x = rng.normal(size=6) * 0.2

    def build():
        par = Parameter([Variable(Component("intercept", default_value=1.0))] + [Variable(f"c{j}") for j in range(1, 6)], transform=Expit())
        return par, Model("binomial", DataExample("obs", "obs_se"), [par])

    par_pd, model_pd = build()
    model_pd.attach(pd.DataFrame(cols))
    par_pa, model_pa = build()
    table = pa.table(cols)
    model_pa.attach(table)
    assert table.column_names == list(cols)  # the immutable table is left alone
    np.testing.assert_array_equal(par_pa.design_mat, par_pd.design_mat)
    a, b = model_pd.evaluate(x), model_pa.evaluate(x)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    np.testing.assert_array_equal(par_pa.get_params(x, table), par_pd.get_params(x))


@pytest.mark.gpu
@pytest.mark.parametrize("link", ["log", "logit", "exp", "expit", "identity"])
def test_device_link_values_match_the_reference_grid(golden, link):
    """SmoothMapping values of every order computed ON THE DEVICE (Parameter.get_params through a one-column
    design whose linear predictor is the grid itself) against the imported reference's outputs on the same grid
    (tests/golden: link/<name>/order{0,1,2}; smoothmapping.py:59-156).  Rows a4-a7 of SURVEY.md section 8."""
    from anml_b200 import parameter as P
    from anml_b200.variable import Variable

    grid = np.ascontiguousarray(golden[f"link/{link}/x"], dtype=np.float64)
    par = P.Parameter([Variable("v")], transform=getattr(P, LINKS[link])())
    par.attach(pd.DataFrame({"v": grid}))
    for order in (0, 1, 2):
        want = golden[f"link/{link}/order{order}"]
        got = par.get_params(np.ones(1), order=order)
        if order == 1:
            got, want = got[:, 0], want * grid             # J = f'(y) * x, one column
        elif order == 2:
            got, want = got[:, 0, 0], want * (grid * grid)  # T = f''(y) * x x^T
        # libdevice and NumPy's libm each round exp / log to within an ulp or two; products add a few more
        np.testing.assert_allclose(got, want, rtol=4e-15, atol=1e-300)
