"""Host-side API behaviour (no GPU): constructors, validation, prior assembly,
smooth mappings.  Mirrors what the reference's unit tests pin
(tests/{data,prior,variable,parameter}/ of the reference)."""
from functools import partial

import numpy as np
import pandas as pd
import pytest

from anml_b200.data import Component, DataExample, DataPrototype, NoNans, Positive, Unique
from anml_b200.getter import SplineGetter, SplinePriorGetter
from anml_b200.parameter import Exp, Expit, Identity, Log, Logit, Parameter
from anml_b200.prior import GaussianPrior, Prior, UniformPrior, filter_priors, get_prior_type
from anml_b200.variable import SplineVariable, Variable
from anml_b200.xspline import XSpline


def complex_step(fun, x, out_shape=(), eps=1e-10):
    c = x + 0j
    g = np.zeros((*out_shape, *x.shape))
    if not out_shape:
        for i in np.ndindex(x.shape):
            c[i] += eps * 1j
            g[i] = fun(c).imag / eps
            c[i] -= eps * 1j
        return g
    for j in np.ndindex(out_shape):
        for i in np.ndindex(x.shape):
            c[i] += eps * 1j
            g[j][i] = fun(c)[j].imag / eps
            c[i] -= eps * 1j
    return g


# ---- data ---------------------------------------------------------------------
def test_component_attach_default_and_errors():
    df = pd.DataFrame({"a": [1.0, 2.0]})
    c = Component("a", [NoNans()])
    assert c.value is None
    c.attach(df)
    assert np.array_equal(c.value, [1.0, 2.0])
    c.clear()
    assert c.value is None
    with pytest.raises(KeyError):
        Component("missing").attach(df)
    filled = Component("intercept", default_value=1.0)
    filled.attach(df)
    assert "intercept" in df and np.all(filled.value == 1.0)  # the caller's frame gains the column
    with pytest.raises(TypeError):
        Component(1)
    with pytest.raises(TypeError):
        Component("a", validators=[1])
    with pytest.raises(TypeError):
        Component("a", validators=3)
    assert repr(c) == "Component(key=a)"


def test_validators():
    with pytest.raises(ValueError):
        NoNans()("k", np.array([1.0, np.nan]))
    with pytest.raises(ValueError):
        Positive()("k", np.array([1.0, 0.0]))
    with pytest.raises(ValueError):
        Unique()("k", np.array([1.0, 1.0]))
    NoNans()("k", np.ones(3)); Positive()("k", np.ones(3)); Unique()("k", np.arange(3.0))
    assert repr(NoNans()) == "NoNans()"


def test_prototype_and_example():
    d = DataExample("obs", "obs_se")
    df = pd.DataFrame({"obs": [0.1, 0.2]})
    d.attach(df)
    assert np.all(d.obs_se.value == 1.0) and np.array_equal(d.obs.value, [0.1, 0.2])
    d.clear()
    assert d.obs.value is None
    with pytest.raises(ValueError):
        DataExample("obs", "se").attach(pd.DataFrame({"obs": [1.0], "se": [-1.0]}))
    with pytest.raises(TypeError):
        DataPrototype({1: Component("a")})
    with pytest.raises(TypeError):
        DataPrototype({"a": "not a component"})


# ---- priors -------------------------------------------------------------------
@pytest.mark.parametrize("params", [[1.0, np.ones(2)], [np.ones(2), np.ones(2)]])
def test_prior_param_broadcast(params):
    assert Prior(params).params.shape == (2, 2)


def test_prior_shapes_and_mat_rules():
    with pytest.raises(ValueError):
        Prior([np.ones(3), np.ones(2)])
    assert Prior([1.0, np.ones(3)]).shape == (3, 3)
    assert Prior([1.0, np.ones(3)], np.ones((3, 2))).shape == (3, 2)
    for bad in ([], np.ones((2, 3))):
        with pytest.raises(ValueError):
            Prior([1.0, np.ones(3)], bad)
    assert Prior([[], []]).params.shape == (2, 0)
    p = Prior([0.0, 1.0], np.ones((4, 2)))  # single-column params repeat to mat's rows
    assert p.params.shape == (2, 4)
    for m in ("objective", "gradient", "hessian"):
        with pytest.raises(NotImplementedError):
            getattr(UniformPrior(0.0, 5.0), m)(None)


def test_gaussian_prior_values_and_derivatives():
    mean, sd = 2.0, 0.5
    for x in (np.array([1.0]), np.array([2.0])):
        assert GaussianPrior(mean, sd).objective(x) == 0.5 * np.sum(((x - mean) / sd) ** 2)
        for mat in (None, np.ones((2, 1))):
            pr = GaussianPrior(mean, sd, mat=mat)
            assert np.allclose(pr.gradient(x), complex_step(pr.objective, x))
            assert np.allclose(pr.hessian(x), complex_step(pr.gradient, x, out_shape=(1,)))
    for bad in (-1.0, 0.0):
        with pytest.raises(ValueError):
            GaussianPrior(1.0, bad)
    with pytest.raises(ValueError):
        UniformPrior(1.0, -1.0)
    assert np.isinf(GaussianPrior.default_params[1, 0]) and UniformPrior.default_params[0, 0] == -np.inf


def test_prior_utils():
    assert get_prior_type("GaussianPrior") is GaussianPrior
    pri = [GaussianPrior(0.0, 1.0), GaussianPrior(0.0, 1.0, mat=np.ones((1, 1))), UniformPrior(0.0, 1.0)]
    assert len(filter_priors(pri, "GaussianPrior")) == 2
    assert len(filter_priors(pri, "GaussianPrior", with_mat=True)) == 1
    assert len(filter_priors(pri, "GaussianPrior", with_mat=False)) == 1
    assert len(filter_priors(pri, "UniformPrior", with_mat=True)) == 0


# ---- variables ----------------------------------------------------------------
def test_variable_design_and_prior_params():
    df = pd.DataFrame({"cov": np.arange(5.0)})
    v = Variable("cov")
    assert isinstance(v.component, Component) and v.size == 1
    assert np.array_equal(v.get_design_mat(df), df[["cov"]].to_numpy())
    assert np.array_equal(v.get_direct_prior_params("GaussianPrior"), [[0.0], [np.inf]])
    assert np.array_equal(v.get_direct_prior_params("UniformPrior"), [[-np.inf], [np.inf]])
    params, mat = v.get_linear_prior_params("GaussianPrior")
    assert params.shape == (2, 0) and mat.shape == (0, 1)
    stacked = Variable("cov", priors=[GaussianPrior(0.0, 1.0, mat=np.ones((3, 1))), GaussianPrior(0.0, 1.0, mat=np.ones((3, 1)))])
    params, mat = stacked.get_linear_prior_params("GaussianPrior")
    assert params.shape == (2, 6) and np.array_equal(mat, np.ones((6, 1)))
    with pytest.raises(ValueError):
        Variable("cov", priors=[GaussianPrior(0.0, 1.0), GaussianPrior(0.0, 1.0)]).get_direct_prior_params("GaussianPrior")
    with pytest.raises(ValueError):
        Variable("cov", priors=[GaussianPrior(np.zeros(2), 1.0)]).get_direct_prior_params("GaussianPrior")
    with pytest.raises(ValueError):
        Variable("cov", priors=[GaussianPrior(0.0, 1.0, mat=np.ones((1, 2)))]).get_linear_prior_params("GaussianPrior")
    with pytest.raises(TypeError):
        Variable(1)
    with pytest.raises(TypeError):
        Variable("cov", priors=[1])


# ---- smooth mappings -----------------------------------------------------------
FUNS = [Identity(), Exp(), Log(), Expit(), Logit()]


@pytest.mark.parametrize("fun", FUNS)
def test_mapping_orders_and_inverse(fun):
    np.random.seed(123)
    x = np.random.rand(5)
    for order in (-1, 3):
        with pytest.raises(ValueError):
            fun(x, order=order)
    assert np.allclose(fun(x, order=1), np.diag(complex_step(fun, x, out_shape=(x.size,))))
    assert np.allclose(fun(x, order=2), np.diag(complex_step(partial(fun, order=1), x, out_shape=(x.size,))))
    assert np.allclose(x, fun(fun.inverse(x))) and np.allclose(x, fun.inverse(fun(x)))


def test_mapping_domains_and_codes(golden):
    for bad in (np.array([0.0]), np.array([-1.0])):
        with pytest.raises(ValueError):
            Log()(bad)
    for bad in (-1.0, 0.0, 1.0, 2.0):
        with pytest.raises(ValueError):
            Logit()(np.array([bad]))
    x = np.ones(3)
    assert Identity()(x) is x
    assert [f.device_code for f in FUNS] == [0, 1, 2, 3, 4]
    names = {"identity": Identity, "exp": Exp, "log": Log, "expit": Expit, "logit": Logit}
    for name, cls in names.items():  # bit-exact against the reference's recorded values
        for order in (0, 1, 2):
            with np.errstate(all="ignore"):
                got = cls()(golden[f"link/{name}/x"].copy(), order=order)
            np.testing.assert_array_equal(got, golden[f"link/{name}/order{order}"])


# ---- parameter (host logic only) -------------------------------------------------
def test_parameter_constructor_contract():
    vs = [Variable("cov1"), Variable("cov2")]
    p = Parameter(vs)
    assert repr(p.transform) == "Identity()" and p.offset is None and p.size == 2 and p.design_mat is None
    assert repr(Parameter(vs, transform=Exp()).transform) == "Exp()"
    assert isinstance(Parameter(vs, offset="cov").offset, Component)
    assert isinstance(Parameter(vs, offset=Component("cov")).offset, Component)
    for bad in ([1], ["a"]):
        with pytest.raises(TypeError):
            Parameter(bad)
    for bad in (1, np.exp):
        with pytest.raises(TypeError):
            Parameter(vs, transform=bad)
    with pytest.raises(TypeError):
        Parameter(vs, offset=1)
    with pytest.raises(TypeError):
        Parameter(vs, priors=[1, 2])
    assert len(Parameter(vs, priors=[GaussianPrior(np.zeros(2), np.ones(2))]).priors) == 1
    with pytest.raises(ValueError, match="Must provide a data frame"):
        p.get_params(np.ones(2))


def test_parameter_prior_assembly_matches_reference(golden):
    df = pd.DataFrame({"cov0": np.zeros(3), "cov1": np.zeros(3), "cov2": np.zeros(3)})
    Mp = golden["prior/param/Mp"]
    vs = [
        Variable(Component("intercept", default_value=1.0), priors=[GaussianPrior(mean=0.5, sd=2.0), UniformPrior(lb=-3.0, ub=3.0)]),
        Variable("cov0", priors=[GaussianPrior(mean=[0.0, 1.0], sd=[1.0, 3.0], mat=np.array([[1.0], [2.0]]))]),
        Variable("cov1"),
        Variable("cov2", priors=[GaussianPrior(mean=-0.3, sd=0.7)]),
    ]
    par = Parameter(vs, priors=[GaussianPrior(mean=[0.2, -0.1], sd=[0.9, 1.1], mat=Mp), UniformPrior(lb=[-1.0], ub=[1.0], mat=np.ones((1, 4))),
                                GaussianPrior(mean=np.zeros(4), sd=np.ones(4))])
    for v in vs:
        v.attach(df)
    for cat in ("direct", "linear"):
        for typ in ("UniformPrior", "GaussianPrior"):
            getattr(par, f"_get_{cat}_prior")(typ)
            pr = par.prior_dict[cat][typ]
            np.testing.assert_array_equal(pr.params, golden[f"prior/param/{cat}/{typ}/params"])
            if pr.mat is not None:
                np.testing.assert_array_equal(pr.mat, golden[f"prior/param/{cat}/{typ}/mat"])
    x = golden["prior/param/x"]
    assert par.prior_objective(x) == float(golden["prior/param/obj"])
    np.testing.assert_array_equal(par.prior_gradient(x), golden["prior/param/grad"])
    np.testing.assert_array_equal(par.prior_hessian(x), golden["prior/param/hess"])
    xc = x + 1e-20j  # dtype-generic like the reference (complex-step differentiation works)
    assert par.prior_gradient(xc).dtype == np.complex128


def test_default_priors_contribute_nothing():
    par = Parameter([Variable("a"), Variable("b")])
    for cat in ("direct", "linear"):
        for typ in ("UniformPrior", "GaussianPrior"):
            getattr(par, f"_get_{cat}_prior")(typ)
    x = np.ones(2)
    assert par.prior_dict["direct"]["GaussianPrior"].shape == (2, 2)
    lin = par.prior_dict["linear"]["GaussianPrior"]
    assert lin.params.shape == (2, 0) and lin.mat.shape == (0, 2)
    assert par.prior_objective(x) == 0.0
    assert np.allclose(par.prior_gradient(x), 0.0) and np.allclose(par.prior_hessian(x), 0.0)


# ---- spline getters (validation; basis values need the GPU) -----------------------
def test_spline_getter_contract():
    g = SplineGetter(np.linspace(0, 1, 5), degree=3)
    assert g.num_spline_bases == 7 and g.ldegree == 0 and g.rdegree == 0
    with pytest.raises(ValueError):
        SplineGetter(np.linspace(0, 1, 5), knots_type="quantile")
    data = np.array([2.0, 4.0, 10.0])
    sp = SplineGetter(np.array([0.0, 0.5, 1.0]), knots_type="rel_domain").get_spline(data)
    assert isinstance(sp, XSpline) and np.allclose(sp.knots, [2.0, 6.0, 10.0])
    sp = SplineGetter(np.array([0.0, 0.5, 1.0]), knots_type="rel_freq").get_spline(data)
    assert np.allclose(sp.knots, np.quantile(data, [0.0, 0.5, 1.0]))
    assert SplineGetter(np.array([0.0, 1.0]), knots_type="abs").get_spline(data).knots[1] == 1.0
    sv = SplineVariable("x", g)
    assert sv.size == 7
    with pytest.raises(TypeError):
        SplineVariable("x", "not a spline")
    assert SplineVariable("x", XSpline(np.linspace(0, 1, 4), 2)).size == 5


def test_spline_prior_getter_contract():
    with pytest.raises(TypeError):
        SplinePriorGetter("p")
    with pytest.raises(ValueError):
        SplinePriorGetter(GaussianPrior(0.0, 1.0, mat=np.ones((1, 1))))
    ok = GaussianPrior(0.0, 1.0)
    for kw in ({"size": 0}, {"order": -1}, {"domain": (0.0, 1.0, 2.0)}, {"domain": (1.0, 0.0)}, {"domain_type": "x"}):
        with pytest.raises(ValueError):
            SplinePriorGetter(ok, **kw)
    g = SplinePriorGetter(ok, size=5, domain=(0.25, 0.75))
    pts = g.sample_points(XSpline(np.array([0.0, 1.0, 4.0]), 3))
    assert np.allclose(pts, np.linspace(1.0, 3.0, 5))
    with pytest.raises(TypeError):
        g.get_prior("not a spline")
    SplineVariable("x", SplineGetter(np.linspace(0, 1, 4)), priors=[g, ok])  # both prior kinds are legal


def test_component_attaches_to_a_pyarrow_table():
    """Arrow ingest: a pyarrow.Table stands in for the pandas frame;
    a single-chunk null-free float64 column is handed on as a view of the Arrow buffer."""
    pa = pytest.importorskip("pyarrow")
    from anml_b200.data.component import Component
    from anml_b200.data.frame import ArrowFrame, as_frame
    from anml_b200.data.validator import NoNans

    a = np.arange(5, dtype=np.float64)
    table = pa.table({"a": a, "b": pa.array([1.0, None, 3.0, 4.0, 5.0]), "c": pa.chunked_array([[1.0, 2.0], [3.0, 4.0, 5.0]])})
    frame = as_frame(table)
    assert isinstance(frame, ArrowFrame) and as_frame(frame) is frame and len(frame) == 5
    comp = Component("a", validators=[NoNans()])
    comp.attach(frame)
    np.testing.assert_array_equal(comp.value, a)
    assert np.shares_memory(comp.value, table.column("a").chunk(0).to_numpy(zero_copy_only=True))
    with pytest.raises(ValueError):
        Component("b", validators=[NoNans()]).attach(frame)  # the null arrives as NaN
    multi = Component("c")
    multi.attach(table)  # a raw table is wrapped on the fly
    np.testing.assert_array_equal(multi.value, [1.0, 2.0, 3.0, 4.0, 5.0])
    icpt = Component("intercept", default_value=1.0)
    icpt.attach(frame)
    np.testing.assert_array_equal(icpt.value, np.ones(5))
    assert "intercept" in frame and "intercept" not in table.column_names
    with pytest.raises(KeyError):
        Component("missing").attach(frame)


def test_quantile_interpolation_restatement_matches_numpy(monkeypatch):
    """The host half of the device quantiles (anml_b200._native.vector_quantile): virtual index,
    neighbour ranks and NumPy's lerp, with the device order statistics replaced by a host sort."""
    from anml_b200 import _native as N

    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 10, 1001, 65537):
        x = np.round(rng.normal(size=n) * 5.0, 2)  # ties
        srt = np.sort(x)
        monkeypatch.setattr(N, "vector_minmax", lambda x_=None, dev_ptr=None, n=None, device=0, _s=srt: (float(_s[0]), float(_s[-1])))
        monkeypatch.setattr(N, "vector_order_stats", lambda ranks, x_=None, dev_ptr=None, n=None, device=0, _s=srt: _s[np.asarray(ranks)])
        q = np.concatenate([np.linspace(0, 1, 23), rng.uniform(size=40), [0.5, 1 / 3, 2 / 3, 1e-12, 1 - 1e-12]])
        np.testing.assert_array_equal(N.vector_quantile(q, x), np.quantile(x, q))
        np.testing.assert_array_equal(N.vector_quantile(q.reshape(4, -1), x), np.quantile(x, q.reshape(4, -1)))
    with pytest.raises(ValueError):
        N.vector_quantile([-0.1], np.arange(3.0))
