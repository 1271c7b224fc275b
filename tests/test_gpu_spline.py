"""GPU spline-basis kernel: knot-interval indices bit-exact, basis values
bit-exact inside the knot span, derivatives and extrapolation within 1e-13."""
import numpy as np
import pytest

from oracle import anml_oracle as O

pytestmark = pytest.mark.gpu


def sample_points(knots, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(knots[0], knots[-1], size=n)
    # exact knots, end points, and the floats adjacent to each knot
    special = np.concatenate([knots, np.nextafter(knots, -np.inf), np.nextafter(knots, np.inf)])
    special = special[(special >= knots[0]) & (special <= knots[-1])]
    return np.concatenate([x, special])


@pytest.mark.parametrize("degree", [1, 2, 3, 5])
@pytest.mark.parametrize("nk", [2, 5, 20])
def test_basis_and_interval_index_bit_exact(degree, nk):
    from anml_b200.xspline import XSpline

    knots = np.sort(np.random.default_rng(nk).uniform(-2, 3, size=nk))
    x = sample_points(knots, 100_000 if nk == 20 else 5000, seed=degree)
    D, idx = XSpline(knots, degree).get_design_mat(x, return_index=True)
    want, widx = O.spline_design(knots, degree, x, return_index=True)
    assert idx.dtype == np.int32
    np.testing.assert_array_equal(idx, widx)
    np.testing.assert_array_equal(D, want)
    assert (np.count_nonzero(D, axis=1) <= degree + 1).all()
    np.testing.assert_allclose(D.sum(axis=1), 1.0, atol=1e-13)


def test_config3_interval_indices_one_million_rows():
    """indices compared bit-exactly on a 1e6-row sample incl.
    exact-knot and end-point values (20 knots on [0, 1], degree 3 -> 22 bases)."""
    from anml_b200.xspline import XSpline

    knots = np.linspace(0.0, 1.0, 20)
    x = sample_points(knots, 1_000_000, seed=20261019 + 3)
    sp = XSpline(knots, 3)
    idx = sp.interval_index(x)
    np.testing.assert_array_equal(idx, O.spline_interval_index(knots, x))
    np.testing.assert_array_equal(sp.interval_index(np.array([knots[0], knots[-1], knots[5]])), [0, 18, 5])


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_derivative_orders(order):
    from anml_b200.xspline import XSpline

    knots = np.linspace(0.0, 2.0, 9)
    x = sample_points(knots, 3000, seed=order)
    D = XSpline(knots, 3).get_design_mat(x, order=order)
    want = O.spline_design(knots, 3, x, order=order)
    np.testing.assert_allclose(D, want, rtol=0, atol=1e-13 * max(np.abs(want).max(), 1.0))
    if order > 3:
        assert not D.any()


@pytest.mark.parametrize("ldeg,rdeg", [(0, 0), (1, 1), (2, 3), (3, 0)])
@pytest.mark.parametrize("order", [0, 1])
def test_extrapolation(ldeg, rdeg, order):
    from anml_b200.xspline import XSpline

    knots = np.array([0.0, 0.3, 0.7, 1.5])
    x = np.concatenate([np.linspace(-1.0, 2.5, 71), knots])
    D, idx = XSpline(knots, 3, ldegree=ldeg, rdegree=rdeg).get_design_mat(x, order=order, return_index=True)
    want, widx = O.spline_design(knots, 3, x, order=order, ldegree=ldeg, rdegree=rdeg, return_index=True)
    np.testing.assert_array_equal(idx, widx)
    np.testing.assert_allclose(D, want, rtol=0, atol=1e-12 * max(np.abs(want).max(), 1.0))


def test_spline_argument_errors():
    from anml_b200 import _native
    from anml_b200.xspline import XSpline

    with pytest.raises(ValueError):
        XSpline([1.0], 3)
    with pytest.raises(ValueError):
        XSpline([1.0, 0.0], 3)
    with pytest.raises(ValueError):
        _native.spline_design([0.5], [0.0, 0.0], 3)  # empty knot span
    with pytest.raises(NotImplementedError):
        _native.spline_design([0.5], [0.0, 1.0], 9)  # degree above the kernel's limit
    out = _native.spline_design(np.empty(0), [0.0, 1.0], 3)
    assert out.shape == (0, 4)


@pytest.mark.parametrize("n", [1, 2, 7, 100_003, (1 << 20) + 3])
def test_device_minmax_and_quantiles_carry_numpy_bits(n):
    """Knot placement from data (getter/spline.py:92-99): data.min()/max() and
    np.quantile(data, k) recomputed on the device are the same doubles."""
    from anml_b200 import _native as N

    rng = np.random.default_rng(n)
    x = rng.normal(size=n) * 3.0
    if n > 10:
        x[: n // 3] = np.round(x[: n // 3], 1)  # plenty of ties
    q = np.concatenate([np.linspace(0.0, 1.0, 20), [0.5, 1.0 / 3.0, 0.999999, 1e-9]])
    lo, hi = N.vector_minmax(x)
    assert lo == x.min() and hi == x.max()
    got = N.vector_quantile(q, x)
    want = np.quantile(x, q)
    np.testing.assert_array_equal(got, want)
    # same through a device-resident copy
    buf = N.DeviceBuffer(0, x)
    np.testing.assert_array_equal(N.vector_quantile(q, dev_ptr=buf.ptr, n=n), want)
    assert N.vector_minmax(dev_ptr=buf.ptr, n=n) == (x.min(), x.max())
    # order statistics themselves
    ranks = np.unique(rng.integers(0, n, size=min(n, 50)))
    np.testing.assert_array_equal(N.vector_order_stats(ranks, x), np.sort(x)[ranks])


def test_device_minmax_and_quantiles_with_nan_and_errors():
    from anml_b200 import _native as N

    x = np.array([1.0, np.nan, -2.0, 5.0])
    lo, hi = N.vector_minmax(x)
    assert np.isnan(lo) and np.isnan(hi)
    assert np.isnan(N.vector_quantile([0.0, 0.5], x)).all() and np.isnan(np.quantile(x, [0.0, 0.5])).all()
    with pytest.raises(ValueError):
        N.vector_minmax(np.empty(0))
    with pytest.raises(ValueError):
        N.vector_quantile([1.5], np.arange(4.0))


@pytest.mark.parametrize("knots_type", ["rel_domain", "rel_freq"])
def test_spline_variable_resolves_data_dependent_knots_on_device(knots_type):
    """Above 65536 rows SplineVariable places rel_domain / rel_freq knots from the
    uploaded abscissae; they equal the host rule bit for bit and the basis follows."""
    import pandas as pd

    from anml_b200.getter.spline import SplineGetter
    from anml_b200.parameter import Parameter
    from anml_b200.variable import SplineVariable

    rng = np.random.default_rng(8)
    n = 150_001
    x = rng.gamma(2.0, 1.5, size=n)
    df = pd.DataFrame({"x": x})
    fractions = np.linspace(0.0, 1.0, 9)
    getter = SplineGetter(fractions, degree=3, knots_type=knots_type)
    want_knots = getter.resolve_knots(x)
    var = SplineVariable("x", getter)
    par = Parameter([var])
    par.attach(df)
    np.testing.assert_array_equal(var.spline.knots, want_knots)
    assert par.size == 11 and par.design_mat.shape == (n, 11)
    want = O.spline_design(want_knots, 3, x)
    np.testing.assert_array_equal(par.design_mat, want)


def _spline_model(family, n, seed, knots_type="rel_domain", with_extras=False, degree=3):
    import pandas as pd

    from anml_b200.data import Component, DataExample, DataPrototype
    from anml_b200.getter.prior import SplinePriorGetter
    from anml_b200.getter.spline import SplineGetter
    from anml_b200.model import Model
    from anml_b200.parameter import Expit, Identity, Parameter
    from anml_b200.prior import GaussianPrior
    from anml_b200.variable import SplineVariable

    rng = np.random.default_rng(seed)
    x = rng.beta(2.0, 3.0, size=n) * 4.0 - 1.0
    df = pd.DataFrame({"x": x})
    f = np.sin(1.5 * x)
    if family == "binomial":
        df["obs"] = (rng.uniform(size=n) < 1 / (1 + np.exp(-2 * f))).astype(float)
    else:
        df["obs"] = f + 0.1 * rng.normal(size=n)
        df["obs_se"] = rng.uniform(0.5, 1.5, size=n)
    offset = None
    if with_extras:
        df["off"] = 0.1 * rng.normal(size=n)
        offset = "off"
    var = SplineVariable("x", SplineGetter(np.linspace(0.0, 1.0, 12), degree=degree, knots_type=knots_type),
                         priors=[SplinePriorGetter(GaussianPrior(mean=0.0, sd=2.0), size=50, order=2)])
    par = Parameter([var], transform=Expit() if family == "binomial" else Identity(), offset=offset)
    model = Model(family, DataExample("obs", "obs_se"), [par])
    return model, par, var, df


@pytest.mark.parametrize("family,extras,degree", [("gaussian", False, 3), ("gaussian", True, 2), ("binomial", True, 3), ("binomial", False, 1)])
def test_one_spline_model_runs_on_the_banded_evaluator(family, extras, degree):
    """BASELINE config #3 shape: the model whose only variable is a spline is evaluated from the
    abscissae (spline_eval.cuh).  Same objective / gradient / Hessian as the dense design route
    (tolerances of BASELINE.json) and as the oracle on the materialised basis."""
    n = 120_007
    model, par, var, df = _spline_model(family, n, seed=degree + 10 * extras, with_extras=extras, degree=degree)
    model.attach(df)
    assert model.banded_spline is True
    nb = par.size
    x = np.random.default_rng(5).normal(size=nb) * 0.3
    o, g, H = model.evaluate(x)
    assert np.array_equal(H, H.T)
    i, j = np.indices(H.shape)
    lin = par.prior_dict["linear"]["GaussianPrior"]
    prior_H = (lin.mat.T / lin.params[1] ** 2) @ lin.mat
    assert np.all((H - prior_H)[np.abs(i - j) > degree] == 0.0)  # likelihood part is banded
    # dense design route of the same model
    model.native.set_option("force_generic", 0)
    model.use_banded_spline = False
    model.attach(df)
    assert model.banded_spline is False
    od, gd, Hd = model.evaluate(x * (1 + 1e-16))
    assert abs(o - od) <= 1e-12 * abs(od)
    assert np.abs(g - gd).max() <= 1e-12 * np.abs(gd).max()
    assert np.abs(H - Hd).max() <= 1e-12 * np.abs(Hd).max()
    # oracle on the host basis
    B = O.spline_design(var.spline.knots, degree, df["x"].to_numpy())
    link = "expit" if family == "binomial" else "identity"
    priors = [{"linear": (lin.params[0], lin.params[1], lin.mat)}]
    want = O.model_eval_fused(family, df["obs"].to_numpy(), [B], [link], x, offsets=[df["off"].to_numpy() if extras else None],
                              obs_se=df["obs_se"].to_numpy() if family == "gaussian" else None, priors=priors)
    assert abs(o - want[0]) <= 1e-10 * abs(want[0])
    assert np.abs(g - want[1]).max() <= 1e-10 * np.abs(want[1]).max()
    assert np.abs(H - want[2]).max() <= 1e-9 * np.abs(want[2]).max()


def test_banded_evaluator_declines_when_rows_leave_the_knot_span():
    """Absolute knots narrower than the data: extrapolated rows are not handled by the banded
    path, the library says so and the model stays on the dense design."""
    import pandas as pd

    from anml_b200.data import DataExample
    from anml_b200.getter.spline import SplineGetter
    from anml_b200.model import Model
    from anml_b200.parameter import Parameter
    from anml_b200.variable import SplineVariable

    rng = np.random.default_rng(3)
    n = 70_001
    df = pd.DataFrame({"x": rng.uniform(-0.2, 1.2, size=n)})
    df["obs"] = df["x"] ** 2 + 0.1 * rng.normal(size=n)
    var = SplineVariable("x", SplineGetter(np.linspace(0.0, 1.0, 8), degree=3, knots_type="abs"))
    model = Model("gaussian", DataExample("obs", "obs_se"), [Parameter([var])])
    model.attach(df)
    assert model.banded_spline is False
    x = rng.normal(size=model.size)
    B = O.spline_design(var.spline.knots, 3, df["x"].to_numpy())
    want = O.model_eval_fused("gaussian", df["obs"].to_numpy(), [B], ["identity"], x, obs_se=np.ones(n))
    o, g, H = model.evaluate(x)
    assert abs(o - want[0]) <= 1e-10 * abs(want[0]) and np.abs(H - want[2]).max() <= 1e-9 * np.abs(want[2]).max()


def test_banded_evaluator_poisson_with_weights_and_offset():
    """The banded evaluator with per-row likelihood weights, an offset and the Poisson/exp pair."""
    import pandas as pd

    from anml_b200.data import Component, DataPrototype
    from anml_b200.data.validator import NoNans
    from anml_b200.getter.spline import SplineGetter
    from anml_b200.model import Model
    from anml_b200.parameter import Exp, Parameter
    from anml_b200.variable import SplineVariable

    rng = np.random.default_rng(12)
    n = 90_011
    xv = rng.uniform(0.0, 3.0, size=n)
    off = 0.2 * rng.normal(size=n)
    wts = rng.uniform(0.0, 2.0, size=n)
    wts[::7] = 0.0  # trimmed rows
    obs = rng.poisson(np.exp(0.5 * np.sin(xv) + off)).astype(float)
    df = pd.DataFrame({"x": xv, "off": off, "w": wts, "obs": obs})
    var = SplineVariable("x", SplineGetter(np.linspace(0.0, 1.0, 10), degree=3, knots_type="rel_freq"))
    par = Parameter([var], transform=Exp(), offset="off")
    data = DataPrototype({"obs": Component("obs", [NoNans()]), "weights": Component("w", [NoNans()])})
    model = Model("poisson", data, [par])
    model.attach(df)
    assert model.banded_spline is True
    x = rng.normal(size=par.size) * 0.2
    o, g, H = model.evaluate(x)
    B = O.spline_design(var.spline.knots, 3, xv)
    want = O.model_eval_fused("poisson", obs, [B], ["exp"], x, offsets=[off], weights=wts)
    assert abs(o - want[0]) <= 1e-10 * abs(want[0])
    assert np.abs(g - want[1]).max() <= 1e-10 * np.abs(want[1]).max()
    assert np.abs(H - want[2]).max() <= 1e-9 * np.abs(want[2]).max()
    # gradient-only evaluation agrees
    o2, g2, _ = model.evaluate(x * (1 + 1e-16), hessian=False)
    assert abs(o2 - o) <= 1e-13 * abs(o) and np.abs(g2 - g).max() <= 1e-12 * np.abs(g).max()


def test_device_spline_against_recorded_xspline():
    """Device basis / derivative / extrapolation values and the SplineVariable / SplinePriorGetter outputs against
    outputs recorded from the REAL xspline + reference (tests/golden/make_spline_golden.py).  Skips until that
    file exists: xspline is not importable in this image, so these rows stay 'parity unpinned' (DESIGN.md 5)."""
    import os

    import pandas as pd

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xspline_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/xspline_golden.npz not recorded: spline parity unpinned")
    from anml_b200 import _native as N
    from anml_b200.getter.prior import SplinePriorGetter
    from anml_b200.getter.spline import SplineGetter
    from anml_b200.prior import GaussianPrior
    from anml_b200.variable import SplineVariable

    g = np.load(path, allow_pickle=False)
    for name in [str(c) for c in g["cases"]]:
        knots, x = g[f"{name}/knots"], g[f"{name}/x"]
        degree, ldeg, rdeg = (int(v) for v in g[f"{name}/params"])
        for order in range(degree + 1):
            want = g[f"{name}/order{order}"]
            got = N.spline_design(x, knots, degree, ldeg, rdeg, order)
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(want).max()), err_msg=f"{name} order {order}")
    data = g["getter/data"]
    sp = SplineGetter(np.linspace(0.0, 1.0, 7), degree=3, knots_type="rel_domain").get_spline(data)
    for order in (0, 1, 2):
        pr = SplinePriorGetter(GaussianPrior(0.0, 1.0), size=50, order=order).get_prior(sp)
        want = g[f"priorgetter/order{order}/mat"]
        np.testing.assert_allclose(pr.mat, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    sv = SplineVariable("v", SplineGetter(np.linspace(0.0, 1.0, 7), degree=3, knots_type="rel_freq"))
    np.testing.assert_allclose(sv.get_design_mat(pd.DataFrame({"v": data})), g["variable/design"], rtol=1e-12, atol=1e-14)
