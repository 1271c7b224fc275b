"""Parity at BASELINE.json's full single-GPU sizes, through checks that do not
need the CPU oracle to hold the data: an independent FP64 evaluation with
torch/cuBLAS on the same device-resident inputs, shard additivity, determinism,
and (splines) partition of unity + interval indices against torch.searchsorted.
Inputs are generated on the device (torch), as bench.py does."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261019


def _torch():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch


def make_logistic(torch, n, p, seed):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    beta_true = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.1
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 1 << 21
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g)
        X[lo:hi, 0] = 1.0
        obs[lo:hi] = (torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(X[lo:hi] @ beta_true)).double()
    return X, obs, beta_true


def native_model(X, obs, lo, hi, prior=True):
    from anml_b200 import _native as N

    n, p = hi - lo, X.shape[1]
    Xs = X[lo:hi]
    d = N.Design.wrap(Xs.data_ptr(), n, p, p, keepalive=Xs)
    m = N.NativeModel(n, "binomial", 1)
    m.set_param(0, d, N.LINK_CODES["Expit"])
    os_ = obs[lo:hi]
    m.set_obs(dev_ptr=os_.data_ptr(), keepalive=os_)
    if prior:
        m.set_gaussian_prior(0, np.zeros(p), np.ones(p))
    return m


def test_config2_logistic_10m_rows_against_torch_fp64():
    """BASELINE config #2: binomial/expit, 1e7 rows x 64, GaussianPrior(0,1)."""
    torch = _torch()
    n, p = 10_000_000, 64
    X, obs, beta_true = make_logistic(torch, n, p, SEED + 2)
    x = (0.5 * beta_true).cpu().numpy()
    model = native_model(X, obs, 0, n)
    o, g, H = model.eval(x)
    o2, g2, H2 = model.eval(x)
    assert o == o2 and np.array_equal(g, g2) and np.array_equal(H, H2)  # bitwise reproducible
    assert np.array_equal(H, H.T)

    # independent evaluation: torch FP64 (cuBLAS DGEMV/DGEMM), chunked
    xb = torch.from_numpy(x).cuda()
    obj = torch.zeros((), dtype=torch.float64, device="cuda")
    grad = torch.zeros(p, dtype=torch.float64, device="cuda")
    hess = torch.zeros((p, p), dtype=torch.float64, device="cuda")
    chunk = 1 << 21
    for lo in range(0, n, chunk):
        Xc, oc = X[lo:lo + chunk], obs[lo:lo + chunk]
        y = Xc @ xb
        th = torch.sigmoid(y)
        obj += -(oc * torch.log(th) + (1 - oc) * torch.log1p(-th)).sum()
        grad += Xc.t() @ (th - oc)
        hess += (Xc.t() * (th * (1 - th))) @ Xc
    obj += 0.5 * (xb**2).sum()
    grad += xb
    hess += torch.eye(p, dtype=torch.float64, device="cuda")
    ow, gw, Hw = float(obj), grad.cpu().numpy(), hess.cpu().numpy()
    assert abs(o - ow) <= 1e-10 * abs(ow)
    assert np.abs(g - gw).max() <= 1e-10 * np.abs(gw).max()
    assert np.abs(H - Hw).max() <= 1e-9 * np.abs(Hw).max()

    # shard additivity (what the multi-GPU path relies on): halves sum to the whole
    a = native_model(X, obs, 0, n // 2 + 7, prior=True)   # "rank 0" carries the prior
    b = native_model(X, obs, n // 2 + 7, n, prior=False)
    oa, ga, Ha = a.eval(x)
    ob, gb, Hb = b.eval(x)
    assert abs((oa + ob) - o) <= 1e-12 * abs(o)
    assert np.abs(ga + gb - g).max() <= 1e-12 * np.abs(g).max()
    assert np.abs(Ha + Hb - H).max() <= 1e-12 * np.abs(H).max()

    # gradient pass alone (no Hessian) returns the same objective and gradient
    from anml_b200 import _native as N
    o3, g3, _ = model.eval(x, N.WANT_OBJ | N.WANT_GRAD)
    assert abs(o3 - o) <= 1e-13 * abs(o) and np.abs(g3 - g).max() <= 1e-12 * np.abs(g).max()


def test_config3_spline_50m_rows_partition_of_unity_and_indices():
    """BASELINE config #3: 5e7 rows, 20 knots, degree 3 -> 22 bases built on the device."""
    torch = _torch()
    from anml_b200 import _native as N

    n, nk, deg = 50_000_000, 20, 3
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + 3)
    xs = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    knots = np.linspace(0.0, 1.0, nk)
    xs[:nk] = torch.from_numpy(knots).cuda()  # exact knots and both end points
    nb = nk - 1 + deg
    design = N.Design(n, nb)
    idx = design.set_spline(0, None, knots, deg, want_index=True, x_dev_ptr=xs.data_ptr())
    want_idx = torch.clamp(torch.searchsorted(torch.from_numpy(knots).cuda(), xs, right=True) - 1, 0, nk - 2).cpu().numpy()
    np.testing.assert_array_equal(idx, want_idx.astype(np.int32))
    # partition of unity: design @ 1 == 1 for every row (get_params order 0, identity)
    ones = design.params(np.ones(nb), None, N.LINK_CODES["Identity"], 0)
    assert np.abs(ones - 1.0).max() <= 4e-16 * 4
    # 4 non-zeros per row, located at the interval: column sums of |basis| outside [idx, idx+3] are zero
    import ctypes as C
    n_rows, n_cols, ld, ptr = C.c_int64(), C.c_int(), C.c_int64(), C.c_void_p()
    N.check(N.load().anml_design_shape(design.handle, C.byref(n_rows), C.byref(n_cols), C.byref(ld), C.byref(ptr)))
    assert (n_rows.value, n_cols.value, ld.value) == (n, nb, nb)
    # Gaussian model on the spline design: Hessian is banded (bandwidth degree) and independent of x
    obs = torch.sin(2 * np.pi * xs) + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    m = N.NativeModel(n, "gaussian", 1)
    m.set_param(0, design, N.LINK_CODES["Identity"])
    m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
    o0, g0, H0 = m.eval(np.zeros(nb))
    o1, g1, H1 = m.eval(np.linspace(-1, 1, nb))
    assert np.array_equal(H0, H1)
    i, j = np.indices(H0.shape)
    assert np.all(H0[np.abs(i - j) > deg] == 0.0) and np.all(np.diag(H0) > 0)
    assert abs(H0.sum() - n) <= 1e-10 * n  # sum_ij H_ij = sum_rows (sum_j B_j)^2 = n
    # the banded evaluator (basis on the fly from the sorted abscissae) on the same 5e7 rows
    assert m.enable_banded_spline(knots, deg, x_dev_ptr=xs.data_ptr()) is True
    ob, gb, Hb = m.eval(np.linspace(-1, 1, nb))
    ob2, gb2, Hb2 = m.eval(np.linspace(-1, 1, nb))
    assert ob == ob2 and np.array_equal(gb, gb2) and np.array_equal(Hb, Hb2)  # bitwise reproducible
    assert abs(ob - o1) <= 1e-12 * abs(o1)
    assert np.abs(gb - g1).max() <= 1e-12 * np.abs(g1).max()
    assert np.abs(Hb - H1).max() <= 1e-12 * np.abs(H1).max()
    assert np.array_equal(Hb, Hb.T) and np.all(Hb[np.abs(i - j) > deg] == 0.0)
    assert abs(o0 - 0.5 * float((obs**2).sum())) <= 1e-12 * o0


def _free(torch):
    import gc

    gc.collect()
    torch.cuda.empty_cache()


def test_north_star_logistic_100m_rows_against_torch_fp64():
    """BASELINE north star: binomial/expit, 1e8 rows x 64 on one GPU (51.2 GB resident)."""
    torch = _torch()
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs a 180 GB device")
    n, p = 100_000_000, 64
    X, obs, beta_true = make_logistic(torch, n, p, SEED)
    x = (0.5 * beta_true).cpu().numpy()
    model = native_model(X, obs, 0, n)
    o, g, H = model.eval(x)
    assert np.array_equal(H, H.T)
    xb = torch.from_numpy(x).cuda()
    obj = torch.zeros((), dtype=torch.float64, device="cuda")
    grad = torch.zeros(p, dtype=torch.float64, device="cuda")
    hess = torch.zeros((p, p), dtype=torch.float64, device="cuda")
    chunk = 1 << 21
    for lo in range(0, n, chunk):
        Xc, oc = X[lo:lo + chunk], obs[lo:lo + chunk]
        y = Xc @ xb
        # the reference's own expressions (smoothmapping.py:119-126 for expit and its derivatives, binomial NLL through
        # log(theta), log(1 - theta)), not a numerically friendlier restatement
        z = torch.exp(-y)
        th = 1.0 / (1.0 + z)
        d1 = z / (1.0 + z) ** 2
        d2 = z * (z - 1.0) / (z + 1.0) ** 3
        obj += -(oc * torch.log(th) + (1 - oc) * torch.log(1 - th)).sum()
        dl = -oc / th + (1 - oc) / (1 - th)
        d2l = oc / th**2 + (1 - oc) / (1 - th) ** 2
        grad += Xc.t() @ (dl * d1)
        hess += (Xc.t() * (d2l * d1 * d1 + dl * d2)) @ Xc
    obj += 0.5 * (xb**2).sum()
    grad += xb
    hess += torch.eye(p, dtype=torch.float64, device="cuda")
    ow, gw, Hw = float(obj), grad.cpu().numpy(), hess.cpu().numpy()
    assert abs(o - ow) <= 1e-10 * abs(ow)
    assert np.abs(g - gw).max() <= 1e-10 * np.abs(gw).max()
    assert np.abs(H - Hw).max() <= 1e-9 * np.abs(Hw).max()
    del model, X, obs
    _free(torch)


def test_config4_two_parameter_shared_design_25m_rows_against_torch_fp64():
    """BASELINE config #4 (mean identity + variance exp over the same 64 columns) at the
    2.5e7-row shard of a 4-GPU run: l = 1/2 log v + 1/2 (obs - mu)^2 / v, v = exp(X b_s)."""
    torch = _torch()
    from anml_b200 import _native as N

    n, p = 25_000_000, 64
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 4)
    bm = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.1
    bs = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.05
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 1 << 21
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g)
        X[lo:hi, 0] = 1.0
        obs[lo:hi] = X[lo:hi] @ bm + torch.exp(0.5 * (X[lo:hi] @ bs)) * torch.randn(hi - lo, dtype=torch.float64, device=dev, generator=g)
    d = N.Design.wrap(X.data_ptr(), n, p, p, keepalive=X)
    m = N.NativeModel(n, "gaussian_meanvar", 2)
    m.set_param(0, d, N.LINK_CODES["Identity"])
    m.set_param(1, d, N.LINK_CODES["Exp"])
    m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
    xm, xs = 0.5 * bm, 0.5 * bs
    x = torch.cat([xm, xs]).cpu().numpy()
    o, gr, H = m.eval(x)
    assert H.shape == (2 * p, 2 * p) and np.array_equal(H, H.T)
    obj = torch.zeros((), dtype=torch.float64, device=dev)
    grad = torch.zeros(2 * p, dtype=torch.float64, device=dev)
    hess = torch.zeros((2 * p, 2 * p), dtype=torch.float64, device=dev)
    for lo in range(0, n, chunk):
        Xc, oc = X[lo:lo + chunk], obs[lo:lo + chunk]
        mu, ls = Xc @ xm, Xc @ xs
        iv = torch.exp(-ls)
        res = oc - mu
        obj += (0.5 * ls + 0.5 * res * res * iv).sum()
        r0 = -res * iv                       # dl/dmu
        r1 = 0.5 - 0.5 * res * res * iv      # dl/dlog v
        w00, w01, w11 = iv, res * iv, 0.5 * res * res * iv
        grad[:p] += Xc.t() @ r0
        grad[p:] += Xc.t() @ r1
        hess[:p, :p] += (Xc.t() * w00) @ Xc
        hess[:p, p:] += (Xc.t() * w01) @ Xc
        hess[p:, p:] += (Xc.t() * w11) @ Xc
    hess[p:, :p] = hess[:p, p:].t()
    ow, gw, Hw = float(obj), grad.cpu().numpy(), hess.cpu().numpy()
    assert abs(o - ow) <= 1e-10 * abs(ow)
    assert np.abs(gr - gw).max() <= 1e-10 * np.abs(gw).max()
    assert np.abs(H - Hw).max() <= 1e-9 * np.abs(Hw).max()
    del m, d, X, obs
    _free(torch)


def test_config5_wide_design_shard_against_torch_fp64():
    """BASELINE config #5 at the per-GPU shard of the 8-GPU run: 2.5e6 rows x 1024 columns
    (20.5 GB), binomial/expit, Hessian on the wide DMMA kernel."""
    torch = _torch()
    from anml_b200 import _native as N

    n, p = 2_500_000, 1024
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 5)
    beta_true = torch.randn(p, dtype=torch.float64, device=dev, generator=g) * 0.025
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    obs = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 1 << 17
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        X[lo:hi].normal_(generator=g)
        X[lo:hi, 0] = 1.0
        obs[lo:hi] = (torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(X[lo:hi] @ beta_true)).double()
    d = N.Design.wrap(X.data_ptr(), n, p, p, keepalive=X)
    m = N.NativeModel(n, "binomial", 1)
    m.set_param(0, d, N.LINK_CODES["Expit"])
    m.set_obs(dev_ptr=obs.data_ptr(), keepalive=obs)
    xb = 0.5 * beta_true
    o, gr, H = m.eval(xb.cpu().numpy())
    assert np.array_equal(H, H.T)
    obj = torch.zeros((), dtype=torch.float64, device=dev)
    grad = torch.zeros(p, dtype=torch.float64, device=dev)
    hess = torch.zeros((p, p), dtype=torch.float64, device=dev)
    for lo in range(0, n, chunk):
        Xc, oc = X[lo:lo + chunk], obs[lo:lo + chunk]
        y = Xc @ xb
        th = torch.sigmoid(y)
        obj += (torch.nn.functional.softplus(-y) + (1 - oc) * y).sum()
        grad += Xc.t() @ (th - oc)
        hess += (Xc.t() * (th * (1 - th))) @ Xc
    ow, gw, Hw = float(obj), grad.cpu().numpy(), hess.cpu().numpy()
    assert abs(o - ow) <= 1e-10 * abs(ow)
    assert np.abs(gr - gw).max() <= 1e-10 * np.abs(gw).max()
    assert np.abs(H - Hw).max() <= 1e-9 * np.abs(Hw).max()
    del m, d, X, obs
    _free(torch)
