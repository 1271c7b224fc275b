"""NumPy restatement of anml's likelihood-evaluation path (the oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites
the reference lines (relative to ``/root/reference``) whose arithmetic it
restates.  The restatement is plain functions over arrays (no classes), so the
CUDA path can be checked term by term.

Parity status
-------------
* links, ``get_params`` orders 0/1/2, Gaussian prior terms: PINNED -- checked
  against the imported reference (``tests/golden/make_golden.py`` writes the
  fixtures; ``tests/test_oracle_golden.py`` replays them on every run).
* model layer (family NLL + chain rule): not in the reference;
  it is composed here from the pinned ``get_params``/prior restatements and
  cross-checked between the literal (n,p,p) route and the fused route.
* spline basis (``xspline`` is a third-party package whose source is absent
  and which is not installed): PARITY UNPINNED against xspline itself.  The
  restatement is the clamped B-spline basis (Cox-de Boor) and is cross-checked
  against ``scipy.interpolate.BSpline`` inside the knot span; extrapolation
  (``ldegree``/``rdegree``) follows xspline's documented rule (Taylor
  polynomial of that degree about the end knot) and is unpinned.
"""
from __future__ import annotations

import math

import numpy as np

LINKS = ("identity", "exp", "log", "expit", "logit")
FAMILIES = ("gaussian", "binomial", "poisson", "gaussian_meanvar")


# --------------------------------------------------------------------------
# SmoothMapping family -- src/anml/parameter/smoothmapping.py
# --------------------------------------------------------------------------
def link_eval(link: str, y: np.ndarray, order: int = 0) -> np.ndarray:
    """f(y), f'(y) or f''(y) for one of anml's five smooth mappings.

    identity: smoothmapping.py:59-65; exp: :76-78; log: :96-104 (domain check
    :98-99); expit: :119-126 (the exp(-y) formulas, kept literally);
    logit: :146-156 (domain check :148-151); order check :32-33.
    """
    if order not in (0, 1, 2):
        raise ValueError("Order must be 0, 1 or 2.")
    y = np.asarray(y)
    if link == "identity":
        if order == 0:
            return y
        return np.ones(y.shape) if order == 1 else np.zeros(y.shape)
    if link == "exp":
        return np.exp(y)
    if link == "log":
        if not (y > 0).all():
            raise ValueError("All values for log function must be positive.")
        if order == 0:
            return np.log(y)
        return 1 / y if order == 1 else -1 / y**2
    if link == "expit":
        z = np.exp(-y)
        if order == 0:
            return 1 / (1 + z)
        if order == 1:
            return z / (1 + z) ** 2
        return z * (z - 1) / (z + 1) ** 3
    if link == "logit":
        if not ((y > 0).all() and (y < 1).all()):
            raise ValueError("All values for logit function must be strictly between 0 and 1.")
        if order == 0:
            return np.log(y / (1 - y))
        if order == 1:
            return 1 / (y * (1 - y))
        return (2 * y - 1) / (y * (1 - y)) ** 2
    raise KeyError(link)


# --------------------------------------------------------------------------
# Parameter.get_params -- src/anml/parameter/main.py:265-280
# --------------------------------------------------------------------------
def linear_predictor(X: np.ndarray, beta: np.ndarray, offset=None) -> np.ndarray:
    """y = X @ beta (+ offset): parameter/main.py:269-271."""
    y = X.dot(beta)
    if offset is not None:
        y = y + offset
    return y


def get_params(X, beta, link="identity", offset=None, order=0):
    """theta / Jacobian / second-order tensor: parameter/main.py:272-280."""
    z = link_eval(link, linear_predictor(X, beta, offset), order)
    if order == 0:
        return z
    if order == 1:
        return z[:, None] * X
    return z[:, None, None] * (X[..., None] * X[:, None, :])


# --------------------------------------------------------------------------
# GaussianPrior terms -- src/anml/prior/main.py:178-197, summed as
# Parameter.prior_objective/gradient/hessian do (parameter/main.py:282-340)
# --------------------------------------------------------------------------
def gaussian_prior_terms(x, mean, sd, mat=None):
    """(objective, gradient, hessian) of one Gaussian prior.

    mat None  -> prior/main.py:179-180, :186-187, :193-194
    mat empty -> :181-182, :188-189, :195-196
    otherwise -> :183, :190, :197
    """
    x = np.asarray(x)
    p = x.size
    if mat is None:
        return (
            0.5 * np.sum(((x - mean) / sd) ** 2),
            (x - mean) / sd**2,
            np.diag(1 / sd**2) * np.ones((p, p)),
        )
    mat = np.asarray(mat)
    if mat.size == 0:
        return 0.0, np.zeros(p), np.zeros((p, p))
    res = mat.dot(x) - mean
    scaled_t = mat.T / sd**2
    return 0.5 * np.sum((res / sd) ** 2), scaled_t.dot(res), scaled_t.dot(mat)


def parameter_prior_terms(x, direct=None, linear=None):
    """Direct + linear Gaussian prior of one Parameter (parameter/main.py:296-339).

    ``direct`` = (mean, sd) of length p (defaults N(0, inf): prior/main.py:166);
    ``linear`` = (mean, sd, mat) with mat (m, p), possibly m == 0.
    """
    x = np.asarray(x, dtype=float)
    p = x.size
    obj, grad, hess = 0.0, np.zeros(p), np.zeros((p, p))
    if direct is None:
        direct = (np.zeros(p), np.full(p, np.inf))
    o, g, h = gaussian_prior_terms(x, np.asarray(direct[0], float), np.asarray(direct[1], float))
    obj, grad, hess = obj + o, grad + g, hess + h
    if linear is not None:
        o, g, h = gaussian_prior_terms(
            x, np.asarray(linear[0], float), np.asarray(linear[1], float), np.asarray(linear[2], float)
        )
        obj, grad, hess = obj + o, grad + g, hess + h
    return obj, grad, hess


# --------------------------------------------------------------------------
# Model layer: per-row negative log-likelihood in the
# distribution parameters theta_k and its first/second partials.  Not in the
# reference; defined by BASELINE.json's configs.
# --------------------------------------------------------------------------
def family_num_params(family: str) -> int:
    return 2 if family == "gaussian_meanvar" else 1


def family_terms(family, obs, thetas, obs_se=None):
    """Return (l, [dl/dtheta_k], {(k,l): d2l/dtheta_k dtheta_l for k<=l}) per row.

    gaussian          l = 1/2 ((obs - mu)/se)^2        (data/example.py:7-8 "least-square")
    binomial          l = -(obs log p + (1-obs) log(1-p))
    poisson           l = lam - obs log lam
    gaussian_meanvar  l = 1/2 log v + 1/2 (obs-mu)^2 / v
    """
    obs = np.asarray(obs, dtype=float)
    if family == "gaussian":
        (mu,) = thetas
        se = np.ones_like(obs) if obs_se is None else np.asarray(obs_se, float)
        res = (obs - mu) / se
        return 0.5 * res**2, [-(obs - mu) / se**2], {(0, 0): 1.0 / se**2}
    if family == "binomial":
        (p,) = thetas
        l = -(obs * np.log(p) + (1 - obs) * np.log(1 - p))
        return l, [-obs / p + (1 - obs) / (1 - p)], {(0, 0): obs / p**2 + (1 - obs) / (1 - p) ** 2}
    if family == "poisson":
        (lam,) = thetas
        return lam - obs * np.log(lam), [1 - obs / lam], {(0, 0): obs / lam**2}
    if family == "gaussian_meanvar":
        mu, v = thetas
        res = obs - mu
        l = 0.5 * np.log(v) + 0.5 * res**2 / v
        d = [-res / v, 0.5 / v - 0.5 * res**2 / v**2]
        h = {(0, 0): 1.0 / v, (0, 1): res / v**2, (1, 1): -0.5 / v**2 + res**2 / v**3}
        return l, d, h
    raise KeyError(family)


def _split(x, sizes):
    out, at = [], 0
    for s in sizes:
        out.append(np.asarray(x[at : at + s], dtype=float))
        at += s
    return out


def model_eval_literal(family, obs, Xs, links, x, offsets=None, obs_se=None, priors=None):
    """obj/grad/Hessian by the reference's literal route.

    theta_k, J_k (n,p_k) and T_k (n,p_k,p_k) come from ``get_params`` orders
    0/1/2 exactly as parameter/main.py:272-280 materialises them, then
        g_k  = J_k^T dl_k
        H_kl = J_k^T diag(d2l_kl) J_l + [k==l] tensordot(dl_k, T_k, 1)
    plus the Gaussian prior terms of each Parameter.  Only feasible for small n.
    """
    K = len(Xs)
    sizes = [X.shape[1] for X in Xs]
    betas = _split(x, sizes)
    offsets = offsets or [None] * K
    th = [get_params(Xs[k], betas[k], links[k], offsets[k], 0) for k in range(K)]
    J = [get_params(Xs[k], betas[k], links[k], offsets[k], 1) for k in range(K)]
    T = [get_params(Xs[k], betas[k], links[k], offsets[k], 2) for k in range(K)]
    l, dl, d2l = family_terms(family, obs, th, obs_se)
    P = sum(sizes)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    obj = float(np.sum(l))
    grad = np.zeros(P)
    hess = np.zeros((P, P))
    for k in range(K):
        sk = slice(starts[k], starts[k + 1])
        grad[sk] = J[k].T.dot(dl[k])
        for m in range(k, K):
            sm = slice(starts[m], starts[m + 1])
            blk = (J[k].T * d2l[(k, m)]).dot(J[m])
            if k == m:
                blk = blk + np.tensordot(dl[k], T[k], axes=1)
            hess[sk, sm] = blk
            if m != k:
                hess[sm, sk] = blk.T
    return _add_priors(obj, grad, hess, betas, starts, priors)


def model_eval_fused(family, obs, Xs, links, x, offsets=None, obs_se=None, priors=None, chunk=1 << 20, weights=None):
    """Same quantities without the (n,p,p) tensor (last lines):
        r_k  = dl_k f_k'                 g_k  = X_k^T r_k
        w_kl = d2l_kl f_k' f_l' + [k==l] dl_k f_k''     H_kl = X_k^T diag(w_kl) X_l
    evaluated in row chunks; this is the "best-effort NumPy" CPU baseline of
    BASELINE.md section 4 and the big-n oracle.
    """
    K = len(Xs)
    n = Xs[0].shape[0]
    sizes = [X.shape[1] for X in Xs]
    betas = _split(x, sizes)
    offsets = offsets or [None] * K
    P = sum(sizes)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    obj, grad, hess = 0.0, np.zeros(P), np.zeros((P, P))
    obs = np.asarray(obs, float)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        Xc = [X[lo:hi] for X in Xs]
        ys = [linear_predictor(Xc[k], betas[k], None if offsets[k] is None else offsets[k][lo:hi]) for k in range(K)]
        th = [link_eval(links[k], ys[k], 0) for k in range(K)]
        f1 = [link_eval(links[k], ys[k], 1) for k in range(K)]
        f2 = [link_eval(links[k], ys[k], 2) for k in range(K)]
        l, dl, d2l = family_terms(family, obs[lo:hi], th, None if obs_se is None else obs_se[lo:hi])
        if weights is not None:  # per-row likelihood weights scale every term of the row
            wt = np.asarray(weights, float)[lo:hi]
            l = wt * l
            dl = [wt * d for d in dl]
            d2l = {k: wt * v for k, v in d2l.items()}
        obj += float(np.sum(l))
        for k in range(K):
            sk = slice(starts[k], starts[k + 1])
            grad[sk] += Xc[k].T.dot(dl[k] * f1[k])
            for m in range(k, K):
                sm = slice(starts[m], starts[m + 1])
                w = d2l[(k, m)] * f1[k] * f1[m]
                if k == m:
                    w = w + dl[k] * f2[k]
                blk = (Xc[k].T * w).dot(Xc[m])
                hess[sk, sm] += blk
                if m != k:
                    hess[sm, sk] += blk.T
    return _add_priors(obj, grad, hess, betas, starts, priors)


def _add_priors(obj, grad, hess, betas, starts, priors):
    if priors is not None:
        for k, pr in enumerate(priors):
            if pr is None:
                continue
            o, g, h = parameter_prior_terms(betas[k], pr.get("direct"), pr.get("linear"))
            sk = slice(starts[k], starts[k + 1])
            obj += float(o)
            grad[sk] += g
            hess[sk, sk] += h
    return obj, grad, hess


# --------------------------------------------------------------------------
# Spline basis -- restates what anml asks of xspline.XSpline at
# getter/spline.py:101-106, variable/spline.py:111-113, getter/prior.py:191-194
# --------------------------------------------------------------------------
def spline_num_bases(nk: int, degree: int) -> int:
    """len(knots) - 1 + degree for ldegree = rdegree = 0 (getter/spline.py:69-77)."""
    return nk - 1 + degree


def spline_interval_index(knots, x) -> np.ndarray:
    """Knot-interval index i with knots[i] <= x < knots[i+1]; the right end
    point (and anything beyond either end) is clamped into the first / last
    interval.  This is the integer the GPU lookup must match bit-exactly
    (SURVEY.md 8c(iii))."""
    knots = np.asarray(knots, float)
    idx = np.searchsorted(knots, np.asarray(x, float), side="right") - 1
    return np.clip(idx, 0, knots.size - 2).astype(np.int32)


def _local_basis_ders(t, degree, mu, xe, order):
    """Non-zero values of the ``order``-th derivative of the degree-``degree``
    B-spline basis at ``xe`` for span ``mu`` (t[mu] <= xe <= t[mu+1]).
    Returns (len(xe), degree+1): entry j is basis index mu-degree+j.

    Cox-de Boor triangle at degree-order, then ``order`` differencing steps
    B'_{j,q} = q (B_{j,q-1}/(t_{j+q}-t_j) - B_{j+1,q-1}/(t_{j+q+1}-t_{j+1})).
    The CUDA kernel runs the same operations in the same order.
    """
    m = xe.size
    N = np.zeros((m, degree + 1))
    if order > degree:
        return N
    low = degree - order
    N[:, 0] = 1.0
    for j in range(1, low + 1):
        saved = np.zeros(m)
        for r in range(j):
            right = t[mu + r + 1] - xe
            left = xe - t[mu + 1 - j + r]
            temp = N[:, r] / (right + left)
            N[:, r] = saved + right * temp
            saved = left * temp
        N[:, j] = saved
    # N[:, 0..low] are degree-`low` values of basis indices mu-low .. mu
    for q in range(low + 1, degree + 1):
        # new active indices mu-q .. mu (q+1 of them); old: mu-q+1 .. mu
        new = np.zeros((m, degree + 1))
        for a in range(q + 1):
            jidx = mu - q + a
            val = np.zeros(m)
            if a >= 1:  # old index jidx is active (position a-1)
                val = val + N[:, a - 1] / (t[jidx + q] - t[jidx])
            if a <= q - 1:  # old index jidx+1 is active (position a)
                val = val - N[:, a] / (t[jidx + q + 1] - t[jidx + 1])
            new[:, a] = q * val
        N = new
    return N


def spline_design(knots, degree, x, order=0, ldegree=0, rdegree=0, return_index=False):
    """Dense (n, nk-1+degree) design matrix of the clamped B-spline basis, its
    ``order``-th derivative, with polynomial extrapolation of degree
    ``ldegree``/``rdegree`` outside [knots[0], knots[-1]].
    """
    knots = np.asarray(knots, float)
    x = np.asarray(x, float).ravel()
    nk = knots.size
    nb = spline_num_bases(nk, degree)
    t = np.concatenate([np.repeat(knots[0], degree), knots, np.repeat(knots[-1], degree)])
    idx = spline_interval_index(knots, x)
    mu = idx.astype(np.int64) + degree
    out = np.zeros((x.size, nb))
    left = x < knots[0]
    right = x > knots[-1]
    inside = ~(left | right)
    rows = np.arange(x.size)

    def scatter(sel, vals):
        r = rows[sel]
        for j in range(degree + 1):
            out[r, idx[sel] + j] += vals[:, j]

    if inside.any():
        scatter(inside, _local_basis_ders(t, degree, mu[inside], x[inside], order))
    for sel, edge, ext in ((left, knots[0], ldegree), (right, knots[-1], rdegree)):
        if not sel.any():
            continue
        dx = x[sel] - edge
        xe = np.full(dx.shape, edge)
        acc = np.zeros((dx.size, degree + 1))
        for mdeg in range(order, ext + 1):
            ders = _local_basis_ders(t, degree, mu[sel], xe, mdeg)
            acc = acc + ders * (dx ** (mdeg - order) / math.factorial(mdeg - order))[:, None]
        scatter(sel, acc)
    if return_index:
        return out, idx
    return out


def spline_knots(knots, knots_type, data):
    """Knot resolution of SplineGetter.get_spline (getter/spline.py:92-99)."""
    knots = np.asarray(knots, float)
    if knots_type == "abs":
        return knots
    if knots_type == "rel_domain":
        lb, ub = data.min(), data.max()
        return lb + knots * (ub - lb)
    if knots_type == "rel_freq":
        return np.quantile(data, knots)
    raise ValueError("Knots type must be one of 'abs', 'rel_domain' or 'rel_freq'.")


def spline_prior_points(knots, size=100, domain=(0.0, 1.0), domain_type="rel"):
    """Sample points of SplinePriorGetter.get_prior (getter/prior.py:185-190)."""
    lb, ub = knots[0], knots[-1]
    dlb, dub = domain
    if domain_type == "rel":
        dlb, dub = lb + (ub - lb) * dlb, lb + (ub - lb) * dub
    return np.linspace(dlb, dub, size)
