"""Generate golden vectors by running the REAL reference (imported from
/root/reference/src) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/anml_ref_golden.npz.  The reference cannot travel to the
GPU box, the fixtures can.  The spline modules of the reference import the
third-party ``xspline`` package, which is not installed here, so no spline
output of the reference can be recorded (spline parity stays unpinned).
"""
import os
import zlib
import sys

import numpy as np
import pandas as pd

REF = "/root/reference/src"
sys.path.insert(0, REF)

from anml.data.component import Component  # noqa: E402
from anml.parameter.main import Parameter  # noqa: E402
from anml.parameter.smoothmapping import Exp, Expit, Identity, Log, Logit  # noqa: E402
from anml.prior.main import GaussianPrior, UniformPrior  # noqa: E402
from anml.variable.main import Variable  # noqa: E402
from scipy.optimize import Bounds, minimize  # noqa: E402

OUT = {}
LINKS = {"identity": Identity, "exp": Exp, "log": Log, "expit": Expit, "logit": Logit}


def put(name, value):
    OUT[name] = np.asarray(value)


# ---- 1. smooth mappings (smoothmapping.py) -------------------------------
rng = np.random.default_rng(20261019)
grid = {
    "identity": rng.normal(size=64) * 3,
    "exp": rng.normal(size=64) * 3,
    "expit": np.concatenate([rng.normal(size=60) * 4, [-30.0, 30.0, 0.0, -700.0]]),
    "log": rng.uniform(1e-3, 50, size=64),
    "logit": rng.uniform(1e-6, 1 - 1e-6, size=64),
}
for name, cls in LINKS.items():
    put(f"link/{name}/x", grid[name])
    for order in (0, 1, 2):
        with np.errstate(all="ignore"):
            put(f"link/{name}/order{order}", cls()(grid[name].copy(), order=order))

# ---- 2. Parameter.get_params orders 0/1/2 (parameter/main.py:229-280) ------
def make_frame(n, p, seed, with_offset):
    r = np.random.default_rng(seed)
    cols = {f"cov{j}": r.normal(size=n) for j in range(p - 1)}
    df = pd.DataFrame(cols)
    if with_offset:
        df["off"] = r.normal(size=n) * 0.3
    return df


cases = [
    ("p5_identity", 40, 5, "identity", False),
    ("p5_exp_off", 40, 5, "exp", True),
    ("p7_expit", 33, 7, "expit", False),
    ("p17_expit_off", 70, 17, "expit", True),
    ("p64_expit", 48, 64, "expit", False),
]
for name, n, p, link, with_offset in cases:
    df = make_frame(n, p, zlib.crc32(name.encode()) % 1000 + 7, with_offset)
    variables = [Variable(Component("intercept", default_value=1.0))] + [Variable(f"cov{j}") for j in range(p - 1)]
    par = Parameter(variables, transform=LINKS[link](), offset="off" if with_offset else None)
    par.attach(df)
    beta = np.random.default_rng(5).normal(size=p) * 0.3
    put(f"params/{name}/X", par.design_mat)
    put(f"params/{name}/beta", beta)
    put(f"params/{name}/link", link)
    if with_offset:
        put(f"params/{name}/offset", df["off"].to_numpy())
    for order in (0, 1, 2):
        put(f"params/{name}/order{order}", par.get_params(beta, order=order))

# ---- 3. priors (prior/main.py:178-197; parameter/main.py:167-340) ----------
r = np.random.default_rng(11)
x3 = r.normal(size=3)
gp = GaussianPrior(mean=[0.1, -0.2, 0.3], sd=[1.0, 0.5, 2.0])
put("prior/direct/x", x3)
put("prior/direct/obj", gp.objective(x3)); put("prior/direct/grad", gp.gradient(x3)); put("prior/direct/hess", gp.hessian(x3))
M = r.normal(size=(5, 3))
gl = GaussianPrior(mean=r.normal(size=5), sd=r.uniform(0.5, 2, size=5), mat=M)
put("prior/linear/mat", M); put("prior/linear/mean", gl.mean); put("prior/linear/sd", gl.sd)
put("prior/linear/obj", gl.objective(x3)); put("prior/linear/grad", gl.gradient(x3)); put("prior/linear/hess", gl.hessian(x3))

# Parameter-level assembly: variable direct priors + variable linear + parameter linear priors
dfp = make_frame(20, 4, 3, False)
vs = [
    Variable(Component("intercept", default_value=1.0), priors=[GaussianPrior(mean=0.5, sd=2.0), UniformPrior(lb=-3.0, ub=3.0)]),
    Variable("cov0", priors=[GaussianPrior(mean=[0.0, 1.0], sd=[1.0, 3.0], mat=np.array([[1.0], [2.0]]))]),
    Variable("cov1"),
    Variable("cov2", priors=[GaussianPrior(mean=-0.3, sd=0.7)]),
]
Mp = r.normal(size=(2, 4))
par = Parameter(vs, priors=[GaussianPrior(mean=[0.2, -0.1], sd=[0.9, 1.1], mat=Mp), UniformPrior(lb=[-1.0], ub=[1.0], mat=np.ones((1, 4))),
                            GaussianPrior(mean=np.zeros(4), sd=np.ones(4))])  # direct prior at Parameter level: ignored
par.attach(dfp)
x4 = r.normal(size=4)
put("prior/param/x", x4); put("prior/param/Mp", Mp)
put("prior/param/obj", par.prior_objective(x4)); put("prior/param/grad", par.prior_gradient(x4)); put("prior/param/hess", par.prior_hessian(x4))
for cat in ("direct", "linear"):
    for typ in ("GaussianPrior", "UniformPrior"):
        pr = par.prior_dict[cat][typ]
        put(f"prior/param/{cat}/{typ}/params", pr.params)
        if pr.mat is not None:
            put(f"prior/param/{cat}/{typ}/mat", pr.mat)

# ---- 4. model layer composed from the reference's own calls ----
def compose(family, obs, obs_se, pars, x):
    sizes = [p.size for p in pars]
    st = np.concatenate([[0], np.cumsum(sizes)])
    xs = [x[st[k]:st[k + 1]] for k in range(len(pars))]
    th = [p.get_params(xs[k], order=0) for k, p in enumerate(pars)]
    J = [p.get_params(xs[k], order=1) for k, p in enumerate(pars)]
    T = [p.get_params(xs[k], order=2) for k, p in enumerate(pars)]
    if family == "gaussian":
        mu = th[0]; l = 0.5 * ((obs - mu) / obs_se) ** 2; dl = [-(obs - mu) / obs_se**2]; d2 = {(0, 0): 1 / obs_se**2}
    elif family == "binomial":
        p_ = th[0]; l = -(obs * np.log(p_) + (1 - obs) * np.log(1 - p_)); dl = [-obs / p_ + (1 - obs) / (1 - p_)]
        d2 = {(0, 0): obs / p_**2 + (1 - obs) / (1 - p_) ** 2}
    elif family == "poisson":
        lam = th[0]; l = lam - obs * np.log(lam); dl = [1 - obs / lam]; d2 = {(0, 0): obs / lam**2}
    else:
        mu, v = th; res = obs - mu
        l = 0.5 * np.log(v) + 0.5 * res**2 / v
        dl = [-res / v, 0.5 / v - 0.5 * res**2 / v**2]
        d2 = {(0, 0): 1 / v, (0, 1): res / v**2, (1, 1): -0.5 / v**2 + res**2 / v**3}
    P = st[-1]
    obj = l.sum(); g = np.zeros(P); H = np.zeros((P, P))
    for k, p in enumerate(pars):
        sk = slice(st[k], st[k + 1])
        obj += p.prior_objective(xs[k])
        g[sk] = J[k].T @ dl[k] + p.prior_gradient(xs[k])
        for m in range(k, len(pars)):
            sm = slice(st[m], st[m + 1])
            blk = (J[k].T * d2[(k, m)]) @ J[m]
            if k == m:
                blk = blk + np.tensordot(dl[k], T[k], axes=1) + p.prior_hessian(xs[k])
            H[sk, sm] = blk
            if k != m:
                H[sm, sk] = blk.T
    return obj, g, H


def model_case(name, family, n, p, links, seed, prior_sd=None, offset=False):
    r = np.random.default_rng(seed)
    df = make_frame(n, p, seed + 1, offset)
    pars = []
    for link in links:
        pri = [GaussianPrior(mean=0.0, sd=prior_sd)] if prior_sd else None
        variables = [Variable(Component("intercept", default_value=1.0), priors=pri)] + [Variable(f"cov{j}", priors=pri) for j in range(p - 1)]
        pars.append(Parameter(variables, transform=LINKS[link](), offset="off" if offset else None))
        pars[-1].attach(df)
    X = pars[0].design_mat
    K = len(links)
    beta_true = r.normal(size=K * p) * 0.2
    y0 = X @ beta_true[:p]
    if family == "gaussian":
        obs = y0 + 0.1 * r.normal(size=n); se = r.uniform(0.5, 1.5, size=n)
    elif family == "binomial":
        obs = (r.uniform(size=n) < 1 / (1 + np.exp(-y0))).astype(float); se = np.ones(n)
    elif family == "poisson":
        obs = r.poisson(np.exp(y0)).astype(float); se = np.ones(n)
    else:
        obs = y0 + np.exp(0.5 * (X @ beta_true[p:])) * r.normal(size=n); se = np.ones(n)
    x = 0.5 * beta_true
    obj, g, H = compose(family, obs, se, pars, x)
    put(f"model/{name}/family", family); put(f"model/{name}/links", ",".join(links))
    put(f"model/{name}/X", X); put(f"model/{name}/obs", obs); put(f"model/{name}/se", se); put(f"model/{name}/x", x)
    if offset:
        put(f"model/{name}/offset", df["off"].to_numpy())
    put(f"model/{name}/prior_sd", prior_sd if prior_sd else 0.0)
    put(f"model/{name}/obj", obj); put(f"model/{name}/grad", g); put(f"model/{name}/hess", H)
    return pars, obs, se, df


model_case("gauss_p4", "gaussian", 300, 4, ["identity"], 101)
model_case("binom_p6", "binomial", 257, 6, ["expit"], 102, prior_sd=1.0)
model_case("binom_p64", "binomial", 130, 64, ["expit"], 103, prior_sd=1.0)
model_case("poisson_p5_off", "poisson", 199, 5, ["exp"], 104, offset=True)
model_case("meanvar_p5", "gaussian_meanvar", 211, 5, ["identity", "exp"], 105)
model_case("gauss_exp_p3", "gaussian", 64, 3, ["exp"], 106)

# ---- 5. config C1: trust-constr fit with reference-composed callbacks --------
n, p = 10_000, 4
r = np.random.default_rng(20261019 + 1)
Xc = r.normal(size=(n, p - 1))
beta_true = r.normal(size=p) * 0.1
df = pd.DataFrame({f"cov{j}": Xc[:, j] for j in range(p - 1)})
obs = beta_true[0] + Xc @ beta_true[1:] + 0.1 * r.normal(size=n)
par = Parameter([Variable(Component("intercept", default_value=1.0))] + [Variable(f"cov{j}") for j in range(p - 1)])
par.attach(df)
se = np.ones(n)
fun = lambda x: compose("gaussian", obs, se, [par], x)[0]
jac = lambda x: compose("gaussian", obs, se, [par], x)[1]
hes = lambda x: compose("gaussian", obs, se, [par], x)[2]
res = minimize(fun, np.zeros(p), jac=jac, hess=hes, method="trust-constr", options={"gtol": 1e-10, "xtol": 1e-14, "maxiter": 200})
put("c1/seed", 20261019 + 1); put("c1/beta_true", beta_true); put("c1/coef", res.x); put("c1/niter", res.nit)
put("c1/obj_at_zero", fun(np.zeros(p))); put("c1/grad_at_zero", jac(np.zeros(p)))

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anml_ref_golden.npz")
np.savez_compressed(path, **OUT)
print("wrote", path, len(OUT), "arrays", os.path.getsize(path), "bytes; C1 fit nit", res.nit, "coef err", np.abs(res.x - beta_true).max())
