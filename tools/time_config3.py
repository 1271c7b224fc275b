"""Dev tool: BASELINE config #3 through the public API: pandas frame -> Model.attach -> evaluate."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from anml_b200.data import DataExample
from anml_b200.getter.prior import SplinePriorGetter
from anml_b200.getter.spline import SplineGetter
from anml_b200.model import Model
from anml_b200.parameter import Parameter
from anml_b200.prior import GaussianPrior
from anml_b200.variable import SplineVariable

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(20261022)
x = rng.random(n)
df = pd.DataFrame({"x": x, "obs": np.sin(2 * np.pi * x) + 0.1 * rng.standard_normal(n)})
out = {"n": n}
for banded in (True, False):
    var = SplineVariable("x", SplineGetter(np.linspace(0, 1, 20), degree=3, knots_type="rel_domain"),
                         priors=[SplinePriorGetter(GaussianPrior(mean=0.0, sd=1.0), size=100, order=2)])
    model = Model("gaussian", DataExample("obs", "obs_se"), [Parameter([var])])
    model.use_banded_spline = banded
    t0 = time.perf_counter(); model.attach(df); t_attach = time.perf_counter() - t0
    t0 = time.perf_counter(); model.attach(df); t_attach2 = time.perf_counter() - t0
    b = np.linspace(-1, 1, model.size)
    for k in range(3): model.evaluate(b * (1 + 1e-3 * k))
    t0 = time.perf_counter()
    for k in range(10): o, g, H = model.evaluate(b * (1.01 + 1e-3 * k))
    t_eval = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter(); res = model.fit(np.zeros(model.size), options={"gtol": 1e-8}); t_fit = time.perf_counter() - t0
    out["banded" if banded else "dense"] = {"uses_banded": model.banded_spline, "attach_s": round(t_attach, 4), "reattach_s": round(t_attach2, 4),
                                              "eval_ms": round(t_eval * 1e3, 3), "fit_s": round(t_fit, 3), "fit_evals": model.num_evaluations, "fit_obj": float(res.fun)}
print(json.dumps(out))
