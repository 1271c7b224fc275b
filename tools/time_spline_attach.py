"""Dev tool: BASELINE config #3 attach (pandas column -> knots from data -> device basis) per knots_type."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from anml_b200.getter.prior import SplinePriorGetter
from anml_b200.getter.spline import SplineGetter
from anml_b200.parameter import Parameter
from anml_b200.prior import GaussianPrior
from anml_b200.variable import SplineVariable

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(3)
x = rng.random(n)
df = pd.DataFrame({"x": x})
out = {"n": n}
for kt in ("abs", "rel_domain", "rel_freq"):
    ts = []
    for rep in range(3):
        var = SplineVariable("x", SplineGetter(np.linspace(0, 1, 20), degree=3, knots_type=kt),
                             priors=[SplinePriorGetter(GaussianPrior(mean=0.0, sd=1.0), size=100, order=2)])
        par = Parameter([var])
        t0 = time.perf_counter(); par.attach(df); ts.append(time.perf_counter() - t0)
    out[kt + "_attach_s"] = [round(t, 4) for t in ts]
t0 = time.perf_counter(); lo, hi = x.min(), x.max(); out["numpy_minmax_s"] = round(time.perf_counter() - t0, 4)
t0 = time.perf_counter(); np.quantile(x, np.linspace(0, 1, 20)); out["numpy_quantile_s"] = round(time.perf_counter() - t0, 4)
print(json.dumps(out))
