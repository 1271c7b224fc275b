"""Dev tool: cProfile of Model.attach on a pandas frame."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from anml_b200.data import Component, DataExample
from anml_b200.model import Model
from anml_b200.parameter import Expit, Parameter
from anml_b200.variable import Variable

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
p = 64
rng = np.random.default_rng(0)
df = pd.DataFrame({f"cov{j}": rng.standard_normal(n) for j in range(p - 1)})
df["obs"] = (rng.random(n) < 0.5).astype(float)
par = Parameter([Variable(Component("intercept", default_value=1.0))] + [Variable(f"cov{j}") for j in range(p - 1)], transform=Expit())
model = Model("binomial", DataExample("obs", "obs_se"), [par])
model.attach(df.copy())
for rep in range(2):
    t0 = time.perf_counter(); model.attach(df); print("attach_s", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); model.attach(df); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
