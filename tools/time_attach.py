"""Dev tool: ingest (attach) and small-n API timings."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from anml_b200.data import Component, DataExample
from anml_b200.model import Model
from anml_b200.parameter import Expit, Parameter
from anml_b200.variable import Variable

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
p = 64
rng = np.random.default_rng(0)
cols = {f"cov{j}": rng.standard_normal(n) for j in range(p - 1)}
df = pd.DataFrame(cols)
df["obs"] = (rng.random(n) < 0.5).astype(float)
par = Parameter([Variable(Component("intercept", default_value=1.0))] + [Variable(f"cov{j}") for j in range(p - 1)], transform=Expit())
model = Model("binomial", DataExample("obs", "obs_se"), [par])
model.attach(df.copy())  # warm-up (library load, context)
t0 = time.perf_counter(); model.attach(df); t1 = time.perf_counter()
x = rng.standard_normal(p) * 0.05
model.evaluate(x)
t2 = time.perf_counter(); model.evaluate(x * 1.01); t3 = time.perf_counter()
print(json.dumps({"n": n, "p": p, "attach_s": t1 - t0, "attach_GBps": n * p * 8 / (t1 - t0) / 1e9, "rows_per_s": n / (t1 - t0), "eval_ms": (t3 - t2) * 1e3}))
