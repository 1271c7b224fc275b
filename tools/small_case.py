import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from anml_b200 import _native as N
rng = np.random.default_rng(0)
n, p = 5000, 64
X = rng.normal(size=(n, p)); obs = (rng.uniform(size=n) < 0.5).astype(float)
d = N.Design(n, p); d.set_columns(0, X)
m = N.NativeModel(n, "binomial", 1); m.set_param(0, d, 3); m.set_obs(obs)
print(m.eval(rng.normal(size=p) * 0.1)[0])
