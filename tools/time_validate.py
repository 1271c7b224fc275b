import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from anml_b200 import _native as N
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
p = 64
rng = np.random.default_rng(0)
cols = [rng.standard_normal(n) for _ in range(p)]
out = {"n": n, "upload_s": [], "validate_s": [], "create_s": []}
d = None
for rep in range(5):
    t0 = time.perf_counter(); d = N.Design(n, p); out["create_s"].append(round(time.perf_counter() - t0, 4))
    t0 = time.perf_counter()
    for c in range(p): d.set_columns(c, cols[c].reshape(-1, 1))
    out["upload_s"].append(round(time.perf_counter() - t0, 4))
    for k in range(2):
        t0 = time.perf_counter(); f = d.validate_columns(0, p); out["validate_s"].append(round(time.perf_counter() - t0, 4))
out["flags"] = [int(f.min()), int(f.max())]
print(json.dumps(out))
