"""Dev tool: host->device upload rate of anml_design_set_columns vs number of host threads."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from anml_b200 import _native as N

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
p = 64
rng = np.random.default_rng(0)
col = rng.standard_normal(n)
block = rng.standard_normal((n // 8, p))
d = N.Design(n, p)
db = N.Design(n // 8, p)
d.set_columns(0, col.reshape(-1, 1)); db.set_columns(0, block)
res = {"n": n, "column_GBps": {}, "block_GBps": {}}
for rep in range(3):
    for t in (1, 2, 3, 4, 6, 8, 12, 16):
        os.environ["ANML_B200_UPLOAD_THREADS"] = str(t)
        t0 = time.perf_counter()
        for c in range(4):
            d.set_columns(c, col.reshape(-1, 1))
        dt = (time.perf_counter() - t0) / 4
        res["column_GBps"].setdefault(t, []).append(round(n * 8 / dt / 1e9, 2))
        t0 = time.perf_counter()
        db.set_columns(0, block)
        dt = time.perf_counter() - t0
        res["block_GBps"].setdefault(t, []).append(round(block.nbytes / dt / 1e9, 2))
# sanity: what was uploaded is what is there
back = db.download()
res["block_roundtrip_exact"] = bool(np.array_equal(back, block))
print(json.dumps(res))
