"""Dev tool: latency of cudaMalloc/cudaFree through the C-ABI for large buffers."""
import sys, os, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anml_b200 import _native as N
lib = N.load(); N.require_device(0)
out = {}
for gb in (0.4, 5.0, 20.0):
    ta, tf = [], []
    for rep in range(8):
        p = C.c_void_p()
        t0 = time.perf_counter(); N.check(lib.anml_dev_alloc(0, int(gb * 1e9), C.byref(p))); ta.append(round((time.perf_counter() - t0) * 1e3, 2))
        t0 = time.perf_counter(); N.check(lib.anml_dev_free(0, p)); tf.append(round((time.perf_counter() - t0) * 1e3, 2))
    out[f"{gb}GB"] = {"alloc_ms": ta, "free_ms": tf}
print(json.dumps(out))
