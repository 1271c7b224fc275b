"""Dev tool: where the host-visible time of one evaluation goes at the bench's workload (1 GPU)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from anml_b200.distributed import NativeLocalEvaluator, ShardedModel

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
cfg = dict(bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "northstar"], n=n)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
data = bench.make_shard(torch, dev, cfg, 0, n)
model, pars = bench.build_model(cfg, data, 0, 0, 1)
sharded = ShardedModel(NativeLocalEvaluator(model, dev))
xs = [data["x0"] * (1 + 1e-3 * k) for k in range(40)]
out = {"n": n}
def loop(name, fn, reps=10):
    for k in range(3): fn(xs[k])
    torch.cuda.synchronize(); model.native.kernel_time(reset=True)
    t0 = time.perf_counter()
    for k in range(reps): fn(xs[3 + k])
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    kms, nl, al = model.native.kernel_time(reset=True)
    out[name] = {"host_ms": round(ms, 4), "main_kernel_ms": round(kms / reps, 4), "launches": al / reps}
loop("collective_async_back_to_back", lambda x: sharded.evaluate_collective(x, True))
loop("native_eval_sync", lambda x: model.native.eval(x))
loop("model_evaluate", lambda x: model.evaluate(x))
loop("sharded_evaluate", lambda x: sharded.evaluate(x, True))
loop("collective_then_sync", lambda x: (sharded.evaluate_collective(x, True), torch.cuda.synchronize()))
loop("collective_async_back_to_back_again", lambda x: sharded.evaluate_collective(x, True))
print(json.dumps(out))
