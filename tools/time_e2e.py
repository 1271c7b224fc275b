"""Dev tool: where the host-visible time of one evaluation goes at the bench's workload (1 GPU)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from anml_b200.data import Component, DataExample
from anml_b200.parameter import Expit, Parameter
from anml_b200.prior import GaussianPrior
from anml_b200.variable import Variable
from anml_b200.model import Model
from anml_b200.distributed import NativeLocalEvaluator, ShardedModel

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
p = 64
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
X, obs, beta_true = bench.make_shard(torch, dev, n, p, 0)
prior = [GaussianPrior(mean=0.0, sd=1.0)]
variables = [Variable(Component("intercept", default_value=1.0), priors=prior)] + [Variable(f"cov{j}", priors=prior) for j in range(p - 1)]
par = Parameter(variables, transform=Expit(), device=0); par.attach_device(X.data_ptr(), n, ld=p, keepalive=X)
model = Model("binomial", DataExample("obs", "obs_se"), [par], device=0); model.attach_device(n, obs.data_ptr(), keepalive=obs)
sharded = ShardedModel(NativeLocalEvaluator(model, dev))
xs = [0.5 * beta_true * (1 + 1e-3 * k) for k in range(40)]
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
print(json.dumps(out))
