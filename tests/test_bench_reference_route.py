"""bench.py's CPU arm drives the REFERENCE's own classes (baseline/_ref, the pip-installed unmodified
package).  Here its arithmetic is checked against the oracle on a small sample, so that the timed route
is known to compute the same objective / gradient / Hessian as the GPU path it is quoted beside."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import anml_oracle as O  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "anml")),
                               reason="baseline/_ref (pip install of the reference) not present")


@needs_ref
@pytest.mark.parametrize("name", ["northstar", "c4"])
def test_reference_route_matches_the_oracle(name):
    cfg = dict(bench.CONFIGS[name], n=3001, p=7)
    route, x0 = bench.cpu_route(cfg, 3001, seed=5)
    assert route.kind == "reference"
    assert type(route.pars[0]).__module__ == "anml.parameter.main"  # the reference's class, not anml_b200's
    got = route.evaluate(x0, chunk=1000)
    data = bench.cpu_sample_data(cfg, 3001, seed=5)
    if name == "northstar":
        want = O.model_eval_fused("binomial", data["obs"], [data["X"]], ["expit"], x0, priors=[{"direct": (np.zeros(7), np.ones(7))}])
    else:
        want = O.model_eval_fused("gaussian_meanvar", data["obs"], [data["X"], data["X"]], ["identity", "exp"], x0)
    assert abs(got[0] - want[0]) <= 1e-12 * abs(want[0])
    assert np.abs(got[1] - want[1]).max() <= 1e-11 * np.abs(want[1]).max()
    assert np.abs(got[2] - want[2]).max() <= 1e-11 * np.abs(want[2]).max()


@needs_ref
def test_reference_literal_route_runs():
    cfg = dict(bench.CONFIGS["northstar"], n=500, p=5)
    route, x0 = bench.cpu_route(cfg, 500, seed=6)
    out = route.literal_rows_per_s(x0, 200)
    assert out["rows"] == 200 and out["rows_per_s"] > 0


def test_spline_config_uses_the_oracle_port():
    cfg = dict(bench.CONFIGS["c3"], n=2000)
    route, x0 = bench.cpu_route(cfg, 2000, seed=7)
    assert route.kind == "port"  # xspline is absent: the reference's spline classes cannot be imported
    o, g, h = route.evaluate(x0)
    assert np.isfinite(o) and g.shape == (22,) and h.shape == (22, 22) and np.array_equal(h, h.T)


def test_both_arms_emit_the_same_config_object():
    for name, cfg in bench.CONFIGS.items():
        for world in (1, 2, 8):
            assert bench.config_dict(cfg, name, world) == bench.config_dict(dict(cfg), name, world)
