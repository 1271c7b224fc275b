"""Pin the spline rows against the REAL xspline -- the day it is importable.

`xspline` is a third-party dependency of the reference (pyproject.toml:22, unpinned); its source is not in
/root/reference, it is not installed in this image and there is no network, so today the B-spline
restatement (oracle/anml_oracle.py::spline_design, anml_b200/csrc/spline.cuh) is "parity unpinned".  Where
`import xspline` works (and /root/reference, or a pip-installed `anml`, is importable), run

    python tests/golden/make_spline_golden.py

It writes tests/golden/xspline_golden.npz with, for a grid of (knots, degree, ldegree, rdegree, order):
  * XSpline(knots, degree, ldegree=, rdegree=).get_design_mat(x, order=)   (call sites getter/spline.py:101-106,
    variable/spline.py:111-113, getter/prior.py:191-194), x inside AND outside the knot span,
  * the reference's SplineGetter.get_spline knots for 'abs' / 'rel_domain' / 'rel_freq' (getter/spline.py:92-99),
  * the reference's SplinePriorGetter.get_prior matrices (getter/prior.py:163-195),
  * the reference's SplineVariable.get_design_mat on a frame (variable/spline.py:109-113).
tests/test_oracle_golden.py (oracle, CPU) and tests/test_gpu_spline.py (device) consume the file when it
exists and skip otherwise: the pin closes without a code change."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "xspline_golden.npz")
REF_SRC = "/root/reference/src"

CASES = [  # (name, knots, degree, ldegree, rdegree)
    ("c3", np.linspace(0.0, 1.0, 20), 3, 0, 0),
    ("d1", np.array([0.0, 0.3, 1.0]), 1, 0, 0),
    ("d2_uneven", np.array([-1.0, -0.2, 0.1, 0.7, 2.0]), 2, 0, 0),
    ("d3_lin_ext", np.linspace(0.0, 1.0, 6), 3, 1, 1),
    ("d3_mixed_ext", np.linspace(-2.0, 3.0, 7), 3, 2, 0),
    ("two_knots", np.array([0.0, 1.0]), 3, 0, 0),
]


def main():
    try:
        from xspline import XSpline
    except Exception as exc:
        raise SystemExit(f"xspline is not importable here ({exc!r}); nothing written")
    if os.path.isdir(REF_SRC) and REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import pandas as pd
    from anml.getter.prior import SplinePriorGetter
    from anml.getter.spline import SplineGetter
    from anml.prior.main import GaussianPrior
    from anml.variable.spline import SplineVariable

    rng = np.random.default_rng(20261019)
    out = {"cases": np.array([c[0] for c in CASES])}
    for name, knots, degree, ldeg, rdeg in CASES:
        lo, hi = knots[0], knots[-1]
        inside = np.concatenate([knots, rng.uniform(lo, hi, 200), [np.nextafter(hi, lo), np.nextafter(lo, hi)]])
        outside = np.concatenate([lo - rng.uniform(0.01, 1.0, 20), hi + rng.uniform(0.01, 1.0, 20)])
        x = np.concatenate([inside, outside])
        sp = XSpline(knots, degree, ldegree=ldeg, rdegree=rdeg)
        out[f"{name}/knots"], out[f"{name}/x"] = knots, x
        out[f"{name}/params"] = np.array([degree, ldeg, rdeg])
        out[f"{name}/n_inside"] = np.array(inside.size)
        for order in range(0, degree + 1):
            out[f"{name}/order{order}"] = np.asarray(sp.get_design_mat(x, order=order))
    data = rng.gamma(2.0, 1.0, 5001)
    for kt in ("abs", "rel_domain", "rel_freq"):
        g = SplineGetter(np.linspace(0.0, 1.0, 7), degree=3, knots_type=kt)
        out[f"getter/{kt}/knots"] = np.asarray(g.get_spline(data).knots)
    out["getter/data"] = data
    sp = SplineGetter(np.linspace(0.0, 1.0, 7), degree=3, knots_type="rel_domain").get_spline(data)
    for order in (0, 1, 2):
        pr = SplinePriorGetter(GaussianPrior(0.0, 1.0), size=50, order=order).get_prior(sp)
        out[f"priorgetter/order{order}/mat"] = np.asarray(pr.mat)
    sv = SplineVariable("v", SplineGetter(np.linspace(0.0, 1.0, 7), degree=3, knots_type="rel_freq"))
    out["variable/design"] = np.asarray(sv.get_design_mat(pd.DataFrame({"v": data})))
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays")


if __name__ == "__main__":
    main()
