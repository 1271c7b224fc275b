"""Drop-in aliasing: make ``import anml`` resolve to this package.

``install_as_anml()`` registers ``anml`` and its sub-modules
(``anml.data.component``, ``anml.prior.main``, ...) in ``sys.modules`` as
aliases of the matching ``anml_b200`` modules, so code written against the
reference -- including the reference's own test-suite -- runs unchanged on the
B200 path.  ``xspline`` can be aliased as well (the reference imports
``from xspline import XSpline``).
"""
from __future__ import annotations

import importlib
import sys

_SUBMODULES = [
    "data", "data.component", "data.validator", "data.prototype", "data.example",
    "prior", "prior.main", "prior.utils",
    "getter", "getter.spline", "getter.prior",
    "variable", "variable.main", "variable.spline",
    "parameter", "parameter.main", "parameter.smoothmapping",
]


def install_as_anml(alias_xspline: bool = True) -> None:
    import anml_b200

    sys.modules["anml"] = anml_b200
    for name in _SUBMODULES:
        mod = importlib.import_module(f"anml_b200.{name}")
        sys.modules[f"anml.{name}"] = mod
    if alias_xspline and "xspline" not in sys.modules:
        sys.modules["xspline"] = importlib.import_module("anml_b200.xspline")
