"""Device-backed stand-in for ``xspline.XSpline``.

anml builds spline design matrices through the third-party ``xspline`` package
(``XSpline(knots, degree, ldegree=, rdegree=)`` and
``.get_design_mat(x, order=)``; call sites getter/spline.py:101-106,
variable/spline.py:111-113, getter/prior.py:191-194).  This class offers the
same surface and evaluates the basis with the CUDA spline kernel
(``anml_b200/csrc/spline.cuh``): knot-interval lookup + Cox-de Boor on the
clamped knot vector, ``order``-th derivative by differencing, and Taylor
extrapolation of degree ``ldegree`` / ``rdegree`` outside the knot span.

Parity note: ``xspline`` is neither vendored in the reference nor installed
here, so values are pinned against the NumPy oracle and SciPy's ``BSpline``
instead ("parity unpinned" for splines).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native


class XSpline:
    def __init__(self, knots, degree: int, ldegree: Optional[int] = None, rdegree: Optional[int] = None):
        knots = np.asarray(knots, dtype=np.float64).ravel()
        if knots.size < 2:
            raise ValueError("XSpline needs at least two knots.")
        if np.any(np.diff(knots) < 0):
            raise ValueError("XSpline knots must be sorted.")
        degree = int(degree)
        if degree < 0:
            raise ValueError("XSpline degree must be non-negative.")
        self.knots = knots
        self.degree = degree
        self.ldegree = None if ldegree is None else int(ldegree)
        self.rdegree = None if rdegree is None else int(rdegree)

    @property
    def num_spline_bases(self) -> int:
        return self.knots.size - 1 + self.degree

    def get_design_mat(self, x, order: int = 0, return_index: bool = False, device: int = 0):
        """(n, num_spline_bases) dense design matrix of the ``order``-th
        derivative of the basis at ``x``, computed on the GPU."""
        return _native.spline_design(x, self.knots, self.degree, self.ldegree or 0, self.rdegree or 0, int(order),
                                     want_index=return_index, device=device)

    def interval_index(self, x, device: int = 0) -> np.ndarray:
        """Knot-interval index of every ``x`` (int32), from the device lookup."""
        return self.get_design_mat(x, 0, return_index=True, device=device)[1]

    def __repr__(self) -> str:
        return (f"XSpline(knots={self.knots!r}, degree={self.degree}, ldegree={self.ldegree}, "
                f"rdegree={self.rdegree})")
