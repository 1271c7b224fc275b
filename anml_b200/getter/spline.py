"""Knot placement from data (reference: src/anml/getter/spline.py:47-106)."""
from __future__ import annotations

from typing import Optional

import numpy as np
from numpy.typing import NDArray

from ..xspline import XSpline

_KNOT_TYPES = ("abs", "rel_domain", "rel_freq")


class SplineGetter:
    """Settings of a spline whose knots are resolved against data on attach.

    ``knots_type``: ``'abs'`` uses ``knots`` as given; ``'rel_domain'`` places
    them at ``min + k (max - min)`` of the data; ``'rel_freq'`` at the data's
    ``k``-quantiles (NumPy's default linear interpolation).
    ``ldegree`` / ``rdegree`` default to 0 and are capped at ``len(knots) - 2``
    (spline.py:57-58).
    """

    def __init__(self, knots: NDArray, degree: int = 3, ldegree: Optional[int] = None, rdegree: Optional[int] = None,
                 knots_type: str = "abs"):
        self.knots = knots
        self.degree = degree
        cap = len(knots) - 2
        self.ldegree = min(0 if ldegree is None else ldegree, cap)
        self.rdegree = min(0 if rdegree is None else rdegree, cap)
        self.knots_type = knots_type

    @property
    def knots_type(self) -> str:
        return self._knots_type

    @knots_type.setter
    def knots_type(self, knots_type: str):
        if knots_type not in _KNOT_TYPES:
            raise ValueError("Knots type must be one of 'abs', 'rel_domain' or 'rel_freq'.")
        self._knots_type = knots_type

    @property
    def num_spline_bases(self) -> int:
        # spline.py:69-77 (the ldegree/rdegree subtraction is a leftover of xspline's
        # old l_linear/r_linear API; with the defaults it is len(knots) - 1 + degree)
        ld, rd = self.ldegree or 0, self.rdegree or 0
        return len(self.knots[ld: len(self.knots) - rd]) - 1 + self.degree

    def resolve_knots(self, data: NDArray) -> NDArray:
        if self.knots_type == "abs":
            return self.knots
        if self.knots_type == "rel_domain":
            lo, hi = data.min(), data.max()
            return lo + self.knots * (hi - lo)
        return np.quantile(data, self.knots)

    def resolve_knots_device(self, dev_ptr: int, n: int, device: int = 0) -> NDArray:
        """``resolve_knots`` for ``n`` abscissae that already sit in device memory: the
        min/max pass or the sort behind the quantiles runs there (same values as the
        host route: min/max are exact, the quantile interpolation is NumPy's formula)."""
        from .. import _native

        if self.knots_type == "abs":
            return self.knots
        if self.knots_type == "rel_domain":
            lo, hi = _native.vector_minmax(dev_ptr=dev_ptr, n=n, device=device)
            lo, hi = np.float64(lo), np.float64(hi)
            return lo + self.knots * (hi - lo)
        return _native.vector_quantile(self.knots, dev_ptr=dev_ptr, n=n, device=device)

    def get_spline_device(self, dev_ptr: int, n: int, device: int = 0) -> XSpline:
        return XSpline(self.resolve_knots_device(dev_ptr, n, device), self.degree, ldegree=self.ldegree, rdegree=self.rdegree)

    def get_spline(self, data: NDArray) -> XSpline:
        return XSpline(self.resolve_knots(data), self.degree, ldegree=self.ldegree, rdegree=self.rdegree)
