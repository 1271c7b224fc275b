"""A variable whose columns are a B-spline basis (reference:
src/anml/variable/spline.py:57-113).  The basis is built on the device."""
from __future__ import annotations

from typing import List, Optional, Tuple, Type, Union

from numpy.typing import NDArray
from pandas import DataFrame

from ..data.component import Component
from ..getter.prior import SplinePriorGetter
from ..getter.spline import SplineGetter
from ..prior.main import Prior
from ..xspline import XSpline
from .main import Variable

SplineVariablePrior = Union[Prior, SplinePriorGetter]

# below this many rows NumPy resolves data-dependent knots faster than a device round trip
_DEVICE_KNOTS_MIN_ROWS = 1 << 16


class SplineVariable(Variable):
    """``spline`` is an ``XSpline`` or a ``SplineGetter``; the getter becomes an
    ``XSpline`` (and every ``SplinePriorGetter`` in ``priors`` a ``Prior``) in
    place at the first ``attach`` -- knots come from the first frame only
    (spline.py:103-107)."""

    _prior_types: Tuple[Type, ...] = (Prior, SplinePriorGetter)
    _resolves_on_device = True  # Parameter.attach leaves the first attach to _write_design

    def __init__(self, component: Union[str, Component], spline: Union[XSpline, SplineGetter],
                 priors: Optional[List[SplineVariablePrior]] = None):
        super().__init__(component, priors)
        self.spline = spline
        self.last_interval_index = None
        self.device_abscissae = None  # DeviceBuffer with the attached column, set by Parameter.attach

    @property
    def spline(self) -> Union[XSpline, SplineGetter]:
        return self._spline

    @spline.setter
    def spline(self, spline: Union[XSpline, SplineGetter]):
        if not isinstance(spline, (XSpline, SplineGetter)):
            raise TypeError("Spline variable input spline must be an instance of XSpline or SplineGetter.")
        self._spline = spline

    @property
    def size(self) -> int:
        if isinstance(self.spline, XSpline):  # spline.py:78-84
            ld, rd = self.spline.ldegree or 0, self.spline.rdegree or 0
            inner = self.spline.knots[ld: len(self.spline.knots) - rd]
            return len(inner) - 1 + self.spline.degree
        if isinstance(self.spline, SplineGetter):
            return self.spline.num_spline_bases
        raise TypeError("Unknown spline type")

    def attach(self, df: DataFrame) -> None:
        self.component.attach(df)
        if isinstance(self.spline, SplineGetter):
            self.spline = self.spline.get_spline(self.component.value)
        for i, prior in enumerate(self.priors):
            if isinstance(prior, SplinePriorGetter):
                self.priors[i] = prior.get_prior(self.spline)

    def get_design_mat(self, df: DataFrame) -> NDArray:
        """(n, size) basis matrix, evaluated by the CUDA spline kernel."""
        self.attach(df)
        return self.spline.get_design_mat(self.component.value)

    def _device_checks(self):
        return False, False  # the component's validators already ran on the host column in attach()

    def _write_design(self, design, col0: int, df: DataFrame) -> None:
        """Upload the abscissae once; knots that depend on the data (first attach of a
        ``SplineGetter``) are resolved from that device copy, then the basis is built from it."""
        from .. import _native

        self.component.attach(df)
        x = _native.as_f64(self.component.value, self.component.key)
        self.device_abscissae = None  # (a deferred design of the previous attach may still borrow the old copy)
        xbuf = _native.DeviceBuffer(design.device, x)
        try:
            if isinstance(self.spline, SplineGetter):
                if x.size >= _DEVICE_KNOTS_MIN_ROWS:
                    self.spline = self.spline.get_spline_device(xbuf.ptr, x.size, design.device)
                else:
                    self.spline = self.spline.get_spline(x)
            for i, prior in enumerate(self.priors):
                if isinstance(prior, SplinePriorGetter):
                    self.priors[i] = prior.get_prior(self.spline)
            sp = self.spline
            if sp.num_spline_bases != self.size:
                raise NotImplementedError("ldegree/rdegree that change the basis count are not supported")
            self.last_interval_index = design.set_spline(col0, None, sp.knots, sp.degree, sp.ldegree or 0, sp.rdegree or 0, 0,
                                                         want_index=False, x_dev_ptr=xbuf.ptr, keepalive=xbuf)
        except BaseException:
            xbuf.free()
            raise
        # kept for the model layer: a one-spline model is evaluated from the abscissae (8 bytes per row)
        self.device_abscissae = xbuf
