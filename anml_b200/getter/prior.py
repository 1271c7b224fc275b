"""Linear priors on a spline (reference: src/anml/getter/prior.py:109-195).

``SplinePriorGetter(prior, size, order, domain, domain_type)`` samples ``size``
points over (a fraction of) the knot span and sets ``prior.mat`` to the
``order``-th derivative of the spline basis there: ``order=1`` with a
``UniformPrior(0, inf)`` is monotonicity, ``order=2`` with a ``GaussianPrior``
is the smoothing prior BASELINE.json calls "SplineGaussianPrior".
The matrix is built by the CUDA spline kernel (``XSpline.get_design_mat``).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ..prior.main import Prior
from ..xspline import XSpline


class SplinePriorGetter:
    def __init__(self, prior: Prior, size: int = 100, order: int = 0, domain: Tuple[float, float] = (0.0, 1.0),
                 domain_type: str = "rel"):
        self.prior = prior
        self.size = size
        self.order = order
        self.domain = domain
        self.domain_type = domain_type

    @property
    def prior(self) -> Prior:
        return self._prior

    @prior.setter
    def prior(self, prior: Prior):
        if not isinstance(prior, Prior):
            raise TypeError("Prior must be an instance of Prior.")
        if prior.mat is not None:
            raise ValueError("Prior mat exists, cannot assign other values.")
        self._prior = prior

    @property
    def size(self) -> int:
        return self._size

    @size.setter
    def size(self, size: int):
        size = int(size)
        if size <= 0:
            raise ValueError("Size must be positive.")
        self._size = size

    @property
    def order(self) -> int:
        return self._order

    @order.setter
    def order(self, order: int):
        order = int(order)
        if order < 0:
            raise ValueError("Order must be non-negative.")
        self._order = order

    @property
    def domain(self) -> Tuple[float, float]:
        return self._domain

    @domain.setter
    def domain(self, domain: Tuple[float, float]):
        domain = tuple(domain)
        if len(domain) != 2:
            raise ValueError("Domain must contains two numbers for lower and upper bound.")
        if domain[0] > domain[1]:
            raise ValueError("Domain lb must be less than or equal to ub.")
        self._domain = domain

    @property
    def domain_type(self) -> str:
        return self._domain_type

    @domain_type.setter
    def domain_type(self, domain_type: str):
        if domain_type not in ("rel", "abs"):
            raise ValueError("Domain type must be choicen from 'rel' or 'abs'.")
        self._domain_type = domain_type

    def sample_points(self, spline: XSpline) -> np.ndarray:
        lo, hi = spline.knots[0], spline.knots[-1]
        a, b = self.domain
        if self.domain_type == "rel":
            a, b = lo + (hi - lo) * a, lo + (hi - lo) * b
        return np.linspace(a, b, self.size)

    def get_prior(self, spline: XSpline) -> Prior:
        if not isinstance(spline, XSpline):
            raise TypeError("Spline must be an instance of XSpline.")
        self.prior.mat = spline.get_design_mat(self.sample_points(spline), order=self.order)
        return self.prior
