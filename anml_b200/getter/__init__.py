"""Getters that turn settings + data into splines / spline priors."""
from .prior import SplinePriorGetter
from .spline import SplineGetter

__all__ = ["SplineGetter", "SplinePriorGetter"]
