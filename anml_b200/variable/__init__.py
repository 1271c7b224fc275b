"""Variables: a column (or a spline basis) of the design matrix plus its priors."""
from .main import Variable
from .spline import SplineVariable

__all__ = ["Variable", "SplineVariable"]
