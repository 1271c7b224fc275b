"""Parameters: design matrix + transform + priors (device-backed)."""
from .main import Parameter
from .smoothmapping import Exp, Expit, Identity, Log, Logit, SmoothMapping

__all__ = ["Parameter", "SmoothMapping", "Identity", "Exp", "Log", "Expit", "Logit"]
