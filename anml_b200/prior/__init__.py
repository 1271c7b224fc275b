"""Priors (host side; O(m p), dtype-generic)."""
from .main import GaussianPrior, Prior, UniformPrior
from .utils import filter_priors, get_prior_type

__all__ = ["GaussianPrior", "Prior", "UniformPrior", "filter_priors", "get_prior_type"]
