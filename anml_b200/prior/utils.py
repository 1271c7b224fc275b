"""Prior lookup helpers (reference: src/anml/prior/utils.py:7-64)."""
from __future__ import annotations

from typing import List, Optional, Type

from . import main as _prior_main
from .main import Prior


def get_prior_type(prior_type: str) -> Type:
    """Class named ``prior_type`` in :mod:`anml_b200.prior.main` (the reference
    resolves the name the same way, utils.py:27; unknown names raise
    ``AttributeError``)."""
    return getattr(_prior_main, prior_type)


def filter_priors(priors: List[Prior], prior_type: str, with_mat: Optional[bool] = None) -> List[Prior]:
    """Members of ``priors`` that are instances of ``prior_type`` and, when
    ``with_mat`` is given, do (True) / do not (False) carry a linear map."""
    cls = get_prior_type(prior_type)
    picked = []
    for prior in priors:
        if not isinstance(prior, cls):
            continue
        if with_mat is None or (prior.mat is not None) == bool(with_mat):
            picked.append(prior)
    return picked
