"""Column validators (reference: src/anml/data/validator.py:7-42).

A validator is called as ``validator(key, value)`` and raises ``ValueError``
when the column breaks its rule.  Host-side, O(n), run once per attach.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np


class Validator(ABC):
    """Callable check on one column; subclasses raise ``ValueError`` on failure."""

    @abstractmethod
    def __call__(self, key: str, value: np.ndarray) -> None:
        ...

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class NoNans(Validator):
    """No entry may be NaN."""

    def __call__(self, key: str, value: np.ndarray) -> None:
        if np.any(np.isnan(value)):
            raise ValueError(f"Column '{key}' contains nans.")


class Positive(Validator):
    """Every entry must be strictly positive."""

    def __call__(self, key: str, value: np.ndarray) -> None:
        if np.any(value <= 0):
            raise ValueError(f"Column '{key}' contains nonpositive numbers.")


class Unique(Validator):
    """No value may repeat."""

    def __call__(self, key: str, value: np.ndarray) -> None:
        if np.unique(value).size < np.shape(value)[0]:
            raise ValueError(f"Column '{key}' contains duplicated values.")
