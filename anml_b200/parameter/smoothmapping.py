"""Smooth link functions with first and second derivatives.

Reference behaviour: src/anml/parameter/smoothmapping.py:9-156 (same class
names, ``f(x, order=0|1|2)`` call surface, ``.inverse``, error messages).

These host-side callables exist for user code and for complex-step checks.
The evaluation path does NOT go through them: ``Parameter.get_params`` and the
model kernels apply the same formulas on the device
(``anml_b200/csrc/rowmath.cuh::link_eval``), selected by ``device_code``.

Layout differs from the reference on purpose: one ``__call__`` in the base
class validates ``order`` and the domain, then dispatches to the subclass's
``_value`` / ``_first`` / ``_second``.
"""
from __future__ import annotations

import numpy as np
from numpy.typing import NDArray

_VALID_ORDERS = (0, 1, 2)


class SmoothMapping:
    """Base class; concrete mappings provide ``_value``, ``_first``, ``_second``."""

    device_code: int = -1  # ANML_LINK_* of include/anml_b200.h; -1: host only
    _inverse_name: str = ""
    _domain_message: str = ""

    def __init__(self):
        if type(self) is SmoothMapping:
            raise TypeError("SmoothMapping is abstract; use Identity, Exp, Log, Expit or Logit.")

    # -- hooks -------------------------------------------------------------------
    def _in_domain(self, x: NDArray) -> bool:
        return True

    def _value(self, x: NDArray) -> NDArray:
        raise NotImplementedError

    def _first(self, x: NDArray) -> NDArray:
        raise NotImplementedError

    def _second(self, x: NDArray) -> NDArray:
        raise NotImplementedError

    # -- public surface ------------------------------------------------------------
    def _validate_order(self, order: int = 0) -> None:
        if order not in _VALID_ORDERS:
            raise ValueError("Order must be 0, 1 or 2.")

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        if not self._in_domain(x):
            raise ValueError(self._domain_message)
        return (self._value, self._first, self._second)[order](x)

    @property
    def inverse(self) -> "SmoothMapping":
        if not self._inverse_name:
            raise NotImplementedError
        return globals()[self._inverse_name]()

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class Identity(SmoothMapping):
    device_code = 0
    _inverse_name = "Identity"

    def _value(self, x):
        return x  # the argument itself (no copy), like the reference

    def _first(self, x):
        return np.ones(np.shape(x))

    def _second(self, x):
        return np.zeros(np.shape(x))


class Exp(SmoothMapping):
    device_code = 1
    _inverse_name = "Log"

    def _value(self, x):
        return np.exp(x)

    _first = _value
    _second = _value


class Log(SmoothMapping):
    device_code = 2
    _inverse_name = "Exp"
    _domain_message = "All values for log function must be positive."

    def _in_domain(self, x):
        return bool(np.all(x > 0))

    def _value(self, x):
        return np.log(x)

    def _first(self, x):
        return 1 / x

    def _second(self, x):
        return -1 / x**2


class Expit(SmoothMapping):
    """1 / (1 + exp(-x)); derivatives written in z = exp(-x) exactly as the reference
    does, so overflow behaviour (nan at order >= 1 for x < -709.78) is the same."""

    device_code = 3
    _inverse_name = "Logit"

    def _value(self, x):
        return 1 / (1 + np.exp(-x))

    def _first(self, x):
        z = np.exp(-x)
        return z / (1 + z) ** 2

    def _second(self, x):
        z = np.exp(-x)
        return z * (z - 1) / (z + 1) ** 3


class Logit(SmoothMapping):
    device_code = 4
    _inverse_name = "Expit"
    _domain_message = "All values for logit function must be strictly between 0 and 1."

    def _in_domain(self, x):
        return bool(np.all(x > 0) and np.all(x < 1))

    def _value(self, x):
        return np.log(x / (1 - x))

    def _first(self, x):
        return 1 / (x * (1 - x))

    def _second(self, x):
        return (2 * x - 1) / (x * (1 - x)) ** 2
