"""Smooth link functions with first and second derivatives (reference:
src/anml/parameter/smoothmapping.py:9-156).

These host-side callables keep the reference's call surface
(``f(x, order=0|1|2)``, ``.inverse``) for user code and for complex-step
checks.  The evaluation path does NOT go through them: ``Parameter.get_params``
and the model kernels apply the same formulas on the device
(``anml_b200/csrc/rowmath.cuh``), selected by ``device_code``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
from numpy.typing import NDArray


class SmoothMapping(ABC):
    device_code: int = -1  # ANML_LINK_* of include/anml_b200.h

    @staticmethod
    def _validate_order(order: int = 0) -> None:
        if order not in (0, 1, 2):
            raise ValueError("Order must be 0, 1 or 2.")

    @property
    def inverse(self) -> "SmoothMapping":
        raise NotImplementedError

    @abstractmethod
    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        ...

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class Identity(SmoothMapping):
    device_code = 0

    @property
    def inverse(self) -> SmoothMapping:
        return Identity()

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        if order == 0:
            return x  # the argument itself, as the reference does (:61-62)
        return np.ones(x.shape) if order == 1 else np.zeros(x.shape)


class Exp(SmoothMapping):
    device_code = 1

    @property
    def inverse(self) -> SmoothMapping:
        return Log()

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        return np.exp(x)


class Log(SmoothMapping):
    device_code = 2

    @property
    def inverse(self) -> SmoothMapping:
        return Exp()

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        if not np.all(x > 0):
            raise ValueError("All values for log function must be positive.")
        if order == 0:
            return np.log(x)
        return 1 / x if order == 1 else -1 / x**2


class Expit(SmoothMapping):
    """1 / (1 + exp(-x)), derivatives in the reference's exp(-x) form (:121-126)."""

    device_code = 3

    @property
    def inverse(self) -> SmoothMapping:
        return Logit()

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        z = np.exp(-x)
        if order == 0:
            return 1 / (1 + z)
        if order == 1:
            return z / (1 + z) ** 2
        return z * (z - 1) / (z + 1) ** 3


class Logit(SmoothMapping):
    device_code = 4

    @property
    def inverse(self) -> SmoothMapping:
        return Expit()

    def __call__(self, x: NDArray, order: int = 0) -> NDArray:
        self._validate_order(order)
        if not (np.all(x > 0) and np.all(x < 1)):
            raise ValueError("All values for logit function must be strictly between 0 and 1.")
        if order == 0:
            return np.log(x / (1 - x))
        if order == 1:
            return 1 / (x * (1 - x))
        return (2 * x - 1) / (x * (1 - x)) ** 2
