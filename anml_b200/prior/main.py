"""Prior containers (reference: src/anml/prior/main.py:10-248).

The prior layer stays in host NumPy on purpose: it is O(m p) per evaluation,
and the reference's own tests differentiate it with *complex* inputs
(tests/prior/test_main.py:74-87), which a float64 device path could not honour.
The model layer adds the same terms on the device (``add_priors_kernel``) and
``tests/`` checks the two agree.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
from numpy.typing import ArrayLike, NDArray


class Prior:
    """Distribution parameters ``params`` (k, m) and an optional linear map
    ``mat`` (m, p); ``mat is None`` means the prior acts on the variable itself.
    """

    default_params: Optional[NDArray] = None

    def __init__(self, params: List[ArrayLike], mat: Optional[ArrayLike] = None):
        self.params = params
        self.mat = mat

    # params: broadcast every entry against the others -> (len(params), m); main.py:51-56
    @property
    def params(self) -> NDArray:
        return self._params

    @params.setter
    def params(self, params: List[ArrayLike]):
        arrays = [np.asarray(p) for p in params]
        if all(a.size == 0 for a in arrays):
            self._params = np.empty((len(arrays), 0))
        else:
            self._params = np.vstack([np.ravel(b) for b in np.broadcast_arrays(*arrays)])

    # mat: single-column params are repeated to mat's row count; row counts must agree; main.py:58-66
    @property
    def mat(self) -> Optional[NDArray]:
        return self._mat

    @mat.setter
    def mat(self, mat: Optional[ArrayLike]):
        if mat is not None:
            mat = np.asarray(mat)
            rows = mat.shape[0]
            if self._params.shape[1] == 1:
                self._params = np.repeat(self._params, rows, axis=1)
            if self._params.shape[1] != rows:
                raise ValueError("Prior mat and params shape don't match.")
        self._mat = mat

    @property
    def shape(self) -> Tuple[int, int]:
        """(number of priors, size of the variable they act on)."""
        if self.mat is None:
            m = self.params.shape[1]
            return (m, m)
        return self.mat.shape

    def objective(self, x: NDArray) -> float:
        raise NotImplementedError

    def gradient(self, x: NDArray) -> NDArray:
        raise NotImplementedError

    def hessian(self, x: NDArray) -> NDArray:
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(params={self.params!r}, mat={self.mat!r})"


class GaussianPrior(Prior):
    """N(mean, sd^2) on ``x`` (or on ``mat @ x``).  ``sd`` must be positive;
    ``mean`` / ``sd`` are snapshots of ``params`` taken at construction
    (main.py:175-176)."""

    default_params: Optional[NDArray] = np.array([[0.0], [np.inf]])

    def __init__(self, mean: ArrayLike, sd: ArrayLike, mat: Optional[ArrayLike] = None):
        super().__init__([mean, sd], mat=mat)
        if not np.all(self.params[1] > 0.0):
            raise ValueError("Gaussian prior standard deviations must be positive.")
        self.mean = self.params[0]
        self.sd = self.params[1]

    def _residual(self, x):
        return self.mat.dot(x) - self.mean

    def objective(self, x: NDArray) -> float:  # main.py:178-183
        if self.mat is None:
            return 0.5 * np.sum(((x - self.mean) / self.sd) ** 2)
        if self.mat.size == 0:
            return 0.0
        return 0.5 * np.sum((self._residual(x) / self.sd) ** 2)

    def gradient(self, x: NDArray) -> NDArray:  # main.py:185-190
        if self.mat is None:
            return (x - self.mean) / self.sd**2
        if self.mat.size == 0:
            return np.zeros(x.size)
        return (self.mat.T / self.sd**2).dot(self._residual(x))

    def hessian(self, x: NDArray) -> NDArray:  # main.py:192-197
        if self.mat is None:
            return np.diag(1 / self.sd**2)
        if self.mat.size == 0:
            return np.zeros((x.size, x.size))
        return (self.mat.T / self.sd**2).dot(self.mat)


class UniformPrior(Prior):
    """Box [lb, ub] on ``x`` (or on ``mat @ x``); only read by the solver as
    bounds / linear constraints -- objective, gradient and Hessian stay the base
    class's ``NotImplementedError`` (main.py:200-248)."""

    default_params: Optional[NDArray] = np.array([[-np.inf], [np.inf]])

    def __init__(self, lb: ArrayLike, ub: ArrayLike, mat: Optional[ArrayLike] = None):
        super().__init__([lb, ub], mat=mat)
        if not np.all(self.params[0] <= self.params[1]):
            raise ValueError("Uniform prior lower bounds have to be less than or equal to the upper bounds.")
        self.lb = self.params[0]
        self.ub = self.params[1]
