"""Parameter = f(offset + X beta) with X resident in HBM.

Reference: src/anml/parameter/main.py:84-340.  Same constructor, properties,
``attach`` / ``get_params`` / ``prior_*`` surface and error behaviour; what
changes is where the design matrix lives and who evaluates it:

* ``attach(df)`` does not ``np.hstack`` on the host (:160-162).  Every variable
  writes its column(s) into a row-major FP64 matrix on the GPU
  (``anml_design``); spline variables build their basis there with the CUDA
  spline kernel.  ``design_mat`` downloads a host copy on demand.
* ``get_params(x, order)`` (:265-280) runs on the device and returns host
  arrays of the reference's shapes -- (n,), (n, p), (n, p, p).  It exists for
  API parity at small n; models never materialise the order-1/2 objects, they
  use the fused evaluator of :mod:`anml_b200.model`.
* ``prior_objective / gradient / hessian`` stay host NumPy (dtype-generic; the
  reference tests them with complex inputs).
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
from numpy.typing import NDArray
from pandas import DataFrame
from scipy.linalg import block_diag

from ..data.frame import as_frame
from .. import _native
from ..data.component import Component
from ..data.validator import NoNans
from ..prior.main import Prior
from ..prior.utils import filter_priors, get_prior_type
from ..variable.main import Variable
from .smoothmapping import Identity, SmoothMapping

_CATEGORIES = ("direct", "linear")
_TYPES = ("UniformPrior", "GaussianPrior")


class Parameter:
    def __init__(self, variables: List[Variable], transform: Optional[SmoothMapping] = None,
                 offset: Optional[Union[str, Component]] = None, priors: Optional[List[Prior]] = None, device: int = 0):
        self.variables = variables
        self.transform = transform
        self.offset = offset
        self.priors = priors
        self.device = device

        self._design = None        # _native.Design (device)
        self._design_host = None   # lazily downloaded copy, or a user-assigned array
        self._offset_dev = None    # _native.DeviceBuffer
        self.prior_dict = {"direct": {}, "linear": {}}

    # ---- validated attributes (main.py:99-138) --------------------------------
    @property
    def variables(self) -> List[Variable]:
        return self._variables

    @variables.setter
    def variables(self, variables: List[Variable]):
        variables = list(variables)
        if any(not isinstance(v, Variable) for v in variables):
            raise TypeError("Parameter input variables must be a list of instances of Variable.")
        self._variables = variables

    @property
    def transform(self) -> SmoothMapping:
        return self._transform

    @transform.setter
    def transform(self, transform: Optional[SmoothMapping]):
        if transform is None:
            transform = Identity()
        elif not isinstance(transform, SmoothMapping):
            raise TypeError("Parameter input transform must be an instance of SmoothMapping or None.")
        self._transform = transform

    @property
    def offset(self) -> Optional[Component]:
        return self._offset

    @offset.setter
    def offset(self, offset: Optional[Union[str, Component]]):
        if isinstance(offset, str):
            offset = Component(offset, validators=[NoNans()])
        elif offset is not None and not isinstance(offset, Component):
            raise TypeError("Parameter input offset has to be a string or an instance of Component.")
        self._offset = offset

    @property
    def priors(self) -> List[Prior]:
        return self._priors

    @priors.setter
    def priors(self, priors: Optional[List[Prior]]):
        priors = [] if priors is None else list(priors)
        if any(not isinstance(p, Prior) for p in priors):
            raise TypeError("Parameter input priors must be a list of instances of Prior.")
        self._priors = priors

    @property
    def size(self) -> int:
        return sum(v.size for v in self.variables)

    # ---- design matrix ----------------------------------------------------------
    @property
    def design_mat(self) -> Optional[NDArray]:
        """Host copy (n, p) of the cached design matrix; ``None`` before attach."""
        if self._design_host is None and self._design is not None:
            self._design_host = self._design.download()
        return self._design_host

    @design_mat.setter
    def design_mat(self, value):
        # the reference exposes a plain attribute; assigning an array re-uploads it
        if value is None:
            self._design, self._design_host = None, None
            return
        value = np.ascontiguousarray(value, dtype=np.float64)
        design = _native.Design(value.shape[0], value.shape[1], self.device)
        design.set_columns(0, value)
        self._design, self._design_host = design, value

    @property
    def device_design(self):
        """The ``anml_design`` handle wrapper (for the model layer)."""
        return self._design

    @property
    def device_offset(self):
        return self._offset_dev

    def attach(self, df: DataFrame) -> None:
        """Cache offset, design matrix (on the device) and the four combined
        priors; rebuilt on every call like the reference (main.py:148-165)."""
        _native.require_device(self.device)
        df = as_frame(df)  # pandas frame, or a pyarrow.Table behind the same few methods
        n = len(df)
        self._offset_dev = None
        if self.offset is not None:
            self.offset.attach(df)
            self._offset_dev = _native.DeviceBuffer(self.device, np.asarray(self.offset.value, dtype=np.float64))
        # sizes can depend on the first attach of a spline variable, so resolve those first;
        # plain variables attach inside _write_design (validators run on the device), spline
        # variables too (their size is known from the getter; knots are resolved from the device copy)
        for v in self.variables:
            if type(v) is not Variable and not getattr(v, "_resolves_on_device", False):
                v.attach(df)
        # A parameter that is exactly one spline variable gets a *deferred* design: the basis block is only
        # recorded, and the dense (n, nb) matrix is built if and when something reads it (design_mat,
        # get_params, a dense evaluation).  A model on the banded evaluator never does (BASELINE config #3).
        lone_spline = len(self.variables) == 1 and getattr(self.variables[0], "_resolves_on_device", False)
        design = _native.Design(n, self.size, self.device, deferred=lone_spline)
        col = 0
        checks = []
        for v in self.variables:
            v._write_design(design, col, df)
            want_nan, want_pos = v._device_checks() if hasattr(v, "_device_checks") else (False, False)
            if want_nan or want_pos:
                checks.append((col, v.size, v.component.key, want_nan, want_pos))
            col += v.size
        if checks:
            # NoNans / Positive (data/validator.py:21-34) in one pass over the uploaded matrix
            flags = design.validate_columns(0, self.size)
            for col0, size, key, want_nan, want_pos in checks:
                f = int(np.bitwise_or.reduce(flags[col0: col0 + size]))
                if want_nan and (f & 1):
                    raise ValueError(f"Column '{key}' contains nans.")
                if want_pos and (f & 2):
                    raise ValueError(f"Column '{key}' contains nonpositive numbers.")
        self._design, self._design_host = design, None
        for category in _CATEGORIES:
            for prior_type in _TYPES:
                getattr(self, f"_get_{category}_prior")(prior_type)

    def attach_device(self, dev_ptr: int, n_rows: int, ld: Optional[int] = None, keepalive=None,
                      offset_dev_ptr: Optional[int] = None) -> None:
        """Adopt a design matrix that already lives in device memory (row-major
        FP64, ``n_rows`` x ``size``, leading dimension ``ld``), e.g. a torch CUDA
        tensor; priors are assembled from the variables as in ``attach``."""
        _native.require_device(self.device)
        ld = self.size if ld is None else int(ld)
        self._design = _native.Design.wrap(dev_ptr, n_rows, self.size, ld, self.device, keepalive)
        self._design_host = None
        self._offset_dev = None
        if offset_dev_ptr is not None:
            self._offset_dev = _BorrowedBuffer(offset_dev_ptr, keepalive)
        for category in _CATEGORIES:
            for prior_type in _TYPES:
                getattr(self, f"_get_{category}_prior")(prior_type)

    # ---- combined priors (main.py:167-227) -------------------------------------
    def _get_direct_prior(self, prior_type: str) -> None:
        """Direct priors come from the variables only; mat-less priors passed to
        the Parameter itself are ignored, as in the reference (:170-172)."""
        params = np.hstack([v.get_direct_prior_params(prior_type) for v in self.variables])
        self.prior_dict["direct"][prior_type] = get_prior_type(prior_type)(params[0], params[1])

    def _get_linear_prior(self, prior_type: str) -> None:
        per_var = [v.get_linear_prior_params(prior_type) for v in self.variables]
        params = np.hstack([p for p, _ in per_var])
        mat = block_diag(*[m for _, m in per_var])
        own = filter_priors(self.priors, prior_type, with_mat=True)
        if own:
            params = np.hstack([params] + [p.params for p in own])
            mat = np.vstack([mat] + [p.mat for p in own])
        else:
            mat = np.vstack([mat, np.empty((0, self.size))])
        self.prior_dict["linear"][prior_type] = get_prior_type(prior_type)(params[0], params[1], mat)

    # ---- evaluation ---------------------------------------------------------------
    def get_params(self, x: NDArray, df: Optional[DataFrame] = None, order: int = 0) -> NDArray:
        """theta = f(offset + X x) (order 0), its Jacobian (order 1) or the
        second-order tensor (order 2), computed on the device."""
        if df is not None:
            self.attach(df)
        if self._design is None:
            raise ValueError("Must provide a data frame to attach data.")
        if order not in (0, 1, 2):
            raise ValueError("Order must be 0, 1 or 2.")
        code = self.transform.device_code
        if code < 0:
            raise NotImplementedError(f"{self.transform!r} has no device implementation")
        return self._design.params(np.asarray(x, dtype=np.float64), self._offset_dev, code, order)

    def _gaussian_terms(self, method: str, x: NDArray):
        return [getattr(self.prior_dict[c]["GaussianPrior"], method)(x) for c in _CATEGORIES]

    def prior_objective(self, x: NDArray) -> float:
        value = 0.0
        for term in self._gaussian_terms("objective", x):
            value += term
        return value

    def prior_gradient(self, x: NDArray) -> NDArray:
        value = np.zeros(x.size, dtype=x.dtype)
        for term in self._gaussian_terms("gradient", x):
            value += term
        return value

    def prior_hessian(self, x: NDArray) -> NDArray:
        value = np.zeros((x.size, x.size), dtype=x.dtype)
        for term in self._gaussian_terms("hessian", x):
            value += term
        return value

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(variables={self.variables}, transform={self.transform}, "
                f"offset={self.offset}, priors={self.priors})")


class _BorrowedBuffer:
    """A device vector owned by someone else (duck-types ``DeviceBuffer.ptr``)."""

    def __init__(self, ptr: int, keepalive=None):
        import ctypes

        self.ptr = ctypes.c_void_p(int(ptr))
        self._keepalive = keepalive
