"""The thin model layer anml leaves to its callers:
``objective(x)``, ``gradient(x)``, ``hessian(x)`` for a SciPy solver, built on
``Parameter`` objects whose design matrices live on the GPU.

    objective(x) = sum_i l(obs_i, theta_1i, .., theta_Ki) + sum_k prior_k(x_k)
    theta_k      = f_k(offset_k + X_k x_k)

Where the reference route would materialise ``get_params(order=1)`` (n, p) and
``order=2`` (n, p, p) per parameter (src/anml/parameter/main.py:276-280), one
fused CUDA pass over X produces the objective, ``X^T r`` and ``X^T diag(w) X``
(``anml_b200/csrc/fused_hess.cuh``); the Gaussian prior terms are added on the
device.  ``x`` is the concatenation of the parameters' coefficient blocks in
order.  The three callbacks share one cached evaluation per ``x`` (trust-constr
asks for all three at the same point).

Families: ``gaussian`` (fixed ``obs_se``), ``binomial``, ``poisson`` -- one
parameter each -- and ``gaussian_meanvar`` (mean, variance).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
from pandas import DataFrame
from scipy.linalg import block_diag

from .data.frame import as_frame
from . import _native
from .data.example import DataExample
from .data.prototype import DataPrototype
from .parameter.main import Parameter
from .variable.main import Variable

FAMILY_PARAMS = {"gaussian": 1, "binomial": 1, "poisson": 1, "gaussian_meanvar": 2}


class Model:
    def __init__(self, family: str, data: DataPrototype, parameters: Sequence[Parameter], device: int = 0):
        if family not in FAMILY_PARAMS:
            raise ValueError(f"unknown family '{family}'; choose from {sorted(FAMILY_PARAMS)}")
        parameters = list(parameters)
        if any(not isinstance(p, Parameter) for p in parameters):
            raise TypeError("Model parameters must be instances of Parameter.")
        if len(parameters) != FAMILY_PARAMS[family]:
            raise ValueError(f"family '{family}' takes {FAMILY_PARAMS[family]} parameter(s), got {len(parameters)}")
        if not isinstance(data, DataPrototype) or not hasattr(data, "obs"):
            raise TypeError("Model data must be a DataPrototype with an 'obs' component (e.g. DataExample).")
        self.family = family
        self.data = data
        self.parameters = parameters
        self.device = device
        self._native: Optional[_native.NativeModel] = None
        self.banded_spline = False  # set by attach: the one-spline model runs on the banded evaluator
        self._cache_key = None
        self._cache = None
        self.prior_enabled = True
        self.has_gaussian_prior = False  # set by attach: some parameter carries a finite Gaussian prior
        self.num_evaluations = 0

    # ---- sizes ---------------------------------------------------------------
    @property
    def sizes(self) -> List[int]:
        return [p.size for p in self.parameters]

    @property
    def size(self) -> int:
        return sum(self.sizes)

    def split(self, x) -> List[np.ndarray]:
        x = np.asarray(x)
        edges = np.cumsum([0] + self.sizes)
        return [x[edges[k]: edges[k + 1]] for k in range(len(self.parameters))]

    # ---- attach --------------------------------------------------------------
    @staticmethod
    def _same_plain_columns(a: Parameter, b: Parameter) -> bool:
        if len(a.variables) != len(b.variables):
            return False
        for va, vb in zip(a.variables, b.variables):
            if type(va) is not Variable or type(vb) is not Variable:
                return False
            ca, cb = va.component, vb.component
            if ca.key != cb.key or ca.default_value != cb.default_value:
                return False
        return True

    def attach(self, df: DataFrame) -> None:
        """Attach observations and every parameter (design matrices go to the
        GPU), then configure the native model."""
        _native.require_device(self.device)
        df = as_frame(df)
        self.data.attach(df)
        first = self.parameters[0]
        first.attach(df)
        for p in self.parameters[1:]:
            if self._same_plain_columns(first, p):
                # same columns -> one matrix in HBM serves both parameters
                if p.offset is not None:
                    p.offset.attach(df)
                for v in p.variables:
                    v.attach(df)
                p._design, p._design_host = first._design, first._design_host
                p._offset_dev = None
                if p.offset is not None:
                    p._offset_dev = _native.DeviceBuffer(self.device, np.asarray(p.offset.value, dtype=np.float64))
                for cat in ("direct", "linear"):
                    for typ in ("UniformPrior", "GaussianPrior"):
                        getattr(p, f"_get_{cat}_prior")(typ)
            else:
                p.attach(df)
        obs = np.asarray(self.data.obs.value, dtype=np.float64)
        se = None
        if self.family == "gaussian" and hasattr(self.data, "obs_se"):
            se = np.asarray(self.data.obs_se.value, dtype=np.float64)
            if np.all(se == 1.0):
                se = None  # the default column (data/example.py:21): skip reading it
        weights = None
        if hasattr(self.data, "weights"):  # optional component: per-row likelihood weights (trimming)
            weights = np.asarray(self.data.weights.value, dtype=np.float64)
        self._configure(len(df), obs=obs, se=se, weights=weights)

    def attach_device(self, n_rows: int, obs_dev_ptr: int, se_dev_ptr: Optional[int] = None, keepalive=None) -> None:
        """Configure from parameters that were given device-resident designs
        (``Parameter.attach_device``) and a device-resident ``obs`` vector."""
        _native.require_device(self.device)
        self._configure(n_rows, obs_ptr=obs_dev_ptr, se_ptr=se_dev_ptr, keepalive=keepalive)

    def _configure(self, n, obs=None, se=None, obs_ptr=None, se_ptr=None, keepalive=None, weights=None):
        nm = _native.NativeModel(n, self.family, len(self.parameters), self.device)
        for k, p in enumerate(self.parameters):
            if p.device_design is None:
                raise ValueError("Must provide a data frame to attach data.")
            code = p.transform.device_code
            if code < 0:
                raise NotImplementedError(f"{p.transform!r} has no device implementation")
            design = p.device_design
            # two parameters that wrap the SAME device matrix (Parameter.attach_device) are one design to the
            # library: the mean / variance model then takes the shared-design path (X read three times)
            first = self.parameters[0].device_design
            if k > 0 and getattr(design, "wrapped", None) is not None and getattr(first, "wrapped", None) == design.wrapped:
                design = first
            nm.set_param(k, design, code, p.device_offset)
        if obs_ptr is not None:
            nm.set_obs(dev_ptr=obs_ptr, keepalive=keepalive)
            if se_ptr is not None:
                nm.set_obs_se(dev_ptr=se_ptr, keepalive=keepalive)
        else:
            nm.set_obs(obs)
            if se is not None:
                nm.set_obs_se(se)
        if weights is not None:
            nm.set_weights(weights)
        self.has_gaussian_prior = False
        for k, p in enumerate(self.parameters):
            direct = p.prior_dict["direct"]["GaussianPrior"]
            linear = p.prior_dict["linear"]["GaussianPrior"]
            has_direct = bool(np.any(np.isfinite(direct.sd)))
            has_linear = linear.mat is not None and linear.mat.shape[0] > 0
            if has_direct or has_linear:
                self.has_gaussian_prior = True
                nm.set_gaussian_prior(k, direct.mean, direct.sd,
                                      linear.mat if has_linear else None,
                                      linear.mean if has_linear else None,
                                      linear.sd if has_linear else None)
        nm.set_prior_enabled(self.prior_enabled)
        self.banded_spline = self._enable_banded_spline(nm)
        self._native = nm
        self._cache_key = None

    # A parameter that is exactly one spline variable (BASELINE config #3) is evaluated from the
    # abscissae, the basis computed on the fly, instead of streaming the dense design (spline_eval.cuh).
    use_banded_spline = True

    def _enable_banded_spline(self, nm) -> bool:
        if not self.use_banded_spline or len(self.parameters) != 1 or len(self.parameters[0].variables) != 1:
            return False
        var = self.parameters[0].variables[0]
        x_dev = getattr(var, "device_abscissae", None)
        if x_dev is None or not (1 <= var.spline.degree <= 3) or x_dev.size != nm.n_rows:
            return False
        return nm.enable_banded_spline(var.spline.knots, var.spline.degree, x_dev_ptr=x_dev.ptr)

    def set_prior_enabled(self, enabled: bool) -> None:
        """Row-sharded runs keep the prior on one rank only."""
        self.prior_enabled = bool(enabled)
        if self._native is not None:
            self._native.set_prior_enabled(self.prior_enabled)
        self._cache_key = None

    @property
    def native(self) -> _native.NativeModel:
        if self._native is None:
            raise ValueError("Must provide a data frame to attach data.")
        return self._native

    # ---- evaluation ------------------------------------------------------------
    def evaluate(self, x, hessian: bool = True):
        """(objective, gradient, Hessian-or-None) at ``x``; one device pass."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        key = (x.tobytes(), bool(hessian))
        if self._cache_key is not None and self._cache_key[0] == key[0] and (self._cache_key[1] or not hessian):
            return self._cache
        want = _native.WANT_OBJ | _native.WANT_GRAD | (_native.WANT_HESS if hessian else 0)
        self._cache = self.native.eval(x, want)
        self._cache_key = key
        self.num_evaluations += 1
        return self._cache

    def objective(self, x) -> float:
        return self.evaluate(x, hessian=self._cache_wants_hessian(x))[0]

    def gradient(self, x) -> np.ndarray:
        return self.evaluate(x, hessian=self._cache_wants_hessian(x))[1].copy()

    def hessian(self, x) -> np.ndarray:
        return self.evaluate(x, hessian=True)[2].copy()

    def _cache_wants_hessian(self, x) -> bool:
        # objective/gradient calls ride on a cached Hessian evaluation when there is one;
        # otherwise they take the cheaper (HBM-bound) gradient pass unless `eager_hessian`
        return getattr(self, "eager_hessian", True)

    # ---- solver wiring -----------------------------------
    def bounds(self):
        """scipy ``Bounds`` from the direct UniformPriors (parameter/main.py:186-188)."""
        from scipy.optimize import Bounds

        lb = np.hstack([p.prior_dict["direct"]["UniformPrior"].lb for p in self.parameters])
        ub = np.hstack([p.prior_dict["direct"]["UniformPrior"].ub for p in self.parameters])
        if np.all(np.isinf(lb)) and np.all(np.isinf(ub)):
            return None
        return Bounds(lb, ub)

    def constraints(self):
        """scipy ``LinearConstraint`` from the linear UniformPriors (:225-227)."""
        from scipy.optimize import LinearConstraint

        mats, lbs, ubs = [], [], []
        for p in self.parameters:
            pr = p.prior_dict["linear"]["UniformPrior"]
            mats.append(pr.mat)
            lbs.append(pr.lb)
            ubs.append(pr.ub)
        mat = block_diag(*mats)
        if mat.shape[0] == 0:
            return []
        return [LinearConstraint(mat, np.hstack(lbs), np.hstack(ubs))]

    def fit(self, x0=None, options=None, method: str = "trust-constr"):
        """Minimise the objective with SciPy (trust-constr by default) using the
        device callbacks; returns the ``OptimizeResult``."""
        from scipy.optimize import minimize

        x0 = np.zeros(self.size) if x0 is None else np.asarray(x0, dtype=np.float64)
        opts = {"gtol": 1e-10, "xtol": 1e-14, "maxiter": 200}
        opts.update(options or {})
        kwargs = {}
        b = self.bounds()
        if b is not None:
            kwargs["bounds"] = b
        c = self.constraints()
        if c:
            kwargs["constraints"] = c
        return minimize(self.objective, x0, jac=self.gradient, hess=self.hessian, method=method, options=opts, **kwargs)


def gaussian_model(obs: str, parameter: Parameter, obs_se: str = "obs_se", device: int = 0) -> Model:
    return Model("gaussian", DataExample(obs, obs_se), [parameter], device)


def binomial_model(obs: str, parameter: Parameter, device: int = 0) -> Model:
    return Model("binomial", DataExample(obs, "__unit_se__"), [parameter], device)
