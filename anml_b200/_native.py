"""ctypes binding of ``libanml_b200.so`` (the C-ABI in ``include/anml_b200.h``).

The shared library is the product: there is no Python or NumPy fallback for
anything that touches the design matrix.  If the library is missing or no
sm_100 device is visible, calls raise ``RuntimeError`` -- loudly, by design.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
# ANML_B200_LIB: another build of the same library (kernel A/B experiments, tools/run_variants.sh)
LIB_PATH = os.environ.get("ANML_B200_LIB") or os.path.join(_LIB_DIR, "libanml_b200.so")

OK, EINVAL, ECUDA, ENOMEM, ESTATE, EUNSUPPORTED, EDOMAIN = range(7)

LINK_CODES = {"Identity": 0, "Exp": 1, "Log": 2, "Expit": 3, "Logit": 4}
FAMILY_CODES = {"gaussian": 0, "binomial": 1, "poisson": 2, "gaussian_meanvar": 3}
WANT_OBJ, WANT_GRAD, WANT_HESS = 1, 2, 4

_lib = None
_lock = threading.Lock()

_c_double_p = C.POINTER(C.c_double)
_c_int32_p = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/anml_b200.h declares
SIGNATURES = {
    "anml_last_error": (C.c_char_p, []),
    "anml_version": (C.c_int, []),
    "anml_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "anml_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "anml_dev_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "anml_dev_free": (C.c_int, [C.c_int, C.c_void_p]),
    "anml_dev_upload": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "anml_dev_download": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "anml_upload_release": (C.c_int, [C.c_int]),
    "anml_design_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "anml_design_create_deferred": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "anml_design_is_materialized": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "anml_design_wrap_device": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "anml_design_destroy": (C.c_int, [C.c_void_p]),
    "anml_design_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]),
    "anml_design_set_columns": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "anml_design_set_spline": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _c_int32_p]),
    "anml_design_validate_columns": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "anml_design_download": (C.c_int, [C.c_void_p, _c_double_p, C.c_int64]),
    "anml_design_params": (C.c_int, [C.c_void_p, _c_double_p, C.c_void_p, C.c_int, C.c_int, _c_double_p]),
    "anml_spline_design": (C.c_int, [C.c_int, _c_double_p, C.c_int64, _c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _c_double_p, _c_int32_p]),
    "anml_vector_minmax": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int, _c_double_p, _c_double_p, C.POINTER(C.c_int)]),
    "anml_vector_order_stats": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.c_int, _c_double_p]),
    "anml_model_create": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "anml_model_destroy": (C.c_int, [C.c_void_p]),
    "anml_model_set_param": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "anml_model_set_obs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "anml_model_set_obs_se": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "anml_model_set_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "anml_model_enable_banded_spline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _c_double_p, C.c_int, C.c_int]),
    "anml_model_set_gaussian_prior": (C.c_int, [C.c_void_p, C.c_int, _c_double_p, _c_double_p, C.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "anml_model_set_prior_enabled": (C.c_int, [C.c_void_p, C.c_int]),
    "anml_model_num_coefs": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "anml_model_eval": (C.c_int, [C.c_void_p, _c_double_p, C.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "anml_model_eval_async": (C.c_int, [C.c_void_p, _c_double_p, C.c_int, C.c_void_p, C.c_void_p]),
    "anml_model_kernel_time": (C.c_int, [C.c_void_p, C.c_int, _c_double_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "anml_model_main_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "anml_model_last_kernels": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "anml_model_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "anml_comm_handle_bytes": (C.c_int, [C.POINTER(C.c_int)]),
    "anml_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "anml_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "anml_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "anml_comm_allreduce_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "anml_comm_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "anml_comm_destroy": (C.c_int, [C.c_void_p]),
}


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(anml_b200 has no CPU fallback)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().anml_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


_EXC = {EINVAL: ValueError, ECUDA: RuntimeError, ENOMEM: MemoryError, ESTATE: ValueError,
        EUNSUPPORTED: NotImplementedError, EDOMAIN: ValueError}


def check(status: int):
    if status != OK:
        raise _EXC.get(status, RuntimeError)(last_error())


def device_count() -> int:
    n = C.c_int(0)
    rc = load().anml_device_count(C.byref(n))
    return n.value if rc == OK else 0


def require_device(device: int = 0):
    if device_count() <= device:
        raise RuntimeError(
            "anml_b200 evaluates on a CUDA device (B200, sm_100) and none is visible; there is no CPU fallback"
        )


def as_f64(a, name="array") -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if arr.dtype != np.float64:
        raise TypeError(f"{name} must be float64")
    return arr


def dptr(arr: np.ndarray):
    return arr.ctypes.data_as(_c_double_p)


class DeviceBuffer:
    """A raw device allocation owned by Python (offset / obs columns)."""

    def __init__(self, device: int, host: np.ndarray):
        host = as_f64(host)
        self.device = device
        self.nbytes = host.nbytes
        self.size = host.size
        p = C.c_void_p()
        check(load().anml_dev_alloc(device, host.nbytes, C.byref(p)))
        self.ptr = p
        if host.nbytes:
            check(load().anml_dev_upload(device, p, host.ctypes.data_as(C.c_void_p), host.nbytes))

    def download(self) -> np.ndarray:
        out = np.empty(self.size, dtype=np.float64)
        if self.nbytes:
            check(load().anml_dev_download(self.device, out.ctypes.data_as(C.c_void_p), self.ptr, self.nbytes))
        return out

    def free(self):
        if getattr(self, "ptr", None) is not None and _lib is not None:
            _lib.anml_dev_free(self.device, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Design:
    """Row-major FP64 design matrix resident in HBM (``anml_design``)."""

    def __init__(self, n_rows: int, n_cols: int, device: int = 0, _wrap=None, deferred: bool = False):
        require_device(device)
        self.device = device
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols)
        h = C.c_void_p()
        if deferred:
            # nothing allocated until something needs the dense matrix (include/anml_b200.h)
            check(load().anml_design_create_deferred(device, self.n_rows, self.n_cols, C.byref(h)))
        elif _wrap is None:
            check(load().anml_design_create(device, self.n_rows, self.n_cols, C.byref(h)))
        else:
            ptr, ld = _wrap
            check(load().anml_design_wrap_device(device, self.n_rows, self.n_cols, C.c_void_p(ptr), ld, C.byref(h)))
        self.handle = h
        self._keepalive = None

    @classmethod
    def wrap(cls, dev_ptr: int, n_rows: int, n_cols: int, ld: int, device: int = 0, keepalive=None):
        """Adopt a caller-owned device matrix (e.g. a torch CUDA tensor) without copying."""
        d = cls(n_rows, n_cols, device, _wrap=(int(dev_ptr), int(ld)))
        d._keepalive = keepalive
        d.wrapped = (int(dev_ptr), int(n_rows), int(n_cols), int(ld), device)  # identity of the borrowed matrix
        return d

    def set_columns(self, col0: int, block: np.ndarray):
        block = np.asarray(block, dtype=np.float64)
        if block.ndim == 1:
            block = block[:, None]
        if block.shape[0] != self.n_rows:
            raise ValueError(f"column block has {block.shape[0]} rows, design has {self.n_rows}")
        ncols = block.shape[1]
        if block.flags.c_contiguous:
            src, src_ld = block, ncols
        elif ncols == 1 and block.strides[0] % 8 == 0 and block.strides[0] > 0:
            src, src_ld = block, block.strides[0] // 8  # a strided view of one column
        else:
            src = np.ascontiguousarray(block)
            src_ld = ncols
        check(load().anml_design_set_columns(self.handle, col0, ncols, src.ctypes.data_as(C.c_void_p), src_ld, 0))

    def set_columns_device(self, col0: int, ncols: int, dev_ptr: int, src_ld: int):
        check(load().anml_design_set_columns(self.handle, col0, ncols, C.c_void_p(int(dev_ptr)), src_ld, 1))

    def set_spline(self, col0, x, knots, degree, ldegree=0, rdegree=0, order=0, want_index=False, x_dev_ptr=None, keepalive=None):
        knots = as_f64(knots, "knots")
        if keepalive is not None:  # a deferred design borrows the abscissae until it is materialised
            self._borrowed = getattr(self, "_borrowed", []) + [keepalive]
        idx = np.empty(self.n_rows, dtype=np.int32) if want_index else None
        idx_p = idx.ctypes.data_as(_c_int32_p) if want_index else None
        if x_dev_ptr is not None:
            xp, on_dev = _as_ptr(x_dev_ptr), 1
        else:
            x = as_f64(x, "x")
            if x.size != self.n_rows:
                raise ValueError("spline abscissae length does not match the design")
            xp, on_dev = x.ctypes.data_as(C.c_void_p), 0
        check(load().anml_design_set_spline(self.handle, col0, xp, on_dev, dptr(knots), knots.size, int(degree),
                                            int(ldegree or 0), int(rdegree or 0), int(order), idx_p))
        return idx

    @property
    def materialized(self) -> bool:
        out = C.c_int(0)
        check(load().anml_design_is_materialized(self.handle, C.byref(out)))
        return bool(out.value)

    def validate_columns(self, col0: int, ncols: int) -> np.ndarray:
        """Per-column flags from the device validators: bit 0 = has NaN, bit 1 = has a value <= 0."""
        flags = (C.c_int * ncols)()
        check(load().anml_design_validate_columns(self.handle, col0, ncols, flags))
        return np.frombuffer(flags, dtype=np.int32).copy()

    def download(self) -> np.ndarray:
        out = np.empty((self.n_rows, self.n_cols), dtype=np.float64)
        check(load().anml_design_download(self.handle, dptr(out), self.n_cols))
        return out

    def params(self, beta, offset_buf, link_code: int, order: int) -> np.ndarray:
        beta = as_f64(beta, "x")
        if beta.size != self.n_cols:
            raise ValueError(f"shapes ({self.n_rows},{self.n_cols}) and ({beta.size},) not aligned")
        if order not in (0, 1, 2):
            raise ValueError("Order must be 0, 1 or 2.")
        shape = {0: (self.n_rows,), 1: (self.n_rows, self.n_cols), 2: (self.n_rows, self.n_cols, self.n_cols)}[order]
        out = np.empty(shape, dtype=np.float64)
        off = offset_buf.ptr if offset_buf is not None else None
        check(load().anml_design_params(self.handle, dptr(beta), off, link_code, order, dptr(out)))
        return out

    def free(self):
        if getattr(self, "handle", None) is not None and _lib is not None:
            _lib.anml_design_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def spline_design(x, knots, degree, ldegree=0, rdegree=0, order=0, want_index=False, device=0):
    """XSpline.get_design_mat on the device for a host vector ``x``."""
    require_device(device)
    x = as_f64(np.asarray(x).ravel(), "x")
    knots = as_f64(knots, "knots")
    nb = knots.size - 1 + int(degree)
    out = np.empty((x.size, nb), dtype=np.float64)
    idx = np.empty(x.size, dtype=np.int32) if want_index else None
    check(load().anml_spline_design(device, dptr(x), x.size, dptr(knots), knots.size, int(degree), int(ldegree or 0),
                                    int(rdegree or 0), int(order), dptr(out),
                                    idx.ctypes.data_as(_c_int32_p) if want_index else None))
    return (out, idx) if want_index else out


def _as_ptr(p) -> C.c_void_p:
    """A device address given as an int, a ``c_void_p`` or a ``DeviceBuffer``."""
    if isinstance(p, DeviceBuffer):
        p = p.ptr
    if isinstance(p, C.c_void_p):
        return p
    return C.c_void_p(int(p))


def _vector_arg(x, dev_ptr, n):
    if dev_ptr is not None:
        return _as_ptr(dev_ptr), int(n), 1, None
    x = as_f64(np.asarray(x).ravel(), "x")
    return x.ctypes.data_as(C.c_void_p), x.size, 0, x


def vector_minmax(x=None, dev_ptr=None, n=None, device=0):
    """(min, max) of a host vector or of ``n`` doubles at ``dev_ptr``, reduced on the
    device; NaN anywhere gives (nan, nan) like ``ndarray.min()/max()``."""
    require_device(device)
    xp, n, on_dev, keep = _vector_arg(x, dev_ptr, n)
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    lo, hi, nan = C.c_double(), C.c_double(), C.c_int()
    check(load().anml_vector_minmax(device, xp, n, on_dev, C.byref(lo), C.byref(hi), C.byref(nan)))
    if nan.value:
        return float("nan"), float("nan")
    return lo.value, hi.value


def vector_order_stats(ranks, x=None, dev_ptr=None, n=None, device=0) -> np.ndarray:
    """``sorted(x)[ranks]`` from a device radix sort."""
    require_device(device)
    xp, n, on_dev, keep = _vector_arg(x, dev_ptr, n)
    ranks = np.ascontiguousarray(ranks, dtype=np.int64).ravel()
    out = np.empty(ranks.size, dtype=np.float64)
    if ranks.size:
        check(load().anml_vector_order_stats(device, xp, n, on_dev, ranks.ctypes.data_as(C.POINTER(C.c_int64)), ranks.size, dptr(out)))
    return out


def vector_quantile(q, x=None, dev_ptr=None, n=None, device=0) -> np.ndarray:
    """``np.quantile(x, q)`` (default ``method="linear"``): the order statistics come from
    the device sort, the interpolation is NumPy's own formula on the host
    (numpy/lib/_function_base_impl.py: virtual index ``(n-1)*q``, ``_get_indexes``,
    ``_lerp``), so the result carries the same bits."""
    q = np.asarray(q, dtype=np.float64)
    if q.size and not (np.all(q >= 0) and np.all(q <= 1)):
        raise ValueError("Quantiles must be in the range [0, 1]")
    if dev_ptr is None:
        x = as_f64(np.asarray(x).ravel(), "x")
        n = x.size
    n = int(n)
    if n == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    lo, hi = vector_minmax(x, dev_ptr, n, device)
    if lo != lo:  # NaNs sort to the end and poison every quantile
        return np.full(q.shape, np.nan)
    vi = (n - 1) * q.ravel()
    prev = np.floor(vi)
    nxt = prev + 1
    above = vi >= n - 1
    prev[above] = n - 1
    nxt[above] = n - 1
    ranks = np.concatenate([prev, nxt]).astype(np.int64)
    vals = vector_order_stats(ranks, x, dev_ptr, n, device)
    a, b = vals[: vi.size], vals[vi.size:]
    t = vi - np.where(above, -1.0, np.floor(vi))  # NumPy's gamma uses the wrapped index -1 there; a == b, so it is inert
    diff = np.subtract(b, a)
    out = np.add(a, diff * t)
    np.subtract(b, diff * (1 - t), out=out, where=t >= 0.5)
    return out.reshape(q.shape)


class NativeModel:
    """``anml_model``: observations + K parameters + priors -> obj / grad / Hessian."""

    def __init__(self, n_rows: int, family: str, n_params: int, device: int = 0):
        require_device(device)
        if family not in FAMILY_CODES:
            raise ValueError(f"unknown family '{family}'; choose from {sorted(FAMILY_CODES)}")
        self.device = device
        self.n_rows = int(n_rows)
        h = C.c_void_p()
        check(load().anml_model_create(device, self.n_rows, FAMILY_CODES[family], n_params, C.byref(h)))
        self.handle = h
        self._refs = []  # designs / buffers the C side borrows
        self.num_coefs = None

    def set_param(self, k: int, design: Design, link_code: int, offset_buf=None):
        check(load().anml_model_set_param(self.handle, k, design.handle, link_code, offset_buf.ptr if offset_buf is not None else None))
        self._refs.append((design, offset_buf))
        self.num_coefs = None

    def set_obs(self, obs=None, dev_ptr=None, keepalive=None):
        if dev_ptr is not None:
            check(load().anml_model_set_obs(self.handle, C.c_void_p(int(dev_ptr)), 1))
            self._refs.append(keepalive)
        else:
            obs = as_f64(obs, "obs")
            if obs.size != self.n_rows:
                raise ValueError("obs length does not match the model")
            check(load().anml_model_set_obs(self.handle, obs.ctypes.data_as(C.c_void_p), 0))

    def set_obs_se(self, se=None, dev_ptr=None, keepalive=None):
        if dev_ptr is not None:
            check(load().anml_model_set_obs_se(self.handle, C.c_void_p(int(dev_ptr)), 1))
            self._refs.append(keepalive)
        elif se is None:
            check(load().anml_model_set_obs_se(self.handle, None, 0))
        else:
            se = as_f64(se, "obs_se")
            if se.size != self.n_rows:
                raise ValueError("obs_se length does not match the model")
            check(load().anml_model_set_obs_se(self.handle, se.ctypes.data_as(C.c_void_p), 0))

    def set_weights(self, weights=None, dev_ptr=None, keepalive=None):
        if dev_ptr is not None:
            check(load().anml_model_set_weights(self.handle, C.c_void_p(int(dev_ptr)), 1))
            self._refs.append(keepalive)
        elif weights is None:
            check(load().anml_model_set_weights(self.handle, None, 0))
        else:
            weights = as_f64(weights, "weights")
            if weights.size != self.n_rows:
                raise ValueError("weights length does not match the model")
            if np.any(weights < 0):
                raise ValueError("weights must be non-negative")
            check(load().anml_model_set_weights(self.handle, weights.ctypes.data_as(C.c_void_p), 0))

    def set_gaussian_prior(self, k, mean, sd, lin_mat=None, lin_mean=None, lin_sd=None):
        mean, sd = as_f64(mean), as_f64(sd)
        if lin_mat is None or np.asarray(lin_mat).shape[0] == 0:
            check(load().anml_model_set_gaussian_prior(self.handle, k, dptr(mean), dptr(sd), 0, None, None, None))
        else:
            lin_mat, lin_mean, lin_sd = as_f64(lin_mat), as_f64(lin_mean), as_f64(lin_sd)
            check(load().anml_model_set_gaussian_prior(self.handle, k, dptr(mean), dptr(sd), lin_mat.shape[0],
                                                       dptr(lin_mat), dptr(lin_mean), dptr(lin_sd)))

    def enable_banded_spline(self, knots, degree: int, x=None, x_dev_ptr=None) -> bool:
        """Evaluate a one-spline model from the abscissae instead of the dense design
        (csrc/spline_eval.cuh).  Returns False -- and leaves the dense path in place --
        when the library declines (abscissae outside the knot span, degree > 3)."""
        knots = as_f64(knots, "knots")
        if x_dev_ptr is not None:
            xp, on_dev, keep = _as_ptr(x_dev_ptr), 1, None
        else:
            keep = as_f64(np.asarray(x).ravel(), "x")
            if keep.size != self.n_rows:
                raise ValueError("spline abscissae length does not match the model")
            xp, on_dev = keep.ctypes.data_as(C.c_void_p), 0
        status = load().anml_model_enable_banded_spline(self.handle, xp, on_dev, dptr(knots), knots.size, int(degree))
        if status == 5:  # ANML_EUNSUPPORTED
            return False
        check(status)
        return True

    def set_prior_enabled(self, enabled: bool):
        check(load().anml_model_set_prior_enabled(self.handle, 1 if enabled else 0))

    def set_option(self, name: str, value: int):
        check(load().anml_model_set_option(self.handle, name.encode(), int(value)))

    def _p(self) -> int:
        if self.num_coefs is None:
            n = C.c_int(0)
            check(load().anml_model_num_coefs(self.handle, C.byref(n)))
            self.num_coefs = n.value
        return self.num_coefs

    def eval(self, x, want=WANT_OBJ | WANT_GRAD | WANT_HESS):
        P = self._p()
        x = as_f64(x, "x")
        if x.size != P:
            raise ValueError(f"x has {x.size} entries, the model has {P} coefficients")
        obj = C.c_double(0.0)
        grad = np.empty(P, dtype=np.float64)
        hess = np.empty((P, P), dtype=np.float64) if (want & WANT_HESS) else None
        check(load().anml_model_eval(self.handle, dptr(x), want | WANT_OBJ | WANT_GRAD, C.byref(obj), dptr(grad),
                                     dptr(hess) if hess is not None else None))
        return obj.value, grad, hess

    def packed_size(self) -> int:
        P = self._p()
        return 2 + P + P * P

    def eval_async(self, x, want, packed_dev_ptr: int, stream_ptr: int = 0):
        x = as_f64(x, "x")
        if x.size != self._p():
            raise ValueError("x has the wrong length")
        check(load().anml_model_eval_async(self.handle, dptr(x), want, C.c_void_p(int(packed_dev_ptr)),
                                           C.c_void_p(int(stream_ptr)) if stream_ptr else None))

    def kernel_time(self, reset=False):
        ms, n, a = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
        check(load().anml_model_kernel_time(self.handle, 1 if reset else 0, C.byref(ms), C.byref(n), C.byref(a)))
        return ms.value, n.value, a.value

    def main_kernel(self) -> str:
        buf = C.create_string_buffer(200)
        check(load().anml_model_main_kernel(self.handle, buf, 200))
        return buf.value.decode()

    def last_kernels(self):
        buf = C.create_string_buffer(1000)
        check(load().anml_model_last_kernels(self.handle, buf, 1000))
        return [k for k in buf.value.decode().split(";") if k]

    def free(self):
        if getattr(self, "handle", None) is not None and _lib is not None:
            _lib.anml_model_destroy(self.handle)
        self.handle = None
        self._refs = []

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PeerComm:
    """``anml_comm``: all-reduce of the packed record over peer memory (one launch per evaluation)."""

    def __init__(self, device: int, rank: int, world: int, max_elems: int):
        require_device(device)
        self.device, self.rank, self.world, self.max_elems = device, rank, world, int(max_elems)
        h = C.c_void_p()
        check(load().anml_comm_create(device, rank, world, self.max_elems, C.byref(h)))
        self.handle = h

    @staticmethod
    def handle_bytes() -> int:
        n = C.c_int(0)
        check(load().anml_comm_handle_bytes(C.byref(n)))
        return n.value

    def export(self) -> bytes:
        buf = C.create_string_buffer(self.handle_bytes())
        check(load().anml_comm_export(self.handle, buf))
        return buf.raw

    def connect(self, all_handles: bytes):
        if len(all_handles) != self.world * self.handle_bytes():
            raise ValueError("need one exported handle per rank, in rank order")
        check(load().anml_comm_connect(self.handle, C.c_char_p(all_handles)))

    def allreduce_async(self, dev_ptr: int, n_elems: int, stream_ptr: int = 0):
        check(load().anml_comm_allreduce_async(self.handle, C.c_void_p(int(dev_ptr)), int(n_elems),
                                               C.c_void_p(int(stream_ptr)) if stream_ptr else None))

    def failed_epoch(self) -> int:
        v = C.c_int64(0)
        check(load().anml_comm_error(self.handle, C.byref(v)))
        return v.value

    def free(self):
        if getattr(self, "handle", None) is not None and _lib is not None:
            _lib.anml_comm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _release_upload_lanes():
    if _lib is not None:
        try:
            _lib.anml_upload_release(-1)
        except Exception:
            pass


import atexit  # noqa: E402

atexit.register(_release_upload_lanes)
