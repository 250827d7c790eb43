"""ctypes front end of the native MIGRAD driver (``csrc/znll_migrad.cpp``, ``znll_migrad`` in ``include/znll.h``).

Same call signature and result object as ``_migrad.minimize`` (the Python statement of the algorithm, kept as the
cross-check): the C++ driver owns the iteration, this module only adapts the three Python callbacks.  An exception raised
inside a callback (``MaximumIterationReached``, a NaN strategy giving up ...) aborts the C++ loop and is re-raised here.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _znll as Z
from ._migrad import MigradResult

_DP = C.POINTER(C.c_double)
FCN_T = C.CFUNCTYPE(C.c_int, C.c_void_p, _DP, C.c_int, _DP, _DP)
INFO_T = C.CFUNCTYPE(C.c_int, C.c_void_p, _DP, _DP)
VALUE, GRADIENT = 1, 2


class Options(C.Structure):
    _fields_ = [("edm_goal", C.c_double), ("maxfcn", C.c_int32), ("hesse", C.c_int32), ("has_gradient", C.c_int32),
                ("has_value_gradient", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("fval", C.c_double), ("edm", C.c_double), ("niter", C.c_int32), ("nfcn", C.c_int32), ("ngrad", C.c_int32),
                ("converged", C.c_int32), ("valid", C.c_int32), ("message", C.c_char * 96)]


def _bound(v, inf):
    return inf if v is None else float(v)


def minimize(value, gradient, x0, steps, lower, upper, *, edm_goal, maxfcn=10000, hesse=True, verbose=False,
             value_gradient=None, information=None):
    lib = Z.lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    steps = np.ascontiguousarray(steps, dtype=np.float64)
    lo = np.array([_bound(v, -np.inf) for v in lower], dtype=np.float64)
    up = np.array([_bound(v, np.inf) for v in upper], dtype=np.float64)
    error = []
    # the driver's x / gradient buffers are copied to and from two persistent arrays with memmove: wrapping the raw
    # pointers in numpy arrays on every callback costs more than a small fit's kernel
    xbuf, gbuf = np.zeros(n), np.zeros(n)
    x_addr, g_addr, nbytes = xbuf.ctypes.data, gbuf.ctypes.data, 8 * n
    memmove = C.memmove

    def fcn(_user, xp, what, vp, gp):
        try:
            memmove(x_addr, xp, nbytes)
            x = xbuf.copy()
            if what == VALUE | GRADIENT and value_gradient is not None:
                v, g = value_gradient(x)
                vp[0] = float(v)
                gbuf[:] = g
                memmove(gp, g_addr, nbytes)
            elif what & GRADIENT:
                gbuf[:] = gradient(x)
                memmove(gp, g_addr, nbytes)
                if what & VALUE:
                    vp[0] = float(value(x))
            else:
                vp[0] = float(value(x))
            return 0
        except BaseException as exc:  # noqa: BLE001 -- carried across the C frame and re-raised below
            error.append(exc)
            return 1

    def info(_user, xp, hp):
        try:
            memmove(x_addr, xp, nbytes)
            H = information(xbuf.copy())
            if H is None:
                return 1
            H = np.ascontiguousarray(H, dtype=np.float64)
            if H.shape != (n, n):
                return 1
            memmove(hp, H.ctypes.data, 8 * n * n)
            return 0
        except Exception:  # the seed is an optimisation; any trouble falls back to Minuit's own seed
            return 1

    opts = Options(float(edm_goal), int(min(maxfcn, 2**31 - 1)), int(bool(hesse)), int(gradient is not None),
                   int(gradient is not None and value_gradient is not None))
    out = Result()
    x = np.zeros(n)
    cov = np.zeros((n, n))
    grad = np.zeros(n)
    cb = FCN_T(fcn)
    icb = INFO_T(info) if information is not None else C.cast(None, INFO_T)
    rc = lib.znll_migrad(cb, icb, None, n, x0.ctypes.data_as(_DP), steps.ctypes.data_as(_DP), lo.ctypes.data_as(_DP),
                         up.ctypes.data_as(_DP), C.byref(opts), x.ctypes.data_as(_DP), cov.ctypes.data_as(_DP),
                         grad.ctypes.data_as(_DP), C.byref(out))
    if error:
        raise error[0]
    if rc:
        msg = f"znll_migrad failed (rc={rc}): {out.message.decode(errors='replace')}"
        raise RuntimeError(msg)
    res = MigradResult()
    res.x, res.fval, res.edm = x, out.fval, out.edm
    res.covariance, res.grad = cov, grad
    res.valid, res.converged = bool(out.valid), bool(out.converged)
    res.nfcn, res.ngrad, res.niter = out.nfcn, out.ngrad, out.niter
    res.message = out.message.decode(errors="replace")
    return res


def minimize_theta(plan, fit_index, theta0, steps, lower, upper, *, edm_goal, maxfcn=10000, hesse=True, full=False,
                   log_offset=0.0, use_information=True, stream=0):
    """A whole MIGRAD minimisation inside ``libznll.so`` (``znll_migrad_theta``): the driver calls ``znll_eval_theta``
    directly, so no Python runs between the first and the last fused pass.  ``plan``: a ``_znll.Plan`` with a theta map;
    ``fit_index``: positions of the fitted parameters in its theta vector, ``theta0`` the full start vector.
    -> (MigradResult | None, theta_out): None when an evaluation was not finite (the caller then takes the evaluator
    route with its NaN strategy)."""
    lib = Z.lib()
    fn = lib.znll_migrad_theta
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), _DP, _DP, _DP, _DP, C.POINTER(Options), C.c_uint32, C.c_double,
                   C.c_int, _DP, _DP, _DP, C.POINTER(Result), C.c_void_p]
    n = len(fit_index)
    idx = (C.c_int32 * n)(*[int(i) for i in fit_index])
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    steps = np.ascontiguousarray(steps, dtype=np.float64)
    lo = np.array([_bound(v, -np.inf) for v in lower], dtype=np.float64)
    up = np.array([_bound(v, np.inf) for v in upper], dtype=np.float64)
    opts = Options(float(edm_goal), int(min(maxfcn, 2**31 - 1)), int(bool(hesse)), 1, 1)
    out = Result()
    theta_out = np.zeros_like(theta0)
    cov = np.zeros((n, n))
    grad = np.zeros(n)
    rc = fn(plan._h, n, idx, theta0.ctypes.data_as(_DP), steps.ctypes.data_as(_DP), lo.ctypes.data_as(_DP),
            up.ctypes.data_as(_DP), C.byref(opts), Z.FULL if full else 0, float(log_offset), int(bool(use_information)),
            theta_out.ctypes.data_as(_DP), cov.ctypes.data_as(_DP), grad.ctypes.data_as(_DP), C.byref(out),
            C.c_void_p(stream or 0))
    if rc == 2:
        return None, theta_out
    if rc:
        msg = f"znll_migrad_theta failed (rc={rc}): {Z.lib().znll_last_error().decode(errors='replace')} {out.message.decode(errors='replace')}"
        raise RuntimeError(msg)
    res = MigradResult()
    res.x = theta_out[np.asarray(fit_index, dtype=np.int64)]
    res.fval, res.edm = out.fval, out.edm
    res.covariance, res.grad = cov, grad
    res.valid, res.converged = bool(out.valid), bool(out.converged)
    res.nfcn, res.ngrad, res.niter = out.nfcn, out.ngrad, out.niter
    res.message = out.message.decode(errors="replace")
    return res, theta_out
