"""Array backends for the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

The PDF / NLL restatement in ``oracle/pdfs.py`` and ``oracle/nll.py`` is written
once against this tiny op table so that the *same* op graph runs

* on numpy fp64         -> value oracle (reference op order),
* on numpy longdouble   -> extended-precision adjudication of rounding noise,
* on torch fp64 + autograd -> gradient oracle (reverse-mode tape over the same
  ops, like ``tf.GradientTape`` in ``src/zfit/z/math.py:236-261``) and the
  multi-threaded CPU baseline.
"""

from __future__ import annotations

import math

import numpy as np


class NumpyBackend:
    name = "numpy"

    def __init__(self, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self._ext = self.dtype == np.dtype(np.longdouble) and np.finfo(np.longdouble).eps < 1e-18

    def asarray(self, x):
        return np.asarray(x, dtype=self.dtype)

    def const(self, x):
        return self.dtype.type(x)

    exp = staticmethod(np.exp)
    log = staticmethod(np.log)
    abs = staticmethod(np.abs)
    sign = staticmethod(np.sign)
    where = staticmethod(np.where)
    maximum = staticmethod(np.maximum)
    square = staticmethod(np.square)
    power = staticmethod(np.power)
    ones_like = staticmethod(np.ones_like)
    zeros_like = staticmethod(np.zeros_like)

    def erf(self, x):
        if self._ext:  # scipy has no longdouble erf: go through mpmath for the O(1) normalisation terms
            import mpmath

            mpmath.mp.prec = 113
            f = np.frompyfunc(lambda v: np.longdouble(str(mpmath.erf(mpmath.mpf(str(v))))), 1, 1)
            return np.asarray(f(np.asarray(x, dtype=np.longdouble)), dtype=np.longdouble)
        from scipy import special

        return special.erf(x)

    def erfc(self, x):
        if self._ext:
            import mpmath

            mpmath.mp.prec = 113
            f = np.frompyfunc(lambda v: np.longdouble(str(mpmath.erfc(mpmath.mpf(str(v))))), 1, 1)
            return np.asarray(f(np.asarray(x, dtype=np.longdouble)), dtype=np.longdouble)
        from scipy import special

        return special.erfc(x)

    def sum(self, x):
        return np.sum(x)

    def to_float(self, x):
        return float(x)


class TorchBackend:
    name = "torch"

    def __init__(self):
        import torch

        self.torch = torch
        self.dtype = torch.float64
        t = torch
        self.exp = t.exp
        self.log = t.log
        self.abs = t.abs
        self.sign = t.sign
        self.where = t.where
        self.maximum = t.maximum
        self.square = t.square
        self.ones_like = t.ones_like
        self.zeros_like = t.zeros_like
        self.erf = t.erf
        self.erfc = t.erfc

    def asarray(self, x):
        t = self.torch
        if isinstance(x, t.Tensor):
            return x.to(t.float64)
        return t.as_tensor(np.asarray(x, dtype=np.float64))

    def const(self, x):
        return self.torch.tensor(float(x), dtype=self.torch.float64)

    def power(self, a, b):
        t = self.torch
        if not isinstance(a, t.Tensor):
            a = self.const(a)
        return t.pow(a, b)

    def sum(self, x):
        return self.torch.sum(x)

    def to_float(self, x):
        return float(x)


SQRT2 = math.sqrt(2.0)
