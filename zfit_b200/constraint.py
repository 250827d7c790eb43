"""Constraints: O(P) host-side scalar terms added to the NLL (``src/zfit/core/constraint.py``,
call site ``core/loss.py:1155-1157``).  Not per-event work; kept minimal."""

from __future__ import annotations

import numpy as np

from .parameter import OrderedSet, ZfitParameter, convert_to_parameter


class BaseConstraint:
    def __init__(self, params):
        self._params = [convert_to_parameter(p) for p in params]

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        out = OrderedSet()
        for p in self._params:
            for leaf in p._leaves():
                if floating is None or leaf.floating == floating:
                    out.add(leaf)
        return out

    def value(self):
        raise NotImplementedError

    def _partials(self):
        raise NotImplementedError


class GaussianConstraint(BaseConstraint):
    """``zfit.constraint.GaussianConstraint(params, observation, *, uncertainty=None, sigma=None, cov=None)``:
    ``-log N(observation | params, cov)`` (``constraint.py:266-392``)."""

    def __init__(self, params, observation, *, uncertainty=None, sigma=None, cov=None):
        params = [params] if isinstance(params, ZfitParameter) else list(params)
        super().__init__(params)
        self.observation = np.atleast_1d(np.asarray(observation, dtype=np.float64))
        if cov is not None and (sigma is not None or uncertainty is not None):
            msg = "Either `sigma` or `cov` can be given, not both."
            raise ValueError(msg)
        if cov is None:
            s = uncertainty if uncertainty is not None else sigma
            if s is None:
                msg = "one of uncertainty / sigma / cov is required"
                raise ValueError(msg)
            s = np.atleast_1d(np.asarray(s, dtype=np.float64))
            if s.ndim == 1 and s.size == 1 and len(params) > 1:
                s = np.full(len(params), s[0])
            cov = np.diag(s**2) if s.ndim == 1 else s  # legacy `uncertainty`: 2-D means a covariance (constraint.py:311-320)
        else:
            cov = np.atleast_1d(np.asarray(cov, dtype=np.float64))
            if cov.ndim == 1:  # 1-D: the diagonal of the covariance matrix (constraint.py:296-299)
                cov = np.diag(np.full(len(params), cov[0]) if cov.size == 1 and len(params) > 1 else cov)
        self.cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
        if self.cov.shape != (len(params), len(params)) or self.observation.size != len(params):
            from .exception import ShapeIncompatibleError

            msg = f"params ({len(params)}), observation ({self.observation.size}) and covariance {self.cov.shape} do not match"
            raise ShapeIncompatibleError(msg)
        self._icov = np.linalg.inv(self.cov)
        self._lognorm = 0.5 * (len(params) * np.log(2 * np.pi) + np.linalg.slogdet(self.cov)[1])

    def value(self):
        d = np.array([p.value() for p in self._params]) - self.observation
        return float(0.5 * d @ self._icov @ d + self._lognorm)

    def _partials(self):
        d = np.array([p.value() for p in self._params]) - self.observation
        g = self._icov @ d
        out = {}
        for p, gi in zip(self._params, g):
            for leaf, dd in p._partials().items():
                out[leaf] = out.get(leaf, 0.0) + float(gi) * dd
        return out


class SimpleConstraint(BaseConstraint):
    """``zfit.constraint.SimpleConstraint(func, params)``: an arbitrary penalty term; gradient by central differences."""

    def __init__(self, func, params=None):
        if params is None:
            msg = "params are required"
            raise ValueError(msg)
        params = [params] if isinstance(params, ZfitParameter) else list(params)
        super().__init__(params)
        self._func = func

    def value(self):
        try:
            return float(self._func(self._params))
        except TypeError:
            return float(self._func())

    def _partials(self):
        out = {}
        for leaf in self.get_params(floating=None):
            v = leaf.value()
            h = 1e-6 * max(1.0, abs(v))
            leaf.assign(v + h)
            up = self.value()
            leaf.assign(v - h)
            dn = self.value()
            leaf.assign(v)
            out[leaf] = (up - dn) / (2 * h)
        return out


class PoissonConstraint(BaseConstraint):
    """``zfit.constraint.PoissonConstraint(params, observation)`` (``constraint.py:407-440``):
    ``-sum log Poisson(rate=observation).prob(params)`` -- the parameters are the counts, as in the reference."""

    def __init__(self, params, observation):
        params = [params] if isinstance(params, ZfitParameter) else list(params)
        super().__init__(params)
        self.observation = np.atleast_1d(np.asarray(observation, dtype=np.float64))
        if self.observation.size != len(params):
            from .exception import ShapeIncompatibleError

            msg = "params and observation have to have the same length"
            raise ShapeIncompatibleError(msg)

    def value(self):
        from scipy import special

        k = np.array([p.value() for p in self._params])
        return float(-np.sum(k * np.log(self.observation) - self.observation - special.gammaln(k + 1.0)))

    def _partials(self):
        from scipy import special

        out = {}
        for p, obs in zip(self._params, self.observation):
            g = -(np.log(obs) - special.digamma(p.value() + 1.0))
            for leaf, dd in p._partials().items():
                out[leaf] = out.get(leaf, 0.0) + float(g) * dd
        return out


class LogNormalConstraint(BaseConstraint):
    """``zfit.constraint.LogNormalConstraint(params, observation, uncertainty)`` (``constraint.py:457-523``):
    ``-sum log LogNormal(loc=observation, scale=uncertainty).prob(params)``."""

    def __init__(self, params, observation, uncertainty):
        params = [params] if isinstance(params, ZfitParameter) else list(params)
        super().__init__(params)
        self.observation = np.atleast_1d(np.asarray(observation, dtype=np.float64))
        self.uncertainty = np.broadcast_to(np.atleast_1d(np.asarray(uncertainty, dtype=np.float64)), self.observation.shape).copy()
        if self.observation.size != len(params):
            from .exception import ShapeIncompatibleError

            msg = "params, observation and uncertainty have to have the same length"
            raise ShapeIncompatibleError(msg)

    def value(self):
        x = np.array([p.value() for p in self._params])
        s = self.uncertainty
        lp = -np.log(x) - np.log(s) - 0.5 * np.log(2 * np.pi) - 0.5 * ((np.log(x) - self.observation) / s) ** 2
        return float(-np.sum(lp))

    def _partials(self):
        out = {}
        for p, loc, s in zip(self._params, self.observation, self.uncertainty):
            x = p.value()
            g = 1.0 / x + (np.log(x) - loc) / (s * s * x)
            for leaf, dd in p._partials().items():
                out[leaf] = out.get(leaf, 0.0) + float(g) * dd
        return out
