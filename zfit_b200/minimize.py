"""Minimizer layer: the *callers* of the hot path (``src/zfit/minimizers/``).

``LossEval`` (``evaluation.py:51-406``) is the only thing a minimizer calls on a loss; it is restated
here so that every driver -- ``Minuit`` (iminuit when importable, else the native stand-in in
``_migrad.py``) and the scipy wrappers -- runs unchanged against the CUDA loss.
"""

from __future__ import annotations

import contextlib
import copy
import math
import time
from collections.abc import Mapping

import numpy as np

from . import _migrad, _migrad_native
from .exception import DerivativeCalculationError, FailMinimizeNaN, MaximumIterationReached
from .parameter import ZfitParameter, assign_values, set_values


# ------------------------------------------------------------------------------------------------
# strategies (strategy.py:42-144)
# ------------------------------------------------------------------------------------------------
class BaseStrategy:
    def __init__(self):
        self.fit_result = None
        self.error = None

    def minimize_nan(self, loss, params, values=None):
        raise FailMinimizeNaN

    def callback(self, value, gradient, hessian, params, loss):
        return value, gradient, hessian

    def __str__(self):
        return type(self).__name__


class PushbackStrategy(BaseStrategy):
    def __init__(self, nan_penalty=100, nan_tol=30):
        super().__init__()
        self.nan_penalty = nan_penalty
        self.nan_tol = nan_tol

    def minimize_nan(self, loss, params, values=None):
        nan_counter = values["nan_counter"]
        if nan_counter < self.nan_tol:
            last_loss = values.get("old_loss")
            last_grad = values.get("old_grad")
            if last_grad is not None:
                last_grad = -last_grad
            if last_loss is not None:
                loss_evaluated = last_loss + self.nan_penalty * nan_counter
            else:
                loss_evaluated = values.get("loss")
            if isinstance(loss_evaluated, str):
                msg = "Loss starts already with NaN, cannot minimize."
                raise RuntimeError(msg)
            return loss_evaluated, last_grad
        return super().minimize_nan(loss=loss, params=params, values=values)


class ToyStrategyFail(BaseStrategy):
    def minimize_nan(self, loss, params, values=None):
        self.fit_result = FitResult(
            loss=loss, params={p: p.value() for p in params}, minimizer=None, valid=False, converged=False,
            edm=None, fminopt=None, message="Failed on too many NaNs", info={},
        )
        raise FailMinimizeNaN


class DefaultToyStrategy(PushbackStrategy, ToyStrategyFail):
    pass


DefaultStrategy = PushbackStrategy


# ------------------------------------------------------------------------------------------------
# LossEval (evaluation.py:51-406)
# ------------------------------------------------------------------------------------------------
def check_derivative_none_raise(values, params):
    if isinstance(values, np.ndarray) and values.dtype != object:  # a float array holds no None
        return
    if values is None or any(v is None for v in values):
        none = [p for p, v in zip(params, values or [None] * len(params)) if v is None]
        msg = f"The derivative of the following parameters is None: {none}."
        raise DerivativeCalculationError(msg)


class LossEval:
    def __init__(self, loss, params, strategy, do_print, maxiter, grad_fn=None, hesse_fn=None, numpy_converter=None,
                 full=None):
        self.full = False if full is None else full
        self.maxiter = maxiter
        self._ignoring_maxiter = False
        self._nfunc_eval = self._ngrad_eval = self._nhess_eval = 0
        self.npass = 0  # calls that reached the loss (one fused pass over the events each, hessian calls excepted)
        self.loss = loss
        self.hesse_fn = loss.hessian if hesse_fn is None else hesse_fn
        if grad_fn is not None:
            def value_gradients_fn(params):
                return loss.value(full=self.full), grad_fn(params)
        else:
            def value_gradients_fn(params):
                return loss.value_gradient(params, full=self.full)
            grad_fn = loss.gradient
        self.gradients_fn = grad_fn
        self.value_gradients_fn = value_gradients_fn
        self.last_value = None
        self.last_gradient = None
        self.last_hessian = None
        self.nan_counter = 0
        self.params = list(params)
        self.strategy = strategy
        self.do_print = do_print
        self.numpy_converter = False if numpy_converter is None else numpy_converter

    @property
    def niter(self):
        return max(self._nfunc_eval, self._ngrad_eval, self._nhess_eval)

    @property
    def ignoring_maxiter(self):
        return self._ignoring_maxiter or self.maxiter is None

    @property
    def maxiter_reached(self):
        return not self.ignoring_maxiter and self.niter > self.maxiter

    def _check_maxiter_reached(self):
        if self.maxiter_reached:
            raise MaximumIterationReached

    @contextlib.contextmanager
    def ignore_maxiter(self):
        old = self._ignoring_maxiter
        self._ignoring_maxiter = True
        yield
        self._ignoring_maxiter = old

    nfunc_eval = property(lambda self: self._nfunc_eval)
    ngrad_eval = property(lambda self: self._ngrad_eval)
    nhess_eval = property(lambda self: self._nhess_eval)

    def _count(self, f=0, g=0, h=0):
        self.npass += 1
        if not self._ignoring_maxiter:
            self._nfunc_eval += f
            self._ngrad_eval += g
            self._nhess_eval += h
            self._check_maxiter_reached()

    def _nan(self, loss_value, gradient, want):
        self.nan_counter += 1
        info = {
            "loss": loss_value,
            "grad": gradient,
            "old_loss": self.last_value,
            "old_grad": self.last_gradient,
            "nan_counter": self.nan_counter,
        }
        v, g = self.strategy.minimize_nan(loss=self.loss, params=self.params, values=info)
        return v, g

    def value_gradient(self, values):
        self._count(f=1, g=1)
        params = self.params
        assign_values(params, values)
        loss_value, gradient = self.value_gradients_fn(params)
        loss_value, gradient, _ = self.strategy.callback(loss_value, gradient, None, params, self.loss)
        if self.do_print:
            print_gradient(params, values, gradient, loss_value)
        check_derivative_none_raise(gradient, params)
        if loss_value != loss_value or np.isnan(gradient).any():
            loss_value, gradient = self._nan(loss_value, gradient, "vg")
        else:
            self.nan_counter = 0
            self.last_value = loss_value
            self.last_gradient = gradient
        if self.numpy_converter:
            loss_value = float(loss_value)
            gradient = self.numpy_converter(gradient)
        return loss_value, gradient

    def value(self, values):
        self._count(f=1)
        params = self.params
        assign_values(params, values)
        loss_value = self.loss.value(full=self.full)
        loss_value, _, _ = self.strategy.callback(loss_value, None, None, params, self.loss)
        if self.do_print:
            print_params(params, values, loss_value)
        if np.isnan(loss_value):
            self.nan_counter += 1
            info = {"loss": loss_value, "old_loss": self.last_value, "nan_counter": self.nan_counter}
            loss_value, _ = self.strategy.minimize_nan(loss=self.loss, params=params, values=info)
        else:
            self.nan_counter = 0
            self.last_value = loss_value
        if self.numpy_converter:
            loss_value = float(loss_value)
        return loss_value

    def gradient(self, values):
        self._count(g=1)
        params = self.params
        assign_values(params, values)
        gradient = self.gradients_fn(params)
        _, gradient, _ = self.strategy.callback(None, gradient, None, params, self.loss)
        if self.do_print:
            print_gradient(params, values, gradient, None)
        check_derivative_none_raise(gradient, params)
        if np.any(np.isnan(gradient)):
            self.nan_counter += 1
            info = {"loss": self.last_value, "grad": gradient, "old_loss": self.last_value,
                    "old_grad": self.last_gradient, "nan_counter": self.nan_counter}
            _, gradient = self.strategy.minimize_nan(loss=self.loss, params=params, values=info)
        else:
            self.nan_counter = 0
            self.last_gradient = gradient
        if self.numpy_converter:
            gradient = self.numpy_converter(gradient)
        return gradient

    def hessian(self, values):
        self._count(h=1)
        params = self.params
        assign_values(params, values)
        hessian = self.hesse_fn(params)
        _, _, hessian = self.strategy.callback(None, None, hessian, params, self.loss)
        if np.any(np.isnan(hessian)):
            self.nan_counter += 1
        else:
            self.nan_counter = 0
            self.last_hessian = hessian
        if self.numpy_converter:
            hessian = self.numpy_converter(hessian)
        return hessian


def print_params(params, values, loss=None):
    print("  ".join(f"{p.name}={v:.6g}" for p, v in zip(params, values)) + (f"  loss={loss}" if loss is not None else ""))


def print_gradient(params, values, gradient, loss=None):
    print_params(params, values, loss)
    print("  grad: " + "  ".join(f"{g:.6g}" if not isinstance(g, str) else g for g in gradient))


# ------------------------------------------------------------------------------------------------
# convergence criterion (termination.py)
# ------------------------------------------------------------------------------------------------
class ConvergenceCriterion:
    def __init__(self, tol, loss, params, name):
        self.loss = loss
        self.tol = tol * loss.errordef
        self.params = params
        self.name = name
        self.last_value = None

    def converged(self, result):
        return self.calculate(result) < self.tol

    def calculate(self, result):
        value = self._calculate(result)
        self.last_value = value
        return value


class EDM(ConvergenceCriterion):
    """``EDM = g^T H^-1 g / 2`` (``termination.py:67-128``)."""

    def __init__(self, tol, loss, params, name="edm"):
        super().__init__(tol, loss, params, name)

    def _calculate(self, result):
        params = list(result.params)
        grad = result.approx.get("gradient")
        if grad is None:
            grad = result.loss.gradient(params)
        grad = np.asarray(grad, dtype=np.float64)
        inv_h = result.approx.get("inv_hessian")
        try:
            if inv_h is None:
                h = result.approx.get("hessian")
                if h is None:
                    h = result.loss.hessian(params)
                edm = float(grad @ np.linalg.solve(h, grad) / 2)
            else:
                edm = float(grad @ inv_h @ grad / 2)
        except np.linalg.LinAlgError:
            edm = 999_999
        return 999_999 if edm < 0 else edm


# ------------------------------------------------------------------------------------------------
# FitResult (fitresult.py:425-700, 1339-1671, 1782-1796)
# ------------------------------------------------------------------------------------------------
class _ParamsView(dict):
    """``result.params``: ``{param: {"value": v, <error name>: {...}}}``, also addressable by name."""

    def __getitem__(self, key):
        if isinstance(key, str):
            for p in self:
                if p.name == key:
                    return dict.__getitem__(self, p)
            raise KeyError(key)
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        if isinstance(key, str):
            return any(p.name == key for p in self)
        return dict.__contains__(self, key)


class FitResult:
    def __init__(self, loss, params, minimizer, valid, converged, edm, fminopt, message="", info=None, approx=None,
                 criterion=None, status=None, niter=None, evaluator=None, fminfull=None):
        self.loss = loss
        self.minimizer = minimizer
        self.params = _ParamsView({p: {"value": float(v)} for p, v in params.items()})
        self._valid = bool(valid)
        self._converged = bool(converged)
        self.edm = edm
        self.fminopt = fminopt
        self._fmin = fminfull
        self.message = message
        self.info = {} if info is None else dict(info)
        self.approx = {} if approx is None else dict(approx)
        self.criterion = criterion
        self.status = status
        self.niter = niter
        self.evaluator = evaluator
        self._covariance = {}

    @property
    def valid(self):
        return self._valid and not self.params_at_limit and self._converged

    @property
    def converged(self):
        return self._converged

    @property
    def params_at_limit(self):
        return any(getattr(p, "at_limit", False) for p in self.params)

    @property
    def values(self):
        return _Values(self.params)

    @property
    def fmin(self):
        """Full loss value at the minimum (``fitresult.py``: ``loss.value(full=True)``)."""
        if self._fmin is None and self.loss is not None:
            with set_values(list(self.params), [v["value"] for v in self.params.values()]):
                self._fmin = float(self.loss.value(full=True))
        return self._fmin

    fminfull = fmin

    def update_params(self):
        """Set the parameters to the fitted values (``fitresult.py:1782-1796``)."""
        set_values(list(self.params), [v["value"] for v in self.params.values()])
        return self

    def __enter__(self):
        self._restore = set_values(list(self.params), [v["value"] for v in self.params.values()])
        return self

    def __exit__(self, *exc):
        self._restore.__exit__(*exc)
        return False

    # -- uncertainties -----------------------------------------------------------------------------
    def covariance(self, params=None, method=None, as_dict=False, weightcorr=None):
        """Covariance = inverse Hessian of the NLL at the minimum (``hesse_np``, ``fitresult.py:234-244, 1600-1671``)
        -- 2P fused-kernel gradient calls; ``method="approx"`` returns the minimizer's own estimate.  For weighted
        data the estimate is corrected (``weightcorr``: ``"asymptotic"`` (default) -> ``H^-1 C H^-1`` with
        ``C = sum w_i^2 grad grad^T`` from one extra fused pass, ``"sumw2"`` -> ``H^-1 * sum w^2 / sum w``, ``False`` ->
        none; ``minimizers/errors.py:333-464``)."""
        allp = list(self.params)
        params = allp if params is None else ([params] if isinstance(params, ZfitParameter) else list(params))
        method = method or "hesse_np"
        weighted = bool(getattr(self.loss, "is_weighted", False)) and method != "approx"
        if weightcorr is None:
            weightcorr = "asymptotic" if weighted else False
        if weightcorr not in (False, "asymptotic", "sumw2"):
            msg = f"Unknown method {weightcorr}, has to be one of False, 'asymptotic', 'sumw2'"
            raise ValueError(msg)
        key = (method, weightcorr if weighted else False)
        if key not in self._covariance:
            if method == "approx":
                cov = self.approx.get("inv_hessian")
                if cov is None:
                    method = "hesse_np"
            if method != "approx":
                with set_values(allp, [self.params[p]["value"] for p in allp]):
                    hess = np.asarray(self.loss.hessian(allp), dtype=np.float64)
                    cov = np.linalg.inv(hess) * (2.0 * self.loss.errordef)
                    if weighted and weightcorr == "asymptotic":
                        C = self.loss.weighted_score_outer(allp)
                        cov = cov @ C @ cov
                    elif weighted and weightcorr == "sumw2":
                        sw = sum(float(np.sum(d.weights)) if d.has_weights else float(d.num_entries) for d in self.loss.data)
                        sw2 = sum(float(np.sum(d.weights**2)) if d.has_weights else float(d.num_entries) for d in self.loss.data)
                        ncon = len(self.loss.constraints)
                        cov = cov * ((sw2 + ncon) / (sw + ncon))
            self._covariance[key] = np.asarray(cov)
        cov = self._covariance[key]
        idx = [allp.index(p) for p in params]
        sub = cov[np.ix_(idx, idx)]
        if as_dict:
            return {(p1, p2): sub[i, j] for i, p1 in enumerate(params) for j, p2 in enumerate(params)}
        return sub

    def correlation(self, params=None, method=None, as_dict=False):
        cov = self.covariance(params, method)
        d = np.sqrt(np.diag(cov))
        return cov / np.outer(d, d)

    def hesse(self, params=None, method=None, cl=None, name=None, weightcorr=None):
        """Symmetric (parabolic) uncertainties from the covariance (``fitresult.py:1339-1462``)."""
        allp = list(self.params)
        params = allp if params is None else ([params] if isinstance(params, ZfitParameter) else list(params))
        method = method or "hesse_np"
        name = name or ("hesse" if method != "approx" else "approx")
        scale = 1.0
        if cl is not None:
            from scipy import stats

            scale = stats.norm.ppf(0.5 + cl / 2)
        cov = self.covariance(params, method, weightcorr=weightcorr)
        out = {}
        for i, p in enumerate(params):
            err = {"error": float(math.sqrt(cov[i, i])) * scale, "cl": 0.68268949 if cl is None else cl}
            self.params[p][name] = err
            out[p] = err
        return out

    def errors(self, params=None, method=None, cl=None, name=None, *, sigma=None, rtol=None, covariance_method=None, **_):
        """Asymmetric (profile-likelihood) uncertainties (``fitresult.py:1464-1598``).

        ``method``: ``"zfit_errors"`` (default here; alias ``"zfit_error"``) is the reference's iminuit-free route,
        ``compute_errors`` (``minimizers/errors.py:50-277``, restated in ``zfit_b200/errors.py``): one root search over all
        parameters per interval end, every function call one fused value+gradient pass; ``"refit"`` brackets the crossing
        with a fit per profile point instead (each with the parameter fixed).  The reference's default ``minuit_minos``
        needs iminuit.  Returns ``(errors, new_result)``: ``new_result`` is a fresh ``FitResult`` when the scan met a
        point below this result's minimum -- this result is then marked invalid, as in the reference."""
        allp = list(self.params)
        params = allp if params is None else ([params] if isinstance(params, ZfitParameter) else list(params))
        method = method or "zfit_errors"
        if method in ("zfit_error", "zfit_errors"):
            from .errors import compute_errors

            name = name or "errors"
            cl_used = 0.68268949 if cl is None and sigma is None else cl
            with set_values(allp, [self.params[p]["value"] for p in allp]):
                out, new_result = compute_errors(self, params, cl=cl, rtol=rtol, sigma=sigma, covariance_method=covariance_method)
            for p in out:
                out[p]["cl"] = None if cl_used is None else round(cl_used, 3)
            if new_result is not None:
                msg = "Invalid, a new minimum was found."
                for p in params:
                    self.params[p][name] = msg
                    if p in new_result.params and p in out:
                        new_result.params[p][name] = out[p]
                self._valid = False
                self.message = msg
                return {p: self.params[p][name] for p in params}, new_result
            for p in params:
                self.params[p][name] = out[p]
            return {p: self.params[p][name] for p in params}, None
        if callable(method):
            out, new_result = method(result=self, params=params, cl=cl)
            for p in out:
                self.params[p][name or "errors"] = out[p]
            return out, new_result
        if method != "refit":
            msg = f"The following method is not a valid, implemented method: {method}. Use one of ['zfit_errors', 'refit']"
            raise KeyError(msg)
        return self._errors_by_refits(params, cl, name)

    def _errors_by_refits(self, params, cl, name):
        """Roots of ``NLL_profiled(theta_i) - NLL_min = errordef * n_sigma^2`` by bracketing, each profile point being a
        fit with ``theta_i`` fixed (what MINOS does; costs a fit per function call)."""
        from scipy import optimize, stats

        allp = list(self.params)
        name = name or "errors"
        cl = 0.68268949 if cl is None else cl
        nsig = stats.norm.ppf(0.5 + cl / 2)
        target = self.loss.errordef * nsig**2
        hesse = self.hesse(params=allp, name="_hesse_for_errors")
        best = {p: self.params[p]["value"] for p in allp}
        fmin = None
        with set_values(allp, list(best.values())):
            fmin = float(self.loss.value(full=False))
        minimizer = self.minimizer if self.minimizer is not None else Minuit(gradient="zfit")
        out = {}
        new_result = None
        for p in params:
            sigma = hesse[p]["error"]
            res = {}
            for sign, key in ((-1.0, "lower"), (1.0, "upper")):
                def profile(v, p=p):
                    others = [q for q in allp if q is not p]
                    with set_values(allp, [best[q] for q in allp]):
                        p.set_value(v, clip=True)
                        if others:
                            was = p.floating
                            p.floating = False
                            try:
                                r = minimizer.minimize(self.loss, params=others)
                                val = float(r.fminopt)
                            finally:
                                p.floating = was
                        else:
                            val = float(self.loss.value(full=False))
                    return val - fmin - target

                a = best[p]
                b = best[p] + sign * sigma
                fb = profile(b)
                k = 0
                while fb < 0 and k < 8:
                    b = best[p] + sign * sigma * (1.5 ** (k + 1)) * 1.2
                    fb = profile(b)
                    k += 1
                root = optimize.brentq(profile, min(a, b), max(a, b), xtol=1e-4 * sigma) if fb > 0 else b
                res[key] = root - best[p]
            res["cl"] = cl
            self.params[p][name] = res
            out[p] = res
        self.update_params()
        for p in allp:
            self.params[p].pop("_hesse_for_errors", None)
        return out, new_result

    def __str__(self):
        lines = [f"FitResult of {self.loss!r}", f"with {self.minimizer!r}", ""]
        lines.append(f"valid={self.valid}  converged={self.converged}  param at limit={self.params_at_limit}  "
                     f"edm={self.edm if self.edm is None else format(self.edm, '.2g')}  "
                     f"approx. fmin (full | opt.)={self._fmin if self._fmin is None else format(self._fmin, '.6f')} | "
                     f"{self.fminopt if self.fminopt is None else format(self.fminopt, '.6f')}")
        lines.append("")
        lines.append("Parameters")
        errnames = sorted({k for v in self.params.values() for k in v if k != "value"})
        head = f"{'name':<14}{'value':>16}" + "".join(f"{e:>24}" for e in errnames) + f"{'at limit':>10}"
        lines.append(head)
        lines.append("-" * len(head))
        for p, v in self.params.items():
            row = f"{p.name:<14}{v['value']:>16.8g}"
            for e in errnames:
                ev = v.get(e)
                if ev is None:
                    row += f"{'':>24}"
                elif "error" in ev:
                    row += f"{'+/- ' + format(ev['error'], '.3g'):>24}"
                else:
                    row += f"{format(ev['lower'], '.3g') + ' ' + format(ev['upper'], '+.3g'):>24}"
            row += f"{getattr(p, 'at_limit', False)!s:>10}"
            lines.append(row)
        return "\n".join(lines)

    __repr__ = __str__


class _Values(Mapping):
    def __init__(self, params):
        self._p = params

    def __getitem__(self, key):
        if isinstance(key, int):
            return list(self._p.values())[key]["value"]
        return self._p[key]["value"]

    def __iter__(self):
        return iter(self._p)

    def __len__(self):
        return len(self._p)

    def __array__(self, dtype=None, copy=None):
        return np.array([v["value"] for v in self._p.values()], dtype=dtype or np.float64)


# ------------------------------------------------------------------------------------------------
# minimizers (baseminimizer.py:136-711)
# ------------------------------------------------------------------------------------------------
class BaseMinimizer:
    _DEFAULTS = {"tol": 1e-3, "verbosity": 0, "strategy": DefaultStrategy, "criterion": EDM, "maxiter": "auto"}

    def __init__(self, tol=None, verbosity=None, criterion=None, strategy=None, minimizer_options=None, maxiter=None,
                 name=None):
        self.tol = self._DEFAULTS["tol"] if tol is None else tol
        self.verbosity = self._DEFAULTS["verbosity"] if verbosity is None else verbosity
        self.criterion = self._DEFAULTS["criterion"] if criterion is None else criterion
        strategy = self._DEFAULTS["strategy"] if strategy is None else strategy
        self.strategy = strategy() if isinstance(strategy, type) else strategy
        self.minimizer_options = {} if minimizer_options is None else dict(minimizer_options)
        self._maxiter = self._DEFAULTS["maxiter"] if maxiter is None else maxiter
        self.name = name or type(self).__name__

    def get_maxiter(self, n=None):
        if self._maxiter == "auto":
            return 5000 + 1000 * (n or 5)
        return self._maxiter

    def copy(self):
        return copy.copy(self)

    def create_evaluator(self, loss, params, numpy_converter=None, strategy=None):
        return LossEval(
            loss=loss, params=params, strategy=self.strategy if strategy is None else strategy,
            do_print=self.verbosity > 9, maxiter=self.get_maxiter(len(params)), numpy_converter=numpy_converter, full=False,
        )

    def create_criterion(self, loss, params):
        return self.criterion(tol=self.tol, loss=loss, params=params)

    def _check_convert_input(self, loss, params, init):
        if callable(loss) and not hasattr(loss, "get_params"):
            from .loss import SimpleLoss

            if params is None:
                msg = "params have to be given if the loss is a plain callable"
                raise ValueError(msg)
            loss = SimpleLoss(loss, params, errordef=getattr(loss, "errordef", 0.5))
        if isinstance(loss, FitResult):
            init = loss
            loss = init.loss
        if params is None:
            params = list(loss.get_params(floating=True))
        else:
            params = [params] if isinstance(params, ZfitParameter) else list(params)
            params = [p for p in params if p.floating]
        if not params:
            msg = "No parameter for minimization given/found. Cannot minimize."
            raise RuntimeError(msg)
        return loss, params, init

    def minimize(self, loss, params=None, init=None):
        """Fully minimize ``loss`` with respect to ``params`` (``baseminimizer.py:438-534``)."""
        loss, params, init = self._check_convert_input(loss, params, init)
        if init is not None:
            set_values(params, init, allow_partial=True)
        start = [p.value() for p in params]
        t0 = time.perf_counter()
        try:
            result = self._minimize(loss, params, init)
        except (FailMinimizeNaN, MaximumIterationReached) as error:
            if getattr(self.strategy, "fit_result", None) is not None:
                result = self.strategy.fit_result
            else:
                result = FitResult(
                    loss=loss, params={p: p.value() for p in params}, minimizer=self.copy(), valid=False, converged=False,
                    edm=None, fminopt=None, message=f"{type(error).__name__}", info={},
                )
        result.info["time_s"] = time.perf_counter() - t0
        from . import settings

        if settings.options.get("auto_update_params", True):
            result.update_params()
        else:
            assign_values(params, start)
        return result

    def _minimize(self, loss, params, init):
        raise NotImplementedError

    def __repr__(self):
        return f"<{type(self).__name__} {self.name} tol={self.tol}>"


_IMINUIT = None


def _have_iminuit() -> bool:
    """Probed once per process: a failing ``import`` is not cached by Python and costs a file-system search every time
    (~0.3 ms of a 1.4 ms host budget for a small fit)."""
    global _IMINUIT
    if _IMINUIT is None:
        try:
            import iminuit  # noqa: F401

            _IMINUIT = True
        except ImportError:
            _IMINUIT = False
    return _IMINUIT


class Minuit(BaseMinimizer):
    """``zfit.minimize.Minuit(tol=None, mode=None, gradient=None, verbosity=None, options=None, maxiter=None,
    criterion=None, strategy=None, name=None)`` (``minimizer_minuit.py:35-165``).

    ``gradient=None/True`` -> the driver's own numerical gradient (Minuit's default, :139-141);
    ``gradient="zfit"/False`` -> ``loss.gradient`` (the fused kernel's analytic gradient).
    Uses iminuit when it is importable; this image has no iminuit, so the native MIGRAD-protocol driver runs instead
    (C++: ``csrc/znll_migrad.cpp`` through ``_migrad_native.py``; ``options={"driver": "python"}`` selects the Python
    statement of the same algorithm in ``_migrad.py``) and ``result.info["driver"]`` says which one ran.
    """

    def __init__(self, tol=None, mode=None, gradient=None, verbosity=None, options=None, maxiter=None, criterion=None,
                 strategy=None, name=None, use_minuit_grad=None, minuit_grad=None, minimize_strategy=None, ncall=None):
        if minimize_strategy is not None:
            mode = minimize_strategy
        if use_minuit_grad is not None:
            gradient = use_minuit_grad
        if minuit_grad is not None:
            gradient = minuit_grad
        self.mode = 1 if mode is None else mode
        if self.mode not in (0, 1, 2):
            msg = f"mode has to be 0, 1 or 2, not {mode}"
            raise ValueError(msg)
        self.minuit_grad = True if gradient is None else (gradient not in ("zfit", False))
        options = {} if options is None else dict(options)
        options.setdefault("ncall", ncall)
        options.setdefault("strategy", self.mode)
        super().__init__(tol=tol, verbosity=verbosity, criterion=criterion, strategy=strategy,
                         minimizer_options=options, maxiter=maxiter, name=name or "Minuit")

    def _minimize(self, loss, params, init):
        if _have_iminuit():
            return self._minimize_iminuit(loss, params, init)
        return self._minimize_native(loss, params, init)

    # -- the reference's route, used whenever iminuit exists (minimizer_minuit.py:175-319)
    def _minimize_iminuit(self, loss, params, init):
        import iminuit

        evaluator = self.create_evaluator(loss, params, numpy_converter=np.array)
        criterion = self.create_criterion(loss, params)
        grad_func = evaluator.gradient if not self.minuit_grad else None
        m = iminuit.Minuit(evaluator.value, np.array([p.value() for p in params]), grad=grad_func,
                           name=[p.name for p in params])
        for p in params:
            if p.has_stepsize:
                m.errors[p.name] = p.stepsize
            if p.has_limits:
                m.limits[p.name] = (p.lower, p.upper)
        m.errordef = loss.errordef
        m.strategy = self.mode
        m.tol = self.tol / 0.002 / loss.errordef
        valid, message, maxiter_reached = False, "", False
        for _ in range(20):
            try:
                m = m.migrad(ncall=self.minimizer_options.get("ncall"))
            except MaximumIterationReached:
                maxiter_reached, message = True, "Maxiter reached"
            criterion.last_value = m.fmin.edm
            converged = not m.fmin.is_above_max_edm
            if converged or maxiter_reached:
                assign_values(params, np.array(m.values))
                valid = not maxiter_reached
                break
        cov = np.array(m.covariance) if m.covariance is not None else None
        return FitResult(
            loss=loss, params=dict(zip(params, np.array(m.values))), minimizer=self.copy(), valid=valid and m.valid,
            converged=converged, edm=m.fmin.edm, fminopt=m.fval, message=message, criterion=criterion,
            approx={"inv_hessian": cov}, evaluator=evaluator,
            info={"n_eval": evaluator.niter, "driver": "iminuit", "original": m, "minuit": m},
        )

    @staticmethod
    def _information_seed(loss, params, evaluator):
        """Seed metric for the native driver: the loss's summed score outer product (one extra fused pass)."""
        fn = getattr(loss, "information_matrix", None)
        if fn is None:
            return None

        def information(x):
            assign_values(params, x)
            evaluator.npass += 1
            return fn(params)

        return information

    def _minimize_in_native_code(self, loss, params, steps, evaluator, criterion):
        return _native_loop_fit(self, loss, params, steps, evaluator, criterion)

    def _minimize_native(self, loss, params, init):
        evaluator = self.create_evaluator(loss, params, numpy_converter=np.array)
        criterion = self.create_criterion(loss, params)
        grad = None if self.minuit_grad else evaluator.gradient
        steps = [p.stepsize if p.has_stepsize else max(0.01, 0.1 * abs(p.value())) * (0.1 if p.has_limits else 1.0) for p in params]
        for i, p in enumerate(params):
            if p.has_limits and p.lower is not None and p.upper is not None and not p.has_stepsize:
                steps[i] = min(steps[i], 0.1 * (p.upper - p.lower))
        # the C++ driver runs unless the Python statement of the same algorithm is asked for (or its per-iteration trace)
        use_python = self.minimizer_options.get("driver") == "python" or self.verbosity > 6
        if not use_python and not self.minuit_grad and self.minimizer_options.get("native_loop", True):
            res = self._minimize_in_native_code(loss, params, steps, evaluator, criterion)
            if res is not None:
                return res
        driver = _migrad if use_python else _migrad_native
        res = driver.minimize(
            evaluator.value, grad, [p.value() for p in params], steps, [p.lower for p in params],
            [p.upper for p in params], edm_goal=self.tol, maxfcn=evaluator.maxiter or 100000, hesse=self.mode >= 1,
            verbose=self.verbosity > 6, value_gradient=None if self.minuit_grad else evaluator.value_gradient,
            information=None if self.minuit_grad else self._information_seed(loss, params, evaluator),
        )
        assign_values(params, res.x)
        criterion.last_value = res.edm
        return FitResult(
            loss=loss, params=dict(zip(params, res.x)), minimizer=self.copy(), valid=res.valid, converged=res.converged,
            edm=res.edm, fminopt=res.fval, message=res.message, criterion=criterion,
            approx={"inv_hessian": res.covariance * (2.0 * loss.errordef), "gradient": res.grad}, evaluator=evaluator,
            info={"n_eval": evaluator.niter, "n_value": evaluator.nfunc_eval, "n_gradient": evaluator.ngrad_eval,
                  "n_pass": evaluator.npass, "n_iter": res.niter,
                  "driver": ("python-migrad" if driver is _migrad else "native-migrad (C++)") + " (iminuit not installed)"},
        )


def _native_loop_fit(minimizer, loss, params, steps, evaluator, criterion):
    """The whole MIGRAD minimisation inside ``libznll.so`` (``znll_migrad_theta``) when the loss has the C-level theta
    path (built-in parameter kinds, no constraints, CUDA engine) and the minimizer's strategy is the default one: the
    driver, the slot map, the chain rule and the extended term are all native, so no interpreter runs between the
    fit's first and last fused pass.  Returns None when that does not apply or an evaluation was not finite -- the
    evaluator route with its NaN strategy then runs from the same start values."""
    if type(minimizer.strategy) is not DefaultStrategy or getattr(loss, "constraints", None) or not hasattr(loss, "_theta_path"):
        return None
    if evaluator.do_print:
        return None
    try:
        loss.check_precompile()
        eng = loss._get_engine()
    except Exception:
        return None
    if not getattr(eng, "is_cuda_engine", False) or eng._lazy_sort_in is not None:
        return None
    if eng.distributed and eng.reduce_mode != "fused-nvlink":  # every rank runs the same native loop on the numbers the
        return None                                             # kernel reduced over NVLink; the NCCL route is driven from Python
    path = loss._theta_path(eng)
    if path is None:
        return None
    idx = path.index_of(params)
    if np.any(idx < 0):
        return None
    log_offset = loss.options.get("subtr_const_value", False)
    full = log_offset is False or log_offset is None
    theta0 = np.array([p.value() for p in path.indep], dtype=np.float64)
    start = theta0[idx].copy()
    launches0 = eng.launch_count()
    res, theta = _migrad_native.minimize_theta(
        eng.plan, idx, theta0, steps, [p.lower for p in params], [p.upper for p in params], edm_goal=minimizer.tol,
        maxfcn=evaluator.maxiter or 100000, hesse=minimizer.mode >= 1, full=full, log_offset=0.0 if full else float(log_offset),
        use_information=True, stream=eng._raw_stream())
    if res is None:
        assign_values(params, start)
        return None
    assign_values(params, res.x)
    criterion.last_value = res.edm
    npass = (eng.launch_count() - launches0) // max(1, len(loss.model))
    evaluator._nfunc_eval, evaluator._ngrad_eval, evaluator.npass = int(res.nfcn), int(res.ngrad), int(npass)
    return FitResult(
        loss=loss, params=dict(zip(params, res.x)), minimizer=minimizer.copy(), valid=res.valid, converged=res.converged,
        edm=res.edm, fminopt=res.fval, message=res.message, criterion=criterion,
        approx={"inv_hessian": res.covariance * (2.0 * loss.errordef), "gradient": res.grad}, evaluator=evaluator,
        info={"n_eval": max(int(res.nfcn), int(res.ngrad)), "n_value": int(res.nfcn), "n_gradient": int(res.ngrad),
              "n_pass": int(npass), "n_iter": res.niter,
              "driver": "native-migrad (C++), whole fit in native code (iminuit not installed)"},
    )


class ScipyBaseMinimizer(BaseMinimizer):
    """scipy.optimize.minimize driven through ``LossEval`` (``minimizers_scipy.py:230-241``)."""

    _METHOD = None
    _USE_GRAD = True
    _USE_HESS = False
    _USE_BOUNDS = True

    def __init__(self, tol=None, gradient=None, hessian=None, verbosity=None, maxiter=None, criterion=None,
                 strategy=None, name=None, options=None, **kw):
        self._grad_choice = gradient
        super().__init__(tol=tol, verbosity=verbosity, criterion=criterion, strategy=strategy,
                         minimizer_options=options or {}, maxiter=maxiter, name=name or type(self).__name__)

    def _minimize(self, loss, params, init):
        from scipy import optimize

        evaluator = self.create_evaluator(loss, params, numpy_converter=np.array)
        criterion = self.create_criterion(loss, params)
        x0 = np.array([p.value() for p in params])
        bounds = None
        if self._USE_BOUNDS and any(p.has_limits for p in params):
            bounds = [(p.lower, p.upper) for p in params]
        kwargs = {"method": self._METHOD, "options": dict(self.minimizer_options)}
        if bounds is not None:
            kwargs["bounds"] = bounds
        use_grad = self._USE_GRAD and self._grad_choice not in (False, "2-point", "3-point")
        if use_grad:
            fun, kwargs["jac"] = evaluator.value_gradient, True
        else:
            fun = evaluator.value
            if self._grad_choice in ("2-point", "3-point"):
                kwargs["jac"] = self._grad_choice
        if self._USE_HESS:
            kwargs["hess"] = evaluator.hessian
        edm, converged, valid, message = None, False, False, ""
        result_params = None
        inv_hess = None
        tol_internal = self.tol
        for _ in range(12):
            kwargs["tol"] = tol_internal * 1e-2
            opt = optimize.minimize(fun, x0, **kwargs)
            x0 = opt.x
            assign_values(params, x0)
            approx = {}
            hinv = getattr(opt, "hess_inv", None)
            if hinv is not None and not isinstance(hinv, np.ndarray) and hasattr(hinv, "todense"):
                hinv = hinv.todense()
            if hinv is not None:  # the minimizer's own metric (minimizers_scipy.py: approx of the temporary and final result)
                hinv = np.asarray(hinv, dtype=np.float64)
                if hinv.shape == (len(params), len(params)) and np.all(np.isfinite(hinv)):
                    inv_hess = hinv
                    approx = {"inv_hessian": inv_hess}
                    jac = _dense_vector(getattr(opt, "jac", None), len(params))
                    if jac is not None:
                        approx["gradient"] = jac
            tmp = FitResult(loss=loss, params=dict(zip(params, x0)), minimizer=self, valid=True, converged=True,
                            edm=None, fminopt=opt.fun, approx=approx)
            with evaluator.ignore_maxiter():
                edm = criterion.calculate(tmp)
            converged = edm < criterion.tol
            message = str(opt.message)
            if converged or evaluator.maxiter_reached:
                valid = converged
                break
            tol_internal *= 0.1
        return FitResult(
            loss=loss, params=dict(zip(params, x0)), minimizer=self.copy(), valid=valid, converged=converged, edm=edm,
            fminopt=float(opt.fun), message=message, criterion=criterion,
            approx={"inv_hessian": inv_hess, "gradient": _dense_vector(getattr(opt, "jac", None), len(params))},
            evaluator=evaluator, info={"n_eval": evaluator.niter, "original": opt, "driver": f"scipy:{self._METHOD}"},
        )


def _dense_vector(v, n):
    """scipy's ``OptimizeResult.jac`` as a dense length-n vector, or None (trust-constr may hand back sparse / stacked forms)."""
    if v is None:
        return None
    try:
        if hasattr(v, "toarray"):
            v = v.toarray()
        a = np.asarray(v, dtype=np.float64).reshape(-1)
    except (TypeError, ValueError):
        return None
    return a if a.size == n and np.all(np.isfinite(a)) else None


def _scipy(name, method, grad=True, hess=False, bounds=True):
    return type(name, (ScipyBaseMinimizer,), {"_METHOD": method, "_USE_GRAD": grad, "_USE_HESS": hess, "_USE_BOUNDS": bounds})


ScipyLBFGSB = _scipy("ScipyLBFGSB", "L-BFGS-B")
ScipyLBFGSBV1 = ScipyLBFGSB
ScipyBFGS = _scipy("ScipyBFGS", "BFGS", bounds=False)
ScipyTrustConstr = _scipy("ScipyTrustConstr", "trust-constr")
ScipyTrustConstrV1 = ScipyTrustConstr
ScipySLSQP = _scipy("ScipySLSQP", "SLSQP")
ScipySLSQPV1 = ScipySLSQP
ScipyTruncNC = _scipy("ScipyTruncNC", "TNC")
ScipyTruncNCV1 = ScipyTruncNC
ScipyNewtonCG = _scipy("ScipyNewtonCG", "Newton-CG", hess=True, bounds=False)
ScipyNewtonCGV1 = ScipyNewtonCG
ScipyPowell = _scipy("ScipyPowell", "Powell", grad=False)
ScipyPowellV1 = ScipyPowell
ScipyNelderMead = _scipy("ScipyNelderMead", "Nelder-Mead", grad=False)
ScipyNelderMeadV1 = ScipyNelderMead
