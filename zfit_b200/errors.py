"""Profile-likelihood intervals without a profiling minimizer: ``compute_errors`` (``src/zfit/minimizers/errors.py:50-277``).

For one parameter of interest at a time the method solves, over ALL fit parameters at once, the square system

    loss(theta) - fmin - errordef * n_sigma^2 = 0
    d loss / d theta_j                        = 0    for every j other than the parameter of interest

with ``scipy.optimize.root`` (default ``hybr``) -- i.e. it walks along the profile and looks for the crossing in one
root search; every function evaluation is ONE fused value+gradient pass (``loss.value_gradient(params=all, full=False)``),
there are no inner fits.  Start values come from the covariance (a step of ``n_sigma`` along the parameter's row), the
residual is mirrored on the far side of the minimum so that only one zero is left, the search ends as soon as the
first component stays below ``rtol`` for more than three consecutive calls, and a point noticeably BELOW the fit's
minimum aborts everything: the loss is minimised again from there and the errors are computed for the new result, which
is handed back to the caller.
"""

from __future__ import annotations

import warnings

import numpy as np


class NewMinimum(Exception):
    """A point with a lower loss than the fit's minimum was met while profiling (``errors.py:29-31``)."""


class _RootFound(Exception):
    pass


def compute_errors(result, params, *, cl=None, rtol=None, method=None, covariance_method=None, sigma=None):
    """-> ({param: {"lower": ..., "upper": ...}}, new_result | None); see the module docstring."""
    from scipy import optimize, stats

    from .parameter import ZfitParameter

    def set_values(ps, vals):  # assign_values (evaluation.py:27-30): no limit check, Parameter.value() clips
        for p, v in zip(ps, vals):
            p.assign(float(v))

    rtol = 0.003 if rtol is None else rtol
    method = "hybr" if method is None else method
    if cl is None:
        sigma = 1 if sigma is None else sigma
    elif sigma is None:
        sigma = stats.chi2(1).ppf(cl) ** 0.5
    else:
        msg = "Cannot specify both sigma and cl."
        raise ValueError(msg)
    params = [params] if isinstance(params, ZfitParameter) else list(params)
    all_params = list(result.params.keys())
    loss = result.loss
    errordef = loss.errordef
    fmin = result.fminopt
    rtol_abs = rtol * errordef
    minimizer = result.minimizer

    cov = result.covariance(method=covariance_method, as_dict=True)
    if cov is None:
        cov = result.covariance(method="hesse_np", as_dict=True)
    if cov is None:
        msg = "Could not compute covariance matrix. Check if the minimum is valid."
        raise RuntimeError(msg)
    param_errors = {p: cov[(p, p)] ** 0.5 for p in all_params}
    param_scale = np.array(list(param_errors.values()))
    best = np.array([result.params[p]["value"] for p in all_params], dtype=np.float64)
    loss_min_tol = (minimizer.tol if minimizer is not None else 1e-3) * errordef * 2  # "2 is just to be tolerant"
    new_result = None
    out = {}
    try:
        for param in params:
            set_values(all_params, best)
            perr = param_errors[param]
            pval = result.params[param]["value"]
            ipoi = all_params.index(param)
            start = {"lower": [], "upper": []}
            for ap, apv in zip(all_params, best):
                factor = cov[(param, ap)] * (2 * errordef / perr**2) ** 0.5
                for d, sgn in (("lower", -1.0), ("upper", 1.0)):
                    step = sgn * factor * sigma
                    for _ in range(50):
                        cand = apv + step
                        lo = -np.inf if getattr(ap, "lower", None) is None else ap.lower
                        up = np.inf if getattr(ap, "upper", None) is None else ap.upper
                        if cand < lo or cand > up:
                            step *= 0.8
                        else:
                            break
                    else:
                        msg = f"Could not find a valid initial value for {ap} in {d} direction."
                        raise RuntimeError(msg)
                    start[d].append(cand)
            loss.value_gradient(params=all_params, full=False)  # offsets fixed, kernels loaded
            state = {"root": None, "ntol": 999}

            def residual(values, swap_sign, ipoi=ipoi, param=param, state=state):
                set_values(all_params, values)
                value, grad = loss.value_gradient(params=all_params, full=False)
                value, grad = float(value), np.asarray(grad, dtype=np.float64)
                if not np.isfinite(value):  # the reference maps a failed evaluation to a huge loss and a random push
                    value, grad = 9999999999.0, np.random.default_rng(0).normal(scale=0.1, size=grad.shape)
                grad = np.concatenate([grad[:ipoi], grad[ipoi + 1:]])
                zeroed = value - fmin
                if swap_sign(values[ipoi]):  # mirror on the far side of the minimum: one zero only
                    zeroed, grad = -zeroed, -grad
                elif zeroed < -loss_min_tol:
                    raise NewMinimum("A new minimum is found.")  # the parameters stay at the new point
                shifted = zeroed - errordef * sigma**2
                if abs(shifted) < rtol_abs:
                    if state["ntol"] > 3:
                        state["root"] = values[ipoi]
                        raise _RootFound
                    state["ntol"] += 1
                else:
                    state["ntol"] = 0
                return np.concatenate([[shifted], grad])

            swaps = {"lower": lambda v, pval=pval: v > pval, "upper": lambda v, pval=pval: v < pval}
            out[param] = {}
            for d in ("lower", "upper"):
                state["root"], state["ntol"] = None, 999
                try:
                    sol = optimize.root(residual, x0=np.array(start[d]), args=(swaps[d],), tol=rtol_abs * 0.1, method=method,
                                        options={"diag": 1 / param_scale} if method in ("hybr", "lm") else {})
                except _RootFound:
                    root = state["root"]
                else:
                    warnings.warn(f"The root finding did not converge below {rtol_abs} but stopped by its own criteria.",
                                  RuntimeWarning, stacklevel=2)
                    root = sol.x[ipoi]
                out[param][d] = float(root - pval)
        set_values(all_params, best)
    except NewMinimum:
        new_found = float(loss.value(full=False))  # the parameters sit at the lower point
        new_result = minimizer.minimize(loss)
        if new_result.fminopt >= new_found + loss_min_tol:
            msg = ("A new minimum was discovered but the minimizer was not able to find this on himself. "
                   "This behavior is currently an exception but will most likely change in the future.")
            raise RuntimeError(msg) from None
        out, again = compute_errors(new_result, params, cl=cl, rtol=rtol, method=method, covariance_method=covariance_method)
        if again is not None:
            new_result = again
    return out, new_result
