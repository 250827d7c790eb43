"""NLL assembly restated from ``src/zfit/core/loss.py`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

``unbinned_nll``        -> ``_unbinned_nll_tf`` + ``_nll_calc_unbinned_tf`` (``loss.py:73-142``) and the
                           two ``_loss_func`` bodies (``loss.py:1134-1158`` / ``1291-1318``).
``initial_log_offset``  -> ``check_precompile`` (``loss.py:383-411``).
``value_and_gradient``  -> ``autodiff_value_gradient`` (``z/math.py:236-261``) on torch's tape.
``exact_nll``           -> extended-precision adjudication (longdouble per-event terms + ``math.fsum``).
"""

from __future__ import annotations

import math

import numpy as np

from .backend import NumpyBackend, TorchBackend

LOG_EPS = 1e-307  # loss.py:117


def kahan_sum_np(x):
    """``tfp.math.reduce_kahan_sum`` restated (third-party, tensorflow-probability >=0.24,<1.0): returns
    ``(total, correction)`` whose difference ``total - correction`` is the compensated sum.  The compensated
    sum equals the correctly-rounded sum to within an ulp, so it is restated as ``math.fsum`` here:
    ``total`` = the plain fp64 sum, ``correction`` = ``total - fsum(x)``."""
    x = np.asarray(x, dtype=np.float64)
    total = np.sum(x)
    return total, total - math.fsum(x)


def _channel_logprobs(B, model, x, P):
    probs = model.pdf(B, x, P)  # loss.py:115
    return B.log(probs + B.const(LOG_EPS))  # loss.py:117


def log_poisson_loss(B, targets, log_input, compute_full_loss):
    """``tf.nn.log_poisson_loss`` restated (third-party, tensorflow >=2.16.2,<3.0):
    ``exp(log_input) - log_input*targets`` [+ Stirling ``t*log t - t + 0.5*log(2*pi*t)``, zeroed for 0<=t<=1]."""
    result = B.exp(log_input) - log_input * targets
    if compute_full_loss:
        t = float(targets)
        if not (0.0 <= t <= 1.0):
            tt = B.const(t)
            result = result + (tt * B.log(tt) - tt + B.const(0.5) * B.log(B.const(2.0 * math.pi) * tt))
    return result


def unbinned_nll(models, datas, P, *, weights=None, extended=False, log_offset=False, kahan=False, B=None):
    """NLL of a (simultaneous) unbinned fit.

    models: list of oracle pdf trees; datas: list of {obs: 1-D array}; weights: list of 1-D array | None;
    ``log_offset``: ``False`` (== ``full=True``) or a number; ``kahan``: UnbinnedNLL's ``sumtype=="kahan"``.
    Extended never uses Kahan (``_value`` never forwards ``sumtype``, loss.py:617-624 vs 1291).
    """
    B = B or NumpyBackend()
    if weights is None:
        weights = [None] * len(models)
    nll = None
    nll_corr = None
    for model, x, w in zip(models, datas, weights):
        x = {k: B.asarray(v) for k, v in x.items()}
        log_probs = _channel_logprobs(B, model, x, P)
        if w is not None:
            log_probs = log_probs * B.asarray(w)  # loss.py:134-135
        if log_offset is not False:
            log_probs = log_probs - B.const(log_offset) if not hasattr(log_offset, "dtype") else log_probs - log_offset
        if kahan and not extended and B.name == "numpy":
            s, c = kahan_sum_np(log_probs)
        else:
            s, c = B.sum(log_probs), B.const(0.0)
        nll = -s if nll is None else nll + (-s)
        nll_corr = -c if nll_corr is None else nll_corr + (-c)
    if extended:  # loss.py:1301-1317
        for model, x, w in zip(models, datas, weights):
            n = float(np.sum(np.asarray(w, dtype=np.float64))) if w is not None else float(len(next(iter(x.values()))))
            y = model.get_yield(B, P)
            term = log_poisson_loss(B, B.const(n), B.log(y), compute_full_loss=log_offset is False)
            if log_offset is not False:
                term = term + (B.const(log_offset) if not hasattr(log_offset, "dtype") else log_offset)
            nll = nll + term
    return nll - nll_corr


def initial_log_offset(models, datas, P, *, weights=None, extended=False, kahan=False):
    """``check_precompile`` (loss.py:393-409): ``-(NLL(theta0; log_offset=0.0) - 10000) / sum(num_entries)``."""
    n_tot = sum(len(next(iter(x.values()))) for x in datas)
    v = unbinned_nll(models, datas, P, weights=weights, extended=extended, log_offset=0.0, kahan=kahan)
    return -float((v - 10000.0) / n_tot)


def value_and_gradient(models, datas, theta, names, *, weights=None, extended=False, log_offset=False, threads=None):
    """Value + reverse-mode gradient on torch's tape over the same op graph (fp64, CPU)."""
    import torch

    if threads:
        torch.set_num_threads(int(threads))
    B = TorchBackend()
    leaves = [torch.tensor(float(v), dtype=torch.float64, requires_grad=True) for v in theta]
    P = dict(zip(names, leaves))
    tdatas = [{k: B.asarray(v) for k, v in x.items()} for x in datas]
    tweights = None if weights is None else [None if w is None else B.asarray(w) for w in weights]
    val = unbinned_nll(models, tdatas, P, weights=tweights, extended=extended, log_offset=log_offset, B=B)
    grads = torch.autograd.grad(val, leaves, allow_unused=True)
    g = np.array([0.0 if gi is None else float(gi) for gi in grads])
    return float(val.detach()), g


def exact_nll(models, datas, P, *, weights=None, extended=False):
    """Extended-precision ``value(full=True)``: per-event terms in longdouble (integrals via mpmath erf),
    summed with ``math.fsum``-like exactness (longdouble pairwise sum of longdouble terms)."""
    B = NumpyBackend(np.longdouble)
    PL = {k: np.longdouble(v) for k, v in P.items()}
    if weights is None:
        weights = [None] * len(models)
    total = np.longdouble(0.0)
    abs_total = np.longdouble(0.0)
    for model, x, w in zip(models, datas, weights):
        xl = {k: np.asarray(v, dtype=np.longdouble) for k, v in x.items()}
        lp = _channel_logprobs(B, model, xl, PL)
        if w is not None:
            lp = lp * np.asarray(w, dtype=np.longdouble)
        total = total - np.sum(lp, dtype=np.longdouble)
        abs_total = abs_total + np.sum(np.abs(lp), dtype=np.longdouble)
    if extended:
        for model, x, w in zip(models, datas, weights):
            n = float(np.sum(np.asarray(w, dtype=np.float64))) if w is not None else float(len(next(iter(x.values()))))
            y = model.get_yield(B, PL)
            total = total + log_poisson_loss(B, np.longdouble(n), np.log(y), True)
    return total, abs_total


def numeric_gradient(fun, theta, rel=1e-6):
    """Central differences with Richardson extrapolation (independent check of the tape)."""
    theta = np.asarray(theta, dtype=np.float64)
    g = np.zeros_like(theta)
    for i in range(theta.size):
        h = rel * max(1.0, abs(theta[i]))

        def d(hh):
            tp, tm = theta.copy(), theta.copy()
            tp[i] += hh
            tm[i] -= hh
            return (fun(tp) - fun(tm)) / (2 * hh)

        g[i] = (4.0 * d(h) - d(2 * h)) / 3.0
    return g
