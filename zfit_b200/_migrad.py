"""Native variable-metric minimizer with Minuit's conventions (stand-in driver when ``iminuit`` is absent).

iminuit (C++ Minuit2) is what ``zfit.minimize.Minuit`` wraps (``src/zfit/minimizers/minimizer_minuit.py:175-319``);
it is not installed in this image and cannot be.  This module implements the same *protocol* so that
"Minuit fit wall time" is reportable and the loss is driven the way MIGRAD drives it:

* parameter limits through Minuit's internal transformations (sin for two-sided, sqrt for one-sided),
* a variable-metric (DFP with Minuit's rank-2 correction, ``DavidonErrorUpdator``) iteration with a
  parabolic line search,
* convergence on the estimated distance to minimum ``EDM = g^T V g / 2 < 0.002 * tol * errordef``
  (with zfit's ``tol`` pre-divided by ``0.002 * errordef``, :313-315, that is ``EDM < zfit tol``),
* Minuit2's bookkeeping of the metric's uncertainty (``Dcovar``: 1 for a seed, 0 after HESSE, averaged with the relative
  size of every DFP update): the stop test is ``EDM (1 + 3 Dcovar) < goal`` and, as in strategy 1, HESSE (differences of
  the analytic gradient: P passes forward, 2P central when the forward matrix is not symmetric to 1e-3) verifies the
  result only while ``Dcovar > 0.05``; MIGRAD continues from the HESSE metric when the verified EDM is above the goal,
* when the loss can supply its information matrix (summed outer product of the per-event scores, ONE fused pass), that
  seeds the metric instead of Minuit's diagonal seed (P passes, no correlations) and refreshes it once when the
  inflated EDM first drops below 1 (about one sigma from the minimum, where the information matrix is the Hessian up
  to O(1/sqrt(N))).  The refreshed metric is credited with ``Dcovar = 0.05``, the threshold of the strategy-1 rule: if
  the DFP updates that follow are small (relative size <= 0.05, i.e. the metric was right) the fit ends without HESSE,
  exactly as Minuit2 ends a fit whose metric it trusts; larger updates push ``Dcovar`` back over the threshold and HESSE
  verifies as usual.  A B200-side addition: the pass is cheap here and removes most DFP iterations and the P HESSE passes,
* ``grad=None`` -> the driver's own finite-difference gradient of ``value`` (Minuit's default, :139-141).

The driver only sees ``value(x)`` and ``gradient(x)`` callbacks of a ``LossEval``.
"""

from __future__ import annotations

import math

import numpy as np


FAR = 0.5  # a second line-search probe only when the secant puts the line minimum this far from the unit step
REFRESH_EDM = 1.0  # refresh the metric from the information matrix once the inflated EDM drops below this
REFRESH_DCOVAR = 0.05  # relative uncertainty credited to the refreshed metric (the HESSE threshold of strategy 1)


class _Transform:
    """Minuit2 ``SinParameterTransformation`` / ``SqrtLow/UpParameterTransformation``."""

    def __init__(self, lower, upper):
        self.lo = np.array([-np.inf if l is None else l for l in lower], dtype=np.float64)
        self.up = np.array([np.inf if u is None else u for u in upper], dtype=np.float64)
        self.both = np.isfinite(self.lo) & np.isfinite(self.up)
        self.low_only = np.isfinite(self.lo) & ~np.isfinite(self.up)
        self.up_only = ~np.isfinite(self.lo) & np.isfinite(self.up)

    def to_internal(self, x):
        x = np.asarray(x, dtype=np.float64)
        y = x.copy()
        b = self.both
        if b.any():
            yy = 2.0 * (x[b] - self.lo[b]) / (self.up[b] - self.lo[b]) - 1.0
            y[b] = np.arcsin(np.clip(yy, -1.0, 1.0))
        m = self.low_only
        if m.any():
            y[m] = np.sqrt(np.maximum((x[m] - self.lo[m] + 1.0) ** 2 - 1.0, 0.0))
        m = self.up_only
        if m.any():
            y[m] = np.sqrt(np.maximum((self.up[m] - x[m] + 1.0) ** 2 - 1.0, 0.0))
        return y

    def to_external(self, y):
        y = np.asarray(y, dtype=np.float64)
        x = y.copy()
        b = self.both
        if b.any():
            x[b] = self.lo[b] + 0.5 * (self.up[b] - self.lo[b]) * (np.sin(y[b]) + 1.0)
        m = self.low_only
        if m.any():
            x[m] = self.lo[m] - 1.0 + np.sqrt(y[m] ** 2 + 1.0)
        m = self.up_only
        if m.any():
            x[m] = self.up[m] + 1.0 - np.sqrt(y[m] ** 2 + 1.0)
        return np.minimum(np.maximum(x, self.lo), self.up)  # rounding must not step an ulp outside the limits

    def dext_dint(self, y):
        y = np.asarray(y, dtype=np.float64)
        d = np.ones_like(y)
        b = self.both
        if b.any():
            d[b] = 0.5 * (self.up[b] - self.lo[b]) * np.cos(y[b])
        m = self.low_only
        if m.any():
            d[m] = y[m] / np.sqrt(y[m] ** 2 + 1.0)
        m = self.up_only
        if m.any():
            d[m] = -y[m] / np.sqrt(y[m] ** 2 + 1.0)
        return d


class MigradResult:
    def __init__(self):
        self.x = None
        self.fval = None
        self.edm = None
        self.covariance = None
        self.hessian = None
        self.grad = None
        self.valid = False
        self.converged = False
        self.nfcn = 0
        self.ngrad = 0
        self.message = ""
        self.niter = 0


def _num_grad(value, x, steps, fx):
    g = np.zeros_like(x)
    for i in range(x.size):
        h = steps[i]
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        fp, fm = value(xp), value(xm)
        g[i] = (fp - fm) / (2 * h)
        # adapt the step like MnNumericalGradientCalculator: aim at |f''| h^2 ~ 8 eps^(2/3) |f| scale
        g2 = (fp + fm - 2 * fx) / (h * h)
        if np.isfinite(g2) and abs(g2) > 0:
            hopt = math.sqrt(8.0 * 1e-11 * (abs(fx) + 1.0) / abs(g2))
            steps[i] = min(max(hopt, 1e-9 * max(1.0, abs(x[i]))), 10 * h)
    return g


def minimize(value, gradient, x0, steps, lower, upper, *, edm_goal, maxfcn=10000, hesse=True, verbose=False,
             value_gradient=None, information=None):
    """Variable-metric minimisation in Minuit's internal coordinates.

    value(x_ext) -> float; gradient(x_ext) -> array or None (None: finite differences of value);
    value_gradient(x_ext) -> (float, array) or None: when the loss returns both in one pass (the fused kernel does),
    line-search probes use it, so an accepted step costs ONE pass instead of a value pass plus a gradient pass.
    information(x_ext) -> (P, P) array or None: an approximate Hessian of ``value`` in external coordinates that costs
    about one pass (for a likelihood: the summed outer product of the per-event scores).  When given it seeds the
    metric instead of Minuit's diagonal seed (which costs P gradient passes and knows no correlations).
    """
    tr = _Transform(lower, upper)
    res = MigradResult()
    counter = {"f": 0, "g": 0}

    def f_int(y):
        counter["f"] += 1
        return float(value(tr.to_external(y)))

    num_steps = None

    def g_int(y, fy=None):
        if gradient is not None:
            counter["g"] += 1
            return np.asarray(gradient(tr.to_external(y)), dtype=np.float64) * tr.dext_dint(y)
        if fy is None:
            fy = f_int(y)
        return _num_grad(f_int, y, num_steps, fy)

    def fg_int(y):
        counter["f"] += 1
        counter["g"] += 1
        fv, gv = value_gradient(tr.to_external(y))
        return float(fv), np.asarray(gv, dtype=np.float64) * tr.dext_dint(y)

    x0 = np.asarray(x0, dtype=np.float64)
    y = tr.to_internal(x0)
    n = y.size
    d0 = np.abs(tr.dext_dint(y))
    d0[d0 < 1e-8] = 1e-8
    ysteps = np.abs(np.asarray(steps, dtype=np.float64)) / d0
    ysteps = np.minimum(ysteps, 0.5)
    num_steps = np.maximum(ysteps * 0.1, 1e-8)

    if value_gradient is not None and gradient is not None:
        f, g = fg_int(y)
    else:
        f = f_int(y)
        g = g_int(y, f)

    def information_metric(y):
        if information is None:
            return None
        try:
            H = information(tr.to_external(y))
        except Exception:  # the seed is an optimisation; any trouble falls back to Minuit's own seed
            return None
        if H is None:
            return None
        H = np.asarray(H, dtype=np.float64)
        if H.shape != (n, n) or not np.all(np.isfinite(H)):
            return None
        J = tr.dext_dint(y)
        H = H * J[:, None] * J[None, :]
        H = 0.5 * (H + H.T)
        d = np.sqrt(np.maximum(np.diag(H), 1e-300))
        C = H / d[:, None] / d[None, :]
        try:
            w = np.linalg.eigvalsh(C)
            if w[0] < 1e-10 * w[-1]:
                return None
            return np.linalg.inv(C) / d[:, None] / d[None, :]
        except np.linalg.LinAlgError:
            return None

    def seed_metric(y, g, first=False):
        """Diagonal seed from second derivatives along each axis (MnSeedGenerator)."""
        if first:
            Vi = information_metric(y)
            if Vi is not None:
                return Vi
        V = np.zeros((n, n))
        for i in range(n):
            h = max(ysteps[i] * 0.05, 1e-7)
            yp = y.copy()
            yp[i] += h
            if gradient is not None:
                g2 = (g_int(yp)[i] - g[i]) / h
            else:
                ym = y.copy()
                ym[i] -= h
                g2 = (f_int(yp) + f_int(ym) - 2 * f) / (h * h)
            if not np.isfinite(g2) or g2 <= 0:
                g2 = 1.0 / max(ysteps[i] ** 2, 1e-16)
            V[i, i] = 1.0 / g2
        return V

    def line_search(y, step, f0, gdel):
        """Parabolic line search along `step` (MnLineSearch, simplified): returns (lambda, f) with f < f0 or (0, f0)."""
        pts = []

        def probe(lam):
            fl = f_int(y + lam * step)
            if np.isfinite(fl):
                pts.append((lam, fl))
            return fl

        f1 = probe(1.0)
        if np.isfinite(f1):
            denom = 2.0 * (f1 - f0 - gdel)
            lam2 = -gdel / denom if denom > 0 else 2.0
            lam2 = min(max(lam2, 0.02), 8.0)
        else:
            lam2 = 0.1
        if abs(lam2 - 1.0) > 0.02:
            f2 = probe(lam2)
            # one three-point refinement when both probes are finite and bracket something useful
            if np.isfinite(f1) and np.isfinite(f2):
                l1, l2 = 1.0, lam2
                a = ((f1 - f0) / l1 - (f2 - f0) / l2) / (l1 - l2)
                b = (f1 - f0) / l1 - a * l1
                if a > 0:
                    lam3 = min(max(-b / (2 * a), 0.01), 10.0)
                    if min(abs(lam3 - l1), abs(lam3 - l2)) > 0.05 * lam3:
                        probe(lam3)
        best = min(pts, key=lambda t: t[1]) if pts else (0.0, f0)
        lam = min([t[0] for t in pts], default=1.0)
        tries = 0
        while best[1] >= f0 and tries < 10:  # backtrack
            tries += 1
            lam *= 0.2
            probe(lam)
            best = min(pts, key=lambda t: t[1]) if pts else (0.0, f0)
        if best[1] >= f0:
            return 0.0, f0
        return best

    def line_search_fg(y, step, f0, gdel):
        """Line search with value+gradient probes: the secant of the directional derivative locates the minimum of the
        quadratic model along the step; returns (lambda, f, g) with f < f0, or (0, f0, None)."""
        cands = []

        def probe(lam):
            fl, gl = fg_int(y + lam * step)
            if np.isfinite(fl) and np.all(np.isfinite(gl)):
                cands.append((lam, fl, gl))
                return fl, float(gl @ step)
            return np.inf, None

        f1, d1 = probe(1.0)
        if d1 is not None:
            lam2 = gdel / (gdel - d1) if d1 > gdel else 2.5  # still descending at lambda = 1 -> extend
            lam2 = min(max(lam2, 0.05), 6.0)
            far = abs(lam2 - 1.0) > FAR
            if f1 >= f0 or far:
                probe(lam2)
        lam = min([c[0] for c in cands], default=1.0)
        tries = 0
        while (not cands or min(c[1] for c in cands) >= f0) and tries < 10:
            tries += 1
            lam *= 0.2
            probe(lam)
        if not cands:
            return 0.0, f0, None
        best = min(cands, key=lambda c: c[1])
        if best[1] >= f0:
            return 0.0, f0, None
        return best

    def verify(y, f, g, V):
        """HESSE in internal coordinates; returns (converged, V, edm)."""
        H = _hesse_internal(f_int, g_int, gradient is not None, y, f, g, V)
        try:
            w = np.linalg.eigvalsh(H)
            if not np.all(w > 0):
                return False, None, None
            Vh = np.linalg.inv(H)
        except np.linalg.LinAlgError:
            return False, None, None
        e = 0.5 * float(g @ Vh @ g)
        return e < edm_goal, Vh, e

    V = seed_metric(y, g, first=True)
    edm = 0.5 * float(g @ V @ g)
    dcovar = 1.0  # Minuit2's estimate of the relative error of V: 1 for a seed, 0 after HESSE, updated by DFP
    restarts = 0
    refreshed = False
    use_fg = value_gradient is not None and gradient is not None
    for it in range(1, 100000):
        res.niter = it
        if counter["f"] + counter["g"] > maxfcn:
            res.message = "call limit reached"
            break
        if not (np.isfinite(f) and np.all(np.isfinite(g))):
            res.message = "non-finite value or gradient"
            break
        if edm * (1.0 + 3.0 * dcovar) < edm_goal:  # VariableMetricBuilder: the EDM is inflated by the metric's uncertainty
            if not hesse or dcovar <= 0.05:
                # strategy 1: HESSE verifies the result only while the metric is still uncertain (Dcovar > 0.05)
                res.converged = True
                break
            ok, Vh, e = verify(y, f, g, V)
            if Vh is not None:
                V, edm, dcovar = Vh, e, 0.0
                if ok:
                    res.converged = True
                    break
            elif restarts < 3:
                restarts += 1
                V = seed_metric(y, g)
                edm = 0.5 * float(g @ V @ g)
                dcovar = 1.0
                if edm < edm_goal:  # the seed agrees that we are there; accept
                    res.converged = True
                    res.message = "converged (hesse not positive definite, seed metric used)"
                    break
            else:
                res.message = "hesse failed: matrix not positive definite"
                break
        if information is not None and not refreshed and it > 1 and edm * (1.0 + 3.0 * dcovar) < REFRESH_EDM:
            # within about one sigma of the minimum the score outer product IS the Hessian up to O(1/sqrt(N)): one pass
            # replaces the correlations DFP would need several more iterations to learn
            refreshed = True
            Vi = information_metric(y)
            if Vi is not None and float(g @ Vi @ g) > 0:
                V, dcovar = Vi, REFRESH_DCOVAR
                edm = 0.5 * float(g @ V @ g)
        step = -V @ g
        gdel = float(step @ g)
        if gdel >= 0:  # metric not positive definite: reseed
            V = seed_metric(y, g)
            dcovar = 1.0
            step = -V @ g
            gdel = float(step @ g)
            if gdel >= 0:
                res.message = "no descent direction"
                break
        g_acc = None
        if use_fg:
            lam, f_new, g_acc = line_search_fg(y, step, f, gdel)
        else:
            lam, f_new = line_search(y, step, f, gdel)
        if lam == 0.0:
            if restarts < 3:
                restarts += 1
                V = seed_metric(y, g)
                edm = 0.5 * float(g @ V @ g)
                dcovar = 1.0
                if edm < edm_goal:
                    continue
                step = -V @ g
                gdel = float(step @ g)
                if gdel < 0 and use_fg:
                    lam, f_new, g_acc = line_search_fg(y, step, f, gdel)
                elif gdel < 0:
                    lam, f_new = line_search(y, step, f, gdel)
            if lam == 0.0:
                res.message = "line search failed to improve"
                res.converged = edm < edm_goal
                break
        dy = lam * step
        y_new = y + dy
        g_new = g_acc if g_acc is not None else g_int(y_new, f_new)
        dg = g_new - g
        # ---- DavidonErrorUpdator -----------------------------------------------------------------
        delgam = float(dy @ dg)
        Vg = V @ dg
        gvg = float(dg @ Vg)
        if delgam > 0 and gvg > 0:
            upd = np.outer(dy, dy) / delgam - np.outer(Vg, Vg) / gvg
            if delgam > gvg:
                flnu = dy / delgam - Vg / gvg
                upd = upd + gvg * np.outer(flnu, flnu)
            V = V + upd
            dcovar = 0.5 * (dcovar + float(np.abs(upd).sum()) / max(float(np.abs(V).sum()), 1e-300))
        y, f, g = y_new, f_new, g_new
        edm = 0.5 * float(g @ V @ g)
        if verbose:
            print(f"[migrad] it={it} f={f:.10g} edm={edm:.3g} dcovar={dcovar:.3g} lam={lam:.3g} nf={counter['f']} ng={counter['g']}")
    x = tr.to_external(y)
    res.x = x
    res.fval = f
    res.edm = edm
    res.grad = g / np.where(np.abs(tr.dext_dint(y)) > 1e-300, tr.dext_dint(y), 1.0)
    J = tr.dext_dint(y)
    res.covariance = (V * J[:, None]) * J[None, :]
    res.nfcn, res.ngrad = counter["f"], counter["g"]
    res.valid = bool(res.converged and np.isfinite(f))
    if res.converged and not res.message:
        res.message = "converged"
    return res


def _hesse_internal(f_int, g_int, have_grad, y, f, g, V):
    n = y.size
    H = np.zeros((n, n))
    sig = np.sqrt(np.maximum(np.diag(V), 1e-300))
    if have_grad:
        # rows of the Hessian from differences of the analytic gradient.  Forward differences reuse g(y): P passes.
        # Their O(h) truncation error shows up as asymmetry of the matrix; when that exceeds 1e-3 (small samples, far from
        # quadratic) the backward points are added and the differences become central (2P passes).
        hs = np.maximum(0.02 * sig, 1e-8)
        gp = []
        for i in range(n):
            yp = y.copy()
            yp[i] += hs[i]
            gp.append(g_int(yp))
            H[i, :] = (gp[i] - g) / hs[i]
        d = np.sqrt(np.maximum(np.abs(np.diag(H)), 1e-300))
        C = H / d[:, None] / d[None, :]
        if np.all(np.isfinite(C)) and np.linalg.norm(C - C.T) <= 1e-3 * np.linalg.norm(C):
            return 0.5 * (H + H.T)
        for i in range(n):
            ym = y.copy()
            ym[i] -= hs[i]
            H[i, :] = (gp[i] - g_int(ym)) / (2 * hs[i])
        return 0.5 * (H + H.T)
    hs = np.maximum(0.1 * sig, 1e-7)
    fp = np.zeros(n)
    fm = np.zeros(n)
    for i in range(n):
        yp, ym = y.copy(), y.copy()
        yp[i] += hs[i]
        ym[i] -= hs[i]
        fp[i], fm[i] = f_int(yp), f_int(ym)
        H[i, i] = (fp[i] + fm[i] - 2 * f) / hs[i] ** 2
    for i in range(n):
        for j in range(i + 1, n):
            ypp = y.copy()
            ypp[i] += hs[i]
            ypp[j] += hs[j]
            H[i, j] = H[j, i] = (f_int(ypp) + f - fp[i] - fp[j]) / (hs[i] * hs[j])
    return H
