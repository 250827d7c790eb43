"""Op-for-op restatement of the reference PDFs on the hot path (TEST INFRASTRUCTURE).

All ``file:line`` citations are relative to the reference tree (zfit 0.28.0).
Every function takes a backend ``B`` (``oracle/backend.py``) so the same op graph runs in
numpy fp64, numpy longdouble and torch fp64 (autograd).

Parameters are referred to by *name* (looked up in the ``P`` mapping at call time) or given
as plain numbers (constants) -- this mirrors ``convert_to_parameter`` turning numbers into
``ConstantParameter`` (``src/zfit/core/parameter.py:1485-1561``).
"""

from __future__ import annotations

import math

import numpy as np

from .backend import SQRT2


def _val(B, P, p):
    """Value of a parameter reference: name -> P[name], number -> constant."""
    if isinstance(p, str):
        return P[p]
    if callable(p):
        return p(P)
    return B.const(p)


def _names(p):
    return [p] if isinstance(p, str) else []


class _Leaf:
    """Common dispatch for 1-D leaves (``src/zfit/core/basepdf.py:398-461``).

    ``pdf(x, norm)``: unnormalised value divided by the integral over ``norm`` (the
    ``_fallback_pdf`` / ``NormNotImplemented`` routes both end in value / normalization,
    ``basepdf.py:438-461``; ``normalization`` = ``integrate(limits=norm, norm=False)``,
    ``basepdf.py:198-251``).  ``norm=False`` gives the unnormalised value.
    """

    def __init__(self, obs, limits, yield_=None):
        self.obs = obs
        self.limits = (float(limits[0]), float(limits[1]))  # the pdf's own space == default norm
        self.yield_ = yield_

    # -- to be provided by subclasses
    def unnorm(self, B, x, P, norm):
        raise NotImplementedError

    def integral(self, B, P, lo, hi, norm):
        raise NotImplementedError

    def param_names(self):
        raise NotImplementedError

    # -- reference dispatch
    def pdf(self, B, x, P, norm=None):
        norm = self.limits if norm is None else norm
        value = self.unnorm(B, x[self.obs], P, norm)
        if norm is False:
            return value
        return value / self.integral(B, P, norm[0], norm[1], norm)

    def integrate(self, B, P, limits, norm=False):
        """``pdf.integrate(limits, norm=False)`` as used by ProductPDF (``functor.py:423-433``)."""
        assert norm is False
        return self.integral(B, P, limits[0], limits[1], norm)

    def all_param_names(self, with_yield=True):
        out = []
        if with_yield and self.yield_ is not None:
            out += _names(self.yield_)
        out += self.param_names()
        return out


class Gauss(_Leaf):
    """``zfit.pdf.Gauss`` = ``WrapDistribution`` over ``tfp.distributions.Normal``
    (``src/zfit/models/dist_tfp.py:68-123,187-268``).

    Third-party arithmetic (tensorflow-probability >=0.24,<1.0; not vendored) restated from
    its published source: ``Normal._log_prob``: ``-0.5*squared_difference(x/scale, loc/scale)
    - (0.5*log(2*pi) + log(scale))``; ``prob = exp(log_prob)``; ``cdf = special_math.ndtr((x-loc)/scale)``.
    """

    def __init__(self, mu, sigma, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.mu, self.sigma = mu, sigma

    def param_names(self):
        return _names(self.mu) + _names(self.sigma)

    def unnorm(self, B, x, P, norm):  # dist_tfp.py:108-110
        mu, sigma = _val(B, P, self.mu), _val(B, P, self.sigma)
        d = x / sigma - mu / sigma
        log_unnormalized = B.const(-0.5) * (d * d)
        log_normalization = B.const(0.5 * math.log(2.0 * math.pi)) + B.log(sigma)
        return B.exp(log_unnormalized - log_normalization)

    @staticmethod
    def ndtr(B, x):
        """``tfp.math.special.ndtr`` (``tensorflow_probability/python/internal/special_math.py``)."""
        half_sqrt_2 = B.const(0.5 * SQRT2)
        w = x * half_sqrt_2
        z = B.abs(w)
        y = B.where(
            z < half_sqrt_2,
            B.const(1.0) + B.erf(w),
            B.where(w > 0, B.const(2.0) - B.erfc(z), B.erfc(z)),
        )
        return B.const(0.5) * y

    def integral(self, B, P, lo, hi, norm):  # dist_tfp.py:113-120
        mu, sigma = _val(B, P, self.mu), _val(B, P, self.sigma)
        up = self.ndtr(B, (B.const(hi) - mu) / sigma)
        low = self.ndtr(B, (B.const(lo) - mu) / sigma)
        return up - low


class Exponential(_Leaf):
    """``zfit.pdf.Exponential`` (``src/zfit/models/basic.py:34-243``).

    The data shift is the centre of the *norm handed to the hook* (``basic.py:131-154,198-200``)
    and 0 when that norm has no limits (``basic.py:98``) -- i.e. inside ``ProductPDF``.
    """

    def __init__(self, lam, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.lam = lam

    def param_names(self):
        return _names(self.lam)

    @staticmethod
    def _shift(B, norm):
        if norm is False or norm is None:
            return B.const(0.0)
        return (B.const(norm[1]) + B.const(norm[0])) / B.const(2.0)  # basic.py:149

    def unnorm(self, B, x, P, norm):  # basic.py:111-126
        lam = _val(B, P, self.lam)
        return B.exp(lam * (x - self._shift(B, norm)))

    def integral(self, B, P, lo, hi, norm):  # basic.py:211-229
        lam = _val(B, P, self.lam)
        s = self._shift(B, norm)
        upper_int = B.exp(lam * (B.const(hi) - s)) / lam
        lower_int = B.exp(lam * (B.const(lo) - s)) / lam
        return upper_int - lower_int


class CrystalBall(_Leaf):
    """``zfit.pdf.CrystalBall`` (``src/zfit/models/physics.py:31-45,83-142,233-336``)."""

    def __init__(self, mu, sigma, alpha, n, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.mu, self.sigma, self.alpha, self.n = mu, sigma, alpha, n

    def param_names(self):
        return _names(self.mu) + _names(self.sigma) + _names(self.alpha) + _names(self.n)

    def unnorm(self, B, x, P, norm):  # crystalball_func, physics.py:31-45
        mu, sigma = _val(B, P, self.mu), _val(B, P, self.sigma)
        alpha, n = _val(B, P, self.alpha), _val(B, P, self.n)
        t = (x - mu) / sigma * B.sign(alpha)
        abs_alpha = B.abs(alpha)
        a = B.power(n / abs_alpha, n) * B.exp(B.const(-0.5) * B.square(alpha))
        b = n / abs_alpha - abs_alpha
        cond = t < -abs_alpha
        # z.safe_where (zextension.py:149-174): the dead branch sees the safe value b-2
        t_tail = B.where(cond, t, B.ones_like(t) * (b - B.const(2.0)))
        t_core = B.where(cond, B.ones_like(t) * (b - B.const(2.0)), t)
        tail = a * B.power(b - t_tail, -n)  # _powerlaw, physics.py:27-28
        core = B.exp(B.const(-0.5) * B.square(t_core))
        func = B.where(cond, tail, core)
        return B.maximum(func, B.zeros_like(func))

    def integral(self, B, P, lo, hi, norm):  # crystalball_integral_func, physics.py:83-142
        mu, sigma = _val(B, P, self.mu), _val(B, P, self.sigma)
        alpha, n = _val(B, P, self.alpha), _val(B, P, self.n)
        one = B.const(1.0)
        sqrt_pi_over_two = B.const(math.sqrt(math.pi / 2))
        sqrt2 = B.const(SQRT2)
        use_log = B.abs(n - one) < 1e-05
        abs_sigma = B.abs(sigma)
        abs_alpha = B.abs(alpha)
        tmin = ((lo if hasattr(lo, "dtype") else B.const(lo)) - mu) / abs_sigma
        tmax = ((hi if hasattr(hi, "dtype") else B.const(hi)) - mu) / abs_sigma
        alpha_negative = alpha < 0
        tmax, tmin = B.where(alpha_negative, -tmin, tmax), B.where(alpha_negative, -tmax, tmin)

        if_true_4 = abs_sigma * sqrt_pi_over_two * (B.erf(tmax / sqrt2) - B.erf(tmin / sqrt2))
        a = B.power(n / abs_alpha, n) * B.exp(B.const(-0.5) * B.square(abs_alpha))
        b = n / abs_alpha - abs_alpha
        b_tmin = b - tmin
        safe_b_tmin_ones = B.where(b_tmin > 0, b_tmin, B.ones_like(b_tmin))
        b_tmax = b - tmax
        safe_b_tmax_ones = B.where(b_tmax > 0, b_tmax, B.ones_like(b_tmax))
        if_true_1 = a * abs_sigma * (B.log(safe_b_tmin_ones) - B.log(safe_b_tmax_ones))
        if_false_1 = (
            a
            * abs_sigma
            / (one - n)
            * (one / B.power(safe_b_tmin_ones, n - one) - one / B.power(safe_b_tmax_ones, n - one))
        )
        if_true_3 = B.where(use_log, if_true_1, if_false_1)
        if_true_2 = a * abs_sigma * (B.log(safe_b_tmin_ones) - B.log(n / abs_alpha))
        if_false_2 = (
            a
            * abs_sigma
            / (one - n)
            * (one / B.power(safe_b_tmin_ones, n - one) - one / B.power(n / abs_alpha, n - one))
        )
        term1 = B.where(use_log, if_true_2, if_false_2)
        term2 = abs_sigma * sqrt_pi_over_two * (B.erf(tmax / sqrt2) - B.erf(-abs_alpha / sqrt2))
        if_false_3 = term1 + term2
        if_false_4 = B.where(tmax <= -abs_alpha, if_true_3, if_false_3)
        return B.where(tmin >= -abs_alpha, if_true_4, if_false_4)


class Chebyshev(_Leaf):
    """``zfit.pdf.Chebyshev`` (``src/zfit/models/polynomials.py:32-43,46-144,164-175,334-482``).

    ``coeffs`` are c_1..c_d; c_0 is the constant 1 unless ``coeff0`` is given (``polynomials.py:97-100``).
    Rescaling uses the pdf's *space* limits (``polynomials.py:112-114``), also inside the integral.
    """

    def __init__(self, coeffs, obs, limits, coeff0=None, yield_=None):
        super().__init__(obs, limits, yield_)
        self.coeffs = [1.0 if coeff0 is None else coeff0, *coeffs]
        self.degree = len(self.coeffs) - 1

    def param_names(self):
        out = []
        for c in self.coeffs:
            out += _names(c)
        return out

    def _rescale(self, B, x):  # rescale_minus_plus_one, polynomials.py:32-43
        lim_low, lim_high = B.const(self.limits[0]), B.const(self.limits[1])
        return (B.const(2.0) * x - lim_low - lim_high) / (lim_high - lim_low)

    @staticmethod
    def _recurrence(B, x, degree):  # do_recurrence + chebyshev_recurrence, polynomials.py:171-175,338-343
        polys = [B.ones_like(x), x]
        for _ in range(1, degree):
            polys.append(B.const(2.0) * (x * polys[-1]) - polys[-2])
        return polys

    def unnorm(self, B, x, P, norm):  # _pdf + create_poly, polynomials.py:134-139,164-168
        z = self._rescale(B, x)
        polys = self._recurrence(B, z, self.degree)
        coeffs = [_val(B, P, c) for c in self.coeffs]
        total = None
        for c, p in zip(coeffs, polys):  # znp.sum over the stacked list == sequential adds
            term = c * p
            total = term if total is None else total + term
        return total

    def integral(self, B, P, lo, hi, norm):  # func_integral_chebyshev1, polynomials.py:443-478
        coeffs = [_val(B, P, c) for c in self.coeffs]
        lower = self._rescale(B, B.const(lo))
        upper = self._rescale(B, B.const(hi))
        integral = coeffs[0] * (upper - lower)
        if self.degree >= 1:
            integral = integral + coeffs[1] * B.const(0.5) * (upper * upper - lower * lower)
        if self.degree >= 2:

            def indefinite(lim):
                max_degree = self.degree + 1
                polys = self._recurrence(B, lim, max_degree)
                tot = None
                for degree in range(2, max_degree):
                    nf = B.const(float(degree))
                    one_int = nf * polys[degree + 1] / (B.square(nf) - B.const(1.0)) - lim * polys[degree] / (
                        nf - B.const(1.0)
                    )
                    term = coeffs[degree] * one_int
                    tot = term if tot is None else tot + term
                return tot

            integral = integral + (indefinite(upper) - indefinite(lower))
        volume = B.const(self.limits[1] - self.limits[0])
        return integral * (B.const(0.5) * volume)


class Legendre(Chebyshev):
    """``zfit.pdf.Legendre`` (``src/zfit/models/polynomials.py:181-332``)."""

    @staticmethod
    def _recurrence(B, x, degree):  # legendre_recurrence, polynomials.py:181-190
        polys = [B.ones_like(x), x]
        for n in range(1, degree):
            polys.append((B.const(2.0 * n + 1.0) * (x * polys[-1]) - B.const(float(n)) * polys[-2]) / B.const(n + 1.0))
        return polys

    def integral(self, B, P, lo, hi, norm):  # legendre_integral, polynomials.py:193-228
        coeffs = [_val(B, P, c) for c in self.coeffs]
        lower = self._rescale(B, B.const(lo))
        upper = self._rescale(B, B.const(hi))
        integral = coeffs[0] * (upper - lower)
        if self.degree >= 1:

            def indefinite(lim):
                polys = self._recurrence(B, lim, self.degree + 1)
                tot = None
                for degree in range(1, self.degree + 1):
                    term = coeffs[degree] * (polys[degree + 1] - polys[degree - 1]) / B.const(2.0 * degree + 1.0)
                    tot = term if tot is None else tot + term
                return tot

            integral = indefinite(upper) - indefinite(lower) + integral
        return integral * (B.const(0.5) * B.const(self.limits[1] - self.limits[0]))


class Chebyshev2(Chebyshev):
    """``zfit.pdf.Chebyshev2`` (``src/zfit/models/polynomials.py:487-600``): U_0 = 1, U_1 = 2x, same recurrence."""

    @staticmethod
    def _recurrence(B, x, degree):
        polys = [B.ones_like(x), x * B.const(2.0)]
        for _ in range(1, degree):
            polys.append(B.const(2.0) * (x * polys[-1]) - polys[-2])
        return polys

    def integral(self, B, P, lo, hi, norm):  # func_integral_chebyshev2, polynomials.py:565-590
        coeffs = [_val(B, P, c) for c in self.coeffs]
        lower = self._rescale(B, B.const(lo))
        upper = self._rescale(B, B.const(hi))

        def indefinite(lim):  # integral of U_n is T_{n+1}/(n+1): a first-kind sum with c_0 = 0
            polys = Chebyshev._recurrence(B, lim, self.degree + 1)
            tot = None
            for n, c in enumerate(coeffs):
                term = (c / B.const(n + 1.0)) * polys[n + 1]
                tot = term if tot is None else tot + term
            return tot

        integral = indefinite(upper) - indefinite(lower)
        return integral * (B.const(0.5) * B.const(self.limits[1] - self.limits[0]))


class Hermite(Chebyshev):
    """``zfit.pdf.Hermite`` (``src/zfit/models/polynomials.py:752-869``): physicists' H_0 = 1, H_1 = 2x,
    H_{n+1} = 2 (x H_n - n H_{n-1})."""

    @staticmethod
    def _recurrence(B, x, degree):  # hermite_polys + hermite_recurrence, polynomials.py:752-760
        polys = [B.ones_like(x), x * B.const(2.0)]
        for n in range(1, degree):
            polys.append(B.const(2.0) * (x * polys[-1] - B.const(float(n)) * polys[-2]))
        return polys

    def integral(self, B, P, lo, hi, norm):  # func_integral_hermite, polynomials.py:841-866
        coeffs = [_val(B, P, c) for c in self.coeffs]
        lower = self._rescale(B, B.const(lo))
        upper = self._rescale(B, B.const(hi))

        def indefinite(lim):  # the integral of H_n is H_{n+1} / (2 (n+1)): a Hermite sum with shifted coefficients, c_0 = 0
            polys = self._recurrence(B, lim, self.degree + 1)
            tot = None
            for n, c in enumerate(coeffs):
                term = (c / B.const((n + 1) * 2.0)) * polys[n + 1]
                tot = term if tot is None else tot + term
            return tot

        integral = indefinite(upper) - indefinite(lower)
        return integral * (B.const(0.5) * B.const(self.limits[1] - self.limits[0]))


class Laguerre(Chebyshev):
    """``zfit.pdf.Laguerre`` (``src/zfit/models/polynomials.py:600-750``): L_0 = 1, L_1 = 1 - x,
    (n+1) L_{n+1} = (2n + 1 - x) L_n - n L_{n-1}."""

    @staticmethod
    def _general(B, x, degree, alpha):  # generalized_laguerre_polys_factory / _recurrence_factory, polynomials.py:600-619
        polys = [B.ones_like(x), B.const(1.0 + alpha) - x]
        for n in range(1, degree):
            polys.append(((B.const(2.0 * n + 1.0 + alpha) - x) * polys[-1] - B.const(n + alpha) * polys[-2]) / B.const(n + 1.0))
        return polys

    @classmethod
    def _recurrence(cls, B, x, degree):
        return cls._general(B, x, degree, 0.0)

    def integral(self, B, P, lo, hi, norm):  # func_integral_laguerre, polynomials.py:711-746
        coeffs = [_val(B, P, c) for c in self.coeffs]
        lower = self._rescale(B, B.const(lo))
        upper = self._rescale(B, B.const(hi))

        def indefinite(lim):  # the integral of L_n is -L^(-1)_{n+1}: a generalised (alpha = -1) sum with c_0 = 0
            polys = self._general(B, lim, self.degree + 1, -1.0)
            tot = None
            for n, c in enumerate(coeffs):
                term = c * polys[n + 1]
                tot = term if tot is None else tot + term
            return B.const(-1.0) * tot

        integral = indefinite(upper) - indefinite(lower)
        return integral * (B.const(0.5) * B.const(self.limits[1] - self.limits[0]))


class Bernstein(_Leaf):
    """``zfit.pdf.Bernstein`` (``src/zfit/models/polynomials.py:872-1028``): De Casteljau's algorithm on the observable
    rescaled to (0, 1) with the pdf's *space* limits; every coefficient is a parameter."""

    def __init__(self, coeffs, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.coeffs = list(coeffs)
        self.degree = len(self.coeffs) - 1

    def param_names(self):
        out = []
        for c in self.coeffs:
            out += _names(c)
        return out

    def _rescale(self, B, x):  # rescale_zero_one, polynomials.py:872-885
        lim_low, lim_high = B.const(self.limits[0]), B.const(self.limits[1])
        return (x - lim_low) / (lim_high - lim_low)

    @staticmethod
    def _de_casteljau(B, x, coeffs):  # polynomials.py:888-900
        beta = list(coeffs)
        n = len(beta)
        for j in range(1, n):
            for k in range(n - j):
                beta[k] = beta[k] * (B.const(1.0) - x) + beta[k + 1] * x
        if n == 1:
            beta[0] = beta[0] * B.ones_like(x)
        return beta[0]

    def unnorm(self, B, x, P, norm):  # _pdf, polynomials.py:969-978
        return self._de_casteljau(B, self._rescale(B, x), [_val(B, P, c) for c in self.coeffs])

    def integral(self, B, P, lo, hi, norm):  # func_integral_bernstein + _coeffs_int, polynomials.py:1000-1025
        coeffs = [_val(B, P, c) for c in self.coeffs]
        n = len(coeffs)
        r = [B.const(0.0)] * (n + 1)
        for j in range(1, n + 1):
            for k in range(j):
                r[j] = r[j] + coeffs[k]
        r = [rj / B.const(float(n)) for rj in r]
        volume = B.const(self.limits[1]) - B.const(self.limits[0])
        upper = self._de_casteljau(B, self._rescale(B, B.const(hi)), r) * volume
        lower = self._de_casteljau(B, self._rescale(B, B.const(lo)), r) * volume
        return upper - lower


class TruncatedGauss(_Leaf):
    """``zfit.pdf.TruncatedGauss`` = ``WrapDistribution`` over ``tfp.distributions.TruncatedNormal``
    (``src/zfit/models/dist_tfp.py:357-417``; value ``distribution.prob``, :108-110; integral ``cdf(upper) - cdf(lower)``,
    :113-120).

    Third-party arithmetic (tensorflow-probability >=0.24,<1.0; not vendored) restated from its published source
    (``truncated_normal.py``): ``log_prob = -(0.5*((x-loc)/scale)**2 + 0.5*log(2*pi) + log(scale) + log Z)`` with
    ``Z = ndtr(std_high) - ndtr(std_low)``, ``-inf`` outside ``[low, high]``; ``cdf = clip((ndtr((x-loc)/scale) -
    ndtr(std_low)) / Z, 0, 1)``.  (Newer releases form ``log Z`` from ``log_ndtr`` differences; same value.)
    """

    def __init__(self, mu, sigma, low, high, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.mu, self.sigma, self.low, self.high = mu, sigma, low, high

    def param_names(self):
        return _names(self.mu) + _names(self.sigma) + _names(self.low) + _names(self.high)

    def _parts(self, B, P):
        mu, sigma = _val(B, P, self.mu), _val(B, P, self.sigma)
        low, high = _val(B, P, self.low), _val(B, P, self.high)
        z = Gauss.ndtr(B, (high - mu) / sigma) - Gauss.ndtr(B, (low - mu) / sigma)
        return mu, sigma, low, high, z

    def unnorm(self, B, x, P, norm):
        mu, sigma, low, high, z = self._parts(B, P)
        t = (x - mu) / sigma
        log_prob = -(B.const(0.5) * B.square(t) + B.const(0.5 * math.log(2.0 * math.pi)) + B.log(sigma) + B.log(z))
        prob = B.exp(log_prob)
        return B.where((x > high) | (x < low), B.zeros_like(prob), prob)

    def integral(self, B, P, lo, hi, norm):
        mu, sigma, low, high, z = self._parts(B, P)

        def cdf(v):
            c = (Gauss.ndtr(B, (B.const(v) - mu) / sigma) - Gauss.ndtr(B, (low - mu) / sigma)) / z
            one, zero = B.const(1.0), B.const(0.0)
            return B.where(c < zero, zero * c, B.where(c > one, one + zero * c, c))

        return cdf(hi) - cdf(lo)


class GeneralizedCB(_Leaf):
    """``zfit.pdf.GeneralizedCB`` (``src/zfit/models/physics.py:47-80,186-232``); ``DoubleCB`` = ``sigmal is sigmar``."""

    def __init__(self, mu, sigmal, alphal, nl, sigmar, alphar, nr, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.p = (mu, sigmal, alphal, nl, sigmar, alphar, nr)

    def param_names(self):
        out = []
        for q in self.p:
            for n in _names(q):
                if n not in out:
                    out.append(n)
        return out

    def _sides(self, B, P):
        mu, sl, al, nl, sr, ar, nr = (_val(B, P, q) for q in self.p)
        left = CrystalBall(lambda _: mu, lambda _: sl, lambda _: al, lambda _: nl, self.obs, self.limits)
        right = CrystalBall(lambda _: mu, lambda _: sr, lambda _: -ar, lambda _: nr, self.obs, self.limits)
        return mu, left, right

    def unnorm(self, B, x, P, norm):  # generalized_crystalball_func, physics.py:69-80
        mu, left, right = self._sides(B, P)
        return B.where(x < mu, left.unnorm(B, x, P, norm), right.unnorm(B, x, P, norm))

    def integral(self, B, P, lo, hi, norm):  # generalized_crystalball_mu_integral_func, physics.py:218-232
        mu, left, right = self._sides(B, P)
        lo_c, hi_c = B.const(lo), B.const(hi)
        upper_of_lowerint = B.where(mu < hi_c, mu, hi_c)
        lower_of_upperint = B.where(mu > lo_c, mu, lo_c)
        il = left.integral(B, P, lo_c, upper_of_lowerint, norm)
        ir = right.integral(B, P, lower_of_upperint, hi_c, norm)
        il = B.where(mu < lo_c, B.zeros_like(il), il)
        ir = B.where(mu > hi_c, B.zeros_like(ir), ir)
        return il + ir


class GaussExpTail(_Leaf):
    """``zfit.pdf.GaussExpTail`` (``src/zfit/models/physics.py:608-623,636-686``)."""

    def __init__(self, mu, sigma, alpha, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.mu, self.sigma, self.alpha = mu, sigma, alpha

    def param_names(self):
        return _names(self.mu) + _names(self.sigma) + _names(self.alpha)

    def unnorm(self, B, x, P, norm):  # gaussexptail_func, physics.py:610-623
        mu, sigma, alpha = _val(B, P, self.mu), _val(B, P, self.sigma), _val(B, P, self.alpha)
        t = (x - mu) / sigma * B.sign(alpha)
        abs_alpha = B.abs(alpha)
        cond = t < -abs_alpha
        t_tail = B.where(cond, t, B.ones_like(t) * abs_alpha)
        t_core = B.where(cond, B.ones_like(t) * abs_alpha, t)
        func = B.where(cond, B.const(0.5) * B.square(abs_alpha) + abs_alpha * t_tail, B.const(-0.5) * B.square(t_core))
        return B.exp(func)

    def integral(self, B, P, lo, hi, norm):  # gaussexptail_integral_func, physics.py:647-686
        mu, sigma, alpha = _val(B, P, self.mu), _val(B, P, self.sigma), _val(B, P, self.alpha)
        sqrt_pi_over_two = B.const(math.sqrt(math.pi / 2))
        sqrt2 = B.const(SQRT2)
        abs_sigma, abs_alpha = B.abs(sigma), B.abs(alpha)
        tmin = ((lo if hasattr(lo, "dtype") else B.const(lo)) - mu) / abs_sigma
        tmax = ((hi if hasattr(hi, "dtype") else B.const(hi)) - mu) / abs_sigma
        alpha_negative = alpha < 0
        tmax, tmin = B.where(alpha_negative, -tmin, tmax), B.where(alpha_negative, -tmax, tmin)
        gauss_tmin_tmax = abs_sigma * sqrt_pi_over_two * (B.erf(tmax / sqrt2) - B.erf(tmin / sqrt2))
        pref = abs_sigma / abs_alpha * B.exp(B.const(0.5) * B.square(abs_alpha))
        exp_tmin_tmax = pref * (B.exp(abs_alpha * tmax) - B.exp(abs_alpha * tmin))
        gauss_minus_a_tmax = abs_sigma * sqrt_pi_over_two * (B.erf(tmax / sqrt2) - B.erf(-abs_alpha / sqrt2))
        exp_tmin_minus_a = pref * (B.exp(-B.square(abs_alpha)) - B.exp(abs_alpha * tmin))
        integral_sum = exp_tmin_minus_a + gauss_minus_a_tmax
        conditional = B.where(tmax <= -abs_alpha, exp_tmin_tmax, integral_sum)
        return B.where(tmin >= -abs_alpha, gauss_tmin_tmax, conditional)


class GeneralizedGaussExpTail(_Leaf):
    """``zfit.pdf.GeneralizedGaussExpTail`` (``src/zfit/models/physics.py:625-634,712-724``)."""

    def __init__(self, mu, sigmal, alphal, sigmar, alphar, obs, limits, yield_=None):
        super().__init__(obs, limits, yield_)
        self.p = (mu, sigmal, alphal, sigmar, alphar)

    def param_names(self):
        out = []
        for q in self.p:
            for n in _names(q):
                if n not in out:
                    out.append(n)
        return out

    def _sides(self, B, P):
        mu, sl, al, sr, ar = (_val(B, P, q) for q in self.p)
        left = GaussExpTail(lambda _: mu, lambda _: sl, lambda _: al, self.obs, self.limits)
        right = GaussExpTail(lambda _: mu, lambda _: sr, lambda _: -ar, self.obs, self.limits)
        return mu, left, right

    def unnorm(self, B, x, P, norm):
        mu, left, right = self._sides(B, P)
        return B.where(x < mu, left.unnorm(B, x, P, norm), right.unnorm(B, x, P, norm))

    def integral(self, B, P, lo, hi, norm):
        mu, left, right = self._sides(B, P)
        lo_c, hi_c = B.const(lo), B.const(hi)
        il = left.integral(B, P, lo_c, B.where(mu < hi_c, mu, hi_c), norm)
        ir = right.integral(B, P, B.where(mu > lo_c, mu, lo_c), hi_c, norm)
        il = B.where(mu < lo_c, B.zeros_like(il), il)
        ir = B.where(mu > hi_c, B.zeros_like(ir), ir)
        return il + ir


class SumPDF:
    """``zfit.pdf.SumPDF`` (``src/zfit/models/functor.py:74-313``, ``basefunctor.py:160-255``).

    ``fracs``: list of k-1 (remainder = 1 - sum, ``basefunctor.py:189-205``) or k parameter refs;
    ``None`` with all daughters extended -> ``frac_i = yield_i / sum(yields)`` and the sum is extended
    with ``yield = sum(yields)`` (``basefunctor.py:228-244``).
    """

    def __init__(self, pdfs, fracs=None, yield_=None):
        self.pdfs = list(pdfs)
        self.obs = self.pdfs[0].obs
        self.limits = self.pdfs[0].limits
        self.fracs_in = None if not fracs else list(fracs)
        self.from_yields = self.fracs_in is None
        if self.from_yields:
            assert all(p.yield_ is not None for p in self.pdfs)
            self.yield_ = lambda P: sum(P[y] if isinstance(y, str) else y for y in [p.yield_ for p in self.pdfs])
            self._yield_names = [n for p in self.pdfs for n in _names(p.yield_)]
        else:
            self.yield_ = yield_
            self._yield_names = _names(yield_) if yield_ is not None else []

    def fracs(self, B, P):
        if self.from_yields:
            ys = [_val(B, P, p.yield_) for p in self.pdfs]
            tot = None
            for y in ys:
                tot = y if tot is None else tot + y
            return [y / tot for y in ys]
        fr = [_val(B, P, f) for f in self.fracs_in]
        if len(fr) == len(self.pdfs) - 1:
            s = None
            for f in fr:
                s = f if s is None else s + f
            fr.append(B.const(1.0) - s)
        return fr

    def get_yield(self, B, P):
        if self.from_yields:
            ys = [_val(B, P, p.yield_) for p in self.pdfs]
            tot = None
            for y in ys:
                tot = y if tot is None else tot + y
            return tot
        return _val(B, P, self.yield_)

    def pdf(self, B, x, P, norm=None):  # functor.py:197-206 (python sum(): 0 + p0 + p1 ...)
        fr = self.fracs(B, P)
        prob = None
        for pdf, f in zip(self.pdfs, fr):
            term = pdf.pdf(B, x, P) * f
            prob = term if prob is None else prob + term
        return prob

    def all_param_names(self, with_yield=True):
        out = []
        if with_yield:
            out += self._yield_names
        if not self.from_yields:
            for f in self.fracs_in:
                out += _names(f)
        for p in self.pdfs:
            for n in p.all_param_names(with_yield=False):
                if n not in out:
                    out.append(n)
        return out


class ProductPDF:
    """``zfit.pdf.ProductPDF`` over disjoint observables (``src/zfit/models/functor.py:331-483``):
    ``_pdf`` raises (1-D daughter norms != N-D norm, :413-421) -> ``_fallback_pdf`` =
    ``prod_i pdf_i.pdf(x, norm=False)`` (:408-411) / ``prod_i pdf_i.integrate(limits_i, norm=False)`` (:423-433).
    """

    def __init__(self, pdfs, yield_=None):
        self.pdfs = list(pdfs)
        self.yield_ = yield_

    limits = None  # a product integrates each daughter over the daughter's own limits

    def pdf(self, B, x, P, norm=None):
        prob = None
        for pdf in self.pdfs:
            v = pdf.pdf(B, x, P, norm=False)
            prob = v if prob is None else prob * v
        if norm is False:  # as a daughter of another product: ``_unnormalized_pdf`` (functor.py:408-411)
            return prob
        return prob / self.integrate(B, P, None, norm=False)

    def integrate(self, B, P, limits, norm=False):  # functor.py:423-433: product of the daughters' integrals
        integral = None
        for pdf in self.pdfs:
            i = pdf.integrate(B, P, pdf.limits, norm=False)
            integral = i if integral is None else integral * i
        return integral

    def get_yield(self, B, P):
        return _val(B, P, self.yield_)

    def all_param_names(self, with_yield=True):
        out = []
        if with_yield and self.yield_ is not None:
            out += _names(self.yield_)
        for p in self.pdfs:
            for n in p.all_param_names(with_yield=False):
                if n not in out:
                    out.append(n)
        return out


def leaf_get_yield(pdf, B, P):
    return _val(B, P, pdf.yield_)


for _cls in (Gauss, Exponential, CrystalBall, Chebyshev, Legendre, Chebyshev2, Hermite, Laguerre, Bernstein, TruncatedGauss, GeneralizedCB,
             GaussExpTail, GeneralizedGaussExpTail):
    _cls.get_yield = lambda self, B, P: _val(B, P, self.yield_)


def inside_limits(x, lo, hi):
    """Inclusive cut on both ends (``src/zfit/core/space.py:218-231``)."""
    x = np.asarray(x)
    return (x >= lo) & (x <= hi)
