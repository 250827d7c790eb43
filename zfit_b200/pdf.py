"""Model layer of the hot path: ``zfit.pdf.{Gauss, Exponential, CrystalBall, Chebyshev, SumPDF, ProductPDF}``.

A PDF object here is *descriptive*: it knows its observables, its normalisation range, its
parameters and how to lower itself into the pre-order ``znll_node`` array + slot list that the
fused CUDA kernel consumes (``include/znll.h``).  The per-event arithmetic lives in
``csrc/znll_device.cuh``; the analytic normalisation integrals in ``csrc/znll_host.cu``.  What this
module mirrors from the reference is the *dispatch semantics* of ``BasePDF.pdf``
(``src/zfit/core/basepdf.py:398-461``): which range each node is normalised over, the Exponential's
data shift (``models/basic.py:128-208``), fractions from yields (``models/basefunctor.py:160-255``),
and the parameter ordering of ``get_params`` (``core/baseobject.py:213-278``, ``core/basepdf.py:829-863``,
``models/basefunctor.py:82-100``).
"""

from __future__ import annotations

import numpy as np

from . import _znll as Z
from .data import Data, convert_to_data
from .exception import (
    AlreadyExtendedPDFError,
    ModelIncompatibleError,
    ModelNotOnDeviceError,
    NotExtendedPDFError,
    ObsIncompatibleError,
    ParamNameNotUniqueError,
)
from .parameter import (
    ComposedParameter,
    FractionOfSum,
    OneMinusSum,
    OrderedSet,
    SumOfParameters,
    ZfitParameter,
    convert_to_parameter,
)
from .space import Space, combine_spaces, convert_to_obs_str


def _filter(params, floating, extract_independent=True):
    out = OrderedSet()
    for p in params:
        leaves = p._leaves() if extract_independent is not False else [p]
        if extract_independent is False and p.independent:
            continue
        for leaf in leaves:
            if floating is None or leaf.floating == floating:
                out.add(leaf)
    return out


class BasePDF:
    """Common PDF surface (``core/basepdf.py``, ``core/basemodel.py``)."""

    _N_OBS = None

    def __init__(self, obs, params: dict, *, extended=None, norm=None, name=None, label=None):
        self._space = obs if isinstance(obs, Space) else Space(obs)
        if self._N_OBS is not None and self._space.n_obs != self._N_OBS:
            msg = f"{type(self).__name__} is {self._N_OBS}-dimensional, got obs {self._space.obs}"
            raise ObsIncompatibleError(msg)
        self._norm = None
        if norm is not None and norm is not False:
            self._norm = norm if isinstance(norm, Space) else self._space.with_limits(norm)
        self._params = dict(params)
        self.name = name or type(self).__name__
        self.label = label or self.name
        self._yield = None
        if extended is not None and extended is not False and extended is not True:
            self._set_yield(extended)
        self._check_unique_names()

    # ---------------------------------------------------------------- description
    @property
    def space(self) -> Space:
        return self._space

    @property
    def obs(self):
        return self._space.obs

    @property
    def n_obs(self):
        return self._space.n_obs

    @property
    def norm(self) -> Space:
        """Normalisation range: by default the space of the PDF (``basepdf.py:143-146``)."""
        return self._norm if self._norm is not None else self._space

    norm_range = norm

    @property
    def params(self) -> dict:
        return self._params

    @property
    def is_extended(self) -> bool:
        return self._yield is not None

    def get_yield(self):
        return self._yield

    def _set_yield(self, value):
        if self.is_extended:
            msg = f"Cannot extend {self}, is already extended."
            raise AlreadyExtendedPDFError(msg)
        self._yield = convert_to_parameter(value)

    def create_extended(self, yield_, name=None):
        """Copy sharing the parameters, with a yield (``basepdf.py:639-689``)."""
        if self.is_extended:
            msg = "This PDF is already extended, cannot create an extended one."
            raise AlreadyExtendedPDFError(msg)
        new = self._copy()
        new._yield = convert_to_parameter(yield_)
        if name:
            new.name = name
        new._check_unique_names()
        return new

    def _copy(self):
        import copy

        new = copy.copy(self)
        new._params = dict(self._params)
        return new

    def _check_unique_names(self):
        seen = {}
        for p in self.get_params(floating=None, is_yield=None, extract_independent=True):
            if p.name in seen and seen[p.name] is not p:
                msg = f"The following parameter names appear more than once in {self}: {p.name}."
                raise ParamNameNotUniqueError(msg)
            seen[p.name] = p

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        if is_yield is True:
            if not self.is_extended:
                msg = "PDF is not extended but only yield parameters were requested."
                raise NotExtendedPDFError(msg)
            return _filter([self._yield], floating, extract_independent)
        own = _filter(self._params.values(), floating, extract_independent)
        if is_yield is not False and self.is_extended:
            out = _filter([self._yield], floating, extract_independent)
            return out.union(own)
        return own

    # ---------------------------------------------------------------- lowering
    def _emit(self, colmap, ctx, limits, nodes, slots):
        raise NotImplementedError

    def _lower(self, obs_order=None, norm=None):
        """-> (nodes, slot parameter objects); ``norm``: explicit normalisation Space or None."""
        obs_order = self.obs if obs_order is None else tuple(obs_order)
        colmap = {o: i for i, o in enumerate(obs_order)}
        missing = [o for o in self.obs if o not in colmap]
        if missing:
            msg = f"data lacks the observables {missing} of {self}"
            raise ObsIncompatibleError(msg)
        nodes, slots = [], []
        self._emit(colmap, "norm", norm, nodes, slots)
        return nodes, slots

    # ---------------------------------------------------------------- evaluation (device)
    def _convert_sort_x(self, x):
        if not isinstance(x, Data):
            x = convert_to_data(x, obs=Space(self.obs)) if not hasattr(x, "keys") else Data(x, obs=Space(self.obs))
        return x.with_obs(self.obs)

    def pdf(self, x, norm=None, *, params=None):
        """Normalised probability density at ``x`` (``basepdf.py:398-429``), evaluated by the per-event CUDA
        kernel; returns a numpy array of shape ``(n_events,)``."""
        x = self._convert_sort_x(x)
        nspace = self._check_norm(norm)
        with _maybe_set(params, self):
            nodes, slot_params = self._lower(self.obs, nspace)
            slots = [p.value() for p in slot_params]
        n = x.num_entries
        if n == 0:
            return np.zeros(0)
        plan = Z.Plan(1)
        try:
            plan.set_model(0, nodes)
            cols, _ = x.device_columns(plan.device)
            import torch

            torch.cuda.synchronize(plan.device)  # columns come from torch's stream, the plan has its own
            plan.bind_device_ptrs(0, [c.data_ptr() for c in cols], 0, n, keepalive=cols)
            return plan.pdf_values(0, slots, n)
        finally:
            plan.close()

    def ext_pdf(self, x, norm=None, *, params=None):
        if not self.is_extended:
            msg = f"{self} is not extended, cannot call `ext_pdf`"
            raise NotExtendedPDFError(msg)
        with _maybe_set(params, self):
            return self.pdf(x, norm) * self._yield.value()

    def log_pdf(self, x, norm=None, *, params=None):
        return np.log(self.pdf(x, norm, params=params))

    def ext_log_pdf(self, x, norm=None, *, params=None):
        return np.log(self.ext_pdf(x, norm, params=params))

    def _check_norm(self, norm):
        if norm is None:
            return None
        if norm is False:
            msg = "unnormalised pdf values are internal to ProductPDF on this path"
            raise NotImplementedError(msg)
        return norm if isinstance(norm, Space) else self._space.with_limits(norm)

    # ---------------------------------------------------------------- integrals (host, analytic)
    def integrate(self, limits, norm=None, *, params=None):
        """Integral over ``limits`` of the pdf normalised over ``norm`` (``basemodel.py:356-450``)."""
        limits = limits if isinstance(limits, Space) else self._space.with_limits(limits)
        with _maybe_set(params, self):
            raw = self._raw_integral(limits.with_obs(self.obs), shift_norm=(self.norm if norm is not False else False))
            if norm is False:
                return np.atleast_1d(raw)
            nspace = self.norm if norm is None else (norm if isinstance(norm, Space) else self._space.with_limits(norm))
            return np.atleast_1d(raw / self._raw_integral(nspace.with_obs(self.obs), shift_norm=nspace))

    def normalization(self, norm=None, *, params=None):
        nspace = self.norm if norm is None else (norm if isinstance(norm, Space) else self._space.with_limits(norm))
        with _maybe_set(params, self):
            return self._raw_integral(nspace.with_obs(self.obs), shift_norm=nspace)

    def ext_integrate(self, limits, norm=None, *, params=None):
        if not self.is_extended:
            msg = f"{self} is not extended"
            raise NotExtendedPDFError(msg)
        with _maybe_set(params, self):
            return self.integrate(limits, norm) * self._yield.value()

    def _raw_integral(self, limits: Space, shift_norm):
        raise NotImplementedError

    # ---------------------------------------------------------------- sampling (convenience, accept-reject)
    def sample(self, n=None, limits=None, *, params=None):
        """Accept-reject sample from the pdf values of the CUDA kernel.  Convenience for tests and examples;
        the reference's sampler (``core/sample.py``) is out of scope."""
        limits = self._space if limits is None else (limits if isinstance(limits, Space) else self._space.with_limits(limits))
        if n is None:
            if not self.is_extended:
                msg = "n has to be given for a non-extended pdf"
                raise ValueError(msg)
            n = int(np.random.poisson(self._yield.value()))
        n = int(n)
        rng = np.random.default_rng(np.random.randint(0, 2**31 - 1))
        lo, up = limits._lower, limits._upper
        with _maybe_set(params, self):
            probe = rng.uniform(lo, up, size=(20000, limits.n_obs))
            pmax = float(np.max(self.pdf(Data(probe, obs=limits, guarantee_limits=True)))) * 1.3
            out = np.zeros((0, limits.n_obs))
            while out.shape[0] < n:
                cand = rng.uniform(lo, up, size=(max(4 * (n - out.shape[0]), 10000), limits.n_obs))
                p = self.pdf(Data(cand, obs=limits, guarantee_limits=True))
                pmax = max(pmax, float(p.max()))
                out = np.concatenate([out, cand[rng.uniform(0, pmax, size=p.shape) < p]])
        return Data(out[:n], obs=limits, guarantee_limits=True)

    def create_sampler(self, n=None, limits=None, *, fixed_params=True, params=None):
        """``BasePDF.create_sampler`` (``core/basepdf.py``, ``core/data.py:1247-1606``): a ``SamplerData`` drawn from this
        pdf that ``resample()`` re-draws in place -- on the GPU, with the tree's own ``pdf_kernel`` as the density.
        ``fixed_params=True`` freezes the current parameter values for all later resamples (the reference's default)."""
        from .data import SamplerData
        from .sampler import DeviceSampler

        limits = self._space if limits is None else (limits if isinstance(limits, Space) else self._space.with_limits(limits))
        if n is None and not self.is_extended:
            msg = "n has to be given for a non-extended pdf"
            raise ValueError(msg)
        allp = list(self.get_params(floating=None, is_yield=None))
        if fixed_params is True:
            fixed = {p: p.value() for p in allp}
        elif fixed_params is False or fixed_params is None:
            fixed = {}
        else:
            fixed = {p: p.value() for p in fixed_params}
        if params is not None:
            byname = {p.name: p for p in allp}
            for k, v in dict(params).items():
                fixed[byname[k] if isinstance(k, str) else k] = float(v)
        return SamplerData(DeviceSampler(self, limits), n, fixed_params=fixed, obs=limits)

    def __repr__(self):
        return f"<zfit.pdf.{type(self).__name__} name={self.name} obs={self.obs} params={[p.name for p in self._params.values()]}>"


class _maybe_set:
    def __init__(self, params, pdf):
        self._ctx = None
        if params is not None:
            from .parameter import set_values

            if not hasattr(params, "keys"):
                allp = list(pdf.get_params(floating=None, is_yield=None))
                params = dict(zip(allp, params))
            byname = {p.name: p for p in pdf.get_params(floating=None, is_yield=None)}
            pp, vv = [], []
            for k, v in params.items():
                p = k if isinstance(k, ZfitParameter) else byname[k]
                pp.append(p)
                vv.append(v)
            self._ctx = set_values(pp, vv)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


# ------------------------------------------------------------------------------------------------
# leaves
# ------------------------------------------------------------------------------------------------
class _Leaf1D(BasePDF):
    _N_OBS = 1
    _KIND = None

    def _slot_params(self):
        return list(self._params.values())

    def _node(self, column, lo, hi, shift):
        return Z.Node(self._KIND, 0, column, 0, lo, hi, 0.0, 0.0, shift)

    def _shift(self, ctx, norm_space):
        return 0.0

    def _emit(self, colmap, ctx, limits, nodes, slots):
        # ctx "norm": normalised over `limits` (explicit norm) or the pdf's own norm (basepdf.py:438-461)
        # ctx "prod": value with norm=False, integral over the product's limits for this obs (functor.py:408-433)
        rng = (limits if limits is not None else self.norm).with_obs(self.obs)
        if not rng.has_limits:
            msg = f"{self} needs a normalisation range with limits"
            raise ModelIncompatibleError(msg)
        lo, hi = rng.limit1d
        nodes.append(self._node(colmap[self.obs[0]], lo, hi, self._shift(ctx, rng)))
        slots.extend(self._slot_params())

    def _raw_integral(self, limits: Space, shift_norm):
        lo, hi = limits.limit1d
        shift = 0.0
        if shift_norm is not False and isinstance(shift_norm, Space) and shift_norm.has_limits:
            shift = self._shift("norm", shift_norm.with_obs(self.obs))
        node = self._node(0, lo, hi, shift)
        return Z.leaf_integral(node, [p.value() for p in self._slot_params()])[0]


class Gauss(_Leaf1D):
    """``zfit.pdf.Gauss(mu, sigma, obs, *, extended=None, norm=None, name="Gauss", label=None)``
    (``models/dist_tfp.py:187-268``)."""

    _KIND = Z.GAUSS

    def __init__(self, mu, sigma, obs, *, extended=None, norm=None, name="Gauss", label=None):
        params = {"mu": convert_to_parameter(mu), "sigma": convert_to_parameter(sigma)}
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


class TruncatedGauss(_Leaf1D):
    """``zfit.pdf.TruncatedGauss(mu, sigma, low, high, obs, ...)`` (``models/dist_tfp.py:357-417``): the Gauss shape,
    zero outside ``[low, high]``."""

    _KIND = Z.TGAUSS

    def __init__(self, mu, sigma, low, high, obs, *, extended=None, norm=None, name="TruncatedGauss", label=None):
        params = {"mu": convert_to_parameter(mu), "sigma": convert_to_parameter(sigma),
                  "low": convert_to_parameter(low), "high": convert_to_parameter(high)}
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


class Exponential(_Leaf1D):
    """``zfit.pdf.Exponential(lam=None, obs=None, *, extended, norm, name, lambda_=None, label)``
    (``models/basic.py:34-243``)."""

    _KIND = Z.EXP

    def __init__(self, lam=None, obs=None, *, extended=None, norm=None, name="Exponential", lambda_=None, label=None):
        if lambda_ is not None:
            if lam is not None:
                from .exception import BreakingAPIChangeError

                msg = "The 'lambda' parameter has been renamed from 'lambda_' to 'lam'."
                raise BreakingAPIChangeError(msg)
            lam = lambda_
        lam = convert_to_parameter(lam)
        self._lam = lam
        super().__init__(obs, {"lambda": lam, "lam": lam}, extended=extended, norm=norm, name=name, label=label)

    def _slot_params(self):
        return [self._lam]

    def _shift(self, ctx, norm_space):
        # centre of the norm handed to the hook; 0 with norm=False i.e. inside ProductPDF (basic.py:98,131-154)
        if ctx == "prod":
            return 0.0
        lo, hi = norm_space.limit1d
        return (hi + lo) / 2


class CrystalBall(_Leaf1D):
    """``zfit.pdf.CrystalBall(mu, sigma, alpha, n, obs, *, extended, norm, name, label)``
    (``models/physics.py:233-336``)."""

    _KIND = Z.CB

    def __init__(self, mu, sigma, alpha, n, obs, *, extended=None, norm=None, name="CrystalBall", label=None):
        params = {
            "mu": convert_to_parameter(mu),
            "sigma": convert_to_parameter(sigma),
            "alpha": convert_to_parameter(alpha),
            "n": convert_to_parameter(n),
        }
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


class _RecursivePolynomial(_Leaf1D):
    """``RecursivePolynomial`` (``models/polynomials.py:46-144``): ``sum_k c_k Poly_k(z)`` on the rescaled observable,
    ``c_0 = 1`` unless ``coeff0`` is given."""

    def __init__(self, obs, coeffs, apply_scaling=True, coeff0=None, *, extended=None, norm=None, name=None, label=None):
        if not apply_scaling:
            msg = f"{type(self).__name__} without rescaling is not lowered to the fused kernel"
            raise ModelNotOnDeviceError(msg)
        space = obs if isinstance(obs, Space) else Space(obs)
        if not space.has_limits:
            msg = "obs need to be a Space with exactly one limit if rescaling is requested."
            raise ValueError(msg)
        coeffs = [coeffs] if isinstance(coeffs, (ZfitParameter, int, float)) else list(coeffs)
        c0 = convert_to_parameter(1.0 if coeff0 is None else coeff0)
        allc = [c0] + [convert_to_parameter(c) for c in coeffs]
        if len(allc) - 1 > 16:
            msg = "polynomial degree > 16 is not supported by the fused kernel"
            raise ModelNotOnDeviceError(msg)
        self._degree = len(allc) - 1
        super().__init__(space, {f"c_{i}": c for i, c in enumerate(allc)}, extended=extended, norm=norm,
                         name=name or type(self).__name__, label=label)

    @property
    def degree(self):
        return self._degree

    def _node(self, column, lo, hi, shift):
        slo, shi = self._space.limit1d
        return Z.Node(self._KIND, 0, column, self._degree, lo, hi, slo, shi, 0.0)


class Chebyshev(_RecursivePolynomial):
    """``zfit.pdf.Chebyshev(obs, coeffs, apply_scaling=True, coeff0=None, *, extended, norm, name, label)``
    (``models/polynomials.py:334-482``)."""

    _KIND = Z.CHEB


class Legendre(_RecursivePolynomial):
    """``zfit.pdf.Legendre`` (``models/polynomials.py:181-332``)."""

    _KIND = Z.LEGENDRE


class Chebyshev2(_RecursivePolynomial):
    """``zfit.pdf.Chebyshev2`` -- second kind, ``U_1 = 2x`` (``models/polynomials.py:487-600``)."""

    _KIND = Z.CHEB2


class Hermite(_RecursivePolynomial):
    """``zfit.pdf.Hermite`` -- physicists' ``H_n``, ``H_1 = 2x`` (``models/polynomials.py:752-869``)."""

    _KIND = Z.HERMITE


class Laguerre(_RecursivePolynomial):
    """``zfit.pdf.Laguerre`` -- ``L_1 = 1 - x`` (``models/polynomials.py:600-750``)."""

    _KIND = Z.LAGUERRE


class Bernstein(_Leaf1D):
    """``zfit.pdf.Bernstein(obs, coeffs, apply_scaling=True, *, extended, norm, name, label)``
    (``models/polynomials.py:872-1028``): ``sum_k c_k B_{k,n}(t)`` with ``t`` the observable rescaled to ``(0, 1)``; every
    coefficient is a parameter (there is no implicit ``c_0 = 1``)."""

    _KIND = Z.BERNSTEIN

    def __init__(self, obs, coeffs, apply_scaling=True, *, extended=None, norm=None, name="Bernstein", label=None):
        if not apply_scaling:
            msg = "Bernstein without rescaling is not lowered to the fused kernel"
            raise ModelNotOnDeviceError(msg)
        space = obs if isinstance(obs, Space) else Space(obs)
        if not space.has_limits:
            msg = "obs need to be a Space with exactly one limit if rescaling is requested."
            raise ValueError(msg)
        coeffs = [coeffs] if isinstance(coeffs, (ZfitParameter, int, float)) else list(coeffs)
        if len(coeffs) < 1:
            msg = "Need at least one coefficient in de_casteljau of Bernstein."
            raise ValueError(msg)
        if len(coeffs) - 1 > 16:
            msg = "polynomial degree > 16 is not supported by the fused kernel"
            raise ModelNotOnDeviceError(msg)
        self._degree = len(coeffs) - 1
        super().__init__(space, {f"c_{i}": convert_to_parameter(c) for i, c in enumerate(coeffs)}, extended=extended,
                         norm=norm, name=name, label=label)

    @property
    def degree(self):
        return self._degree

    def _node(self, column, lo, hi, shift):
        slo, shi = self._space.limit1d
        return Z.Node(self._KIND, 0, column, self._degree, lo, hi, slo, shi, 0.0)


class GeneralizedCB(_Leaf1D):
    """``zfit.pdf.GeneralizedCB(mu, sigmal, alphal, nl, sigmar, alphar, nr, obs, ...)`` (``models/physics.py:47-80,
    186-232, 506-640``): CrystalBall(mu, sigmal, alphal, nl) below ``mu``, CrystalBall(mu, sigmar, -alphar, nr) above."""

    _KIND = Z.GCB

    def __init__(self, mu, sigmal, alphal, nl, sigmar, alphar, nr, obs, *, extended=None, norm=None,
                 name="GeneralizedCB", label=None):
        params = {
            "mu": convert_to_parameter(mu),
            "sigmal": convert_to_parameter(sigmal),
            "alphal": convert_to_parameter(alphal),
            "nl": convert_to_parameter(nl),
            "sigmar": convert_to_parameter(sigmar),
            "alphar": convert_to_parameter(alphar),
            "nr": convert_to_parameter(nr),
        }
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


class DoubleCB(_Leaf1D):
    """``zfit.pdf.DoubleCB(mu, sigma, alphal, nl, alphar, nr, obs, ...)`` (``models/physics.py:339-503``): the
    generalized node with one ``sigma`` feeding both sides (two slots, one parameter)."""

    _KIND = Z.GCB

    def __init__(self, mu, sigma, alphal, nl, alphar, nr, obs, *, extended=None, norm=None, name="DoubleCB", label=None):
        params = {
            "mu": convert_to_parameter(mu),
            "sigma": convert_to_parameter(sigma),
            "alphal": convert_to_parameter(alphal),
            "nl": convert_to_parameter(nl),
            "alphar": convert_to_parameter(alphar),
            "nr": convert_to_parameter(nr),
        }
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)

    def _slot_params(self):
        p = self._params
        return [p["mu"], p["sigma"], p["alphal"], p["nl"], p["sigma"], p["alphar"], p["nr"]]


class GaussExpTail(_Leaf1D):
    """``zfit.pdf.GaussExpTail(mu, sigma, alpha, obs, ...)`` (``models/physics.py:727-805``)."""

    _KIND = Z.GET

    def __init__(self, mu, sigma, alpha, obs, *, extended=None, norm=None, name="GaussExpTail", label=None):
        params = {"mu": convert_to_parameter(mu), "sigma": convert_to_parameter(sigma), "alpha": convert_to_parameter(alpha)}
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


class GeneralizedGaussExpTail(_Leaf1D):
    """``zfit.pdf.GeneralizedGaussExpTail(mu, sigmal, alphal, sigmar, alphar, obs, ...)`` (``models/physics.py:819-923``)."""

    _KIND = Z.GGET

    def __init__(self, mu, sigmal, alphal, sigmar, alphar, obs, *, extended=None, norm=None,
                 name="GeneralizedGaussExpTail", label=None):
        params = {
            "mu": convert_to_parameter(mu),
            "sigmal": convert_to_parameter(sigmal),
            "alphal": convert_to_parameter(alphal),
            "sigmar": convert_to_parameter(sigmar),
            "alphar": convert_to_parameter(alphar),
        }
        super().__init__(obs, params, extended=extended, norm=norm, name=name, label=label)


# ------------------------------------------------------------------------------------------------
# functors
# ------------------------------------------------------------------------------------------------
class BaseFunctor(BasePDF):
    def __init__(self, pdfs, obs, params, **kw):
        self.pdfs = list(pdfs)
        super().__init__(obs, params, **kw)

    @property
    def models(self):
        return self.pdfs

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        params = super().get_params(floating, is_yield, extract_independent)
        if is_yield is not True:  # daughters contribute with is_yield=False (basefunctor.py:82-100)
            for m in self.pdfs:
                params = params.union(m.get_params(floating=floating, is_yield=False, extract_independent=extract_independent))
        return params


class SumPDF(BaseFunctor):
    """``zfit.pdf.SumPDF(pdfs, fracs=None, obs=None, extended=None, norm=None, name="SumPDF", label=None)``
    (``models/functor.py:74-313``, ``models/basefunctor.py:160-255``)."""

    def __init__(self, pdfs, fracs=None, obs=None, extended=None, norm=None, name="SumPDF", label=None):
        pdfs = list(pdfs) if isinstance(pdfs, (list, tuple)) else [pdfs]
        if len(pdfs) < 2:
            msg = f"Cannot build a sum of less than two pdfs {pdfs}"
            raise ValueError(msg)
        common = convert_to_obs_str(obs) if obs is not None else pdfs[0].obs
        if any(frozenset(p.obs) != frozenset(common) for p in pdfs):
            msg = "Currently, sums are only supported in the same observables"
            raise ObsIncompatibleError(msg)
        space = obs if isinstance(obs, Space) else pdfs[0].space.with_obs(common)
        all_ext = all(p.is_extended for p in pdfs)
        if fracs is None:
            fracs = []
        elif isinstance(fracs, (ZfitParameter, int, float)):
            fracs = [fracs]
        fracs = [convert_to_parameter(f) for f in fracs]
        self._frac_param_created = False
        sum_yields = None
        if fracs:
            if len(fracs) == len(pdfs) - 1:
                self._frac_param_created = True
                remaining = OneMinusSum("Composed_autoparam_remaining_frac", fracs)
                param_fracs = [*fracs, remaining]
            elif len(fracs) == len(pdfs):
                param_fracs = fracs
            else:
                msg = (
                    f"If all PDFs are not extended {pdfs}, the fracs {fracs} have to be of"
                    f" the same length as pdf or one less."
                )
                raise ModelIncompatibleError(msg)
            fracs_cleaned = param_fracs
        else:
            if not all_ext:
                msg = f"Not all pdf {pdfs} are extended and no fracs {fracs} are provided."
                raise ModelIncompatibleError(msg)
            yields = [p.get_yield() for p in pdfs]
            sum_yields = SumOfParameters("Composed_autoparam_sum_yields", yields)
            param_fracs = [FractionOfSum(f"Composed_autoparam_frac_{i}", y, sum_yields) for i, y in enumerate(yields)]
            fracs_cleaned = None
        self._fracs = param_fracs
        self._original_fracs = fracs_cleaned
        self._automatically_extended = False
        if extended is None:
            extended = all_ext and not fracs_cleaned
        if extended is True and all_ext and not fracs_cleaned:
            self._automatically_extended = True
            extended = sum_yields
        if extended is True or extended is False:
            extended = None
        params = {f"frac_{i}": f for i, f in enumerate(param_fracs)}
        super().__init__(pdfs, space, params, extended=extended, norm=norm, name=name, label=label)

    @property
    def fracs(self):
        return self._fracs

    @property
    def _fracs_sum_to_one(self):
        return self._frac_param_created or self._original_fracs is None

    def _emit(self, colmap, ctx, limits, nodes, slots):
        if ctx == "norm" and limits is None and self.norm is not None and any(p.norm != self.norm.with_obs(p.obs) for p in self.pdfs):
            # no explicit range requested, but the sum's own norm differs from a daughter's: the reference's _pdf then
            # raises SpecificFunctionNotImplemented as well (equal_norm_ranges is False, functor.py:199-201) and the
            # value is _unnormalized_pdf / normalization(self.norm) -- the same branch as below, over self.norm
            limits = self.norm
        if ctx == "norm" and limits is not None and any(p.norm != limits.with_obs(p.obs) for p in self.pdfs):
            # reference: SpecificFunctionNotImplemented -> _unnormalized_pdf / normalization(norm) (functor.py:189-206):
            #   sum_i f_i pdf_i(x) / sum_j f_j int_norm pdf_j
            # == sum_i f'_i U_i(x)/I_i(norm)  with  f'_i = f_i r_i / sum_j f_j r_j,  r_i = I_i(norm)/I_i(own norm):
            # leaves normalised over the requested range + composed fractions (their slot derivatives come from the
            # analytic integral derivatives of the C library's host prologue).
            if not all(isinstance(p, _Leaf1D) for p in self.pdfs):
                msg = "a SumPDF of functors with a norm that differs from its daughters' norm is not lowered yet"
                raise ModelNotOnDeviceError(msg)
            rng = limits.with_obs(self.obs)
            nodes.append(Z.Node(Z.SUM, len(self.pdfs), 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0))
            slots.extend(_renormalised_fracs(self, rng))
            for p in self.pdfs:
                p._emit(colmap, "norm", rng, nodes, slots)
            return
        if ctx == "prod":
            rng = limits.with_obs(self.obs)
            if not self._fracs_sum_to_one or any(p.norm != rng for p in self.pdfs):
                msg = (
                    "a SumPDF inside a ProductPDF is lowered only when its fractions sum to one and its daughters"
                    " are normalised over the product's range"
                )
                raise ModelNotOnDeviceError(msg)
        nodes.append(Z.Node(Z.SUM, len(self.pdfs), 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0))
        slots.extend(self._fracs)
        for p in self.pdfs:  # each daughter normalised over its own norm: pdf.pdf(x) (functor.py:204)
            p._emit(colmap, "norm", None, nodes, slots)

    def _raw_integral(self, limits, shift_norm):
        # sum_i frac_i * pdf_i.integrate(limits)  -- daughters normalised (functor.py:218-229)
        tot = 0.0
        for p, f in zip(self.pdfs, self._fracs):
            tot += f.value() * float(p.integrate(limits)[0])
        return tot


def _renormalised_fracs(sumpdf, rng):
    """f'_i = f_i r_i / sum_j f_j r_j with r_i = I_i(rng) / I_i(own norm), as ComposedParameters with analytic partials."""
    pdfs, fracs = sumpdf.pdfs, sumpdf._fracs
    k = len(pdfs)
    deps = {}
    layout = []  # per daughter: (keys of its slot params, keys of its frac)
    for i, (p, f) in enumerate(zip(pdfs, fracs)):
        skeys = []
        for j, sp in enumerate(p._slot_params()):
            key = f"s{i}_{j}"
            deps[key] = sp
            skeys.append(key)
        deps[f"f{i}"] = f
        layout.append(skeys)

    def ratios(vals):
        r, dr = [], []
        for i, p in enumerate(pdfs):
            sv = [vals[key] for key in layout[i]]
            lo, hi = rng.limit1d
            olo, ohi = p.norm.limit1d
            shift = p._shift("norm", p.norm)  # both integrals of the SAME unnormalised function (one data shift)
            I1, d1 = Z.leaf_integral(p._node(0, lo, hi, shift), sv)
            I0, d0 = Z.leaf_integral(p._node(0, olo, ohi, shift), sv)
            r.append(I1 / I0)
            dr.append((d1 - (I1 / I0) * d0) / I0)
        return r, dr

    def make(i):
        def value(vals):
            r, _ = ratios(vals)
            tot = sum(float(vals[f"f{j}"]) * r[j] for j in range(k))
            return float(vals[f"f{i}"]) * r[i] / tot

        def grad(vals):
            r, dr = ratios(vals)
            f = [float(vals[f"f{j}"]) for j in range(k)]
            tot = sum(f[j] * r[j] for j in range(k))
            fi = f[i] * r[i] / tot
            out = {}
            for j in range(k):
                out[f"f{j}"] = ((r[i] if j == i else 0.0) - fi * r[j]) / tot
                for key, d in zip(layout[j], dr[j]):
                    out[key] = ((f[i] * d if j == i else 0.0) - fi * f[j] * d) / tot
            return out

        return ComposedParameter(f"Composed_autoparam_renorm_frac_{i}_{sumpdf.name}", value, params=deps, grad=grad)

    return [make(i) for i in range(k)]


class ProductPDF(BaseFunctor):
    """``zfit.pdf.ProductPDF(pdfs, obs=None, extended=None, norm=None, name="ProductPDF", label=None)`` over
    disjoint observables (``models/functor.py:331-483``)."""

    def __init__(self, pdfs, obs=None, extended=None, norm=None, name="ProductPDF", label=None):
        pdfs = list(pdfs) if isinstance(pdfs, (list, tuple)) else [pdfs]
        seen = set()
        for p in pdfs:
            if seen & set(p.obs):
                msg = "ProductPDF with overlapping observables is not lowered to the fused kernel (numeric integration)"
                raise ModelNotOnDeviceError(msg)
            seen |= set(p.obs)
        space = combine_spaces(*[p.space for p in pdfs])
        if obs is not None:
            space = space.with_obs(convert_to_obs_str(obs))
        if extended is True or extended is False:
            extended = None
        super().__init__(pdfs, space, {}, extended=extended, norm=norm, name=name, label=label)

    def _emit(self, colmap, ctx, limits, nodes, slots):
        rng = limits if limits is not None else self.norm
        nodes.append(Z.Node(Z.PROD, len(self.pdfs), 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0))
        for p in self.pdfs:  # pdf_i.pdf(x, norm=False) / pdf_i.integrate(limits_i, norm=False) (functor.py:408-433)
            p._emit(colmap, "prod", rng.with_obs(p.obs), nodes, slots)

    def _raw_integral(self, limits, shift_norm):
        tot = 1.0
        for p in self.pdfs:
            tot *= p._raw_integral(limits.with_obs(p.obs), shift_norm=False)
        return tot
