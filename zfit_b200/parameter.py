"""Parameters for the hot path (``src/zfit/core/parameter.py``).

``Parameter`` is a host-side fp64 scalar (the reference's is a ``tf.Variable``).  PDF arguments
("slots" of the fused kernel) may also be ``ConstantParameter`` or ``ComposedParameter``; the loss
differentiates the kernel's slot gradient through them with the chain rule
(``d slot / d theta``), which replaces the reference's tape through ``ComposedParameter``
(``parameter.py:1067-1218``) -- see SURVEY.md section 7 "Composed parameters".
"""

from __future__ import annotations

import contextlib
from collections.abc import Iterable, Mapping

import numpy as np

from .exception import ParameterNotIndependentError


class ZfitParameter:
    """Marker base (``_interfaces.py``)."""

    _independent = False
    floating = False

    @property
    def independent(self) -> bool:
        return self._independent

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        out = []
        for p in self._leaves():
            if floating is not None and p.floating != floating:
                continue
            if p not in out:
                out.append(p)
        return OrderedSet(out)

    # arithmetic -> ComposedParameter, so that things like ``mu + 5`` work (parameter.py:241-340)
    def _binop(self, other, op, sym):
        other_p = convert_to_parameter(other)
        return ComposedParameter(
            f"{self.name}{sym}{getattr(other_p, 'name', other)}",
            lambda params: op(params["a"], params["b"]),
            params={"a": self, "b": other_p},
        )

    def __add__(self, o):
        return self._binop(o, lambda a, b: a + b, "+")

    def __radd__(self, o):
        return convert_to_parameter(o)._binop(self, lambda a, b: a + b, "+")

    def __sub__(self, o):
        return self._binop(o, lambda a, b: a - b, "-")

    def __rsub__(self, o):
        return convert_to_parameter(o)._binop(self, lambda a, b: a - b, "-")

    def __mul__(self, o):
        return self._binop(o, lambda a, b: a * b, "*")

    def __rmul__(self, o):
        return convert_to_parameter(o)._binop(self, lambda a, b: a * b, "*")

    def __truediv__(self, o):
        return self._binop(o, lambda a, b: a / b, "/")

    def __rtruediv__(self, o):
        return convert_to_parameter(o)._binop(self, lambda a, b: a / b, "/")

    def __neg__(self):
        return ComposedParameter(f"-{self.name}", lambda params: -params["a"], params={"a": self})

    def __float__(self):
        return float(self.value())

    def numpy(self):
        return np.float64(self.value())

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value(), dtype=dtype or np.float64)


class OrderedSet(list):
    """Insertion-ordered set with the little API the loss / minimizers use (``ordered_set.OrderedSet``).
    Membership is by identity (parameters are unique objects), O(1) through an id index."""

    def __init__(self, it=()):
        super().__init__()
        self._ids = set()
        for x in it:
            self.add(x)

    def add(self, x):
        if id(x) not in self._ids:
            self._ids.add(id(x))
            self.append(x)

    def union(self, *others):
        out = OrderedSet(self)
        for o in others:
            for x in o:
                out.add(x)
        return out

    def __contains__(self, x):
        return id(x) in self._ids

    def __sub__(self, other):
        return OrderedSet(x for x in self if x not in other)

    def __or__(self, other):
        return self.union(other)

    def __eq__(self, other):
        if isinstance(other, (set, frozenset)):
            return len(self) == len(other) and all(x in other for x in self)
        return list.__eq__(self, other)

    __hash__ = None


class TemporarilySet:
    """Result of ``set_value`` / ``set_values``: already applied; restores on context exit
    (``src/zfit/util/temporary.py``)."""

    def __init__(self, restore):
        self._restore = restore

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._restore()
        return False


class Parameter(ZfitParameter):
    """``zfit.Parameter(name, value, lower=None, upper=None, stepsize=None, floating=True, *, label=None)``
    (``parameter.py:350-436``)."""

    _independent = True
    DEFAULT_stepsize = 0.01

    def __init__(self, name, value, lower=None, upper=None, stepsize=None, floating=True, *, label=None,
                 step_size=None, prior=None):
        if prior is not None:
            msg = "priors / Bayesian inference are outside the hot path (SURVEY.md section 2, row 23)"
            raise NotImplementedError(msg)
        self.name = str(name)
        self.label = label if label is not None else self.name
        self.lower = None if lower is None else float(lower)
        self.upper = None if upper is None else float(upper)
        self.floating = bool(floating)
        self._stepsize = None
        self.stepsize = stepsize if stepsize is not None else step_size
        self._value = float(value)
        self.set_value(value)  # checks the limits

    # -- state
    @property
    def has_limits(self) -> bool:
        return self.lower is not None or self.upper is not None

    @property
    def at_limit(self) -> bool:
        if not self.has_limits:
            return False
        lo = -np.inf if self.lower is None else self.lower
        up = np.inf if self.upper is None else self.upper
        tol = min((up - lo) * 1e-5, 1e-15) if np.isfinite(up - lo) else 1e-15
        return bool(self._value <= lo + tol or self._value >= up - tol)

    @property
    def stepsize(self) -> float:
        return self.DEFAULT_stepsize if self._stepsize is None else self._stepsize

    @stepsize.setter
    def stepsize(self, v):
        self._stepsize = None if v is None else float(v)

    step_size = stepsize

    @property
    def has_stepsize(self) -> bool:
        return self._stepsize is not None

    def value(self) -> float:
        """Current value, clipped into the limits with a pass-through gradient
        (``tfp.math.clip_by_value_preserve_gradient``, ``parameter.py:548-557``)."""
        v = self._value
        if self.lower is not None and v < self.lower:
            v = self.lower
        if self.upper is not None and v > self.upper:
            v = self.upper
        return v

    def read_value(self):
        return self.value()

    def _outside(self, v) -> bool:
        tol = 1e-15
        if self.lower is not None and v <= self.lower - tol:
            return True
        return bool(self.upper is not None and v >= self.upper + tol)

    def set_value(self, value, *, clip=None):
        value = float(np.asarray(value, dtype=np.float64).reshape(()))
        if self.has_limits:
            if clip:
                if self.lower is not None:
                    value = max(value, self.lower)
                if self.upper is not None:
                    value = min(value, self.upper)
            elif self._outside(value):
                msg = (
                    f"Setting value {value} invalid for parameter {self.name} with limits "
                    f"{self.lower} - {self.upper}. In order to silence this and clip the value, you can use clip=True"
                )
                raise ValueError(msg)
        old = self._value
        self._value = value

        def restore():
            self._value = old

        return TemporarilySet(restore)

    def assign(self, value, **_):
        """Unchecked fast path (``parameter.py:680-700``)."""
        self._value = float(value)

    def randomize(self, minval=None, maxval=None, sampler=np.random.uniform):
        lo = self.lower if minval is None else minval
        up = self.upper if maxval is None else maxval
        if lo is None or up is None:
            msg = "Cannot randomize a parameter without limits"
            raise ValueError(msg)
        v = float(sampler(lo, up))
        self.set_value(v)
        return v

    def _leaves(self):
        return [self]

    def _partials(self):
        """{independent parameter: d value / d parameter}"""
        return {self: 1.0}

    def __repr__(self):
        return f"<zfit.Parameter '{self.name}' floating={self.floating} value={self._value:.4g}>"

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other


class ConstantParameter(ZfitParameter):
    """Numbers given as PDF arguments (``convert_to_parameter``, ``parameter.py:982-1060``)."""

    _independent = False
    floating = False

    def __init__(self, name, value, *, label=None):
        self.name = str(name)
        self.label = label or self.name
        self._value = float(value)
        self.lower = self.upper = None

    def value(self):
        return self._value

    @property
    def has_limits(self):
        return False

    def _leaves(self):
        return []

    def _partials(self):
        return {}

    def __repr__(self):
        return f"<zfit.ConstantParameter '{self.name}' value={self._value:.4g}>"

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other


class ComposedParameter(ZfitParameter):
    """``zfit.ComposedParameter(name, func, params)`` (``parameter.py:1067-1218``): a Python function of other
    parameters.  ``grad`` may supply analytic partials ``{key: d value/d params[key]}``; otherwise the partials
    come from torch autograd on fp64 scalars when ``func`` is torch-traceable, else from central differences.
    """

    _independent = False
    floating = False

    def __init__(self, name, func=None, params=None, *, grad=None, label=None, value_fn=None, **kwargs):
        if func is None and value_fn is not None:
            func = value_fn
        if "dependents" in kwargs and params is None:
            params = kwargs.pop("dependents")
        self.name = str(name)
        self.label = label or self.name
        if params is None:
            params = {}
        if isinstance(params, Mapping):
            self.params = {k: convert_to_parameter(v) for k, v in params.items()}
            self._as_mapping = True
        else:
            if isinstance(params, ZfitParameter) or not isinstance(params, Iterable):
                params = [params]
            self.params = {f"param_{i}": convert_to_parameter(p) for i, p in enumerate(params)}
            self._as_mapping = False
        self._func = func
        self._grad = grad
        self.lower = self.upper = None

    @property
    def has_limits(self):
        return False

    def _call(self, values: dict):
        if self._as_mapping:
            return self._func(values)
        return self._func(list(values.values()))

    def value(self) -> float:
        return float(self._call({k: p.value() for k, p in self.params.items()}))

    def _leaves(self):
        out = []
        for p in self.params.values():
            for leaf in p._leaves():
                if not any(leaf is o for o in out):
                    out.append(leaf)
        return out

    def _local_partials(self) -> dict:
        vals = {k: p.value() for k, p in self.params.items()}
        if self._grad is not None:
            return {k: float(v) for k, v in self._grad(vals).items()}
        try:  # exact: torch autograd on fp64 scalars (plumbing only, O(1) work per call)
            import torch

            tv = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in vals.items()}
            out = self._call(tv)
            if isinstance(out, torch.Tensor) and out.requires_grad:
                gs = torch.autograd.grad(out, list(tv.values()), allow_unused=True)
                return {k: 0.0 if g is None else float(g) for k, g in zip(tv, gs)}
        except Exception:
            pass
        partials = {}
        for k, v in vals.items():  # central differences with one Richardson step
            h = 1e-4 * max(1.0, abs(v))

            def d(hh, k=k, v=v):
                up, dn = dict(vals), dict(vals)
                up[k], dn[k] = v + hh, v - hh
                return (float(self._call(up)) - float(self._call(dn))) / (2 * hh)

            partials[k] = (4.0 * d(h) - d(2 * h)) / 3.0
        return partials

    def _partials(self, memo=None):
        """{independent leaf: d value / d leaf}; ``memo`` shares sub-results within one loss evaluation."""
        if memo is not None:
            hit = memo.get(id(self))
            if hit is not None:
                return hit
        local = self._local_partials()
        out: dict = {}
        for k, p in self.params.items():
            lk = local.get(k, 0.0)
            if p._independent:
                out[p] = out.get(p, 0.0) + lk
            elif lk != 0.0 or True:
                sub = p._partials(memo) if isinstance(p, ComposedParameter) else p._partials()
                for leaf, d in sub.items():
                    out[leaf] = out.get(leaf, 0.0) + lk * d
        if memo is not None:
            memo[id(self)] = out
        return out

    def set_value(self, value):
        msg = "Cannot set the value of a composed parameter"
        raise ParameterNotIndependentError(msg)

    def __repr__(self):
        return f"<zfit.ComposedParameter '{self.name}' params={[p.name for p in self.params.values()]} value={self.value():.4g}>"

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other


class SumOfParameters(ComposedParameter):
    """``sum_yields`` of an automatically extended SumPDF (``models/basefunctor.py:231-236``), with a direct fast path."""

    def __init__(self, name, terms):
        terms = [convert_to_parameter(t) for t in terms]
        super().__init__(name, lambda params: sum(params.values()), params={f"yield_{i}": t for i, t in enumerate(terms)},
                         grad=lambda vals: dict.fromkeys(vals, 1.0))
        self._terms = terms
        self._plain = all(t._independent for t in terms)

    def value(self):
        tot = 0.0
        for t in self._terms:
            tot += t.value()
        return tot

    def _partials(self, memo=None):
        if self._plain:
            out = {}
            for t in self._terms:
                out[t] = out.get(t, 0.0) + 1.0
            return out
        return super()._partials(memo)


class FractionOfSum(ComposedParameter):
    """``frac_i = yield_i / sum(yields)`` (``models/basefunctor.py:237-244``), with a direct fast path."""

    def __init__(self, name, term, total: SumOfParameters):
        term = convert_to_parameter(term)
        super().__init__(name, lambda params: params["yield_"] / params["sum_yields"],
                         params={"sum_yields": total, "yield_": term},
                         grad=lambda v: {"yield_": 1.0 / v["sum_yields"], "sum_yields": -v["yield_"] / v["sum_yields"] ** 2})
        self._term, self._total = term, total
        self._plain = term._independent and total._plain

    def value(self):
        return self._term.value() / self._total.value()

    def _partials(self, memo=None):
        if not self._plain:
            return super()._partials(memo)
        tot = self._total.value()
        y = self._term.value()
        out = {}
        for t in self._total._terms:
            out[t] = out.get(t, 0.0) - y / (tot * tot)
        out[self._term] = out.get(self._term, 0.0) + 1.0 / tot
        return out


class OneMinusSum(ComposedParameter):
    """Remaining fraction ``1 - sum(fracs)`` (``models/basefunctor.py:189-205``), with a direct fast path."""

    def __init__(self, name, terms):
        terms = [convert_to_parameter(t) for t in terms]
        super().__init__(name, lambda params: 1.0 - sum(params.values()), params={f"frac_{i}": t for i, t in enumerate(terms)},
                         grad=lambda vals: dict.fromkeys(vals, -1.0))
        self._terms = terms
        self._plain = all(t._independent for t in terms)

    def value(self):
        tot = 1.0
        for t in self._terms:
            tot -= t.value()
        return tot

    def _partials(self, memo=None):
        if self._plain:
            out = {}
            for t in self._terms:
                out[t] = out.get(t, 0.0) - 1.0
            return out
        return super()._partials(memo)


_auto_names = {"n": 0}


def convert_to_parameter(value, name=None, prefer_constant=True, params=None) -> ZfitParameter:
    """``convert_to_parameter`` (``parameter.py:1485-1561``): parameter passthrough, number -> constant,
    callable + ``params`` -> composed."""
    if isinstance(value, ZfitParameter):
        return value
    if callable(value):
        _auto_names["n"] += 1
        return ComposedParameter(name or f"Composed_autoparam_{_auto_names['n']}", value, params=params)
    _auto_names["n"] += 1
    name = name or f"FIXED_autoparam_{_auto_names['n']}"
    if prefer_constant:
        return ConstantParameter(name, float(np.asarray(value, dtype=np.float64).reshape(())))
    return Parameter(name, value, floating=False)


def convert_to_parameters(values, **_):
    return [convert_to_parameter(v) for v in values]


def assign_values(params, values, **_):
    """Unchecked fast path used by ``LossEval`` (``parameter.py:1564-1613``)."""
    if isinstance(values, Mapping):
        for p, v in values.items():
            p.assign(v)
        return
    if isinstance(params, ZfitParameter):
        params, values = [params], [values]
    for p, v in zip(params, np.asarray(values, dtype=np.float64).reshape(-1).tolist()):
        p.assign(v)


assign_values_jit = assign_values


def set_values(params, values=None, allow_partial=None, clip=None):
    """``zfit.param.set_values`` (``parameter.py:1616-1659``): plain call or context manager."""
    if values is None and isinstance(params, Mapping):
        params, values = list(params.keys()), list(params.values())
    if isinstance(params, ZfitParameter):
        params = [params]
    params = list(params)
    if hasattr(values, "params") and hasattr(values, "values") and not isinstance(values, Mapping):  # FitResult
        result = values
        new_params, new_values = [], []
        for p in params:
            if p in result.params:
                new_params.append(p)
                new_values.append(result.params[p]["value"])
            elif not allow_partial:
                msg = f"Parameter {p} not in the FitResult; use allow_partial=True"
                raise ValueError(msg)
        params, values = new_params, new_values
    elif isinstance(values, Mapping):
        byname = {(k.name if isinstance(k, ZfitParameter) else k): v for k, v in values.items()}
        new_params, new_values = [], []
        for p in params:
            if p.name in byname:
                new_params.append(p)
                new_values.append(byname[p.name])
            elif not allow_partial:
                msg = f"Parameter {p} not in values {values}"
                raise ValueError(msg)
        params, values = new_params, new_values
    else:
        values = np.atleast_1d(np.asarray(values, dtype=np.float64)).reshape(-1)
        if len(values) != len(params):
            msg = f"Incompatible length of parameters and values: {len(params)} vs {len(values)}"
            raise ValueError(msg)
    restores = [p.set_value(v, clip=clip) for p, v in zip(params, values)]

    def restore():
        for r in reversed(restores):
            r._restore()

    return TemporarilySet(restore)


@contextlib.contextmanager
def temporarily(params, values):
    with set_values(params, values):
        yield
