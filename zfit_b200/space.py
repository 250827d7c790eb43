"""``zfit.Space`` for the hot path: named observables with (optional) rectangular limits.

Mirrors the part of ``src/zfit/core/space.py:1246-2239`` the five configs and the acceptance
scripts touch: 1-D and N-D rectangular spaces, the ``*`` product, ``with_obs``, ``inside`` (inclusive on
both ends, ``space.py:218-231``), ``limit1d``, ``volume``, equality on obs + limits.  ``MultiSpace``,
functional limits and binning are out of scope.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from .exception import BreakingAPIChangeError, LimitsIncompatibleError, ObsIncompatibleError


def convert_to_obs_str(obs) -> tuple[str, ...]:
    if obs is None:
        return ()
    if isinstance(obs, Space):
        return obs.obs
    if isinstance(obs, str):
        return (obs,)
    return tuple(str(o) for o in obs)


class Space:
    def __init__(self, obs=None, *args, limits=None, lower=None, upper=None, name=None, label=None, **kwargs):
        if kwargs.get("binning") is not None:
            msg = "binned spaces are outside the unbinned hot path"
            raise NotImplementedError(msg)
        if len(args) > 2:
            msg = "Space takes lower, upper as the first two arguments after obs."
            raise BreakingAPIChangeError(msg)
        if len(args) == 2:
            limits = args
        elif len(args) == 1:
            limits = args[0]
        elif lower is not None and upper is not None:
            limits = (lower, upper)
        elif (lower is None) != (upper is None):
            msg = "Both lower and upper have to be given (or none of them)."
            raise ValueError(msg)
        if isinstance(obs, Space):
            if limits is None and obs.has_limits:
                limits = (obs._lower, obs._upper)
            obs = obs.obs
        self._obs = convert_to_obs_str(obs)
        if len(set(self._obs)) != len(self._obs):
            msg = f"Observables have to be unique: {self._obs}"
            raise ObsIncompatibleError(msg)
        self.name = name or "Space"
        self.label = label
        n = len(self._obs)
        if limits is None or limits is False:
            self._lower = self._upper = None
        else:
            lo, up = limits
            lo = np.atleast_1d(np.asarray(lo, dtype=np.float64)).reshape(-1)
            up = np.atleast_1d(np.asarray(up, dtype=np.float64)).reshape(-1)
            if lo.size != n or up.size != n:
                msg = f"limits {limits} do not match the {n} observables {self._obs}"
                raise LimitsIncompatibleError(msg)
            if np.any(lo > up):
                msg = f"lower limits {lo} have to be smaller than upper limits {up}"
                raise ValueError(msg)
            self._lower, self._upper = lo, up

    # -- coordinates
    @property
    def obs(self) -> tuple[str, ...]:
        return self._obs

    @property
    def n_obs(self) -> int:
        return len(self._obs)

    @property
    def axes(self):
        return tuple(range(self.n_obs))

    # -- limits
    @property
    def has_limits(self) -> bool:
        return self._lower is not None

    @property
    def limits_are_set(self) -> bool:
        return self.has_limits

    @property
    def n_limits(self) -> int:
        return 1 if self.has_limits else 0

    _depr_n_limits = n_limits

    def __bool__(self):
        return self.has_limits

    def _need_limits(self):
        if not self.has_limits:
            msg = f"Space {self} has no limits"
            raise LimitsIncompatibleError(msg)

    @property
    def lower(self) -> np.ndarray:
        self._need_limits()
        return self._lower[None, :].copy()

    @property
    def upper(self) -> np.ndarray:
        self._need_limits()
        return self._upper[None, :].copy()

    @property
    def rect_limits(self):
        return self.lower, self.upper

    rect_limits_np = rect_limits

    @property
    def limits(self):
        return self.rect_limits

    @property
    def v1(self):
        return _V1(self)

    @property
    def limit1d(self) -> tuple[float, float]:
        self._need_limits()
        if self.n_obs != 1:
            msg = f"limit1d needs a 1-D space, this one has obs {self.obs}"
            raise RuntimeError(msg)
        return float(self._lower[0]), float(self._upper[0])

    @property
    def limit2d(self):
        self._need_limits()
        if self.n_obs != 2:
            msg = "limit2d needs a 2-D space"
            raise RuntimeError(msg)
        return (float(self._lower[0]), float(self._lower[1]), float(self._upper[0]), float(self._upper[1]))

    @property
    def volume(self) -> float:
        self._need_limits()
        return float(np.prod(self._upper - self._lower))

    def rect_area(self):
        return self.volume

    area = rect_area

    # -- operations
    def with_limits(self, limits=None, *, lower=None, upper=None, name=None):
        if limits is None and lower is not None:
            limits = (lower, upper)
        return Space(self._obs, limits=limits, name=name or self.name)

    def with_obs(self, obs, allow_superset=True, allow_subset=True):
        obs = convert_to_obs_str(obs)
        if obs == self._obs:
            return self
        if not set(obs) <= set(self._obs):
            if not allow_superset:
                msg = f"obs {obs} not a subset of {self._obs}"
                raise ObsIncompatibleError(msg)
            if not set(self._obs) <= set(obs):
                msg = f"cannot reorder/select {obs} from {self._obs}"
                raise ObsIncompatibleError(msg)
            return self  # only a reorder of a subset is meaningful here
        idx = [self._obs.index(o) for o in obs]
        if not self.has_limits:
            return Space(obs, name=self.name)
        return Space(obs, limits=(self._lower[idx], self._upper[idx]), name=self.name)

    def get_subspace(self, obs=None, axes=None, name=None):
        if obs is None and axes is not None:
            obs = [self._obs[a] for a in (axes if isinstance(axes, Iterable) else [axes])]
        return self.with_obs(obs)

    def get_limits(self, obs):
        return self.with_obs(obs).rect_limits

    def inside(self, x, guarantee_limits: bool = False):
        """Boolean mask of events inside the limits, inclusive on both ends (``space.py:229-230``)."""
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None] if self.n_obs == 1 else x[None, :]
        if guarantee_limits or not self.has_limits:
            return np.ones(x.shape[0], dtype=bool)
        return np.all((x >= self._lower[None, :]) & (x <= self._upper[None, :]), axis=-1)

    def filter(self, x, guarantee_limits: bool = False):
        x = np.asarray(x, dtype=np.float64)
        return x[self.inside(x, guarantee_limits)]

    def __mul__(self, other):
        if not isinstance(other, Space):
            return NotImplemented
        return combine_spaces(self, other)

    def __eq__(self, other):
        if other is False or other is None:
            return not self.has_limits and other is False
        if not isinstance(other, Space):
            return NotImplemented
        if self._obs != other._obs or self.has_limits != other.has_limits:
            return False
        if not self.has_limits:
            return True
        return bool(np.array_equal(self._lower, other._lower) and np.array_equal(self._upper, other._upper))

    def __hash__(self):
        lim = None if not self.has_limits else (tuple(self._lower), tuple(self._upper))
        return hash((self._obs, lim))

    def __iter__(self):
        yield self

    def __repr__(self):
        if self.has_limits:
            return f"<zfit Space obs={self._obs}, limits=({self._lower.tolist()}, {self._upper.tolist()})>"
        return f"<zfit Space obs={self._obs}, limits=None>"


class _V1:
    def __init__(self, space):
        self._s = space

    @property
    def limits(self):
        s = self._s
        s._need_limits()
        if s.n_obs == 1:
            return np.array([s._lower[0]]), np.array([s._upper[0]])
        return s._lower.copy(), s._upper.copy()

    @property
    def lower(self):
        return self.limits[0]

    @property
    def upper(self):
        return self.limits[1]


def combine_spaces(*spaces: Space) -> Space:
    """Product of spaces with disjoint observables; identical observables must agree on limits
    (``space.py:2309``)."""
    obs: list[str] = []
    lower: list[float] = []
    upper: list[float] = []
    have = [s.has_limits for s in spaces]
    if any(have) and not all(have):
        msg = "cannot combine spaces with and without limits"
        raise LimitsIncompatibleError(msg)
    for s in spaces:
        for i, o in enumerate(s.obs):
            if o in obs:
                j = obs.index(o)
                if s.has_limits and (lower[j] != s._lower[i] or upper[j] != s._upper[i]):
                    msg = f"limits of obs {o} differ between the spaces"
                    raise LimitsIncompatibleError(msg)
                continue
            obs.append(o)
            if s.has_limits:
                lower.append(float(s._lower[i]))
                upper.append(float(s._upper[i]))
    if all(have):
        return Space(obs, limits=(lower, upper))
    return Space(obs)


def convert_to_space(obs=None, limits=None) -> Space | None:
    if obs is None:
        return None
    if isinstance(obs, Space):
        return obs if limits is None else obs.with_limits(limits)
    return Space(obs, limits=limits)
