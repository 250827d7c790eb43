"""``zfit.Data`` for the hot path (``src/zfit/core/data.py:115-1128, 1203-1244, 1772-1937``).

Events are held as structure-of-arrays fp64 columns (one contiguous array per observable) plus an
optional 1-D weight column -- the layout the fused kernel reads with coalesced ``double2`` loads.
Columns live either on the host (numpy) or already on the GPU (torch CUDA tensors; torch is used
for device memory only).  Events outside the limits of ``obs`` are dropped once, at construction,
inclusive on both ends (``check_cut_data_weights``, ``data.py:1203-1244``; ``space.py:229-230``).
``SamplerData`` (``data.py:1247-1606``) is the in-place resampled dataset of toy studies; its events are generated on the
GPU (``sampler.py``).  ``from_root`` reads branches through ``uproot`` when it is installed; binning is out of scope.
"""

from __future__ import annotations

from collections.abc import Mapping

import numpy as np

from .exception import ObsIncompatibleError, OutsideLimitsError, ShapeIncompatibleError
from .space import Space, convert_to_obs_str


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "device")


def _is_pandas(x) -> bool:
    return type(x).__module__.startswith("pandas") and hasattr(x, "columns")


class Data:
    BATCH_SIZE = 1_000_000  # name kept for parity (data.py:125); unused

    def __init__(self, data, *, obs=None, weights=None, name=None, label=None, guarantee_limits=False,
                 use_hash=None, dtype=None):
        if isinstance(data, Data):
            src = data
            obs = src.space if obs is None else obs
            cols = {o: c for o, c in zip(src.obs, src._cols)}
            weights = src._weights if weights is None else weights
            data = cols
        space = obs if isinstance(obs, Space) else (Space(obs) if obs is not None else None)
        cols, weights = self._ingest(data, space, weights)
        if space is None:
            msg = "obs (a Space or observable names) is required"
            raise ValueError(msg)
        if len(cols) != space.n_obs:
            msg = f"data has {len(cols)} columns but obs {space.obs} has {space.n_obs}"
            raise ShapeIncompatibleError(msg)
        n = int(cols[0].shape[0]) if cols else 0
        if weights is not None and int(weights.shape[0]) != n:
            msg = f"Weights have to have the same length as the data, not {int(weights.shape[0])} != {n}."
            raise ValueError(msg)
        if space.has_limits and not guarantee_limits and n > 0:
            cols, weights = _cut(cols, weights, space)
        self._space = space
        self._cols = cols
        self._weights = weights
        self.name = name or "Data"
        self.label = label or self.name
        self._dev_cache = None
        self._version = 0

    # ------------------------------------------------------------------ ingestion
    @staticmethod
    def _ingest(data, space, weights):
        wname = None
        if isinstance(weights, str):
            wname, weights = weights, None
        if _is_pandas(data):
            names = space.obs if space is not None else tuple(data.columns)
            if wname is not None:
                weights = data[wname].to_numpy()
            cols = [np.ascontiguousarray(data[o].to_numpy(), dtype=np.float64) for o in names]
        elif isinstance(data, Mapping):
            names = space.obs if space is not None else tuple(data.keys())
            missing = [o for o in names if o not in data]
            if missing:
                msg = f"observables {missing} not in the data mapping"
                raise ObsIncompatibleError(msg)
            if wname is not None:
                weights = data[wname]
            cols = [_as_col(data[o]) for o in names]
        elif _is_torch(data):
            t = data.detach()
            if t.dim() == 1:
                t = t[:, None]
            if t.dim() != 2:
                msg = f"Data has to be 2-D, i.e. (nevents, nobs)., not {tuple(t.shape)}"
                raise ValueError(msg)
            import torch

            t = t.to(torch.float64)
            cols = [t[:, i].contiguous() for i in range(t.shape[1])]
        else:
            arr = np.asarray(data, dtype=np.float64)
            arr = np.atleast_1d(arr)
            if arr.ndim == 1:
                n_obs = space.n_obs if space is not None else 1
                arr = arr[:, None] if n_obs == 1 else arr[None, :]
            if arr.ndim != 2:
                msg = f"Data has to be 2-D, i.e. (nevents, nobs)., not {arr.shape}, with data={arr}."
                raise ValueError(msg)
            cols = [np.ascontiguousarray(arr[:, i]) for i in range(arr.shape[1])]
        if weights is not None:
            weights = _as_col(weights)
            if weights.ndim != 1:
                msg = f"Weights have to be 1-D, not {tuple(weights.shape)}."
                raise ValueError(msg)
            if cols and _is_torch(cols[0]) and not _is_torch(weights):
                import torch

                weights = torch.from_numpy(np.ascontiguousarray(weights)).to(cols[0].device)
        return cols, weights

    @classmethod
    def from_numpy(cls, obs, array, weights=None, *, name=None, label=None, **kw):
        return cls(np.asarray(array), obs=obs, weights=weights, name=name, label=label, **kw)

    @classmethod
    def from_tensor(cls, obs, tensor, weights=None, *, name=None, label=None, **kw):
        return cls(tensor, obs=obs, weights=weights, name=name, label=label, **kw)

    @classmethod
    def from_pandas(cls, df, obs=None, weights=None, *, name=None, label=None, **kw):
        if obs is None:
            obs = [c for c in df.columns if c != weights]
        return cls(df, obs=obs, weights=weights, name=name, label=label, **kw)

    @classmethod
    def from_mapping(cls, mapping, obs=None, weights=None, *, name=None, label=None, **kw):
        if obs is None:
            obs = [k for k in mapping if k != weights]
        return cls(mapping, obs=obs, weights=weights, name=name, label=label, **kw)

    @classmethod
    def from_root(cls, path, treepath, obs=None, *, weights=None, obs_alias=None, name=None, label=None,
                  root_dir_options=None, branches=None, branches_alias=None, **kw):
        """``Data.from_root`` (``core/data.py:612-698``): the branches named by ``obs`` (or ``obs_alias[obs]``) of one
        tree, read with ``uproot`` straight into per-observable columns; ``weights`` may name a branch.  ``uproot`` is an
        optional dependency, imported here."""
        from .exception import BreakingAPIChangeError

        if branches:
            msg = "Use `obs` instead of `branches`."
            raise BreakingAPIChangeError(msg)
        if branches_alias is not None:
            msg = "Use `obs_alias` instead of `branches_alias`."
            raise BreakingAPIChangeError(msg)
        if obs is None and obs_alias is None:
            msg = "Either branches or branches_alias has to be specified."
            raise ValueError(msg)
        obs_alias = {} if obs_alias is None else dict(obs_alias)
        if obs is None:
            obs = list(obs_alias.values())
        space = obs if isinstance(obs, Space) else Space(obs)
        wanted = [obs_alias.get(o, o) for o in space.obs]
        weight_branch = weights if isinstance(weights, str) else None
        try:
            import uproot
        except ImportError as error:
            msg = "Data.from_root needs the optional dependency `uproot`, which is not installed"
            raise ImportError(msg) from error
        with uproot.open(path, **(root_dir_options or {}))[treepath] as tree:
            arrays = tree.arrays(expressions=tuple([*wanted, weight_branch] if weight_branch else wanted), library="np")
        columns = {o: np.asarray(arrays[b], dtype=np.float64) for o, b in zip(space.obs, wanted)}
        if weight_branch:
            weights = np.asarray(arrays[weight_branch], dtype=np.float64)
        return cls(columns, obs=space, weights=weights, name=name, label=label, **kw)

    # ------------------------------------------------------------------ properties
    @property
    def space(self) -> Space:
        return self._space

    data_range = space

    @property
    def obs(self):
        return self._space.obs

    @property
    def n_obs(self) -> int:
        return self._space.n_obs

    @property
    def num_entries(self) -> int:
        return int(self._cols[0].shape[0]) if self._cols else 0

    nentries = num_entries
    nevents = num_entries

    @property
    def n_events(self) -> int:
        return self.num_entries

    @property
    def has_weights(self) -> bool:
        return self._weights is not None

    @property
    def weights(self):
        if self._weights is None:
            return None
        return _to_numpy(self._weights)

    @property
    def samplesize(self) -> float:
        """Effective sample size: sum of weights, or N (``data.py:268-275``)."""
        if self._weights is None:
            return float(self.num_entries)
        w = self._weights
        return float(w.sum().item()) if _is_torch(w) else float(np.sum(w))

    @property
    def on_device(self) -> bool:
        return bool(self._cols) and _is_torch(self._cols[0]) and self._cols[0].is_cuda

    @property
    def hashint(self):
        """Changes whenever the events change (``update_data``): losses re-bind and recompute their offset
        (``loss.py:247-262``)."""
        return id(self) + self._version  # cheap, and distinct per (object, version)

    def update_data(self, sample, weights=None, *, guarantee_limits=False):
        """Replace the events in place (``data.py:1031-1060``); objects holding this Data see the new sample."""
        cols, weights = self._ingest(sample, self._space, weights)
        if len(cols) != self._space.n_obs:
            msg = f"data has {len(cols)} columns but obs {self._space.obs} has {self._space.n_obs}"
            raise ShapeIncompatibleError(msg)
        n = int(cols[0].shape[0]) if cols else 0
        if weights is not None and int(weights.shape[0]) != n:
            msg = f"Weights have to have the same length as the data, not {int(weights.shape[0])} != {n}."
            raise ValueError(msg)
        if self._space.has_limits and not guarantee_limits and n > 0:
            cols, weights = _cut(cols, weights, self._space)
        self._cols, self._weights = cols, weights
        self._dev_cache = None
        self._version += 1

    def __len__(self):
        return self.num_entries

    # ------------------------------------------------------------------ access
    def value(self, obs=None):
        cols = self._select(obs)
        if not cols:
            return np.zeros((0, 0))
        out = np.stack([_to_numpy(c) for c in cols], axis=-1)
        if isinstance(obs, str):
            return out[:, 0]
        return out

    def numpy(self):
        return self.value()

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value(), dtype=dtype or np.float64)

    def unstack_x(self, obs=None, always_list=None):
        cols = [_to_numpy(c) for c in self._select(obs)]
        if len(cols) == 1 and not always_list:
            return cols[0]
        return cols

    def __getitem__(self, item):
        if isinstance(item, str):
            return _to_numpy(self._select([item])[0])
        return self.with_obs(item)

    def _select(self, obs):
        if obs is None:
            return list(self._cols)
        obs = convert_to_obs_str(obs)
        missing = [o for o in obs if o not in self.obs]
        if missing:
            msg = f"observables {missing} not in data obs {self.obs}"
            raise ObsIncompatibleError(msg)
        return [self._cols[self.obs.index(o)] for o in obs]

    def with_obs(self, obs, *, guarantee_limits=True):
        """Column select / reorder (``data.py:876-904``); shares the columns."""
        if isinstance(obs, Space):
            space = obs
            names = obs.obs
        else:
            names = convert_to_obs_str(obs)
            space = self._space.with_obs(names)
        new = object.__new__(Data)
        new._cols = self._select(names)
        new._weights = self._weights
        new._space = space
        new.name, new.label = self.name, self.label
        new._dev_cache = None
        new._version = 0
        if isinstance(obs, Space) and obs.has_limits and obs != self._space.with_obs(names) and not guarantee_limits:
            new._cols, new._weights = _cut(new._cols, new._weights, obs)
        return new

    def with_weights(self, weights):
        new = object.__new__(Data)
        new._cols, new._space, new.name, new.label = self._cols, self._space, self.name, self.label
        new._weights = None if weights is None else Data._ingest(np.zeros((0, self.n_obs)), self._space, weights)[1]
        if new._weights is not None and int(new._weights.shape[0]) != self.num_entries:
            msg = "Weights have to have the same length as the data"
            raise ValueError(msg)
        new._dev_cache = None
        new._version = 0
        return new

    def to_pandas(self, obs=None, weightsname=None):
        import pandas as pd

        names = self.obs if obs is None else convert_to_obs_str(obs)
        df = pd.DataFrame({o: _to_numpy(c) for o, c in zip(names, self._select(names))})
        if self.has_weights:
            df[weightsname or ""] = self.weights
        return df

    # ------------------------------------------------------------------ device residency
    def device_columns(self, device: int = 0):
        """(list of torch CUDA fp64 columns, weights | None) -- resident for the life of the Data."""
        if self._dev_cache is not None and self._dev_cache[0] == device:
            return self._dev_cache[1], self._dev_cache[2]
        import torch

        dev = torch.device("cuda", device)

        def up(c):
            if _is_torch(c):
                # a pinned host tensor goes straight through the copy engine, asynchronously on torch's current stream
                # (the kernels are launched on the same stream); anything else takes torch's plain copy
                t = c.to(dev, non_blocking=c.device.type == "cpu" and _is_pinned(c)).contiguous()
            else:
                t = _upload_numpy(np.ascontiguousarray(c), dev)
            if t.data_ptr() % 16:  # e.g. a view starting at an odd element: the kernel reads double2
                t = t.clone()
            return t

        cols = [up(c) for c in self._cols]
        w = None if self._weights is None else up(self._weights)
        self._dev_cache = (device, cols, w)
        return cols, w

    def pinned_host_columns(self):
        """(columns, weights | None) as torch CPU tensors when EVERY column lives in page-locked host memory and no device
        copy exists yet, else None: the engine then streams them to the GPU in chunks and partitions each chunk while the
        next one is on the wire (``CudaEngine._stream_and_partition``)."""
        if self._dev_cache is not None:
            return None
        import torch

        out = []
        for c in [*self._cols, *([] if self._weights is None else [self._weights])]:
            t = c if _is_torch(c) else torch.from_numpy(np.ascontiguousarray(c))
            if t.device.type != "cpu" or t.dtype != torch.float64 or not t.is_contiguous() or not _is_pinned(t):
                return None
            out.append(t)
        if self._weights is None:
            return out, None
        return out[:-1], out[-1]

    def __repr__(self):
        return f"<zfit.Data: {self.name} obs={self.obs} shape=({self.num_entries}, {self.n_obs})>"


_STAGE = {}  # device index -> (pinned staging buffers, their last-use events)
_STAGE_ELEMS = 2 << 20  # 16 MB of fp64 per buffer
_STAGE_BUFS = 4


def _copy_threads():
    """Host threads for filling the staging buffers: the cores this process may use, shared among the ranks of the node."""
    import os

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(16, cores // local_world))


def _is_pinned(t) -> bool:
    try:
        return bool(t.is_pinned())
    except Exception:  # no CUDA context yet / CPU-only build
        return False


def _upload_numpy(a, dev):
    """Pageable host column -> HBM through a ring of pinned staging buffers: host threads fill one buffer while the
    copy engine drains the others (measured on the B200 box: 800 MB in 16 ms = 51 GB/s, against 70 ms for a plain
    pageable ``.to(device)``).  Small columns take the plain path."""
    import torch

    src = torch.from_numpy(a)
    n = src.numel()
    if n < 2 * _STAGE_ELEMS or src.dtype != torch.float64:
        return src.to(dev)
    if _is_pinned(src):  # page-locked already (cudaHostAlloc / cudaHostRegister): the copy engine reads it directly
        return src.to(dev, non_blocking=True)
    key = dev.index or 0
    if key not in _STAGE:
        _STAGE[key] = ([torch.empty(_STAGE_ELEMS, dtype=torch.float64).pin_memory() for _ in range(_STAGE_BUFS)],
                       [None] * _STAGE_BUFS)
    bufs, events = _STAGE[key]
    out = torch.empty(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    # the fill is torch's parallel CPU copy; torchrun exports OMP_NUM_THREADS=1, which would leave one thread to move
    # every rank's gigabytes (measured 0.46 s instead of 0.05 s for 2 GB), so the thread count is raised for the upload
    prev = torch.get_num_threads()
    want = _copy_threads()
    if want != prev:
        torch.set_num_threads(want)
    try:
        for k, off in enumerate(range(0, n, _STAGE_ELEMS)):
            m = min(_STAGE_ELEMS, n - off)
            b = k % _STAGE_BUFS
            if events[b] is not None:
                events[b].synchronize()
            bufs[b][:m].copy_(src[off : off + m])
            out[off : off + m].copy_(bufs[b][:m], non_blocking=True)
            events[b] = torch.cuda.Event()
            events[b].record(stream)
    finally:
        if want != prev:
            torch.set_num_threads(prev)
    return out


class SamplerData(Data):
    """``SamplerData`` (``data.py:1247-1606``): a Data whose events are drawn from a pdf and re-drawn in place by
    ``resample``.  The parameter values used for sampling are those at creation (``fixed_params``) unless overridden per
    call.  Events are generated, accepted and kept on the GPU; a loss built on it re-binds without any host copy."""

    def __init__(self, sampler, n, *, fixed_params, obs, name=None, label=None):
        self._sampler = sampler
        self._n = n
        self.params = dict(fixed_params)
        super().__init__(sampler.draw(n, self.params), obs=obs, name=name or "SamplerData", label=label, guarantee_limits=True)

    @property
    def n(self):
        return self._n

    def resample(self, params=None, *, n=None, param_values=None):
        if param_values is not None:
            if params is not None:
                msg = "Cannot specify both `fixed_params` and `params`."
                raise ValueError(msg)
            params = param_values
        values = dict(self.params)
        if params is not None:
            byname = {p.name: p for p in values}
            for k, v in dict(params).items():
                key = byname.get(k, k) if isinstance(k, str) else k
                values[key] = float(v.value() if hasattr(v, "value") and callable(v.value) else v)
        self.update_data(self._sampler.draw(self._n if n is None else n, values), guarantee_limits=True)


def _as_col(x):
    if _is_torch(x):
        import torch

        return x.detach().to(torch.float64).contiguous()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _to_numpy(c):
    if _is_torch(c):
        return c.detach().cpu().numpy()
    return c


def _upload_for_cut(cols, weights):
    """Large host columns that still need their cut: move them to HBM first (pinned staging ring, ~50 GB/s) and cut
    there -- three numpy passes per observable over 1e8 events cost ~0.3 s on the host, the device pass ~1 ms, and the
    loss binds the resident columns without another copy.  Single-process only: with one process per GPU every rank
    uploads just its own shard of the host columns at bind time (``distributed.shard``), so the cut stays on the host.
    Returns (device columns, device weights) or None when the cut should run on the host."""
    from . import distributed as _dist
    from . import settings

    threshold = settings.options.get("device_cut_min_events")
    import os

    multi = _dist.is_distributed() or int(os.environ.get("WORLD_SIZE", "1") or 1) > 1  # torchrun, group not up yet
    if threshold is None or int(cols[0].shape[0]) < int(threshold) or multi:
        return None
    try:
        import torch

        if not torch.cuda.is_available():
            return None
    except Exception:
        return None
    dev = torch.device("cuda", _dist.local_device())

    def up(c):
        if _is_torch(c):
            return c.to(dev, non_blocking=_is_pinned(c))
        return _upload_numpy(np.ascontiguousarray(c, dtype=np.float64), dev)

    return [up(c) for c in cols], (None if weights is None else up(weights))


def _cut_on_device(cols, weights, lo, up):
    """Columns that already live in HBM are cut there: one predicate + stable compaction pass of libznll
    (``znll_cut_events``, csrc/znll_prep.cu); only the survivor count comes back to the host."""
    import torch

    from . import _znll as Z

    n = int(cols[0].shape[0])
    if n == 0:
        return cols, weights
    dev = cols[0].device
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    cols = [c.contiguous() for c in cols]
    w = None if weights is None else weights.to(dev).contiguous()
    with torch.cuda.device(index):
        out = [torch.empty_like(c) for c in cols]
        out_w = None if w is None else torch.empty_like(w)
        stream = torch.cuda.current_stream(index).cuda_stream
        kept = Z.cut_events(index, [c.data_ptr() for c in cols], [c.data_ptr() for c in out], 0 if w is None else w.data_ptr(),
                            0 if out_w is None else out_w.data_ptr(), [float(v) for v in lo], [float(v) for v in up], n,
                            stream=stream)
    if kept == n:
        return cols, weights
    # a view of the compacted buffer; when most events went, give the memory back
    shrink = (lambda t: t[:kept].clone()) if kept < n // 2 else (lambda t: t[:kept])
    return [shrink(c) for c in out], (None if out_w is None else shrink(out_w))


def _cut(cols, weights, space: Space):
    lo, up = space._lower, space._upper
    if _is_torch(cols[0]) and cols[0].is_cuda:
        return _cut_on_device(cols, weights, lo, up)
    moved = _upload_for_cut(cols, weights)
    if moved is not None:
        return _cut_on_device(moved[0], moved[1], lo, up)
    if _is_torch(cols[0]):
        inside = None
        for i, c in enumerate(cols):
            m = (c >= float(lo[i])) & (c <= float(up[i]))
            inside = m if inside is None else inside & m
        if bool(inside.all()):
            return cols, weights
        cols = [c[inside].contiguous() for c in cols]
        if weights is not None:
            weights = weights[inside].contiguous()
        return cols, weights
    inside = np.ones(cols[0].shape[0], dtype=bool)
    for i, c in enumerate(cols):
        inside &= (c >= lo[i]) & (c <= up[i])
    if inside.all():
        return cols, weights
    cols = [np.ascontiguousarray(c[inside]) for c in cols]
    if weights is not None:
        weights = np.ascontiguousarray(weights[inside])
    return cols, weights


def convert_to_data(data, obs=None, *, check_limits=False) -> Data:
    """``convert_to_data`` (``data.py:48-81``).  With ``check_limits`` raw arrays must lie inside ``obs``."""
    if isinstance(data, Data):
        return data
    if obs is None:
        msg = "obs is required to convert raw arrays to Data"
        raise ValueError(msg)
    space = obs if isinstance(obs, Space) else Space(obs)
    if check_limits and space.has_limits:
        raw = Data(data, obs=space, guarantee_limits=True)
        cut = Data(data, obs=space)
        if cut.num_entries != raw.num_entries:
            msg = f"Data is not fully within the limits of {space}"
            raise OutsideLimitsError(msg)
        return cut
    return Data(data, obs=space)


def concat(datasets, *, obs=None, axis=None, name=None, label=None):
    """``zfit.data.concat`` along the event axis (``data.py:1609``)."""
    datasets = list(datasets)
    if axis in (1, "obs"):
        space = None
        cols = {}
        for d in datasets:
            for o, c in zip(d.obs, d._cols):
                cols[o] = c
            space = d.space if space is None else space * d.space
        return Data(cols, obs=space, weights=datasets[0]._weights, name=name, label=label, guarantee_limits=True)
    space = datasets[0].space if obs is None else (obs if isinstance(obs, Space) else datasets[0].space.with_obs(obs))
    cols = [np.concatenate([_to_numpy(d._select(space.obs)[i]) for d in datasets]) for i in range(space.n_obs)]
    if any(d.has_weights for d in datasets):
        w = np.concatenate([d.weights if d.has_weights else np.ones(d.num_entries) for d in datasets])
    else:
        w = None
    return Data(dict(zip(space.obs, cols)), obs=space, weights=w, name=name, label=label)
