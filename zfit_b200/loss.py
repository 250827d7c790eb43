"""``zfit.loss.UnbinnedNLL`` / ``ExtendedUnbinnedNLL`` on the fused CUDA kernel (``src/zfit/core/loss.py``).

The minimizer-facing contract of ``BaseLoss`` is kept verbatim (``value``, ``gradient``,
``value_gradient``, ``hessian``, ``value_gradient_hessian``, ``__call__``, ``get_params``, ``errordef``,
``model`` / ``data`` / ``fit_range`` / ``constraints``, ``__add__``, ``create_new``).
What changes is *how* a value + gradient is produced: instead of tracing ~25 TensorFlow ops per
event and replaying them on a tape (``loss.py:73-142``, ``z/math.py:236-261``), one call

  1. reads the current parameter values into the model trees' *slots*,
  2. runs ONE fused kernel per channel (``znll_eval``: per-event composite pdf, analytic slot
     derivatives, normalisation, log, weights, deterministic fp64 reduction),
  3. applies d(slot)/d(theta) on the host (shared parameters, fractions from yields, user
     ``ComposedParameter``), adds the extended Poisson term and the constraints.

There is no CPU fallback: without the CUDA library / a GPU, evaluating raises.
"""

from __future__ import annotations

import copy
import math
import warnings
from collections.abc import Iterable, Mapping

import numpy as np

from . import distributed as _dist
from .data import Data, convert_to_data
from .exception import (
    BreakingAPIChangeError,
    IntentionAmbiguousError,
    NotExtendedPDFError,
    OutsideLimitsError,
    ParamNameNotUniqueError,
)
from .parameter import OrderedSet, ZfitParameter, set_values
from .pdf import BasePDF, CrystalBall, DoubleCB, GeneralizedCB
from .space import Space

DEFAULT_FULL_ARG = True  # loss.py:70
_PIECEWISE = (CrystalBall, DoubleCB, GeneralizedCB)  # leaves whose kernels branch per warp on the event's side of a threshold


# ------------------------------------------------------------------------------------------------
# device engine
# ------------------------------------------------------------------------------------------------
class CudaEngine:
    """Owns the ``znll_plan`` of a loss: lowered model trees, HBM-resident event columns, launches.

    ``eval(slots, grad, log_offset) -> (values[n_channels], grad_slots[total_slots])`` -- already summed
    over all ranks when ``torch.distributed`` is initialised with more than one rank (one NCCL all-reduce
    of the raw ``sum(1 + n_slots)`` fp64 accumulators per call, issued on the launch stream).
    """

    is_cuda_engine = True  # (test doubles with the same interface lack it: no C-level theta path, no native fit loop)

    def __init__(self, models, datas, fit_ranges, *, device=None, sort_events=True, deterministic_reduce=False,
                 distributed=None, fused_reduce=True):
        from . import _znll as Z

        self.Z = Z
        sort_events = "lazy" if sort_events == "lazy" else bool(sort_events)
        self.distributed = _dist.is_distributed() if distributed is None else bool(distributed)
        self.reduce_mode = "none"
        self.deterministic_reduce = deterministic_reduce
        if device is None:
            device = _dist.local_device()
        self.device = device
        import os as _os
        import time as _time

        trace = [] if _os.environ.get("ZNLL_BIND_TRACE") else None  # stage times of the bind (tools/bind_breakdown.py)

        def mark(name):
            if trace is not None:
                trace.append((name, _time.perf_counter()))

        mark("start")
        self.plan = Z.Plan(len(models), device=device)
        mark("plan_create")
        self.slot_params: list[ZfitParameter] = []
        self.channel_slots: list[tuple[int, int]] = []
        self._keep = []
        self.n_local = []
        import torch

        self.torch = torch
        for k, (model, data, rng) in enumerate(zip(models, datas, fit_ranges)):
            if rng is not None:
                data = data.with_obs(rng, guarantee_limits=False)
            nodes, slots = model._lower(data.obs, rng)
            mark("lower")
            self.plan.set_model(k, nodes)
            mark("set_model")
            start = len(self.slot_params)
            self.slot_params.extend(slots)
            self.channel_slots.append((start, len(self.slot_params)))
            n = data.num_entries
            streamed = self._stream_and_partition(model, data) if sort_events is True else None
            if streamed is not None:
                cols, w = streamed
                mark("stream+partition")
            else:
                cols, w = data.device_columns(device)
                mark("device_columns")
                if sort_events is True and n > 1:
                    cols, w = self._sort_for_coherence(model, data, cols, w)
            self._keep.append((cols, w))
            mark("partition(enqueue)")
            # the columns were produced by torch kernels on torch's stream (upload, cut, sort); the plan launches
            # on its own stream, so make them visible first
            torch.cuda.synchronize(device)
            mark("sync")
            self.plan.bind_device_ptrs(k, [c.data_ptr() for c in cols], 0 if w is None else w.data_ptr(), n)
            mark("bind+sumw")
            self.n_local.append(n)
        # sort_events="lazy" (experimental, not the default): bind unsorted and sort only once the loss has been evaluated
        # LAZY_SORT_AFTER times -- a bind-time sort of 1e8 events costs about as much as a whole fit's kernel time, so a
        # loss that is fitted once should not pay it, one that is profiled or refitted should (ski-rental rule)
        self._lazy_sort_in = self.LAZY_SORT_AFTER if sort_events == "lazy" else None
        self.samplesize = [self.plan.sumw(k) for k in range(len(models))]
        self.n_entries = list(self.n_local)
        self._bind_args = (list(models), list(fit_ranges), sort_events)
        self._bound_datas = list(datas)
        if self.distributed:
            t = torch.tensor(self.samplesize + [float(n) for n in self.n_local], dtype=torch.float64, device=f"cuda:{device}")
            _dist.all_reduce_sum_(t, deterministic=True)
            t = t.cpu().numpy()
            nch = len(models)
            self.samplesize = [float(v) for v in t[:nch]]
            self.n_entries = [int(round(v)) for v in t[nch:]]
            for k in range(nch):
                self.plan.set_sumw(k, self.samplesize[k])
            self.reduce_mode = "nccl"
            if fused_reduce:
                try:
                    self._setup_fused_exchange(device)
                    self.reduce_mode = "fused-nvlink"
                except Exception as exc:  # symmetric memory unavailable: keep the NCCL all-reduce, and say so
                    self.reduce_fallback_reason = repr(exc)
                    self.plan.set_peers(0, 1, [], 0)
                    warnings.warn(
                        "zfit_b200: the fused NVLink exchange is unavailable on this rank, the cross-GPU reduction uses "
                        f"an NCCL all-reduce per evaluation instead ({exc!r})", RuntimeWarning, stacklevel=2)
            if self.reduce_mode == "nccl":
                self._acc = _dist.device_view(self.plan.acc_device_ptr(), self.plan.total_acc(), device)
            # the passes the kernel does not reduce itself (information matrix); a weak reference: the plan must not keep
            # its engine alive through the callback
            import weakref

            me = weakref.ref(self)
            self.plan.set_reduce_hook(lambda buf: me()._sum_over_ranks(buf))
        self.kernel_names = [self.plan.kernel_name(k) for k in range(len(models))]
        self._raw_stream = self._pick_stream_getter()
        mark("rest")
        if trace is not None:
            print("bind trace [ms]: " + " | ".join(f"{n} {1e3 * (t - p):.2f}" for (n, t), (_, p) in zip(trace[1:], trace[:-1])), flush=True)

    def _sum_over_ranks(self, buf):
        """In-place world sum of a small host float64 array through the communicator (all-gather + fixed-order sum:
        identical on every rank) -- what ``libznll`` calls for its rank-local passes (``znll_plan_set_reduce_hook``)."""
        t = self.torch.from_numpy(buf).to(f"cuda:{self.device}")
        _dist.all_reduce_sum_(t, deterministic=True)
        buf[:] = t.cpu().numpy()

    def _pick_stream_getter(self):
        """torch's current stream on this device as a raw ``cudaStream_t``.  ``torch.cuda.current_stream(d).cuda_stream``
        builds a Stream object per call (~2 us on the step path); torch's own raw getter (the one its compiled graphs
        use) returns the handle directly.  It is used only if it exists and agrees with the public API right now."""
        torch, dev = self.torch, self.device

        def public():
            return torch.cuda.current_stream(dev).cuda_stream

        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        try:
            if raw is not None and int(raw(dev)) == int(public()):
                return lambda: raw(dev)
        except Exception:
            pass
        return public

    LAZY_SORT_AFTER = 48

    def _sort_now(self):
        """Deferred bind-time sort: reorder the resident columns and bind them again (same kernels, same plan)."""
        models, fit_ranges, _ = self._bind_args
        torch = self.torch
        for k, (model, data, rng) in enumerate(zip(models, self._bound_datas, fit_ranges)):
            cols, w = self._keep[k]
            n = self.n_local[k]
            if n > 1:
                if rng is not None:
                    data = data.with_obs(rng, guarantee_limits=False)
                cols, w = self._sort_for_coherence(model, data, cols, w)
                self._keep[k] = (cols, w)
                torch.cuda.synchronize(self.device)
                self.plan.bind_device_ptrs(k, [c.data_ptr() for c in cols], 0 if w is None else w.data_ptr(), n)
                if self.distributed:
                    self.plan.set_sumw(k, self.samplesize[k])  # binding recomputed the local sum; the global one holds

    def rebind(self, datas):
        """The events of the bound ``Data`` objects changed in place (``SamplerData.resample``): bind the new columns to
        the existing plan -- same compiled kernels, no host copy when the data was born on the device."""
        if self.distributed:
            msg = "in-place resampling of a sharded loss is not supported"
            raise NotImplementedError(msg)
        models, fit_ranges, sort_events = self._bind_args
        torch = self.torch
        self._keep = []
        for k, (model, data, rng) in enumerate(zip(models, datas, fit_ranges)):
            if rng is not None:
                data = data.with_obs(rng, guarantee_limits=False)
            cols, w = data.device_columns(self.device)
            n = data.num_entries
            if sort_events is True and n > 1:
                cols, w = self._sort_for_coherence(model, data, cols, w)
            self._keep.append((cols, w))
            torch.cuda.synchronize(self.device)
            self.plan.bind_device_ptrs(k, [c.data_ptr() for c in cols], 0 if w is None else w.data_ptr(), n)
            self.n_local[k] = n
        self.samplesize = [self.plan.sumw(k) for k in range(len(models))]
        self.n_entries = list(self.n_local)
        self._bound_datas = list(datas)
        self._lazy_sort_in = self.LAZY_SORT_AFTER if sort_events == "lazy" else None

    def _sort_for_coherence(self, model, data, cols, w):
        """Events are static for the whole fit and the NLL is a sum, so they may be reordered once at bind time: a stable
        partition into 32 buckets of the CrystalBall observable (``znll_partition_events``, csrc/znll_prep.cu) makes
        every warp uniformly core or tail -- all but those of the one bucket that holds the moving threshold.  Costs
        about two evaluations' worth of memory traffic (a full ``torch.sort`` cost 25 evaluations' worth of time)."""
        target = _first_leaf(model, _PIECEWISE)
        if target is None:
            return cols, w
        torch = self.torch
        i = data.obs.index(target.obs[0])
        lo, hi = data.space.with_obs(target.obs[0]).limit1d if data.space.has_limits else (None, None)
        if lo is None or not (hi > lo):
            lo, hi = float(cols[i].min()), float(cols[i].max())
            if not (hi > lo):
                return cols, w
        n = int(cols[0].numel())
        out = [torch.empty_like(c) for c in cols]
        wout = None if w is None else torch.empty_like(w)
        self.Z.partition_events(self.device, [c.data_ptr() for c in cols], [c.data_ptr() for c in out],
                                0 if w is None else w.data_ptr(), 0 if wout is None else wout.data_ptr(), i, float(lo),
                                float(hi), n, stream=self._raw_stream_now())
        return out, wout

    def _raw_stream_now(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    STREAM_CHUNK = 1 << 23  # events per chunk of the pipelined upload (64 MB per column)

    def _stream_and_partition(self, model, data):
        """Upload and partition in ONE pipeline when the events still sit in pinned host memory: the columns cross PCIe
        in chunks of ``STREAM_CHUNK`` events on a copy stream, and every chunk is bucketed
        (``znll_partition_events``) on the compute stream while the next one is on the wire, straight into its slice of
        the final columns.  Warp coherence only needs LOCAL order, so the buckets are per chunk; the whole bind then
        costs the PCIe transfer plus one chunk's partition instead of transfer + a pass over everything, and the
        unpartitioned copy never exists in HBM.  Returns None when the case does not apply (no piecewise leaf, data
        already on the device or in pageable memory, a small sample)."""
        target = _first_leaf(model, _PIECEWISE)
        n = data.num_entries
        if target is None or n < 2 * self.STREAM_CHUNK or not data.space.has_limits:
            return None
        host = data.pinned_host_columns()
        if host is None:
            return None
        hcols, hw = host
        torch, dev = self.torch, self.torch.device("cuda", self.device)
        key = data.obs.index(target.obs[0])
        lo, hi = data.space.with_obs(target.obs[0]).limit1d
        if not (hi > lo):
            return None
        streams = [*hcols, *([] if hw is None else [hw])]
        out = [torch.empty(n, dtype=torch.float64, device=dev) for _ in streams]
        chunk = self.STREAM_CHUNK
        stage = [[torch.empty(chunk, dtype=torch.float64, device=dev) for _ in streams] for _ in range(2)]
        compute = torch.cuda.current_stream(self.device)
        copy = torch.cuda.Stream(device=self.device)
        free = [None, None]
        copy.wait_stream(compute)
        nc = len(hcols)
        for k, start in enumerate(range(0, n, chunk)):
            m = min(chunk, n - start)
            b = k & 1
            with torch.cuda.stream(copy):
                if free[b] is not None:
                    copy.wait_event(free[b])  # the partition that read this staging buffer two chunks ago is done
                for s, h in zip(stage[b], streams):
                    s[:m].copy_(h[start : start + m], non_blocking=True)
                arrived = torch.cuda.Event()
                arrived.record(copy)
            compute.wait_event(arrived)
            self.Z.partition_events(self.device, [s.data_ptr() for s in stage[b][:nc]],
                                    [o.data_ptr() + 8 * start for o in out[:nc]],
                                    0 if hw is None else stage[b][nc].data_ptr(),
                                    0 if hw is None else out[nc].data_ptr() + 8 * start, key, float(lo), float(hi), m,
                                    stream=compute.cuda_stream)
            free[b] = torch.cuda.Event()
            free[b].record(compute)
        for s in stage[0] + stage[1]:
            s.record_stream(compute)
        return out[:nc], (None if hw is None else out[nc])

    def _setup_fused_exchange(self, device):
        """Peer-mapped exchange buffer through torch's symmetric memory (allocation + handle exchange only); the
        reduction itself runs inside the NLL kernel (csrc/znll_device.cuh::peer_exchange).  ONE buffer per process and
        device serves every plan (it is sized for the largest accumulator block the kernels support and the evaluation
        sequence number is process-wide), so binding a loss costs no rendezvous after the first."""
        nbytes, ptrs, keep = _exchange_buffer(self.torch, device, self.Z.exchange_bytes_max(_dist.world_size()))
        self.plan.set_peers(_dist.rank(), _dist.world_size(), ptrs, nbytes, keepalive=keep)

    def eval(self, slots, grad=True, log_offset=0.0, kahan=False):
        # kernels go to torch's current stream, so torch ops (NCCL all-reduce, data preparation) order with them
        stream = self._raw_stream()
        if self._lazy_sort_in is not None:
            self._lazy_sort_in -= 1
            if self._lazy_sort_in < 0:
                self._lazy_sort_in = None
                self._sort_now()
        if not self.distributed or self.reduce_mode == "fused-nvlink":
            return self.plan.eval(slots, grad=grad, log_offset=log_offset, kahan=kahan, stream=stream)
        self.plan.launch(slots, grad=grad, log_offset=log_offset, stream=stream)
        _dist.all_reduce_sum_(self._acc, deterministic=self.deterministic_reduce)
        return self.plan.collect(grad=grad, stream=stream)

    def eval_cov(self, slots):
        """Per channel the (n_slots+1)^2 matrix sum_i w_i^2 [g_i;1][g_i;1]^T (g_i = d log pdf_i/d slot), world-summed."""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        _, _, mats = self.plan.eval_cov(slots, stream=stream)
        if self.distributed:
            flat = self.torch.from_numpy(np.concatenate([m.ravel() for m in mats])).to(f"cuda:{self.device}")
            _dist.all_reduce_sum_(flat, deterministic=True)
            flat = flat.cpu().numpy()
            out, off = [], 0
            for m in mats:
                out.append(flat[off : off + m.size].reshape(m.shape).copy())
                off += m.size
            mats = out
        return mats

    def eval_hess(self, slots, log_offset=0.0):
        """(values[n_channels], d value / d slot, [per channel d2 value / d slot^2]) from ONE fused pass per channel
        (``znll_eval_hess``), world-summed; ``None`` when a channel's tree has no one-pass Hessian (not flat, an
        unsupported node, built through NVRTC) -- the loss then differences the analytic gradient."""
        if not self.plan.has_hess():
            return None
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        got = self.plan.eval_hess(slots, log_offset=log_offset, stream=stream)
        if got is None:
            return None
        values, gs, mats = got
        if self.distributed:  # every term of the assembly is linear in the per-rank accumulators and in sum(w)
            flat = self.torch.from_numpy(np.concatenate([values, gs, *[m.ravel() for m in mats]])).to(f"cuda:{self.device}")
            _dist.all_reduce_sum_(flat, deterministic=True)
            flat = flat.cpu().numpy()
            nv, ng = len(values), len(gs)
            values, gs = flat[:nv].copy(), flat[nv : nv + ng].copy()
            out, off = [], nv + ng
            for m in mats:
                out.append(flat[off : off + m.size].reshape(m.shape).copy())
                off += m.size
            mats = out
        return values, gs, mats

    def launch_count(self):
        return self.plan.launch_count()

    def close(self):
        self.plan.close()


_XCHG = {}  # device -> (nbytes, peer pointers, keepalive) | Exception: the process-wide exchange buffer


def _exchange_buffer(torch, device, nbytes):
    """Symmetric-memory exchange buffer of this process on ``device``, created (collectively) on first use.  Whether the
    fused path is usable is agreed on by ALL ranks (a MIN all-reduce of a success flag): a rank-local failure must not
    leave one rank on NCCL while its peers spin in the kernel."""
    import torch.distributed as td

    if device in _XCHG:
        got = _XCHG[device]
        if isinstance(got, Exception):
            raise got
        return got
    err = None
    try:
        import torch.distributed._symmetric_memory as symm

        world = _dist.world_size()
        n = (nbytes + 7) // 8
        buf = symm.empty(n, dtype=torch.float64, device=torch.device("cuda", device))
        buf.zero_()
        handle = symm.rendezvous(buf, td.group.WORLD)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        if len(ptrs) != world:
            msg = f"symmetric memory returned {len(ptrs)} peers for world {world}"
            raise RuntimeError(msg)
        torch.cuda.synchronize(device)
    except Exception as exc:
        err = exc
    ok = torch.tensor([0.0 if err is not None else 1.0], dtype=torch.float64, device=f"cuda:{device}")
    td.all_reduce(ok, op=td.ReduceOp.MIN)  # also the barrier: every rank has zeroed its flags before anyone publishes
    if float(ok.item()) < 1.0:
        _XCHG[device] = err if err is not None else RuntimeError("a peer rank could not set up the symmetric-memory exchange")
        raise _XCHG[device]
    _XCHG[device] = (n * 8, ptrs, (buf, handle))
    return _XCHG[device]


class ThetaPath:
    """The loss in theta space behind the C ABI (``znll_plan_set_theta_map`` / ``znll_eval_theta``): ``indep`` are the
    independent parameters the loss depends on, in the order of the C side's theta vector."""

    def __init__(self, indep, extended):
        self.indep = indep
        self.extended = extended
        self._index_cache = {}

    def index_of(self, params):
        key = tuple(id(p) for p in params)
        got = self._index_cache.get(key)
        if got is None:
            pos = {id(p): i for i, p in enumerate(self.indep)}
            idx = np.array([pos.get(id(p), -1) for p in params], dtype=np.int64)
            got = (idx, bool(np.all(idx >= 0)))
            self._index_cache[key] = got
        return got[0]

    def gather(self, g, params):
        """d value / d params from the gradient over ``indep`` (0 for a parameter the loss does not depend on)."""
        self.index_of(params)
        idx, complete = self._index_cache[tuple(id(p) for p in params)]
        if complete:
            return g[idx]
        return np.where(idx >= 0, g[np.maximum(idx, 0)], 0.0) if len(idx) else np.zeros(0)


def _build_theta_path(loss, eng):
    """Describe every slot (and, for an extended loss, every channel's yield) as a rule over the independent parameters
    -- a plain value, ``yield_i / sum(yields)``, ``1 - sum(fracs)``, a sum of yields -- and hand the map to the plan.
    Returns None when a slot depends on anything else (a user ``ComposedParameter``, renormalised fractions): the loss
    then keeps the Python chain rule."""
    from . import _znll as Z
    from .parameter import ConstantParameter, FractionOfSum, OneMinusSum, SumOfParameters

    indep, refs = [], []

    def ref_of(p):
        if p.independent:
            for i, q in enumerate(indep):
                if q is p:
                    return (1, i, 0.0)
            indep.append(p)
            return (1, len(indep) - 1, 0.0)
        if isinstance(p, ConstantParameter) or not p._leaves():
            return (0, 0, float(p.value()))
        return None

    def rule_of(p):
        if p.independent or isinstance(p, ConstantParameter) or not p._leaves():
            terms, kind = [p], Z.RULE_VALUE
        elif isinstance(p, FractionOfSum) and p._plain:
            terms, kind = [p._term, *p._total._terms], Z.RULE_FRACTION
        elif isinstance(p, OneMinusSum) and p._plain:
            terms, kind = list(p._terms), Z.RULE_REMAINDER
        elif isinstance(p, SumOfParameters) and p._plain:
            terms, kind = list(p._terms), Z.RULE_SUM
        else:
            return None
        first = len(refs)
        for t in terms:
            r = ref_of(t)
            if r is None:
                return None
            refs.append(r)
        return (kind, first, len(terms))

    slot_rules = []
    for sp in eng.slot_params:
        ru = rule_of(sp)
        if ru is None:
            return None
        slot_rules.append(ru)
    yield_rules = None
    if loss.is_extended:
        yield_rules = []
        for mod in loss.model:
            if not mod.is_extended:
                return None
            ru = rule_of(mod.get_yield())
            if ru is None:
                return None
            yield_rules.append(ru)
    n = len(indep)
    # theta arrives clipped (Parameter.value): no bounds on the C side, so a limit changed later cannot go stale there
    eng.plan.set_theta_map(n, np.full(n, -np.inf), np.full(n, np.inf), refs, slot_rules, yield_rules)
    return ThetaPath(indep, yield_rules is not None)


def _first_leaf(model, cls):
    if isinstance(model, cls):
        return model
    for m in getattr(model, "pdfs", []):
        r = _first_leaf(m, cls)
        if r is not None:
            return r
    return None


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def _to_list(x):
    if x is None:
        return []
    if isinstance(x, (list,)):
        return list(x)
    return [x]


class BaseLoss:
    _name = "BaseLoss"

    def __init__(self, model, data, fit_range=None, constraints=None, options=None):
        if fit_range is not None and all(fr is not None for fr in _to_list(fit_range)):
            warnings.warn(
                "The fit_range argument is depreceated and will maybe removed in future releases. "
                "It is preferred to define the range in the space"
                " when creating the data and the model or directly cut the data correctly.",
                stacklevel=2,
            )
        model, data, fit_range = self._input_check(model, data, fit_range)
        self._model = model
        self._data = data
        self._fit_range = fit_range
        self._options = self._check_init_options(options, data)
        self._constraints = _constraint_check_convert(_to_list(constraints))
        self._errordef = None
        self._engine = None
        self._engine_factory = None
        self._data_hashes = None
        self._all_params = None
        self._slot_plan = None
        self._assert_params_unique()

    # ------------------------------------------------------------------ init helpers
    def _check_init_options(self, options, data):  # loss.py:264-290
        nevents = sum(d.num_entries for d in data)
        options = {} if options is None else copy.copy(dict(options))
        clean = copy.copy(options)
        if options.pop("numhess", None) is None:
            clean["numhess"] = True
        if options.pop("numgrad", None) is None:
            from . import settings

            clean["numgrad"] = settings.options["numerical_grad"]
        if options.pop("subtr_const", None) is None:
            clean["subtr_const"] = True
        if options.pop("sumtype", None) is None:
            clean["sumtype"] = "kahan" if nevents > 200_000 else None
        return clean

    def _input_check(self, pdf, data, fit_range):  # loss.py:332-381, 413-480
        if isinstance(pdf, tuple):
            msg = "`pdf` has to be a pdf or a list of pdfs, not a tuple."
            raise TypeError(msg)
        if isinstance(data, tuple):
            msg = "`data` has to be a data or a list of data, not a tuple."
            raise TypeError(msg)
        pdf, data = _to_list(pdf), _to_list(data)
        is_simple = isinstance(self, SimpleLoss)
        if not pdf and not is_simple:
            msg = "At least one model must be provided to create a loss."
            raise ValueError(msg)
        if not data and not is_simple:
            msg = "At least one dataset must be provided to create a loss."
            raise ValueError(msg)
        if len(pdf) != len(data):
            msg = f"pdf, data don't have the same number of components:\npdf: {pdf}\ndata: {data}"
            raise ValueError(msg)
        pdf_checked, data_checked = [], []
        for i, (mod, dat) in enumerate(zip(pdf, data)):
            if not isinstance(mod, BasePDF):
                msg = f"Model at index {i} must be a ZfitPDF, got {type(mod).__name__}"
                raise TypeError(msg)
            if not isinstance(dat, Data):
                if fit_range is not None:
                    msg = "Fit range should not be used if data is not ZfitData."
                    raise TypeError(msg)
                try:
                    dat = convert_to_data(dat, obs=mod.space, check_limits=True)
                except OutsideLimitsError as error:
                    msg = (
                        f"Data {dat} is not a zfit Data (and therefore has no Space that defines possible limits) "
                        f"and is not fully within the limits {mod.space} of the model {mod}."
                        f"If the data should be what it is, please convert to zfit Data (`zfit.Data(data, obs=obs)`) "
                        f"or remove events outside the space"
                    )
                    raise IntentionAmbiguousError(msg) from error
            if dat.num_entries == 0 and not _dist.is_distributed():
                msg = f"Dataset at index {i} is empty (has 0 entries). Cannot create loss with empty data."
                raise ValueError(msg)
            if mod.space.obs != dat.space.obs:
                if set(mod.space.obs) <= set(dat.space.obs) and len(mod.space.obs) == len(dat.space.obs):
                    # same observables in another order: keep the caller's Data object (a reordered copy would not see
                    # SamplerData.resample / update_data on the original); the engine addresses columns by NAME at bind
                    # time (model._lower(data.obs, ...)), so no reordering is needed
                    pass
                else:
                    msg = (
                        f"Model at index {i} has observables {mod.space.obs} but data has "
                        f"observables {dat.space.obs}. The observables must match."
                    )
                    raise ValueError(msg)
            pdf_checked.append(mod)
            data_checked.append(dat)
        if fit_range is None:
            fit_range = [None] * len(pdf_checked)
        else:
            fit_range = fit_range if isinstance(fit_range, list) else [fit_range]
            if len(fit_range) != len(pdf_checked):
                msg = "pdf, data and fit_range don't have the same number of components"
                raise ValueError(msg)
            fit_range = [
                None if r is None else (r if isinstance(r, Space) else p.space.with_limits(r))
                for p, r in zip(pdf_checked, fit_range)
            ]
        return pdf_checked, data_checked, fit_range

    def _assert_params_unique(self):
        seen = {}
        for p in self.get_params(floating=None, is_yield=None, extract_independent=True):
            if p.name in seen and seen[p.name] is not p:
                msg = f"The following parameter names appear more than once in {self}: {p.name}."
                raise ParamNameNotUniqueError(msg)
            seen[p.name] = p

    # ------------------------------------------------------------------ description
    @property
    def name(self):
        return type(self).__name__

    @property
    def model(self):
        return self._model

    @property
    def data(self):
        return self._data

    @property
    def fit_range(self):
        return self._fit_range

    @property
    def constraints(self):
        return self._constraints

    @property
    def errordef(self):
        return self._errordef

    @property
    def is_weighted(self) -> bool:
        return any(d.has_weights for d in self.data)

    @property
    def is_extended(self) -> bool:
        return False

    @property
    def options(self):
        return self._options

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        """OrderedSet union over the models in list order, then constraints (``loss.py:296-321``);
        the order defines the gradient layout."""
        params = OrderedSet()
        for m in self.model:
            params = params.union(m.get_params(floating=floating, is_yield=is_yield, extract_independent=extract_independent))
        for c in self.constraints:
            params = params.union(c.get_params(floating=floating, is_yield=False, extract_independent=extract_independent))
        return params

    # ------------------------------------------------------------------ engine
    def _check_data_changed(self):
        """The events of a bound Data changed in place (``SamplerData.resample``): re-bind them, and the automatic offset
        is stale (``is_precompiled``, ``loss.py:247-262``)."""
        hashes = [d.hashint for d in self.data]
        if self._engine is not None and hashes != self._data_hashes:
            if hasattr(self._engine, "rebind"):
                self._engine.rebind(self.data)
            else:
                self._engine = None
            if self._options.get("subtr_const") is True:
                self._options["subtr_const_value"] = None
            self._options["_precompiled"] = False
        self._data_hashes = hashes

    def _get_engine(self, check=True):
        if check:  # the evaluation entry points have already checked (check_precompile): their inner calls skip it
            self._check_data_changed()
        if self._engine is None:
            if self._engine_factory is not None:
                self._engine = self._engine_factory(self.model, self.data, self.fit_range)
            else:
                self._engine = CudaEngine(
                    self.model,
                    self.data,
                    self.fit_range,
                    sort_events=self._options.get("sort_events", True),
                    deterministic_reduce=self._options.get("deterministic_reduce", False),
                    distributed=self._options.get("distributed"),
                    fused_reduce=self._options.get("fused_reduce", True),
                )
        return self._engine

    def _set_engine_factory(self, factory):
        """Test seam: ``factory(models, datas, fit_ranges) -> engine`` with the ``CudaEngine`` interface."""
        self._engine_factory = factory
        self._engine = None

    # ------------------------------------------------------------------ core evaluation
    def _nll_channels(self, want_grad, log_offset):
        eng = self._get_engine(check=False)
        if self._slot_plan is None or self._slot_plan[0] is not eng:
            sp = eng.slot_params
            direct = [(j, p) for j, p in enumerate(sp) if p.independent]
            composed = [(j, p) for j, p in enumerate(sp) if not p.independent and p._leaves()]
            self._slot_plan = (eng, direct, composed, np.empty(len(sp), dtype=np.float64))
        _, direct, composed, slots = self._slot_plan
        for j, p in enumerate(eng.slot_params):
            slots[j] = p.value()
        # the kernels always accumulate with error-free merges of the threads' partial sums; the reference's sumtype
        # switch (loss.py:138-141) is forwarded for interface parity only
        kahan = self._options.get("sumtype") == "kahan"
        values, gslots = eng.eval(slots, grad=want_grad, log_offset=0.0 if log_offset is False else float(log_offset), kahan=kahan)
        total = float(values.sum()) if len(values) > 1 else float(values[0])
        grad = {}
        if want_grad:
            for j, p in direct:
                grad[p] = grad.get(p, 0.0) + gslots[j]
            memo = {}
            for j, p in composed:
                g = gslots[j]
                for leaf, d in p._partials(memo).items():
                    grad[leaf] = grad.get(leaf, 0.0) + g * d
        return total, grad

    def _extra_terms(self, want_grad, log_offset):
        """Terms beyond the per-event sum (extended Poisson term); overridden by ExtendedUnbinnedNLL."""
        return 0.0, {}

    def _theta_path(self, eng):
        """The C-level theta path for this engine (built once), or None: only the CUDA engine has one, a sharded loss only
        with the fused exchange (the NCCL route reduces between launch and collect), and only while every slot follows a
        built-in rule; ``options={"theta_path": False}`` keeps the Python chain rule (the parity tests compare the two)."""
        cached = getattr(self, "_theta_cache", None)
        if cached is not None and cached[0] is eng:
            return cached[1]
        path = None
        if (getattr(eng, "is_cuda_engine", False) and self._options.get("theta_path", True)
                and (not eng.distributed or eng.reduce_mode == "fused-nvlink")):
            path = _build_theta_path(self, eng)
        self._theta_cache = (eng, path)
        return path

    def _evaluate(self, params, want_grad, log_offset):
        eng = self._get_engine(check=False)
        path = self._theta_path(eng)
        if path is not None and getattr(eng, "_lazy_sort_in", None) is None:
            # slots from theta, the fused pass, the chain rule and the extended term in ONE native call
            full = log_offset is False
            theta = [p.value() for p in path.indep]
            value, g = eng.plan.eval_theta(theta, grad=want_grad, full=full, log_offset=0.0 if full else float(log_offset),
                                           stream=eng._raw_stream())
            grad = None
            if want_grad:
                grad = path.gather(g, params)
            for c in self.constraints:
                value += float(c.value())
                if want_grad:
                    pos = {id(p): i for i, p in enumerate(params)}
                    for k, v in c._partials().items():
                        i = pos.get(id(k))
                        if i is not None:
                            grad[i] += v
            return value, grad
        value, grad = self._nll_channels(want_grad, log_offset)
        extra, egrad = self._extra_terms(want_grad, log_offset)
        value += extra
        for k, v in egrad.items():
            grad[k] = grad.get(k, 0.0) + v
        for c in self.constraints:
            value += float(c.value())
            if want_grad:
                for k, v in c._partials().items():
                    grad[k] = grad.get(k, 0.0) + v
        if not want_grad:
            return value, None
        return value, np.array([grad.get(p, 0.0) for p in params], dtype=np.float64)

    # ------------------------------------------------------------------ weighted-fit covariance ingredient
    def information_matrix(self, params=None):
        """Observed-information estimate ``sum_i w_i grad(v_i) grad(v_i)^T`` (+ the constraints' outer products): the
        expected Hessian of the NLL at the true parameters, from the same single pass as ``weighted_score_outer``
        (weighted channels are rescaled by ``sum w / sum w^2``).  The native MIGRAD driver seeds its metric with it."""
        return self.weighted_score_outer(params, _unit_scale=True)

    def weighted_score_outer(self, params=None, _unit_scale=False):
        """``C = sum_i w_i^2 grad_theta(v_i) grad_theta(v_i)^T`` with ``v_i = log pdf(x_i) [+ log yield]`` plus one
        unit-weight term per constraint -- the matrix ``covariance_with_weights`` builds from a per-event Jacobian
        (``minimizers/errors.py:406-452``); here it is ONE extra fused pass over the events (``znll_eval_cov``)."""
        params = list(self._input_check_params(params))
        eng = self._get_engine()
        slots = np.array([p.value() for p in eng.slot_params], dtype=np.float64)
        mats = eng.eval_cov(slots)
        index = {id(p): i for i, p in enumerate(params)}
        npar = len(params)
        C = np.zeros((npar, npar))
        for k, (E, (lo, hi)) in enumerate(zip(mats, self._channel_slot_ranges(eng))):
            ns = hi - lo
            B = np.zeros((ns + 1, npar))
            for j, sp in enumerate(eng.slot_params[lo:hi]):
                for leaf, d in (sp._partials().items() if not sp.independent else [(sp, 1.0)]):
                    if id(leaf) in index:
                        B[j, index[id(leaf)]] += d
            if self.is_extended:
                yparam = self.model[k].get_yield()
                y = yparam.value()
                for leaf, d in yparam._partials().items():
                    if id(leaf) in index:
                        B[ns, index[id(leaf)]] += d / y
            if _unit_scale and self.data[k].has_weights and E[ns, ns] > 0:
                E = E * (eng.samplesize[k] / E[ns, ns])  # sum w / sum w^2
            C += B.T @ E @ B
        for con in self.constraints:
            g = np.zeros(npar)
            for leaf, d in con._partials().items():
                if id(leaf) in index:
                    g[index[id(leaf)]] += d
            C += np.outer(g, g)
        return C

    def _channel_slot_ranges(self, eng):
        if hasattr(eng, "channel_slots"):
            return eng.channel_slots
        out, off = [], 0
        for model, data, rng in zip(self.model, self.data, self.fit_range):
            n = len(model._lower(data.obs, rng)[1])
            out.append((off, off + n))
            off += n
        return out

    # ------------------------------------------------------------------ offsets (loss.py:383-411)
    def check_precompile(self, *, params=None, force=False):
        self._check_data_changed()
        if self._options.get("_precompiled") and not force:
            return params, False
        with self._check_set_input_params(params):
            do_subtr = self._options.get("subtr_const", False)
            if do_subtr:
                if do_subtr is not True:
                    self._options["subtr_const_value"] = do_subtr
                if self._options.get("subtr_const_value") is None:
                    eng = self._get_engine()
                    nevents_tot = float(sum(eng.n_entries))
                    allp = list(self.get_params(floating=None))
                    v0, _ = self._evaluate(allp, False, 0.0)
                    self._options["subtr_const_value"] = -((v0 - 10000.0) / nevents_tot)
        self._options["_precompiled"] = True
        return params, True

    def _check_set_input_params(self, params):
        if params is None:
            return _NullCtx()
        if isinstance(params, Mapping):
            byname = {p.name: p for p in self.get_params(floating=None, is_yield=None)}
            keys = [k if isinstance(k, ZfitParameter) else byname[k] for k in params]
            return set_values(keys, list(params.values()))
        if hasattr(params, "params") and hasattr(params, "loss"):  # FitResult
            return set_values(list(params.params), [v["value"] for v in params.params.values()])
        if isinstance(params, Iterable):
            allp = list(self.get_params(floating=None, is_yield=None))
            vals = list(params)
            if len(vals) != len(allp):
                msg = "Length of params does not match the length of all parameters or provide a Mapping."
                raise ValueError(msg)
            return set_values(allp, vals)
        msg = f"cannot interpret params {params}"
        raise TypeError(msg)

    def _all_params_cached(self):
        # membership of a parameter in the loss never changes after construction (only `floating` may)
        if self._all_params is None:
            self._all_params = self.get_params(floating=None, extract_independent=None, is_yield=None).union(
                self.get_params(floating=None, extract_independent=True, is_yield=None)
            )
        return self._all_params

    def _input_check_params(self, params, only_independent=False):
        if params is None:
            return tuple(self.get_params())
        if isinstance(params, ZfitParameter):
            params = (params,)
        params = tuple(params)
        if only_independent and (bad := [p for p in params if not p.independent]):
            msg = f"Parameters {bad} are not independent and `only_independent` is set to True."
            raise ValueError(msg)
        allp = self._all_params_cached()
        if not_in := [p for p in params if p not in allp]:
            msg = f"Parameters {not_in} are not in the model."
            raise ValueError(msg)
        return params

    # ------------------------------------------------------------------ public contract
    def __call__(self, _x=None):
        if isinstance(_x, dict):
            msg = "Dicts are not supported when calling a loss, only array-like values."
            raise TypeError(msg)
        if _x is None:
            return self.value(full=True)
        params = list(self.get_params())
        vals = np.atleast_1d(np.asarray(_x, dtype=np.float64)).reshape(-1)
        if len(vals) != len(params):
            msg = f"Incompatible length of parameters and values: {len(params)} vs {len(vals)}"
            raise ValueError(msg)
        with set_values(params, vals):
            return self.value(full=True)

    def value(self, *, params=None, full=None):
        params, _ = self.check_precompile(params=params)
        if full is None:
            full = DEFAULT_FULL_ARG
        log_offset = False if full else self._options.get("subtr_const_value", False)
        if log_offset is None:
            log_offset = False
        with self._check_set_input_params(params):
            return self._call_value(log_offset)

    def _call_value(self, log_offset):
        return self._evaluate((), False, log_offset)[0]

    def gradient(self, params=None, *, numgrad=None, paramvals=None):
        params = self._input_check_params(params, only_independent=True)
        numgrad = self._options["numgrad"] if numgrad is None else numgrad
        paramvals, _ = self.check_precompile(params=paramvals)
        with self._check_set_input_params(paramvals):
            return self._call_value_gradient(params, numgrad, full=False)[1]

    def value_gradient(self, params=None, *, full=None, numgrad=None, paramvals=None):
        params = self._input_check_params(params)
        numgrad = self._options["numgrad"] if numgrad is None else numgrad
        if full is None:
            full = DEFAULT_FULL_ARG
        paramvals, _ = self.check_precompile(params=paramvals)
        with self._check_set_input_params(paramvals):
            return self._call_value_gradient(params, numgrad, full)

    def _call_value_gradient(self, params, numgrad, full):
        log_offset = False if full else self._options.get("subtr_const_value", False)
        if log_offset is None:
            log_offset = False
        if numgrad:
            value = self._evaluate((), False, log_offset)[0]
            return value, _numerical_gradient(lambda: self._evaluate((), False, log_offset)[0], params)
        return self._evaluate(params, True, log_offset)

    def hessian(self, params=None, hessian=None, *, numgrad=None, paramvals=None):
        params = self._input_check_params(params)
        numgrad = self._options["numgrad"] if numgrad is None else numgrad
        paramvals, _ = self.check_precompile(params=paramvals)
        with self._check_set_input_params(paramvals):
            return self._call_hessian(params, hessian, numgrad)

    def _call_hessian(self, params, hessian, numgrad):
        """The exact Hessian from ONE fused pass over the events (``znll_eval_hess``: per-event second derivatives of flat
        trees accumulated in the kernel, the chain rule through normalisation integrals, composed parameters, the
        extended term and the constraints on the host) whenever every channel's tree supports it -- all five BASELINE
        configs do; otherwise, or with ``numgrad=True`` / ``options["hessian"] == "numeric"``, central differences of
        the analytic gradient (2P fused-kernel calls).  The reference's default is a numdifftools Hessian over
        ``loss.value`` (``numhess=True``, ``loss.py:271-272``, ``z/math.py:104-151``); its exact route differentiates
        the whole graph a second time (``z/math.py:288-352``)."""
        log_offset = self._options.get("subtr_const_value", False) or False
        if not numgrad and self._options.get("hessian", "exact") != "numeric":
            H = self._exact_hessian(params, log_offset)
            if H is not None:
                self.last_hessian_method = "one-pass"
                return np.diag(H).copy() if hessian == "diag" else H
        self.last_hessian_method = "gradient-differences"

        def grad_fn():
            if numgrad:
                return _numerical_gradient(lambda: self._evaluate((), False, log_offset)[0], params)
            return self._evaluate(params, True, log_offset)[1]

        H = _numerical_jacobian(grad_fn, params)
        H = 0.5 * (H + H.T)
        if hessian == "diag":
            return np.diag(H).copy()
        return H

    def _exact_hessian(self, params, log_offset):
        """d2 loss / d theta^2 for ``params`` from the one-pass slot Hessian, or None when the engine has none."""
        eng = self._get_engine(check=False)
        if not hasattr(eng, "eval_hess"):
            return None
        slot_params = eng.slot_params
        slots = np.array([p.value() for p in slot_params], dtype=np.float64)
        got = eng.eval_hess(slots, 0.0 if log_offset is False else float(log_offset))
        if got is None:
            return None
        _, gs, mats = got
        params = list(params)
        index = {id(p): i for i, p in enumerate(params)}
        npar = len(params)
        H = np.zeros((npar, npar))
        second_cache = {}

        def jac_row(p):  # d p / d theta over `params`
            row = np.zeros(npar)
            for leaf, d in (p._partials().items() if not p.independent else [(p, 1.0)]):
                j = index.get(id(leaf))
                if j is not None:
                    row[j] += d
            return row

        def second(p):  # d2 p / d theta^2 over `params` (zero for independent and constant parameters)
            if p.independent or not p._leaves():
                return None
            if id(p) not in second_cache:
                second_cache[id(p)] = _second_partials(p, params, index)
            return second_cache[id(p)]

        for (lo, hi), Hs in zip(self._channel_slot_ranges(eng), mats):
            J = np.stack([jac_row(sp) for sp in slot_params[lo:hi]]) if hi > lo else np.zeros((0, npar))
            H += J.T @ Hs @ J
            for j, sp in enumerate(slot_params[lo:hi]):
                S = second(sp)
                if S is not None and gs[lo + j] != 0.0:
                    H += gs[lo + j] * S
        H += self._extra_hessian(params, index, jac_row, second)
        for con in self.constraints:
            H += _constraint_hessian(con, params, index)
        return 0.5 * (H + H.T)

    def _extra_hessian(self, params, index, jac_row, second):
        """Second derivatives of the terms beyond the per-event sum (extended Poisson term); none here."""
        return np.zeros((len(params), len(params)))

    def value_gradient_hessian(self, params=None, *, hessian=None, full=None, numgrad=None, paramvals=None):
        params = self._input_check_params(params)
        if full is None:
            full = DEFAULT_FULL_ARG
        paramvals, _ = self.check_precompile(params=paramvals)
        with self._check_set_input_params(paramvals):
            v, g = self._call_value_gradient(params, self._options["numgrad"] if numgrad is None else numgrad, full)
            h = self._call_hessian(params, hessian, False)
        return v, g, h

    def gradients(self, *_, **__):
        msg = "`gradients` is deprecated, use `gradient` instead."
        raise BreakingAPIChangeError(msg)

    def value_gradients(self, *_, **__):
        msg = "`value_gradients` is deprecated, use `value_gradient` instead."
        raise BreakingAPIChangeError(msg)

    def __add__(self, other):
        if not isinstance(other, BaseLoss):
            msg = "Has to be a subclass of `BaseLoss` or overwrite `__add__`."
            raise TypeError(msg)
        if type(other) is not type(self):
            msg = "cannot safely add two different kind of loss."
            raise ValueError(msg)
        kwargs = {
            "model": self.model + other.model,
            "data": self.data + other.data,
            "constraints": self.constraints + other.constraints,
        }
        fit_range = self.fit_range + other.fit_range
        if any(fr is not None for fr in fit_range):
            kwargs["fit_range"] = fit_range
        return type(self)(**kwargs)

    def create_new(self, model=None, data=None, fit_range=None, constraints=None, options=None):
        """New loss of the same kind taking over constants and options (``loss.py:923-1017``)."""
        model = self.model if model is None else model
        data = self.data if data is None else data
        if fit_range is None and any(fr is not None for fr in self.fit_range):
            fit_range = self.fit_range
        constraints = self.constraints if constraints is None else constraints
        if options is None:
            options = {k: v for k, v in self._options.items() if not k.startswith("_")}
        kwargs = {"model": model, "data": data, "constraints": constraints, "options": options}
        if fit_range is not None:
            kwargs["fit_range"] = fit_range
        return type(self)(**kwargs)

    def __repr__(self):
        return (
            f"<{type(self).__name__} model={[m.name for m in self.model]} data={[d.name for d in self.data]} "
            f'constraints={self.constraints if len(self.constraints) <= 3 else "True"}>'
        )


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _constraint_check_convert(constraints):
    for c in constraints:
        if not (hasattr(c, "value") and hasattr(c, "get_params") and hasattr(c, "_partials")):
            msg = (
                "Constraints have to be of type `Constraint`, a simple constraint from a function can be constructed"
                " with `SimpleConstraint`."
            )
            raise BreakingAPIChangeError(msg)
    return list(constraints)


def _steps(params):
    return [1e-5 * max(1.0, abs(p.value())) if not getattr(p, "has_stepsize", False) else min(p.stepsize * 1e-2, 1e-4 * max(1.0, abs(p.value()))) for p in params]


def _numerical_gradient(func, params):
    """Central differences with two Richardson levels, O(h^6) (stands in for numdifftools, ``z/math.py:43-79``)."""
    g = np.zeros(len(params))
    for i, p in enumerate(params):
        v = p.value()
        h = 1e-4 * max(1.0, abs(v))
        if p.has_limits:  # stay inside the limits (value() clips)
            room = min(v - p.lower if p.lower is not None else np.inf, p.upper - v if p.upper is not None else np.inf)
            if room > 0:
                h = min(h, room / 4.5)

        def d(hh, p=p, v=v):
            p.assign(v + hh)
            up = func()
            p.assign(v - hh)
            dn = func()
            p.assign(v)
            return (up - dn) / (2 * hh)

        d1, d2, d4 = d(h), d(2 * h), d(4 * h)
        r1, r2 = (4.0 * d1 - d2) / 3.0, (4.0 * d2 - d4) / 3.0
        g[i] = (16.0 * r1 - r2) / 15.0
    return g


def _numerical_jacobian(grad_fn, params):
    n = len(params)
    H = np.zeros((n, n))
    for i, p in enumerate(params):
        v = p.value()
        h = 1e-4 * max(abs(v), p.stepsize if hasattr(p, "stepsize") else 1.0, 1e-3)
        if p.has_limits:
            lo = -np.inf if p.lower is None else p.lower
            up = np.inf if p.upper is None else p.upper
            h = min(h, 0.5 * (up - v) if np.isfinite(up) and up > v else h, 0.5 * (v - lo) if np.isfinite(lo) and v > lo else h)
        p.assign(v + h)
        gp = np.asarray(grad_fn(), dtype=np.float64)
        p.assign(v - h)
        gm = np.asarray(grad_fn(), dtype=np.float64)
        p.assign(v)
        H[i, :] = (gp - gm) / (2 * h)
    return H


class UnbinnedNLL(BaseLoss):
    """``zfit.loss.UnbinnedNLL(model, data, fit_range=None, constraints=None, options=None)``
    (``loss.py:1020-1170``): ignores yields (``is_yield=False``, :1160-1170)."""

    def __init__(self, model, data, fit_range=None, constraints=None, options=None):
        super().__init__(model=model, data=data, fit_range=fit_range, constraints=constraints, options=options)
        self._errordef = 0.5
        extended_pdfs = [pdf for pdf in self.model if pdf.is_extended]
        if extended_pdfs and type(self) is UnbinnedNLL:
            warnings.warn(
                f"Extended PDFs ({extended_pdfs}) are given to a normal UnbinnedNLL. This won't take the yield "
                "into account and simply treat the PDFs as non-extended PDFs. To create an "
                "extended NLL, use the `ExtendedUnbinnedNLL`.",
                stacklevel=2,
            )

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **kw):
        if not self.is_extended:
            is_yield = False  # the loss does not depend on the yields
        return super().get_params(floating, is_yield, extract_independent, **kw)


class ExtendedUnbinnedNLL(UnbinnedNLL):
    """``zfit.loss.ExtendedUnbinnedNLL`` (``loss.py:1178-1332``): adds, per channel,
    ``tf.nn.log_poisson_loss(samplesize, log(yield), compute_full_loss=(log_offset is False))`` and, when an
    offset is in use, ``+ log_offset`` (``:1301-1317``)."""

    @property
    def is_extended(self) -> bool:
        return True

    def _extra_terms(self, want_grad, log_offset):
        eng = self._get_engine(check=False)
        total = 0.0
        grad = {}
        for k, mod in enumerate(self.model):
            if not mod.is_extended:
                msg = f"The pdf {mod} is not extended but has to be (for an extended fit)"
                raise NotExtendedPDFError(msg)
            t = float(eng.samplesize[k])
            yparam = mod.get_yield()
            y = yparam.value()
            log_y = math.log(y) if y > 0 else (float("-inf") if y == 0 else float("nan"))
            term = math.exp(log_y) - log_y * t
            if log_offset is False:
                if not (0.0 <= t <= 1.0):
                    term += t * math.log(t) - t + 0.5 * math.log(2.0 * math.pi * t)
            else:
                term += float(log_offset)
            total += term
            if want_grad:
                dterm = 1.0 - t / y if y != 0 else float("nan")
                for leaf, d in yparam._partials().items():
                    grad[leaf] = grad.get(leaf, 0.0) + dterm * d
        return total, grad

    def _extra_hessian(self, params, index, jac_row, second):
        """y - T log y per channel: (T / y^2) dy dy^T + (1 - T / y) d2y."""
        eng = self._get_engine(check=False)
        n = len(params)
        H = np.zeros((n, n))
        for k, mod in enumerate(self.model):
            t = float(eng.samplesize[k])
            yparam = mod.get_yield()
            y = yparam.value()
            row = jac_row(yparam)
            H += (t / (y * y)) * np.outer(row, row)
            S = second(yparam)
            if S is not None:
                H += (1.0 - t / y) * S
        return H


def _fd_step(p):
    v = p.value()
    h = 1e-4 * max(abs(v), 1e-3)
    if getattr(p, "has_limits", False):
        lo = -np.inf if p.lower is None else p.lower
        up = np.inf if p.upper is None else p.upper
        room = min(up - v, v - lo)
        if room > 0:
            h = min(h, 0.5 * room)
    return v, h


def _second_partials(p, params, index):
    """d2 p / d theta_i d theta_j of a dependent parameter over ``params``: central differences of its analytic first
    partials (host scalars only; fractions from yields, user ``ComposedParameter``s)."""
    n = len(params)
    S = np.zeros((n, n))
    for leaf in p._leaves():
        i = index.get(id(leaf))
        if i is None:
            continue
        v, h = _fd_step(leaf)
        rows = []
        for sgn in (1.0, -1.0):
            leaf.assign(v + sgn * h)
            row = np.zeros(n)
            for l2, d in p._partials().items():
                j = index.get(id(l2))
                if j is not None:
                    row[j] += d
            rows.append(row)
        leaf.assign(v)
        S[i, :] = (rows[0] - rows[1]) / (2.0 * h)
    return 0.5 * (S + S.T)


def _constraint_hessian(con, params, index):
    """d2 constraint / d theta^2 by central differences of its analytic gradient (host scalars only)."""
    n = len(params)
    S = np.zeros((n, n))
    for leaf in con.get_params(floating=None):
        i = index.get(id(leaf))
        if i is None:
            continue
        v, h = _fd_step(leaf)
        rows = []
        for sgn in (1.0, -1.0):
            leaf.assign(v + sgn * h)
            row = np.zeros(n)
            for l2, d in con._partials().items():
                j = index.get(id(l2))
                if j is not None:
                    row[j] += d
            rows.append(row)
        leaf.assign(v)
        S[i, :] = (rows[0] - rows[1]) / (2.0 * h)
    return 0.5 * (S + S.T)


class SimpleLoss(BaseLoss):
    """``zfit.loss.SimpleLoss(func, params, errordef, *, gradient=None, hessian=None)`` (``loss.py:1340-1633``):
    the generic callable slot -- lets any function (+ gradient) be driven by the minimizers."""

    def __init__(self, func, params=None, errordef=None, *, gradient=None, hessian=None, jit=None, deps=None):
        if deps is not None and params is None:
            params = deps
        if errordef is None:
            errordef = getattr(func, "errordef", None)
        if errordef is None:
            msg = "errordef has to be given (or be an attribute of func): 0.5 for a NLL, 1 for a chi2"
            raise ValueError(msg)
        if params is None:
            msg = "params have to be given"
            raise ValueError(msg)
        if isinstance(params, ZfitParameter):
            params = [params]
        self._simple_params = OrderedSet(params)
        self._simple_func = func
        self._simple_grad = None if gradient in (None, "num", "zfit") else gradient
        self._simple_hess = None if hessian in (None, "num", "zfit") else hessian
        super().__init__(model=[], data=[], options={"subtr_const": False})
        self._errordef = errordef

    def get_params(self, floating=True, is_yield=None, extract_independent=True, **_):
        return OrderedSet(p for p in self._simple_params if floating is None or p.floating == floating)

    def check_precompile(self, *, params=None, force=False):
        return params, False

    def _call_func(self):
        try:
            return float(self._simple_func(list(self._simple_params)))
        except TypeError:
            return float(self._simple_func())

    def _evaluate(self, params, want_grad, log_offset):
        value = self._call_func()
        if not want_grad:
            return value, None
        if self._simple_grad is not None:
            allp = list(self._simple_params)
            g = np.asarray(self._simple_grad(allp), dtype=np.float64)
            return value, np.array([g[allp.index(p)] for p in params])
        return value, _numerical_gradient(self._call_func, params)
