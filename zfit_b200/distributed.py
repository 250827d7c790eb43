"""Event sharding across GPUs (one process per GPU, ``torch.distributed`` for the plumbing).

The NLL and every gradient component are plain sums over events (``src/zfit/core/loss.py:134-142``), so the
path shards naturally: each rank keeps a contiguous slice of every channel's columns resident in its
own HBM for the whole fit and, per evaluation, the ranks combine their ``2 + n_slots`` raw fp64
accumulators with ONE all-reduce (NCCL over NVLink when the tensors are CUDA tensors, gloo on CPU
tensors in the host-logic tests).  Every rank runs the same minimizer on the same reduced numbers
(SPMD), so no parameter broadcast is needed as long as the reduced result is bit-identical on all
ranks; ``deterministic=True`` guarantees that with an all-gather followed by a fixed rank-order sum.
"""

from __future__ import annotations

import os


def _td():
    import torch.distributed as td

    return td


def is_distributed() -> bool:
    try:
        td = _td()
    except Exception:
        return False
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank() -> int:
    return _td().get_rank() if is_distributed() else 0


def world_size() -> int:
    return _td().get_world_size() if is_distributed() else 1


def local_device() -> int:
    """CUDA device of this rank: LOCAL_RANK under torchrun, else 0."""
    return int(os.environ.get("LOCAL_RANK", "0")) if is_distributed() else 0


def shard_bounds(n: int, r: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Contiguous, near-equal slice [start, stop) of n events for rank r; start is even so that every
    shard of a 16-byte aligned column stays 16-byte aligned for the double2 loads."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    per = -(-n // world)
    per += per & 1
    start = min(n, r * per)
    stop = min(n, start + per)
    return start, stop


def shard(data, r: int | None = None, world: int | None = None):
    """Rank-local slice of a ``zfit_b200.Data`` (events and weights)."""
    from .data import Data

    start, stop = shard_bounds(data.num_entries, r, world)
    new = object.__new__(Data)
    new._cols = [c[start:stop] for c in data._cols]
    new._weights = None if data._weights is None else data._weights[start:stop]
    new._space, new.name, new.label = data._space, data.name, data.label
    new._dev_cache = None
    new._version = 0
    return new


def all_reduce_sum_(t, deterministic: bool = False):
    """In-place sum of a 1-D fp64 tensor over all ranks (NCCL for CUDA tensors, gloo for CPU tensors).

    ``deterministic``: all-gather + fixed rank-order sum, bit-identical on every rank whatever algorithm
    the backend would pick for an all-reduce."""
    if not is_distributed():
        return t
    td = _td()
    if not deterministic:
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return t
    import torch

    parts = [torch.empty_like(t) for _ in range(td.get_world_size())]
    td.all_gather(parts, t.contiguous())
    acc = parts[0].clone()
    for p in parts[1:]:
        acc += p
    t.copy_(acc)
    return t


class _CudaArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {
            "shape": (n,),
            "typestr": "<f8",
            "data": (ptr, False),
            "version": 2,
            "strides": None,
        }


def device_view(ptr: int, n: int, device: int):
    """Zero-copy torch view of a device fp64 buffer owned by libznll (the raw accumulators)."""
    import torch

    return torch.as_tensor(_CudaArray(ptr, n), device=f"cuda:{device}")
