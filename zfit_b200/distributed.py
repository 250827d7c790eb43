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


def bind_to_gpu_numa_node(device: int | None = None) -> dict:
    """Pin this process to the CPU cores that are local to its GPU (``/sys/bus/pci/devices/<bdf>/local_cpulist``).

    With one process per GPU on a multi-socket host, every rank stages its event columns through pinned host memory;
    Linux places those pages on the NUMA node of the allocating thread, and a rank that runs on the far socket pushes
    its whole shard across the inter-socket link (and shares it with the ranks that live there).  Binding the process
    before its first pinned allocation keeps each rank's staging memory, its copy threads and its interpreter next to
    its own PCIe root.  Returns what was done (for the bench record); a no-op when the topology cannot be read."""
    info = {"bound": False}
    try:
        import torch

        dev = local_device() if device is None else int(device)
        props = torch.cuda.get_device_properties(dev)
        bdf = None
        if hasattr(props, "pci_bus_id"):
            bdf = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0"
        if bdf is None:  # older torch: ask NVML (CUDA_VISIBLE_DEVICES permitting, NVML indices are the CUDA indices)
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[dev]) if visible and all(t.strip().isdigit() for t in visible.split(",")) else dev
            raw = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            raw = raw.decode() if isinstance(raw, bytes) else raw
            bdf = raw.lower()[-12:]  # "00000000:1b:00.0" -> "0000:1b:00.0"
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        if not os.path.exists(path):
            return info
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            info.update({"pci": bdf, "note": "every allowed core is local already"})
            return info
        # ranks that share a node split its cores, so that their copy threads do not sit on each other
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        os.sched_setaffinity(0, cpus)
        numa = f"/sys/bus/pci/devices/{bdf}/numa_node"
        info.update({"bound": True, "pci": bdf, "cores": len(cpus), "local_world": local_world,
                     "numa_node": int(open(numa).read().strip()) if os.path.exists(numa) else None})
    except Exception as exc:  # never fatal: this is a placement hint
        info["error"] = repr(exc)
    return info
