"""Throughput of the bind-time preprocessing kernels of csrc/znll_prep.cu on resident columns (one GPU).

    python tools/prep_bench.py [n_events] [repeats]   ->   one JSON line

* ``znll_partition_events`` (stable 32-bucket partition): reads the key column twice and every stream once, writes every
  stream once -> (2 + 2 S) x 8 bytes per event for S streams with the key among them ... reported as moved GB/s.
* ``znll_cut_events`` (inclusive cut + stable compaction): reads every column twice (count, move), the weights once, and
  writes the survivors -> (2 C + W + keep x (C + W)) x 8 bytes per event.
Timed with CUDA events on the launch stream, after two warm-up calls; the cut includes its host synchronisation (the
survivor count comes back) so it is also timed by wall clock."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zfit_b200 import _znll as Z

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
stream = torch.cuda.current_stream(0).cuda_stream
out = {"n_events": n, "repeats": reps}


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms, wall = [], []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2], sorted(wall)[len(wall) // 2]


for n_cols, weighted in ((1, False), (1, True), (3, False)):
    cols = [torch.empty(n, dtype=torch.float64, device=dev).uniform_(-1.2, 1.2, generator=gen) for _ in range(n_cols)]
    w = torch.empty(n, dtype=torch.float64, device=dev).normal_(1.0, 0.5, generator=gen) if weighted else None
    dst = [torch.empty_like(c) for c in cols]
    dw = None if w is None else torch.empty_like(w)
    src_p, dst_p = [c.data_ptr() for c in cols], [c.data_ptr() for c in dst]
    wp, dwp = (0, 0) if w is None else (w.data_ptr(), dw.data_ptr())
    streams = n_cols + (1 if weighted else 0)
    tag = f"{n_cols}col" + ("+w" if weighted else "")

    ms, _ = timed(lambda: Z.partition_events(0, src_p, dst_p, wp, dwp, 0, -1.2, 1.2, n, stream=stream))
    out[f"partition_{tag}"] = {"ms": ms, "bytes": (1 + 2 * streams) * 8 * n, "gbs": (1 + 2 * streams) * 8 * n / ms / 1e6}

    for name, lo, hi in (("keep_all", -2.0, 2.0), ("keep_83pct", -1.0, 1.0)):
        kept = []
        ms, wall = timed(lambda: kept.append(Z.cut_events(0, src_p, dst_p, wp, dwp, [lo] * n_cols, [hi] * n_cols, n, stream=stream)))
        frac = kept[-1] / n
        nbytes = (2 * n_cols + (frac if weighted else 0) + frac * streams) * 8 * n
        out[f"cut_{tag}_{name}"] = {"ms": ms, "wall_ms": wall, "kept_fraction": frac, "bytes": nbytes, "gbs": nbytes / ms / 1e6}
    del cols, dst, w, dw
    torch.cuda.empty_cache()

print(json.dumps(out))
