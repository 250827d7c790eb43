"""Where the bind of a loss built from pinned host columns spends its time (cfg given on the command line).

    python tools/bind_breakdown.py 2 [repeats]

Stages are separated by torch.cuda.synchronize(), so their sum exceeds the pipelined bind bench.py measures; the point is
to see which stage is worth attacking.  Output: one line per repeat with the stage times in ms."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zfit_b200 as zfit
from zfit_b200 import workloads as W
from zfit_b200.loss import CudaEngine

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = W.build(cfg, device="cuda:0")
params = list(w["loss"].get_params())
w["loss"].value_gradient(params, full=False)
off = w["loss"].options.get("subtr_const_value") or 0.0

def pinned(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h
host_cols = [[pinned(c) for c in d._cols] for d in w["data"]]
host_w = [None if d._weights is None else pinned(d._weights) for d in w["data"]]
nbytes = sum(c.numel() * 8 for cols in host_cols for c in cols) + sum(0 if x is None else x.numel() * 8 for x in host_w)
sync = torch.cuda.synchronize

def stamp(marks, name):
    sync(); marks.append((name, time.perf_counter()))

for r in range(reps):
    marks = []
    stamp(marks, "start")
    datas = [zfit.Data(dict(zip(d.obs, cols)), obs=d.space, weights=hw, guarantee_limits=True)
             for d, cols, hw in zip(w["data"], host_cols, host_w)]
    loss = type(w["loss"])(model=w["model"], data=datas, options={"subtr_const_value": off} if off else None)
    stamp(marks, "Data+loss objects")
    for d in datas:
        d.device_columns(0)
    stamp(marks, "H2D %.0f MB" % (nbytes / 1e6))
    eng = loss._get_engine()
    stamp(marks, "engine (plan, partition, sumw)")
    loss.value_gradient(params, full=False)
    stamp(marks, "first value_gradient")
    loss.value_gradient(params, full=False)
    stamp(marks, "second value_gradient")
    t0 = marks[0][1]
    print(f"cfg{cfg} rep{r}: " + " | ".join(f"{n} {1e3 * (t - p):.2f}" for (n, t), (_, p) in zip(marks[1:], marks[:-1])) +
          f" | total {1e3 * (marks[-2][1] - t0):.2f} ms (H2D alone at {nbytes / 1e9 / max(1e-9, (marks[2][1] - marks[1][1])):.1f} GB/s)", flush=True)
    # engine internals once more, separately
    if r == reps - 1:
        models, datas2, frs = loss.model, loss.data, loss.fit_range
        t = time.perf_counter(); e2 = CudaEngine(models, datas2, frs, sort_events=False); sync(); t_nosort = time.perf_counter() - t
        t = time.perf_counter(); e3 = CudaEngine(models, datas2, frs, sort_events=True); sync(); t_sort = time.perf_counter() - t
        print(f"cfg{cfg}: engine without partition {1e3 * t_nosort:.2f} ms, with {1e3 * t_sort:.2f} ms")
        e2.close(); e3.close()
    del loss, datas, eng
