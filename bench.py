#!/usr/bin/env python
"""bench.py -- UnbinnedNLL value+grad evaluations/s at 1e8 events, fp64 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config all|1..5] [--events E] [--impl reference]

A "step" is one ``loss.value_gradient(params)`` of a config's loss: every event of every channel is
evaluated (composite pdf, analytic derivatives, normalisation, log, weights, reduction).  N > 1 (torchrun):
weak scaling, every rank holds its own shard of the config's events, ONE cross-GPU reduction of the raw
accumulators per step (fused into the kernel's last block over NVLink peer memory; ``--nccl`` selects an NCCL
all-reduce instead); ``value`` = (N x events / 1e8) x steps / time, i.e. whole-job throughput in the metric's unit
(at N=1 and 1e8 events exactly "evals/s at 1e8 events").

The default run (``--config all``) measures the headline config 2 (signal+background extended fit, 1e8 events)
with everything the contract asks for -- value, e2e from pinned HOST buffers, roofline, cpu_baseline, Minuit fit --
and then configs 3, 4, 5 and 1 with the same legs minus the CPU baseline; their numbers land in the line's
``configs`` object.  At N > 1 every config is first checked for parity: the world-reduced accumulators must be
bitwise identical on all ranks and agree with the sum of every rank's own single-GPU evaluation of its shard
(``parity``).  With ``--gpus 8`` config 5 is BASELINE.json's 1e9-event sWeighted fit (1.25e8 events per GPU).

``--impl reference`` times the CPU oracle port of the reference's op graph (torch-CPU fp64 restatement with
a reverse-mode tape, all host threads); TensorFlow / zfit cannot be installed here.  It evaluates the config's
FULL event count per step (in chunks, so the tape fits host RAM) when the run fits ``--ref-budget`` seconds, and
otherwise a bounded sample, extrapolated linearly in N and flagged as such.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "UnbinnedNLL value+grad evals/s at 1e8 events fp64"
UNIT = "evals/s"
HEADLINE = 2
ALL_CONFIGS = (2, 3, 4, 5, 1)


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    ONE ``nvidia-smi -lms`` child is started when the sampler is created -- before the warm-up steps -- and prints a
    timestamped row every few milliseconds; ``with sampler:`` only records the wall-clock window of the timed region and
    the rows inside it are picked afterwards.  (Spawning a process per sample from a thread, as round 1 did, forks this
    CUDA process inside the timed region: milliseconds of stalled interpreter in an 11 ms measurement.)"""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=20, enabled=True):
        self.index = index
        self.t_begin = self.t_end = None
        self.rows = []
        self.proc = None
        self._lines = []
        if not enabled:  # N > 1: rank 0 samples its GPU; eight pollers on one driver delay every rank's launches
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", str(period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self._reader = threading.Thread(target=self._read, daemon=True)
        self._reader.start()
        t0 = time.perf_counter()
        while not self._lines and time.perf_counter() - t0 < 3.0:  # NVML start-up (0.1 - 0.3 s) must not fall into the timed region
            time.sleep(0.01)

    def _read(self):
        try:
            for line in self.proc.stdout:
                self._lines.append(line)
        except Exception:
            pass

    def __enter__(self):
        import datetime

        self.t_begin = datetime.datetime.now()
        return self

    def __exit__(self, *exc):
        import datetime

        self.t_end = datetime.datetime.now()
        return False

    def close(self):
        if self.proc is None:
            return
        time.sleep(0.05)  # let the row that covers the end of the window appear
        try:
            self.proc.terminate()  # the exact child started above
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self._reader.join(timeout=2)
        self.proc = None
        import datetime

        for line in list(self._lines):
            cells = [c.strip() for c in line.split(",")]
            if len(cells) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(cells[0], "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                continue
            self.rows.append((ts, cells[1:]))

    def summary(self):
        self.close()
        if not self.rows or self.t_begin is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable or not sampled on this rank"]}
        import datetime

        inside = [r for ts, r in self.rows if self.t_begin <= ts <= self.t_end]
        window = "inside the timed region"
        if not inside:  # a region shorter than one sampling period: the rows right around it
            pad = datetime.timedelta(milliseconds=150)
            inside = [r for ts, r in self.rows if self.t_begin - pad <= ts <= self.t_end + pad]
            window = "within 150 ms of the timed region (it is shorter than one sampling period)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample near the timed region"]}
        sm = sorted(float(r[1]) for r in inside if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        pw = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(inside), "window": window}


def _step_spread(marks):
    """Spread of this rank's per-step wall times inside the timed region (a step returns when its result is on the host):
    a mean well above the median means a few stalled steps (an interrupted rank stalls all its peers in the exchange)."""
    d = sorted((b - a) * 1e6 for a, b in zip(marks[:-1], marks[1:]))
    if not d:
        return None
    return {"median": d[len(d) // 2], "p90": d[min(len(d) - 1, (9 * len(d)) // 10)], "max": d[-1], "mean": sum(d) / len(d)}


class NvmlSampler(ClockSampler):
    """The same record taken in-process through NVML (the library ``nvidia-smi`` is a front end of): a thread reads SM
    clock, power and the clock-event reasons every ``period_ms``.  ``--clock-source nvml``; for A/B runs of how much the
    external ``nvidia-smi -lms`` child perturbs the launches."""

    def __init__(self, index=0, period_ms=20, enabled=True):
        self.index = index
        self.t_begin = self.t_end = None
        self.rows = []
        self.proc = None
        self._stop = threading.Event()
        self._thread = None
        if not enabled:
            return
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(index)
            try:
                bdf = f"{getattr(props, 'pci_domain_id', 0):08x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0"
                self._h = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nv = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        self._period = period_ms * 1e-3
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < 3.0:
            time.sleep(0.01)

    def _loop(self):
        import datetime

        nv, h = self._nv, self._h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                mask = int(get_reasons(h))
                flags = ["Active" if mask & bits[n] else "Not Active" for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
                self.rows.append((datetime.datetime.now(), [str(self.index), f"{sm:.0f}", f"{self._max:.0f}", f"{pw:.2f}", hex(mask), *flags]))
            except Exception:
                pass
            self._stop.wait(self._period)

    def close(self):
        if self._thread is not None:
            time.sleep(0.03)
            self._stop.set()
            self._thread.join(timeout=2)
            self._thread = None


# ------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline leg and --impl reference): the checker, timed -- never the product
# ------------------------------------------------------------------------------------------------
ORACLE_CHUNK = 10_000_000  # events per tape: ~40 N-vectors of fp64 -> ~3 GB of host RAM per chunk


def _oracle_case(cfg, n_sample, chunk=None):
    from tests._oracle_engine import OracleEngine
    from zfit_b200 import workloads as W

    w = W.build(cfg, n=n_sample, device="cpu")
    loss = w["loss"]
    loss._set_engine_factory(lambda m, d, r: OracleEngine(m, d, r, chunk=chunk))
    return w, loss


def _host_ram_available():
    try:
        import psutil

        return int(psutil.virtual_memory().available)
    except Exception:
        return 0


def time_cpu_port(cfg, events_full, budget_s=20.0, threads=None):
    """Time the torch-CPU fp64 restatement (same op graph as the reference + reverse-mode tape) on a bounded
    sample and extrapolate linearly in N to the metric's 1e8-event evaluation."""
    import torch

    cores = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    n_sample = 1_000_000
    w, loss = _oracle_case(cfg, n_sample)
    params = list(loss.get_params())
    loss.value_gradient(params, full=False)  # warm-up (also computes the offset)
    t0 = time.perf_counter()
    loss.value_gradient(params, full=False)
    t1 = time.perf_counter() - t0
    # scale the sample so that ~budget_s is spent in total over 3 repetitions, bounded by host RAM
    n_big = int(min(max(n_sample, n_sample * budget_s / (3 * max(t1, 1e-3))), 10_000_000, events_full))
    if n_big > n_sample * 1.5:
        w, loss = _oracle_case(cfg, n_big)
        params = list(loss.get_params())
        loss.value_gradient(params, full=False)
    else:
        n_big = n_sample
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        loss.value_gradient(params, full=False)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    n_eval = w["n_events"]
    evals_per_s_at_full = 1.0 / (t * (events_full / n_eval))
    return {
        "value": evals_per_s_at_full,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"torch-CPU fp64 restatement (autograd tape) of config {cfg} on {n_eval} events, median of 3 "
                  f"value+grad calls = {t:.3f} s, extrapolated linearly to {events_full} events",
        "seconds_per_call_sample": t,
        "n_sample": n_eval,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = HEADLINE if args.config == "all" else int(args.config)
    from zfit_b200 import workloads as W

    events = args.events or W.FULL_SIZES[cfg]
    nch = 2 if cfg == 3 else 1
    total_events = events * nch
    import torch

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    # probe the per-event cost on 1e6 events, then decide: the FULL event count per step (chunked tape) when the whole
    # (warmup + steps) run fits the budget and the columns fit host RAM, else a bounded sample extrapolated in N
    probe_n = min(1_000_000, total_events)
    w, loss = _oracle_case(cfg, probe_n // nch)
    params = list(loss.get_params())
    loss.value_gradient(params, full=False)
    t0 = time.perf_counter()
    loss.value_gradient(params, full=False)
    t_probe = max(time.perf_counter() - t0, 1e-4)
    calls = args.steps + max(1, args.warmup) + 1  # + the offset evaluation of the first call
    t_full = t_probe * total_events / w["n_events"] * calls
    ram = _host_ram_available()
    need = 8 * total_events * 6 + 50 * 8 * min(ORACLE_CHUNK, total_events)  # columns incl. generation copies + one tape
    full = (t_full <= args.ref_budget and (ram == 0 or need < 0.6 * ram)) and not args.ref_events
    if full:
        n_sample = total_events
    else:
        per_call = args.ref_budget / calls
        cap = args.ref_events or 10_000_000
        n_sample = int(min(total_events, cap, max(200_000, w["n_events"] * per_call / t_probe)))
    per_channel = n_sample // nch
    w, loss = _oracle_case(cfg, per_channel, chunk=ORACLE_CHUNK)
    params = list(loss.get_params())
    theta0 = np.array([p.value() for p in params])
    for _ in range(max(1, args.warmup)):
        loss.value_gradient(params, full=False)
    t0 = time.perf_counter()
    for k in range(args.steps):
        for p, v in zip(params, theta0):
            p.assign(v * (1.0 + 1e-9 * (k + 1)))
        loss.value_gradient(params, full=False)
    dt = time.perf_counter() - t0
    n_eval = w["n_events"]
    scale = total_events / n_eval
    value = args.steps / (dt * scale) * (total_events / 1e8 if cfg != 1 else 1.0)
    how = ("every step evaluates all %d events (tape over chunks of %d)" % (n_eval, ORACLE_CHUNK) if n_eval == total_events
           else "each step evaluates a bounded sample of %d events; ms_per_step and value are extrapolated linearly in N "
                "to %d events (full size would take %.0f s > --ref-budget %.0f s, or not fit host RAM)"
                % (n_eval, total_events, t_full, args.ref_budget))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps * scale,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": W.NAMES[cfg], "events_per_gpu": total_events, "sample_events": n_eval,
                   "extrapolated": n_eval != total_events,
                   "note": "CPU oracle port of the reference op graph (TensorFlow/zfit not installable here); " + how},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_eval} events per step, {args.steps} steps, torch-CPU fp64 + autograd tape"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# the product arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def _barrier(ctx):
    ctx.torch.cuda.synchronize(ctx.local)
    if ctx.world > 1:
        ctx.td.barrier()
    ctx.torch.cuda.synchronize(ctx.local)


def _max_over_ranks(ctx, values):
    t = ctx.torch.tensor(list(values), dtype=ctx.torch.float64, device=f"cuda:{ctx.local}")
    if ctx.world > 1:
        ctx.td.all_reduce(t, op=ctx.td.ReduceOp.MAX)
    return [float(v) for v in t]


def parity_check(ctx, loss, w, off):
    """N > 1, before anything is timed: (a) the world-reduced (values, slot gradient) of one evaluation are bitwise
    identical on every rank; (b) they agree with the sum over ranks of each rank's own single-GPU evaluation of its
    shard (a second engine with ``distributed=False`` on the same resident columns) to 1e-12 (value) / 1e-10 (gradient)."""
    import torch

    from zfit_b200.loss import CudaEngine

    td = ctx.td
    eng = loss._get_engine()
    slots = np.array([p.value() for p in eng.slot_params], dtype=np.float64)
    vals, gs = eng.eval(slots, grad=True, log_offset=off)
    mine = torch.from_numpy(np.concatenate([vals, gs])).to(f"cuda:{ctx.local}")
    gathered = [torch.empty_like(mine) for _ in range(ctx.world)]
    td.all_gather(gathered, mine)
    bitwise = all(bool((g.view(torch.int64) == gathered[0].view(torch.int64)).all()) for g in gathered)
    local = CudaEngine(loss.model, loss.data, loss.fit_range, distributed=False,
                       sort_events=loss.options.get("sort_events", True))
    lv, lg = local.eval(slots, grad=True, log_offset=off)
    loc = torch.from_numpy(np.concatenate([lv, lg])).to(f"cuda:{ctx.local}")
    parts = [torch.empty_like(loc) for _ in range(ctx.world)]
    td.all_gather(parts, loc)
    tot = np.sum(np.stack([p.cpu().numpy() for p in parts]), axis=0)  # fixed rank order
    local.close()
    nv = len(vals)
    ref_v, ref_g = tot[:nv], tot[nv:]
    rel_v = float(np.max(np.abs(vals - ref_v) / np.maximum(np.abs(ref_v), 1e-300)))
    rel_g = float(np.max(np.abs(gs - ref_g)) / max(float(np.max(np.abs(ref_g))), 1e-300))
    ok = bool(bitwise and rel_v <= 1e-12 and rel_g <= 1e-10)
    return {"ranks_bitwise": bool(bitwise), "vs_local_sum_rel": rel_v, "grad_vs_local_sum_rel": rel_g,
            "tolerance": {"value": 1e-12, "gradient": 1e-10}, "ok": ok, "reduce": eng.reduce_mode}


def measure_config(ctx, cfg, args, headline):
    """All legs of one config on this rank (all ranks in lock-step); returns the config's result dict."""
    import zfit_b200 as zfit
    from zfit_b200 import _znll as Z  # noqa: F401
    from zfit_b200 import workloads as W

    torch, td, world, rank, local = ctx.torch, ctx.td, ctx.world, ctx.rank, ctx.local
    events = args.events or W.FULL_SIZES[cfg]
    w = W.build(cfg, n=events, device=f"cuda:{local}", rank=rank)
    loss = w["loss"]
    if args.nccl:
        loss.options["fused_reduce"] = False
    if args.sort is not None:
        loss.options["sort_events"] = {"1": True, "0": False, "lazy": "lazy"}[args.sort]
    params = list(loss.get_params())
    theta0 = np.array([p.value() for p in params])
    n_events = w["n_events"]  # events evaluated per step on this rank
    loss.value_gradient(params, full=False)  # binds data, compiles nothing (AOT), computes the offset
    eng = loss._get_engine()
    plan = eng.plan
    stream = ctx.stream
    slots = np.array([p.value() for p in eng.slot_params])
    off = loss.options.get("subtr_const_value") or 0.0
    res = {"workload": W.NAMES[cfg], "events_per_gpu": n_events, "n_params": len(params), "kernel": eng.kernel_names}

    if world > 1:
        res["parity"] = parity_check(ctx, loss, w, off)
        if not res["parity"]["ok"]:
            msg = f"multi-GPU parity failed for config {cfg}: {res['parity']}"
            raise RuntimeError(msg)

    def step(k):
        for p, v in zip(params, theta0):  # a new theta every step, as a minimizer would
            p.assign(v * (1.0 + 1e-9 * (k + 1)))
        return loss.value_gradient(params, full=False)

    # ---- value: K steps through the loss, inputs resident in HBM, CUDA events on the launch stream ------------
    sampler = (NvmlSampler if args.clock_source == "nvml" else ClockSampler)(
        local, period_ms=max(1, args.clock_period_ms), enabled=(rank == 0 and args.clock_period_ms > 0))  # its nvidia-smi child starts here, outside the timed region
    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            step(k)
        launches0 = plan.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import gc

        gc.collect()
        gc.disable()  # a collection pause inside an 8 ms timed region is a 10 % outlier; re-enabled right after
        try:
            with sampler as clocks:
                marks = [0.0] * (args.steps + 1)
                # nothing rank-dependent between the barrier and the first step: with the reduction inside the kernel a
                # rank that starts late (the collection above takes milliseconds and differs per rank) makes every peer
                # wait in its first launch, and their event pair would time that wait
                _barrier(ctx)
                t_wall0 = marks[0] = time.perf_counter()
                e0.record(stream)
                for k in range(args.steps):
                    step(k)
                    marks[k + 1] = time.perf_counter()
                e1.record(stream)
                _barrier(ctx)
                t_wall = time.perf_counter() - t_wall0
        finally:
            gc.enable()
        dev_ms = e0.elapsed_time(e1)
        launches = plan.launch_count() - launches0
    dev_ms, wall_ms = _max_over_ranks(ctx, [dev_ms, t_wall * 1e3])
    units = world * n_events / 1e8 if cfg != 1 else float(world)  # 1e8-event evaluations per step, all ranks
    res.update({
        "value": units * args.steps / (dev_ms * 1e-3),
        "ms_per_step": dev_ms / args.steps,
        "units_per_step": units,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "step_wall_us": _step_spread(marks),
        "l2": "inputs (%.0f MB per GPU) are larger than the 126 MB L2; no flush needed" % (n_events * w["bytes_per_event"] / 1e6)
              if n_events * w["bytes_per_event"] > 130e6 else
              "inputs (%.2f MB) fit the L2: this config is launch-latency-bound by design (report us per call)" % (n_events * w["bytes_per_event"] / 1e6),
        "parallelism": f"event-sharded x{world}, {plan.total_acc()} fp64 accumulators reduced per step via {eng.reduce_mode}" if world > 1 else "single GPU",
    })
    steady = units * args.steps / (wall_ms * 1e-3)
    res["e2e"] = {
        "value": steady,
        "unit": UNIT,
        "h2d_bytes_per_step": int(8 * len(params) + sum(112 + 8 * 16 for _ in eng.kernel_names)),
        "d2h_bytes_per_step": int(8 * plan.total_acc()),
        "steady_state": {
            "value": steady,
            "note": "wall clock of the same K loss.value_gradient(params) calls with the event columns already "
                    "bound (theta from host numpy in, accumulators back to pinned host memory out)",
        },
    }

    # ---- roofline of the dominant kernel: back-to-back launches, CUDA events on the launch stream -----------
    # every rank runs it in lock-step (with the fused exchange a launch waits for its peers); rank 0 reports its own
    peaks, peak_src = ctx.peaks
    _barrier(ctx)
    with torch.cuda.stream(stream):
        reps = max(5, min(args.steps, 50))
        for _ in range(3):
            plan.launch(slots, grad=True, log_offset=off, stream=stream.cuda_stream)
        torch.cuda.synchronize(local)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(reps):
            plan.launch(slots, grad=True, log_offset=off, stream=stream.cuda_stream)
        k1.record(stream)
        torch.cuda.synchronize(local)
        kern_ms = k0.elapsed_time(k1) / reps
    fp64_peak = ctx.fp64_peak
    flops = W.F_ALG[cfg] * n_events
    bytes_alg = w["bytes_per_event"] * n_events
    achieved_tf = flops / (kern_ms * 1e-3) / 1e12
    achieved_gbs = bytes_alg / (kern_ms * 1e-3) / 1e9
    t_fp64 = flops / (fp64_peak * 1e12)
    t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
    traffic, dp_per_event, capture = None, None, None
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            captured = json.loads(tr.read_text())
            traffic = captured.get(f"cfg{cfg}")
            dp_per_event = captured.get("dp_instr_per_event", {}).get(f"cfg{cfg}")
            capture = captured.get("source")
        except Exception:
            traffic, dp_per_event = None, None
    # executed FP64 instructions (ncu capture of this kernel under profiles/) over the pipe's measured issue peak
    # (one DFMA = 2 flop): what the hardware did, as opposed to what a counting convention says it had to do
    pipe_frac = None if dp_per_event is None else dp_per_event * n_events / (kern_ms * 1e-3) / (fp64_peak * 1e12 / 2.0)
    hbm_frac = achieved_gbs / peaks["hbm_gbs"]
    t_pipe = None if dp_per_event is None else dp_per_event * n_events / (fp64_peak * 1e12 / 2.0)
    # bound: the slower of the two hardware limits for the work the kernel executes
    bound = "hbm" if (t_pipe is None and t_hbm > t_fp64) or (t_pipe is not None and t_hbm >= t_pipe) else "fp64"
    if bound == "hbm":
        frac, achieved, peak, unit = hbm_frac, achieved_gbs, peaks["hbm_gbs"], "GB/s"
    elif pipe_frac is not None:
        frac, achieved, peak, unit = pipe_frac, pipe_frac * fp64_peak, fp64_peak, "TFLOP/s"
    else:
        frac, achieved, peak, unit = max(t_fp64, t_hbm) / (kern_ms * 1e-3), achieved_tf, fp64_peak, "TFLOP/s"
    res["roofline"] = {
        "bound": bound,
        "achieved": achieved,
        "peak": peak,
        "unit": unit,
        "frac": frac,
        "traffic": traffic,
        "kernel": eng.kernel_names[0],
        "kernel_ms": kern_ms,
        "launches_per_step": len(eng.kernel_names),
        "algorithmic_bytes_per_event": w["bytes_per_event"],
        "executed_dp_instr_per_event": dp_per_event,
        "fp64_pipe_frac": pipe_frac,
        "hbm_frac": hbm_frac,
        "hbm_achieved_gbs": achieved_gbs,
        "hbm_peak_gbs": peaks["hbm_gbs"],
        "hbm_peak_source": peak_src,
        "fp64_peak_tflops": fp64_peak,
        "fp64_peak_source": "measured live: register-resident DFMA loop (znll_fp64_peak), 300 ms; clocked burst / "
                            "sustained record: profiles/r2_fp64_peak.json",
        "frac_convention": max(t_fp64, t_hbm) / (kern_ms * 1e-3),
        "algorithmic_flops_per_event_convention": W.F_ALG[cfg],
        "ncu_capture": capture,
        "note": "frac = achieved / peak of the binding hardware unit: for an FP64-bound kernel the DP instructions it "
                "EXECUTES per event (ncu capture under profiles/) x events / kernel time over the measured DFMA issue "
                "peak (fp64_peak / 2 instructions/s); for an HBM-bound kernel algorithmic bytes / kernel time over the "
                "measured copy bandwidth.  frac_convention is SURVEY.md 8d's frozen flop convention (EXP 30, LOG 45, "
                "RCP 14 flop), which over-prices the table-driven functions and can exceed 1; kept for continuity.  No "
                "tensor cores: elementwise transcendental map-reduce",
    }

    if headline and rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res["cpu_baseline"] = time_cpu_port(cfg, n_events, budget_s=args.cpu_budget)
        except Exception as exc:  # the checker failing must not void the GPU number
            res["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                                   "sample": f"failed: {exc!r}"}

    if not args.no_fit:
        # ---- end-to-end fit wall time (all ranks run the same minimizer in lock-step) ---------------------------
        minimizer = zfit.minimize.Minuit(gradient="zfit")
        fit_times, result = [], None
        try:
            for _ in range(2):  # the first fit pays one-time costs (lazy load of the covariance-pass kernel, callbacks)
                for p, v in zip(params, theta0):
                    p.assign(v)
                _barrier(ctx)
                t0 = time.perf_counter()
                result = minimizer.minimize(loss)
                _barrier(ctx)
                fit_times.append(time.perf_counter() - t0)
        except Exception as exc:  # a failed fit must not cost the evaluation numbers of this line
            result = None
            res["fit"] = {"error": repr(exc)}
        if result is not None:
            res["fit"] = {
                "wall_s": fit_times[-1],
                "cold_wall_s": fit_times[0],
                "n_eval": int(result.info.get("n_eval", -1)),
                "n_pass": int(result.info.get("n_pass", -1)),
                "valid": bool(result.valid),
                "edm": None if result.edm is None else float(result.edm),
                "driver": result.info.get("driver"),
                "minimizer": "zfit.minimize.Minuit(gradient='zfit'), tol=1e-3",
                "fitted": {p.name: result.params[p]["value"] for p in result.params},
            }
        # ---- Hessian at the fit's end point: ONE fused pass per channel (flat trees) against 2P gradient passes -------
        try:
            loss.hessian(params)  # first call loads the kernel
            _barrier(ctx)
            l0 = plan.launch_count()
            t0 = time.perf_counter()
            H_one = np.asarray(loss.hessian(params), dtype=np.float64)
            torch.cuda.synchronize(local)
            t_one = time.perf_counter() - t0
            res["hessian"] = {"method": getattr(loss, "last_hessian_method", None), "seconds": t_one,
                              "launches": int(plan.launch_count() - l0)}
            if headline:
                loss.options["hessian"] = "numeric"
                l0 = plan.launch_count()
                t0 = time.perf_counter()
                H_num = np.asarray(loss.hessian(params), dtype=np.float64)
                torch.cuda.synchronize(local)
                res["hessian"]["gradient_differences"] = {"seconds": time.perf_counter() - t0,
                                                          "launches": int(plan.launch_count() - l0)}
                # full-size agreement of the two routes, in units of sqrt(H_aa H_bb) (the differences carry O(h^2) error)
                scale = np.sqrt(np.abs(np.outer(np.diag(H_num), np.diag(H_num))))
                res["hessian"]["max_scaled_difference"] = float(np.max(np.abs(H_one - H_num) / np.where(scale > 0, scale, 1.0)))
                # positive definiteness in correlation units: yields of 1e7 next to shape parameters of O(1) make the
                # raw matrix's smallest eigenvalue a rounding artefact
                dg = np.sqrt(np.abs(np.diag(H_one)))
                dg = np.where(dg > 0, dg, 1.0)
                eig = np.linalg.eigvalsh(0.5 * (H_one + H_one.T) / np.outer(dg, dg))
                res["hessian"]["positive_definite"] = bool(eig[0] > 0)
                res["hessian"]["min_eigenvalue_scaled"] = float(eig[0])
                loss.options["hessian"] = "exact"
        except Exception as exc:
            res["hessian"] = {"error": repr(exc)}
        for p, v in zip(params, theta0):
            p.assign(v)

    if not args.no_bind:
        # ---- e2e: the same K steps through the public API starting from pinned HOST buffers.  The timed region holds
        # the host->HBM copy of every event column, the bind (bucket partition for warp coherence, sum of weights, for
        # N > 1 the exchange set-up), and per step theta in / results out.  All ranks in lock-step, wall clock, max over ranks.
        try:
            def pinned(t):
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t)
                return h

            host_cols = [[pinned(c) for c in d._cols] for d in w["data"]]
            host_w = [None if d._weights is None else pinned(d._weights) for d in w["data"]]
            h2d = sum(c.numel() * 8 for cols in host_cols for c in cols) + sum(0 if x is None else x.numel() * 8 for x in host_w)

            def from_host():
                datas = [zfit.Data(dict(zip(d.obs, cols)), obs=d.space, weights=hw, guarantee_limits=True)
                         for d, cols, hw in zip(w["data"], host_cols, host_w)]
                opts = {k: v for k, v in loss.options.items() if k in ("sort_events", "fused_reduce")}
                if off:
                    opts["subtr_const_value"] = off
                return type(loss)(model=w["model"], data=datas, options=opts or None)

            with torch.cuda.stream(stream):
                warm = from_host()  # warms the allocator and the exchange set-up; untimed like the W warm-up steps
                warm.value_gradient(params, full=False)
                del warm
                _barrier(ctx)
                t0 = time.perf_counter()
                loss2 = from_host()
                loss2.value_gradient(params, full=False)
                torch.cuda.synchronize(local)
                t_bind = time.perf_counter() - t0
                for k in range(args.steps):
                    for p, v in zip(params, theta0):
                        p.assign(v * (1.0 + 1e-9 * (k + 1)))
                    loss2.value_gradient(params, full=False)
                _barrier(ctx)
                t_all = time.perf_counter() - t0
            t_all, t_bind = _max_over_ranks(ctx, [t_all, t_bind])
            per_step = res["e2e"]["h2d_bytes_per_step"]
            res["e2e"].update({
                "value": units * args.steps / t_all,
                "h2d_bytes_per_step": int(h2d / args.steps + per_step),
                "seconds": t_all,
                "bind": {"seconds": t_bind, "h2d_bytes": int(h2d), "h2d_gbs": h2d / t_bind / 1e9,
                         "note": "zfit.Data(pinned host columns) -> HBM + plan bind + first value_gradient, paid once per "
                                 "fit; it is inside the e2e timed region"},
                "note": "K steps of loss.value_gradient(params) on a loss built from pinned HOST columns inside the timed "
                        "region: H2D of all event columns once (they then stay resident, as the reference's Data tensors "
                        "do), theta in and 1+P results out every step; wall clock, max over ranks",
            })
            del loss2, host_cols, host_w
        except Exception as exc:
            res["e2e"]["bind"] = {"error": repr(exc)}
        for p, v in zip(params, theta0):
            p.assign(v)
    eng.close()
    del loss, w, eng, plan
    torch.cuda.empty_cache()
    return res


def _summary(res):
    r = res["roofline"]
    out = {
        "workload": res["workload"],
        "events_per_gpu": res["events_per_gpu"],
        "value": res["value"],
        "unit": UNIT if "cfg1" not in res["workload"] else "calls/s",
        "ms_per_step": res["ms_per_step"],
        "kernel_ms": r["kernel_ms"],
        "roofline": {k: r[k] for k in ("bound", "frac", "fp64_pipe_frac", "hbm_frac", "frac_convention", "achieved", "peak", "unit")},
        "e2e": {"value": res["e2e"]["value"], "bind_seconds": res["e2e"].get("bind", {}).get("seconds"),
                "steady_state": res["e2e"]["steady_state"]["value"]},
        "clocks": res["clocks"],
        "gpu_launches": res["gpu_launches"],
        "step_wall_us": res.get("step_wall_us"),
    }
    if "fit" in res:
        out["fit"] = {k: res["fit"].get(k) for k in ("wall_s", "n_pass", "valid", "edm", "error") if k in res["fit"]}
    if "hessian" in res:
        out["hessian"] = res["hessian"]
    if "parity" in res:
        out["parity"] = res["parity"]
    return out


def run_product(args):
    import torch
    import torch.distributed as td

    ctx = Ctx()
    ctx.torch, ctx.td = torch, td
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: zfit_b200 has no CPU fallback"}))
        return 2
    torch.cuda.set_device(ctx.local)
    ctx.numa = None
    if ctx.world > 1 and not args.no_numa:
        from zfit_b200 import distributed as _D

        ctx.numa = _D.bind_to_gpu_numa_node(ctx.local)  # before any pinned allocation of this rank
    if ctx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        td.init_process_group("nccl", device_id=torch.device("cuda", ctx.local), timeout=datetime.timedelta(seconds=300))

    import zfit_b200 as zfit  # noqa: F401
    from zfit_b200 import _znll as Z

    ctx.stream = torch.cuda.Stream(device=ctx.local)
    ctx.peaks = _peaks()
    ctx.fp64_peak = Z.fp64_peak(ctx.local, 300.0)

    if args.config == "all":
        cfgs = list(ALL_CONFIGS if ctx.world == 1 else (2, 3, 4, 5))  # cfg 1 (5.4k events) stays on one GPU
    else:
        cfgs = [int(args.config)]
    results, errors = {}, {}
    for i, cfg in enumerate(cfgs):
        try:
            results[cfg] = measure_config(ctx, cfg, args, headline=(i == 0))
        except Exception as exc:
            if i == 0:
                raise
            errors[cfg] = repr(exc)  # a secondary config failing must not cost the headline line
            _barrier(ctx)
    head = results[cfgs[0]]
    line = {
        "metric": METRIC,
        "value": head["value"],
        "unit": UNIT,
        "n_gpus": ctx.world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": head["workload"],
            "events_per_gpu": head["events_per_gpu"],
            "n_params": head["n_params"],
            "kernel": head["kernel"],
            "l2": head["l2"],
            "parallelism": head["parallelism"],
            "minuit_driver": "n/a (evaluation benchmark)",
            "also_measured": [f"cfg{c}" for c in cfgs[1:]],
        },
        "clocks": head["clocks"],
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"],
        "roofline": head["roofline"],
    }
    for key in ("cpu_baseline", "fit", "hessian", "parity", "step_wall_us"):
        if key in head:
            line[key] = head[key]
    if ctx.numa is not None:
        line["config"]["host_placement"] = ctx.numa
    line["configs"] = {f"cfg{c}": _summary(r) for c, r in results.items()}
    for c, e in errors.items():
        line["configs"][f"cfg{c}"] = {"error": e}
    if ctx.rank == 0:
        print(json.dumps(line))
    if ctx.world > 1:
        td.barrier()
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="all", help="'all' (headline cfg 2, then 3, 4, 5, 1) or one of 1..5")
    ap.add_argument("--events", type=int, default=None, help="events per GPU (per channel for config 3)")
    ap.add_argument("--ref-events", type=int, default=0, help="force a bounded sample of this many events in the reference arm")
    ap.add_argument("--ref-budget", type=float, default=420.0, help="CPU seconds the reference arm may spend in total")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="skip the e2e leg that starts from host buffers (e2e then reports the steady state only)")
    ap.add_argument("--nccl", action="store_true", help="multi-GPU: NCCL all-reduce instead of the fused NVLink exchange")
    ap.add_argument("--no-fit", action="store_true", help="skip the end-to-end Minuit fit after the timed region")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not pin each rank to the cores next to its GPU")
    ap.add_argument("--clock-period-ms", type=int, default=20, help="nvidia-smi sampling period during the timed region; 0: no sampler (A/B runs only)")
    ap.add_argument("--clock-source", default="smi", choices=["smi", "nvml"], help="clock record: the nvidia-smi -lms child (recipe) or in-process NVML")
    ap.add_argument("--sort", default=None, choices=["1", "0", "lazy"], help="override the loss option sort_events (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
