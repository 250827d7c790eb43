"""Clocked FP64-peak record for the roofline denominator (writes JSON to stdout; keep it under profiles/).

    python tools/fp64_peak_record.py > gpurun_out/r2_fp64_peak.json

`znll_fp64_peak` runs a register-resident DFMA loop (every thread 8 independent chains, two distinct register operands
per DFMA -- the shape the FP64 pipe sustains at full rate, tools/fp64_pipe_bench.cu) for the requested time and reports
2 x DFMA / s.  Legs: a 50 ms burst (what a kernel timed alone sees) and a >= 4 s sustained run (what a kernel inside a long
step sees), each with the SM clock, power draw and throttle reasons sampled by one nvidia-smi child DURING the leg."""
import datetime
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zfit_b200 import _znll as Z  # noqa: E402

Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.hw_slowdown,"
     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def leg(ms):
    proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", "0", "-lms", "20"],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    time.sleep(0.3)
    t0 = datetime.datetime.now()
    tf = Z.fp64_peak(0, ms)
    t1 = datetime.datetime.now()
    time.sleep(0.05)
    proc.terminate()
    out, _ = proc.communicate(timeout=5)
    rows = []
    for line in out.splitlines():
        c = [x.strip() for x in line.split(",")]
        try:
            ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
        except ValueError:
            continue
        if t0 - datetime.timedelta(milliseconds=40) <= ts <= t1 + datetime.timedelta(milliseconds=40):
            rows.append(c)
    sm = sorted(float(r[1]) for r in rows) if rows else []
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
    return {"requested_ms": ms, "tflops": tf, "dfma_per_s": tf * 1e12 / 2, "samples": len(rows),
            "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
            "sm_max_mhz": max((float(r[2]) for r in rows), default=None),
            "power_w_max": max((float(r[3]) for r in rows), default=None),
            "temperature_c_max": max((float(r[4]) for r in rows), default=None), "reasons": reasons}


def main():
    Z.fp64_peak(0, 20.0)  # load the module, wake the clocks
    rec = {"what": "FP64 DFMA peak of this B200 (znll_fp64_peak: register-resident DFMA loop, 2 flop per DFMA)",
           "datasheet_tflops": 148 * 64 * 2 * 1.965e9 / 1e12,
           "burst": [leg(50.0) for _ in range(3)], "sustained": [leg(4000.0)], "half_second": [leg(500.0)]}
    b = max(x["tflops"] for x in rec["burst"])
    s = rec["sustained"][0]["tflops"]
    rec["summary"] = {"burst_tflops": b, "sustained_tflops": s, "sustained_over_burst": s / b,
                      "note": "bench.py uses a live 300 ms measurement of the same loop as the denominator of fp64_pipe_frac"}
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
