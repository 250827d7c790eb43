"""Condense an `ncu --set full` report (.ncu-rep) of the step kernel into the JSON kept under profiles/.

usage: python tools/ncu_summary.py REPORT.ncu-rep EVENTS_PER_LAUNCH OUT.json "capture command"
Reads the report with `ncu -i ... --page raw --csv` (metric values) and `--page source --csv` (instruction mix).
"""
import collections, csv, io, json, subprocess, sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "launch__block_size",
    "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, events, out, capture = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
    raw = page(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}
    metrics = {k: {"value": row[col[k]], "unit": units[col[k]]} for k in KEEP if k in col}

    def num(k):
        m = metrics[k]
        return float(m["value"].replace(",", "")) * UNIT.get(m["unit"], 1.0)

    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    cyc = num("sm__cycles_elapsed.avg")
    dp = sum(num(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed") for o in ("dadd", "dfma", "dmul")) * cyc
    src = page(rep, "source")
    h2 = src[1]
    ix, isrc = h2.index("Instructions Executed"), h2.index("Source")
    mix = collections.Counter()
    for r in src[2:]:
        if len(r) <= ix or not r[ix].isdigit():
            continue
        tok = r[isrc].split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        mix[op.rstrip(";")] += int(r[ix])
    pairs = events / 2 / 32
    summary = {
        "kernel": row[col["Kernel Name"]],
        "events_per_launch": events,
        "dram_bytes_per_launch": dram,
        "dram_bytes_per_event": dram / events,
        "dp_thread_instructions_per_event": dp / events,
        "warp_instructions_x32_per_event": num("smsp__inst_executed.sum") * 32 / events,
        "warp_instructions_per_event_pair_by_opcode": {k: round(v / pairs, 2) for k, v in mix.most_common(24)},
        "capture": capture,
        "metrics": metrics,
    }
    json.dump(summary, open(out, "w"), indent=1)
    print(out, "dram B/event %.3f" % (dram / events), "DP/event %.1f" % (dp / events),
          "instr/event %.1f" % summary["warp_instructions_x32_per_event"], "time", metrics["gpu__time_duration.sum"])


if __name__ == "__main__":
    main()
