"""Instruction mix of a kernel's event loop, read from the SASS of the built objects (no GPU needed).

    python tools/sass_loop_stats.py 'nll_kernel<Sum<CBall<0>,Expo<0>>,false,true>' [--object zfit_b200/_build/znll_catalog_b.o]

The step kernels are issue-bound: a DP instruction holds the sub-partition's issue port for 2
cycles (3 when a DFMA reads three distinct registers), every other instruction for 1.  So the cycles one warp spends on
an event pair can be estimated from the loop body alone:  2*DP2 + 3*DP3 + nonDP.  This tool finds the backward branches
of the kernel, takes each loop body [target, branch], and prints the instruction histogram and that estimate, so that
instruction-count changes can be judged before GPU time is spent on them.  Blocks inside the address range that are
only reached through rare-path branches are still counted: read the numbers as an upper bound and compare like with
like.
"""

from __future__ import annotations

import argparse
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DP_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
    return out.splitlines()


def functions(obj: Path):
    """-> {mangled: [(addr, opcode, text)]}"""
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            text = m.group(2).strip()
            parts = text.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            cur.append((int(m.group(1), 16), op, text))
    return funcs


def distinct_regs(text):
    """Number of distinct 64-bit register sources of a DFMA (constant-bank / immediate operands do not count)."""
    ops = text.split(None, 1)[1] if not text.startswith("@") else text.split(None, 2)[2]
    srcs = [o.strip() for o in ops.split(",")][1:]
    regs = {re.sub(r"[-|~]|\.reuse", "", s) for s in srcs if re.match(r"[-|~]?\|?R\d+", s.strip())}
    return len(regs)


def loops(instrs):
    out = []
    for addr, op, text in instrs:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr:
                out.append((int(m.group(1), 16), addr))
    return out


def stats(instrs, lo, hi):
    body = [(a, op, t) for a, op, t in instrs if lo <= a <= hi]
    hist = collections.Counter(op.split(".")[0] for _, op, _ in body)
    dp2 = dp3 = 0
    for _, op, t in body:
        base = op.split(".")[0]
        if base in DP_OPS:
            if base == "DFMA" and distinct_regs(t) >= 3:
                dp3 += 1
            else:
                dp2 += 1
    n = len(body)
    return {"n": n, "dp": dp2 + dp3, "dp3": dp3, "non_dp": n - dp2 - dp3, "cycles": 2 * dp2 + 3 * dp3 + (n - dp2 - dp3),
            "hist": hist}


# operands that only occur in libdevice's special-case handling (inf / NaN results, subnormal rescaling): blocks holding
# them are rare paths, however short -- the walker must not take them as shortcuts
RARE_MARKS = ("+INF", "-INF", "QNAN", "e+16", "CALL")


def cost(op, text):
    if any(m in text for m in RARE_MARKS):
        return 100000
    base = op.split(".")[0]
    if base in DP_OPS:
        return 3 if base == "DFMA" and distinct_regs(text) >= 3 else 2
    return 1


def hot_path(instrs, lo, hi):
    """Cheapest path (in issue cycles) from the loop head to its back edge through the control-flow graph of the body:
    every range test of the event code is a hammock whose fast side is the cheaper one, so this is the instruction
    stream of an iteration that takes no rare path (for a CrystalBall: a pair on the Gaussian core).
    -> (path as a list of instruction indices, its cost)"""
    import heapq

    body = [(a, op, t) for a, op, t in instrs if lo <= a <= hi]
    index = {a: i for i, (a, _, _) in enumerate(body)}
    n = len(body)
    dist = [None] * n
    prev = [None] * n
    heap = [(cost(body[0][1], body[0][2]), 0, None)]
    while heap:
        d, i, p = heapq.heappop(heap)
        if dist[i] is not None:
            continue
        dist[i], prev[i] = d, p
        a, op, text = body[i]
        if a == hi:
            break
        succ = []
        base = op.split(".")[0]
        if base == "BRA":
            m = re.search(r"0x([0-9a-f]+)", text)
            tgt = int(m.group(1), 16) if m else None
            if tgt in index:
                succ.append(index[tgt])
            if text.startswith("@") and i + 1 < n:  # predicated: may fall through
                succ.append(i + 1)
        elif base in ("EXIT", "RET", "BRX", "JMP"):
            if text.startswith("@") and i + 1 < n:
                succ.append(i + 1)
        elif i + 1 < n:
            succ.append(i + 1)
        for j in succ:
            if dist[j] is None:
                heapq.heappush(heap, (d + cost(body[j][1], body[j][2]), j, i))
    end = index[hi]
    if dist[end] is None:
        return [], None
    path, i = [], end
    while i is not None:
        path.append(i)
        i = prev[i]
    path.reverse()
    return [body[i] for i in path], dist[end]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("kernel", help="substring of the demangled kernel name (spaces ignored)")
    ap.add_argument("--object", action="append", help="object file(s); default: every zfit_b200/_build/znll_catalog_*.o")
    ap.add_argument("--min-body", type=int, default=40, help="ignore loops shorter than this many instructions")
    ap.add_argument("--top", type=int, default=12, help="opcodes to list")
    ap.add_argument("--hot", action="store_true", help="also walk the cheapest path through each loop body (no rare path taken)")
    ap.add_argument("--dump-hot", action="store_true", help="print that path")
    args = ap.parse_args()
    objs = [Path(o) for o in args.object] if args.object else sorted((ROOT / "zfit_b200" / "_build").glob("znll_catalog_*.o"))
    want = args.kernel.replace(" ", "")
    found = False
    for obj in objs:
        funcs = functions(obj)
        names = list(funcs)
        for mangled, pretty in zip(names, demangle(names)):
            if want not in pretty.replace(" ", ""):
                continue
            found = True
            instrs = funcs[mangled]
            print(f"{pretty.split('(')[0]}   [{obj.name}]  {len(instrs)} instructions in total")
            for lo, hi in sorted(set(loops(instrs))):
                s = stats(instrs, lo, hi)
                if s["n"] < args.min_body:
                    continue
                top = ", ".join(f"{k} {v}" for k, v in s["hist"].most_common(args.top))
                print(f"  loop 0x{lo:04x}-0x{hi:04x}: {s['n']} instr, DP {s['dp']} (3-register DFMA {s['dp3']}), "
                      f"non-DP {s['non_dp']}, issue cycles/iteration ~ {s['cycles']}")
                print(f"    {top}")
                if args.hot or args.dump_hot:
                    path, c = hot_path(instrs, lo, hi)
                    if c is not None:
                        dp = sum(1 for _, op, _ in path if op.split(".")[0] in DP_OPS)
                        hist = collections.Counter(op.split(".")[0] for _, op, _ in path)
                        print(f"    cheapest path: {len(path)} instr, DP {dp}, non-DP {len(path) - dp}, issue cycles ~ {c}")
                        print("      " + ", ".join(f"{k} {v}" for k, v in hist.most_common(args.top)))
                        if args.dump_hot:
                            for a, _, t in path:
                                print(f"        {a:05x}  {t}")
    if not found:
        sys.exit(f"no kernel matching {args.kernel!r}")


if __name__ == "__main__":
    main()
