"""Instruction budget of the four full-size step kernels, read from the SASS of the built catalogue objects (no GPU).

DESIGN.md section 3.1 rests on these numbers: the step kernels are bound by the instructions they issue (a DP
instruction occupies its sub-partition's issue port for 2 cycles, 3 when a DFMA reads three distinct registers), and the
model "cycles per event pair = 2 DP2 + 3 DP3 + non-DP" of ``tools/sass_loop_stats.py`` matched the measured kernels to
within 4 %.  The test walks the cheapest path through each kernel's event loop (an iteration that takes no rare branch)
and holds it to the counts the measured kernels of round 1 had, so that a change of the device header which makes the
hot path longer is seen on the CPU, before GPU time is spent."""

import importlib.util
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "zfit_b200" / "_build"

spec = importlib.util.spec_from_file_location("sass_loop_stats", ROOT / "tools" / "sass_loop_stats.py")
S = importlib.util.module_from_spec(spec)
spec.loader.exec_module(S)

# kernel (demangled, no spaces) -> (DP instructions, modelled issue cycles) of one trip through the event loop.
# (round 1's cfg 4 ran the loop body twice per trip -- two register buffers --; its old numbers were for two pairs.)
BUDGET = {
    # round 2 (optimistic sweep with one rare-path test per pair, fixed-reference log-sum-exp, plain per-thread sums,
    # 32-bit pair index, table addresses in registers); round 1 had (94, 311), (109, 325), (111, 341), (170, 513)
    # (every tree now streams its columns through a cp.async ring whose loop body exists once: a few more address
    # instructions per pair than the register prefetch, paid back several times over by the hidden HBM latency)
    # (unweighted trees: split evaluation -- the root log's arguments go into a running product per thread, one log per
    # few hundred pairs instead of one per event --, pair reciprocals from one rcp, range tests as running maxima;
    # before these cfg 2 / 3 / 4 stood at (86, 270), (105, 305), (81, 239))
    "nll_kernel<znll::Sum<znll::CBall<0>,znll::Expo<0>>,false,true>": (66, 205),  # cfg 2, pair on the Gaussian core
    "nll_kernel<znll::Sum<znll::Gauss<0>,znll::Poly<0,4,0>>,false,true>": (86, 244),  # cfg 3
    "nll_kernel<znll::Sum<znll::Gauss<0>,znll::Poly<0,4,0>>,true,true>": (107, 315),  # cfg 5 (weighted)
    "nll_kernel<znll::Prod<znll::Gauss<0>,znll::Expo<1>,znll::Poly<2,4,0>>,false,true>": (61, 181),  # cfg 4, ONE pair
}


@pytest.fixture(scope="module")
def kernels():
    objs = sorted(BUILD.glob("znll_catalog_*.o"))
    if not objs:
        pytest.skip("catalogue objects not built (python -m zfit_b200.build)")
    found = {}
    for obj in objs:
        funcs = S.functions(obj)
        names = list(funcs)
        for mangled, pretty in zip(names, S.demangle(names)):
            found[pretty.split("(")[0].replace("void ", "").replace(" ", "")] = funcs[mangled]
    return found


@pytest.mark.parametrize("name", list(BUDGET))
def test_event_loop_stays_within_its_instruction_budget(kernels, name):
    key = "znll::" + name
    assert key in kernels, sorted(k for k in kernels if "nll_kernel" in k)[:5]
    instrs = kernels[key]
    best = None
    for lo, hi in sorted(set(S.loops(instrs))):
        path, cycles = S.hot_path(instrs, lo, hi)
        if cycles is None or cycles > 50_000:  # loops whose every path holds a rare-path marker (table fill)
            continue
        dp = sum(1 for _, op, _ in path if op.split(".")[0] in S.DP_OPS)
        if dp >= 50 and (best is None or dp > best[0]):  # the event loop is the one long DP-heavy loop
            best = (dp, cycles, len(path))
    assert best is not None, "event loop not found"
    dp, cycles, n = best
    dp_budget, cycle_budget = BUDGET[name]
    assert dp <= dp_budget, f"{name}: {dp} DP instructions on the hot path (budget {dp_budget}); {n} instructions in total"
    assert cycles <= cycle_budget, f"{name}: modelled issue cycles {cycles} (budget {cycle_budget})"
