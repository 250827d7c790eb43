"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances are north_star's: NLL 1e-12 relative (adjudicated on the full value against the
extended-precision oracle, relative to sum|terms|), each gradient component 1e-10 relative to the
gradient scale sum_i |per-event contribution| (components go to 0 at the minimum).
"""

import numpy as np
import pytest

from oracle import nll as ON
from tests._cases import ALL_CASES
from zfit_b200 import _znll as Z

pytestmark = pytest.mark.gpu

NLL_RTOL = 1e-12
GRAD_RTOL = 1e-10


def _run(case, log_offset=0.0, grad=True, device_bind=False):
    plan = Z.Plan(1)
    plan.set_model(0, case.nodes)
    plan.bind_host(0, case.cols, case.weights)
    vals, g = plan.eval(case.slots, grad=grad, log_offset=log_offset)
    name, origin = plan.kernel_name(0), plan.kernel_origin(0)
    sumw = plan.sumw(0)
    plan.close()
    return float(vals[0]), g, name, origin, sumw


def _oracle(case, log_offset=False):
    w = None if case.weights is None else [case.weights]
    idx, names, theta = case.free()
    val, grad = ON.value_and_gradient([case.oracle_model], [case.data], theta, names, weights=w, log_offset=log_offset)
    return val, idx, grad


@pytest.mark.parametrize("key", list(ALL_CASES))
def test_value_and_gradient_match_oracle(key):
    case = ALL_CASES[key]()
    v, g, name, origin, sumw = _run(case)
    assert name == case.kernel
    ov, idx, og = _oracle(case)
    w = None if case.weights is None else [case.weights]
    exact, abs_terms = ON.exact_nll([case.oracle_model], [case.data], case.P, weights=w)
    # value: against the extended-precision sum, relative to sum|terms| (and the fp64 oracle must agree too)
    assert abs(v - float(exact)) <= NLL_RTOL * float(abs_terms)
    assert abs(ov - float(exact)) <= 1e-11 * float(abs_terms)
    # gradient: each component against the reverse-mode tape over the reference op graph
    n = len(case.cols[0])
    wabs = n if case.weights is None else float(np.sum(np.abs(case.weights)))
    for j, i in enumerate(idx):
        scale = max(abs(og[j]), wabs * 1e-3 * max(1.0, abs(og[j]) / wabs))
        assert abs(g[i] - og[j]) <= GRAD_RTOL * max(scale, abs(og).max()), (case.slot_names[i], g[i], og[j])
    expect_sumw = n if case.weights is None else float(np.sum(case.weights))
    assert abs(sumw - expect_sumw) <= 1e-12 * max(1.0, wabs)


def test_rtc_origin_and_catalogue_origin():
    c = ALL_CASES["rtc"](n=2000)
    _, _, name, origin, _ = _run(c)
    assert origin == 2, "shape outside the catalogue must be NVRTC-instantiated"
    c = ALL_CASES["cb_exp"](n=2000)
    _, _, name, origin, _ = _run(c)
    assert origin == 1


def test_rtc_image_cache_across_processes(tmp_path):
    """A tree outside the catalogue in two fresh processes sharing one ``ZNLL_CACHE_DIR``: the first compiles with NVRTC and
    stores the image, the second loads the stored cubin (no compilation) -- same kernel, bit-identical results."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    code = (
        "import json, sys, time\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "from tests._cases import ALL_CASES\n"
        "from zfit_b200 import _znll as Z\n"
        "c = ALL_CASES['rtc'](n=3000)\n"
        "plan = Z.Plan(1)\n"
        "t0 = time.perf_counter(); plan.set_model(0, c.nodes); dt = time.perf_counter() - t0\n"
        "plan.bind_host(0, c.cols, c.weights)\n"
        "v, g = plan.eval(c.slots, grad=True)\n"
        "_, _, mats = plan.eval_cov(c.slots)\n"
        "print(json.dumps({'origin': plan.kernel_origin(0), 'v': float(v[0]).hex(), 'g': [float(x).hex() for x in g],"
        " 'cov': float(mats[0].sum()).hex(), 'set_model_s': dt}))\n"
    )
    env = dict(os.environ, ZNLL_CACHE_DIR=str(tmp_path / "rtc"))
    runs = []
    for _ in range(2):
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
        assert len(list((tmp_path / "rtc").glob("rtc-*.cubin"))) == 1
    first, second = runs
    assert first["origin"] == second["origin"] == 2
    assert (first["v"], first["g"], first["cov"]) == (second["v"], second["g"], second["cov"])
    assert second["set_model_s"] < 0.5 * first["set_model_s"], (first["set_model_s"], second["set_model_s"])


@pytest.mark.parametrize("key", ["cb_exp", "gauss_cheb_weighted"])
def test_log_offset_is_subtracted_per_event_unweighted(key):
    case = ALL_CASES[key](n=20_000)
    v0, g0, *_ = _run(case, log_offset=0.0)
    off = -3.25
    v1, g1, *_ = _run(case, log_offset=off)
    n = len(case.cols[0])
    assert abs((v1 - v0) - n * off) <= 1e-9 * abs(n * off)  # -(sum(l) - n*off)
    np.testing.assert_allclose(g0, g1, rtol=0, atol=1e-9 * np.abs(g0).max())


def test_value_only_equals_value_with_grad():
    case = ALL_CASES["cb_exp"](n=30_001)  # odd: exercises the scalar tail event
    v0, *_ = _run(case, grad=False)
    v1, *_ = _run(case, grad=True)
    assert v0 == v1


def test_deterministic_bitwise_repeat():
    case = ALL_CASES["product"](n=100_003)
    a = _run(case)
    b = _run(case)
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1])


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 1023])
def test_ragged_sizes(n):
    case = ALL_CASES["gauss_cheb"](n=4096)
    case.cols = [case.cols[0][:n]]
    v, g, *_ = _run(case)
    ov, idx, og = _oracle(case)
    assert abs(v - ov) <= 1e-12 * max(1.0, abs(ov)) * 10
    for j, i in enumerate(idx):
        assert abs(g[i] - og[j]) <= 1e-9 * max(1.0, np.abs(og).max())


@pytest.mark.parametrize("n", [1, 2, 511, 512, 513, 10_001])
def test_no_value_outside_the_bound_range_is_consumed(n):
    """Canary check (compute-sanitizer is not available on the pool): the bound columns sit inside larger buffers whose
    surroundings are NaN; a read past either end that reached the arithmetic would poison value, gradient or the
    covariance pass.  Results must equal those of tightly allocated copies bit for bit."""
    import torch

    case = ALL_CASES["product"](n=12_000)  # three columns: exercises the two-buffer loop
    pad = 64
    tight, padded, keep = [], [], []
    for c in case.cols:
        buf = torch.full((n + 2 * pad,), float("nan"), dtype=torch.float64, device="cuda:0")
        buf[pad : pad + n] = torch.from_numpy(np.ascontiguousarray(c[:n])).cuda()
        keep.append(buf)
        padded.append(buf[pad : pad + n])
        tight.append(buf[pad : pad + n].clone())
    out = []
    for cols in (tight, padded):
        plan = Z.Plan(1)
        plan.set_model(0, case.nodes)
        torch.cuda.synchronize()
        plan.bind_device_ptrs(0, [c.data_ptr() for c in cols], 0, n)
        v, g = plan.eval(case.slots, grad=True)
        _, _, mats = plan.eval_cov(case.slots)
        out.append((float(v[0]), np.array(g), np.array(mats[0])))
        plan.close()
    assert np.isfinite(out[1][0]) and np.all(np.isfinite(out[1][1])) and np.all(np.isfinite(out[1][2]))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_nan_propagates_not_raises():
    # negative Chebyshev pdf -> log of a negative number -> NaN must reach the caller (evaluation.py:212-227)
    case = ALL_CASES["gauss_cheb"](n=5000)
    s = case.slots.copy()
    s[0], s[1] = 0.0, 1.0  # pure polynomial
    s[5] = 3.0             # c1 = 3 -> negative on the left side
    case.slots = s
    v, g, *_ = _run(case)
    assert np.isnan(v)
