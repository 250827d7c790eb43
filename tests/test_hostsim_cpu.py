"""The kernels' per-event arithmetic on the CPU: ``znll_device.cuh`` compiled for the host (``tests/_hostsim.py``)
between the product library's own prologue and epilogue, against the oracle, to the tolerances of the GPU parity tests.

This is a check of the device *code*, not a way to run the product without a GPU: the launch machinery, the block and
cross-GPU reductions and everything that touches CUDA stay with the ``-m gpu`` tests."""

import numpy as np
import pytest

from oracle import nll as ON
from tests._cases import ALL_CASES
from tests._hostsim import HostSim

NLL_RTOL = 1e-12
GRAD_RTOL = 1e-10
N = 20_000


@pytest.fixture(scope="module")
def cases():
    return {k: (f() if k == "gauss" else f(n=N)) for k, f in ALL_CASES.items()}


@pytest.fixture(scope="module")
def sim(cases):
    return HostSim([c.kernel for c in cases.values()])


def _oracle(case, log_offset=False):
    w = None if case.weights is None else [case.weights]
    idx, names, theta = case.free()
    val, grad = ON.value_and_gradient([case.oracle_model], [case.data], theta, names, weights=w, log_offset=log_offset)
    return val, idx, grad


@pytest.mark.parametrize("key", list(ALL_CASES))
def test_device_code_matches_oracle(key, cases, sim):
    case = cases[key]
    v, g, name = sim.eval(case.nodes, case.slots, case.cols, case.weights)
    assert name == case.kernel
    ov, idx, og = _oracle(case)
    w = None if case.weights is None else [case.weights]
    exact, abs_terms = ON.exact_nll([case.oracle_model], [case.data], case.P, weights=w)
    assert abs(v - float(exact)) <= NLL_RTOL * float(abs_terms)
    n = len(case.cols[0])
    wabs = n if case.weights is None else float(np.sum(np.abs(case.weights)))
    for j, i in enumerate(idx):
        scale = max(abs(og[j]), wabs * 1e-3 * max(1.0, abs(og[j]) / wabs))
        assert abs(g[i] - og[j]) <= GRAD_RTOL * max(scale, abs(og).max()), (case.slot_names[i], g[i], og[j])


@pytest.mark.parametrize("key", ["cb_exp", "gauss_cheb_weighted", "product"])
def test_value_only_instantiation_and_odd_tail(key, cases, sim):
    case = cases[key]
    cols = [c[: N - 1] for c in case.cols]  # odd: the scalar instantiation handles the last event
    w = None if case.weights is None else case.weights[: N - 1]
    v0, g0, _ = sim.eval(case.nodes, case.slots, cols, w, grad=False)
    v1, g1, _ = sim.eval(case.nodes, case.slots, cols, w, grad=True)
    assert g0 is None and v0 == v1
    # the odd event's contribution equals the difference to the even prefix
    v2, g2, _ = sim.eval(case.nodes, case.slots, [c[: N - 2] for c in cols], None if w is None else w[: N - 2])
    last = [c[N - 2 : N - 1] for c in cols]
    v3, g3, _ = sim.eval(case.nodes, case.slots, last, None if w is None else w[N - 2 : N - 1], grad=False)
    assert abs((v1 - v2) - v3) <= 1e-9 * max(1.0, abs(v3))


def test_log_offset_after_the_weights(cases, sim):
    case = cases["gauss_cheb_weighted"]
    off = -3.25
    v0, g0, _ = sim.eval(case.nodes, case.slots, case.cols, case.weights)
    v1, g1, _ = sim.eval(case.nodes, case.slots, case.cols, case.weights, log_offset=off)
    assert abs((v1 - v0) - N * off) <= 1e-9 * abs(N * off)  # unweighted offset per event (loss.py:134-137)
    np.testing.assert_allclose(g0, g1, rtol=0, atol=1e-9 * np.abs(g0).max())


def test_pdf_values_match_oracle(cases, sim):
    from oracle.backend import NumpyBackend

    NP = NumpyBackend()

    for key in ("cb_exp", "product", "gauss_cheb", "nested"):
        case = cases[key]
        got = sim.pdf(case.nodes, case.slots, case.cols)
        ref = np.asarray(case.oracle_model.pdf(NP, case.data, case.P), dtype=np.float64)
        np.testing.assert_allclose(got, ref, rtol=2e-13, atol=0)


def test_negative_pdf_gives_nan_not_an_error(cases, sim):
    case = cases["gauss_cheb"]
    s = case.slots.copy()
    s[0], s[1], s[5] = 0.0, 1.0, 3.0  # pure polynomial with c1 = 3: negative on the left side
    v, g, _ = sim.eval(case.nodes, s, case.cols)
    assert np.isnan(v)


def test_tiny_probabilities_take_the_epsilon_path(sim, cases):
    """p + 1e-307 (loss.py:117) matters when p underflows: events 60 sigma out must give log(1e-307), not -inf."""
    case = cases["gauss"]
    x = np.array([2.0, 2.5, 200.0, -190.0, 1.0])
    nodes = list(case.nodes)
    nodes[0].norm_lo, nodes[0].norm_hi = -300.0, 300.0
    v, g, _ = sim.eval(nodes, [2.0, 0.5], [x])
    assert np.isfinite(v) and np.all(np.isfinite(g))
    lp = -0.5 * ((x - 2.0) / 0.5) ** 2 - np.log(0.5 * np.sqrt(2 * np.pi))
    with np.errstate(under="ignore"):
        ref = -np.sum(np.log(np.exp(lp) + 1e-307))
    assert abs(v - ref) <= 1e-12 * abs(ref)


# ---- the hand-rolled fp64 elementary functions, same bounds as tests/test_math_gpu.py -----------------------------
def _ulp_err(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


@pytest.mark.parametrize("pairs", [False, True])
def test_fast_exp(sim, pairs):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-699, 699, 100_000), rng.uniform(-40, 5, 100_000), rng.normal(0, 1e-3, 50_000),
                        [0.0, -0.0, 1.0, -1.0, 699.999, -699.999]])
    assert _ulp_err(sim.math(0, x, pairs), np.exp(x)).max() <= 2.0
    xs = np.array([710.0, -746.0, -720.0, 800.0, -800.0, np.inf, -np.inf, np.nan, 700.0, -700.0, 0.5, -0.5])
    got = sim.math(0, xs, pairs)
    with np.errstate(all="ignore"):
        ref = np.exp(xs)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=4e-16)


@pytest.mark.parametrize("pairs", [False, True])
def test_fast_log(sim, pairs):
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 100_000)), rng.uniform(0.5, 2.0, 100_000), rng.uniform(1e-6, 1.0, 50_000),
                        [1.0, 2.0, 0.5, 1e-307, 1e300, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 3.0]])
    got, ref = sim.math(1, x, pairs), np.log(x)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2)) <= 4e-16
    assert np.max(np.abs(got - ref)) <= 3e-13
    xs = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, 2.0, 0.25])
    got = sim.math(1, xs, pairs)
    with np.errstate(all="ignore"):
        ref = np.log(xs)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(got[ok][np.isinf(ref[ok])], ref[ok][np.isinf(ref[ok])])
    fin = ok & np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=4e-16)


@pytest.mark.parametrize("pairs", [False, True])
def test_fast_rcp(sim, pairs):
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 100_000)) * rng.choice([-1.0, 1.0], 100_000), rng.uniform(1e-3, 10, 100_000)])
    assert _ulp_err(sim.math(2, x, pairs), 1.0 / x).max() <= 1.0
    xs = np.array([0.0, np.inf, -np.inf, np.nan, 1e-310, 4.0])
    got = sim.math(2, xs, pairs)
    with np.errstate(all="ignore"):
        ref = 1.0 / xs
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=4e-16)


def test_small_table_option(cases):
    """ZNLL_TAB_BITS=7 (128-entry tables with the longer polynomials, znll_device.cuh): same accuracy bounds, same parity."""
    case = cases["gauss_cheb"]
    small = HostSim([case.kernel], defines=("ZNLL_TAB_BITS=7",))
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-699, 699, 100_000), rng.uniform(-40, 5, 100_000), rng.normal(0, 1e-3, 50_000), [0.0, -0.0, 1.0, -1.0]])
    y = np.concatenate([np.exp(rng.uniform(-700, 700, 100_000)), rng.uniform(0.5, 2.0, 100_000)])
    for pairs in (False, True):
        assert _ulp_err(small.math(0, x, pairs), np.exp(x)).max() <= 2.0
        got, ref = small.math(1, y, pairs), np.log(y)
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2)) <= 4e-16
    v, g, _ = small.eval(case.nodes, case.slots, case.cols, case.weights)
    ov, idx, og = _oracle(case)
    assert abs(v - ov) <= 1e-12 * abs(ov)
    np.testing.assert_allclose(g[idx], og, rtol=0, atol=GRAD_RTOL * np.abs(og).max())
