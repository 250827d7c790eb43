"""CPU: the oracle against the reference's golden vector and known-answer tests."""

import json
from pathlib import Path

import numpy as np
import pytest
from scipy import integrate, stats

from oracle import nll as ON
from oracle import pdfs as O
from oracle.backend import NumpyBackend, TorchBackend

GOLDEN = Path(__file__).parent / "golden"
B = NumpyBackend()


def _golden_model():
    g = O.Gauss("mu", "sigma", "x", (-10, 10), yield_="n_sig")
    e = O.Exponential("lambda", "x", (-10, 10), yield_="n_bkg")
    return O.SumPDF([g, e])


def test_golden_fit_fixture_pins_the_formula():
    """tests/truth/data/deterministic_fit_data.npz: at the stored fit result the gradient must vanish to Minuit's
    EDM tolerance, the yields must sum to N (tests/highstats/test_exact_yields.py:82), the value is pinned."""
    d = np.load(GOLDEN / "deterministic_fit_data.npz", allow_pickle=True)
    x = d["data"][:, 0]
    fp = json.loads(str(d["fitted_params_json"]))
    assert list(fp) == ["n_sig", "n_bkg", "mu", "sigma", "lambda"]
    m = _golden_model()
    names = m.all_param_names()
    assert names == list(fp)  # parameter ordering rule
    P = {k: np.float64(v) for k, v in fp.items()}
    v = float(ON.unbinned_nll([m], [{"x": x}], P, extended=True))
    assert abs(v - 62323.843426662) < 1e-6
    val, g = ON.value_and_gradient([m], [{"x": x}], [fp[n] for n in names], names, extended=True)
    assert abs(val - v) < 1e-8
    # EDM = g^T H^-1 g / 2 with a numerical Hessian of the oracle gradient
    th = np.array([fp[n] for n in names])
    H = np.zeros((5, 5))
    for i in range(5):
        h = 1e-4 * max(1.0, abs(th[i]))
        tp, tm = th.copy(), th.copy()
        tp[i] += h
        tm[i] -= h
        H[i] = (ON.value_and_gradient([m], [{"x": x}], tp, names, extended=True)[1]
                - ON.value_and_gradient([m], [{"x": x}], tm, names, extended=True)[1]) / (2 * h)
    edm = 0.5 * g @ np.linalg.solve(0.5 * (H + H.T), g)
    assert 0 <= edm < 1e-3  # the reference's Minuit(tol=1e-3) stop rule
    assert abs(fp["n_sig"] + fp["n_bkg"] - 21200) < 2.0
    ex, _ = ON.exact_nll([m], [{"x": x}], P, extended=True)
    assert abs(float(ex) - v) < 1e-9  # fp64 op-order restatement vs extended precision


def test_gradient_tape_vs_central_differences():
    d = np.load(GOLDEN / "deterministic_fit_data.npz", allow_pickle=True)
    x = d["data"][:5000, 0]
    m = _golden_model()
    names = m.all_param_names()
    th = np.array([1000.0, 4100.0, 0.9, 1.1, -0.07])
    _, g = ON.value_and_gradient([m], [{"x": x}], th, names, extended=True)
    ng = ON.numeric_gradient(lambda t: float(ON.unbinned_nll([m], [{"x": x}], dict(zip(names, t)), extended=True)), th)
    np.testing.assert_allclose(g, ng, rtol=2e-6, atol=1e-6)


def test_crystalball_shape_vs_scipy():
    """tests/test_pdf_cb.py:129-147: ratio to scipy.stats.crystalball constant, rtol 5e-7."""
    mu, sigma, alpha, n = -0.3, 1.1, 0.8, 2.0
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "x", (-6, 4))
    P = {"mu": mu, "sigma": sigma, "alpha": alpha, "n": n}
    x = np.linspace(-6, 4, 1001)
    ours = cb.unnorm(B, x, {k: np.float64(v) for k, v in P.items()}, None)
    ref = stats.crystalball.pdf(x, beta=alpha, m=n, loc=mu, scale=sigma)
    ratio = ours / ref
    np.testing.assert_allclose(ratio, ratio[0], rtol=5e-7)
    # normalised pdf integrates to one; the mirrored (alpha < 0) shape is the reflection
    Pn = {k: np.float64(v) for k, v in P.items()}
    integral, _ = integrate.quad(lambda t: float(cb.pdf(B, {"x": np.array([t])}, Pn)[0]), -6, 4, points=[mu - alpha * sigma])
    assert abs(integral - 1.0) < 1e-9
    Pm = dict(Pn, alpha=np.float64(-alpha))
    np.testing.assert_allclose(cb.unnorm(B, 2 * mu - x, Pm, None), ours, rtol=1e-12)


@pytest.mark.parametrize("lims", [(-6, 4), (-6, -2), (-0.5, 3), (-1.18, 4), (-6, -1.18)])
def test_crystalball_integral_vs_quad_and_additivity(lims):
    """tests/test_pdf_cb.py:53-78: analytic vs numeric integral; additivity over sub-ranges."""
    P = {"mu": np.float64(-0.3), "sigma": np.float64(1.1), "alpha": np.float64(0.8), "n": np.float64(2.0)}
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "x", (-6, 4))
    lo, hi = lims
    ana = float(cb.integral(B, P, lo, hi, None))
    num, _ = integrate.quad(lambda t: float(cb.unnorm(B, np.array([t]), P, None)[0]), lo, hi, points=[-0.3 - 0.8 * 1.1] if lo < -1.18 < hi else None)
    assert abs(ana - num) < 1e-9 * max(1.0, abs(num))
    mid = 0.5 * (lo + hi)
    assert abs(float(cb.integral(B, P, lo, mid, None)) + float(cb.integral(B, P, mid, hi, None)) - ana) < 1e-12


def test_crystalball_log_switch():
    P = {"mu": np.float64(0.0), "sigma": np.float64(1.0), "alpha": np.float64(1.5), "n": np.float64(1.0 + 2e-6)}
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "x", (-6, 3))
    ana = float(cb.integral(B, P, -6, 3, None))
    num, _ = integrate.quad(lambda t: float(cb.unnorm(B, np.array([t]), P, None)[0]), -6, 3, points=[-1.5])
    assert abs(ana - num) < 2e-6 * num  # the reference's own n ~ 1 approximation (physics.py:89,115-131): O(|n-1|)


def test_gauss_sum_hand_formula():
    """tests/test_pdfs.py:53-72 `true_gaussian_sum`."""
    mus, sigmas, fracs = [1.0, -2.0, 0.5], [1.2, 0.7, 2.0], [0.3, 0.5]
    lo, hi = -15.0, 15.0
    gs = [O.Gauss(f"m{i}", f"s{i}", "x", (lo, hi)) for i in range(3)]
    m = O.SumPDF(gs, fracs=["f0", "f1"])
    P = {f"m{i}": np.float64(mus[i]) for i in range(3)} | {f"s{i}": np.float64(sigmas[i]) for i in range(3)}
    P |= {"f0": np.float64(fracs[0]), "f1": np.float64(fracs[1])}
    x = np.linspace(-10, 10, 301)
    ours = m.pdf(B, {"x": x}, P)
    fr = [*fracs, 1 - sum(fracs)]
    true = sum(f * stats.norm.pdf(x, mu, s) / (stats.norm.cdf(hi, mu, s) - stats.norm.cdf(lo, mu, s)) for f, mu, s in zip(fr, mus, sigmas))
    np.testing.assert_allclose(ours, true, rtol=1e-12)


@pytest.mark.parametrize("coeffs", [(0.2, -0.1, 0.05, 0.02), (0.5,), (0.1, 0.3), (-0.15, 0.1, 0.0, 0.03, 0.2, -0.1)])
@pytest.mark.parametrize("lims,space", [((-10, 10), (-10, 10)), ((-3, 7), (-10, 10)), ((0.5, 1.5), (0, 2))])
def test_chebyshev_value_and_integral(coeffs, lims, space):
    """tests/test_pdf_polynomials.py:41-155: analytic integral vs numeric, values vs numpy's chebval."""
    names = [f"c{i + 1}" for i in range(len(coeffs))]
    ch = O.Chebyshev(names, "x", space)
    P = {n: np.float64(c) for n, c in zip(names, coeffs)}
    x = np.linspace(space[0], space[1], 201)
    z = (2 * x - space[0] - space[1]) / (space[1] - space[0])
    np.testing.assert_allclose(ch.unnorm(B, x, P, None), np.polynomial.chebyshev.chebval(z, [1.0, *coeffs]), rtol=1e-13, atol=1e-15)
    ana = float(ch.integral(B, P, lims[0], lims[1], None))
    num, _ = integrate.quad(lambda t: float(ch.unnorm(B, np.array([t]), P, None)[0]), *lims)
    assert abs(ana - num) < 1e-12 * max(1.0, abs(num))


def test_exponential_shift_rule_and_stability():
    """tests/test_pdfs.py:302-321: no NaN at x ~ 5500 thanks to the shift; shift is 0 with norm=False."""
    ex = O.Exponential("lam", "x", (5000, 6000))
    P = {"lam": np.float64(-0.3)}  # exp(-0.3*5500) underflows without the shift
    x = np.linspace(5000, 6000, 11)
    p = ex.pdf(B, {"x": x}, P)
    assert np.all(np.isfinite(p)) and np.all(p >= 0)
    xs = np.linspace(5000, 6000, 200001)
    assert abs(np.trapezoid(ex.pdf(B, {"x": xs}, P), xs) - 1.0) < 1e-4
    ex2 = O.Exponential("lam", "y", (0, 5))
    P2 = {"lam": np.float64(-0.7)}
    y = np.array([0.0, 1.0, 5.0])
    np.testing.assert_allclose(ex2.pdf(B, {"y": y}, P2, norm=False), np.exp(-0.7 * y), rtol=1e-15)
    np.testing.assert_allclose(float(ex2.integrate(B, P2, (0, 5))), (np.exp(-3.5) - 1) / -0.7, rtol=1e-15)


def test_product_pdf_is_product_of_normalised_factors():
    """tests/test_pdfs.py:204-227: prod.pdf == prod_i pdf_i.pdf(norm_i) (rtol 1e-5 there)."""
    g = O.Gauss("mu", "sigma", "x", (-4, 4))
    e = O.Exponential("lam", "y", (0, 5))
    c = O.Chebyshev(["c1", "c2"], "z", (-2, 4))
    m = O.ProductPDF([g, e, c])
    P = {k: np.float64(v) for k, v in {"mu": 0.3, "sigma": 1.2, "lam": -0.7, "c1": 0.2, "c2": -0.1}.items()}
    rng = np.random.default_rng(0)
    x = {"x": rng.uniform(-4, 4, 100), "y": rng.uniform(0, 5, 100), "z": rng.uniform(-2, 4, 100)}
    np.testing.assert_allclose(m.pdf(B, x, P), g.pdf(B, x, P) * e.pdf(B, x, P) * c.pdf(B, x, P), rtol=1e-13)


def test_nll_assembly_weights_offset_kahan_extended():
    rng = np.random.default_rng(3)
    x = rng.normal(0.2, 1.1, 20000)
    x = x[(x > -5) & (x < 5)]
    w = rng.normal(1.0, 0.6, x.size)
    g = O.Gauss("mu", "sigma", "x", (-5, 5), yield_="y")
    P = {"mu": np.float64(0.1), "sigma": np.float64(1.2), "y": np.float64(x.size * 0.9)}
    lp = np.log(g.pdf(B, {"x": x}, P) + 1e-307)
    # weights multiply first, the (unweighted) offset is subtracted afterwards (loss.py:134-137)
    v = float(ON.unbinned_nll([g], [{"x": x}], P, weights=[w], log_offset=0.25))
    assert abs(v - (-(np.sum(lp * w - 0.25)))) < 1e-8
    # Kahan = plain sum to fp64 rounding
    a = float(ON.unbinned_nll([g], [{"x": x}], P, kahan=True))
    b = float(ON.unbinned_nll([g], [{"x": x}], P, kahan=False))
    assert abs(a - b) < 1e-8 and abs(a + np.sum(lp)) < 1e-8
    # extended: + y - T log y (+ Stirling when full, + offset otherwise); T = sum of weights (data.py:268-275)
    T, y = float(np.sum(w)), float(P["y"])
    full = float(ON.unbinned_nll([g], [{"x": x}], P, weights=[w], extended=True))
    expect = -np.sum(lp * w) + y - T * np.log(y) + T * np.log(T) - T + 0.5 * np.log(2 * np.pi * T)
    assert abs(full - expect) < 1e-7
    part = float(ON.unbinned_nll([g], [{"x": x}], P, weights=[w], extended=True, log_offset=0.5))
    assert abs(part - (-np.sum(lp * w - 0.5) + y - T * np.log(y) + 0.5)) < 1e-7
    # check_precompile's constant (loss.py:393-409)
    off = ON.initial_log_offset([g], [{"x": x}], P)
    v0 = float(ON.unbinned_nll([g], [{"x": x}], P, log_offset=off))
    assert abs(v0 - 10000.0) < 1e-6


def test_torch_and_numpy_backends_agree():
    from tests._cases import ALL_CASES

    for key in ("cb_exp", "product", "gauss_cheb_weighted"):
        c = ALL_CASES[key](n=3000)
        w = None if c.weights is None else [c.weights]
        a = float(ON.unbinned_nll([c.oracle_model], [c.data], c.P, weights=w))
        idx, names, theta = c.free()
        b, _ = ON.value_and_gradient([c.oracle_model], [c.data], theta, names, weights=w)
        assert abs(a - b) < 1e-9 * abs(a)


def test_bernstein_and_truncated_gauss_against_scipy():
    """Known answers for the last two leaves of SURVEY section 8f row 3: scipy.interpolate.BPoly (the Bernstein basis) and
    scipy.stats.truncnorm (the reference compares its TruncatedGauss with a Gauss x Uniform product, tests/test_pdfs.py)."""
    from scipy import integrate, interpolate, stats

    from oracle.backend import NumpyBackend

    B = NumpyBackend()
    x = np.linspace(-1.9, 2.9, 41)
    c = {"c0": 0.8, "c1": 1.3, "c2": 0.4, "c3": 1.1}
    bn = O.Bernstein(list(c), "x", (-2, 3))
    bp = interpolate.BPoly(np.array(list(c.values()))[:, None], [-2, 3])
    assert np.abs(bn.unnorm(B, x, c, None) - bp(x)).max() < 1e-15
    assert abs(bn.integral(B, c, -1.0, 2.5, None) - integrate.quad(bp, -1.0, 2.5)[0]) < 1e-13
    assert np.abs(bn.pdf(B, {"x": x}, c) - bp(x) / integrate.quad(bp, -2, 3)[0]).max() < 1e-14
    P = {"mu": 0.3, "sigma": 1.2, "low": -1.0, "high": 2.0}
    tg = O.TruncatedGauss("mu", "sigma", "low", "high", "x", (-2, 3))
    ref = stats.truncnorm.pdf(x, (P["low"] - P["mu"]) / P["sigma"], (P["high"] - P["mu"]) / P["sigma"], loc=P["mu"], scale=P["sigma"])
    assert np.abs(tg.pdf(B, {"x": x}, P) - ref).max() < 1e-15
    inner = O.TruncatedGauss("mu", "sigma", "low", "high", "x", (-0.5, 1.5))  # norm range inside the truncation: plain Gauss
    xs = np.array([0.0, 1.0])
    gauss = stats.norm.pdf(xs, 0.3, 1.2) / (stats.norm.cdf(1.5, 0.3, 1.2) - stats.norm.cdf(-0.5, 0.3, 1.2))
    assert np.abs(inner.pdf(B, {"x": xs}, P) - gauss).max() < 1e-15
