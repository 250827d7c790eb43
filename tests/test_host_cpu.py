"""CPU: host layer (API mirror, lowering, chain rule, minimizer protocol) and the C-ABI surface -- no GPU."""

import json
import re
from pathlib import Path

import numpy as np
import pytest

import zfit_b200 as zfit
from oracle import nll as ON
from oracle import pdfs as O
from oracle.backend import TorchBackend
from tests._oracle_engine import OracleEngine, use_oracle
from zfit_b200 import _znll as Z
from zfit_b200 import exception as zexc

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).parent / "golden"


# ------------------------------------------------------------------------------------------------ C ABI
def test_cabi_library_loads_and_exports_every_declared_symbol():
    header = (ROOT / "include" / "znll.h").read_text()
    declared = set(re.findall(r"\b(znll_[a-z0-9_]+)\s*\(", header))
    declared -= {"znll_plan", "znll_node", "znll_kind"}
    lib = Z.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/znll.h but not exported by libznll.so"
    assert declared == set(Z.SYMBOLS), (declared ^ set(Z.SYMBOLS))
    assert lib.znll_abi_version() == 1


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(Z.ZnllError, match="no CUDA device"):
        Z.Plan(1)
    obs = zfit.Space("x", -1, 1)
    g = zfit.pdf.Gauss(zfit.Parameter("mu_nf", 0.0), zfit.Parameter("s_nf", 1.0), obs)
    nll = zfit.loss.UnbinnedNLL(g, zfit.Data(np.zeros(10), obs=obs))
    with pytest.raises(Z.ZnllError):
        nll.value()


def _autograd_integral(opdf, names, vals, lo, hi, norm):
    import torch

    B = TorchBackend()
    leaves = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in vals]
    I = opdf.integral(B, dict(zip(names, leaves)), lo, hi, norm)
    g = torch.autograd.grad(I, leaves, allow_unused=True)
    return float(I.detach()), np.array([0.0 if x is None else float(x) for x in g])


@pytest.mark.parametrize(
    "mu,s,a,n,lo,hi",
    [(5275, 25, 1.2, 2.5, 5000, 5600), (5275, 25, -1.2, 2.5, 5000, 5600), (-0.3, 1.1, 0.8, 2, -6, 4),
     (0, 1, 1.5, 1.000001, -6, -3), (0, 1, 1.5, 3.0, -6, -3), (0, 1, 1.5, 3.0, -1, 3), (0, 1, 1.5, 1.000002, -5, 3),
     (0.0, -1.0, 1.5, 3.0, -5, 3)],
)
def test_host_prologue_crystalball_integral_and_derivatives(mu, s, a, n, lo, hi):
    node = Z.Node(Z.CB, 0, 0, 0, lo, hi, 0, 0, 0)
    I, dI = Z.leaf_integral(node, [mu, s, a, n])
    rI, rdI = _autograd_integral(O.CrystalBall("mu", "s", "a", "n", "x", (lo, hi)), ["mu", "s", "a", "n"], [mu, s, a, n], lo, hi, (lo, hi))
    assert abs(I - rI) <= 1e-14 * abs(rI)
    np.testing.assert_allclose(dI, rdI, rtol=1e-12, atol=1e-14 * abs(rI))


def test_host_prologue_other_leaf_integrals():
    I, dI = Z.leaf_integral(Z.Node(Z.GAUSS, 0, 0, 0, -2, 3, 0, 0, 0), [1.2, 1.3])
    rI, rdI = _autograd_integral(O.Gauss("mu", "sigma", "x", (-2, 3)), ["mu", "sigma"], [1.2, 1.3], -2, 3, (-2, 3))
    assert abs(I - rI) <= 1e-15 and np.allclose(dI, rdI, rtol=1e-13)
    I, dI = Z.leaf_integral(Z.Node(Z.EXP, 0, 0, 0, 5000, 5600, 0, 0, 5300.0), [-0.003])
    rI, rdI = _autograd_integral(O.Exponential("l", "x", (5000, 5600)), ["l"], [-0.003], 5000, 5600, (5000, 5600))
    assert abs(I - rI) <= 1e-13 * rI and np.allclose(dI, rdI, rtol=1e-12)
    I, dI = Z.leaf_integral(Z.Node(Z.EXP, 0, 0, 0, 0, 5, 0, 0, 0.0), [-0.7])  # inside a ProductPDF: no shift
    rI, rdI = _autograd_integral(O.Exponential("l", "x", (0, 5)), ["l"], [-0.7], 0, 5, False)
    assert abs(I - rI) <= 1e-15 and np.allclose(dI, rdI, rtol=1e-13)
    cn = ["c0", "c1", "c2", "c3", "c4"]
    cv = [1.0, 0.2, -0.1, 0.05, 0.02]
    I, dI = Z.leaf_integral(Z.Node(Z.CHEB, 0, 0, 4, -3, 7, -10, 10, 0), cv)
    rI, rdI = _autograd_integral(O.Chebyshev(cn[1:], "x", (-10, 10), coeff0="c0"), cn, cv, -3, 7, (-3, 7))
    assert abs(I - rI) <= 1e-14 * abs(rI) and np.allclose(dI, rdI, rtol=1e-13)


def test_host_prologue_next_row_leaves():
    """Legendre / Chebyshev2 / Hermite / Laguerre / GeneralizedCB integrals + derivatives vs the
    oracle tape."""
    cn, cv = ["c0", "c1", "c2", "c3"], [1.0, 0.3, -0.2, 0.1]
    for kind, cls in ((Z.LEGENDRE, O.Legendre), (Z.CHEB2, O.Chebyshev2), (Z.HERMITE, O.Hermite), (Z.LAGUERRE, O.Laguerre)):
        I, dI = Z.leaf_integral(Z.Node(kind, 0, 0, 3, -3, 7, -10, 10, 0), cv)
        rI, rdI = _autograd_integral(cls(cn[1:], "x", (-10, 10), coeff0="c0"), cn, cv, -3, 7, (-3, 7))
        assert abs(I - rI) <= 1e-14 * abs(rI) and np.allclose(dI, rdI, rtol=1e-13)
    gn = ["mu", "sl", "al", "nl", "sr", "ar", "nr"]
    for vals, lo, hi in [([0.2, 1.1, 1.2, 2.5, 0.9, 1.5, 3.0], -8, 7), ([0.2, 1.1, 1.2, 2.5, 0.9, 1.5, 3.0], 0.5, 7),
                         ([0.2, 1.1, 1.2, 2.5, 0.9, 1.5, 3.0], -8, 0.1), ([5280, 20, 1.3, 2.2, 25, 1.1, 4.0], 5000, 5600)]:
        I, dI = Z.leaf_integral(Z.Node(Z.GCB, 0, 0, 0, lo, hi, 0, 0, 0), vals)
        rI, rdI = _autograd_integral(O.GeneralizedCB(*gn, "x", (lo, hi)), gn, vals, lo, hi, (lo, hi))
        assert abs(I - rI) <= 1e-14 * abs(rI)
        np.testing.assert_allclose(dI, rdI, rtol=1e-11, atol=1e-13 * abs(rI))


def test_host_prologue_gaussexptail_family():
    for vals, lo, hi in [([0.2, 1.1, 1.2], -8, 7), ([0.2, 1.1, -1.2], -8, 7), ([0.2, 1.1, 1.2], -8, -2), ([0.2, 1.1, 1.2], 0, 7)]:
        I, dI = Z.leaf_integral(Z.Node(Z.GET, 0, 0, 0, lo, hi, 0, 0, 0), vals)
        rI, rdI = _autograd_integral(O.GaussExpTail("mu", "s", "a", "x", (lo, hi)), ["mu", "s", "a"], vals, lo, hi, (lo, hi))
        assert abs(I - rI) <= 1e-14 * abs(rI)
        np.testing.assert_allclose(dI, rdI, rtol=1e-11, atol=1e-13 * abs(rI))
    gn = ["mu", "sl", "al", "sr", "ar"]
    for vals, lo, hi in [([0.2, 1.1, 1.2, 0.9, 1.5], -8, 7), ([0.2, 1.1, 1.2, 0.9, 1.5], 0.5, 7), ([0.2, 1.1, 1.2, 0.9, 1.5], -8, 0.1)]:
        I, dI = Z.leaf_integral(Z.Node(Z.GGET, 0, 0, 0, lo, hi, 0, 0, 0), vals)
        rI, rdI = _autograd_integral(O.GeneralizedGaussExpTail(*gn, "x", (lo, hi)), gn, vals, lo, hi, (lo, hi))
        assert abs(I - rI) <= 1e-14 * abs(rI)
        np.testing.assert_allclose(dI, rdI, rtol=1e-10, atol=1e-13 * abs(rI))
    obs = zfit.Space("x", -8, 7)
    pdf = zfit.pdf.GeneralizedGaussExpTail(0.2, 1.1, 1.2, 0.9, 1.5, obs)
    assert abs(float(pdf.integrate((-8, 7))[0]) - 1.0) < 1e-13 and pdf._lower()[0][0].kind == Z.GGET


def test_host_prologue_bernstein_and_truncated_gauss():
    """SURVEY section 8f row 3, last two leaves: integrals + derivatives vs the oracle tape, lowering, validation."""
    cn, cv = ["c0", "c1", "c2", "c3"], [0.8, 1.3, 0.4, 1.1]
    for lo, hi in [(-10, 10), (-3, 7), (2.5, 9.0)]:
        I, dI = Z.leaf_integral(Z.Node(Z.BERNSTEIN, 0, 0, 3, lo, hi, -10, 10, 0), cv)
        rI, rdI = _autograd_integral(O.Bernstein(cn, "x", (-10, 10)), cn, cv, lo, hi, (lo, hi))
        assert abs(I - rI) <= 1e-14 * abs(rI) and np.allclose(dI, rdI, rtol=1e-13)
    tn = ["mu", "sigma", "low", "high"]
    for vals, lo, hi in [([0.3, 1.2, -1.0, 2.0], -2, 3), ([0.3, 1.2, -5.0, 2.0], -2, 3), ([0.3, 1.2, -1.0, 8.0], -2, 3),
                         ([0.3, 1.2, -5.0, 8.0], -2, 3), ([5280.0, 20.0, 5100.0, 5500.0], 5000, 5600)]:
        I, dI = Z.leaf_integral(Z.Node(Z.TGAUSS, 0, 0, 0, lo, hi, 0, 0, 0), vals)
        # the oracle's integral carries the truncation's own normaliser Z, which cancels in pdf = prob / integral:
        # compare I / Z and the derivatives of ln(I / Z)... i.e. differentiate the oracle's  integral * Z
        tg = O.TruncatedGauss(*tn, "x", (lo, hi))

        class _TimesZ:
            limits = (lo, hi)

            def integral(self, B, P, a, b, norm):
                return tg.integral(B, P, a, b, norm) * tg._parts(B, P)[4]

        rI, rdI = _autograd_integral(_TimesZ(), tn, vals, lo, hi, (lo, hi))
        assert abs(I - rI) <= 1e-14 * abs(rI)
        np.testing.assert_allclose(dI, rdI, rtol=1e-11, atol=1e-14)
    obs = zfit.Space("x", -2, 3)
    tgp = zfit.pdf.TruncatedGauss(0.3, 1.2, -1.0, 2.0, obs)
    assert abs(float(tgp.integrate((-2, 3))[0]) - 1.0) < 1e-13 and tgp._lower()[0][0].kind == Z.TGAUSS
    bern = zfit.pdf.Bernstein(obs, cv)
    nodes, slots = bern._lower()
    assert nodes[0].kind == Z.BERNSTEIN and nodes[0].degree == 3 and len(slots) == 4 and bern.degree == 3
    assert abs(float(bern.integrate((-2, 3))[0]) - 1.0) < 1e-13
    with pytest.raises(ValueError, match="at least one coefficient"):
        zfit.pdf.Bernstein(obs, [])


def test_double_cb_shares_sigma_slot():
    obs = zfit.Space("m", 5000, 5600)
    mu, sigma = zfit.Parameter("mu_dcb", 5280.0), zfit.Parameter("sigma_dcb", 20.0)
    pars = [zfit.Parameter(n + "_dcb", v) for n, v in (("al", 1.2), ("nl", 2.5), ("ar", 1.4), ("nr", 3.0))]
    dcb = zfit.pdf.DoubleCB(mu, sigma, pars[0], pars[1], pars[2], pars[3], obs)
    nodes, slots = dcb._lower()
    assert nodes[0].kind == Z.GCB and slots[1] is sigma and slots[4] is sigma and len(slots) == 7
    assert [p.name for p in dcb.get_params()] == ["mu_dcb", "sigma_dcb", "al_dcb", "nl_dcb", "ar_dcb", "nr_dcb"]
    rng = np.random.default_rng(0)
    x = rng.normal(5280, 30, 3000)
    nll = use_oracle(zfit.loss.UnbinnedNLL(dcb, zfit.Data(x, obs=obs)))
    v, g = nll.value_gradient(full=True)
    gn = nll.gradient(numgrad=True)
    np.testing.assert_allclose(g, gn, rtol=2e-5, atol=1e-5 * np.abs(g).max())  # sigma: two slots, one theta
    assert abs(float(dcb.integrate((5000, 5600))[0]) - 1.0) < 1e-13


# ------------------------------------------------------------------------------------------------ API mirror
def test_space():
    s = zfit.Space("x", -2, 3)
    assert s.obs == ("x",) and s.n_obs == 1 and s.limit1d == (-2.0, 3.0) and s.volume == 5.0 and bool(s)
    assert zfit.Space("x", (-2, 3)) == s and zfit.Space("x", limits=(-2, 3)) == s and zfit.Space(obs="x", lower=-2, upper=3) == s
    assert not zfit.Space("x").has_limits
    xy = s * zfit.Space("y", 0, 5)
    assert xy.obs == ("x", "y") and xy.volume == 25.0 and xy.with_obs("y").limit1d == (0.0, 5.0)
    assert xy.with_obs(["y", "x"]).obs == ("y", "x")
    np.testing.assert_array_equal(s.inside(np.array([-2.0, -2.0000001, 3.0, 3.1])), [True, False, True, False])  # inclusive
    assert s.lower.shape == (1, 1) and s.v1.limits[0][0] == -2.0
    with pytest.raises(Exception):
        zfit.Space("x", 3, -2)


def test_parameter_semantics():
    p = zfit.Parameter("p_sem", 1.0, 0, 2)
    assert p.floating and p.independent and p.has_limits and p.stepsize == 0.01 and not p.has_stepsize
    with pytest.raises(ValueError):
        p.set_value(2.5)
    p.set_value(2.5, clip=True)
    assert p.value() == 2.0
    with p.set_value(0.5):
        assert p.value() == 0.5
    assert p.value() == 2.0
    q = zfit.Parameter("q_sem", 3.0)
    with zfit.param.set_values([p, q], [1.5, -1.0]):
        assert (p.value(), q.value()) == (1.5, -1.0)
    assert (p.value(), q.value()) == (2.0, 3.0)
    with pytest.raises(ValueError):
        zfit.param.set_values([p, q], [1.0])
    p.assign(5.0)  # unchecked fast path; value() clips with identity gradient (parameter.py:548-557)
    assert p.value() == 2.0
    c = zfit.param.convert_to_parameter(3.0)
    assert not c.floating and not c.independent and list(c.get_params()) == []


def test_composed_parameter_chain_rule():
    a, b = zfit.Parameter("a_cp", 2.0), zfit.Parameter("b_cp", 3.0)
    prod = zfit.ComposedParameter("prod_cp", lambda p: p["a"] * p["b"] ** 2, params={"a": a, "b": b})  # torch-traceable
    assert prod.value() == 18.0
    d = prod._partials()
    assert abs(d[a] - 9.0) < 1e-12 and abs(d[b] - 12.0) < 1e-12
    import math

    nt = zfit.ComposedParameter("nt_cp", lambda p: math.sin(float(p["a"])) + float(p["b"]), params={"a": a, "b": b})  # numeric
    d = nt._partials()
    assert abs(d[a] - math.cos(2.0)) < 1e-8 and abs(d[b] - 1.0) < 1e-8
    nested = zfit.ComposedParameter("nest_cp", lambda p: p["x"] + p["y"], params={"x": prod, "y": a})
    assert abs(nested._partials()[a] - 10.0) < 1e-12
    assert [p.name for p in nested.get_params()] == ["a_cp", "b_cp"]
    with pytest.raises(zexc.ParameterNotIndependentError):
        prod.set_value(1.0)


def test_data_cut_weights_obs():
    obs = zfit.Space("x", -1, 1) * zfit.Space("y", 0, 2)
    arr = np.array([[0.0, 1.0], [-1.0, 2.0], [1.5, 1.0], [0.5, -0.1], [1.0, 0.0]])
    w = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    d = zfit.Data(arr, obs=obs, weights=w)
    assert d.num_entries == 3 and d.nentries == 3 and d.has_weights  # inclusive cut, weights cut alike
    np.testing.assert_array_equal(d.weights, [1.0, 2.0, 5.0])
    assert d.samplesize == 8.0
    np.testing.assert_array_equal(d.value("y"), [1.0, 2.0, 0.0])
    np.testing.assert_array_equal(d.with_obs(["y", "x"]).value()[:, 0], [1.0, 2.0, 0.0])
    d1 = zfit.Data.from_numpy(obs="x", array=np.arange(5.0))
    assert d1.value().shape == (5, 1) and d1.samplesize == 5.0
    import pandas as pd

    dp = zfit.Data.from_pandas(pd.DataFrame({"x": [0.0, 0.5], "y": [1.0, 1.5], "w": [2.0, 3.0]}), obs=obs, weights="w")
    assert dp.num_entries == 2 and dp.samplesize == 5.0
    with pytest.raises(ValueError):
        zfit.Data(arr, obs=obs, weights=w[:3])
    import torch

    dt = zfit.Data(torch.tensor(arr), obs=obs, weights=torch.tensor(w))
    assert dt.num_entries == 3 and dt.samplesize == 8.0


def _sig_bkg(tag, extended=True):
    obs = zfit.Space("m", 5000, 5600)
    mu = zfit.Parameter("mu" + tag, 5275.0, 5200, 5350)
    sigma = zfit.Parameter("sigma" + tag, 25.0, 5, 60)
    alpha = zfit.Parameter("alpha" + tag, 1.2, 0.1, 5)
    n = zfit.Parameter("n" + tag, 2.5, 1.05, 20)
    lam = zfit.Parameter("lambda" + tag, -0.003, -0.02, -1e-5)
    n_sig = zfit.Parameter("n_sig" + tag, 2500.0, 0, 1e5)
    n_bkg = zfit.Parameter("n_bkg" + tag, 7500.0, 0, 1e5)
    cb = zfit.pdf.CrystalBall(mu, sigma, alpha, n, obs, extended=n_sig if extended else None)
    ex = zfit.pdf.Exponential(lam, obs, extended=n_bkg if extended else None)
    return obs, cb, ex, [n_sig, n_bkg, mu, sigma, alpha, n, lam]


def test_lowering_and_param_order():
    obs, cb, ex, params = _sig_bkg("_lo")
    model = zfit.pdf.SumPDF([cb, ex])
    assert model.is_extended and abs(model.get_yield().value() - 10000.0) < 1e-9
    assert [p.name for p in model.get_params()] == [p.name for p in params]  # yields, then daughters in order
    assert [p.name for p in model.get_params(is_yield=False)] == [p.name for p in params]  # fracs depend on the yields
    nodes, slots = model._lower()
    assert [n.kind for n in nodes] == [Z.SUM, Z.CB, Z.EXP] and nodes[0].n_children == 2
    assert nodes[2].shift == 5300.0 and (nodes[1].norm_lo, nodes[1].norm_hi) == (5000.0, 5600.0)
    assert abs(slots[0].value() - 0.25) < 1e-15 and abs(slots[1].value() - 0.75) < 1e-15
    assert [s.name for s in slots[2:]] == [p.name for p in params[2:]]
    # product: Exponential unshifted, integrals over the product's range
    oy = zfit.Space("y", 0, 5)
    g = zfit.pdf.Gauss(zfit.Parameter("mu_p", 0.0), zfit.Parameter("s_p", 1.0), zfit.Space("x", -4, 4))
    e = zfit.pdf.Exponential(zfit.Parameter("lam_p", -0.7), oy)
    c = zfit.pdf.Chebyshev(zfit.Space("z", -2, 4), [zfit.Parameter("c1_p", 0.1), zfit.Parameter("c2_p", 0.2)])
    prod = zfit.pdf.ProductPDF([g, e, c])
    assert prod.obs == ("x", "y", "z")
    nodes, slots = prod._lower(("z", "x", "y"))
    assert [n.kind for n in nodes] == [Z.PROD, Z.GAUSS, Z.EXP, Z.CHEB]
    assert [n.column for n in nodes[1:]] == [1, 2, 0] and nodes[2].shift == 0.0 and nodes[3].degree == 2
    assert slots[3].value() == 1.0 and not slots[3].floating  # c_0 == 1 constant (polynomials.py:97-100)
    assert [p.name for p in prod.get_params()] == ["mu_p", "s_p", "lam_p", "c1_p", "c2_p"]
    # remainder fraction
    s3 = zfit.pdf.SumPDF([g, zfit.pdf.Gauss(zfit.Parameter("mu_q", 1.0), zfit.Parameter("s_q", 2.0), zfit.Space("x", -4, 4))],
                         fracs=zfit.Parameter("f_q", 0.3))
    _, sl = s3._lower()
    assert abs(sl[1].value() - 0.7) < 1e-15 and sl[1]._partials()[s3.params["frac_0"]] == -1.0


def test_model_validation_errors():
    obs = zfit.Space("x", -1, 1)
    g = zfit.pdf.Gauss(zfit.Parameter("mu_v", 0.0), zfit.Parameter("s_v", 1.0), obs)
    with pytest.raises(ValueError):
        zfit.pdf.SumPDF([g])
    with pytest.raises(zexc.ModelIncompatibleError):
        zfit.pdf.SumPDF([g, g])  # not extended, no fracs
    with pytest.raises(zexc.ObsIncompatibleError):
        zfit.pdf.SumPDF([g, zfit.pdf.Gauss(1.0, 1.0, zfit.Space("y", -1, 1))], fracs=0.5)
    with pytest.raises(zexc.ParamNameNotUniqueError):
        zfit.pdf.Gauss(zfit.Parameter("dup", 0.0), zfit.Parameter("dup", 1.0), obs)
    with pytest.raises(ValueError):
        zfit.pdf.Chebyshev(zfit.Space("x"), [0.1])
    ge = g.create_extended(zfit.Parameter("y_v", 10.0))
    assert ge.is_extended and not g.is_extended and [p.name for p in ge.get_params()] == ["y_v", "mu_v", "s_v"]
    with pytest.raises(zexc.AlreadyExtendedPDFError):
        ge.create_extended(5.0)
    # analytic integrals through the C library's host prologue
    assert abs(float(g.integrate((-1, 1))[0]) - 1.0) < 1e-15
    assert abs(float(g.integrate((-1, 0))[0]) - 0.5) < 1e-15


def test_loss_validation_errors():
    obs = zfit.Space("x", -1, 1)
    g = zfit.pdf.Gauss(zfit.Parameter("mu_lv", 0.0), zfit.Parameter("s_lv", 1.0), obs)
    d = zfit.Data(np.linspace(-0.9, 0.9, 50), obs=obs)
    with pytest.raises(TypeError):
        zfit.loss.UnbinnedNLL((g,), d)
    with pytest.raises(TypeError):
        zfit.loss.UnbinnedNLL(g, (d,))
    with pytest.raises(ValueError):
        zfit.loss.UnbinnedNLL([], [])
    with pytest.raises(ValueError):
        zfit.loss.UnbinnedNLL([g, g], [d])
    with pytest.raises(TypeError):
        zfit.loss.UnbinnedNLL("nope", d)
    with pytest.raises(ValueError):
        zfit.loss.UnbinnedNLL(g, zfit.Data(np.array([5.0]), obs=obs))  # empty after the cut
    with pytest.raises(zexc.IntentionAmbiguousError):
        zfit.loss.UnbinnedNLL(g, np.array([0.0, 3.0]))  # raw array outside the model space
    with pytest.raises(ValueError):
        zfit.loss.UnbinnedNLL(g, zfit.Data(np.zeros(3), obs=zfit.Space("y", -1, 1)))
    nll = use_oracle(zfit.loss.ExtendedUnbinnedNLL(g, d))
    with pytest.raises(zexc.NotExtendedPDFError):
        nll.value()


# ------------------------------------------------------------------------------------------------ loss host logic
def _oracle_direct(extended, log_offset, x, vals):
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "m", (5000, 5600), yield_="n_sig")
    ex = O.Exponential("lam", "m", (5000, 5600), yield_="n_bkg")
    m = O.SumPDF([cb, ex])
    names = ["n_sig", "n_bkg", "mu", "sigma", "alpha", "n", "lam"]
    return ON.value_and_gradient([m], [{"m": x}], vals, names, extended=extended, log_offset=log_offset)


def test_loss_value_gradient_against_direct_oracle():
    """Host layer (lowering to slots + chain rule through frac = yield/sum(yields) + Poisson term) == the oracle
    evaluated directly on the reference's parametrisation."""
    from tests._cases import cb_exp_case

    x = cb_exp_case(n=4000).cols[0]
    obs, cb, ex, params = _sig_bkg("_vg")
    model = zfit.pdf.SumPDF([cb, ex])
    vals = [p.value() for p in params]
    ext = use_oracle(zfit.loss.ExtendedUnbinnedNLL(model, zfit.Data(x, obs=obs)))
    assert [p.name for p in ext.get_params()] == [p.name for p in params]
    v, g = ext.value_gradient(full=True)
    rv, rg = _oracle_direct(True, False, x, vals)
    assert abs(v - rv) < 1e-9 * abs(rv)
    np.testing.assert_allclose(g, rg, rtol=1e-9, atol=1e-9 * np.abs(rg).max())
    # full=False: constant from the first evaluation; extended adds the offset once per channel
    off = ext.options["subtr_const_value"]
    v2, g2 = ext.value_gradient(full=False)
    rv2, rg2 = _oracle_direct(True, off, x, vals)
    assert abs(v2 - rv2) < 1e-7 and np.allclose(g2, rg2, rtol=1e-9, atol=1e-9 * np.abs(rg).max())
    assert abs(off - ON.initial_log_offset([_oracle_tree()], [{"m": x}], dict(zip(_names(), vals)), extended=True)) < 1e-12
    # UnbinnedNLL ignores the yields as yields but still depends on them through the fractions
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plain = use_oracle(zfit.loss.UnbinnedNLL(model, zfit.Data(x, obs=obs)))
    assert not plain.is_extended and [p.name for p in plain.get_params()] == [p.name for p in params]
    pv, pg = plain.value_gradient(full=True)
    rpv, rpg = _oracle_direct(False, False, x, vals)
    assert abs(pv - rpv) < 1e-9 * abs(rpv) and np.allclose(pg, rpg, rtol=1e-9, atol=1e-9 * np.abs(rpg).max())
    # call protocol (loss.py:532-568)
    assert ext(np.array(vals)) == ext.value(full=True)
    with pytest.raises(ValueError):
        ext(vals[:-1])
    with pytest.raises(TypeError):
        ext({"a": 1})
    sub = [params[3], params[0]]
    gs = ext.gradient(sub)
    assert gs[0] == g2[3] and gs[1] == g2[0]
    with pytest.raises(ValueError):
        ext.gradient([zfit.Parameter("alien", 1.0)])
    # numerical gradient option and Hessian
    gn = ext.gradient(numgrad=True)
    np.testing.assert_allclose(gn, g2, rtol=2e-5, atol=2e-5 * np.abs(g2).max())
    H = ext.hessian(params[2:5])
    assert H.shape == (3, 3) and np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H[:2, :2]) > 0)
    assert ext.hessian(params[2:5], hessian="diag").shape == (3,)


def _names():
    return ["n_sig", "n_bkg", "mu", "sigma", "alpha", "n", "lam"]


def _oracle_tree():
    cb = O.CrystalBall("mu", "sigma", "alpha", "n", "m", (5000, 5600), yield_="n_sig")
    ex = O.Exponential("lam", "m", (5000, 5600), yield_="n_bkg")
    return O.SumPDF([cb, ex])


def test_simultaneous_shared_parameters_and_loss_addition():
    rng = np.random.default_rng(0)
    o1, o2 = zfit.Space("x", -5, 5), zfit.Space("y", -5, 5)
    mu, sigma = zfit.Parameter("mu_sh", 0.3, -2, 2), zfit.Parameter("sigma_sh", 1.2, 0.2, 4)
    g1 = zfit.pdf.Gauss(mu, sigma, o1)
    g2 = zfit.pdf.Gauss(mu, zfit.Parameter("sigma2_sh", 2.0, 0.2, 6), o2)
    d1 = zfit.Data(rng.normal(0.2, 1.0, 500), obs=o1)
    d2 = zfit.Data(rng.normal(0.2, 2.2, 700), obs=o2)
    both = use_oracle(zfit.loss.UnbinnedNLL([g1, g2], [d1, d2]))
    l1, l2 = use_oracle(zfit.loss.UnbinnedNLL(g1, d1)), use_oracle(zfit.loss.UnbinnedNLL(g2, d2))
    assert [p.name for p in both.get_params()] == ["mu_sh", "sigma_sh", "sigma2_sh"]
    v, g = both.value_gradient(full=True)
    v1, ga = l1.value_gradient(full=True)
    v2, gb = l2.value_gradient(full=True)
    assert abs(v - (v1 + v2)) < 1e-9
    np.testing.assert_allclose(g, [ga[0] + gb[0], ga[1], gb[1]], rtol=1e-10)  # shared mu: two slots, one theta
    added = l1 + l2
    assert len(added.model) == 2 and isinstance(added, zfit.loss.UnbinnedNLL)
    with pytest.raises(ValueError):
        l1 + use_oracle(zfit.loss.ExtendedUnbinnedNLL(g1.create_extended(zfit.Parameter("y_sh", 500.0)), d1))


def test_constraint_enters_value_and_gradient():
    rng = np.random.default_rng(1)
    obs = zfit.Space("x", -5, 5)
    mu, sigma = zfit.Parameter("mu_c", 0.3, -2, 2), zfit.Parameter("sigma_c", 1.2, 0.2, 4)
    g = zfit.pdf.Gauss(mu, sigma, obs)
    d = zfit.Data(rng.normal(0.2, 1.0, 300), obs=obs)
    plain = use_oracle(zfit.loss.UnbinnedNLL(g, d))
    constr = zfit.constraint.GaussianConstraint(mu, observation=0.1, uncertainty=0.05)
    con = use_oracle(zfit.loss.UnbinnedNLL(g, d, constraints=constr))
    v0, g0 = plain.value_gradient(full=True)
    v1, g1 = con.value_gradient(full=True)
    assert abs((v1 - v0) - constr.value()) < 1e-10
    assert abs((g1[0] - g0[0]) - (0.3 - 0.1) / 0.05**2) < 1e-8 and abs(g1[1] - g0[1]) < 1e-10


# ------------------------------------------------------------------------------------------------ minimizers
def test_golden_fit_through_host_layer_with_oracle_engine():
    d = np.load(GOLDEN / "deterministic_fit_data.npz", allow_pickle=True)
    stored = json.loads(str(d["fitted_params_json"]))
    obs = zfit.Space("x", -10, 10)
    mu = zfit.Parameter("mu", 0.5, -4, 6)
    sigma = zfit.Parameter("sigma", 1.2, 0.1, 10)
    lambd = zfit.Parameter("lambda", -0.05, -1, -0.01)
    n_bkg = zfit.Parameter("n_bkg", 20000, 0, 50000)
    n_sig = zfit.Parameter("n_sig", 1000, 0, 30000)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs, extended=n_sig),
                             zfit.pdf.Exponential(lambd, obs=obs, extended=n_bkg)])
    nll = use_oracle(zfit.loss.ExtendedUnbinnedNLL(model=model, data=zfit.Data(d["data"], obs=obs)))
    assert [p.name for p in nll.get_params()] == list(stored)
    result = zfit.minimize.Minuit(gradient="zfit").minimize(nll).update_params()
    assert result.valid and result.converged and result.edm < 1e-3
    errs = result.hesse()
    for p in result.params:
        assert abs(result.params[p]["value"] - stored[p.name]) < 0.1 * errs[p]["error"], p.name
    assert abs(result.fmin - 62323.84) < 0.05
    assert "driver" in result.info and result.info["n_eval"] < 200
    assert str(result).startswith("FitResult")
    assert abs(result.params["mu"]["value"] - mu.value()) == 0.0


@pytest.mark.parametrize("factory", [
    lambda: zfit.minimize.Minuit(),                      # the driver's own numerical gradient (Minuit default)
    lambda: zfit.minimize.Minuit(gradient="zfit"),
    lambda: zfit.minimize.ScipyLBFGSB(),
    lambda: zfit.minimize.ScipyTrustConstr(),
    lambda: zfit.minimize.ScipySLSQP(),
])
def test_minimizers_acceptance_problem(factory):
    """tests/minimize/test_minimizer.py:24-48: Gauss + Exponential SumPDF (frac 0.8), mu, sigma, lambda within 0.1."""
    rng = np.random.default_rng(5)
    n = 6000
    sig = rng.normal(1.0, 0.5, int(0.8 * n))
    u = rng.uniform(0, 1, n - sig.size)
    lam_true, lo, hi = -0.4, -5.0, 5.0  # truncated exponential by inverse CDF
    bkg = lo + np.log1p(u * np.expm1(lam_true * (hi - lo))) / lam_true
    obs = zfit.Space("obs1", lo, hi)
    tag = f"_{abs(hash(factory)) % 10**6}"
    mu = zfit.Parameter("mu" + tag, 0.6, -2, 3)
    sigma = zfit.Parameter("sigma" + tag, 0.8, 0.1, 3)
    lam = zfit.Parameter("lambda" + tag, -0.7, -3, -0.01)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Exponential(lam, obs)], fracs=0.8)
    nll = use_oracle(zfit.loss.UnbinnedNLL(model, zfit.Data(np.concatenate([sig, bkg]), obs=obs)))
    result = factory().minimize(nll)
    assert result.valid, result.message
    assert abs(mu.value() - 1.0) < 0.1 and abs(sigma.value() - 0.5) < 0.1 and abs(lam.value() - lam_true) < 0.1


def test_nan_strategy_and_maxiter():
    a = zfit.Parameter("a_nan", 1.0)

    def f(params):
        x = params[0].value()
        return float("nan") if x > 1.5 else (x - 3.0) ** 2

    loss = zfit.loss.SimpleLoss(f, [a], errordef=0.5)
    ev = zfit.minimize.LossEval(loss, [a], zfit.minimize.PushbackStrategy(), False, maxiter=50)
    assert ev.value(np.array([1.0])) == 4.0
    assert ev.value(np.array([2.0])) == 4.0 + 100  # pushed back: last value + penalty (strategy.py:104-126)
    assert ev.nan_counter == 1
    ev2 = zfit.minimize.LossEval(loss, [a], zfit.minimize.PushbackStrategy(), False, maxiter=2)
    ev2.value(np.array([1.0]))
    ev2.value(np.array([1.0]))
    with pytest.raises(zexc.MaximumIterationReached):
        ev2.value(np.array([1.0]))


def test_errors_profile_interval():
    rng = np.random.default_rng(7)
    obs = zfit.Space("x", -6, 6)
    mu, sigma = zfit.Parameter("mu_er", 0.2, -2, 2), zfit.Parameter("sigma_er", 1.1, 0.3, 4)
    nll = use_oracle(zfit.loss.UnbinnedNLL(zfit.pdf.Gauss(mu, sigma, obs), zfit.Data(rng.normal(0.0, 1.0, 400), obs=obs)))
    result = zfit.minimize.Minuit(gradient="zfit", tol=1e-6).minimize(nll)
    h = result.hesse()
    e, _ = result.errors(params=[mu])
    assert abs(h[mu]["error"] - 1.0 / np.sqrt(400)) < 0.01
    assert abs(e[mu]["upper"] - h[mu]["error"]) < 0.1 * h[mu]["error"] and abs(e[mu]["lower"] + h[mu]["error"]) < 0.1 * h[mu]["error"]
    cov = result.covariance()
    assert cov.shape == (2, 2) and abs(cov[0, 0] - h[mu]["error"] ** 2) < 1e-12


def test_weighted_covariance_correction_host_logic():
    """covariance_with_weights (errors.py:333-464) through the host layer: C from the engine's per-event score outer
    product, mapped slots -> theta (incl. the log-yield term of extended fits), cov = H^-1 C H^-1."""
    rng = np.random.default_rng(12)
    obs = zfit.Space("x", -6, 6)
    x = np.concatenate([rng.normal(0.3, 1.0, 1500), rng.uniform(-6, 6, 1500)])
    w = rng.normal(1.0, 0.5, x.size)
    mu, sigma = zfit.Parameter("mu_wc", 0.2, -2, 2), zfit.Parameter("sigma_wc", 1.1, 0.3, 4)
    lam = zfit.Parameter("lam_wc", -0.05, -1, 1)
    ns, nb = zfit.Parameter("ns_wc", 1500.0, 0, 1e5), zfit.Parameter("nb_wc", 1500.0, 0, 1e5)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs, extended=ns), zfit.pdf.Exponential(lam, obs, extended=nb)])
    data = zfit.Data(x, obs=obs, weights=w)
    nll = use_oracle(zfit.loss.ExtendedUnbinnedNLL(model, data))
    params = list(nll.get_params())
    C = nll.weighted_score_outer(params)
    # independent: finite-difference Jacobian of v_i = log pdf(x_i) + log(yield) in theta space
    from oracle.backend import NumpyBackend

    NB = NumpyBackend()
    tree = O.SumPDF([O.Gauss("mu", "sigma", "x", (-6, 6), yield_="ns"), O.Exponential("lam", "x", (-6, 6), yield_="nb")])
    names = ["ns", "nb", "mu", "sigma", "lam"]
    th = np.array([p.value() for p in params])

    def v(t):
        P = dict(zip(names, map(np.float64, t)))
        return np.log(tree.pdf(NB, {"x": data.value("x")}, P)) + np.log(P["ns"] + P["nb"])

    J = np.zeros((len(th), data.num_entries))
    for i in range(len(th)):
        h = 1e-6 * max(1.0, abs(th[i]))
        tp, tm = th.copy(), th.copy()
        tp[i] += h
        tm[i] -= h
        J[i] = (v(tp) - v(tm)) / (2 * h)
    Cref = (J * data.weights**2) @ J.T
    np.testing.assert_allclose(C, Cref, rtol=2e-6, atol=1e-8 * np.abs(Cref).max())
    result = zfit.minimize.Minuit(gradient="zfit", tol=1e-6).minimize(nll)
    cov_corr = result.covariance()
    cov_none = result.covariance(weightcorr=False)
    cov_sumw2 = result.covariance(weightcorr="sumw2")
    assert np.all(np.diag(cov_corr) > 0) and not np.allclose(cov_corr, cov_none)
    np.testing.assert_allclose(cov_sumw2, cov_none * (np.sum(w**2) / np.sum(w)), rtol=1e-12)
    h = result.hesse()
    assert abs(h[mu]["error"] - np.sqrt(cov_corr[2, 2])) < 1e-12


def test_sum_pdf_with_explicit_norm_fit_range():
    """fit_range / explicit norm on a SumPDF: sum_i f_i pdf_i(x) / sum_j f_j int_norm pdf_j (functor.py:189-206,
    loss.py:112-114), lowered as leaves normalised over the range + renormalised composed fractions."""
    from oracle.backend import NumpyBackend

    NB = NumpyBackend()
    rng = np.random.default_rng(21)
    obs = zfit.Space("x", -10, 10)
    x = np.concatenate([rng.normal(1.0, 1.5, 1500), rng.exponential(4.0, 2500) - 10])
    mu, sigma = zfit.Parameter("mu_fr", 0.8, -5, 5), zfit.Parameter("sigma_fr", 1.7, 0.3, 6)
    lam, frac = zfit.Parameter("lam_fr", -0.2, -2, 2), zfit.Parameter("frac_fr", 0.4, 0, 1)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Exponential(lam, obs)], fracs=frac)
    data = zfit.Data(x, obs=obs)
    with pytest.warns(UserWarning, match="fit_range"):
        nll = use_oracle(zfit.loss.UnbinnedNLL(model, data, fit_range=(-4, 6)))
    params = list(nll.get_params())
    v, g = nll.value_gradient(params, full=True)
    # direct restatement of the reference route on the cut data
    xin = data.value("x")
    xin = xin[(xin >= -4) & (xin <= 6)]
    og, oe = O.Gauss("mu", "sigma", "x", (-10, 10)), O.Exponential("lam", "x", (-10, 10))

    def ref(theta):
        P = dict(zip(["frac", "mu", "sigma", "lam"], map(np.float64, theta)))
        f = [P["frac"], 1 - P["frac"]]
        num = f[0] * og.pdf(NB, {"x": xin}, P) + f[1] * oe.pdf(NB, {"x": xin}, P)
        den = f[0] * og.integral(NB, P, -4, 6, (-10, 10)) / og.integral(NB, P, -10, 10, (-10, 10)) + f[1] * oe.integral(
            NB, P, -4, 6, (-10, 10)) / oe.integral(NB, P, -10, 10, (-10, 10))
        return float(-np.sum(np.log(num / den + 1e-307)))

    th = [p.value() for p in params]
    assert [p.name for p in params] == ["frac_fr", "mu_fr", "sigma_fr", "lam_fr"]
    assert abs(v - ref(th)) < 1e-9 * abs(v)
    np.testing.assert_allclose(g, ON.numeric_gradient(ref, th), rtol=1e-6, atol=1e-6 * np.abs(g).max())


def test_constraints_values_and_partials():
    from scipy import stats

    a, b = zfit.Parameter("a_con", 4.2, 0, 100), zfit.Parameter("b_con", 7.5, 0, 100)
    g = zfit.constraint.GaussianConstraint([a, b], observation=[4.0, 8.0], cov=[[0.04, 0.01], [0.01, 0.09]])
    assert abs(g.value() + stats.multivariate_normal([4.2, 7.5], [[0.04, 0.01], [0.01, 0.09]]).logpdf([4.0, 8.0])) < 1e-12
    g1 = zfit.constraint.GaussianConstraint([a, b], observation=[4.0, 8.0], cov=[0.04, 0.09])  # 1-D: diagonal of the covariance
    g2 = zfit.constraint.GaussianConstraint([a, b], observation=[4.0, 8.0], sigma=[0.2, 0.3])
    assert abs(g1.value() - g2.value()) < 1e-14
    p = zfit.constraint.PoissonConstraint([a, b], observation=[5.0, 6.0])
    assert abs(p.value() + np.sum(stats.gamma.logpdf([5.0, 6.0], 1) * 0 + np.array([4.2, 7.5]) * np.log([5.0, 6.0]) - np.array([5.0, 6.0])
                                   - np.array([stats.loggamma.logpdf(0, 1) * 0 + __import__("scipy").special.gammaln(v + 1) for v in (4.2, 7.5)]))) < 1e-12
    ln = zfit.constraint.LogNormalConstraint([a], observation=[1.4], uncertainty=[0.2])
    assert abs(ln.value() + stats.lognorm.logpdf(4.2, s=0.2, scale=np.exp(1.4))) < 1e-12
    for con in (g, p, ln):
        num = {}
        for leaf in con.get_params():
            v = leaf.value()
            leaf.assign(v + 1e-6)
            up = con.value()
            leaf.assign(v - 1e-6)
            dn = con.value()
            leaf.assign(v)
            num[leaf] = (up - dn) / 2e-6
        ana = con._partials()
        for leaf in num:
            assert abs(ana[leaf] - num[leaf]) < 1e-6 * max(1.0, abs(num[leaf]))
    with pytest.raises(ValueError):
        zfit.constraint.GaussianConstraint([a], observation=[1.0], sigma=[0.1], cov=[0.01])


def test_update_data_in_place_invalidates_binding_and_offset():
    """``Data.update_data`` (what ``SamplerData.resample`` calls, data.py:1031-1060, 1563-1599): a loss holding the Data
    sees the new events, and its automatic offset is recomputed (``is_precompiled``, loss.py:247-262)."""
    rng = np.random.default_rng(5)
    obs = zfit.Space("x", -5, 5)
    mu, sigma = zfit.Parameter("mu_upd", 0.2, -2, 2), zfit.Parameter("sigma_upd", 1.1, 0.2, 4)
    model = zfit.pdf.Gauss(mu, sigma, obs)
    data = zfit.Data(rng.normal(0.0, 1.0, 3000), obs=obs)
    nll = use_oracle(zfit.loss.UnbinnedNLL(model, data))
    h0 = data.hashint
    v0 = nll.value(full=False)
    off0 = nll.options["subtr_const_value"]
    assert abs(v0 - 10000.0) < 1e-6
    x2 = rng.normal(0.5, 1.3, 2500)
    data.update_data(x2)
    assert data.hashint != h0 and data.num_entries == np.sum(np.abs(x2) <= 5)
    v1 = nll.value(full=False)
    assert abs(v1 - 10000.0) < 1e-6 and nll.options["subtr_const_value"] != off0  # offset recomputed for the new sample
    fresh = use_oracle(zfit.loss.UnbinnedNLL(model, zfit.Data(x2, obs=obs)))
    assert abs(nll.value(full=True) - fresh.value(full=True)) < 1e-9
    g1, g2 = nll.gradient(), fresh.gradient()
    np.testing.assert_allclose(g1, g2, rtol=1e-12)
    with pytest.raises(ValueError, match="same length"):
        data.update_data(x2, weights=np.ones(3))


def test_plan_step_path_marshalling(tmp_path, monkeypatch):
    """``Plan.eval / launch / collect`` keep persistent argument buffers and cached ctypes pointers (step-path overhead);
    checked here without a GPU against a C stub with the signatures of ``znll_eval`` / ``znll_launch`` / ``znll_collect``
    (``include/znll.h``) that echoes what it received."""
    import ctypes as C
    import subprocess

    src = tmp_path / "stub.c"
    src.write_text(
        """
        static double seen[16]; static unsigned seen_flags; static double seen_off; static void* seen_stream; static int nslots = 3;
        int stub_eval(void* p, const double* s, unsigned f, double off, double* v, double* g, void* st) {
          if (p != (void*)0x1234) return 1;
          for (int i = 0; i < nslots; ++i) g[i] = 10.0 * s[i];
          v[0] = off + f; v[1] = (double)(long)st; return 0; }
        int stub_launch(void* p, const double* s, unsigned f, double off, void* st) {
          for (int i = 0; i < nslots; ++i) seen[i] = s[i];
          seen_flags = f; seen_off = off; seen_stream = st; return 0; }
        int stub_collect(void* p, unsigned f, double* v, double* g, void* st) {
          if (f != seen_flags || st != seen_stream) return 1;
          for (int i = 0; i < nslots; ++i) g[i] = -seen[i];
          v[0] = seen_off; v[1] = 7.0; return 0; }
        int stub_fail(void* p, const double* s, unsigned f, double off, double* v, double* g, void* st) { return 1; }
        """
    )
    so = tmp_path / "libstub.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    stub = C.CDLL(str(so))
    L = Z.lib()
    for name, sym in (("stub_eval", "znll_eval"), ("stub_launch", "znll_launch"), ("stub_collect", "znll_collect"),
                      ("stub_fail", "znll_eval")):
        fn = getattr(stub, name)
        fn.restype, fn.argtypes = Z.SYMBOLS[sym]
    monkeypatch.setattr(L, "znll_eval", stub.stub_eval, raising=False)
    monkeypatch.setattr(L, "znll_launch", stub.stub_launch, raising=False)
    monkeypatch.setattr(L, "znll_collect", stub.stub_collect, raising=False)
    monkeypatch.setattr(L, "znll_plan_total_slots", lambda h: 3, raising=False)
    plan = Z.Plan.__new__(Z.Plan)  # no device: only the argument path is under test
    plan._h, plan.n_channels, plan._keep = C.c_void_p(0x1234), 2, []
    try:
        plan._buffers()
        for slots in (np.array([1.0, 2.0, 3.0]), [1.0, 2.0, 3.0], np.array([1, 2, 3]), np.arange(6.0)[::2] * 1.0 + 1 - np.arange(3.0)):
            v, g = plan.eval(slots, grad=True, log_offset=0.25, stream=77)
            assert v.tolist() == [0.25 + Z.WANT_GRAD, 77.0]
            assert g.tolist() == [10.0, 20.0, 30.0]
        v0, g0 = plan.eval([4.0, 5.0, 6.0], grad=False, log_offset=np.float64(1.5), kahan=True)
        assert g0 is None and v0.tolist() == [1.5 + Z.KAHAN, 0.0]
        assert v.tolist() == [0.25 + Z.WANT_GRAD, 77.0]  # results handed out earlier are copies
        plan.launch(np.array([7.0, 8.0, 9.0]), grad=True, log_offset=2.0, stream=5)
        v, g = plan.collect(grad=True, stream=5)
        assert v.tolist() == [2.0, 7.0] and g.tolist() == [-7.0, -8.0, -9.0]
        with pytest.raises(ValueError, match="expected 3 slots"):
            plan.eval([1.0, 2.0])
        monkeypatch.setattr(L, "znll_eval", stub.stub_fail, raising=False)
        plan._buffers()
        with pytest.raises(Z.ZnllError):
            plan.eval([1.0, 2.0, 3.0])
    finally:
        plan._h = C.c_void_p()  # nothing to destroy


def test_data_from_root_reads_branches_through_uproot(monkeypatch):
    """``Data.from_root`` (``core/data.py:612-698``): obs / obs_alias -> branches, weights by branch name, inclusive cut;
    ``uproot`` is optional and absent here, so a stand-in module records what is asked of it."""
    import sys
    import types

    r = np.random.default_rng(5)
    branches = {"B_M": r.uniform(5000, 5600, 50), "pt": r.uniform(0, 10, 50), "sw": r.normal(1, 0.3, 50)}
    asked = {}

    class _Tree:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def arrays(self, expressions, library):
            asked["expressions"], asked["library"] = tuple(expressions), library
            return {k: branches[k] for k in expressions}

    class _File(dict):
        pass

    def _open(path, **options):
        asked["path"], asked["options"] = path, options
        return _File({"DecayTree": _Tree()})

    with pytest.raises(ImportError, match="uproot"):
        zfit.Data.from_root("f.root", "DecayTree", obs="B_M")
    monkeypatch.setitem(sys.modules, "uproot", types.SimpleNamespace(open=_open))
    obs = zfit.Space("mass", 5100, 5500) * zfit.Space("pt", 0, 10)
    d = zfit.Data.from_root("f.root", "DecayTree", obs=obs, weights="sw", obs_alias={"mass": "B_M"}, root_dir_options={"timeout": 3})
    assert asked == {"path": "f.root", "options": {"timeout": 3}, "expressions": ("B_M", "pt", "sw"), "library": "np"}
    keep = (branches["B_M"] >= 5100) & (branches["B_M"] <= 5500)
    assert d.obs == ("mass", "pt") and d.num_entries == keep.sum() and d.has_weights
    np.testing.assert_array_equal(d["mass"], branches["B_M"][keep])
    np.testing.assert_array_equal(d.weights, branches["sw"][keep])
    d2 = zfit.Data.from_root("f.root", "DecayTree", obs=["pt"])
    assert d2.obs == ("pt",) and d2.num_entries == 50 and not d2.has_weights
    with pytest.raises(ValueError):
        zfit.Data.from_root("f.root", "DecayTree")
    with pytest.raises(zexc.BreakingAPIChangeError):
        zfit.Data.from_root("f.root", "DecayTree", obs="pt", branches=["pt"])


def test_host_prologue_crystalball_integral_fuzz():
    """Seeded sweep over the CrystalBall integral's whole case structure (limits left of / right of / across the tail
    boundary, either sign of alpha, n near 1 where the reference switches to its log form, physics.py:83-142): the host
    prologue's closed form and its four analytic derivatives against the oracle's tape, and the integral itself against
    adaptive quadrature of the oracle's shape."""
    from scipy.integrate import quad

    r = np.random.default_rng(20261020)
    B = TorchBackend()
    import torch

    checked_branches = set()
    for _ in range(120):
        mu, s = r.uniform(-2, 2), r.uniform(0.3, 3.0)
        a = r.uniform(0.2, 3.0) * r.choice([-1.0, 1.0])
        n = 1.0 + r.choice([3e-6, -4e-6, 1e-3]) if r.uniform() < 0.25 else r.uniform(1.05, 8.0)
        edge = mu - abs(a) * s if a > 0 else mu + abs(a) * s  # where core and tail meet
        kind = r.integers(3)
        if kind == 0:  # across the boundary
            lo, hi = edge - r.uniform(0.5, 6) * s, edge + r.uniform(0.5, 6) * s
        elif (kind == 1) == (a > 0):  # tail only
            lo, hi = sorted([edge - np.sign(a) * r.uniform(0.1, 3) * s, edge - np.sign(a) * r.uniform(3.5, 8) * s])
        else:  # core only
            lo, hi = sorted([edge + np.sign(a) * r.uniform(0.1, 2) * s, edge + np.sign(a) * r.uniform(2.5, 7) * s])
        checked_branches.add((int(kind), bool(a > 0), bool(abs(n - 1) < 1e-5)))
        I, dI = Z.leaf_integral(Z.Node(Z.CB, 0, 0, 0, lo, hi, 0, 0, 0), [mu, s, a, n])
        opdf = O.CrystalBall("mu", "s", "a", "n", "x", (lo, hi))
        rI, rdI = _autograd_integral(opdf, ["mu", "s", "a", "n"], [mu, s, a, n], lo, hi, (lo, hi))
        assert abs(I - rI) <= 2e-13 * abs(rI), (mu, s, a, n, lo, hi)
        if abs(n - 1) >= 1e-5:  # the log form's tape differentiates a different expression: compared by value only
            np.testing.assert_allclose(dI, rdI, rtol=2e-9, atol=1e-12 * abs(rI), err_msg=str((mu, s, a, n, lo, hi)))
        P = {k: torch.tensor(v, dtype=torch.float64) for k, v in zip(["mu", "s", "a", "n"], [mu, s, a, n])}
        shape = lambda x: float(opdf.unnorm(B, torch.tensor([x], dtype=torch.float64), P, (lo, hi))[0])
        pts = [edge] if lo < edge < hi else None
        q, _ = quad(shape, lo, hi, points=pts, epsabs=0, epsrel=1e-11, limit=200)
        # inside |n - 1| < 1e-5 the reference integrates the tail as if n were exactly 1 (its `use_log` switch,
        # physics.py:88,123,132): mirrored on purpose, so there the closed form is off the true integral by O(|n - 1|)
        tol = 2e-9 if abs(n - 1) >= 1e-5 else 4.0 * abs(n - 1)
        assert abs(I - q) <= tol * abs(q), (mu, s, a, n, lo, hi, I, q)
    assert len(checked_branches) >= 10


def test_lazy_sort_countdown_of_the_engine():
    """``options={"sort_events": "lazy"}`` (experimental): the engine binds unsorted and sorts exactly once, after
    ``LAZY_SORT_AFTER`` evaluations; the default mode never takes that path.  Device work is stubbed out."""
    import importlib

    CudaEngine = importlib.import_module("zfit_b200.loss").CudaEngine

    class _Plan:
        def eval(self, slots, **kw):
            return np.zeros(1), np.zeros(len(slots))

    for mode, expect in (("lazy", 1), (None, 0)):
        eng = CudaEngine.__new__(CudaEngine)
        eng.plan, eng.distributed, eng.reduce_mode = _Plan(), False, "none"
        eng._raw_stream = lambda: 0
        eng._lazy_sort_in = CudaEngine.LAZY_SORT_AFTER if mode == "lazy" else None
        calls = []
        eng._sort_now = lambda: calls.append(eng._lazy_sort_in)
        for _ in range(CudaEngine.LAZY_SORT_AFTER + 30):
            eng.eval(np.zeros(3))
            if not calls:
                assert mode is None or eng._lazy_sort_in >= 0
        assert len(calls) == expect and eng._lazy_sort_in is None


def test_large_host_cut_stays_on_the_host_without_a_gpu_or_under_a_launcher(monkeypatch):
    """``data._upload_for_cut``: the HBM route for the cut of large host arrays needs a CUDA device, a single-process job and
    the ``device_cut_min_events`` option; otherwise the inclusive cut runs on the host exactly as for small arrays."""
    import importlib

    import zfit_b200 as zfit

    D = importlib.import_module("zfit_b200.data")  # (the package attribute `zfit.data` is the public namespace)
    rng = np.random.default_rng(4)
    cols = [rng.normal(size=1000)]
    monkeypatch.setitem(zfit.settings.options, "device_cut_min_events", 100)
    monkeypatch.setenv("WORLD_SIZE", "2")  # torchrun before init_process_group: every rank would upload everything
    assert D._upload_for_cut(cols, None) is None
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setitem(zfit.settings.options, "device_cut_min_events", None)
    assert D._upload_for_cut(cols, None) is None
    monkeypatch.setitem(zfit.settings.options, "device_cut_min_events", 100)
    import torch

    if not torch.cuda.is_available():
        assert D._upload_for_cut(cols, None) is None
    d = zfit.Data(cols[0], obs=zfit.Space("x", -1, 1)) if not torch.cuda.is_available() else None
    if d is not None:
        assert not d.on_device and d.num_entries == int(((cols[0] >= -1) & (cols[0] <= 1)).sum())
