"""GPU parity tests through the zfit-facing API (loss.value / value_gradient / minimizers), against the same
host layer driven by the CPU oracle engine (tests/_oracle_engine.py)."""

import json
from pathlib import Path

import numpy as np
import pytest

import zfit_b200 as zfit
from tests._oracle_engine import OracleEngine
from zfit_b200 import workloads as W

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _twin(loss):
    """Same models/data/parameters, evaluated by the oracle engine."""
    twin = loss.create_new()
    twin._set_engine_factory(OracleEngine)
    return twin


@pytest.mark.parametrize("cfg,n", [(1, 10_000), (2, 100_000), (3, 60_000), (4, 100_000), (5, 100_000)])
def test_config_value_gradient_parity(cfg, n):
    w = W.build(cfg, n=n, device="cuda:0")
    loss = w["loss"]
    ref = _twin(loss)
    params = list(loss.get_params())
    assert [p.name for p in params] == [p.name for p in ref.get_params()]
    v_full, g = loss.value_gradient(params, full=True)
    r_full, rg = ref.value_gradient(params, full=True)
    nev = w["n_events"]
    # value(full=True): 1e-12 relative (north_star); the sum of |terms| is >= |value|
    assert abs(v_full - r_full) <= 1e-12 * max(abs(r_full), nev), (v_full, r_full)
    # each gradient component: 1e-10 relative to the gradient scale (components vanish at the minimum)
    scale = max(np.abs(rg).max(), 1.0)
    assert np.all(np.abs(g - rg) <= 1e-10 * np.maximum(np.abs(rg), scale)), (g, rg)
    # the optimiser-facing value (full=False) subtracts a constant fixed at the first call (loss.py:393-409)
    ref.options["subtr_const_value"] = loss.options["subtr_const_value"]
    v, _ = loss.value_gradient(params, full=False)
    r, _ = ref.value_gradient(params, full=False)
    assert abs(v - r) <= 1e-12 * max(abs(r_full), nev)
    # the offset brings the first value to ~1e4; the extended term adds the offset once more per channel (:1314-1316)
    expect = 10000.0 + (loss.options["subtr_const_value"] * len(loss.model) if loss.is_extended else 0.0)
    assert abs(v - expect) < 1e-9 * nev
    assert loss.value(full=False) == v and abs(loss.gradient(params) - g).max() <= 1e-10 * scale


def test_loss_call_protocol_and_param_subset():
    w = W.build(2, n=20_000, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    x = np.array([p.value() for p in params])
    assert loss(x) == loss.value(full=True)
    with pytest.raises(ValueError):
        loss(x[:-1])
    with pytest.raises(TypeError):
        loss({"a": 1})
    sub = [params[4], params[2]]
    g_all = loss.gradient(params)
    g_sub = loss.gradient(sub)
    assert g_sub[0] == g_all[4] and g_sub[1] == g_all[2]
    assert loss.errordef == 0.5 and loss.is_extended


def test_pdf_values_match_oracle():
    from oracle import pdfs as O
    from oracle.backend import NumpyBackend

    obs = zfit.Space("m", 5000, 5600)
    mu, sg = zfit.Parameter("mu_pv", 5279.0), zfit.Parameter("sg_pv", 21.0)
    al, nn = zfit.Parameter("al_pv", 1.3), zfit.Parameter("nn_pv", 2.7)
    lam, fr = zfit.Parameter("lam_pv", -0.0025), zfit.Parameter("fr_pv", 0.3)
    cb = zfit.pdf.CrystalBall(mu, sg, al, nn, obs)
    ex = zfit.pdf.Exponential(lam, obs)
    model = zfit.pdf.SumPDF([cb, ex], fracs=fr)
    x = np.linspace(5000, 5600, 1001)
    p = model.pdf(x)
    ocb = O.CrystalBall("mu", "sg", "al", "nn", "m", (5000, 5600))
    oex = O.Exponential("lam", "m", (5000, 5600))
    om = O.SumPDF([ocb, oex], fracs=["fr"])
    P = {"mu": 5279.0, "sg": 21.0, "al": 1.3, "nn": 2.7, "lam": -0.0025, "fr": 0.3}
    ref = om.pdf(NumpyBackend(), {"m": x}, {k: np.float64(v) for k, v in P.items()})
    np.testing.assert_allclose(p, ref, rtol=5e-14)
    assert abs(np.trapezoid(p, x) - 1.0) < 1e-4
    assert abs(float(model.integrate((5000, 5600))[0]) - 1.0) < 1e-12


def test_golden_fit_matches_reference_fixture_and_oracle_fit():
    """tests/truth/data/deterministic_fit_data.npz (reference: tests/cli/test_plotting_extended.py:25-71)."""
    d = np.load(GOLDEN / "deterministic_fit_data.npz", allow_pickle=True)
    stored = json.loads(str(d["fitted_params_json"]))

    def setup(tag):
        obs = zfit.Space("x", -10, 10)
        mu = zfit.Parameter("mu", 0.5, -4, 6)
        sigma = zfit.Parameter("sigma", 1.2, 0.1, 10)
        lambd = zfit.Parameter("lambda", -0.05, -1, -0.01)
        n_bkg = zfit.Parameter("n_bkg", 20000, 0, 50000)
        n_sig = zfit.Parameter("n_sig", 1000, 0, 30000)
        model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs, extended=n_sig),
                                 zfit.pdf.Exponential(lambd, obs=obs, extended=n_bkg)])
        return zfit.loss.ExtendedUnbinnedNLL(model=model, data=zfit.Data(d["data"], obs=obs))

    nll = setup("gpu")
    assert [p.name for p in nll.get_params()] == list(stored)  # ordering rule
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(nll)
    assert res.valid
    ref = setup("cpu")
    ref._set_engine_factory(OracleEngine)
    rres = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(ref)
    errs = res.hesse()
    for p, rp in zip(res.params, rres.params):
        a, b = res.params[p]["value"], rres.params[rp]["value"]
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (p.name, a, b)  # north_star: fitted parameters to 1e-8
        # the stored reference fit stopped at Minuit's EDM tolerance: agreement well inside the uncertainty
        assert abs(a - stored[p.name]) < 0.05 * errs[p]["error"], (p.name, a, stored[p.name])
    assert abs(res.fmin - 62323.84) < 0.01
    assert abs(sum(res.params[p]["value"] for p in list(res.params)[:2]) - 21200) < 0.6  # test_exact_yields.py:82


def test_simple_fit_example_config1():
    """examples/simple_fit.py: the fitted mu, sigma agree between the GPU loss and the oracle loss."""
    w = W.build(1, device="cuda:0")
    loss = w["loss"]
    ref = _twin(loss)
    start = [p.value() for p in loss.get_params()]
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(loss)
    fitted = [res.params[p]["value"] for p in res.params]
    zfit.param.set_values(list(loss.get_params()), start)
    rres = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(ref)
    rfit = [rres.params[p]["value"] for p in rres.params]
    assert res.valid and rres.valid
    np.testing.assert_allclose(fitted, rfit, rtol=1e-8)
    # Minuit's default numerical gradient and a scipy driver reach the same minimum through LossEval
    for minimizer in (zfit.minimize.Minuit(tol=1e-6), zfit.minimize.ScipyLBFGSB(tol=1e-6)):
        zfit.param.set_values(list(loss.get_params()), start)
        r2 = minimizer.minimize(loss)
        np.testing.assert_allclose([r2.params[p]["value"] for p in r2.params], fitted, rtol=2e-4)


@pytest.mark.parametrize("cfg,n", [(2, 100_000), (3, 60_000), (4, 100_000), (5, 100_000)])
def test_config_fit_parity(cfg, n):
    """north_star: fitted parameters agree to 1e-8.  Minuit(gradient="zfit", tol=1e-9) from identical start values on the
    GPU loss and on its oracle twin (same host layer, CPU restatement of the reference's op graph underneath)."""
    w = W.build(cfg, n=n, device="cuda:0")
    loss = w["loss"]
    ref = _twin(loss)
    params = list(loss.get_params())
    start = [p.value() for p in params]
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(loss)
    fitted = np.array([res.params[p]["value"] for p in res.params])
    zfit.param.set_values(params, start)
    rres = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(ref)
    rfit = np.array([rres.params[p]["value"] for p in rres.params])
    assert res.valid and rres.valid
    assert [p.name for p in res.params] == [p.name for p in rres.params]
    assert np.all(np.abs(fitted - rfit) <= 1e-8 * np.maximum(1.0, np.abs(rfit))), (fitted, rfit)
    assert abs(res.fmin - rres.fmin) <= 1e-9 * max(1.0, abs(rres.fmin))


def test_config2_fit_closes_on_the_generation_parameters():
    """true parameters = generation parameters.  A 2e6-event signal+background sample is fitted from the
    start values of the config; every parameter must come back within four Hesse errors of what generated the data
    (round 1's CrystalBall sampler inflated the tail share and failed this for alpha, n and the signal yield)."""
    n = 2_000_000
    w = W.build(2, n=n, device="cuda:0")
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-6).minimize(w["loss"])
    assert res.valid
    errs = res.hesse()
    truth = {"mu_c2": 5280.0, "sigma_c2": 20.0, "alpha_c2": 1.5, "n_c2": 3.0, "lambda_c2": -0.002,
             "n_sig_c2": 0.3 * n, "n_bkg_c2": 0.7 * n}
    for p in res.params:
        pull = (res.params[p]["value"] - truth[p.name]) / errs[p]["error"]
        assert abs(pull) < 4.0, (p.name, res.params[p]["value"], truth[p.name], errs[p]["error"])


def _sliced(data, sl):
    cols = {o: c[sl] for o, c in zip(data.obs, data._cols)}
    wts = None if data._weights is None else data._weights[sl]
    return zfit.Data(cols, obs=data.space, weights=wts, guarantee_limits=True)


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_full_size_properties(cfg):
    """BASELINE sizes of configs 3 (2 x 5e7), 4 (1e8, three columns) and 5 (1.25e8 weighted, one GPU's shard): bitwise
    run-to-run determinism, additivity over halves, and a 1e6-event subsample against the oracle."""
    import torch

    w = W.build(cfg, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    v1, g1 = loss.value_gradient(params, full=True)
    v2, g2 = loss.value_gradient(params, full=True)
    assert v1 == v2 and np.array_equal(g1, g2)
    assert np.isfinite(v1) and np.all(np.isfinite(g1))
    halves = []
    for part in (0, 1):
        datas = []
        for d in w["data"]:
            h = d.num_entries // 2
            h -= h & 1
            datas.append(_sliced(d, slice(0, h) if part == 0 else slice(h, None)))
        lh = zfit.loss.UnbinnedNLL(model=w["model"], data=datas)
        halves.append(lh.value_gradient(params, full=True))
    assert abs(v1 - (halves[0][0] + halves[1][0])) <= 1e-12 * abs(v1)
    np.testing.assert_allclose(g1, halves[0][1] + halves[1][1], rtol=0, atol=1e-10 * np.abs(g1).max())
    nsub = 1_000_000 // len(w["data"])
    sub = zfit.loss.UnbinnedNLL(model=w["model"], data=[_sliced(d, slice(0, nsub)) for d in w["data"]])
    a, ga = sub.value_gradient(params, full=True)
    b, gb = _twin(sub).value_gradient(params, full=True)
    assert abs(a - b) <= 1e-12 * abs(b)
    assert np.all(np.abs(ga - gb) <= 1e-10 * np.maximum(np.abs(gb), np.abs(gb).max()))
    del w, loss, sub
    torch.cuda.empty_cache()


def test_full_size_properties_1e8():
    """BASELINE size (1e8 events, config 2): properties that need no oracle pass over the full data."""
    import torch

    w = W.build(2, n=100_000_000, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    v1, g1 = loss.value_gradient(params, full=True)
    v2, g2 = loss.value_gradient(params, full=True)
    assert v1 == v2 and np.array_equal(g1, g2)  # bitwise deterministic
    assert np.isfinite(v1) and np.all(np.isfinite(g1))
    # additivity: the NLL of the whole sample is the sum of the NLLs of two halves (same Poisson term handling)
    data = w["data"][0]
    x = data._cols[0]
    halves = []
    for sl in (slice(0, 50_000_000), slice(50_000_000, None)):
        dh = zfit.Data(x[sl], obs=data.space, guarantee_limits=True)
        lh = zfit.loss.UnbinnedNLL(model=w["model"], data=dh)
        halves.append(lh.value_gradient(list(lh.get_params()), full=True))
    whole = zfit.loss.UnbinnedNLL(model=w["model"], data=data)
    wp = list(whole.get_params())
    vw, gw = whole.value_gradient(wp, full=True)
    assert abs(vw - (halves[0][0] + halves[1][0])) <= 1e-12 * abs(vw)
    np.testing.assert_allclose(gw, halves[0][1] + halves[1][1], rtol=0, atol=1e-10 * np.abs(gw).max())
    # a 1e6-event subsample agrees with the oracle (same parameters, per-event average)
    sub = zfit.Data(x[:1_000_000], obs=data.space, guarantee_limits=True)
    ls = zfit.loss.UnbinnedNLL(model=w["model"], data=sub)
    lr = _twin(ls)
    sp = list(ls.get_params())
    a, ga = ls.value_gradient(sp, full=True)
    b, gb = lr.value_gradient(sp, full=True)
    assert abs(a - b) <= 1e-12 * abs(b)
    assert np.all(np.abs(ga - gb) <= 1e-10 * np.maximum(np.abs(gb), np.abs(gb).max()))
    del x, data
    torch.cuda.empty_cache()


@pytest.mark.parametrize("cfg,n", [(5, 6_000), (4, 6_000), (2, 6_000), (3, 4_000)])
def test_weighted_score_outer_matches_oracle(cfg, n):
    """One fused pass (znll_eval_cov) against the reference-style per-event Jacobian (errors.py:406-452)."""
    w = W.build(cfg, n=n, device="cuda:0")
    loss = w["loss"]
    ref = _twin(loss)
    params = list(loss.get_params())
    C = loss.weighted_score_outer(params)
    R = ref.weighted_score_outer(params)
    scale = np.sqrt(np.outer(np.diag(R), np.diag(R)))
    assert np.all(np.abs(C - R) <= 1e-9 * scale + 1e-12 * np.abs(R).max()), np.abs(C - R).max()
    assert np.allclose(C, C.T, rtol=1e-13, atol=0)


def test_weighted_fit_asymptotic_covariance():
    w = W.build(5, n=200_000, device="cuda:0")
    loss = w["loss"]
    result = zfit.minimize.Minuit(gradient="zfit", tol=1e-5).minimize(loss)
    assert result.valid
    corr = result.covariance()
    plain = result.covariance(weightcorr=False)
    # sWeight-like weights with mean 1 and rms 0.6: the corrected errors are larger by roughly sum w^2 / sum w
    ratio = np.diag(corr) / np.diag(plain)
    assert np.all(ratio > 1.05) and np.all(ratio < 2.5)


def test_sum_pdf_with_fit_range_parity():
    """Deprecated fit_range on a SumPDF (loss.py:112-114): cut to the range and normalise over it."""
    rng = np.random.default_rng(21)
    obs = zfit.Space("x", -10, 10)
    x = np.concatenate([rng.normal(1.0, 1.5, 15000), rng.exponential(4.0, 25000) - 10])
    mu, sigma = zfit.Parameter("mu_frg", 0.8, -5, 5), zfit.Parameter("sigma_frg", 1.7, 0.3, 6)
    lam, frac = zfit.Parameter("lam_frg", -0.2, -2, 2), zfit.Parameter("frac_frg", 0.4, 0, 1)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Exponential(lam, obs)], fracs=frac)
    with pytest.warns(UserWarning, match="fit_range"):
        loss = zfit.loss.UnbinnedNLL(model, zfit.Data(x, obs=obs), fit_range=(-4, 6))
    with pytest.warns(UserWarning, match="fit_range"):
        ref = _twin(loss)
    params = list(loss.get_params())
    v, g = loss.value_gradient(params, full=True)
    r, rg = ref.value_gradient(params, full=True)
    assert abs(v - r) <= 1e-12 * abs(r)
    assert np.all(np.abs(g - rg) <= 1e-10 * np.maximum(np.abs(rg), np.abs(rg).max()))


def test_misaligned_view_and_device_born_data():
    import torch

    obs = zfit.Space("x", -5, 5)
    base = torch.randn(20_001, dtype=torch.float64, device="cuda:0").clamp_(-4.9, 4.9)
    view = base[1:]  # starts 8 bytes into the allocation: not 16-byte aligned
    mu, sigma = zfit.Parameter("mu_mis", 0.1, -1, 1), zfit.Parameter("sigma_mis", 1.1, 0.3, 3)
    g = zfit.pdf.Gauss(mu, sigma, obs)
    a = zfit.loss.UnbinnedNLL(g, zfit.Data(view, obs=obs)).value(full=True)
    b = zfit.loss.UnbinnedNLL(g, zfit.Data(view.cpu().numpy(), obs=obs)).value(full=True)
    assert a == b
