"""The reference's own loss tests, re-expressed against the CUDA loss (behavioural spec: tests/test_loss.py).

Same problems, same statistical tolerances as upstream; the arithmetic under test is the fused kernel."""

import numpy as np
import pytest

import zfit_b200 as zfit
from zfit_b200.minimize import Minuit

pytestmark = pytest.mark.gpu

rng = np.random.default_rng(519)  # tests/conftest.py:20 seeds with 519
mu_true, sigma_true = 1.2, 4.1
mu_true2, sigma_true2 = 1.01, 3.5
yield_true = 3000
test_values_np = rng.normal(loc=mu_true, scale=sigma_true, size=(yield_true, 1))
test_values_np2 = rng.normal(loc=mu_true2, scale=sigma_true2, size=yield_true)
obs1 = zfit.Space("obs1", (min(test_values_np.min(), test_values_np2.min()) - 1.4, max(test_values_np.max(), test_values_np2.max()) + 2.4))
mu_constr, sigma_constr = [1.6, 0.02], [3.5, 0.01]

_n = {"i": 0}


def _tag():
    _n["i"] += 1
    return f"_rb{_n['i']}"


def create_params(prefix):
    t = _tag()
    mu = zfit.Parameter(prefix + "mu" + t, mu_true - 0.2, mu_true - 1.0, mu_true + 1.0)
    sigma = zfit.Parameter(prefix + "sigma" + t, sigma_true - 0.3, sigma_true - 2.0, sigma_true + 2.0)
    return mu, sigma


def create_gauss(prefix):
    mu, sigma = create_params(prefix)
    return zfit.pdf.Gauss(mu, sigma, obs=obs1, name="gaussian" + prefix), mu, sigma


@pytest.mark.parametrize("size", [None, 5000, 50000, 300000, 3_000_000])
def test_extended_unbinned_nll(size):  # tests/test_loss.py:133-156
    if size is None:
        values = test_values_np
        size = values.shape[0]
    else:
        values = np.random.default_rng(size).normal(mu_true, sigma_true, size=(size, 1))
    data = zfit.Data.from_tensor(obs=obs1, tensor=values)
    gauss, mu3, sigma3 = create_gauss("3")
    yield3 = zfit.Parameter("yield35" + _tag(), yield_true + 300, 0, 10000000)
    gaussian3 = gauss.create_extended(yield3)
    with pytest.warns(UserWarning, match="fit_range"):
        nll = zfit.loss.ExtendedUnbinnedNLL(model=gaussian3, data=data, fit_range=(-20, 20))
    assert {mu3, sigma3, yield3} == nll.get_params()
    status = Minuit(tol=1e-4).minimize(loss=nll)  # Minuit's default: the driver's own numerical gradient
    params = status.params
    inside = values[(values >= -20) & (values <= 20)]
    assert pytest.approx(np.mean(data.value()), rel=0.05) == params[mu3]["value"]
    assert pytest.approx(np.std(data.value()), rel=0.05) == params[sigma3]["value"]
    assert pytest.approx(inside.size, rel=0.005) == params[yield3]["value"]


def test_unbinned_simultaneous_nll():  # tests/test_loss.py:159-169
    data1 = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np)
    data2 = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np2)
    gaussian1, mu1, sigma1 = create_gauss("1")
    gaussian2, mu2, sigma2 = create_gauss("2")
    gaussian2 = gaussian2.create_extended(zfit.Parameter("yield_gauss2" + _tag(), 5))
    with pytest.warns(UserWarning, match="Extended PDFs"):
        nll = zfit.loss.UnbinnedNLL(model=[gaussian1, gaussian2], data=[data1, data2])
    status = Minuit(tol=1e-5).minimize(loss=nll, params=[mu1, sigma1, mu2, sigma2])
    params = status.params
    assert set(nll.get_params()) == {mu1, mu2, sigma1, sigma2}
    assert pytest.approx(np.mean(test_values_np), rel=0.007) == params[mu1]["value"]
    assert pytest.approx(np.mean(test_values_np2), rel=0.007) == params[mu2]["value"]
    assert pytest.approx(np.std(test_values_np), rel=0.007) == params[sigma1]["value"]
    assert pytest.approx(np.std(test_values_np2), rel=0.007) == params[sigma2]["value"]


@pytest.mark.parametrize("weights", (None, np.random.default_rng(1).normal(loc=1.0, scale=0.2, size=yield_true)), ids=["noweights", "weights"])
@pytest.mark.parametrize("sigma", (lambda: [mu_constr[1] ** 2, sigma_constr[1] ** 2],
                                   lambda: np.array([[mu_constr[1] ** 2, 0], [0, sigma_constr[1] ** 2]])), ids=["cov1d", "cov2d"])
@pytest.mark.parametrize("options", ({"subtr_const": False}, {"subtr_const": True}))
def test_unbinned_nll(weights, sigma, options):  # tests/test_loss.py:172-212
    gaussian1, mu1, sigma1 = create_gauss("1")
    gaussian2, mu2, sigma2 = create_gauss("2")
    data = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np, weights=weights)
    nll_object = zfit.loss.UnbinnedNLL(model=gaussian1, data=data, options=options)
    status = Minuit(tol=1e-5).minimize(loss=nll_object, params=[mu1, sigma1])
    params = status.params
    rel_error = 0.12 if weights is None else 0.1
    assert pytest.approx(np.mean(test_values_np), rel=rel_error) == params[mu1]["value"]
    assert pytest.approx(np.std(test_values_np), rel=rel_error) == params[sigma1]["value"]
    constraints = zfit.constraint.GaussianConstraint(params=[mu2, sigma2], observation=[mu_constr[0], sigma_constr[0]], cov=sigma())
    nll_object = zfit.loss.UnbinnedNLL(model=gaussian2, data=data, constraints=constraints, options=options)
    status = Minuit(tol=1e-4).minimize(loss=nll_object, params=[mu2, sigma2])
    params = status.params
    if weights is None:
        assert params[mu2]["value"] > np.average(test_values_np, weights=weights)
        assert params[sigma2]["value"] < np.std(test_values_np)


def test_add():  # tests/test_loss.py:215-275
    params = [create_params(str(i)) for i in range(4)]
    pdfs = [zfit.pdf.Gauss(mu, sigma, obs=obs1) for mu, sigma in params]
    datas = [zfit.Data.from_tensor(obs=obs1, tensor=rng.normal(1.0, 3.0, size=500)) for _ in range(4)]
    nll = zfit.loss.UnbinnedNLL(model=pdfs[0], data=datas[0])
    nll2 = zfit.loss.UnbinnedNLL(model=pdfs[1], data=datas[1])
    nll3 = zfit.loss.UnbinnedNLL(model=[pdfs[2], pdfs[3]], data=[datas[2], datas[3]])
    simult = nll + nll2 + nll3
    assert simult.model == pdfs and simult.data == datas
    assert abs(simult.value(full=True) - (nll.value(full=True) + nll2.value(full=True) + nll3.value(full=True))) < 1e-9
    assert len(simult.get_params()) == 8
    with pytest.raises(ValueError):
        _ = nll + zfit.loss.ExtendedUnbinnedNLL(pdfs[0].create_extended(zfit.Parameter("y" + _tag(), 400.0)), datas[0])


@pytest.mark.parametrize("numgrad", [True, False], ids=["numgrad", "analytic"])
def test_gradients(numgrad):  # tests/test_loss.py:278-335
    t = _tag()
    initial1, initial2 = 1.0, 2.0
    param1, param2 = zfit.Parameter("param1" + t, initial1), zfit.Parameter("param2" + t, initial2)
    gauss1 = zfit.pdf.Gauss(param1, 4, obs=obs1)
    gauss2 = zfit.pdf.Gauss(param2, 5, obs=obs1)
    data1 = zfit.Data.from_tensor(obs=obs1, tensor=np.full(100, 1.0))
    data2 = zfit.Data.from_tensor(obs=obs1, tensor=np.full(100, 1.0))
    nll = zfit.loss.UnbinnedNLL(model=[gauss1, gauss2], data=[data1, data2])

    def loss_funcparam1and2(values):
        with zfit.param.set_values([param2, param1], values):
            return nll.value()

    def central(f, x0, i, h=1e-4):
        xp, xm = np.array(x0, float), np.array(x0, float)
        xp[i] += h
        xm[i] -= h
        d1 = (f(xp) - f(xm)) / (2 * h)
        xp[i] += h
        xm[i] -= h
        d2 = (f(xp) - f(xm)) / (4 * h)
        return (4 * d1 - d2) / 3

    truth = np.array([central(loss_funcparam1and2, [initial2, initial1], i) for i in range(2)])
    gradient1 = nll.gradient(params=param2, numgrad=numgrad)
    np.testing.assert_allclose(gradient1, truth[:1], rtol=1e-6)
    gradient2 = nll.gradient(params=[param2, param1], numgrad=numgrad)
    np.testing.assert_allclose(gradient2, truth, rtol=1e-5)
    params3 = list(nll.get_params())
    gradient3 = nll.gradient(params3)
    expected = [gradient2[[param2, param1].index(p)] for p in params3]  # order of get_params defines the layout
    np.testing.assert_allclose(gradient3, expected, rtol=1e-5)


def test_simple_loss():  # tests/test_loss.py:338-420
    t = _tag()
    true_a, true_b, true_c = 1.0, 4.0, -0.3
    a = zfit.Parameter("variable_a15151" + t, 1.5, -1.0, 20.0, stepsize=0.1)
    b = zfit.Parameter("variable_b15151" + t, 3.5, 0, 20)
    c = zfit.Parameter("variable_c15151" + t, -0.23, -1, 1)
    params = [a, b, c]

    def loss_func(params):
        a, b, c = (p.value() for p in params)
        return (a - true_a) ** 2 + (b - true_b) ** 2 + (c - true_c) ** 4 + 0.42

    loss_func.errordef = 1.0
    loss = zfit.loss.SimpleLoss(func=loss_func, params=params)
    assert loss.errordef == 1.0 and set(loss.get_params()) == set(params)
    result = Minuit(gradient="zfit").minimize(loss=loss)
    assert result.valid
    assert abs(result.params[a]["value"] - true_a) < 0.05 and abs(result.params[b]["value"] - true_b) < 0.05  # EDM < 1e-3
    assert abs(result.params[c]["value"] - true_c) < 0.2
    assert abs(result.fmin - 0.42) < 1e-3


def test_callable_loss_and_iminuit_protocol():  # tests/test_loss.py:501-547
    gaussian1, mu1, sigma1 = create_gauss("1")
    data = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np)
    nll = zfit.loss.UnbinnedNLL(model=gaussian1, data=data)
    x = np.array([p.value() for p in nll.get_params()])
    assert nll(x) == pytest.approx(nll.value(full=True))
    assert nll(x + 0.1) == pytest.approx(nll.value(params=dict(zip(nll.get_params(), x + 0.1)), full=True))
    with pytest.raises(ValueError):
        nll(x[:1])
    assert nll.errordef == 0.5
    # any driver that only knows f(x) and grad(x) can minimise it (what raw iminuit.Minuit(loss, x) does upstream)
    from scipy import optimize

    ev = zfit.minimize.LossEval(nll, list(nll.get_params()), zfit.minimize.DefaultStrategy(), False, maxiter=500, numpy_converter=np.array)
    opt = optimize.minimize(ev.value, x, jac=ev.gradient, method="BFGS")
    assert opt.success or opt.status == 2
    assert opt.x[0] == pytest.approx(np.mean(test_values_np), rel=0.01) and opt.x[1] == pytest.approx(np.std(test_values_np), rel=0.01)


def test_exact_yields_toys():  # tests/highstats/test_exact_yields.py:12-125 (fewer toys; the fits run on the GPU)
    obs = zfit.Space("x", -10, 10)
    t = _tag()
    mu = zfit.Parameter("mu" + t, 1.0, -4, 6)
    sigma = zfit.Parameter("sigma" + t, 1.0, 0.1, 10)
    lambd = zfit.Parameter("lambda" + t, -0.06, -1, -0.01)
    n_bkg = zfit.Parameter("n_bkg" + t, 20000, 0, 500000)
    n_sig = zfit.Parameter("n_sig" + t, 1000, 0, 300000)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs, extended=n_sig), zfit.pdf.Exponential(lambd, obs, extended=n_bkg)])
    toy_rng = np.random.default_rng(7)
    true_vals = [20000.0, 1000.0, 1.0, 1.0, -0.06]
    minimizer = Minuit(gradient=False, tol=1e-05)  # as upstream: the loss' own gradient, tight tolerance
    for _ in range(15):
        n = toy_rng.poisson(21000)
        nsig = toy_rng.binomial(n, 1000 / 21000)
        sig = toy_rng.normal(1.0, 1.0, nsig)
        u = toy_rng.uniform(0, 1, n - nsig)
        bkg = -10 + np.log1p(u * np.expm1(-0.06 * 20)) / -0.06
        data = zfit.Data(np.concatenate([sig, bkg]), obs=obs)
        zfit.param.set_values([n_bkg, n_sig, mu, sigma, lambd], true_vals)
        nll = zfit.loss.ExtendedUnbinnedNLL(model=model, data=data)
        result = minimizer.minimize(nll)
        assert result.valid
        fitted = result.params[n_sig]["value"] + result.params[n_bkg]["value"]
        assert abs(fitted - data.num_entries) < 0.6  # the extended term pins the total yield to N


def test_weights_one_equal_no_weights():  # tests/result/test_uncertainties.py:180,671
    data = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np)
    dataw = zfit.Data.from_tensor(obs=obs1, tensor=test_values_np, weights=np.ones(yield_true))
    gaussian1, mu1, sigma1 = create_gauss("1")
    a = zfit.loss.UnbinnedNLL(gaussian1, data)
    b = zfit.loss.UnbinnedNLL(gaussian1, dataw)
    va, ga = a.value_gradient(full=True)
    vb, gb = b.value_gradient(full=True)
    assert va == vb and np.array_equal(ga, gb)
    assert b.is_weighted and not a.is_weighted


def test_exponential_numerics_far_from_zero():  # tests/test_pdfs.py:302-321
    t = _tag()
    lam = zfit.Parameter("lambda" + t, -0.3, -5, 5)
    obs = zfit.Space("obs1", 5000, 6000)
    exp1 = zfit.pdf.Exponential(lam, obs=obs)
    x = np.linspace(5000, 6000, 200001)
    p = exp1.pdf(x)  # exp(-0.3 * 5500) would underflow without the data shift of basic.py:128-154
    assert np.all(np.isfinite(p)) and np.all(p >= 0)
    assert abs(np.trapezoid(p, x) - 1.0) < 1e-3
    data = zfit.Data(5000 + np.random.default_rng(3).exponential(1 / 0.3, 10000), obs=obs)
    nll = zfit.loss.UnbinnedNLL(exp1, data)
    v, g = nll.value_gradient(full=True)
    assert np.isfinite(v) and np.all(np.isfinite(g))
    result = Minuit(gradient="zfit").minimize(nll)
    assert abs(result.params[lam]["value"] + 0.3) < 0.02


def test_product_pdf_values_are_product_of_factors():  # tests/test_pdfs.py:204-227
    t = _tag()
    ox, oy, oz = zfit.Space("x", -4, 4), zfit.Space("y", 0, 5), zfit.Space("z", -2, 4)
    g = zfit.pdf.Gauss(zfit.Parameter("mu" + t, 0.3), zfit.Parameter("sigma" + t, 1.2), ox)
    e = zfit.pdf.Exponential(zfit.Parameter("lam" + t, -0.7), oy)
    c = zfit.pdf.Chebyshev(oz, [zfit.Parameter("c1" + t, 0.2), zfit.Parameter("c2" + t, -0.1)])
    prod = zfit.pdf.ProductPDF([g, e, c])
    r = np.random.default_rng(0)
    pts = np.stack([r.uniform(-4, 4, 500), r.uniform(0, 5, 500), r.uniform(-2, 4, 500)], axis=1)
    data = zfit.Data(pts, obs=ox * oy * oz)
    expected = g.pdf(pts[:, 0]) * e.pdf(pts[:, 1]) * c.pdf(pts[:, 2])
    np.testing.assert_allclose(prod.pdf(data), expected, rtol=1e-12)
    np.testing.assert_allclose(prod.pdf(data.with_obs(["z", "x", "y"])), expected, rtol=1e-12)  # obs are matched by name
    assert abs(float(prod.integrate(ox * oy * oz)[0]) - 1.0) < 1e-12


# ------------------------------------------------------------------------------------------------
# the flows of the reference's example scripts (examples/*.py), on seeded data, checked against the truth
# ------------------------------------------------------------------------------------------------
def test_example_flow_extended_fit_with_param_update_disabled():  # examples/signal_bkg_mass_extended_fit.py
    t = _tag()
    r = np.random.default_rng(11)
    obs = zfit.Space("x", -10, 10)
    mu = zfit.Parameter("mu" + t, 1.0, -4, 6)
    sigma = zfit.Parameter("sigma" + t, 1.0, 0.1, 10)
    lambd = zfit.Parameter("lambda" + t, -0.06, -1, -0.01)
    n_bkg = zfit.Parameter("n_bkg" + t, 20000, 0, 50000)
    n_sig = zfit.Parameter("n_sig" + t, 1000, 0, 30000)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs, extended=n_sig),
                             zfit.pdf.Exponential(lambd, obs=obs, extended=n_bkg)])
    bkg = -10.0 - np.log1p(-r.uniform(size=20000) * (1.0 - np.exp(-0.06 * 20.0))) / 0.06  # exp(-0.06 x) on (-10, 10)
    sample = np.concatenate([r.normal(1.0, 1.0, size=1200), bkg])
    sample = sample[(sample > -10) & (sample < 10)]
    data = zfit.Data.from_numpy(obs=obs, array=sample)
    start = {mu: 0.5, sigma: 1.2, lambd: -0.05}
    zfit.param.set_values(start)
    nll = zfit.loss.ExtendedUnbinnedNLL(model=model, data=data)
    with zfit.run.experimental_disable_param_update(True):
        result = Minuit(gradient="zfit").minimize(nll)
        assert [p.value() for p in start] == list(start.values())  # the fit left the parameters alone
        assert result.valid
        assert result.update_params() is result
    assert zfit.settings.options["auto_update_params"] is True
    assert mu.value() == result.params[mu]["value"]
    hesse = result.hesse()
    errors, _ = result.errors()
    n = len(sample)
    # extended fit: the yields sum to the sample size (to 0.6 at tol=1e-5, tests/highstats/test_exact_yields.py:82; this fit
    # runs at the default tol=1e-3, where EDM < 1e-3 allows a few events)
    assert n_sig.value() + n_bkg.value() == pytest.approx(n, abs=5.0)
    for p, truth in ((mu, 1.0), (sigma, 1.0), (lambd, -0.06), (n_sig, 1200.0), (n_bkg, n - 1200.0)):
        err = hesse[p]["error"]
        assert abs(p.value() - truth) < 4.5 * err, (p.name, p.value(), truth, err)
        assert errors[p]["lower"] < 0 < errors[p]["upper"]
        assert 0.5 * (errors[p]["upper"] - errors[p]["lower"]) == pytest.approx(err, rel=0.15)
    zfit.param.set_values(model.get_params(), result)  # set_values accepts a FitResult
    assert sigma.value() == result.params[sigma]["value"]


def test_example_flow_simultaneous_fit_by_loss_addition():  # examples/simultaneous_fit.py
    t = _tag()
    r = np.random.default_rng(12)
    obs = zfit.Space("x", -10, 10)
    mu_shared = zfit.Parameter("mu_shared" + t, 1.0, -4, 6)
    sigma1 = zfit.Parameter("sigma_one" + t, 1.0, 0.1, 10)
    sigma2 = zfit.Parameter("sigma_two" + t, 1.0, 0.1, 10)
    gauss1 = zfit.pdf.Gauss(mu=mu_shared, sigma=sigma1, obs=obs)
    gauss2 = zfit.pdf.Gauss(mu=mu_shared, sigma=sigma2, obs=obs)
    data1 = zfit.Data.from_numpy(obs=obs, array=r.normal(2.0, 3.0, size=10000))
    data2 = zfit.Data.from_numpy(obs=obs, array=r.normal(2.0, 4.0, size=10000))
    together = zfit.loss.UnbinnedNLL(model=[gauss1, gauss2], data=[data1, data2])
    added = zfit.loss.UnbinnedNLL(model=gauss1, data=data1) + zfit.loss.UnbinnedNLL(model=gauss2, data=data2)
    assert list(added.get_params()) == [mu_shared, sigma1, sigma2]
    assert added.value() == pytest.approx(together.value(), rel=1e-13)
    result = Minuit().minimize(added).update_params()
    result.hesse()
    result.errors()
    assert result.valid
    for p, truth in ((mu_shared, 2.0), (sigma1, 3.0), (sigma2, 4.0)):
        assert abs(p.value() - truth) < 4.5 * result.params[p]["hesse"]["error"]
        assert result.params[p]["errors"]["upper"] == pytest.approx(result.params[p]["hesse"]["error"], rel=0.1)
    assert "mu_shared" in str(result)


def test_example_flow_multidim_product_from_dataframe():  # examples/multidim_preprocess_fit.py
    pd = pytest.importorskip("pandas")
    t = _tag()
    r = np.random.default_rng(13)
    xobs, yobs, zobs = zfit.Space("xobs", -4, 4), zfit.Space("yobs", -3, 5), zfit.Space("z", -2, 4)
    obs = xobs * yobs * zobs
    mu1 = zfit.Parameter("mu1" + t, 1.0, -4, 6)
    mu23 = zfit.Parameter("mu_shared" + t, 1.0, -4, 6)
    sigma12 = zfit.Parameter("sigma_shared" + t, 1.0, 0.1, 10)
    sigma3 = zfit.Parameter("sigma3" + t, 1.0, 0.1, 10)
    product = zfit.pdf.ProductPDF([zfit.pdf.Gauss(mu=mu1, sigma=sigma12, obs=xobs), zfit.pdf.Gauss(mu=mu23, sigma=sigma12, obs=yobs),
                                   zfit.pdf.Gauss(mu=mu23, sigma=sigma3, obs=zobs)])
    raw = zfit.Data(r.normal(loc=[2.0, 2.5, 2.5], scale=[3.0, 3.0, 1.5], size=(10000, 3)), obs=obs)  # cut to the space
    df = raw.to_pandas()
    assert isinstance(df, pd.DataFrame) and list(df.columns) == ["xobs", "yobs", "z"] and len(df) == raw.num_entries < 10000
    nll = zfit.loss.UnbinnedNLL(model=product, data=df)  # a dataframe is taken directly
    result = Minuit().minimize(nll).update_params()
    errors, _ = result.errors()
    assert result.valid
    for p, truth in ((mu1, 2.0), (mu23, 2.5), (sigma12, 3.0), (sigma3, 1.5)):
        assert abs(p.value() - truth) < 4.5 * 0.5 * (errors[p]["upper"] - errors[p]["lower"]), (p.name, p.value())
