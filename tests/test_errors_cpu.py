"""``FitResult.errors`` = the reference's ``compute_errors`` (minimizers/errors.py:50-277): one root search over all
parameters per interval end on ``loss.value_gradient``, and a ``new_result`` when the scan meets a lower minimum."""

import warnings

import numpy as np
import pytest

import zfit_b200 as zfit
from tests._oracle_engine import use_oracle
from zfit_b200.minimize import FitResult


def _problem(tag, n=4000, seed=3):
    rng = np.random.default_rng(seed)
    obs = zfit.Space("x", -10, 10)
    x = np.concatenate([rng.normal(1.0, 1.5, int(0.4 * n)), rng.uniform(-10, 10, n - int(0.4 * n))])
    mu = zfit.Parameter("mu" + tag, 0.8, -5, 5)
    sigma = zfit.Parameter("sigma" + tag, 1.7, 0.3, 6)
    frac = zfit.Parameter("frac" + tag, 0.45, 0, 1)
    lam = zfit.Parameter("lam" + tag, 0.01, -1, 1)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Exponential(lam, obs)], fracs=frac)
    return use_oracle(zfit.loss.UnbinnedNLL(model, zfit.Data(x, obs=obs)))


def test_zfit_errors_agree_with_profile_refits_and_hesse():
    loss = _problem("_e1")
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-5).minimize(loss)
    hesse = res.hesse()
    calls0 = loss._get_engine().ncalls
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        errs, new = res.errors()
    assert new is None and res.valid
    calls = loss._get_engine().ncalls - calls0
    refit, _ = res.errors(method="refit", name="refit")
    for p in res.params:
        for side in ("lower", "upper"):
            # the reference stops once the profile is within rtol = 0.003 * errordef of the target: ~0.3 % of a sigma^2
            assert abs(errs[p][side] - refit[p][side]) < 0.02 * hesse[p]["error"], (p.name, side, errs[p], refit[p])
            assert abs(abs(errs[p][side]) - hesse[p]["error"]) < 0.15 * hesse[p]["error"]
        assert errs[p]["lower"] < 0 < errs[p]["upper"] and res.params[p]["errors"] is errs[p]
    assert calls < 400  # value+gradient passes, no inner fits
    with pytest.raises(ValueError):
        res.errors(cl=0.9, sigma=2)
    two_sigma, _ = res.errors(params=list(res.params)[:1], sigma=2, name="two")
    p0 = list(res.params)[0]
    assert 1.7 < two_sigma[p0]["upper"] / errs[p0]["upper"] < 2.4


def test_errors_return_a_new_result_when_a_lower_minimum_is_met():
    loss = _problem("_e2")
    good = zfit.minimize.Minuit(gradient="zfit", tol=1e-5).minimize(loss)
    hesse = good.hesse()
    params = list(good.params)
    # a "result" that sits 1.5 sigma away from the minimum in one parameter: profiling walks through lower values
    shifted = {p: good.params[p]["value"] + (1.5 * hesse[p]["error"] if p is params[0] else 0.0) for p in params}
    zfit.param.set_values(params, list(shifted.values()))
    fake = FitResult(loss=loss, params=shifted, minimizer=zfit.minimize.Minuit(gradient="zfit", tol=1e-5), valid=True,
                     converged=True, edm=1e-6, fminopt=float(loss.value(full=False)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        errs, new = fake.errors(params=[params[0]])  # the lower end lies beyond the true minimum: the scan walks through it
    assert new is not None and new.valid
    assert not fake.valid and "new minimum" in fake.message
    assert isinstance(fake.params[params[0]]["errors"], str)
    for p in params:
        assert abs(new.params[p]["value"] - good.params[p]["value"]) < 0.05 * hesse[p]["error"]
    assert new.params[params[0]]["errors"]["lower"] < 0 < new.params[params[0]]["errors"]["upper"]
