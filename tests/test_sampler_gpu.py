"""On-device toy generation: ``pdf.create_sampler`` / ``SamplerData.resample`` (reference: core/sample.py:148-566,
core/data.py:1247-1606) -- distribution checks against closed forms and a toy-study loop on one loss."""

import numpy as np
import pytest
from scipy import stats

import zfit_b200 as zfit

pytestmark = pytest.mark.gpu


def test_sampler_draws_from_the_pdf_and_stays_on_the_device():
    obs = zfit.Space("x", -3, 5)
    mu, sigma = zfit.Parameter("mu_smp", 0.7, -2, 3), zfit.Parameter("sigma_smp", 1.2, 0.3, 4)
    lam, frac = zfit.Parameter("lam_smp", -0.3, -2, 2), zfit.Parameter("frac_smp", 0.6, 0, 1)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Exponential(lam, obs)], fracs=frac)
    sampler = model.create_sampler(n=400_000)
    sampler._sampler.seed(1234)
    sampler.resample()
    assert sampler.on_device and sampler.num_entries == 400_000 and sampler.n == 400_000
    x = sampler.value("x")
    assert x.min() >= -3 and x.max() <= 5
    # Kolmogorov-Smirnov against the model's own cdf (numerical integral of the device pdf values on a grid)
    grid = np.linspace(-3, 5, 4001)
    pdf = model.pdf(zfit.Data(grid, obs=obs))
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(grid))])
    ks = stats.kstest(x, lambda v: np.interp(v, grid, cdf / cdf[-1]))
    assert ks.pvalue > 1e-3, ks
    # fixed_params: changing the parameters afterwards does not change what is sampled; overrides per call do
    mu.set_value(2.5)
    sampler.resample()
    assert abs(np.mean(sampler.value("x")) - np.mean(x)) < 0.02
    sampler.resample({mu: -1.0}, n=200_000)
    assert sampler.num_entries == 200_000 and np.mean(sampler.value("x")) < np.mean(x) - 0.5
    with pytest.raises(ValueError, match="both"):
        sampler.resample({mu: 0.0}, param_values={mu: 0.0})


def test_extended_sampler_fluctuates_the_yield_and_multidim_product():
    ox, oy = zfit.Space("x", -4, 4), zfit.Space("y", 0, 5)
    mu, sigma, lam = zfit.Parameter("mu_s2", 0.3, -2, 2), zfit.Parameter("sigma_s2", 1.0, 0.3, 3), zfit.Parameter("lam_s2", -0.7, -3, -0.01)
    nyield = zfit.Parameter("n_s2", 50_000, 0, 1e6)
    model = zfit.pdf.ProductPDF([zfit.pdf.Gauss(mu, sigma, ox), zfit.pdf.Exponential(lam, oy)], extended=nyield)
    sampler = model.create_sampler()
    sizes = []
    for _ in range(6):
        sampler.resample()
        sizes.append(sampler.num_entries)
    assert len(set(sizes)) > 1 and abs(np.mean(sizes) - 50_000) < 5 * np.sqrt(50_000 / 6) + 1
    xy = sampler.value()
    assert xy.shape[1] == 2 and abs(np.mean(xy[:, 0]) - 0.3) < 0.03
    # E[y] of a truncated exponential with rate 0.7 on (0, 5)
    r = 0.7
    ey = 1 / r - 5 * np.exp(-r * 5) / (1 - np.exp(-r * 5))
    assert abs(np.mean(xy[:, 1]) - ey) < 0.03
    with pytest.raises(ValueError, match="n has to be given"):
        zfit.pdf.Gauss(mu, sigma, ox).create_sampler()


def test_toy_study_loop_on_one_loss():
    """The reference's toy pattern: one loss on a sampler, resample + reset + minimize per toy; pulls are standard normal."""
    obs = zfit.Space("x", -6, 6)
    mu, sigma = zfit.Parameter("mu_toy", 0.4, -3, 3), zfit.Parameter("sigma_toy", 1.5, 0.3, 5)
    model = zfit.pdf.Gauss(mu, sigma, obs)
    sampler = model.create_sampler(n=20_000)
    nll = zfit.loss.UnbinnedNLL(model, sampler)
    minimizer = zfit.minimize.Minuit(gradient="zfit")
    engine = None
    pulls = []
    for toy in range(30):
        sampler.resample()
        zfit.param.set_values([mu, sigma], [0.1, 1.2])
        result = minimizer.minimize(nll)
        assert result.valid
        if engine is None:
            engine = nll._get_engine()
        assert nll._get_engine() is engine, "the plan is re-bound, not rebuilt"
        err = np.sqrt(np.diag(result.covariance()))
        pulls.append((result.params[mu]["value"] - 0.4) / err[0])
        x = sampler.value("x")
        assert abs(result.params[mu]["value"] - x.mean()) < 5e-3  # nearly the sample mean (slight truncation at +-6)
    pulls = np.array(pulls)
    assert abs(pulls.mean()) < 0.75 and 0.5 < pulls.std() < 1.6
