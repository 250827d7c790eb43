"""CPU twin of tests/test_api_gpu.py::test_config_fit_parity: the kernels' per-event code compiled for the host
(tests/_hostsim.py) against the oracle engine, both driven by Minuit(gradient="zfit") through the same host layer."""

import numpy as np
import pytest

import zfit_b200 as zfit
from tests._hostsim import HostSimEngine
from tests._oracle_engine import OracleEngine
from zfit_b200 import workloads as W


@pytest.mark.parametrize("cfg,n", [(2, 20_000), (4, 20_000), (5, 20_000)])
def test_config_fit_parity_hostsim(cfg, n):
    w = W.build(cfg, n=n, device="cpu")
    loss = w["loss"]
    loss._set_engine_factory(HostSimEngine)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    params = list(loss.get_params())
    start = [p.value() for p in params]
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(loss)
    fitted = np.array([res.params[p]["value"] for p in res.params])
    zfit.param.set_values(params, start)
    rres = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(ref)
    rfit = np.array([rres.params[p]["value"] for p in rres.params])
    assert res.valid and rres.valid
    assert np.all(np.abs(fitted - rfit) <= 1e-8 * np.maximum(1.0, np.abs(rfit))), (fitted, rfit)


def test_crystalball_sampler_follows_the_pdf():
    """The synthetic cfg 2 signal must be drawn from the model it is fitted with: Kolmogorov-Smirnov
    against scipy's crystalball truncated to the observable, and the tail share."""
    import torch
    from scipy import stats

    gen = W._gen(2, torch.device("cpu"))
    x = W.sample_crystalball(400_000, 5280.0, 20.0, 1.5, 3.0, 5000.0, 5600.0, gen, torch.device("cpu")).numpy()
    cb = stats.crystalball(beta=1.5, m=3.0, loc=5280.0, scale=20.0)
    lo, tot = cb.cdf(5000.0), cb.cdf(5600.0) - cb.cdf(5000.0)
    ks = stats.kstest(x, lambda v: (cb.cdf(v) - lo) / tot)
    assert ks.pvalue > 1e-3, ks
    tail = np.mean(x < 5280.0 - 1.5 * 20.0)
    expect = (cb.cdf(5250.0) - lo) / tot
    assert abs(tail - expect) < 5 * np.sqrt(expect * (1 - expect) / x.size)
