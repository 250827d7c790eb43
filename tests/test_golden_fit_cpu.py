"""CPU twin of ``test_api_gpu.py::test_golden_fit_matches_reference_fixture_and_oracle_fit``: the reference's only stored
golden on this path (``tests/truth/data/deterministic_fit_data.npz``, reference: ``tests/cli/test_plotting_extended.py:25-71``)
fitted by ``zfit.minimize.Minuit`` through the whole host layer with the kernel's event loop taken from the host-compiled
device header (``HostSimEngine``), against the fit driven by the oracle and against the stored reference parameters."""

import json
from pathlib import Path

import numpy as np

import zfit_b200 as zfit
from tests._hostsim import HostSimEngine
from tests._oracle_engine import OracleEngine

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_golden_fit_with_the_device_arithmetic():
    d = np.load(GOLDEN / "deterministic_fit_data.npz", allow_pickle=True)
    stored = json.loads(str(d["fitted_params_json"]))

    def setup(factory):
        obs = zfit.Space("x", -10, 10)
        mu = zfit.Parameter("mu", 0.5, -4, 6)
        sigma = zfit.Parameter("sigma", 1.2, 0.1, 10)
        lambd = zfit.Parameter("lambda", -0.05, -1, -0.01)
        n_bkg = zfit.Parameter("n_bkg", 20000, 0, 50000)
        n_sig = zfit.Parameter("n_sig", 1000, 0, 30000)
        model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs, extended=n_sig),
                                 zfit.pdf.Exponential(lambd, obs=obs, extended=n_bkg)])
        loss = zfit.loss.ExtendedUnbinnedNLL(model=model, data=zfit.Data(d["data"], obs=obs))
        loss._set_engine_factory(factory)
        return loss

    nll = setup(HostSimEngine)
    assert [p.name for p in nll.get_params()] == list(stored)  # ordering rule
    res = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(nll)
    assert res.valid
    assert nll._get_engine().kernel_names == ["Sum<Gauss<0>,Expo<0>>"]
    rres = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(setup(OracleEngine))
    errs = res.hesse()
    for p, rp in zip(res.params, rres.params):
        a, b = res.params[p]["value"], rres.params[rp]["value"]
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (p.name, a, b)  # north_star: fitted parameters to 1e-8
        # the stored reference fit stopped at Minuit's EDM tolerance: agreement well inside the uncertainty
        assert abs(a - stored[p.name]) < 0.05 * errs[p]["error"], (p.name, a, stored[p.name])
    assert abs(res.fmin - 62323.84) < 0.01
    assert abs(sum(res.params[p]["value"] for p in list(res.params)[:2]) - 21200) < 0.6  # test_exact_yields.py:82
