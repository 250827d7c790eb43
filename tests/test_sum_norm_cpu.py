"""A SumPDF whose own ``norm`` differs from its daughters' (reference: ``SumPDF._pdf`` raises
``SpecificFunctionNotImplemented`` when the norm ranges are not all equal, models/functor.py:197-201, and the value becomes
``_unnormalized_pdf / normalization(norm)``): the lowering must renormalise even when no explicit range is passed."""

import numpy as np
from scipy import stats

import zfit_b200 as zfit
from tests._hostsim import HostSim


def test_sum_pdf_with_its_own_norm_is_renormalised():
    obs = zfit.Space("x", -10, 10)
    mu, sg = zfit.Parameter("mu_sn", 0.3), zfit.Parameter("sg_sn", 1.2)
    lam, fr = zfit.Parameter("lam_sn", -0.2), zfit.Parameter("fr_sn", 0.4)
    g, e = zfit.pdf.Gauss(mu, sg, obs), zfit.pdf.Exponential(lam, obs)
    model = zfit.pdf.SumPDF([g, e], fracs=fr, norm=(-2, 2))
    nodes, slots = model._lower(model.obs, None)
    assert [(n.norm_lo, n.norm_hi) for n in nodes[1:]] == [(-2.0, 2.0), (-2.0, 2.0)]  # leaves over the sum's norm
    x = np.linspace(-2, 2, 401)
    sim = HostSim(["Sum<Gauss<0>,Expo<0>>"])
    got = sim.pdf(nodes, [p.value() for p in slots], [x])

    def daughters(v):  # each normalised over ITS norm (-10, 10)
        pg = stats.norm.pdf(v, 0.3, 1.2) / (stats.norm.cdf(10, 0.3, 1.2) - stats.norm.cdf(-10, 0.3, 1.2))
        pe = np.exp(-0.2 * v) / ((np.exp(-0.2 * 10) - np.exp(0.2 * 10)) / -0.2)
        return pg, pe

    pg, pe = daughters(x)
    ig = (stats.norm.cdf(2, 0.3, 1.2) - stats.norm.cdf(-2, 0.3, 1.2)) / (stats.norm.cdf(10, 0.3, 1.2) - stats.norm.cdf(-10, 0.3, 1.2))
    ie = (np.exp(-0.2 * 2) - np.exp(0.2 * 2)) / (np.exp(-0.2 * 10) - np.exp(0.2 * 10))
    want = (0.4 * pg + 0.6 * pe) / (0.4 * ig + 0.6 * ie)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert abs(np.trapezoid(got, x) - 1.0) < 1e-4
    assert abs(float(model.integrate((-2, 2))[0]) - 1.0) < 1e-12
    # equal norms keep the plain lowering
    plain = zfit.pdf.SumPDF([g, e], fracs=fr)
    nodes, slots = plain._lower(plain.obs, None)
    assert [(n.norm_lo, n.norm_hi) for n in nodes[1:]] == [(-10.0, 10.0), (-10.0, 10.0)] and slots[0] is fr
