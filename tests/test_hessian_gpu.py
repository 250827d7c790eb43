"""The exact one-pass Hessian on the GPU (``hess_kernel`` through ``znll_eval_hess``): against torch's double backward over
the oracle graph at the C ABI, and through ``loss.hessian`` / ``FitResult.hesse`` for every BASELINE config."""

import numpy as np
import pytest

from tests.test_hessian_cpu import CASES, _oracle_hessian

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_cabi_hessian_matches_double_backward(name):
    from zfit_b200 import _znll as Z

    case = CASES[name]()
    plan = Z.Plan(1)
    plan.set_model(0, case.nodes)
    plan.bind_host(0, case.cols, case.weights)
    assert plan.has_hess()
    values, g, mats = plan.eval_hess(case.slots)
    v1, g1 = plan.eval(case.slots, grad=True)
    assert abs(values[0] - v1[0]) <= 1e-12 * abs(v1[0])
    np.testing.assert_allclose(g, g1, rtol=1e-9, atol=1e-9 * np.abs(g1).max())
    idx, ov, og, oH = _oracle_hessian(case, case.weights)
    H = mats[0][np.ix_(idx, idx)]
    scale = np.sqrt(np.abs(np.outer(np.diag(oH), np.diag(oH))))
    assert np.all(np.abs(H - oH) <= 1e-7 * scale + 1e-9 * np.abs(oH).max()), np.abs(H - oH).max() / np.abs(oH).max()
    plan.close()


@pytest.mark.parametrize("cfg,n", [(1, 10_000), (2, 200_000), (3, 100_000), (4, 200_000), (5, 200_000)])
def test_loss_hessian_one_pass(cfg, n):
    import zfit_b200 as zfit
    from zfit_b200 import workloads as W

    w = W.build(cfg, n=n, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    launches0 = loss._get_engine().launch_count() if loss._engine is not None else 0
    H = loss.hessian(params)
    assert loss.last_hessian_method == "one-pass"
    eng = loss._get_engine()
    loss.options["hessian"] = "numeric"
    before = eng.launch_count()
    Hn = loss.hessian(params)
    passes_numeric = eng.launch_count() - before
    assert passes_numeric >= 2 * len(params)  # what the one pass replaces
    scale = np.sqrt(np.abs(np.outer(np.diag(Hn), np.diag(Hn))))
    assert np.all(np.abs(H - Hn) <= 2e-5 * scale + 1e-7 * np.abs(Hn).max()), np.abs(H - Hn).max()
    loss.options["hessian"] = "exact"
    before = eng.launch_count()
    loss.hessian(params)
    assert eng.launch_count() - before == len(loss.model)  # ONE launch per channel
    # the covariance of a fit built on it is symmetric positive definite at the minimum
    res = zfit.minimize.Minuit(gradient="zfit").minimize(loss)
    cov = res.covariance(method="hesse_np", weightcorr=False)
    assert np.all(np.linalg.eigvalsh(cov) > 0)


def _two_dim_parts(tag):
    import zfit_b200 as zfit

    rng = np.random.default_rng(3)
    ox, oy = zfit.Space("x", -5, 5), zfit.Space("y", 0, 4)
    mu, sg = zfit.Parameter(f"mu_{tag}", 0.1, -1, 1), zfit.Parameter(f"sg_{tag}", 1.1, 0.3, 3)
    l1, l2 = zfit.Parameter(f"l1_{tag}", -0.5, -3, -0.01), zfit.Parameter(f"l2_{tag}", -1.0, -3, -0.01)
    fr = zfit.Parameter(f"fr_{tag}", 0.4, 0, 1)
    data = zfit.Data(np.stack([rng.normal(0, 1.3, 5000).clip(-5, 5), rng.uniform(0, 4, 5000)], axis=1), obs=ox * oy)
    return zfit, ox, oy, mu, sg, l1, l2, fr, data


def test_sum_of_products_has_the_one_pass_hessian():
    """signal(x) signal(y) + background(x) background(y): every branch of the Sum is a Product of leaves, which the
    Hessian pass treats as one generalised leaf (leaf blocks + outer products of the leaves' scores)."""
    zfit, ox, oy, mu, sg, l1, l2, fr, data = _two_dim_parts("sp")
    a = zfit.pdf.ProductPDF([zfit.pdf.Gauss(mu, sg, ox), zfit.pdf.Exponential(l1, oy)])
    b = zfit.pdf.ProductPDF([zfit.pdf.Gauss(mu, 2.0, ox), zfit.pdf.Exponential(l2, oy)])
    loss = zfit.loss.UnbinnedNLL(zfit.pdf.SumPDF([a, b], fracs=fr), data)
    H = loss.hessian()
    assert loss.last_hessian_method == "one-pass"
    loss.options["hessian"] = "numeric"
    Hn = loss.hessian()
    scale = np.sqrt(np.abs(np.outer(np.diag(Hn), np.diag(Hn))))
    assert np.all(np.abs(H - Hn) <= 2e-5 * scale + 1e-7 * np.abs(Hn).max()), np.abs(H - Hn).max()


def test_deeper_tree_falls_back_to_gradient_differences():
    """A Product with a Sum among its factors is outside the shapes of the Hessian pass."""
    zfit, ox, oy, mu, sg, l1, l2, fr, data = _two_dim_parts("dp")
    sx = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sg, ox), zfit.pdf.Gauss(mu, 2.0, ox)], fracs=fr)
    loss = zfit.loss.UnbinnedNLL(zfit.pdf.ProductPDF([sx, zfit.pdf.Exponential(l1, oy)]), data)
    H = loss.hessian()
    assert loss.last_hessian_method == "gradient-differences"
    assert np.all(np.isfinite(H)) and np.allclose(H, H.T)
