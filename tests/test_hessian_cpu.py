"""The exact one-pass Hessian (``hess_kernel`` + ``znll_eval_hess``) on the CPU: the kernel's per-event code
compiled for the host (tests/_hostsim.py) feeds the product library's own host assembly, and the result is held against
torch's double backward over the oracle's op graph -- the reference's route (``loss.hessian``, core/loss.py:765-878, goes
through ``z/math.py:288-352``: a second differentiation of the whole graph)."""

import numpy as np
import pytest
import torch

from oracle import nll as ON
from oracle.backend import TorchBackend
from tests import _cases as CS
from tests._hostsim import HostSim


def _oracle_hessian(case, weights):
    idx, names, theta = case.free()
    B = TorchBackend()
    data = [{k: B.asarray(v) for k, v in case.data.items()}]
    w = None if weights is None else [B.asarray(weights)]

    def f(th):
        P = {n: th[i] for i, n in enumerate(names)}
        for n, v in zip(case.slot_names, case.slots):  # constant slots (a fixed c_0) stay numbers
            if n is None:
                continue
        return ON.unbinned_nll([case.oracle_model], data, P, weights=w, B=B)

    th = torch.tensor(theta, dtype=torch.float64)
    H = torch.autograd.functional.hessian(f, th).numpy()
    g = torch.autograd.functional.jacobian(f, th).numpy()
    return idx, float(f(th)), g, H


CASES = {
    "gauss": lambda: CS.gauss_case(n=4001),
    "cb_exp": lambda: CS.cb_exp_case(n=6001),
    "cb_exp_other_point": lambda: CS.cb_exp_case(n=6000, alpha=1.4, nn=3.2),
    "cb_exp_mirrored": lambda: CS.cb_exp_case(n=6000, alpha=-1.3, nn=2.2),
    "gauss_cheb": lambda: CS.gauss_cheb_case(n=6001),
    "gauss_cheb_weighted": lambda: CS.gauss_cheb_case(n=6000, weighted=True),
    "product": lambda: CS.product_case(n=6001),
    "gcb_exp": lambda: CS.gcb_exp_case(n=6001),
    "get_exp": lambda: CS.get_exp_case(False, n=6001),
    "gget_exp": lambda: CS.get_exp_case(True, n=6000),
    "tgauss": lambda: CS.tgauss_case(n=6001),
    "legendre": lambda: CS.poly_case("legendre", n=4000),
    "bernstein": lambda: CS.bernstein_case(n=4000),
    "three_leaves_outside_catalogue": lambda: CS.rtc_case(n=6001),  # on the GPU this tree is built by NVRTC
    "sum_of_products": lambda: CS.nested_case(n=6001),  # signal(x) signal(y) + background(x) background(y)
}


@pytest.mark.parametrize("name", list(CASES))
def test_one_pass_hessian_matches_double_backward(name):
    case = CASES[name]()
    sim = HostSim([case.kernel])
    got = sim.hess(case.nodes, case.slots, case.cols, case.weights)
    assert got is not None, f"{case.kernel} should have the one-pass Hessian"
    v, g, H = got
    idx, ov, og, oH = _oracle_hessian(case, case.weights)
    assert abs(v - ov) <= 1e-11 * abs(ov)
    np.testing.assert_allclose(g[idx], og, rtol=1e-8, atol=1e-8 * np.abs(og).max())
    Hs = H[np.ix_(idx, idx)]
    scale = np.sqrt(np.abs(np.outer(np.diag(oH), np.diag(oH))))
    assert np.all(np.abs(Hs - oH) <= 1e-7 * scale + 1e-9 * np.abs(oH).max()), np.abs(Hs - oH).max() / np.abs(oH).max()
    assert np.allclose(H, H.T, rtol=0, atol=0)


def test_unsupported_tree_reports_none():
    """A Product whose factor is itself a Sum has no one-pass Hessian (the pass knows a leaf, a Product of leaves, and a Sum
    of leaves and of Products of leaves): the caller falls back to gradient differences."""
    from zfit_b200 import _znll as Z

    nodes = [
        Z.Node(Z.PROD, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.SUM, 2, 0, 0, 0, 0, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, -5, 5, 0, 0, 0),
        Z.Node(Z.GAUSS, 0, 0, 0, -5, 5, 0, 0, 0),
        Z.Node(Z.EXP, 0, 1, 0, 0, 4, 0, 0, 0),
    ]
    slots = [0.4, 0.6, 0.1, 1.0, -0.2, 1.5, -0.8]
    name, _ = Z.test_prologue(nodes, slots)
    sim = HostSim([name])
    rng = np.random.default_rng(0)
    assert sim.hess(nodes, slots, [rng.normal(0, 1, 64).clip(-5, 5), rng.uniform(0, 4, 64)]) is None


@pytest.mark.parametrize("cfg,n", [(1, 4000), (2, 20_000), (3, 10_000), (4, 20_000), (5, 20_000)])
def test_loss_hessian_one_pass_against_gradient_differences(cfg, n):
    """API level, every BASELINE config: ``loss.hessian`` (one pass; chain rule through shared parameters, fractions from
    yields, the extended term) against central differences of the analytic gradient, both on the host-compiled kernels."""
    import zfit_b200 as zfit
    from tests._hostsim import HostSimEngine
    from zfit_b200 import workloads as W

    w = W.build(cfg, n=n, device="cpu")
    loss = w["loss"]
    loss._set_engine_factory(HostSimEngine)
    params = list(loss.get_params())
    H = loss.hessian(params)
    assert loss.last_hessian_method == "one-pass"
    loss.options["hessian"] = "numeric"
    Hn = loss.hessian(params)
    assert loss.last_hessian_method == "gradient-differences"
    scale = np.sqrt(np.abs(np.outer(np.diag(Hn), np.diag(Hn))))
    assert np.all(np.abs(H - Hn) <= 2e-5 * scale + 1e-7 * np.abs(Hn).max()), np.abs(H - Hn).max()
    assert np.array_equal(H, H.T)
    d = loss.hessian(params[:2], hessian="diag")
    assert d.shape == (2,)
    # a Gaussian constraint adds its own second derivative
    loss.options["hessian"] = "exact"
    con = zfit.constraint.GaussianConstraint(params[0], params[0].value() * 1.01, sigma=abs(params[0].value()) * 0.05 + 0.1)
    lc = type(loss)(model=w["model"], data=w["data"], constraints=con)
    lc._set_engine_factory(HostSimEngine)
    Hc = lc.hessian(params)
    assert abs((Hc - H)[0, 0] - 1.0 / (abs(params[0].value()) * 0.05 + 0.1) ** 2) <= 1e-6 * abs(Hc[0, 0])
