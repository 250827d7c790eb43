"""CPU twin of ``test_random_trees_gpu.py``: the same random trees (same seeds), evaluated through the product's model
lowering, prologue and epilogue with the kernel's event loop taken from the host-compiled device header
(``tests/_hostsim.py::HostSimEngine``), against the oracle engine.  Covers on a CPU box what the GPU test covers except
the NVRTC load and the launch itself: pre-order slot layout, per-leaf normalisation constants, the Sum / Product epilogue
of nested trees, the chain rule through composed parameters."""

import numpy as np
import pytest

from tests._hostsim import HostSimEngine
from tests._oracle_engine import OracleEngine
from tests._random_trees import build_case


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112])
def test_random_tree_device_code_matches_oracle(seed):
    n = 4_001  # odd: exercises the scalar tail event
    b, loss = build_case(seed, n)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    loss._set_engine_factory(HostSimEngine)
    params = list(loss.get_params())
    assert len(params) == b.k
    v, g = loss.value_gradient(params, full=True)
    r, rg = ref.value_gradient(params, full=True)
    assert np.isfinite(r) and np.all(np.isfinite(rg)), (loss.model, r)
    names = loss._get_engine().kernel_names
    assert abs(v - r) <= 1e-12 * max(abs(r), n), (v, r, names)
    scale = max(np.abs(rg).max(), 1.0)
    assert np.all(np.abs(g - rg) <= 1e-10 * np.maximum(np.abs(rg), scale)), (g, rg, names)


@pytest.mark.parametrize("name", ["Hermite", "Laguerre"])
def test_hermite_laguerre_through_the_api(name):
    """``zfit.pdf.Hermite`` / ``Laguerre`` (models/polynomials.py:600-869): lowering, normalisation integral and its
    derivative, device arithmetic (host-compiled) against the oracle restatement, and ``integrate`` against quadrature."""
    import zfit_b200 as zfit
    from scipy import integrate as si

    rng = np.random.default_rng(7)
    obs = zfit.Space("x", -6.0, 8.0)
    tag = name.lower()
    mu, sg = zfit.Parameter(f"mu_{tag}", 0.8, -5, 5), zfit.Parameter(f"sg_{tag}", 1.3, 0.3, 5)
    cs = [zfit.Parameter(f"c{k}_{tag}", v, -1, 1) for k, v in enumerate([0.12, -0.07, 0.05], start=1)]
    fr = zfit.Parameter(f"fr_{tag}", 0.3, 0, 1)
    poly = getattr(zfit.pdf, name)(obs, cs)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sg, obs), poly], fracs=fr)
    data = zfit.Data(np.concatenate([rng.normal(1, 1.2, 1500), rng.uniform(-6, 8, 3501)]), obs=obs)
    loss = zfit.loss.UnbinnedNLL(model, data)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    loss._set_engine_factory(HostSimEngine)
    params = list(loss.get_params())
    v, g = loss.value_gradient(params, full=True)
    r, rg = ref.value_gradient(params, full=True)
    assert abs(v - r) <= 1e-12 * abs(r)
    np.testing.assert_allclose(g, rg, rtol=0, atol=1e-10 * np.abs(rg).max())
    # the registered analytic integral (host prologue) against numerical quadrature of the shape
    z = lambda x: (2 * x + 6 - 8) / 14.0
    c = [1.0, 0.12, -0.07, 0.05]
    if name == "Hermite":
        shape = lambda x: np.polynomial.hermite.hermval(z(x), c)
    else:
        shape = lambda x: np.polynomial.laguerre.lagval(z(x), c)
    num = si.quad(shape, -2.0, 5.0)[0] / si.quad(shape, -6.0, 8.0)[0]
    assert abs(float(poly.integrate((-2.0, 5.0))[0]) - num) <= 1e-12


@pytest.mark.parametrize("seed", [102, 104, 106, 111, 112, 113, 115, 116])
def test_random_tree_one_pass_hessian_matches_double_backward(seed):
    """The random trees whose shape the Hessian pass knows (a leaf, a Product of leaves, a Sum of leaves and of Products of
    leaves -- seed 104 and 115 are sums of products): ``eval_hess`` at slot level (host-compiled ``hess_event`` + the
    library's assembly) against torch's double backward over the oracle graph.  (Differences of the analytic gradient are
    NOT a valid check here: a TruncatedGauss bound or a CrystalBall junction moves events across a kink.)"""
    import torch

    from oracle import nll as ON
    from oracle.backend import TorchBackend

    b, loss = build_case(seed, 3_001)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    loss._set_engine_factory(HostSimEngine)
    eng, oeng = loss._get_engine(), ref._get_engine()
    slots = np.array([p.value() for p in eng.slot_params], dtype=np.float64)
    got = eng.eval_hess(slots)
    assert got is not None, eng.kernel_names
    H = got[2][0]
    tree, cols, w, off, ns = oeng.channels[0]
    B = TorchBackend()
    names = [f"s{k}" for k in range(ns)]
    data = [{k: B.asarray(v) for k, v in cols.items()}]

    ww = None if w is None else [B.asarray(w)]

    def f(th):
        return ON.unbinned_nll([tree], data, {n: th[i] for i, n in enumerate(names)}, weights=ww, B=B)

    oH = torch.autograd.functional.hessian(f, torch.tensor(slots, dtype=torch.float64)).numpy()
    scale = np.sqrt(np.abs(np.outer(np.diag(oH), np.diag(oH))))
    assert np.all(np.abs(H - oH) <= 1e-7 * scale + 1e-9 * np.abs(oH).max()), (eng.kernel_names, np.abs(H - oH).max())
