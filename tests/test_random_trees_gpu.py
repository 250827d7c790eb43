"""Randomised model trees: value and gradient of the fused kernel (NVRTC-instantiated for shapes outside the AOT
catalogue) against the oracle engine driving the same host layer.  Guards the parts a fixed catalogue cannot: the
pre-order slot layout, the per-leaf normalisation constants, the Sum/Product epilogue for nested trees and the chain rule
through composed parameters (nested Sum/Product trees)."""

import numpy as np
import pytest

from tests._oracle_engine import OracleEngine
from tests._random_trees import build_case

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112])
def test_random_tree_value_gradient_parity(seed):
    n = 20_001  # odd: exercises the scalar tail event
    b, loss = build_case(seed, n)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    params = list(loss.get_params())
    assert len(params) == b.k
    v, g = loss.value_gradient(params, full=True)
    r, rg = ref.value_gradient(params, full=True)
    assert np.isfinite(r) and np.all(np.isfinite(rg)), (loss.model, r)
    assert abs(v - r) <= 1e-12 * max(abs(r), n), (v, r, loss._get_engine().kernel_names)
    scale = max(np.abs(rg).max(), 1.0)
    assert np.all(np.abs(g - rg) <= 1e-10 * np.maximum(np.abs(rg), scale)), (g, rg, loss._get_engine().kernel_names)


@pytest.mark.parametrize("seed", [102, 104, 106, 111, 112, 113, 115, 116])
def test_random_tree_one_pass_hessian_on_the_gpu(seed):
    """GPU twin of the CPU test of the same name: the trees the Hessian pass knows, their ``hess_kernel`` built by NVRTC,
    against torch's double backward over the oracle graph at slot level.  A tree whose accumulators exceed the reduction's
    256 (seed 104: 22 slots) has no such kernel and must say so."""
    import torch

    from oracle import nll as ON
    from oracle.backend import TorchBackend

    b, loss = build_case(seed, 3_001)
    ref = loss.create_new()
    ref._set_engine_factory(OracleEngine)
    eng, oeng = loss._get_engine(), ref._get_engine()
    slots = np.array([p.value() for p in eng.slot_params], dtype=np.float64)
    got = eng.eval_hess(slots)
    ns = len(slots)
    if 2 + (ns + 1) * (ns + 2) // 2 > 256:
        assert got is None
        loss.hessian()
        assert loss.last_hessian_method == "gradient-differences"
        return
    assert got is not None, eng.kernel_names
    H = got[2][0]
    tree, cols, w, off, n_slots = oeng.channels[0]
    B = TorchBackend()
    names = [f"s{k}" for k in range(n_slots)]
    data = [{k: B.asarray(v) for k, v in cols.items()}]
    ww = None if w is None else [B.asarray(w)]

    def f(th):
        return ON.unbinned_nll([tree], data, {n: th[i] for i, n in enumerate(names)}, weights=ww, B=B)

    oH = torch.autograd.functional.hessian(f, torch.tensor(slots, dtype=torch.float64)).numpy()
    scale = np.sqrt(np.abs(np.outer(np.diag(oH), np.diag(oH))))
    assert np.all(np.abs(H - oH) <= 1e-7 * scale + 1e-9 * np.abs(oH).max()), (eng.kernel_names, np.abs(H - oH).max())
