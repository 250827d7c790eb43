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
