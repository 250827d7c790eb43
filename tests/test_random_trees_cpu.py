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
