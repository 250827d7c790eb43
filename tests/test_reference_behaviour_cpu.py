"""CPU twin of ``test_reference_behaviour_gpu.py``: the reference's own loss tests (``tests/test_loss.py`` etc., re-expressed
there) run unchanged, with ``CudaEngine`` replaced -- in this test process only -- by ``HostSimEngine``: model lowering,
prologue, epilogue and the whole Python host layer are the product's, the kernel's event loop is the host-compiled device
header.  The two tests that need ``pdf_kernel`` launches through a plan stay GPU-only."""

import pytest

import importlib

loss_mod = importlib.import_module("zfit_b200.loss")  # `zfit_b200.loss` the attribute is the zfit-style namespace
from tests import test_reference_behaviour_gpu as G
from tests._hostsim import HostSimEngine


@pytest.fixture(autouse=True)
def _device_code_on_the_host(monkeypatch):
    monkeypatch.setattr(loss_mod, "CudaEngine", lambda models, datas, fit_ranges, **_: HostSimEngine(models, datas, fit_ranges))


_GPU_ONLY = {"test_exponential_numerics_far_from_zero", "test_product_pdf_values_are_product_of_factors"}
for _name in dir(G):
    if _name.startswith("test_") and _name not in _GPU_ONLY:
        globals()[_name] = getattr(G, _name)
del _name
