import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from zfit_b200 import _znll

        return _znll.lib().znll_device_count() > 0
    except Exception:
        return False


def _gpu_box():
    """Is there a GPU even though the library did not load?  Device nodes, else the driver's own tool, else torch."""
    import glob
    import shutil
    import subprocess

    if glob.glob("/dev/nvidia[0-9]*") or os.path.exists("/dev/nvidiactl"):
        return True
    smi = shutil.which("nvidia-smi")
    if smi:
        try:
            out = subprocess.run([smi, "-L"], capture_output=True, text=True, timeout=20)
            if out.returncode == 0 and "GPU " in out.stdout:
                return True
        except Exception:
            pass
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


@pytest.fixture(scope="session")
def have_gpu():
    return _have_gpu()


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a CPU box: they are skipped (visibly) without a device,
    # and they FAIL (not skip) on a GPU box if the native library cannot be loaded.
    if _have_gpu() or _gpu_box():  # a GPU box whose library does not load must show failures
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
