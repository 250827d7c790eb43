"""``zfit.settings`` / ``zfit.run`` for the hot path (``src/zfit/settings.py``, ``util/execution.py``):
fp64 everywhere; the graph / autograd switches of the reference have no meaning without TensorFlow and
are accepted as no-ops."""

from __future__ import annotations

import numpy as np


class _Options(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


# device_cut_min_events: host arrays with at least this many events that still have to be cut to their Space are
# uploaded first and cut in HBM (zfit_b200/data.py::_cut); None keeps every cut on the host
options = _Options(numerical_grad=False, auto_update_params=True, epsilon=1e-8, device_cut_min_events=1 << 22)


class _ZTypes:
    float = np.float64
    int = np.int64
    complex = np.complex128


ztypes = _ZTypes()


def set_seed(seed=None, numpy=None, backend=None):
    if seed is None:
        seed = np.random.randint(0, 2**31 - 1)
    np.random.seed(int(seed) if numpy is None or numpy is True else int(numpy))
    try:
        import torch

        torch.manual_seed(int(seed))
    except Exception:
        pass
    return {"seed": seed}


class RunManager:
    """Subset of ``zfit.run`` (``util/execution.py:23-514``)."""

    numeric_checks = False
    chunking = None

    def __init__(self):
        self._n_cpu = None

    def executing_eagerly(self):
        return True

    def set_graph_mode(self, graph=None):
        return _Null()

    def set_autograd_mode(self, autograd=None):
        return _Null()

    def set_n_cpu(self, n_cpu="auto", strict=False):
        self._n_cpu = n_cpu

    @property
    def n_cpu(self):
        import os

        return self._n_cpu if isinstance(self._n_cpu, int) else len(os.sched_getaffinity(0))

    def experimental_disable_param_update(self, value: bool = True):
        """``util/execution.py:165-181``: minimizers stop writing the fitted values into the parameters (call
        ``result.update_params()`` instead); applied at once, usable as a context manager that restores the old setting."""
        old = options["auto_update_params"]
        options["auto_update_params"] = not value

        class _Restore:
            def __enter__(self):
                return self

            def __exit__(self, *exc):
                options["auto_update_params"] = old
                return False

        return _Restore()

    def get_graph_mode(self):
        return False

    def get_autograd_mode(self):
        return True

    def assert_executing_eagerly(self):
        return None

    def clear_graph_cache(self):
        return None


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


run = RunManager()
