"""The native MIGRAD stand-in (zfit_b200/_migrad.py) on closed-form functions: Minuit's protocol pieces one by one --
limit transformations, EDM stop rule, Dcovar-gated HESSE, forward/central HESSE switch, numerical-gradient mode and the
information-matrix seed with its fall-backs (reference driver: iminuit behind ``minimizers/minimizer_minuit.py:175-319``)."""

import numpy as np
import pytest

from zfit_b200 import _migrad as _migrad_py
from zfit_b200 import _migrad_native


@pytest.fixture(params=["native", "python"], autouse=True)
def _driver(request, monkeypatch):
    """Every test runs on the C++ driver and on its Python statement."""
    global _migrad
    _migrad = _migrad_native if request.param == "native" else _migrad_py
    return request.param


_migrad = _migrad_native


def _quadratic(A, x0):
    A = np.asarray(A, dtype=np.float64)
    x0 = np.asarray(x0, dtype=np.float64)
    calls = {"f": 0, "g": 0, "fg": 0}

    def f(x):
        calls["f"] += 1
        d = np.asarray(x) - x0
        return 0.5 * float(d @ A @ d) + 3.0

    def g(x):
        calls["g"] += 1
        return A @ (np.asarray(x) - x0)

    def fg(x):
        calls["fg"] += 1
        d = np.asarray(x) - x0
        return 0.5 * float(d @ A @ d) + 3.0, A @ d

    return f, g, fg, calls


A3 = np.array([[4.0, 1.2, -0.5], [1.2, 2.0, 0.3], [-0.5, 0.3, 1.0]])
X3 = np.array([0.7, -1.1, 2.5])


def test_quadratic_converges_and_covariance_is_inverse_hessian():
    f, g, fg, calls = _quadratic(A3, X3)
    res = _migrad.minimize(f, g, [0.0, 0.0, 0.0], [0.1] * 3, [None] * 3, [None] * 3, edm_goal=1e-6, value_gradient=fg)
    assert res.valid and res.converged
    assert np.allclose(res.x, X3, atol=2e-3)
    assert res.edm < 1e-6
    assert np.allclose(res.covariance, np.linalg.inv(A3), rtol=1e-5)  # HESSE on a quadratic is exact
    assert calls["f"] == 0, "with value_gradient no pass is spent on a value alone (seed and HESSE call the gradient)"


def test_information_seed_makes_a_quadratic_a_two_step_fit():
    f, g, fg, calls = _quadratic(A3, X3)
    plain = _migrad.minimize(f, g, [0.0] * 3, [0.1] * 3, [None] * 3, [None] * 3, edm_goal=1e-6, value_gradient=fg)
    n_plain = calls["fg"]
    f, g, fg, calls = _quadratic(A3, X3)
    seeded = _migrad.minimize(f, g, [0.0] * 3, [0.1] * 3, [None] * 3, [None] * 3, edm_goal=1e-6, value_gradient=fg,
                              information=lambda x: A3)
    assert seeded.valid and np.allclose(seeded.x, X3, atol=1e-6)  # exact metric: one Newton step lands on the minimum
    assert calls["fg"] < n_plain and seeded.niter <= 3
    assert np.allclose(seeded.covariance, plain.covariance, rtol=1e-5)


@pytest.mark.parametrize("bad", [
    lambda x: np.zeros((3, 3)),                  # singular
    lambda x: -A3,                               # not positive definite
    lambda x: np.full((3, 3), np.nan),           # non-finite
    lambda x: np.eye(2),                         # wrong shape
    lambda x: (_ for _ in ()).throw(RuntimeError("no information")),  # raises
    lambda x: None,
])
def test_unusable_information_falls_back_to_minuit_seed(bad):
    f, g, fg, _ = _quadratic(A3, X3)
    res = _migrad.minimize(f, g, [0.0] * 3, [0.1] * 3, [None] * 3, [None] * 3, edm_goal=1e-6, value_gradient=fg, information=bad)
    assert res.valid and np.allclose(res.x, X3, atol=2e-3)


def test_limits_are_respected_and_minimum_on_the_boundary_side():
    f, g, fg, _ = _quadratic(np.diag([1.0, 4.0]), [3.0, -2.0])
    # the unconstrained minimum (3, -2) lies outside the box for x0, inside for x1
    res = _migrad.minimize(f, g, [0.5, 0.0], [0.1, 0.1], [0.0, -5.0], [1.0, 5.0], edm_goal=1e-5, value_gradient=fg)
    assert 0.0 <= res.x[0] <= 1.0 and abs(res.x[0] - 1.0) < 1e-2
    assert abs(res.x[1] + 2.0) < 1e-2
    one_sided = _migrad.minimize(f, g, [5.0, 0.0], [0.1, 0.1], [4.0, None], [None, None], edm_goal=1e-5, value_gradient=fg)
    assert one_sided.x[0] >= 4.0 and abs(one_sided.x[0] - 4.0) < 1e-2


def test_rosenbrock_with_gradient_and_with_minuit_style_numerical_gradient():
    def f(x):
        return float(100.0 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2)

    def g(x):
        return np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200.0 * (x[1] - x[0] ** 2)])

    res = _migrad.minimize(f, g, [-1.2, 1.0], [0.1, 0.1], [None] * 2, [None] * 2, edm_goal=1e-7)
    assert res.valid and np.allclose(res.x, [1.0, 1.0], atol=2e-3)
    num = _migrad.minimize(f, None, [-1.2, 1.0], [0.1, 0.1], [None] * 2, [None] * 2, edm_goal=1e-7)
    assert num.valid and np.allclose(num.x, [1.0, 1.0], atol=5e-3)
    assert num.nfcn > res.nfcn  # finite differences cost value calls


def test_hesse_switches_to_central_differences_on_a_skewed_function():
    """A strongly non-quadratic function makes the forward-difference Hessian visibly asymmetric: the driver then adds the
    backward points.  Both must end at the analytic inverse Hessian."""
    ngrad = {"n": 0}

    def f(x):
        return float(np.exp(x[0]) - x[0] + 0.5 * (x[1] - 0.3 * x[0] ** 2) ** 2 + 0.1 * x[1] ** 4)

    def g(x):
        ngrad["n"] += 1
        r = x[1] - 0.3 * x[0] ** 2
        return np.array([np.exp(x[0]) - 1.0 - 0.6 * x[0] * r, r + 0.4 * x[1] ** 3])

    res = _migrad.minimize(f, g, [0.8, 0.6], [0.1, 0.1], [None] * 2, [None] * 2, edm_goal=1e-9)
    assert res.valid and np.allclose(res.x, [0.0, 0.0], atol=1e-3)
    H = np.array([[1.0, 0.0], [0.0, 1.0]])  # Hessian at the minimum (0, 0)
    assert np.allclose(res.covariance, np.linalg.inv(H), atol=2e-2)


def test_call_limit_and_nonfinite_value_are_reported_not_raised():
    f, g, fg, _ = _quadratic(A3, X3)
    res = _migrad.minimize(f, g, [50.0, -40.0, 30.0], [0.1] * 3, [None] * 3, [None] * 3, edm_goal=1e-12, maxfcn=3, value_gradient=fg)
    assert not res.valid and "limit" in res.message
    bad = _migrad.minimize(lambda x: float("nan"), lambda x: np.full(3, np.nan), [0.0] * 3, [0.1] * 3, [None] * 3, [None] * 3,
                           edm_goal=1e-6)
    assert not bad.valid and "non-finite" in bad.message


def test_native_and_python_drivers_walk_the_same_path():
    f, g, fg, c1 = _quadratic(A3, X3)
    a = _migrad_native.minimize(f, g, [0.0] * 3, [0.1] * 3, [-5, None, None], [5, None, 9], edm_goal=1e-6, value_gradient=fg)
    n1 = dict(c1)
    f, g, fg, c2 = _quadratic(A3, X3)
    b = _migrad_py.minimize(f, g, [0.0] * 3, [0.1] * 3, [-5, None, None], [5, None, 9], edm_goal=1e-6, value_gradient=fg)
    assert n1 == dict(c2) and a.niter == b.niter and a.nfcn == b.nfcn and a.ngrad == b.ngrad
    assert np.allclose(a.x, b.x, rtol=0, atol=1e-9) and abs(a.fval - b.fval) < 1e-12 and np.allclose(a.covariance, b.covariance, rtol=1e-6)
    assert a.message == b.message == "converged"


def test_exception_in_the_objective_crosses_the_native_driver():
    class Boom(RuntimeError):
        pass

    calls = {"n": 0}

    def f(x):
        calls["n"] += 1
        if calls["n"] > 3:
            raise Boom("stop")
        return float(np.sum(np.asarray(x) ** 2))

    with pytest.raises(Boom):
        _migrad_native.minimize(f, None, [1.0, 2.0], [0.1, 0.1], [None] * 2, [None] * 2, edm_goal=1e-6)
