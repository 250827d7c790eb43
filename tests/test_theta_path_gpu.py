"""The C-level theta path (``znll_eval_theta``: slots from theta, fused pass, chain rule, extended term in one native call)
and the whole-fit native loop (``znll_migrad_theta``) against the Python chain rule / the evaluator route."""

import numpy as np
import pytest

import zfit_b200 as zfit
from zfit_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _pair(cfg, n):
    w = W.build(cfg, n=n, device="cuda:0")
    fast = w["loss"]
    slow = type(fast)(model=w["model"], data=w["data"], options={"theta_path": False})
    return w, fast, slow


@pytest.mark.parametrize("cfg,n", [(1, 10_000), (2, 100_001), (3, 60_000), (4, 100_000), (5, 100_000)])
def test_theta_path_equals_python_chain_rule(cfg, n):
    w, fast, slow = _pair(cfg, n)
    params = list(fast.get_params())
    for full in (True, False):
        v1, g1 = fast.value_gradient(params, full=full)
        slow.options["subtr_const_value"] = fast.options["subtr_const_value"]
        v2, g2 = slow.value_gradient(params, full=full)
        assert fast._theta_path(fast._get_engine()) is not None and slow._theta_path(slow._get_engine()) is None
        assert abs(v1 - v2) <= 1e-13 * max(abs(v2), 1.0)
        np.testing.assert_allclose(g1, g2, rtol=1e-12, atol=1e-12 * np.abs(g2).max())
        assert fast.value(full=full) == v1
    # a subset in another order, with one parameter fixed at a new value
    params[0].floating = False
    params[0].set_value(params[0].value() * 1.001)
    sub = [params[-1], params[1]]
    g1, g2 = fast.gradient(sub), slow.gradient(sub)
    np.testing.assert_allclose(g1, g2, rtol=1e-12, atol=1e-12 * np.abs(g2).max())
    params[0].floating = True


def test_user_composed_parameter_keeps_the_python_chain_rule():
    rng = np.random.default_rng(1)
    obs = zfit.Space("x", -5, 5)
    a = zfit.Parameter("a_tp", 0.2, -1, 1)
    mu = zfit.ComposedParameter("mu_tp", lambda p: 2.0 * p["a"] + 0.1, params={"a": a})
    sg = zfit.Parameter("sg_tp", 1.1, 0.3, 3)
    loss = zfit.loss.UnbinnedNLL(zfit.pdf.Gauss(mu, sg, obs), zfit.Data(rng.normal(0.5, 1.0, 5000).clip(-5, 5), obs=obs))
    v, g = loss.value_gradient([a, sg], full=True)
    assert loss._theta_path(loss._get_engine()) is None
    assert np.all(np.isfinite(g)) and np.isfinite(v)


@pytest.mark.parametrize("cfg,n", [(1, 10_000), (2, 200_000), (3, 100_000), (5, 200_000)])
def test_native_loop_fit_equals_evaluator_route(cfg, n):
    w = W.build(cfg, n=n, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    start = [p.value() for p in params]
    r1 = zfit.minimize.Minuit(gradient="zfit", tol=1e-9).minimize(loss)
    assert "whole fit in native code" in r1.info["driver"], r1.info["driver"]
    x1 = np.array([r1.params[p]["value"] for p in params])
    zfit.param.set_values(params, start)
    r2 = zfit.minimize.Minuit(gradient="zfit", tol=1e-9, options={"native_loop": False}).minimize(loss)
    assert "whole fit" not in r2.info["driver"]
    x2 = np.array([r2.params[p]["value"] for p in params])
    assert r1.valid and r2.valid
    assert np.all(np.abs(x1 - x2) <= 1e-8 * np.maximum(1.0, np.abs(x2))), (x1, x2)
    assert abs(r1.fmin - r2.fmin) <= 1e-9 * max(1.0, abs(r2.fmin))
    assert abs(r1.info["n_pass"] - r2.info["n_pass"]) <= 2
    errs = r1.hesse()
    assert all(np.isfinite(errs[p]["error"]) and errs[p]["error"] > 0 for p in params)


def test_native_loop_hands_over_on_a_non_finite_start():
    """A start value that makes the loss NaN: the native loop returns, the evaluator route with its NaN strategy runs."""
    w = W.build(5, n=50_000, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    loss.value_gradient(params, full=False)  # offset fixed at a sane point
    c1 = [p for p in params if p.name.startswith("c1")][0]
    c1.set_value(-0.999)
    [p for p in params if p.name.startswith("c2")][0].set_value(-0.999)  # a polynomial that goes negative: NaN events
    res = zfit.minimize.Minuit(gradient="zfit").minimize(loss)
    assert "whole fit" not in res.info.get("driver", "")


def test_theta_path_extended_weighted_two_channels_with_constraint():
    """Every term the native call owns at once: two channels, weights, yields -> fractions, the Poisson terms with sum(w) as
    the observed count, a constant slot, a shared parameter -- plus a Gaussian constraint added on the Python side."""
    rng = np.random.default_rng(8)
    mu = zfit.Parameter("mu_tx", 0.4, -2, 2)
    sg = zfit.Parameter("sg_tx", 1.2, 0.3, 4)
    models, datas = [], []
    for ch in range(2):
        obs = zfit.Space(f"x{ch}", -6, 6)
        ns = zfit.Parameter(f"ns_tx{ch}", 3000.0 + 500 * ch, 0, 1e5)
        nb = zfit.Parameter(f"nb_tx{ch}", 6000.0, 0, 1e5)
        lam = zfit.Parameter(f"lam_tx{ch}", -0.1 - 0.05 * ch, -2, 2)
        g = zfit.pdf.Gauss(mu, sg if ch == 0 else 1.5, obs, extended=ns)  # channel 1: a constant width slot
        e = zfit.pdf.Exponential(lam, obs, extended=nb)
        models.append(zfit.pdf.SumPDF([g, e]))
        x = np.concatenate([rng.normal(0.5, 1.3, 3000), rng.uniform(-6, 6, 6000)]).clip(-6, 6)
        datas.append(zfit.Data(x, obs=obs, weights=rng.normal(1.0, 0.4, x.size)))
    con = zfit.constraint.GaussianConstraint(mu, 0.45, sigma=0.2)
    fast = zfit.loss.ExtendedUnbinnedNLL(models, datas, constraints=con)
    slow = zfit.loss.ExtendedUnbinnedNLL(models, datas, constraints=con, options={"theta_path": False})
    params = list(fast.get_params())
    for full in (True, False):
        v1, g1 = fast.value_gradient(params, full=full)
        slow.options["subtr_const_value"] = fast.options["subtr_const_value"]
        v2, g2 = slow.value_gradient(params, full=full)
        assert fast._theta_path(fast._get_engine()) is not None and slow._theta_path(slow._get_engine()) is None
        assert abs(v1 - v2) <= 1e-13 * max(abs(v2), 1.0)
        np.testing.assert_allclose(g1, g2, rtol=1e-12, atol=1e-12 * np.abs(g2).max())
    # constraints keep the evaluator route for fits (they live in Python); the result must not depend on the route
    r = zfit.minimize.Minuit(gradient="zfit").minimize(fast)
    assert r.valid and "whole fit" not in r.info["driver"]
