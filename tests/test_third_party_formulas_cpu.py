"""The arithmetic the reference delegates to TensorFlow / tensorflow-probability (``tfp.distributions.Normal.prob/cdf``,
``tf.nn.log_poisson_loss``, ``tf.math.erf``, ``tfp.math.reduce_kahan_sum``) cannot be executed here (neither package is
installable), and the reference stores no values for it ("parity unpinned" at this boundary).  The
oracle restates the published formulas; this file holds those restatements -- and the product's host prologue where it
computes the same quantity -- to the independent implementations that ARE in this image (scipy, PyTorch, mpmath, the
exact Poisson pmf), so that a transcription slip cannot hide behind "both sides agree".  It narrows the gap; it does not
close it: a TensorFlow-specific rounding difference at the 1e-16 level would still be invisible."""

import math

import numpy as np
import pytest
import scipy.special
import scipy.stats

from oracle import nll as ON
from oracle import pdfs as O
from oracle.backend import NumpyBackend, TorchBackend
from zfit_b200 import _znll as Z

torch = pytest.importorskip("torch")
B = NumpyBackend()
rng = np.random.default_rng(77)


def test_normal_prob_against_scipy_and_torch():
    g = O.Gauss("mu", "sigma", "x", (-10, 10))
    for mu, sigma in [(0.3, 1.2), (5279.9, 20.04), (-3.0, 0.05), (1e-3, 7.5)]:
        x = mu + sigma * rng.normal(size=2000) * rng.choice([0.1, 1.0, 6.0], size=2000)
        mine = g.unnorm(B, x, {"mu": mu, "sigma": sigma}, False)
        # tfp forms the argument as x/sigma - mu/sigma, scipy and torch as (x - mu)/sigma: the two
        # roundings differ by eps * |mu/sigma| in d, i.e. by eps * |mu/sigma| * |d| relative in exp(-d^2/2)
        d = np.abs(x - mu) / sigma
        tol = 4e-15 + 4 * np.finfo(float).eps * (abs(mu / sigma) + d) * np.maximum(d, 1.0)
        ref = scipy.stats.norm.pdf(x, mu, sigma)
        assert np.all(np.abs(mine - ref) <= tol * ref), float(np.max(np.abs(mine - ref) / (tol * ref)))
        tl = torch.distributions.Normal(torch.tensor(mu, dtype=torch.float64), torch.tensor(sigma, dtype=torch.float64))
        ref = tl.log_prob(torch.from_numpy(x)).exp().numpy()
        assert np.all(np.abs(mine - ref) <= tol * ref), float(np.max(np.abs(mine - ref) / (tol * ref)))


def test_ndtr_and_gauss_integral_against_scipy_torch_mpmath_and_the_host_prologue():
    mp = pytest.importorskip("mpmath")
    x = np.concatenate([rng.uniform(-8, 8, 3000), [-0.70710678, 0.70710678, 0.0, -37.0, 8.2]])
    mine = O.Gauss.ndtr(B, x)
    np.testing.assert_allclose(mine, scipy.special.ndtr(x), rtol=4e-15, atol=0)
    # torch's ndtr is 0.5 * (1 + erf(x / sqrt 2)): fine near the centre, absolute accuracy only in the left tail (which is
    # why tfp, scipy and this restatement switch to erfc there)
    np.testing.assert_allclose(mine, torch.special.ndtr(torch.from_numpy(x)).numpy(), rtol=0, atol=2.3e-16)
    mp.mp.dps = 40
    exact = np.array([float(mp.ncdf(mp.mpf(float(v)))) for v in x[:400]])
    # in the far tail the rounding of x / sqrt 2 is amplified by x^2 (d ln erfc(z) / d ln z ~ -2 z^2)
    assert np.all(np.abs(mine[:400] - exact) <= 4e-16 * (2.0 + x[:400] ** 2) * exact)
    g = O.Gauss("mu", "sigma", "x", (-2, 3))
    for mu, sigma, lo, hi in [(1.2, 1.3, -2, 3), (5280.0, 20.0, 5000, 5600), (0.0, 1.0, 3.0, 9.0), (0.0, 1.0, -9.0, -3.0)]:
        ref = float(mp.ncdf((mp.mpf(hi) - mu) / sigma) - mp.ncdf((mp.mpf(lo) - mu) / sigma))
        mine_I = float(g.integral(B, {"mu": mu, "sigma": sigma}, lo, hi, (lo, hi)))
        host_I, _ = Z.leaf_integral(Z.Node(Z.GAUSS, 0, 0, 0, lo, hi, 0, 0, 0), [mu, sigma])
        # a difference of two cdfs: the absolute error is a rounding of numbers of order one (the reference's formula,
        # dist_tfp.py:113-120, has the same cancellation in the far right tail)
        assert abs(mine_I - ref) <= 3e-16 and abs(host_I - ref) <= 3e-16, (mu, sigma, lo, hi, mine_I, host_I, ref)
        assert abs(host_I - mine_I) <= 1e-15 * abs(ref) + 2.3e-16


def test_erf_against_mpmath_and_torch():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    x = np.concatenate([rng.uniform(-6, 6, 500), [0.0, 1e-9, -1e-9, 0.5, 3.2]])
    mine, minec = B.erf(x), B.erfc(x)
    np.testing.assert_allclose(mine, [float(mp.erf(mp.mpf(float(v)))) for v in x], rtol=5e-16, atol=1e-300)
    np.testing.assert_allclose(minec, [float(mp.erfc(mp.mpf(float(v)))) for v in x], rtol=1e-14, atol=0)
    np.testing.assert_allclose(mine, torch.erf(torch.from_numpy(x)).numpy(), rtol=1e-15, atol=1e-300)


def test_log_poisson_loss_against_torch_and_the_exact_pmf():
    F = torch.nn.functional
    for t in [0.0, 0.5, 1.0, 2.0, 17.0, 21200.0, 1.0e8, 123456.789]:  # a weighted sample size need not be an integer
        for y in [max(t, 1.0) * f for f in (0.5, 0.999, 1.0, 1.7)]:
            log_y = math.log(y)
            for full in (False, True):
                mine = float(ON.log_poisson_loss(B, t, np.float64(log_y), full))
                ref = float(F.poisson_nll_loss(torch.tensor(log_y, dtype=torch.float64), torch.tensor(t, dtype=torch.float64),
                                               log_input=True, full=full, reduction="none"))
                assert mine == pytest.approx(ref, rel=1e-15, abs=1e-12 * max(1.0, t)), (t, y, full)
            if float(t).is_integer() and t > 1:  # full loss = -log pmf up to Stirling's 1/(12 t) remainder
                exact = -scipy.stats.poisson.logpmf(int(t), y)
                assert float(ON.log_poisson_loss(B, t, np.float64(log_y), True)) == pytest.approx(exact, abs=1.0 / (12.0 * t) + 1e-9 * t)


def test_kahan_sum_against_fsum():
    terms = rng.normal(size=200_001) * 10.0 ** rng.integers(-6, 7, size=200_001)
    exact = math.fsum(terms)
    s = c = 0.0
    for v in terms[:20001]:  # the compensated recurrence of tfp.math.reduce_kahan_sum, spelt out
        y = v - c
        t = s + y
        c = (t - s) - y
        s = t
    assert abs(s - c - math.fsum(terms[:20001])) <= 4 * np.finfo(float).eps * abs(math.fsum(terms[:20001])) + 1e-9
    # the oracle's value path at fit scale: plain fp64 summation stays within n*eps*sum|terms| of the exact sum, the
    # bound the GPU tests hold the kernel's compensated reduction to (tests/test_cabi_gpu.py)
    assert abs(float(np.sum(terms)) - exact) <= len(terms) * np.finfo(float).eps * float(np.sum(np.abs(terms)))


def test_torch_backend_agrees_with_numpy_backend_on_the_same_formulas():
    T = TorchBackend()
    x = rng.uniform(-5, 5, 1000)
    np.testing.assert_allclose(O.Gauss.ndtr(T, torch.from_numpy(x)).numpy(), O.Gauss.ndtr(B, x), rtol=1e-13, atol=1e-300)
