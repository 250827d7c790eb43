"""Accuracy of the kernel's hand-rolled fp64 exp / log / reciprocal against libm (numpy)."""

import numpy as np
import pytest

from zfit_b200 import _znll as Z

pytestmark = pytest.mark.gpu


def _ulp_err(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_fast_exp():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-699, 699, 400_000), rng.uniform(-40, 5, 400_000), rng.normal(0, 1e-3, 100_000),
                        [0.0, -0.0, 1.0, -1.0, 699.999, -699.999]])
    got = Z.test_math(0, x)
    ref = np.exp(x)
    assert _ulp_err(got, ref).max() <= 2.0
    # outside the fast range the libdevice path answers: overflow, underflow, denormals, NaN
    xs = np.array([710.0, -746.0, -720.0, 800.0, -800.0, np.inf, -np.inf, np.nan, 700.0, -700.0])
    got = Z.test_math(0, xs)
    with np.errstate(all="ignore"):
        ref = np.exp(xs)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=4e-16)


def test_fast_log():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 400_000)), rng.uniform(0.5, 2.0, 400_000), rng.uniform(1e-6, 1.0, 200_000),
                        [1.0, 2.0, 0.5, 1e-307, 1e300, np.nextafter(1.0, 0), np.nextafter(1.0, 2)]])
    got = Z.test_math(1, x)
    ref = np.log(x)
    # absolute error matters for a sum of logs; relative error except in the immediate vicinity of log(1) = 0
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-2)) <= 4e-16
    assert np.max(np.abs(got - ref)) <= 3e-13  # |log| up to 700
    xs = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310])
    got = Z.test_math(1, xs)
    with np.errstate(all="ignore"):
        ref = np.log(xs)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(got[ok][np.isinf(ref[ok])], ref[ok][np.isinf(ref[ok])])
    fin = ok & np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=4e-16)


def test_fast_rcp():
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 500_000)) * rng.choice([-1.0, 1.0], 500_000), rng.uniform(1e-3, 10, 300_000)])
    got = Z.test_math(2, x)
    assert _ulp_err(got, 1.0 / x).max() <= 1.0
    xs = np.array([0.0, np.inf, -np.inf, np.nan, 1e-310])
    got = Z.test_math(2, xs)
    with np.errstate(all="ignore"):
        ref = 1.0 / xs
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=4e-16)
