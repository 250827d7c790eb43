"""Bind-time bucket partition (csrc/znll_prep.cu): stable, deterministic, carries every stream, and the
NLL of a loss does not depend on it beyond the summation order."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _partition(cols, w, key, lo, hi):
    import torch

    from zfit_b200 import _znll as Z

    dev = torch.device("cuda", 0)
    src = [torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in cols]
    sw = None if w is None else torch.from_numpy(np.ascontiguousarray(w)).to(dev)
    dst = [torch.empty_like(c) for c in src]
    dw = None if sw is None else torch.empty_like(sw)
    Z.partition_events(0, [c.data_ptr() for c in src], [c.data_ptr() for c in dst], 0 if sw is None else sw.data_ptr(),
                       0 if dw is None else dw.data_ptr(), key, lo, hi, src[0].numel(),
                       stream=torch.cuda.current_stream(0).cuda_stream)
    torch.cuda.synchronize()
    return [c.cpu().numpy() for c in dst], None if dw is None else dw.cpu().numpy()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65_537, 3_000_001])
@pytest.mark.parametrize("weighted", [False, True])
def test_partition_is_the_stable_bucket_order(n, weighted):
    rng = np.random.default_rng(n)
    lo, hi = -3.0, 5.0
    x = rng.uniform(lo - 0.5, hi + 0.5, n)  # some events outside the range: clamped into the end buckets
    if n > 40:
        x[7], x[11] = np.nan, hi  # NaN goes to bucket 0, the upper edge to the last bucket
    y = rng.normal(size=n)
    w = rng.normal(1.0, 0.5, n) if weighted else None
    (py, px), pw = _partition([y, x], w, 1, lo, hi)
    with np.errstate(invalid="ignore"):
        b = np.floor((x - lo) * (32.0 / (hi - lo)))
    b = np.where(np.isnan(b), 0, np.clip(b, 0, 31)).astype(np.int64)
    order = np.argsort(b, kind="stable")
    np.testing.assert_array_equal(px, x[order])
    np.testing.assert_array_equal(py, y[order])
    if weighted:
        np.testing.assert_array_equal(pw, w[order])


def test_partition_is_deterministic_and_rejects_bad_arguments():
    from zfit_b200 import _znll as Z

    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, 200_003)
    a, _ = _partition([x], None, 0, 0.0, 1.0)
    b, _ = _partition([x], None, 0, 0.0, 1.0)
    np.testing.assert_array_equal(a[0], b[0])
    with pytest.raises(Z.ZnllError):
        _partition([x], None, 0, 1.0, 1.0)


def test_loss_value_does_not_depend_on_the_partition():
    import zfit_b200 as zfit
    from zfit_b200 import workloads as W

    w = W.build(2, n=400_001, device="cuda:0")
    params = list(w["loss"].get_params())
    out = {}
    for sort in (True, False):
        loss = type(w["loss"])(model=w["model"], data=w["data"], options={"sort_events": sort, "subtr_const": False})
        out[sort] = loss.value_gradient(params, full=True)
    v1, g1 = out[True]
    v0, g0 = out[False]
    assert abs(v1 - v0) <= 1e-12 * abs(v0)
    np.testing.assert_allclose(g1, g0, rtol=1e-10, atol=1e-10 * np.abs(g0).max())


def test_streamed_upload_partitions_chunk_by_chunk(monkeypatch):
    """Pinned host columns take the pipelined path (upload chunk k+1 while chunk k is bucketed): same NLL as the
    device-born path, every chunk of the resident column is a bucket-ordered permutation of the chunk that was sent."""
    import torch

    import zfit_b200 as zfit
    from zfit_b200 import workloads as W
    from zfit_b200.loss import CudaEngine

    monkeypatch.setattr(CudaEngine, "STREAM_CHUNK", 1 << 16)
    n = 5 * (1 << 16) + 12_345
    w = W.build(2, n=n, device="cuda:0")
    params = list(w["loss"].get_params())
    ref_v, ref_g = type(w["loss"])(model=w["model"], data=w["data"], options={"subtr_const": False}).value_gradient(params, full=True)
    dev_x = w["data"][0]._cols[0]
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    host.copy_(dev_x)
    data = zfit.Data({"m": host}, obs=w["data"][0].space, guarantee_limits=True)
    loss = type(w["loss"])(model=w["model"], data=data, options={"subtr_const": False})
    v, g = loss.value_gradient(params, full=True)
    assert data._dev_cache is None  # streamed: the unpartitioned copy never existed on the device
    assert abs(v - ref_v) <= 1e-12 * abs(ref_v)
    np.testing.assert_allclose(g, ref_g, rtol=1e-10, atol=1e-10 * np.abs(ref_g).max())
    resident = loss._get_engine()._keep[0][0][0].cpu().numpy()
    sent = host.numpy()
    lo, hi = 5000.0, 5600.0
    for start in range(0, n, 1 << 16):
        a, b = sent[start : start + (1 << 16)], resident[start : start + (1 << 16)]
        bucket = np.clip(np.floor((a - lo) * (32.0 / (hi - lo))), 0, 31).astype(np.int64)
        np.testing.assert_array_equal(b, a[np.argsort(bucket, kind="stable")])
    # pageable host memory keeps the two-step path and gives the same numbers
    data2 = zfit.Data({"m": sent.copy()}, obs=w["data"][0].space, guarantee_limits=True)
    v2, _ = type(w["loss"])(model=w["model"], data=data2, options={"subtr_const": False}).value_gradient(params, full=True)
    assert abs(v2 - ref_v) <= 1e-12 * abs(ref_v)
