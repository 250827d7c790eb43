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


def _cut(cols, w, lo, hi, less=None):
    import torch

    from zfit_b200 import _znll as Z

    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    src = [up(c) for c in cols]
    sw = None if w is None else up(w)
    dst = [torch.full_like(c, -777.0) for c in src]
    dw = None if sw is None else torch.full_like(sw, -777.0)
    la, lb = (0, 0) if less is None else (up(less[0]), up(less[1]))
    kept = Z.cut_events(0, [c.data_ptr() for c in src], [c.data_ptr() for c in dst], 0 if sw is None else sw.data_ptr(),
                        0 if dw is None else dw.data_ptr(), lo, hi, src[0].numel(),
                        less_a=0 if less is None else la.data_ptr(), less_b=0 if less is None else lb.data_ptr(),
                        stream=torch.cuda.current_stream(0).cuda_stream)
    return kept, [c.cpu().numpy() for c in dst], None if dw is None else dw.cpu().numpy()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65_537, 3_000_001])
@pytest.mark.parametrize("n_cols,weighted", [(1, False), (1, True), (3, True), (4, False)])
def test_cut_is_the_inclusive_stable_mask(n, n_cols, weighted):
    """``check_cut_data_weights`` (core/data.py:1203-1244): ``lower <= x <= upper`` on every observable, weights alongside;
    the kept events keep their order, the rest of the destination is not touched."""
    rng = np.random.default_rng(100 * n_cols + n)
    lo = np.array([-1.0, 0.0, 2.0, -5.0][:n_cols])
    hi = np.array([1.5, 4.0, 2.5, 5.0][:n_cols])
    cols = [rng.uniform(lo[c] - 0.2 * (hi[c] - lo[c]), hi[c] + 0.2 * (hi[c] - lo[c]), n) for c in range(n_cols)]
    if n > 40:
        cols[0][3], cols[0][5], cols[0][9] = lo[0], hi[0], np.nan  # the edges are inside (inclusive), NaN is outside
        cols[-1][17] = np.inf
    w = rng.normal(1.0, 0.5, n) if weighted else None
    mask = np.ones(n, dtype=bool)
    for c in range(n_cols):
        with np.errstate(invalid="ignore"):
            mask &= (cols[c] >= lo[c]) & (cols[c] <= hi[c])
    kept, out, ow = _cut(cols, w, lo, hi)
    assert kept == int(mask.sum())
    for c in range(n_cols):
        np.testing.assert_array_equal(out[c][:kept], cols[c][mask])
        assert np.all(out[c][kept:] == -777.0)
    if weighted:
        np.testing.assert_array_equal(ow[:kept], w[mask])


def test_cut_keeps_all_none_and_takes_the_accept_predicate():
    from zfit_b200 import _znll as Z

    rng = np.random.default_rng(8)
    x = rng.uniform(0, 1, 100_001)
    kept, out, _ = _cut([x], None, [0.0], [1.0])
    assert kept == x.size and np.array_equal(out[0], x)
    kept, out, _ = _cut([x], None, [2.0], [3.0])
    assert kept == 0 and np.all(out[0] == -777.0)
    u, p = rng.uniform(0, 1, x.size), rng.uniform(0, 1, x.size)
    kept, out, _ = _cut([x], None, [0.25], [1.0], less=(u, p))
    mask = (x >= 0.25) & (u < p)
    assert kept == int(mask.sum()) and np.array_equal(out[0][:kept], x[mask])
    assert Z.cut_events(0, [8], [16], 0, 0, [0.0], [1.0], 0) == 0  # no events: nothing is read
    with pytest.raises(Z.ZnllError):
        Z.cut_events(0, [8], [8], 0, 0, [0.0], [1.0], 10)  # aliased


def test_data_on_the_device_is_cut_by_the_library():
    """``zfit.Data`` built from CUDA tensors: the cut to the Space runs in HBM and matches the host cut event by event."""
    import torch

    import zfit_b200 as zfit

    rng = np.random.default_rng(9)
    raw = np.stack([rng.normal(0.5, 2.0, 300_007), rng.uniform(-1, 6, 300_007)], axis=1)
    w = rng.normal(1.0, 0.3, raw.shape[0])
    obs = zfit.Space("x", -2, 3) * zfit.Space("y", 0, 5)
    host = zfit.Data(raw, obs=obs, weights=w)
    dev = zfit.Data(torch.from_numpy(raw).to("cuda:0"), obs=obs, weights=torch.from_numpy(w).to("cuda:0"))
    assert dev.num_entries == host.num_entries < raw.shape[0]
    assert dev.on_device
    np.testing.assert_array_equal(np.asarray(dev.value()), np.asarray(host.value()))
    np.testing.assert_array_equal(np.asarray(dev.weights), np.asarray(host.weights))
    inside = zfit.Data(torch.from_numpy(np.asarray(host.value())).to("cuda:0"), obs=obs)
    assert inside.num_entries == host.num_entries  # nothing to cut: the caller's tensors are kept


def test_large_host_arrays_are_cut_in_hbm(monkeypatch):
    """Host arrays above ``settings.options.device_cut_min_events`` that still need their cut are uploaded and cut on the
    device (the Data is then resident); the events and their order equal the host cut's, and a loss built on it agrees."""
    import zfit_b200 as zfit

    rng = np.random.default_rng(10)
    n = 600_011
    raw = rng.normal(0.3, 1.5, n)
    w = rng.normal(1.0, 0.3, n)
    obs = zfit.Space("x", -2, 3)
    host = zfit.Data(raw, obs=obs, weights=w)
    assert not host.on_device
    monkeypatch.setitem(zfit.settings.options, "device_cut_min_events", 100_000)
    dev = zfit.Data(raw, obs=obs, weights=w)
    assert dev.on_device and dev.num_entries == host.num_entries < n
    np.testing.assert_array_equal(np.asarray(dev.value())[:, 0], np.asarray(host.value())[:, 0])
    np.testing.assert_array_equal(np.asarray(dev.weights), np.asarray(host.weights))
    mu, sg = zfit.Parameter("mu_dc", 0.2, -1, 1), zfit.Parameter("sg_dc", 1.4, 0.5, 3)
    model = zfit.pdf.Gauss(mu, sg, obs)
    a = zfit.loss.UnbinnedNLL(model, host).value_gradient([mu, sg])
    b = zfit.loss.UnbinnedNLL(model, dev).value_gradient([mu, sg])
    assert float(a[0]) == float(b[0]) and np.array_equal(np.asarray(a[1]), np.asarray(b[1]))
    monkeypatch.setitem(zfit.settings.options, "device_cut_min_events", None)
    assert not zfit.Data(raw, obs=obs).on_device


@pytest.mark.parametrize("n", [1, 31, 33, 511, 513, 8_191, 100_003])
def test_prep_kernels_stay_inside_their_destination(n):
    """Guard bands around every destination (and the sources read back unchanged): the staged move of the partition and
    the compaction of the cut write exactly their n (resp. kept) elements."""
    import torch

    from zfit_b200 import _znll as Z

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(n)
    G = 64  # guard elements on both sides
    srcs = [torch.from_numpy(rng.uniform(-1.3, 1.3, n)).to(dev) for _ in range(3)]
    keep_src = [s.clone() for s in srcs]
    bufs = [torch.full((n + 2 * G,), 12345.0, dtype=torch.float64, device=dev) for _ in range(3)]
    dsts = [b[G : G + n] for b in bufs]
    stream = torch.cuda.current_stream(0).cuda_stream
    sp, dp = [s.data_ptr() for s in srcs], [d.data_ptr() for d in dsts]
    Z.partition_events(0, sp[:2], dp[:2], sp[2], dp[2], 1, -1.0, 1.0, n, stream=stream)
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.all(b[:G] == 12345.0) and torch.all(b[G + n :] == 12345.0)
        assert not torch.any(b[G : G + n] == 12345.0)
    for b in bufs:
        b.fill_(12345.0)
    kept = Z.cut_events(0, sp[:2], dp[:2], sp[2], dp[2], [-1.0, -1.0], [1.0, 1.0], n, stream=stream)
    for b in bufs:
        assert torch.all(b[:G] == 12345.0) and torch.all(b[G + kept :] == 12345.0)
        assert not torch.any(b[G : G + kept] == 12345.0)
    for s, k in zip(srcs, keep_src):
        assert torch.equal(s, k)
