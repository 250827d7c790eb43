"""Synthetic workloads of BASELINE.json's five configs (shapes, start values: SURVEY.md section 8d).

Data are generated with seeded torch generators -- on the GPU for the full-size configs, so 1e8..1e9
events never cross PCIe -- truncated to the observable range so that no event is cut.  This is data
*generation* for benchmarks and tests (torch is plumbing here); nothing in this file is on the timed path.
"""

from __future__ import annotations

import math

from . import distributed as D

FULL_SIZES = {1: 10_000, 2: 100_000_000, 3: 50_000_000, 4: 100_000_000, 5: 125_000_000}
NAMES = {
    1: "cfg1: Gauss UnbinnedNLL, 10k events (examples/simple_fit.py)",
    2: "cfg2: SumPDF[CrystalBall,Exponential] ExtendedUnbinnedNLL, 1e8 events",
    3: "cfg3: simultaneous 2 x SumPDF[Gauss,Chebyshev4] sharing mu/sigma, 2x5e7 events",
    4: "cfg4: ProductPDF[Gauss x Exponential x Chebyshev4], 1e8 events, 3 columns",
    5: "cfg5: weighted SumPDF[Gauss,Chebyshev4], 1.25e8 events per GPU",
}


def _gen(seed, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _fill(n, device, draw, chunk=1 << 24):
    """Concatenate accepted draws until n events are collected (chunked to bound memory)."""
    import torch

    out = torch.empty(n, dtype=torch.float64, device=device)
    have = 0
    while have < n:
        x = draw(min(chunk, max(2 * (n - have), 4096)))
        k = min(x.numel(), n - have)
        out[have : have + k] = x[:k]
        have += k
    return out


def sample_gauss(n, mu, sigma, lo, hi, gen, device):
    import torch

    def draw(k):
        x = torch.randn(k, dtype=torch.float64, device=device, generator=gen) * sigma + mu
        return x[(x >= lo) & (x <= hi)]

    return _fill(n, device, draw)


def sample_exp(n, lam, lo, hi, gen, device):
    """Truncated exponential exp(lam*x) on [lo, hi] by inverse CDF."""
    import torch

    u = torch.rand(n, dtype=torch.float64, device=device, generator=gen)
    if abs(lam) < 1e-300:
        return lo + u * (hi - lo)
    span = math.expm1(lam * (hi - lo))
    x = lo + torch.log1p(u * span) / lam
    return x.clamp_(lo, hi)


def sample_cheb(n, coeffs, lo, hi, gen, device):
    import torch

    c = [1.0, *coeffs]
    bound = sum(abs(v) for v in c)

    def draw(k):
        x = lo + torch.rand(k, dtype=torch.float64, device=device, generator=gen) * (hi - lo)
        z = (2 * x - lo - hi) / (hi - lo)
        t0, t1 = torch.ones_like(z), z
        v = c[0] * t0 + (c[1] * t1 if len(c) > 1 else 0)
        for ck in c[2:]:
            t0, t1 = t1, 2 * z * t1 - t0
            v = v + ck * t1
        keep = torch.rand(k, dtype=torch.float64, device=device, generator=gen) * bound < v
        return x[keep]

    return _fill(n, device, draw)


def sample_crystalball(n, mu, sigma, alpha, nn, lo, hi, gen, device):
    """CrystalBall (alpha > 0) by composition: with probability p_tail the power-law tail (inverse CDF, t < -alpha), else
    the Gaussian core drawn from the normal TRUNCATED to t >= -alpha (inverse CDF through ndtri) -- every draw of a
    component lands in that component, so the mixture weights are the pdf's own; draws outside [lo, hi] are then
    rejected as a whole, which is sampling from the pdf truncated to the observable.  (Round 1 rejected core draws below
    -alpha without redrawing them, which inflated the tail share from 12.2 % to 12.9 % and biased every closure fit.)"""
    import torch

    a = abs(alpha)
    tail_int = (nn / a) * math.exp(-0.5 * a * a) / (nn - 1.0)
    phi_lo = 0.5 * (1.0 + math.erf(-a / math.sqrt(2)))  # Phi(-a)
    core_int = math.sqrt(2 * math.pi) * (1.0 - phi_lo)
    p_tail = tail_int / (tail_int + core_int)
    B = nn / a - a

    def draw(k):
        u = torch.rand(k, dtype=torch.float64, device=device, generator=gen)
        is_tail = u < p_tail
        v = torch.rand(k, dtype=torch.float64, device=device, generator=gen).clamp_min_(1e-300)
        t_tail = B - (nn / a) * v.pow(-1.0 / (nn - 1.0))
        uc = torch.rand(k, dtype=torch.float64, device=device, generator=gen)
        t_core = torch.special.ndtri((phi_lo + uc * (1.0 - phi_lo)).clamp_(1e-300, 1.0 - 1e-16))
        t = torch.where(is_tail, t_tail, t_core)
        x = mu + sigma * t
        return x[(x >= lo) & (x <= hi)]

    return _fill(n, device, draw)


def _shuffle(x, gen, device):
    import torch

    return x[torch.randperm(x.numel(), device=device, generator=gen)]


def build(cfg: int, n: int | None = None, device="cpu", rank: int = 0, zfit=None):
    """-> dict(loss=..., params=[...], n_events=total events evaluated per call on this rank, ...).

    ``n``: events per channel on this rank (default: the full size of the config).
    """
    import torch

    if zfit is None:
        import zfit_b200 as zfit
    n = FULL_SIZES[cfg] if n is None else int(n)
    dev = torch.device(device)
    tag = f"_c{cfg}"  # parameter names are unique per config so that several can coexist in one process

    if cfg == 1:
        gen = _gen(1, dev)
        obs = zfit.Space("x", -2, 3)
        raw = torch.randn(n, dtype=torch.float64, device=dev, generator=gen) * 3.0 + 2.0
        mu = zfit.Parameter("mu" + tag, 1.2, -4, 6)
        sigma = zfit.Parameter("sigma" + tag, 1.3, 0.5, 10)
        model = zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs)
        data = zfit.Data(raw, obs=obs)  # cut to the space, ~5.4k survive (examples/simple_fit.py)
        loss = zfit.loss.UnbinnedNLL(model=model, data=data)
        return {"loss": loss, "n_events": data.num_entries, "bytes_per_event": 8, "model": [model], "data": [data]}

    if cfg == 2:
        gen = _gen(2 + 1000 * rank, dev)
        lo, hi = 5000.0, 5600.0
        obs = zfit.Space("m", lo, hi)
        nsig = int(round(0.3 * n))
        sig = sample_crystalball(nsig, 5280.0, 20.0, 1.5, 3.0, lo, hi, gen, dev)
        bkg = sample_exp(n - nsig, -0.002, lo, hi, gen, dev)
        x = _shuffle(torch.cat([sig, bkg]), gen, dev)
        del sig, bkg
        mu = zfit.Parameter("mu" + tag, 5275.0, 5200, 5350)
        sigma = zfit.Parameter("sigma" + tag, 25.0, 5, 60)
        alpha = zfit.Parameter("alpha" + tag, 1.2, 0.1, 5)
        nn = zfit.Parameter("n" + tag, 2.5, 1.05, 20)
        lam = zfit.Parameter("lambda" + tag, -0.003, -0.02, -1e-5)
        # the yields count the events of ALL ranks (each rank holds its own n-event shard of one sample)
        ntot = float(n) * D.world_size()
        n_sig = zfit.Parameter("n_sig" + tag, 0.25 * ntot, 0, 2.0 * ntot, stepsize=max(1.0, math.sqrt(ntot)))
        n_bkg = zfit.Parameter("n_bkg" + tag, 0.75 * ntot, 0, 2.0 * ntot, stepsize=max(1.0, math.sqrt(ntot)))
        cb = zfit.pdf.CrystalBall(mu=mu, sigma=sigma, alpha=alpha, n=nn, obs=obs, extended=n_sig)
        ex = zfit.pdf.Exponential(lam, obs=obs, extended=n_bkg)
        model = zfit.pdf.SumPDF([cb, ex])
        data = zfit.Data(x, obs=obs, guarantee_limits=True)
        loss = zfit.loss.ExtendedUnbinnedNLL(model=model, data=data)
        return {"loss": loss, "n_events": n, "bytes_per_event": 8, "model": [model], "data": [data]}

    if cfg == 3:
        lo, hi = -10.0, 10.0
        mu = zfit.Parameter("mu" + tag, 0.8, -5, 5)
        sigma = zfit.Parameter("sigma" + tag, 1.7, 0.3, 6)
        models, datas = [], []
        truth = [(0.40, (0.2, -0.1, 0.05, 0.02), 3), (0.25, (-0.15, 0.1, 0.0, 0.03), 4)]
        for ch, (frac, coeffs, seed) in enumerate(truth):
            gen = _gen(seed + 1000 * rank, dev)
            obs = zfit.Space(f"x{ch}", lo, hi)
            ns = int(round(frac * n))
            sig = sample_gauss(ns, 1.0, 1.5, lo, hi, gen, dev)
            bkg = sample_cheb(n - ns, coeffs, lo, hi, gen, dev)
            x = _shuffle(torch.cat([sig, bkg]), gen, dev)
            del sig, bkg
            fr = zfit.Parameter(f"frac{ch}" + tag, frac + 0.05, 0, 1)
            cs = [zfit.Parameter(f"c{k + 1}_{ch}" + tag, 0.7 * c + 0.01, -1, 1) for k, c in enumerate(coeffs)]
            g = zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs)
            cheb = zfit.pdf.Chebyshev(obs=obs, coeffs=cs)
            models.append(zfit.pdf.SumPDF([g, cheb], fracs=fr))
            datas.append(zfit.Data(x, obs=obs, guarantee_limits=True))
        loss = zfit.loss.UnbinnedNLL(model=models, data=datas)
        return {"loss": loss, "n_events": 2 * n, "bytes_per_event": 8, "model": models, "data": datas}

    if cfg == 4:
        gen = _gen(5 + 1000 * rank, dev)
        ox, oy, oz = zfit.Space("x", -4, 4), zfit.Space("y", 0, 5), zfit.Space("z", -2, 4)
        x = sample_gauss(n, 0.3, 1.2, -4.0, 4.0, gen, dev)
        y = sample_exp(n, -0.7, 0.0, 5.0, gen, dev)
        zc = (0.2, -0.1, 0.05, 0.02)
        z = sample_cheb(n, zc, -2.0, 4.0, gen, dev)
        mu = zfit.Parameter("mu" + tag, 0.2, -2, 2)
        sigma = zfit.Parameter("sigma" + tag, 1.3, 0.3, 4)
        lam = zfit.Parameter("lambda" + tag, -0.6, -3, -0.01)
        cs = [zfit.Parameter(f"c{k + 1}" + tag, 0.7 * c + 0.01, -1, 1) for k, c in enumerate(zc)]
        model = zfit.pdf.ProductPDF(
            [zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=ox), zfit.pdf.Exponential(lam, obs=oy), zfit.pdf.Chebyshev(obs=oz, coeffs=cs)]
        )
        data = zfit.Data(torch.stack([x, y, z], dim=1), obs=ox * oy * oz, guarantee_limits=True)
        del x, y, z
        loss = zfit.loss.UnbinnedNLL(model=model, data=data)
        return {"loss": loss, "n_events": n, "bytes_per_event": 24, "model": [model], "data": [data]}

    if cfg == 5:
        gen = _gen(6 + rank, dev)
        lo, hi = -10.0, 10.0
        obs = zfit.Space("x", lo, hi)
        coeffs = (0.2, -0.1, 0.05, 0.02)
        ns = int(round(0.4 * n))
        sig = sample_gauss(ns, 1.0, 1.5, lo, hi, gen, dev)
        bkg = sample_cheb(n - ns, coeffs, lo, hi, gen, dev)
        x = _shuffle(torch.cat([sig, bkg]), gen, dev)
        del sig, bkg
        w = torch.randn(n, dtype=torch.float64, device=dev, generator=gen) * 0.6 + 1.0  # sWeight-like, ~5% negative
        mu = zfit.Parameter("mu" + tag, 0.8, -5, 5)
        sigma = zfit.Parameter("sigma" + tag, 1.7, 0.3, 6)
        fr = zfit.Parameter("frac" + tag, 0.45, 0, 1)
        cs = [zfit.Parameter(f"c{k + 1}" + tag, 0.7 * c + 0.01, -1, 1) for k, c in enumerate(coeffs)]
        model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu=mu, sigma=sigma, obs=obs), zfit.pdf.Chebyshev(obs=obs, coeffs=cs)], fracs=fr)
        data = zfit.Data(x, obs=obs, weights=w, guarantee_limits=True)
        loss = zfit.loss.UnbinnedNLL(model=model, data=data)
        return {"loss": loss, "n_events": n, "bytes_per_event": 16, "model": [model], "data": [data]}

    msg = f"unknown config {cfg}"
    raise ValueError(msg)


# Algorithmic work per event (frozen counting convention of SURVEY.md section 8d / BASELINE.md section 3):
# add/sub/mul = 1, FMA = 2, compare/select = 0, EXP = 30, LOG = 45, RCP/DIV = 14, POW = 76.
F_ALG = {1: 45.0, 2: 200.0, 3: 145.0, 4: 100.0, 5: 155.0}
B_ALG = {1: 8.0, 2: 8.0, 3: 8.0, 4: 24.0, 5: 16.0}
