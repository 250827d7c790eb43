"""Toy study throughput: resample on the GPU + reset + Minuit fit on ONE loss (SamplerData.resample pattern)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zfit_b200 as zfit

obs = zfit.Space("m", 5000, 5600)
mu, sigma = zfit.Parameter("mu_t", 5280, 5200, 5350), zfit.Parameter("sigma_t", 20, 5, 60)
lam = zfit.Parameter("lam_t", -0.002, -0.02, -1e-5)
for n in (10_000, 1_000_000, 10_000_000):
    nsig, nbkg = zfit.Parameter(f"nsig_t{n}", 0.3 * n, 0, 2 * n), zfit.Parameter(f"nbkg_t{n}", 0.7 * n, 0, 2 * n)
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs, extended=nsig), zfit.pdf.Exponential(lam, obs, extended=nbkg)])
    params = [nsig, nbkg, mu, sigma, lam]
    start = [0.25 * n, 0.75 * n, 5275.0, 25.0, -0.003]
    sampler = model.create_sampler()
    nll = zfit.loss.ExtendedUnbinnedNLL(model, sampler)
    minimizer = zfit.minimize.Minuit(gradient="zfit")
    ntoys = 40 if n <= 1_000_000 else 15
    t_s = t_f = 0.0
    ok = 0
    for toy in range(ntoys + 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sampler.resample()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        zfit.param.set_values(params, start)
        res = minimizer.minimize(nll)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if toy >= 2:
            t_s += t1 - t0; t_f += t2 - t1; ok += bool(res.valid)
    print(f"n={n}: {ntoys} toys, sample {1e3*t_s/ntoys:.2f} ms + fit {1e3*t_f/ntoys:.2f} ms per toy = {ntoys/(t_s+t_f):.1f} toys/s, valid {ok}/{ntoys}, "
          f"last fit passes {res.info['n_pass']}", flush=True)
