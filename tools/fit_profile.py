"""cProfile of the host side of a small fit (cfg 1: the latency-bound config): where the time outside the kernels goes.

    python tools/fit_profile.py [n_fits]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zfit_b200 as zfit
from zfit_b200 import workloads as W

n_fits = int(sys.argv[1]) if len(sys.argv) > 1 else 300
w = W.build(1, device="cuda:0")
loss = w["loss"]
params = list(loss.get_params())
start = [p.value() for p in params]
minimizer = zfit.minimize.Minuit(gradient="zfit")
for _ in range(5):
    zfit.param.set_values(params, start)
    minimizer.minimize(loss)
t0 = time.perf_counter()
for _ in range(n_fits):
    zfit.param.set_values(params, start)
    minimizer.minimize(loss)
print(f"{1e6 * (time.perf_counter() - t0) / n_fits:.1f} us per fit (unprofiled)")
pr = cProfile.Profile()
pr.enable()
for _ in range(n_fits):
    zfit.param.set_values(params, start)
    res = minimizer.minimize(loss)
pr.disable()
print(res.info.get("driver"), res.info.get("n_pass"))
pstats.Stats(pr).sort_stats("cumulative").print_stats(32)
