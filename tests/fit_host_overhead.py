"""Host time of one Minuit fit with the engine taken out: the kernel's event loop runs on the host (tests/_hostsim, test
infrastructure) behind a memo keyed by the slot values, so from the second fit on every engine call is a dictionary
lookup and what remains is the Python / ctypes / C++-driver layer.  CPU only.

    python tests/fit_host_overhead.py [cfg]
"""
import sys, time, numpy as np, cProfile, pstats, io
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import zfit_b200 as zfit
from zfit_b200 import workloads as W
from tests._hostsim import HostSimEngine
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = W.build(cfg, n=10000 if cfg == 1 else 50000, device="cpu")
loss = w["loss"]
class Eng(HostSimEngine):
    memo = {}; t = 0.0
    def eval(self, slots, grad=True, log_offset=0.0, kahan=False):
        t0 = time.perf_counter()
        k = (tuple(np.asarray(slots).tolist()), grad, log_offset)
        r = Eng.memo.get(k)
        if r is None: r = Eng.memo[k] = super().eval(slots, grad=grad, log_offset=log_offset, kahan=kahan)
        Eng.t += time.perf_counter() - t0
        return r[0].copy(), (None if r[1] is None else r[1].copy())
    def eval_cov(self, slots):
        t0 = time.perf_counter()
        k = ("cov", tuple(np.asarray(slots).tolist()))
        r = Eng.memo.get(k)
        if r is None: r = Eng.memo[k] = super().eval_cov(slots)
        Eng.t += time.perf_counter() - t0
        return [m.copy() for m in r]
loss._set_engine_factory(lambda m, d, f: Eng(m, d, f))
params = list(loss.get_params()); th0 = [p.value() for p in params]
mini = zfit.minimize.Minuit(gradient="zfit", tol=1e-3)
def fit():
    for p, v in zip(params, th0): p.set_value(v)
    return mini.minimize(loss)
for _ in range(3): r = fit()
Eng.t = 0; K = 200
t = time.perf_counter()
for _ in range(K): r = fit()
wall = (time.perf_counter() - t) / K
print(f"cfg{cfg}: fit wall {wall*1e6:.0f} us with a memoised engine ({Eng.t/K*1e6:.0f} us inside it), passes {r.info.get('n_pass')}")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fit()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(40); print(s.getvalue()[:6000])
print("=========== cumulative")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:7000])
