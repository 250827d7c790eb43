"""Full-size Minuit fits of the BASELINE configs on cuda:0: wall time, fused passes, information-pass cost.

usage: python tools/fit_timing.py 1,2,3,4,5 [verbose]
"""
import os
sys_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys, time, numpy as np, torch
sys.path.insert(0, sys_path)
import zfit_b200 as zfit
from zfit_b200 import workloads as W
cfgs = [int(c) for c in sys.argv[1].split(',')]
verb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for cfg in cfgs:
    w = W.build(cfg, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    th0 = [p.value() for p in params]
    loss.value_gradient(params, full=False)
    torch.cuda.synchronize()
    t = time.perf_counter(); H = loss.information_matrix(params); torch.cuda.synchronize(); t_info = time.perf_counter() - t
    t = time.perf_counter(); H = loss.information_matrix(params); torch.cuda.synchronize(); t_info2 = time.perf_counter() - t
    t = time.perf_counter(); loss.value_gradient(params, full=False); torch.cuda.synchronize(); t_vg = time.perf_counter() - t
    for rep in range(2):
        for p, v in zip(params, th0): p.assign(v)
        torch.cuda.synchronize(); t = time.perf_counter()
        res = zfit.minimize.Minuit(gradient="zfit", verbosity=7 if (verb and rep == 1) else 0).minimize(loss)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"cfg{cfg} rep{rep} fit {dt*1e3:.2f} ms passes {res.info['n_pass']} iters {res.info['n_iter']} valid {res.valid} edm {res.edm:.3g} fmin {res.fmin!r} info_pass {t_info*1e3:.2f}/{t_info2*1e3:.2f} ms vg {t_vg*1e3:.2f} ms", flush=True)
    print('   ', [float(res.params[p]['value']) for p in res.params])
    del loss, w
