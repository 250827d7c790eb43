"""Where the per-step time goes above the kernel: C-ABI eval vs Python API (cfg 2, 1e8 events)."""
import os, sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zfit_b200 as zfit
from zfit_b200 import workloads as W
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = W.build(cfg, device="cuda:0")
loss = w["loss"]; params = list(loss.get_params()); th0 = np.array([p.value() for p in params])
loss.value_gradient(params, full=False)
eng = loss._get_engine(); plan = eng.plan
slots = np.array([p.value() for p in eng.slot_params]); off = loss.options["subtr_const_value"]
st = torch.cuda.current_stream().cuda_stream
K = 500
def timeit(fn):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(K): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / K * 1e3
def launches():
    plan.launch(slots, grad=True, log_offset=off, stream=st)
t_kernel = None
for _ in range(20): launches()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K): launches()
e1.record(); torch.cuda.synchronize(); t_kernel = e0.elapsed_time(e1) / K
t_eval = timeit(lambda: plan.eval(slots, grad=True, log_offset=off, stream=st))
t_api = timeit(lambda: loss.value_gradient(params, full=False))
def step():
    for p, v in zip(params, th0): p.assign(v * (1 + 1e-9))
    loss.value_gradient(params, full=False)
t_step = timeit(step)
print(f"cfg{cfg}: kernel back-to-back {t_kernel:.4f} ms | C-ABI eval (launch+collect, sync) {t_eval:.4f} | loss.value_gradient {t_api:.4f} | assign+value_gradient {t_step:.4f}")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
