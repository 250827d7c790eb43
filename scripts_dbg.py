import numpy as np, torch, sys
sys.path.insert(0,'.')
import zfit_b200 as zfit
from zfit_b200 import workloads as W
for n in (1_000_000, 20_000_000, 100_000_000):
    w = W.build(2, n=n, device="cuda:0")
    x = w['data'][0]._cols[0]
    print(n, 'nan in data', bool(torch.isnan(x).any()), float(x.min()), float(x.max()), x.numel())
    loss = w['loss']; params = list(loss.get_params())
    print(' full=True ', loss.value_gradient(params, full=True))
    print(' full=False', loss.value_gradient(params, full=False))
    print(' value only', loss.value(full=True), loss.value(full=False), loss.options.get('subtr_const_value'))
    eng = loss._get_engine()
    slots = np.array([p.value() for p in eng.slot_params])
    for off in (0.0, 1.0, 6.0):
        print('  raw off', off, eng.plan.eval(slots, grad=True, log_offset=off))
    del w, loss, x; torch.cuda.empty_cache()
