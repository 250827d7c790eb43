"""One-time cost of a fit that starts from HOST numpy columns: Data -> HBM -> bind -> first value_gradient."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zfit_b200 as zfit
from zfit_b200 import workloads as W

for cfg in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "2,4").split(",")]:
    w = W.build(cfg, device="cuda:0")
    loss = w["loss"]
    params = list(loss.get_params())
    loss.value_gradient(params, full=False)
    off = loss.options["subtr_const_value"]
    host_cols = [[np.ascontiguousarray(c.cpu().numpy()) for c in d._cols] for d in w["data"]]
    host_w = [None if d._weights is None else np.ascontiguousarray(d._weights.cpu().numpy()) for d in w["data"]]
    nbytes = sum(c.nbytes for cols in host_cols for c in cols) + sum(0 if x is None else x.nbytes for x in host_w)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        datas = [zfit.Data(dict(zip(d.obs, cols)), obs=d.space, weights=hw, guarantee_limits=True)
                 for d, cols, hw in zip(w["data"], host_cols, host_w)]
        t1 = time.perf_counter()
        for d in datas: d.device_columns(0)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        loss2 = type(loss)(model=w["model"], data=datas, options={"subtr_const_value": off})
        t3 = time.perf_counter()
        loss2.value_gradient(params, full=False)
        torch.cuda.synchronize(); t4 = time.perf_counter()
        loss2.value_gradient(params, full=False)
        torch.cuda.synchronize(); t5 = time.perf_counter()
        print(f"cfg{cfg} rep{rep}: {nbytes/1e6:.0f} MB  Data() {1e3*(t1-t0):.1f}  upload {1e3*(t2-t1):.1f}  loss() {1e3*(t3-t2):.1f}  "
              f"first eval (bind, sort, sumw) {1e3*(t4-t3):.1f}  second eval {1e3*(t5-t4):.2f}  total {1e3*(t4-t0):.1f} ms", flush=True)
        del loss2, datas
    del loss, w
