"""Launch the second-order kernels of one config once each (for an ncu capture of hess_kernel / cov_kernel).

    python tools/aux_kernels_run.py [cfg]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zfit_b200 import workloads as W

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = W.build(cfg, device="cuda:0")
loss = w["loss"]
params = list(loss.get_params())
loss.value_gradient(params, full=False)
for _ in range(2):
    H = loss.hessian(params)
    C = loss.information_matrix(params)
print(loss.last_hessian_method, H.shape, C.shape)
