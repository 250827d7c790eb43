# achievable read bandwidth of the step kernel's access pattern with (almost) no arithmetic: Prod of Exponentials
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zfit_b200 as zfit
n = 100_000_000
for ncol in (1, 2, 3, 4):
    obs = [zfit.Space(f"x{k}", 0.0, 5.0) for k in range(ncol)]
    lams = [zfit.Parameter(f"lam{k}_s{ncol}", -0.3 - 0.1 * k, -3, 3) for k in range(ncol)]
    pdfs = [zfit.pdf.Exponential(l, o) for l, o in zip(lams, obs)]
    model = pdfs[0] if ncol == 1 else zfit.pdf.ProductPDF(pdfs)
    cols = {f"x{k}": torch.rand(n, dtype=torch.float64, device="cuda:0") * 5.0 for k in range(ncol)}
    sp = obs[0]
    for o in obs[1:]: sp = sp * o
    data = zfit.Data(cols, obs=sp, guarantee_limits=True)
    loss = zfit.loss.UnbinnedNLL(model, data)
    params = list(loss.get_params())
    loss.value_gradient(params, full=False)
    eng = loss._get_engine(); plan = eng.plan
    slots = np.array([p.value() for p in eng.slot_params])
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): plan.launch(slots, grad=True, log_offset=0.0, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): plan.launch(slots, grad=True, log_offset=0.0, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"{ncol} streams: {eng.kernel_names} {ms:.4f} ms  {8*ncol*n/ms/1e6:.0f} GB/s", flush=True)
    del loss, data, cols
