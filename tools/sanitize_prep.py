"""Small driver for compute-sanitizer runs of the bind-time kernels (memcheck / racecheck need small inputs).

    compute-sanitizer --tool memcheck  python tools/sanitize_prep.py
    compute-sanitizer --tool racecheck python tools/sanitize_prep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from zfit_b200 import _znll as Z

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
stream = torch.cuda.current_stream(0).cuda_stream
for n in (1, 33, 1000, 70_001):
    for n_cols, weighted in ((1, False), (2, True), (3, True)):
        cols = [torch.from_numpy(rng.uniform(-1.2, 1.2, n)).to(dev) for _ in range(n_cols)]
        w = torch.from_numpy(rng.normal(1, 0.3, n)).to(dev) if weighted else None
        dst = [torch.empty_like(c) for c in cols]
        dw = None if w is None else torch.empty_like(w)
        sp, dp = [c.data_ptr() for c in cols], [c.data_ptr() for c in dst]
        Z.partition_events(0, sp, dp, 0 if w is None else w.data_ptr(), 0 if dw is None else dw.data_ptr(), n_cols - 1, -1.0, 1.0, n,
                           stream=stream)
        torch.cuda.synchronize()
        x = cols[n_cols - 1].cpu().numpy()
        b = np.clip(np.floor((x + 1.0) * 16.0), 0, 31).astype(np.int64)
        assert np.array_equal(dst[n_cols - 1].cpu().numpy(), x[np.argsort(b, kind="stable")])
        kept = Z.cut_events(0, sp, dp, 0 if w is None else w.data_ptr(), 0 if dw is None else dw.data_ptr(), [-1.0] * n_cols,
                            [1.0] * n_cols, n, stream=stream)
        mask = np.ones(n, dtype=bool)
        for c in cols:
            v = c.cpu().numpy()
            mask &= (v >= -1.0) & (v <= 1.0)
        assert kept == int(mask.sum())
        assert np.array_equal(dst[0].cpu().numpy()[:kept], cols[0].cpu().numpy()[mask])
print("sanitize_prep ok")
