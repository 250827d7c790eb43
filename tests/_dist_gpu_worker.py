"""Multi-GPU worker (NCCL): event-sharded CUDA engine, one all-reduce of the raw accumulators per step."""

import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as td

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import zfit_b200 as zfit  # noqa: E402
from tests._dist_worker import build  # noqa: E402
from zfit_b200 import distributed as D  # noqa: E402


def main():
    out = sys.argv[1]
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    import datetime

    td.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=120))
    rank, world = td.get_rank(), td.get_world_size()
    rng = np.random.default_rng(11)
    n = 2_000_001
    x = np.concatenate([rng.normal(1.0, 1.5, n // 3), rng.uniform(-10, 10, n - n // 3)])
    x = x[(x >= -10) & (x <= 10)]
    w = rng.normal(1.0, 0.6, x.size)
    res = {}
    for mode, opts in (("fused", {}), ("nccl", {"fused_reduce": False}), ("nccl_det", {"fused_reduce": False, "deterministic_reduce": True})):
        model, full, obs = build(x, w, f"_g{mode}")
        nll = zfit.loss.UnbinnedNLL(model, D.shard(full), options=opts)
        params = list(nll.get_params())
        v, g = nll.value_gradient(params, full=True)
        mine = torch.tensor([v, *g], dtype=torch.float64, device=f"cuda:{local_rank}")
        gathered = [torch.empty_like(mine) for _ in range(world)]
        td.all_gather(gathered, mine)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree bitwise"
        single = zfit.loss.UnbinnedNLL(model, full, options={"distributed": False})
        sv, sg = single.value_gradient(params, full=True)
        assert abs(v - sv) <= 1e-12 * abs(sv), (v, sv)
        assert np.all(np.abs(g - sg) <= 1e-10 * np.abs(sg).max()), (g, sg)
        # the one-pass Hessian all-reduces its per-rank results: same matrix as the single-GPU loss on the whole sample
        H, Hs = nll.hessian(params), single.hessian(params)
        assert nll.last_hessian_method == "one-pass" and single.last_hessian_method == "one-pass"
        assert np.all(np.abs(H - Hs) <= 1e-9 * np.sqrt(np.abs(np.outer(np.diag(Hs), np.diag(Hs))))), (H, Hs)
        # the C-level theta path is active exactly where the reduction is fused into the kernel
        assert (nll._theta_path(nll._get_engine()) is not None) == (mode == "fused")
        start = [float(p.value()) for p in params]
        result = zfit.minimize.Minuit(gradient="zfit").minimize(nll)
        fitted = torch.tensor([result.params[p]["value"] for p in result.params], dtype=torch.float64, device=f"cuda:{local_rank}")
        gathered = [torch.empty_like(fitted) for _ in range(world)]
        td.all_gather(gathered, fitted)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        # with the fused exchange every rank runs the whole fit in native code (metric seed all-reduced through the hook);
        # it ends at the minimum the single-GPU fit of the whole sample finds
        native = "whole fit in native code" in result.info["driver"]
        assert native == (mode == "fused"), result.info["driver"]
        if mode == "fused":
            for p, val in zip(params, start):
                p.assign(val)
            alone = zfit.minimize.Minuit(gradient="zfit").minimize(single)
            a = np.array([result.params[p]["value"] for p in params])
            b = np.array([alone.params[p]["value"] for p in params])
            errs = np.sqrt(np.diag(np.linalg.inv(Hs)))
            assert np.all(np.abs(a - b) <= 1e-3 * errs), (a, b, errs)
        eng = nll._get_engine()
        res[mode] = {"value": v, "single": sv, "valid": bool(result.valid), "n_eval": int(result.info["n_eval"]),
                     "kernel": eng.kernel_names, "reduce_mode": eng.reduce_mode,
                     "fallback": getattr(eng, "reduce_fallback_reason", None)}
    if rank == 0:
        json.dump(res, open(out, "w"))
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
