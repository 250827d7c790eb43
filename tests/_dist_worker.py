"""World-size-2 gloo worker (CPU): event sharding + per-step all-reduce, host logic only (oracle engine)."""

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
from tests._oracle_engine import OracleEngine  # noqa: E402
from zfit_b200 import distributed as D  # noqa: E402


def build(x, w, tag):
    obs = zfit.Space("x", -10, 10)
    mu = zfit.Parameter("mu" + tag, 0.8, -5, 5)
    sigma = zfit.Parameter("sigma" + tag, 1.7, 0.3, 6)
    fr = zfit.Parameter("frac" + tag, 0.45, 0, 1)
    cs = [zfit.Parameter(f"c{k + 1}" + tag, v, -1, 1) for k, v in enumerate((0.15, -0.06, 0.04, 0.02))]
    model = zfit.pdf.SumPDF([zfit.pdf.Gauss(mu, sigma, obs), zfit.pdf.Chebyshev(obs, cs)], fracs=fr)
    return model, zfit.Data(x, obs=obs, weights=w), obs


def main():
    out = sys.argv[1]
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    assert D.is_distributed() and D.world_size() == 2
    rng = np.random.default_rng(11)  # every rank draws the same full sample, then keeps its shard
    n = 30_001
    x = np.concatenate([rng.normal(1.0, 1.5, n // 3), rng.uniform(-10, 10, n - n // 3)])
    x = x[(x >= -10) & (x <= 10)]
    w = rng.normal(1.0, 0.6, x.size)
    model, full, obs = build(x, w, "_d")
    local = D.shard(full)
    lo, hi = D.shard_bounds(full.num_entries)
    assert lo % 2 == 0 and local.num_entries == hi - lo
    counts = torch.tensor([float(local.num_entries)])
    td.all_reduce(counts)
    assert int(counts[0]) == full.num_entries

    nll = zfit.loss.UnbinnedNLL(model, local)
    nll._set_engine_factory(OracleEngine)
    params = list(nll.get_params())
    v, g = nll.value_gradient(params, full=True)
    assert nll._get_engine().n_entries == [full.num_entries]
    assert abs(nll._get_engine().samplesize[0] - float(np.sum(full.weights))) < 1e-9

    # every rank must hold bit-identical reduced numbers (SPMD minimizers stay in lock-step)
    mine = torch.tensor([v, *g], dtype=torch.float64)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    td.all_gather(gathered, mine)
    assert all(torch.equal(gathered[0], t) for t in gathered)

    # a short SPMD fit: identical trajectory and result on both ranks
    result = zfit.minimize.Minuit(gradient="zfit").minimize(nll)
    fitted = torch.tensor([result.params[p]["value"] for p in result.params] + [float(result.info["n_eval"])], dtype=torch.float64)
    gathered = [torch.empty_like(fitted) for _ in range(world)]
    td.all_gather(gathered, fitted)
    assert all(torch.equal(gathered[0], t) for t in gathered)

    if rank == 0:
        json.dump({"value": v, "grad": list(map(float, g)), "fitted": list(map(float, fitted[:-1])), "valid": bool(result.valid),
                   "n": int(full.num_entries)}, open(out, "w"))
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
