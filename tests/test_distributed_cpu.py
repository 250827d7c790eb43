"""CPU: the N > 1 path (event sharding + one all-reduce per step) with world_size 2 on gloo."""

import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

import zfit_b200 as zfit
from tests._dist_worker import build
from tests._oracle_engine import use_oracle
from zfit_b200 import distributed as D

ROOT = Path(__file__).resolve().parent.parent


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 2, 7, 100, 101, 12_345_679):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(s[0] % 2 == 0 or s[0] == s[1] for s in spans)  # keeps 16-byte alignment of every shard for double2 loads
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(s for s in sizes if s or True) <= max(2, -(-n // world) + 1)


def test_world_size_2_gloo_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "dist.json"
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "_dist_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = json.loads(out.read_text())
    # single process on the whole sample
    rng = np.random.default_rng(11)
    n = 30_001
    x = np.concatenate([rng.normal(1.0, 1.5, n // 3), rng.uniform(-10, 10, n - n // 3)])
    x = x[(x >= -10) & (x <= 10)]
    w = rng.normal(1.0, 0.6, x.size)
    model, full, _ = build(x, w, "_s")
    assert got["n"] == full.num_entries
    nll = use_oracle(zfit.loss.UnbinnedNLL(model, full))
    v, g = nll.value_gradient(full=True)
    assert abs(got["value"] - v) <= 1e-12 * abs(v)
    np.testing.assert_allclose(got["grad"], g, rtol=0, atol=1e-10 * np.abs(g).max())
    result = zfit.minimize.Minuit(gradient="zfit").minimize(nll)
    assert got["valid"] and result.valid
    np.testing.assert_allclose(got["fitted"], [result.params[p]["value"] for p in result.params], rtol=1e-6, atol=1e-7)
