"""bench.py's reference arm runs on the host alone, so its JSON contract can be checked without a GPU: one line on
stdout with the keys the driver reads, `impl: reference`, a cpu_baseline block and an e2e block that repeats the value."""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*extra):
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--ref-events", "200000", "--ref-budget", "5", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["unit"] == "evals/s" and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_config_and_nonzero_rank_exits_quietly(monkeypatch):
    d = _run("--config", "4")
    assert "cfg4" in d["config"]["workload"]
    env = {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}
    import os

    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, cwd=str(ROOT), env={**os.environ, **env})
    assert p.returncode == 0 and p.stdout.strip() == ""
