"""GPU (needs >= 2 devices): NCCL path of the event-sharded loss against the single-GPU loss."""

import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_gpu_nccl_matches_single_gpu(tmp_path):
    from zfit_b200 import _znll as Z

    if Z.lib().znll_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "_dist_gpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = json.loads(out.read_text())
    assert all(got[m]["valid"] for m in ("fused", "nccl", "nccl_det"))
    assert got["nccl"]["reduce_mode"] == "nccl"
    # the fused NVLink exchange is the product path; if symmetric memory is unavailable the reason must be visible
    assert got["fused"]["reduce_mode"] == "fused-nvlink", got["fused"]["fallback"]
    assert abs(got["fused"]["value"] - got["nccl"]["value"]) <= 1e-12 * abs(got["nccl"]["value"])
