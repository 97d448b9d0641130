"""Data-parallel training over NVLink (needs >= 2 GPUs; skipped on a single-GPU box): launches tests/mgpu_worker.py under
torch.distributed.run, one process per GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_nvlink_allreduce_training(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(lines[-1][len("MGPU_RESULT "):])
    assert res["ok"] and res["peer_timeout_reported"], res
    for mode, m in res["modes"].items():
        assert "unavailable" in m or (m["replicated_bit_identical"] and m["count_ok"]), (mode, m)
    assert "unavailable" not in res["modes"]["ipc"], res
