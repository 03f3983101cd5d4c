"""GPU, needs >= 2 devices: one process per GPU over NCCL; sharded result == oracle (scripts/check_sharded.py)."""
import os
import subprocess
import sys

import pytest

import evo_ssearch_b200 as evs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(evs.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_nccl_equals_oracle():
    g = min(evs.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_sharded.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SHARDED_PARITY_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
