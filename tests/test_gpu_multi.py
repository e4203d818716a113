"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): sharded robust + MLE fits over the in-kernel NVLink
exchange against the CPU oracle — parameters, labels, iteration counts and R's final .Random.seed (the key stream is
generated in per-rank shares).  Runs scripts/mgpu_check.py under torchrun, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fits_match_oracle_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MGPU_QUICK="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU CHECK OK" in out.stdout, out.stdout[-2000:]
