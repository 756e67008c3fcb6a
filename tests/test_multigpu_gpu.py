"""GPU, >= 2 devices: the sharded eval path over NCCL equals the single-device result.
Launched through torchrun (one process per GPU). Skipped on 1-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_eval_matches_single_device():
    world = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "sharded eval ok" in p.stdout
    print([ln for ln in p.stdout.splitlines() if "sharded eval ok" in ln][-1])      # shown with pytest -rP: world size, transports
