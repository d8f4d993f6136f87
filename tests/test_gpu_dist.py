"""Data-parallel training on hardware (reference trainer.py:110-128,722-734 under accelerate / DDP): the
averaged gradients of two ranks equal the mean of the two half-batch gradients, and the replicas stay identical after the fused
optimizer step.  Needs two GPUs (gpurun --gpus 2); skipped on a single-GPU box.  The worker is tests/dist_grad_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_gradients_match_single_process_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dist_grad_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    print(out[-3000:])
    assert r.returncode == 0 and "DIST_OK" in out, out[-3000:]
