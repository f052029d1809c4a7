"""Data-parallel (NCCL) equivalence on a multi-GPU box: skipped when fewer than 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_train_step_equals_single_rank():
    w = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={w}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "dp_equivalence_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dp-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
