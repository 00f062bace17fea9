"""Multi-GPU parity (runs only where >= 2 GPUs are visible, e.g. `gpurun --gpus 2`): batch-sharded ranks + the statistics
all-reduce reproduce the single-GPU losses, perplexity and codebook gradient."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_stats_allreduce_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
