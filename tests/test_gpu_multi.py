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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices_agree():
    """Function attributes and launch plans are cached per device: the same process quantising on cuda:0 and then on cuda:1
    (D = 256 needs > 48 KB of dynamic shared memory in every kernel of the path) must get identical results."""
    import vq_b200
    outs = []
    g = torch.Generator().manual_seed(5)
    z = torch.randn(2, 256, 1536, generator=g)
    cb = torch.randn(700, 256, generator=g)
    Gq = torch.randn(2, 256, 1536, generator=g) * 1e-3
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        vq = vq_b200.VectorQuantizer(700, 256, 0.25).to(dev)
        with torch.no_grad():
            vq.codebook.weight.copy_(cb)
        x = z.to(dev).requires_grad_(True)
        emb, com, q, ppl, enc, idx = vq(x)
        (emb + com + (q * Gq.to(dev)).sum()).backward()
        outs.append((idx.cpu(), q.detach().cpu(), emb.item(), x.grad.cpu()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
        assert abs(o[2] - outs[0][2]) <= 1e-6 * abs(outs[0][2]) and torch.allclose(o[3], outs[0][3], rtol=1e-6, atol=1e-9)
