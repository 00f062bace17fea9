"""World-size-2 `gloo` tests of the host-side multi-GPU logic, runnable without a GPU:
batch sharding, the statistics all-reduce and the fact that [counts | residual sums | SSE | N] summed over ranks yields
exactly the single-process losses, perplexity and codebook gradient.  Per-rank statistics come from the CPU oracle
(test infrastructure); the all-reduce goes through the package's TorchStatsComm."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vq_oracle as O
from vq_b200.distributed import TorchStatsComm, shard_bounds


def test_shard_bounds_partition_the_batch():
    for batch in (1, 2, 7, 8, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, D, W, K, beta = 7, 32, 101, 64, 0.25              # 7 items over 2 ranks: UNEVEN shards (4 + 3)
        rng = np.random.default_rng(7)                       # same data on every rank; each takes its shard
        z = rng.standard_normal((B, D, W), dtype=np.float32)
        cb = rng.standard_normal((K, D), dtype=np.float32)
        lo, hi = shard_bounds(B, world, rank)
        fwd = O.vq_forward(z[lo:hi], cb, beta)
        stats = torch.from_numpy(O.shard_stats(z[lo:hi], cb, fwd.indices).astype(np.float32))
        local = stats.clone()
        glob, done = TorchStatsComm().allreduce_async(stats)  # "overlap" mode: a global copy, the local buffer is untouched
        assert done is None and torch.equal(stats, local)
        TorchStatsComm().allreduce(stats)                    # the one exchange on the path, in place
        assert torch.equal(stats, glob)
        mse, com, ppl, dE = O.finalize_from_stats(stats.numpy().astype(np.float64), K, D, beta)
        # single-process reference on the concatenated batch
        full = O.vq_forward(z, cb, beta)
        _, dE_ref = O.vq_backward(z, cb, full.indices, beta, 1.0, 0.0, None)
        ok = (np.isclose(mse, full.embedding_loss, rtol=1e-5) and np.isclose(com, full.commitment_loss, rtol=1e-5)
              and np.isclose(ppl, full.perplexity, rtol=1e-5)
              and np.allclose(dE, dE_ref, rtol=1e-4, atol=1e-6 * np.abs(dE_ref).max())
              and np.array_equal(fwd.indices, full.indices[lo * W:hi * W]))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_stats_allreduce_world2_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
