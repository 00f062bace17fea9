"""torchrun worker for the multi-GPU parity test: every rank quantises its batch shard through the drop-in module with the
NCCL statistics all-reduce of libvqb_b200.so, and rank 0 compares losses / perplexity / codebook gradient with a
single-GPU run over the whole batch.  Launched by tests/test_gpu_multi.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200  # noqa: E402
from vq_b200.distributed import StatsComm, TorchStatsComm, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, D, W, K, beta = 8 * world, 64, 1000, 512, 0.25
    g = torch.Generator(device="cpu").manual_seed(3)
    z = torch.randn(B, D, W, generator=g).to(dev)
    cb = torch.randn(K, D, generator=g).to(dev)
    lo, hi = shard_bounds(B, world, rank)
    results = {}
    for name, comm in (("vqb_nccl", StatsComm()), ("torch_nccl", TorchStatsComm())):
        vq = vq_b200.VectorQuantizer(K, D, beta, stats_comm=comm).to(dev)
        with torch.no_grad():
            vq.codebook.weight.copy_(cb)
        x = z[lo:hi].clone().requires_grad_(True)
        emb, com, q, ppl, enc, idx = vq(x)
        (emb + com).backward()
        results[name] = (emb.item(), com.item(), ppl.item(), vq.codebook.weight.grad.clone(), idx.reshape(-1).clone())
    ok = True
    if rank == 0:
        ref = vq_b200.VectorQuantizer(K, D, beta).to(dev)
        with torch.no_grad():
            ref.codebook.weight.copy_(cb)
        x = z.clone().requires_grad_(True)
        emb, com, q, ppl, enc, idx = ref(x)
        (emb + com).backward()
        for name, (e, c, p, dE, ix) in results.items():
            ok &= abs(e - emb.item()) <= 1e-5 * abs(emb.item()) and abs(c - com.item()) <= 1e-5 * abs(com.item())
            ok &= abs(p - ppl.item()) <= 1e-5 * abs(ppl.item())
            ok &= torch.allclose(dE, ref.codebook.weight.grad, rtol=1e-4, atol=1e-6 * float(ref.codebook.weight.grad.abs().max()))
            ok &= torch.equal(ix, idx.reshape(-1)[lo * W:hi * W])
        print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
