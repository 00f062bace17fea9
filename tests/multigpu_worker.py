"""torchrun worker for the multi-GPU parity test: every rank quantises its batch shard through the drop-in module with the
NCCL statistics all-reduce of libvqb_b200.so, and rank 0 compares losses / perplexity / codebook gradient / latent gradient
with a single-GPU run over the whole batch.  Launched by tests/test_gpu_multi.py.

Covers: the library's communicator and torch.distributed's, stats_sync "forward" (global losses) and "overlap" (local losses,
side-stream exchange joined in backward), even and UNEVEN shards, and the gradient contract of the module's docstring:
dE is the global mean gradient on every rank; dX is normalised by the rank's own frame count (what DDP expects to average)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vq_b200  # noqa: E402
from vq_b200.distributed import StatsComm, TorchStatsComm, shard_bounds  # noqa: E402


def close(a, b, rtol):
    return abs(a - b) <= rtol * abs(b)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, W, K, beta = 64, 1000, 512, 0.25
    ok = True
    vqb_comm = StatsComm()
    for B in (8 * world, 8 * world + 1):                       # even shards, then one rank with an extra batch item
        g = torch.Generator(device="cpu").manual_seed(3 + B)
        z = torch.randn(B, D, W, generator=g).to(dev)
        cb = torch.randn(K, D, generator=g).to(dev)
        Gq = (torch.randn(B, D, W, generator=g) * 1e-3).to(dev)
        lo, hi = shard_bounds(B, world, rank)
        # single-GPU truth on the concatenated batch (every rank computes it: rank r needs its own slice of dX)
        ref = vq_b200.VectorQuantizer(K, D, beta).to(dev)
        with torch.no_grad():
            ref.codebook.weight.copy_(cb)
        x = z.clone().requires_grad_(True)
        emb, com, q, ppl, enc, idx = ref(x)
        (emb + com + (q * Gq).sum()).backward()
        n_all, n_loc = B * W, (hi - lo) * W
        # dX = Gq + c (x - q) / N: the commitment part scales with 1 / N, the upstream part does not
        # (in float64: the commitment part is ~1e-4 of the upstream part, a float32 difference would keep 3 digits of it)
        dx_ref = Gq[lo:hi].double() + (x.grad.double() - Gq.double())[lo:hi] * (n_all / n_loc)
        for name, comm, sync in (("vqb_nccl/forward", vqb_comm, "forward"), ("torch_nccl/forward", TorchStatsComm(), "forward"),
                                 ("vqb_nccl/overlap", vqb_comm, "overlap"), ("torch_nccl/overlap", TorchStatsComm(), "overlap")):
            vq = vq_b200.VectorQuantizer(K, D, beta, stats_comm=comm, stats_sync=sync).to(dev)
            with torch.no_grad():
                vq.codebook.weight.copy_(cb)
            xs = z[lo:hi].clone().requires_grad_(True)
            e, c, qs, p, _, ix = vq(xs)
            (e + c + (qs * Gq[lo:hi]).sum()).backward()
            checks = {"idx": torch.equal(ix.reshape(-1), idx.reshape(-1)[lo * W:hi * W]), "q": torch.equal(qs, q[lo:hi])}
            dE_ref = ref.codebook.weight.grad
            checks["dE"] = torch.allclose(vq.codebook.weight.grad, dE_ref, rtol=1e-4, atol=1e-6 * float(dE_ref.abs().max()))
            checks["dX"] = torch.allclose(xs.grad.double(), dx_ref, rtol=1e-5, atol=2e-7 * float(dx_ref.abs().max()))
            if sync == "forward":                               # global losses: identical to the single-process run
                checks["losses"] = (close(e.item(), emb.item(), 1e-5) and close(c.item(), com.item(), 1e-5)
                                    and close(p.item(), ppl.item(), 1e-5))
            else:                                               # local losses: this shard alone (reference under DDP)
                loc = vq_b200.VectorQuantizer(K, D, beta).to(dev)
                with torch.no_grad():
                    loc.codebook.weight.copy_(cb)
                el, cl, _, pl, _, _ = loc(z[lo:hi])
                checks["local_losses"] = close(e.item(), el.item(), 1e-6) and close(p.item(), pl.item(), 1e-6)
            good = all(checks.values())
            if not good:
                print(f"rank {rank}: MISMATCH in {name} at B={B}: failed {[k for k, v in checks.items() if not v]}", flush=True)
            ok &= bool(good)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
    if rank == 0:
        print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
    dist.barrier()
    vqb_comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
