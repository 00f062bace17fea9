"""A/B experiment for the round-1 finding that tail_tma_kernel runs +19 % slower whenever WORLD_SIZE > 1 (Next #1c).

Runs the cfg-3 training-mode forward on every rank and prints rank 0's per-stage CUDA-event times.  Variants:
  plain        no torch.distributed at all (single process)
  pg_only      process group initialised (NCCL, eager communicator), nothing else
  comm_idle    + StatsComm created (the library's own NCCL communicator), never called
  inplace      + in-place all-reduce of `stats` every step (round-1 behaviour)
  outofplace   + all-reduce of a COPY of `stats` (the buffer the tail's atomics hit is never touched by NCCL)
  sidestream   + all-reduce on a side stream, joined before finalize
  gloo_pg      process group over gloo (no NCCL in torch), StatsComm in-place all-reduce
usage: [torchrun ...] python scripts/exp_multi_tail.py --variant inplace [--B 256] [--steps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vq_b200  # noqa: F401
from vq_b200 import _lib, functional as F
from vq_b200.distributed import StatsComm

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="plain")
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--D", type=int, default=256)
ap.add_argument("--W", type=int, default=16384)
ap.add_argument("--K", type=int, default=8192)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-resid", action="store_true")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
v = args.variant
comm = None
if v != "plain":
    if v == "gloo_pg":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    if v != "pg_only":
        comm = StatsComm()

B, D, W, K = args.B, args.D, args.W, args.K
g = torch.Generator(device=dev).manual_seed(42 + rank)
z = torch.randn(B, D, W, device=dev, generator=g)
cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
stats = torch.empty(_lib.stats_len(K, D), device=dev)
stats2 = torch.empty_like(stats)
side = torch.cuda.Stream(device=dev)
lib = _lib.lib()


def step():
    idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=not args.no_resid, stats=stats)
    if comm is not None and v in ("inplace", "gloo_pg"):
        comm.allreduce(st)
    elif comm is not None and v == "outofplace":
        stats2.copy_(st)
        comm.allreduce(stats2)
        st = stats2
    elif comm is not None and v == "sidestream":
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            stats2.copy_(st)
            comm.allreduce(stats2)
            done = torch.cuda.Event()
            done.record()
        torch.cuda.current_stream().wait_event(done)
        st = stats2
    return F.vq_finalize(st, K, D, 0.25)


for _ in range(3):
    step()
if dist.is_initialized():
    dist.barrier()
torch.cuda.synchronize()
lib.vqb_debug_kernel_timing(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    losses = step()
e1.record()
torch.cuda.synchronize()
out = {"variant": v, "world": world, "resid": not args.no_resid, "N": B * W, "step_ms": round(e0.elapsed_time(e1) / args.steps, 3)}
for sid, name in enumerate(("search", "prep", "fallback", "tail", "pack")):
    ms, n = C.c_double(0), C.c_int(0)
    lib.vqb_debug_stage_time_ms(sid, C.byref(ms), C.byref(n))
    out[name] = round(ms.value / max(1, n.value), 3)
lib.vqb_debug_kernel_timing(0)
out["env"] = {k: os.environ[k] for k in os.environ if k.startswith(("NCCL_", "VQB_", "OMP_", "CUDA_"))}
if rank == 0:
    print(json.dumps(out), flush=True)
if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
