"""Multi-GPU plumbing for the bottleneck: latents shard by batch, the codebook is replicated and the only exchange is a
sum-all-reduce of the statistics buffer [counts | residual sums | SSE | N].

`StatsComm` owns a NCCL communicator created through the C ABI (vqb_comm_*), bootstrapped over an existing
torch.distributed process group (any backend - only the 128-byte unique id travels through it).
`TorchStatsComm` does the same all-reduce through torch.distributed itself (NCCL on GPUs, gloo in the CPU tests of
the host-side logic).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous batch shard of `rank`; the first `batch % world` ranks take one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class TorchStatsComm:
    """stats all-reduce through torch.distributed (works for CUDA tensors over NCCL and CPU tensors over gloo)."""

    def __init__(self, group=None):
        self.group = group

    def allreduce(self, stats: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return stats

    def allreduce_async(self, stats: torch.Tensor):
        """Global copy of `stats` (the local buffer is left untouched) + the event to wait for (None: already complete)."""
        if stats.is_cuda:
            return _side_stream_allreduce(self, stats)
        out = stats.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out, None


_side_streams: dict = {}


def _side_stream_allreduce(comm, stats: torch.Tensor):
    """All-reduce a COPY of `stats` on a per-device side stream: the caller's stream only pays for recording one event, the
    exchange (and the wait for the slowest rank) overlaps whatever the caller enqueues next, and is joined by waiting on the
    returned event.  comm.timings (if a list) receives a CUDA-event pair around the exchange."""
    dev = stats.device
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    out = stats.clone()                       # on the caller's stream (a few microseconds): `stats` may be reused right away
    ready = torch.cuda.Event()
    ready.record(cur)
    timed = getattr(comm, "timings", None)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        t0 = t1 = None
        if timed is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(side)
        comm.allreduce(out)
        if timed is not None:
            t1.record(side)
            timed.append((t0, t1))
        done = torch.cuda.Event()
        done.record(side)
    out.record_stream(side)
    return out, done


def side_stream(device) -> "torch.cuda.Stream":
    """The per-device side stream the overlapped exchanges run on (created on first use)."""
    dev = torch.device(device)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    return side


class StatsComm:
    """NCCL communicator held by libvqb_b200.so; all-reduces on the caller's current CUDA stream."""

    def __init__(self, group=None, device: Optional[torch.device] = None):
        if not dist.is_initialized():
            raise RuntimeError("StatsComm needs an initialised torch.distributed process group for the bootstrap")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        ident = (C.c_ubyte * L.UNIQUE_ID_BYTES)()
        if self.rank == 0:
            L.check("vqb_comm_unique_id", L.lib().vqb_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * L.UNIQUE_ID_BYTES).from_buffer_copy(box[0])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check("vqb_comm_init", L.lib().vqb_comm_init(ident, self.rank, self.world, C.byref(handle)))
        self._comm = handle

    timings = None     # set to a list to collect (start, stop) CUDA-event pairs of the side-stream exchanges

    def allreduce_async(self, stats: torch.Tensor):
        """Global copy of `stats` on a side stream + the event to wait for (see _side_stream_allreduce)."""
        return _side_stream_allreduce(self, stats)

    @property
    def handle(self):
        """The raw communicator for C-ABI calls that take one (vqb_forward_host)."""
        return self._comm

    def allreduce(self, stats: torch.Tensor) -> torch.Tensor:
        if not stats.is_cuda or stats.dtype != torch.float32 or not stats.is_contiguous():
            raise RuntimeError("stats must be a contiguous fp32 CUDA tensor")
        with torch.cuda.device(stats.device):
            L.check("vqb_allreduce_stats", L.lib().vqb_allreduce_stats(self._comm, stats.data_ptr(), stats.numel(),
                                                                       torch.cuda.current_stream(stats.device).cuda_stream))
        return stats

    def close(self) -> None:
        if getattr(self, "_comm", None):
            L.check("vqb_comm_destroy", L.lib().vqb_comm_destroy(self._comm))
            self._comm = None
