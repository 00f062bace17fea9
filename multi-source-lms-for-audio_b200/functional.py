"""Thin torch wrappers over the C ABI (include/vqb.h): device memory, streams and error translation only.

Every function requires CUDA tensors and raises otherwise - the hot path has no CPU implementation.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

# (device, stream) -> scratch tensor, least recently used first.  Bounded: a process that cycles through many streams must not
# pin the largest workspace it ever used on every one of them.
_workspaces: "collections.OrderedDict" = collections.OrderedDict()
_WORKSPACE_CACHE_ENTRIES = 4


def _require_cuda(name: str, t: torch.Tensor, dtype) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 vector quantiser has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def workspace_bytes(N: int, K: int, D: int, flags: int) -> int:
    out = C.c_size_t(0)
    L.check("vqb_workspace_bytes", L.lib().vqb_workspace_bytes(N, K, D, flags, C.byref(out)))
    return out.value


def workspace_bytes_bw(B: int, D: int, W: int, K: int, flags: int) -> int:
    """Exact scratch size for one [B, D, W] call (no bf16 latent copy when the tensor-core kernel reads the latents itself)."""
    out = C.c_size_t(0)
    L.check("vqb_workspace_bytes_bw", L.lib().vqb_workspace_bytes_bw(B, D, W, K, flags, C.byref(out)))
    return out.value


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) scratch, grown on demand and reused (stream order makes reuse safe); at most
    _WORKSPACE_CACHE_ENTRIES streams are remembered (least recently used dropped first), release_workspaces() drops all.

    Inside CUDA-graph capture the scratch is a FRESH allocation that is not cached: it comes from the graph's private
    memory pool and lives as long as the graph, so a later (larger) eager call can never free memory whose address is
    baked into a captured graph."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, _stream_ptr(device))
    ws = _workspaces.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    _workspaces[key] = ws                       # most recently used last
    while len(_workspaces) > _WORKSPACE_CACHE_ENTRIES:
        _workspaces.popitem(last=False)
    return ws


def release_workspaces() -> None:
    """Drop every cached scratch tensor (they are only a cache: the next call allocates again)."""
    _workspaces.clear()


def _check_codes(idx: torch.Tensor, K: int) -> None:
    """Caller-supplied indices (e.g. BERT-predicted tokens, bert.py:72-78): refuse codes outside [0, K) on the host - the
    reference's scatter_ / one-hot matmul raise a device-side assert there.  One small reduction + sync; pass
    validate=False on a hot path (the kernels are memory-safe either way: out-of-range codes give NaN / an all-zero row)."""
    if idx.numel() and bool(((idx < 0) | (idx >= K)).any()):
        raise IndexError(f"index out of range: codes must lie in [0, {K})")


def vq_forward(z: torch.Tensor, codebook: torch.Tensor, *, precision: str = "bf16", want_q: bool = True,
               want_resid: bool = False, stats: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """z [B, D, W] fp32 (BCW, contiguous), codebook [K, D] fp32 -> (idx [N] int64, q_st [B, D, W] or None, stats).

    Replaces vector_quantizer.py:25-48,52; `stats` is [counts | residual sums | SSE | N] (include/vqb.h)."""
    _require_cuda("z", z, torch.float32)
    _require_cuda("codebook", codebook, torch.float32)
    if z.dim() != 3 or codebook.dim() != 2 or z.shape[1] != codebook.shape[1]:
        raise ValueError(f"expected z [B, D, W] and codebook [K, D], got {tuple(z.shape)} and {tuple(codebook.shape)}")
    if precision not in L.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(L.PRECISIONS)}, got {precision!r}")
    z = z.contiguous()
    codebook = codebook.contiguous()
    B, D, W = z.shape
    K = codebook.shape[0]
    N = B * W
    flags = L.PRECISIONS[precision] | (L.WANT_Q if want_q else 0) | (L.WANT_RESID if want_resid else 0)
    with torch.cuda.device(z.device):
        ws = workspace if workspace is not None else _workspace(z.device, workspace_bytes_bw(B, D, W, K, flags))
        idx = torch.empty(N, dtype=torch.int64, device=z.device)
        q = torch.empty_like(z) if want_q else None
        if stats is None:
            stats = torch.empty(L.stats_len(K, D), dtype=torch.float32, device=z.device)
        L.check("vqb_forward", L.lib().vqb_forward(z.data_ptr(), codebook.data_ptr(), B, D, W, K, flags, idx.data_ptr(), _ptr(q),
                                                   stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(z.device)))
    return idx, q, stats


def vq_finalize(stats: torch.Tensor, K: int, D: int, beta: float) -> torch.Tensor:
    """stats -> tensor [embedding_loss, commitment_loss, perplexity] (vector_quantizer.py:45-46,49-50)."""
    _require_cuda("stats", stats, torch.float32)
    losses = torch.empty(3, dtype=torch.float32, device=stats.device)
    with torch.cuda.device(stats.device):
        L.check("vqb_finalize", L.lib().vqb_finalize(stats.data_ptr(), K, D, float(beta), losses.data_ptr(), _stream_ptr(stats.device)))
    return losses


def vq_backward(z: torch.Tensor, codebook: torch.Tensor, idx: torch.Tensor, stats: torch.Tensor, Gq: Optional[torch.Tensor],
                g_e: Optional[torch.Tensor], g_c: Optional[torch.Tensor], beta: float, need_dx: bool = True,
                need_de: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Gradients of SURVEY.md row a12; g_e / g_c are 0-dim CUDA tensors (or None = 0), read on the device."""
    _require_cuda("z", z, torch.float32)
    _require_cuda("codebook", codebook, torch.float32)
    _require_cuda("idx", idx, torch.int64)
    B, D, W = z.shape
    K = codebook.shape[0]
    z = z.contiguous()
    codebook = codebook.contiguous()
    if Gq is not None:
        Gq = Gq.to(torch.float32).contiguous()
    g_e = None if g_e is None else g_e.to(torch.float32).reshape(1).contiguous()
    g_c = None if g_c is None else g_c.to(torch.float32).reshape(1).contiguous()
    dX = torch.empty_like(z) if need_dx else None
    dE = torch.empty_like(codebook) if need_de else None
    with torch.cuda.device(z.device):
        L.check("vqb_backward", L.lib().vqb_backward(z.data_ptr(), codebook.data_ptr(), idx.data_ptr(), stats.data_ptr(), _ptr(Gq),
                                                     _ptr(g_e), _ptr(g_c), float(beta), B, D, W, K, _ptr(dX), _ptr(dE),
                                                     _stream_ptr(z.device)))
    return dX, dE


def ema_update(stats: torch.Tensor, codebook: torch.Tensor, cluster_size: torch.Tensor, embed_sum: torch.Tensor, decay: float = 0.99,
               eps: float = 1e-5) -> None:
    """EXTENSION (not in the reference): in-place EMA codebook update from a statistics buffer produced with want_resid=True.
    cluster_size is [K + 1] fp32 state (slot K = running total), embed_sum [K, D] fp32 state."""
    for name, t in (("stats", stats), ("codebook", codebook), ("cluster_size", cluster_size), ("embed_sum", embed_sum)):
        _require_cuda(name, t, torch.float32)
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    K, D = codebook.shape
    if cluster_size.numel() != K + 1 or embed_sum.shape != codebook.shape or stats.numel() != L.stats_len(K, D):
        raise ValueError("ema_update: shape mismatch")
    with torch.cuda.device(codebook.device):
        L.check("vqb_ema_update", L.lib().vqb_ema_update(stats.data_ptr(), codebook.data_ptr(), cluster_size.data_ptr(), embed_sum.data_ptr(),
                                                         K, D, float(decay), float(eps), _stream_ptr(codebook.device)))


def onehot(idx: torch.Tensor, K: int, validate: bool = True) -> torch.Tensor:
    """Dense `encodings` [N, K] fp32 (vector_quantizer.py:38-39)."""
    _require_cuda("idx", idx, torch.int64)
    idx = idx.reshape(-1).contiguous()
    if validate:
        _check_codes(idx, K)
    out = torch.empty(idx.numel(), K, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        L.check("vqb_onehot", L.lib().vqb_onehot(idx.data_ptr(), idx.numel(), K, out.data_ptr(), _stream_ptr(idx.device)))
    return out


def gather(codebook: torch.Tensor, idx: torch.Tensor, B: int, W: int, validate: bool = True) -> torch.Tensor:
    """De-quantise: [B, D, W] with out[b, :, w] = codebook[idx[b*W + w]] (vector_quantizer.py:42, bert.py:75-78)."""
    _require_cuda("codebook", codebook, torch.float32)
    _require_cuda("idx", idx, torch.int64)
    codebook = codebook.contiguous()
    idx = idx.reshape(-1).contiguous()
    K, D = codebook.shape
    if validate:
        _check_codes(idx, K)
    if idx.numel() != B * W:
        raise ValueError(f"idx has {idx.numel()} entries, expected B*W = {B * W}")
    out = torch.empty(B, D, W, dtype=torch.float32, device=codebook.device)
    with torch.cuda.device(codebook.device):
        L.check("vqb_gather", L.lib().vqb_gather(codebook.data_ptr(), idx.data_ptr(), B, D, W, K, out.data_ptr(),
                                                 _stream_ptr(codebook.device)))
    return out


def window_indices(idx: torch.Tensor, batch: int, window: int = 512, pad_id: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Index stream -> BERT windows (bert.py:50-69): tokens [B, n_win, window] int64, mask [B, n_win, window] fp32."""
    _require_cuda("idx", idx, torch.int64)
    idx = idx.reshape(-1).contiguous()
    if idx.numel() % batch:
        raise ValueError("index stream length is not a multiple of the batch size")
    Lseq = idx.numel() // batch
    n_win = (Lseq + window - 1) // window
    tokens = torch.empty(batch, n_win, window, dtype=torch.int64, device=idx.device)
    mask = torch.empty(batch, n_win, window, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        L.check("vqb_window_indices", L.lib().vqb_window_indices(idx.data_ptr(), batch, Lseq, window, pad_id, tokens.data_ptr(),
                                                                 mask.data_ptr(), _stream_ptr(idx.device)))
    return tokens, mask


def vq_forward_host(z_host: torch.Tensor, codebook_host: torch.Tensor, *, precision: str = "bf16", want_resid: bool = False,
                    chunk_batches: int = 0, idx_out: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None,
                    want_q: bool = False, q_out: Optional[torch.Tensor] = None, comm=None):
    """Host-buffer path (vqb_forward_host): z/codebook in (ideally pinned) HOST memory -> idx, stats in host memory; with
    want_q also the straight-through output [B, D, W] (returned as a third item).  `comm` (distributed.StatsComm): all-reduce
    the statistics over the ranks before they are copied back.  Runs on the CURRENT CUDA device."""
    if z_host.is_cuda or codebook_host.is_cuda:
        raise RuntimeError("vq_forward_host takes host tensors; use vq_forward for device tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 vector quantiser has no CPU fallback")
    z_host = z_host.contiguous()
    codebook_host = codebook_host.contiguous()
    B, D, W = z_host.shape
    K = codebook_host.shape[0]
    flags = L.PRECISIONS[precision] | (L.WANT_RESID if want_resid else 0) | (L.WANT_Q if want_q else 0)
    if idx_out is None:
        idx_out = torch.empty(B * W, dtype=torch.int64).pin_memory()
    if stats_out is None:
        stats_out = torch.empty(L.stats_len(K, D), dtype=torch.float32).pin_memory()
    if want_q and q_out is None:
        q_out = torch.empty((B, D, W), dtype=torch.float32).pin_memory()
    L.check("vqb_forward_host", L.lib().vqb_forward_host(z_host.data_ptr(), codebook_host.data_ptr(), B, D, W, K, flags,
                                                         idx_out.data_ptr(), q_out.data_ptr() if want_q else None,
                                                         stats_out.data_ptr(), chunk_batches,
                                                         comm.handle if comm is not None else None))
    return (idx_out, stats_out, q_out) if want_q else (idx_out, stats_out)


def debug_counters(device=None) -> dict:
    """Shortlist diagnostics of the last vq_forward on the current stream's cached workspace (synchronises)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ws = _workspaces.get((device.index, _stream_ptr(device)))
    if ws is None:
        raise RuntimeError("no cached workspace on this stream")
    torch.cuda.synchronize(device)
    out = (C.c_int64 * 3)()
    L.check("vqb_debug_counters", L.lib().vqb_debug_counters(ws.data_ptr(), out))
    return {"rescored": out[0], "fallback": out[1], "shortlisted": out[2]}


def debug_tc_scores(z: torch.Tensor, codebook: torch.Tensor, precision: str = "bf16") -> torch.Tensor:
    """Raw tensor-core scores |e|^2 - 2 lp(x).lp(e), [N, K], lp = bf16 or tf32; test hook for the tcgen05 tile."""
    _require_cuda("z", z, torch.float32)
    _require_cuda("codebook", codebook, torch.float32)
    z = z.contiguous()
    codebook = codebook.contiguous()
    B, D, W = z.shape
    K = codebook.shape[0]
    flags = L.PRECISIONS[precision]
    with torch.cuda.device(z.device):
        ws = _workspace(z.device, workspace_bytes(B * W, K, D, flags))
        out = torch.full((B * W, K), float("nan"), dtype=torch.float32, device=z.device)
        L.check("vqb_debug_tc_scores", L.lib().vqb_debug_tc_scores(z.data_ptr(), codebook.data_ptr(), B, D, W, K, flags, out.data_ptr(),
                                                                   ws.data_ptr(), ws.numel(), _stream_ptr(z.device)))
    return out
