"""Drop-in replacement for the reference's `VectorQuantizer` (src/model/components/vector_quantizer.py:6-54).

Same constructor, same attributes (`codebook` is an nn.Embedding, so the checkpoint key
`vector_quantizer.codebook.weight` is unchanged), same `forward(inputs[B, D, W])` 6-tuple, same autograd behaviour -
but the whole path runs in libvqb_b200.so (hand-written sm_100a CUDA) through `torch.autograd.Function`.
CUDA only: a CPU tensor raises instead of silently taking another path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import functional as F


class _VQFunction(torch.autograd.Function):
    """forward: vqb_forward [+ stats all-reduce] + vqb_finalize; backward: vqb_backward."""

    @staticmethod
    def forward(ctx, inputs: torch.Tensor, weight: torch.Tensor, beta: float, precision: str, comm, sync: str):
        z = inputs.contiguous()
        need_dw = ctx.needs_input_grad[1]
        idx, q, stats = F.vq_forward(z, weight.detach(), precision=precision, want_q=True, want_resid=need_dw)
        pending = None
        if comm is not None and sync == "forward":
            comm.allreduce(stats)             # the one exchange on the path (DDP's implicit all-reduce in the reference)
        elif comm is not None and need_dw:    # "overlap": side stream, joined in backward (behind the decoder's fwd + bwd)
            pending = comm.allreduce_async(stats)
        K, D = weight.shape
        losses = F.vq_finalize(stats, K, D, beta)
        emb, com, ppl = losses.unbind(0)
        ctx.beta = float(beta)
        ctx.pending = pending
        ctx.save_for_backward(z, weight, idx, stats)
        ctx.mark_non_differentiable(ppl, idx)
        return emb, com, q, ppl, idx

    @staticmethod
    def backward(ctx, g_emb, g_com, g_q, _g_ppl, _g_idx):
        z, weight, idx, stats = ctx.saved_tensors
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if ctx.pending is not None:           # global statistics from the side stream -> dE with the global N
            stats, done = ctx.pending
            if done is not None:
                torch.cuda.current_stream(z.device).wait_event(done)
        dX, dE = F.vq_backward(z, weight.detach(), idx, stats, g_q, g_emb, g_com, ctx.beta, need_dx=need_dx, need_de=need_dw)
        return dX, dE, None, None, None, None


class VectorQuantizer(nn.Module):
    """B200-native VQ bottleneck with the reference module's interface.

    Extra keyword-only knobs (all default to reference behaviour):
      precision        "bf16": tcgen05 shortlist + fp32 rescoring (same indices as fp32);
                       "tf32": the same with kind::tf32 MMAs straight from the fp32 latents and codebook (no converted
                       copies, ~3x tighter guard band, half the tensor rate: the choice for small D);
                       "fp32": exact CUDA-core search.
      dense_encodings  "auto" (dense one-hot when N*K*4 bytes <= dense_limit_bytes, else a sparse COO tensor of the same
                       shape), True (always dense, like the reference) or False (always sparse).
      stats_comm       object with .allreduce(stats) / .allreduce_async(stats) (see distributed.StatsComm) for batch-sharded
                       multi-GPU training: the latents shard by batch, the codebook is replicated.
      stats_sync       "forward" (default): the statistics are all-reduced inside forward; losses, perplexity and the codebook
                       gradient are GLOBAL (identical on every rank, equal to a single process on the concatenated batch).
                       "overlap": forward returns this rank's LOCAL losses / perplexity (what the reference logs under DDP:
                       no sync_dist, vqvae.py:68-69) and starts the all-reduce on a side stream; backward joins it, so the
                       exchange hides behind the decoder's forward + backward.  The codebook gradient is the same global one.

    Gradient contract under stats_comm (matches Lightning DDP around the reference, configs/trainer/default.yaml:9-10):
    `codebook.weight.grad` is already the global mean gradient on every rank - do NOT let a DDP wrapper average it again with
    something else (averaging identical tensors is harmless).  `inputs.grad` is normalised by this rank's OWN frame count, like
    the reference's per-rank loss: the encoder gradients it produces are expected to be averaged over ranks by the DDP wrapper
    of the encoder.  With unequal shards the two conventions differ by N_local * world / N_global per rank, exactly as they do
    for the reference under DDP.
    """

    def __init__(self, num_embedding: int, embedding_dim: int, commitment_cost: float, *, precision: str = "bf16",
                 dense_encodings="auto", dense_limit_bytes: int = 256 << 20, stats_comm=None, stats_sync: str = "forward"):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_embedding = num_embedding
        # codebook, initialised as the reference does (vector_quantizer.py:18-19)
        self.codebook = nn.Embedding(self.num_embedding, self.embedding_dim)
        self.codebook.weight.data.uniform_(-1 / self.num_embedding, 1 / self.num_embedding)
        self.commitment_cost = commitment_cost
        self.precision = precision
        self.dense_encodings = dense_encodings
        self.dense_limit_bytes = dense_limit_bytes
        self.stats_comm = stats_comm
        if stats_sync not in ("forward", "overlap"):
            raise ValueError(f"stats_sync must be 'forward' or 'overlap', got {stats_sync!r}")
        self.stats_sync = stats_sync

    def _encodings(self, idx: torch.Tensor) -> torch.Tensor:
        N, K = idx.numel(), self.num_embedding
        dense = self.dense_encodings
        if dense == "auto":
            dense = N * K * 4 <= self.dense_limit_bytes
        if dense:
            return F.onehot(idx, K)
        rows = torch.arange(N, device=idx.device)
        return torch.sparse_coo_tensor(torch.stack([rows, idx]), torch.ones(N, device=idx.device), (N, K))

    def forward(self, inputs: torch.Tensor):
        if not inputs.is_cuda:
            raise RuntimeError("VectorQuantizer (B200) needs CUDA inputs: there is no CPU fallback for the hot path")
        if inputs.dim() != 3 or inputs.shape[1] != self.embedding_dim:
            raise ValueError(f"expected inputs [B, {self.embedding_dim}, W], got {tuple(inputs.shape)}")
        emb, com, quantized, ppl, idx = _VQFunction.apply(inputs, self.codebook.weight, float(self.commitment_cost),
                                                          self.precision, self.stats_comm, self.stats_sync)
        encodings = self._encodings(idx)
        return emb, com, quantized, ppl, encodings, idx.unsqueeze(1)

    @torch.no_grad()
    def encode(self, inputs: torch.Tensor) -> torch.Tensor:
        """Index export only (what Quantize.get_encodings_idx needs, transform.py:15-16): [N] int64, no `quantized`."""
        idx, _, _ = F.vq_forward(inputs.contiguous(), self.codebook.weight, precision=self.precision, want_q=False)
        return idx

    @torch.no_grad()
    def decode(self, idx: torch.Tensor, batch: int) -> torch.Tensor:
        """Indices -> codewords in BCW (the one-hot matmul of vector_quantizer.py:42 / bert.py:75-78 as a gather)."""
        idx = idx.reshape(-1)
        return F.gather(self.codebook.weight, idx, batch, idx.numel() // batch)   # raises IndexError on codes outside [0, K)
