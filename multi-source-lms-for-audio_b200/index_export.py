"""Index-stream export for the Transformer / BERT stages.

Mirrors `Quantize` (src/data/transform.py:5-16) and the window preparation inside `AudioBert.forward`
(src/model/bert.py:46-69): indices per clip, cut into 512-token windows, the last one zero-padded and masked out, plus
the optional 15 % [MASK] substitution the reference applies in training."""
from __future__ import annotations

from typing import Tuple

import torch

from . import functional as F


class Quantize:
    """Drop-in for transform.Quantize: wraps a frozen VQ-VAE-like object exposing `get_quantized(x)`."""

    def __init__(self, vqvae):
        self.vqvae = vqvae
        self.vqvae.eval()

    def get_quantized(self, x):
        return self.vqvae.get_quantized(x)[0]

    def get_encodings_idx(self, x):
        return self.vqvae.get_quantized(x)[2]


def export_windows(idx: torch.Tensor, batch: int, window: int = 512, pad_id: int = 0, mask_token: int | None = None,
                   mask_prob: float = 0.15, generator: torch.Generator | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx [B*L] or [B*L, 1] int64 (CUDA) -> (tokens [B, n_win, window] int64, attention_mask [B, n_win, window] fp32).

    With `mask_token` set, a random `mask_prob` of the tokens is replaced first, as bert.py:46-48 does in training."""
    idx = idx.reshape(-1)
    if mask_token is not None:
        drop = torch.rand(idx.numel(), device=idx.device, generator=generator) < mask_prob
        idx = torch.where(drop, torch.full_like(idx, mask_token), idx)
    return F.window_indices(idx, batch, window=window, pad_id=pad_id)
