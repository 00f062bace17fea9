"""Caller of the hot path: the VQ-VAE forward / training step around the bottleneck.

This is NOT a rebuild of the reference's LightningModule: it restates only the wiring of `VQVAE.__init__` / `forward`
(src/model/vqvae.py:39-53, 81-86), the stage-1 loss (vqvae.py:59-66) and the optimiser (vqvae.py:168-171) so that the
fused quantiser can be exercised and timed in its real context (BASELINE configs 1 and 4).  The convolutions are stock
cuDNN `nn.Conv1d` / `nn.ConvTranspose1d` (out of scope as kernels).  Sub-module and parameter names equal the
reference's, so a reference checkpoint's `state_dict` loads with `strict=True`.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .quantizer import VectorQuantizer


class ResidualStack(nn.Module):
    """x + conv1x1(relu(conv3(relu(x)))) repeated, final relu (components/residual_stack.py:5-26)."""

    def __init__(self, in_channel: int, num_hidden: int, num_residual_layer: int, num_residual_hidden: int):
        super().__init__()
        blocks = []
        for i in range(num_residual_layer):
            cin = in_channel if i == 0 else num_hidden
            blocks.append(nn.Sequential(nn.ReLU(True), nn.Conv1d(cin, num_residual_hidden, 3, 1, 1, bias=False), nn.ReLU(True),
                                        nn.Conv1d(num_residual_hidden, num_hidden, 1, 1, bias=False)))
        self.residual_layers = nn.ModuleList(blocks)

    def forward(self, x):
        for block in self.residual_layers:
            x = x + block(x)
        return F.relu(x)


class Encoder(nn.Module):
    """4 stems -> num_hidden channels at T/4 frames (components/encoder.py:7-29)."""

    def __init__(self, in_channel: int, num_hidden: int, num_residual_layer: int, num_residual_hidden: int):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channel, num_hidden // 2, 4, 2, 1)
        self.conv2 = nn.Conv1d(num_hidden // 2, num_hidden, 4, 2, 1)
        self.conv3 = nn.Conv1d(num_hidden, num_hidden, 3, 1, 1)
        self.residual_stack = ResidualStack(num_hidden, num_hidden, num_residual_layer, num_residual_hidden)

    def forward(self, x):
        return self.residual_stack(self.conv3(F.relu(self.conv2(F.relu(self.conv1(x))))))


class Decoder(nn.Module):
    """embedding_dim channels at T/4 frames -> 4 stems at T samples (components/decoder.py:7-33)."""

    def __init__(self, in_channel: int, num_hidden: int, num_residual_layer: int, num_residual_hidden: int):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channel, num_hidden, 3, 1, 1)
        self.residual_stack = ResidualStack(num_hidden, num_hidden, num_residual_layer, num_residual_hidden)
        self.conv1_transpose = nn.ConvTranspose1d(num_hidden, num_hidden // 2, 4, 2, 1)
        self.conv2_transpose = nn.ConvTranspose1d(num_hidden // 2, 4, 4, 2, 1)

    def forward(self, x):
        return self.conv2_transpose(F.relu(self.conv1_transpose(self.residual_stack(self.conv1(x)))))


class VQVAEStep(nn.Module):
    """encoder -> 1x1 conv -> B200 VectorQuantizer -> decoder, defaults of configs/model/vqvae.yaml:3-9."""

    def __init__(self, num_hidden: int = 128, num_residual_layer: int = 2, num_residual_hidden: int = 32, num_embedding: int = 512,
                 embedding_dim: int = 64, commitment_cost: float = 0.25, learning_rate: float = 1e-4, **vq_kwargs):
        super().__init__()
        self.learning_rate = learning_rate
        self.encoder = Encoder(4, num_hidden, num_residual_layer, num_residual_hidden)
        self.conv = nn.Conv1d(num_hidden, embedding_dim, 1, 1)
        self.vector_quantizer = VectorQuantizer(num_embedding, embedding_dim, commitment_cost, **vq_kwargs)
        self.decoder = Decoder(embedding_dim, num_hidden, num_residual_layer, num_residual_hidden)

    def forward(self, x):
        z = self.conv(self.encoder(x))
        embedding_loss, commitment_loss, quantized, perplexity, _, _ = self.vector_quantizer(z)
        return self.decoder(quantized), embedding_loss, commitment_loss, perplexity

    @torch.no_grad()
    def get_quantized(self, x):
        """vqvae.py:88-93: (quantized, encodings, encodings_idx)."""
        z = self.conv(self.encoder(x))
        _, _, quantized, _, encodings, idx = self.vector_quantizer(z)
        return quantized, encodings, idx

    @staticmethod
    def make_batch(instruments: torch.Tensor):
        """Stage-1 batch (the evident intent of datamodule.py:118-119): the mixture replicated on the 4 input channels,
        the 4 stems as targets."""
        mixture = instruments.sum(dim=1, keepdim=True).expand(-1, 4, -1).contiguous()
        return mixture, instruments

    def training_loss(self, batch):
        """embedding + commitment + sum of per-stem L1 (vqvae.py:59-66)."""
        mixed, instruments = batch
        output, embedding_loss, commitment_loss, perplexity = self(mixed)
        loss = embedding_loss + commitment_loss
        for i in range(4):
            loss = loss + F.l1_loss(output[:, i, :], instruments[:, i, :])
        return loss, perplexity

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.learning_rate, amsgrad=False)     # vqvae.py:168-171
