"""Torch-CPU port of the reference quantiser forward, used ONLY as the timed CPU baseline.  TEST INFRASTRUCTURE.

The reference (src/model/components/vector_quantizer.py:23-54) is a torch module that materialises the N x K distance
matrix and the N x K one-hot and gathers codewords with a dense one-hot matmul.  It cannot travel to the GPU box and it
cannot run un-chunked at BASELINE config 3 (N x K fp32 = 550 GB), so this port executes the same torch operations with
the same cost structure (distance sgemm, argmin, one-hot scatter, one-hot sgemm gather, two MSEs, histogram) on chunks of
frames and recombines losses / perplexity from sums and counts.  Precision "highest" (torch default), all host threads.
It is pinned against the golden fixtures in tests/test_oracle_golden.py like the numpy oracle.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def vq_forward_chunked(z_bcw: torch.Tensor, weight: torch.Tensor, beta: float, chunk: int = 32768):
    B, D, W = z_bcw.shape
    K = weight.shape[0]
    flat = z_bcw.permute(0, 2, 1).contiguous().view(-1, D)            # BCW -> BWC -> [N, D]   (:25-29)
    N = flat.shape[0]
    w2 = torch.sum(weight ** 2, dim=1)
    idx = torch.empty(N, dtype=torch.int64)
    quant = torch.empty_like(flat)
    counts = torch.zeros(K)
    sse = 0.0
    for s in range(0, N, chunk):
        x = flat[s:s + chunk]
        dist = torch.sum(x ** 2, dim=1, keepdim=True) + (w2 - 2 * torch.matmul(x, weight.t()))   # (:32-33)
        i = torch.argmin(dist, dim=1)                                                           # (:37)
        onehot = torch.zeros(x.shape[0], K)
        onehot.scatter_(1, i.unsqueeze(1), 1)                                                   # (:38-39)
        q = torch.matmul(onehot, weight)                                                        # (:42)
        sse += float(torch.sum((q - x) ** 2, dtype=torch.float64))                              # (:45-46)
        quant[s:s + chunk] = x + (q - x)                                                        # (:48)
        counts += onehot.sum(0)                                                                 # (:49)
        idx[s:s + chunk] = i
    mse = sse / (N * D)
    p = counts / N
    ppl = torch.exp(-torch.sum(p * torch.log(p + 1e-10)))                                       # (:50)
    out = quant.view(B, W, D).permute(0, 2, 1).contiguous()                                     # (:52)
    return mse, beta * mse, out, float(ppl), idx
