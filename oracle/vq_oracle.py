"""CPU oracle for the VQ-VAE vector-quantiser bottleneck.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference quantiser
(`/root/reference/src/model/components/vector_quantizer.py:23-54`) and of the
backward pass autograd derives from it.  It is the
*checker* for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product package never does, and has no CPU fallback.

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so this restatement is pinned against outputs of the
reference module itself, executed in the build container by
`oracle/make_golden.py` and committed under `tests/golden/` (see
`tests/test_oracle_golden.py`).

Everything is fp32 in the reference's association order.  The arithmetic lives
in numpy (BLAS sgemm for the contraction), chunked over frames so that the
N x K distance matrix the reference materialises never has to exist at once.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- layout
def bcw_to_rows(z_bcw: np.ndarray) -> np.ndarray:
    """[B, D, W] -> [N, D] with n = b*W + w  (vector_quantizer.py:25-29)."""
    B, D, W = z_bcw.shape
    return np.ascontiguousarray(np.transpose(z_bcw, (0, 2, 1))).reshape(B * W, D)


def rows_to_bcw(rows: np.ndarray, B: int, W: int) -> np.ndarray:
    """[N, D] -> [B, D, W] contiguous  (vector_quantizer.py:52)."""
    D = rows.shape[1]
    return np.ascontiguousarray(np.transpose(rows.reshape(B, W, D), (0, 2, 1)))


# --------------------------------------------------------------------------- distances
def code_sqnorms(codebook: np.ndarray) -> np.ndarray:
    """sum(weight**2, dim=1)  (vector_quantizer.py:33) in fp32."""
    e = codebook.astype(F32, copy=False)
    return np.sum(e * e, axis=1, dtype=F32)


def distances(rows: np.ndarray, codebook: np.ndarray, e2: np.ndarray | None = None) -> np.ndarray:
    """d[n,k] = fl(|x_n|^2 + fl(|e_k|^2 - fl(2 * fl(x_n . e_k))))  (vector_quantizer.py:32-33).

    The association order is part of the contract: adding |x|^2 last quantises
    the score to ulp(|x|^2 + ...) and creates exact fp32 ties that argmin
    resolves to the lowest index.
    """
    x = rows.astype(F32, copy=False)
    e = codebook.astype(F32, copy=False)
    if e2 is None:
        e2 = code_sqnorms(e)
    x2 = np.sum(x * x, axis=1, keepdims=True, dtype=F32)
    dot = x @ e.T  # fp32 sgemm
    two_dot = F32(2.0) * dot
    inner = e2[None, :] - two_dot
    return x2 + inner


def argmin_first(d: np.ndarray) -> np.ndarray:
    """torch.argmin(dim=1) semantics (vector_quantizer.py:37):
    lowest index on ties, and a NaN anywhere in the row wins (first NaN)."""
    idx = np.argmin(d, axis=1)  # numpy: first occurrence of the min, NaN propagates as min
    nan_rows = np.isnan(d).any(axis=1)
    if nan_rows.any():
        idx = idx.copy()
        idx[nan_rows] = np.argmax(np.isnan(d[nan_rows]), axis=1)
    return idx.astype(np.int64)


def ulp32(v: np.ndarray) -> np.ndarray:
    return np.spacing(np.abs(v).astype(F32)).astype(F32)


def near_tie_eps(x2: np.ndarray, e2max: float) -> np.ndarray:
    """The stated near-tie tolerance:
    eps_n = 8 * ulp_fp32(|x_n|^2 + max_k |e_k|^2)."""
    return F32(8.0) * ulp32(x2.astype(F32) + F32(e2max))


# --------------------------------------------------------------------------- forward
@dataclass
class VQForward:
    embedding_loss: np.float32
    commitment_loss: np.float32
    quantized: np.ndarray        # [B, D, W] fp32, the straight-through VALUE fl(x + fl(q - x))
    perplexity: np.float32
    indices: np.ndarray          # [N] int64  (reference returns [N, 1])
    counts: np.ndarray           # [K] int64 histogram of indices
    margin: np.ndarray           # [N] fp32 oracle top-2 margin d(2) - d(1) (0 on exact ties)
    eps: np.ndarray              # [N] fp32 near-tie tolerance for that frame
    dmin: np.ndarray             # [N] fp32 oracle minimum distance
    sse: float                   # sum over n,d of (q - x)^2 in float64

    def encodings(self, K: int) -> np.ndarray:
        """Dense one-hot [N, K] fp32 (vector_quantizer.py:38-39). Small cases only."""
        enc = np.zeros((self.indices.shape[0], K), dtype=F32)
        enc[np.arange(self.indices.shape[0]), self.indices] = 1.0
        return enc


def perplexity_from_counts(counts: np.ndarray, N: int) -> np.float32:
    """exp(-sum p log(p + 1e-10)), p = counts/N in fp32 (vector_quantizer.py:49-50)."""
    p = (counts.astype(F32) / F32(N)).astype(F32)
    t = p * np.log(p + F32(1e-10), dtype=F32)
    return F32(np.exp(-np.sum(t, dtype=F32), dtype=F32))


def vq_forward(z_bcw: np.ndarray, codebook: np.ndarray, beta: float, chunk: int = 16384) -> VQForward:
    """Reference forward (vector_quantizer.py:23-54) restated, chunked over frames."""
    z_bcw = np.asarray(z_bcw, dtype=F32)
    e = np.ascontiguousarray(codebook, dtype=F32)
    B, D, W = z_bcw.shape
    K = e.shape[0]
    N = B * W
    rows = bcw_to_rows(z_bcw)
    e2 = code_sqnorms(e)
    e2max = float(np.max(e2)) if np.isfinite(e2).all() else float("inf")

    idx = np.empty(N, dtype=np.int64)
    margin = np.empty(N, dtype=F32)
    dmin = np.empty(N, dtype=F32)
    for s in range(0, N, chunk):
        d = distances(rows[s:s + chunk], e, e2)
        i = argmin_first(d)
        idx[s:s + chunk] = i
        if K > 1:
            part = np.partition(d, 1, axis=1)[:, :2]
            dmin[s:s + chunk] = part[:, 0]
            margin[s:s + chunk] = part[:, 1] - part[:, 0]
        else:
            dmin[s:s + chunk] = d[:, 0]
            margin[s:s + chunk] = np.inf
    x2 = np.sum(rows * rows, axis=1, dtype=F32)
    eps = near_tie_eps(x2, e2max)

    q = e[idx]                                  # vector_quantizer.py:42 (exact gather at "highest")
    diff = (q - rows).astype(F32)
    sse = float(np.sum(diff.astype(np.float64) ** 2))
    mse = F32(sse / (N * D))                    # F.mse_loss (vector_quantizer.py:45-46)
    st = (rows + diff).astype(F32)              # vector_quantizer.py:48, value only
    counts = np.bincount(idx, minlength=K).astype(np.int64)
    return VQForward(
        embedding_loss=mse,
        commitment_loss=F32(F32(beta) * mse),
        quantized=rows_to_bcw(st, B, W),
        perplexity=perplexity_from_counts(counts, N),
        indices=idx, counts=counts, margin=margin, eps=eps, dmin=dmin, sse=sse,
    )


# --------------------------------------------------------------------------- backward
def vq_backward(z_bcw: np.ndarray, codebook: np.ndarray, indices: np.ndarray, beta: float,
                g_e: float, g_c: float, G_q: np.ndarray | None):
    """Gradients autograd derives from vector_quantizer.py:42-48:

      dX[b,:,w] = G_q[b,:,w] + g_c * beta * 2 (x - q) / (N D)
      dE[k,:]   = g_e * (2 / (N D)) * sum_{n: idx_n = k} (e_k - x_n)

    Accumulated in float64 and rounded once, so it is a tight reference for
    the fp32 kernels (tolerances in SURVEY.md section 8c)."""
    z_bcw = np.asarray(z_bcw, dtype=F32)
    e = np.asarray(codebook, dtype=F32)
    B, D, W = z_bcw.shape
    N = B * W
    rows = bcw_to_rows(z_bcw).astype(np.float64)
    q = e[indices].astype(np.float64)
    scale = 2.0 / (N * D)
    dX_rows = g_c * beta * scale * (rows - q)
    dX = rows_to_bcw(dX_rows.astype(F32), B, W)
    if G_q is not None:
        dX = (dX.astype(np.float64) + np.asarray(G_q, dtype=np.float64)).astype(F32)
    dE = np.zeros(e.shape, dtype=np.float64)
    np.add.at(dE, indices, (q - rows))
    dE *= g_e * scale
    return dX, dE.astype(F32)


def adam_step(param: np.ndarray, grad: np.ndarray, lr: float = 1e-4, b1: float = 0.9, b2: float = 0.999,
              eps: float = 1e-8) -> np.ndarray:
    """First Adam step from zero moments, as torch.optim.Adam(amsgrad=False)
    applies it to the codebook (vqvae.py:168-171)."""
    g = grad.astype(F32)
    m = (F32(1 - b1) * g).astype(F32)
    v = (F32(1 - b2) * g * g).astype(F32)
    mhat = m / F32(1 - b1)
    denom = (np.sqrt(v, dtype=F32) / F32(math.sqrt(1 - b2))) + F32(eps)
    return (param.astype(F32) - F32(lr) * (mhat / denom)).astype(F32)


# --------------------------------------------------------------------------- statistics (multi-GPU exchange)
def shard_stats(z_bcw: np.ndarray, codebook: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """The per-rank statistics buffer the CUDA path all-reduces:
    [counts[K] | sum_{n in k}(x_n - e_k) [K, D] | SSE | N] as float64 here."""
    e = np.asarray(codebook, dtype=F32)
    K, D = e.shape
    rows = bcw_to_rows(np.asarray(z_bcw, dtype=F32)).astype(np.float64)
    res = np.zeros((K, D), dtype=np.float64)
    np.add.at(res, indices, rows - e[indices].astype(np.float64))
    counts = np.bincount(indices, minlength=K).astype(np.float64)
    sse = float(np.sum((rows - e[indices].astype(np.float64)) ** 2))
    return np.concatenate([counts, res.reshape(-1), [sse, float(rows.shape[0])]])


def finalize_from_stats(stats: np.ndarray, K: int, D: int, beta: float):
    """Losses, perplexity and dE (for g_e = 1) from a summed statistics buffer."""
    counts = stats[:K]
    res = stats[K:K + K * D].reshape(K, D)
    sse, n = stats[K + K * D], stats[K + K * D + 1]
    mse = F32(sse / (n * D))
    ppl = perplexity_from_counts(counts, int(n))
    dE = (-(2.0 / (n * D)) * res).astype(F32)
    return mse, F32(F32(beta) * mse), ppl, dE


# --------------------------------------------------------------------------- index export (BERT windows)
def window_indices(indices: np.ndarray, batch: int, window: int = 512, pad_id: int = 0):
    """Index stream -> BERT windows, restating bert.py:50-69: reshape to
    [B, L], cut into `window`-token pieces, zero-pad the last one and mask the
    padding out.  Returns (tokens [B, n_win, window] int64, mask [B, n_win, window] fp32)."""
    x = np.asarray(indices).reshape(batch, -1)
    L = x.shape[1]
    n_win = (L + window - 1) // window
    tokens = np.full((batch, n_win, window), pad_id, dtype=np.int64)
    mask = np.zeros((batch, n_win, window), dtype=F32)
    for j in range(n_win):
        piece = x[:, j * window:(j + 1) * window]
        tokens[:, j, :piece.shape[1]] = piece
        mask[:, j, :piece.shape[1]] = 1.0
    return tokens, mask
