"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python oracle/make_golden.py

It imports `src.model.components.vector_quantizer.VectorQuantizer` from
/root/reference (torch default float32 matmul precision "highest", see
SURVEY.md section 8c), runs forward + autograd backward + one Adam step on
seeded inputs, and stores inputs and outputs as .npz.  The committed fixtures
are what pins `oracle/vq_oracle.py` and, through it, the CUDA path.

TEST INFRASTRUCTURE ONLY - nothing in the product package imports this.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present: golden fixtures can only be generated in the build container")
    sys.path.insert(0, REF)
    from src.model.components.vector_quantizer import VectorQuantizer  # noqa: E402
    return VectorQuantizer


def reference_margins(z: torch.Tensor, w: torch.Tensor):
    """Top-2 margin of the reference's own fp32 distance rows, evaluated with the
    very expression of vector_quantizer.py:32-33 (the module does not return it)."""
    x = torch.einsum("bcw -> bwc", z).contiguous().view(-1, w.shape[1])
    d = (torch.sum(x ** 2, dim=1, keepdim=True)) + (torch.sum(w ** 2, dim=1) - 2 * torch.matmul(x, w.t()))
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    return (top2[:, 1] - top2[:, 0]).numpy(), top2[:, 0].numpy()


def run_case(VQ, name: str, z: np.ndarray, codebook: np.ndarray | None, K: int, D: int, beta: float,
             store_inputs: bool = True, extra: dict | None = None):
    torch.manual_seed(1234)
    vq = VQ(num_embedding=K, embedding_dim=D, commitment_cost=beta)
    if codebook is not None:
        with torch.no_grad():
            vq.codebook.weight.copy_(torch.from_numpy(codebook))
    cb0 = vq.codebook.weight.detach().clone()
    zt = torch.from_numpy(z).clone().requires_grad_(True)
    emb, com, q, ppl, enc, idx = vq(zt)
    assert enc.shape == (z.shape[0] * z.shape[2], K) and idx.shape == (enc.shape[0], 1)
    g = torch.Generator().manual_seed(99)
    Gq = torch.randn(q.shape, generator=g) * 1e-3
    opt = torch.optim.Adam(vq.parameters(), lr=1e-4, amsgrad=False)   # vqvae.py:168-171
    loss = emb + com + (q * Gq).sum()
    loss.backward()
    dX = zt.grad.detach().clone()
    dE = vq.codebook.weight.grad.detach().clone()
    opt.step()
    margin, dmin = reference_margins(torch.from_numpy(z), cb0)
    out = dict(
        beta=np.float32(beta), K=np.int64(K), D=np.int64(D),
        embedding_loss=emb.detach().numpy(), commitment_loss=com.detach().numpy(),
        perplexity=ppl.detach().numpy(), indices=idx.detach().numpy().reshape(-1).astype(np.int32),
        margin=margin.astype(np.float32), dmin=dmin.astype(np.float32),
        Gq=Gq.numpy(), dX=dX.numpy(),
        requires_grad=np.array([t.requires_grad for t in (emb, com, q, ppl, enc, idx)]),
        torch_version=np.array(torch.__version__),
    )
    cb1 = vq.codebook.weight.detach()
    if store_inputs:
        out.update(z=z, codebook=cb0.numpy(), quantized=q.detach().numpy(), dE=dE.numpy(),
                   codebook_after_adam=cb1.numpy())
    else:
        # large-K case: inputs are regenerated from seeds on the box (sha256 guards generator drift); only the
        # rows of dE / the updated codebook that belong to selected codes are stored, the rest are asserted here.
        sel = torch.unique(idx.reshape(-1))
        rest = torch.ones(K, dtype=torch.bool)
        rest[sel] = False
        assert bool((dE[rest] == 0).all()) and bool((cb1[rest] == cb0[rest]).all())
        out.update(z_sha=np.array(_sha(z)), codebook_sha=np.array(_sha(cb0.numpy())),
                   quantized_sha=np.array(_sha(q.detach().numpy())), sel_codes=sel.numpy(),
                   dE_sel=dE[sel].numpy(), codebook_after_adam_sel=cb1[sel].numpy())
    if extra:
        out.update(extra)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    n_tie = int((margin == 0).sum())
    print(f"{name}: N={enc.shape[0]} K={K} D={D} emb={float(emb.detach()):.6g} ppl={float(ppl.detach()):.5g} exact_ties={n_tie} "
          f"-> {os.path.getsize(path)/1024:.0f} KiB")


def run_large_case(VQ, name: str, z: np.ndarray, K: int, D: int, beta: float):
    """cfg-1-sized, tie-heavy fixture (VERDICT r01 'Next' #4): default-init codebook U(-1/K, 1/K), N = 176 000 frames.
    Only what pins the index stream is stored: indices (int16), the reference's own fp32 top-2 margins, losses, perplexity, the
    codebook (torch RNG, so it cannot be regenerated from numpy) and its gradient; the latents come back from the numpy seed
    (sha256 guards generator drift)."""
    torch.manual_seed(1234)
    vq = VQ(num_embedding=K, embedding_dim=D, commitment_cost=beta)
    cb0 = vq.codebook.weight.detach().clone()
    zt = torch.from_numpy(z).clone().requires_grad_(True)
    emb, com, q, ppl, enc, idx = vq(zt)
    (emb + com).backward()
    margin, dmin = reference_margins(torch.from_numpy(z), cb0)
    x = torch.einsum("bcw -> bwc", torch.from_numpy(z)).contiguous().view(-1, D)
    d64 = (x.double() ** 2).sum(1, keepdim=True) + ((cb0.double() ** 2).sum(1) - 2 * x.double() @ cb0.double().t())
    idx64 = torch.argmin(d64, dim=1)
    out = dict(beta=np.float32(beta), K=np.int64(K), D=np.int64(D), shape=np.array(z.shape, dtype=np.int64),
               embedding_loss=emb.detach().numpy(), commitment_loss=com.detach().numpy(), perplexity=ppl.detach().numpy(),
               indices=idx.detach().numpy().reshape(-1).astype(np.int16), margin=margin.astype(np.float32),
               codebook=cb0.numpy(), dE=vq.codebook.weight.grad.detach().numpy(), z_sha=np.array(_sha(z)),
               quantized_sha=np.array(_sha(q.detach().numpy())), dX_sha=np.array(_sha(zt.grad.detach().numpy())),
               exact_ties=np.int64(int((margin == 0).sum())), fp32_ne_fp64=np.int64(int((idx64 != idx.reshape(-1)).sum())),
               torch_version=np.array(torch.__version__))
    os.makedirs(os.path.join(OUT, "large"), exist_ok=True)     # its own directory: a different schema than the small fixtures
    path = os.path.join(OUT, "large", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: N={enc.shape[0]} K={K} D={D} emb={float(emb.detach()):.6g} ppl={float(ppl.detach()):.5g} "
          f"exact_ties={int(out['exact_ties'])} fp32!=fp64={int(out['fp32_ne_fp64'])} -> {os.path.getsize(path)/1024:.0f} KiB")


def _sha(a: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded(seed: int, shape, scale: float = 1.0) -> np.ndarray:
    """Inputs come from numpy's PCG64 so the GPU box can regenerate the large ones."""
    return (np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * np.float32(scale))


def main():
    VQ = load_reference()
    torch.set_num_threads(8)
    # (i) default-init codebook U(-1/K, 1/K) + randn latents: the tie-heavy regime
    run_case(VQ, "default_init_k512_d64", seeded(1, (2, 64, 375)), None, 512, 64, 0.25)
    # (ii) randn codebook, W not a multiple of 128, B=3
    run_case(VQ, "randn_k1024_d64", seeded(2, (3, 64, 250)), seeded(3, (1024, 64)), 1024, 64, 0.25)
    # (iii) the reference's committed trained codebook, latents = codeword + noise
    import pandas as pd
    cb = pd.read_csv(os.path.join(REF, "logs/best_checkpoint/codebook.csv")).values.astype(np.float32)
    header = np.array(pd.read_csv(os.path.join(REF, "logs/best_checkpoint/codebook.csv"), header=None, nrows=1).values)
    assert cb.shape == (512, 64), cb.shape
    rng = np.random.default_rng(4)
    pick = rng.integers(0, 512, size=2 * 300)
    rows = cb[pick] + rng.standard_normal((600, 64), dtype=np.float32) * np.float32(0.05)
    z = np.ascontiguousarray(rows.reshape(2, 300, 64).transpose(0, 2, 1))
    run_case(VQ, "trained_codebook_csv", z, cb, 512, 64, 0.25, extra=dict(csv_header_row=header))
    # (iv) adversarial: duplicated codewords (lowest index must win), B=1, K=128, D=128
    cbd = seeded(5, (128, 128))
    cbd[64:] = cbd[:64]
    run_case(VQ, "duplicate_codes_k128_d128", seeded(6, (1, 128, 130)), cbd, 128, 128, 1.0)
    # (v) the headline shape class K=8192, D=256 on a small N; inputs regenerated from seeds on the box
    run_case(VQ, "randn_k8192_d256", seeded(7, (2, 256, 320)), seeded(8, (8192, 256)), 8192, 256, 0.25,
             store_inputs=False)
    # (vi) K=512 D=64 latents of stage-1 shape statistics (small norm codebook after a few steps), beta=0.5
    run_case(VQ, "small_norm_k256_d64", seeded(9, (4, 64, 129), 0.05), seeded(10, (256, 64), 0.02), 256, 64, 0.5)


def vqvae_case():
    """BASELINE config 1 in miniature: the reference's Encoder / 1x1 conv / VectorQuantizer / Decoder wired as
    src/model/vqvae.py:39-53,81-86 with the stage-1 loss of vqvae.py:59-66, seeded weights, batch 2."""
    sys.path.insert(0, REF)
    from src.model.components.encoder import Encoder
    from src.model.components.decoder import Decoder
    from src.model.components.vector_quantizer import VectorQuantizer
    from torch import nn
    import torch.nn.functional as F
    torch.manual_seed(2024)
    enc = Encoder(in_channel=4, num_hidden=128, num_residual_layer=2, num_residual_hidden=32)
    conv = nn.Conv1d(128, 64, kernel_size=1, stride=1)
    vq = VectorQuantizer(num_embedding=512, embedding_dim=64, commitment_cost=0.25)
    dec = Decoder(in_channel=64, num_hidden=128, num_residual_layer=2, num_residual_hidden=32)
    instruments = torch.from_numpy(seeded(11, (2, 4, 4096), 0.1))
    mixed = instruments.sum(dim=1, keepdim=True).expand(-1, 4, -1).contiguous()      # intent of datamodule.py:118-119
    with torch.no_grad():                          # codewords drawn from the latents themselves, so that many codes are in use
        z0 = conv(enc(mixed)).permute(0, 2, 1).reshape(-1, 64)
        pick = torch.randperm(z0.shape[0], generator=torch.Generator().manual_seed(5))[:512]
        vq.codebook.weight.copy_(z0[pick] + 0.002 * torch.randn(512, 64, generator=torch.Generator().manual_seed(6)))
    z = conv(enc(mixed))
    emb, com, q, ppl, _, idx = vq(z)
    out = dec(q)
    loss = emb + com
    for i in range(4):
        loss = loss + F.l1_loss(out[:, i, :], instruments[:, i, :])
    loss.backward()
    state = {}
    for prefix, mod in (("encoder.", enc), ("conv.", conv), ("vector_quantizer.", vq), ("decoder.", dec)):
        for k, v in mod.state_dict().items():
            state["sd:" + prefix + k] = v.detach().numpy()
    margin, _ = reference_margins(z.detach(), vq.codebook.weight.detach())
    path = os.path.join(OUT, "vqvae_step", "vqvae_step_b2_t4096.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, instruments=instruments.numpy(), z=z.detach().numpy(), output=out.detach().numpy(),
                        loss=loss.detach().numpy(), embedding_loss=emb.detach().numpy(), perplexity=ppl.detach().numpy(),
                        indices=idx.reshape(-1).numpy().astype(np.int32), margin=margin.astype(np.float32),
                        grad_conv_weight=conv.weight.grad.numpy(), grad_codebook=vq.codebook.weight.grad.numpy(), **state)
    print(f"vqvae_step: loss={float(loss.detach()):.6g} ppl={float(ppl):.4g} unique codes={idx.unique().numel()} "
          f"-> {os.path.getsize(path)/1024:.0f} KiB")


def large_cases():
    VQ = load_reference()
    torch.set_num_threads(8)
    # BASELINE config 1 at full clip length x 16 clips: N = 176 000, tie-heavy default-init codebook
    run_large_case(VQ, "large_default_init_k512_d64_n176000", seeded(21, (16, 64, 11000)), 512, 64, 0.25)


if __name__ == "__main__":
    if "--large-only" in sys.argv:
        large_cases()
    else:
        main()
        vqvae_case()
        large_cases()
