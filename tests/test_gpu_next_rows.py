"""GPU tests of the rows next to the hot path: VQ-VAE step harness against a reference-generated fixture (f1, BASELINE
config 1 in miniature), index export (f2) and the EMA extension (f4)."""
import os

import numpy as np
import pytest
import torch

import vq_b200
from oracle import vq_oracle as O
from vq_b200 import functional as F
from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vqvae_step_matches_reference_fixture(precision):
    """encoder -> 1x1 conv -> quantiser -> decoder + stage-1 loss (vqvae.py:59-66, 81-86) with the reference's weights.
    cuDNN vs MKL convolutions differ in the last bits, so latents are compared with a tolerance and indices on frames whose
    reference margin is clear of that noise."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN_DIR, "vqvae_step", "vqvae_step_b2_t4096.npz"))
    model = vq_b200.VQVAEStep(precision=precision).to(DEV)
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}, strict=True)
    instruments = torch.from_numpy(g["instruments"]).to(DEV)
    batch = model.make_batch(instruments)
    z = model.conv(model.encoder(batch[0]))
    zn = z.detach().cpu().numpy()
    np.testing.assert_allclose(zn, g["z"], rtol=1e-4, atol=1e-5)
    loss, ppl = model.training_loss(batch)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-4)
    quantized, enc, idx = model.get_quantized(batch[0])
    got = idx.reshape(-1).cpu().numpy()
    # a latent perturbation dz moves a squared distance by <= 2 |x - e| |dz|: frames whose reference margin exceeds twice that
    # (two codes move) must keep their index
    ref = O.vq_forward(g["z"], g["sd:vector_quantizer.codebook.weight"], 0.25)
    dz = np.sqrt(((zn - g["z"]) ** 2).sum(axis=1)).reshape(-1)                   # per-frame |dz| (frames are n = b*W + w)
    tol = 4.0 * (np.sqrt(ref.dmin + ref.margin) + dz) * dz + ref.eps
    clear = ref.margin > tol
    assert np.array_equal(ref.indices, g["indices"].astype(np.int64))
    assert clear.mean() > 0.9 and np.array_equal(got[clear], g["indices"][clear])
    assert (got == g["indices"]).mean() > 0.99
    np.testing.assert_allclose(ppl.item(), g["perplexity"], rtol=2e-2)
    gw = model.conv.weight.grad.cpu().numpy()
    np.testing.assert_allclose(gw, g["grad_conv_weight"], rtol=5e-3, atol=5e-3 * np.abs(g["grad_conv_weight"]).max())
    gc = model.vector_quantizer.codebook.weight.grad.cpu().numpy()
    np.testing.assert_allclose(gc, g["grad_codebook"], rtol=5e-3, atol=5e-3 * np.abs(g["grad_codebook"]).max())
    opt = model.configure_optimizers()
    opt.step()
    assert isinstance(opt, torch.optim.Adam) and opt.defaults["lr"] == 1e-4


def test_index_export_windows_and_masking():
    class Frozen:                                       # stands in for a frozen VQVAE: get_quantized -> (q, enc, idx)
        def __init__(self, vq): self.vq = vq
        def eval(self): return self
        def get_quantized(self, x):
            with torch.no_grad():
                _, _, q, _, enc, idx = self.vq(x)
            return q, enc, idx
    B, D, W, K = 3, 64, 1100, 512
    vq = vq_b200.VectorQuantizer(K, D, 0.25).to(DEV)
    x = torch.randn(B, D, W, device=DEV) * 0.002
    quant = vq_b200.Quantize(Frozen(vq))
    idx = quant.get_encodings_idx(x)                    # transform.py:15-16
    assert idx.shape == (B * W, 1) and torch.equal(idx.reshape(-1), vq.encode(x))
    assert quant.get_quantized(x).shape == (B, D, W)
    tok, mask = vq_b200.export_windows(idx, B)
    wt, wm = O.window_indices(idx.reshape(-1).cpu().numpy(), B, 512)
    assert np.array_equal(tok.cpu().numpy(), wt) and np.array_equal(mask.cpu().numpy(), wm)
    gen = torch.Generator(device=DEV).manual_seed(1)
    tok_m, _ = vq_b200.export_windows(idx, B, mask_token=103, generator=gen)     # bert.py:46-48
    real = torch.from_numpy(wm).bool().to(DEV)
    frac = float((tok_m[real] == 103).float().mean())
    assert 0.10 < frac < 0.20
    assert torch.equal(tok_m[real & (tok_m != 103)], tok[real & (tok_m != 103)])


def test_ema_update_matches_numpy():
    """EXTENSION (not in the reference): EMA codebook update derived from the statistics buffer."""
    B, D, W, K, decay, eps = 4, 32, 600, 128, 0.9, 1e-5
    g = torch.Generator(device=DEV).manual_seed(0)
    z = torch.randn(B, D, W, device=DEV, generator=g)
    cb = torch.randn(K, D, device=DEV, generator=g)
    idx, _, stats = F.vq_forward(z, cb, precision="fp32", want_q=False, want_resid=True)
    rows = z.permute(0, 2, 1).reshape(-1, D).cpu().numpy().astype(np.float64)
    ix = idx.cpu().numpy()
    counts = np.bincount(ix, minlength=K).astype(np.float64)
    sumx = np.zeros((K, D)); np.add.at(sumx, ix, rows)
    cs0 = np.ones(K); es0 = cb.cpu().numpy().astype(np.float64).copy()
    cs1 = decay * cs0 + (1 - decay) * counts
    es1 = decay * es0 + (1 - decay) * sumx
    n = cs1.sum()
    want = es1 / ((cs1 + eps) / (n + K * eps) * n)[:, None]
    cluster = torch.ones(K + 1, device=DEV); embed = cb.clone(); cb_new = cb.clone()
    F.ema_update(stats, cb_new, cluster, embed, decay=decay, eps=eps)
    np.testing.assert_allclose(cluster[:K].cpu().numpy(), cs1, rtol=1e-6)
    np.testing.assert_allclose(embed.cpu().numpy(), es1, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(cb_new.cpu().numpy(), want, rtol=1e-4, atol=1e-4)
