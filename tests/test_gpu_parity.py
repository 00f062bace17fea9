"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) and through the drop-in module, against the
CPU oracle and the reference-generated golden fixtures.  Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest
import torch

import vq_b200
from oracle import vq_oracle as O
from vq_b200 import _lib, functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_RTOL = 1e-5
PRECISIONS = ["fp32", "bf16", "tf32"]


def seeded(seed, shape, scale=1.0):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * np.float32(scale)


def assert_index_parity(got, z, cb, ref_idx, margin, eps):
    """Equal on every frame whose oracle fp32 top-2 margin exceeds eps_n; on near-ties the chosen code must lie within
    eps_n of the minimum oracle distance (the stated near-tie tolerance)."""
    got = np.asarray(got).reshape(-1)
    clear = margin > eps
    n_clear_bad = int((got[clear] != ref_idx[clear]).sum())
    assert n_clear_bad == 0, f"{n_clear_bad} index mismatches on clear-margin frames"
    bad = np.nonzero(got != ref_idx)[0]
    if bad.size:
        d = O.distances(O.bcw_to_rows(z)[bad], cb)
        chosen = d[np.arange(bad.size), got[bad]]
        assert np.all(chosen - d.min(axis=1) <= eps[bad]), "near-tie frame resolved to a code outside the tolerance"
    return int(bad.size)


def run_module(z, cb, beta, precision, Gq=None, **kw):
    K, D = cb.shape
    vq = vq_b200.VectorQuantizer(K, D, beta, precision=precision, **kw).to(DEV)
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(cb))
    zt = torch.from_numpy(z).to(DEV).requires_grad_(True)
    out = vq(zt)
    if Gq is not None:
        emb, com, q = out[0], out[1], out[2]
        (emb + com + (q * torch.from_numpy(Gq).to(DEV)).sum()).backward()
    return vq, zt, out


# ----------------------------------------------------------------------------------------------- tensor-core tile
@pytest.mark.parametrize("B,D,W,K", [(1, 64, 128, 256), (2, 64, 1024, 512), (1, 256, 640, 8192), (3, 128, 1280, 300), (1, 80, 256, 1000),
                                     (1, 16, 2048, 64), (2, 192, 384, 700)])
def test_tcgen05_tf32_scores(B, D, W, K):
    """The kind::tf32 tile: the fp32 [dims][frames] boxes as an MN-major swizzled A operand, fp32 codebook boxes as B.  Raw scores
    equal |e|^2 - 2 tf32(x).tf32(e) in float64, where tf32() is what the tensor core keeps of an fp32 operand (truncation of the low
    13 mantissa bits, or round-to-nearest - the test accepts either and the guard band covers both)."""
    if W % 4 or not (W % 128 == 0 or W >= 1024) or D > 256:
        pytest.skip("shape takes the bf16 shortlist (tf32 reads the latents in place)")
    z = torch.from_numpy(seeded(11, (B, D, W))).to(DEV)
    cb = torch.from_numpy(seeded(12, (K, D))).to(DEV)
    got = F.debug_tc_scores(z, cb, precision="tf32")
    assert not torch.isnan(got).any(), "tile left scores unwritten"
    rows = z.permute(0, 2, 1).reshape(-1, D)

    def trunc(t):
        return (t.view(torch.int32) & -8192).view(torch.float32).double()

    def rne(t):
        i = t.view(torch.int32)
        return ((i + 0x0FFF + ((i >> 13) & 1)) & -8192).view(torch.float32).double()
    e2 = (cb.double() ** 2).sum(1)[None, :]
    errs = {}
    for name, f in (("truncate", trunc), ("round", rne)):
        want = e2 - 2.0 * f(rows) @ f(cb).T
        errs[name] = ((got.double() - want).abs().max().item(), want.abs().max().item())
    best = min(errs, key=lambda k: errs[k][0])
    err, scale = errs[best]
    assert err <= 2e-5 * scale + 1e-4, f"max |score error| {errs} (best: {best})"


@pytest.mark.parametrize("B,D,W,K", [(1, 64, 128, 256), (2, 64, 333, 512), (1, 256, 640, 8192), (3, 128, 100, 300),
                                     (1, 80, 257, 1000), (1, 512, 130, 512), (1, 16, 200, 64)])
def test_tcgen05_scores_match_bf16_matmul(B, D, W, K):
    """Validates TMA boxes, UMMA descriptors, TMEM addressing and the |e|^2 add: raw scores of the tcgen05 tile equal
    |e|^2 - 2 bf16(x).bf16(e) evaluated in float64 from the same bf16-rounded operands."""
    z = torch.from_numpy(seeded(11, (B, D, W))).to(DEV)
    cb = torch.from_numpy(seeded(12, (K, D))).to(DEV)
    got = F.debug_tc_scores(z, cb)
    rows = z.permute(0, 2, 1).reshape(-1, D)
    xb = rows.to(torch.bfloat16).to(torch.float64)
    eb = cb.to(torch.bfloat16).to(torch.float64)
    want = (cb.double() ** 2).sum(1)[None, :] - 2.0 * xb @ eb.T
    assert not torch.isnan(got).any(), "tile left scores unwritten"
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 2e-5 * scale + 1e-4, f"max |score error| {err} (scale {scale})"


# ----------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("precision", PRECISIONS)
def test_golden_forward_backward(golden, precision):
    g = golden
    z, cb, beta = g["z"], g["codebook"], float(g["beta"])
    K, D = cb.shape
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, precision, Gq=g["Gq"])
    assert [t.requires_grad for t in (emb, com, q, ppl, enc, idx)] == [True, True, True, False, False, False]
    assert idx.shape == (z.shape[0] * z.shape[2], 1) and idx.dtype == torch.int64 and enc.shape == (idx.shape[0], K)
    got = idx.reshape(-1).cpu().numpy()
    x2 = (O.bcw_to_rows(z) ** 2).sum(1)
    eps = O.near_tie_eps(x2, float((cb ** 2).sum(1).max()))
    n_bad = assert_index_parity(got, z, cb, g["indices"].astype(np.int64), g["margin"], eps)
    np.testing.assert_allclose(emb.item(), g["embedding_loss"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(com.item(), g["commitment_loss"], rtol=LOSS_RTOL)
    assert torch.equal(enc.argmax(1), idx.reshape(-1)) and float(enc.sum()) == idx.shape[0]
    if n_bad == 0:
        np.testing.assert_allclose(ppl.item(), g["perplexity"], rtol=LOSS_RTOL)
        if "quantized" in g:
            assert np.array_equal(q.detach().cpu().numpy(), g["quantized"]), "straight-through value must be bit-equal"
        np.testing.assert_allclose(zt.grad.cpu().numpy(), g["dX"], rtol=1e-5, atol=1e-7 * np.abs(g["dX"]).max())
        dE = vq.codebook.weight.grad.cpu().numpy()
        if "dE" in g:
            np.testing.assert_allclose(dE, g["dE"], rtol=1e-4, atol=1e-6 * np.abs(g["dE"]).max())
            untouched = np.bincount(got, minlength=K) == 0
            assert np.all(dE[untouched] == 0), "never-selected codes must get an exactly zero gradient"
            opt = torch.optim.Adam(vq.parameters(), lr=1e-4, amsgrad=False)     # vqvae.py:168-171
            opt.step()
            after = vq.codebook.weight.detach().cpu().numpy()
            np.testing.assert_allclose(after, g["codebook_after_adam"], rtol=1e-5, atol=2e-7)
            assert np.array_equal(after[untouched], cb[untouched]), "never-selected rows must stay bit-identical"
        else:
            sel = g["sel_codes"]
            np.testing.assert_allclose(dE[sel], g["dE_sel"], rtol=1e-4, atol=1e-6 * np.abs(g["dE_sel"]).max())
            rest = np.ones(K, bool); rest[sel] = False
            assert np.all(dE[rest] == 0)


# ----------------------------------------------------------------------------------------------- oracle on seeded inputs
CASES = [
    # B, D, W, K, codebook scale (None = reference default init U(-1/K, 1/K)), latent scale
    (2, 64, 11000, 512, None, 1.0),     # cfg-1 shape, tie-heavy default init
    (8, 64, 4099, 1024, 1.0, 1.0),      # W not a multiple of anything
    (1, 256, 3000, 8192, 1.0, 1.0),     # headline shape class
    (5, 128, 777, 384, 0.3, 2.0),       # K not a multiple of 256
    (1, 16, 50, 7, 1.0, 1.0),           # tiny everything
    (3, 512, 200, 1000, 1.0, 0.5),      # largest D
    (1, 64, 1, 512, 1.0, 1.0),          # a single frame
    (2, 48, 500, 1, 1.0, 1.0),          # a single code
    (3, 16, 4, 5, 1.0, 1.0),            # W % 4 == 0 but shorter than one TMA box of the tail (32 frames)
    (2, 32, 36, 64, 1.0, 1.0),          # one full box + a 4-frame remainder per batch item
    (1, 192, 2052, 600, 1.0, 1.0),      # fused operand preparation with a ragged last frame tile, D = 192 (J = 6)
    (1, 256, 256, 512, 1.0, 1.0),       # D = 256 with two frame tiles only: no CTA pairs, whole-tile codebook stages (found by stress_shapes.py)
    (2, 256, 128, 300, 1.0, 1.0),       # same, one tile per batch item
    (1, 512, 256, 40, 1.0, 1.0),        # D = 512 with two frame tiles: unfused operand preparation (found by stress_shapes.py)
    (3, 448, 128, 257, 1.0, 1.0),       # D = 448, three tiles: fused preparation without CTA pairs, two codebook stages
]


@pytest.mark.parametrize("tail_form", [None, "30"])
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B,D,W,K,cb_scale,z_scale", CASES)
def test_oracle_parity(B, D, W, K, cb_scale, z_scale, precision, tail_form, experiment_env):
    """tail_form "30": tail3_kernel wherever the shape allows it (by default small batches take tail2_kernel: too few tiles per block)."""
    if tail_form:
        if precision == "fp32" and D > 256:
            pytest.skip("same kernels as the default run")
        experiment_env(VQB_TAIL_FORM=tail_form)
    z = seeded(100 + D + K, (B, D, W), z_scale)
    if cb_scale is None:
        cb = np.random.default_rng(5).uniform(-1 / K, 1 / K, (K, D)).astype(np.float32)
    else:
        cb = seeded(200 + D + K, (K, D), cb_scale)
    beta = 0.25
    ref = O.vq_forward(z, cb, beta)
    Gq = seeded(7, z.shape, 1e-3)
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, precision, Gq=Gq)
    got = idx.reshape(-1).cpu().numpy()
    n_bad = assert_index_parity(got, z, cb, ref.indices, ref.margin, ref.eps)
    np.testing.assert_allclose(emb.item(), ref.embedding_loss, rtol=LOSS_RTOL)
    np.testing.assert_allclose(com.item(), ref.commitment_loss, rtol=LOSS_RTOL)
    if n_bad == 0:
        np.testing.assert_allclose(ppl.item(), ref.perplexity, rtol=LOSS_RTOL)
        assert np.array_equal(q.detach().cpu().numpy(), ref.quantized)
    dX, dE = O.vq_backward(z, cb, got, beta, 1.0, 1.0, Gq)
    np.testing.assert_allclose(zt.grad.cpu().numpy(), dX, rtol=1e-5, atol=1e-7 * np.abs(dX).max())
    np.testing.assert_allclose(vq.codebook.weight.grad.cpu().numpy(), dE, rtol=1e-4, atol=1e-6 * np.abs(dE).max())


@pytest.mark.parametrize("B,D,W,K,cb_scale", [(2, 64, 11000, 512, None), (1, 256, 3000, 8192, 1.0), (3, 128, 1280, 700, 1.0),
                                              (2, 16, 1100, 40, 1.0)])
def test_fused_tail_variant_matches_oracle(B, D, W, K, cb_scale, experiment_env):
    """VQB_TC_TAIL=1: the search kernel finishes the frames itself (rescoring, gather, straight-through value, statistics)
    and only the exact-search fallback frames go through the list kernel.  Same parity bar as the default path."""
    experiment_env(VQB_TC_TAIL="1")
    z = seeded(300 + D + K, (B, D, W))
    cb = (np.random.default_rng(5).uniform(-1 / K, 1 / K, (K, D)).astype(np.float32) if cb_scale is None
          else seeded(400 + D + K, (K, D), cb_scale))
    beta = 0.25
    ref = O.vq_forward(z, cb, beta)
    Gq = seeded(7, z.shape, 1e-3)
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, "bf16", Gq=Gq)
    got = idx.reshape(-1).cpu().numpy()
    n_bad = assert_index_parity(got, z, cb, ref.indices, ref.margin, ref.eps)
    np.testing.assert_allclose(emb.item(), ref.embedding_loss, rtol=LOSS_RTOL)
    if n_bad == 0:
        np.testing.assert_allclose(ppl.item(), ref.perplexity, rtol=LOSS_RTOL)
        assert np.array_equal(q.detach().cpu().numpy(), ref.quantized)
    dX, dE = O.vq_backward(z, cb, got, beta, 1.0, 1.0, Gq)
    np.testing.assert_allclose(zt.grad.cpu().numpy(), dX, rtol=1e-5, atol=1e-7 * np.abs(dX).max())
    np.testing.assert_allclose(vq.codebook.weight.grad.cpu().numpy(), dE, rtol=1e-4, atol=1e-6 * np.abs(dE).max())


@pytest.mark.parametrize("env", [{"VQB_TAIL_TMA": "0"}, {"VQB_TAIL_FORM": "0"}, {"VQB_TAIL_FORM": "0", "VQB_TAIL_VARIANT": "0"},
                                 {"VQB_TAIL_FORM": "2"}, {"VQB_TAIL_FORM": "216"}, {"VQB_TAIL_FORM": "30"}, {"VQB_TAIL_FORM": "300"}, {"VQB_TC_EPI": "1"}, {"VQB_RESID_REPLICAS": "1"},
                                 {"VQB_RESID_REPLICAS": "8"}, {"VQB_DX_TILES": "1"}, {"VQB_TC_ASLOTS": "6"},
                                 {"VQB_L2_ONCE": "1"}, {"VQB_TC_EHSLOTS": "5"}, {"VQB_TC_MODE": "1"}])
def test_kernel_variants_agree_with_the_default_path(env, experiment_env):
    """The experiment switches select other forms of the same kernels (register-staged tail, round-1 TMA tails, tail2_kernel
    with 32- and 16-frame tiles - the default is tail3_kernel -, one / eight
    residual-sum replicas, tile-staged backward, two spare A chunks, evict-first latent loads, a deeper bias-operand ring,
    cta_group::1 MMAs): indices and `quantized` must be identical, statistics and
    gradients equal up to the summation order of the atomics.  A hot code (a quarter of the frames) stresses the replicas."""
    B, D, W, K, beta = 3, 256, 1536, 700, 0.25
    cb = seeded(11, (K, D))
    z = seeded(12, (B, D, W))
    hot = np.random.default_rng(13).random((B, W)) < 0.25
    z[np.nonzero(hot)[0], :, np.nonzero(hot)[1]] = cb[5] + seeded(14, (int(hot.sum()), D), 0.05)
    Gq = seeded(7, z.shape, 1e-3)

    def run():
        vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, "bf16", Gq=Gq)
        return (idx.reshape(-1).cpu().numpy(), q.detach().cpu().numpy(), emb.item(), ppl.item(), zt.grad.cpu().numpy(),
                vq.codebook.weight.grad.cpu().numpy())
    ref = run()
    experiment_env(**env)
    got = run()
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert (got[0] == 5).mean() > 0.2
    np.testing.assert_allclose(got[2], ref[2], rtol=1e-6)
    np.testing.assert_allclose(got[3], ref[3], rtol=1e-6)
    np.testing.assert_allclose(got[4], ref[4], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got[5], ref[5], rtol=1e-4, atol=1e-6 * np.abs(ref[5]).max())


@pytest.mark.parametrize("D,lpf", [(64, 2), (64, 4), (64, 8), (32, 2), (128, 4), (128, 8), (96, 4)])
def test_tail3_lanes_per_frame_agree_with_the_oracle(D, lpf, experiment_env):
    """tail3_kernel reads the swizzled latent box in place with 8, 4 or 2 lanes per frame (the permutation of the codebook copy and
    of the residual replicas follows): every choice must give the oracle's indices, `quantized`, losses and gradients.  W is not a
    multiple of the 32-frame tile, so the last tile of every item is ragged."""
    experiment_env(VQB_TAIL_LPF=str(lpf), VQB_TAIL_FORM="30")
    B, W, K, beta = 3, 1100, 300, 0.25
    cb = seeded(21, (K, D))
    z = seeded(22, (B, D, W))
    Gq = seeded(23, z.shape, 1e-3)
    ref = O.vq_forward(z, cb, beta)
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, "bf16", Gq=Gq)
    got = idx.reshape(-1).cpu().numpy()
    n_bad = assert_index_parity(got, z, cb, ref.indices, ref.margin, ref.eps)
    if n_bad == 0:
        assert np.array_equal(q.detach().cpu().numpy(), ref.quantized)
    np.testing.assert_allclose(emb.item(), ref.embedding_loss, rtol=1e-5)
    np.testing.assert_allclose(ppl.item(), ref.perplexity, rtol=1e-5)
    dX, dE = O.vq_backward(z, cb, got, beta, 1.0, 1.0, Gq)
    np.testing.assert_allclose(zt.grad.cpu().numpy(), dX, rtol=1e-5, atol=1e-7 * np.abs(dX).max())
    np.testing.assert_allclose(vq.codebook.weight.grad.cpu().numpy(), dE, rtol=1e-4, atol=1e-6 * np.abs(dE).max())


@pytest.mark.parametrize("env", [{}, {"VQB_TC_EPI": "1"}, {"VQB_TC_EVSM": "-2"}, {"VQB_TC_EVSM": "0"}])
@pytest.mark.parametrize("K", [512, 1024, 4096])
def test_adversarial_code_order_floods_the_event_stacks(K, env, experiment_env):
    """Codes ordered so that a frame's score keeps improving along the codebook sweep: nearly every 8-code chunk is a new running
    maximum, the per-thread event stacks run through all their levels (own shared-memory slots, the warp's overflow pool, global
    scratch) and many frames fall back to the exact search - the result must still be the oracle's.  Also with the slab-queue epilogue
    (per-warp queues with a global spill, a flooding lane must only lose itself), without the pool and without shared-memory slots."""
    if env:
        experiment_env(**env)
    B, D, W, beta = 2, 64, 2048, 0.25
    rng = np.random.default_rng(31)
    u = rng.standard_normal(D).astype(np.float32)
    u /= np.linalg.norm(u)
    # random codes SORTED by their projection on the direction u all flooding latents share: a frame's score improves along the sweep
    # (a new running maximum in nearly every chunk) while the final candidates stay few (the top projections are well separated)
    cb = seeded(30, (K, D))
    cb = np.ascontiguousarray(cb[np.argsort(cb @ u)])
    a = rng.uniform(3.0, 6.0, (B, 1, W)).astype(np.float32)
    z = (a * u[None, :, None] + 0.05 * rng.standard_normal((B, D, W))).astype(np.float32)
    z[1, :, ::2] = seeded(32, (D, W // 2))            # half of the second item: ordinary latents between flooding neighbours
    ref = O.vq_forward(z, cb, beta)
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, "bf16")
    got = idx.reshape(-1).cpu().numpy()
    n_bad = assert_index_parity(got, z, cb, ref.indices, ref.margin, ref.eps)
    np.testing.assert_allclose(emb.item(), ref.embedding_loss, rtol=LOSS_RTOL)
    if n_bad == 0:
        assert np.array_equal(q.detach().cpu().numpy(), ref.quantized)
        np.testing.assert_allclose(ppl.item(), ref.perplexity, rtol=LOSS_RTOL)


def test_stage_timing_reports_every_stage_of_the_forward():
    """vqb_debug_stage_time_ms (bench.py's roofline legs): one timed launch per stage and forward, durations positive."""
    import ctypes as C
    lib = _lib.lib()
    z = torch.from_numpy(seeded(21, (2, 64, 2048))).to(DEV)
    cb = torch.from_numpy(seeded(22, (512, 64))).to(DEV)
    F.vq_forward(z, cb, precision="bf16")
    torch.cuda.synchronize()
    lib.vqb_debug_kernel_timing(1)
    for _ in range(3):
        F.vq_forward(z, cb, precision="bf16")
    for stage in range(5):
        ms, n = C.c_double(0), C.c_int(0)
        _lib.check("vqb_debug_stage_time_ms", lib.vqb_debug_stage_time_ms(stage, C.byref(ms), C.byref(n)))
        assert n.value == 3 and ms.value > 0.0, (stage, n.value, ms.value)
    lib.vqb_debug_kernel_timing(0)


@pytest.mark.parametrize("precision,seed", [("bf16", 101), ("fp32", 102), ("tf32", 103)])
def test_random_shapes_against_oracle(precision, seed):
    """scripts/stress_shapes.py: 80 random (B, D, W, K) incl. ragged / tiny / large-D shapes, forward + backward + index export +
    host path against the oracle.  (It found the launch-plan bugs fixed in round 1: D = 256 / 512 with fewer than four frame tiles.)"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "stress_shapes.py"), "80", str(seed), precision],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_trained_like_latents_use_single_candidate_shortlists():
    """Clustered latents (codeword + noise): the shortlist is a single code for almost every frame and nothing falls back."""
    K, D = 2048, 128
    cb = seeded(1, (K, D))
    pick = np.random.default_rng(2).integers(0, K, 4 * 2048)
    rows = cb[pick] + seeded(3, (pick.size, D), 0.1)
    z = np.ascontiguousarray(rows.reshape(4, 2048, D).transpose(0, 2, 1))
    idx, _, _ = F.vq_forward(torch.from_numpy(z).to(DEV), torch.from_numpy(cb).to(DEV), precision="bf16", want_q=False)
    c = F.debug_counters()
    assert np.array_equal(idx.cpu().numpy(), pick)
    assert c["fallback"] == 0 and c["rescored"] < 0.01 * pick.size


# ----------------------------------------------------------------------------------------------- edge semantics
def test_nan_and_inf_rows_follow_torch_argmin_semantics():
    """vector_quantizer.py:37 inherits torch.argmin: a NaN distance wins; +inf rows give index 0."""
    B, D, W, K = 1, 32, 64, 300
    z = seeded(1, (B, D, W))
    cb = seeded(2, (K, D))
    z[0, 3, 5] = np.nan
    z[0, 0, 9] = np.inf
    d = O.distances(O.bcw_to_rows(z), cb)
    want = O.argmin_first(d)
    for precision in PRECISIONS:
        idx, _, _ = F.vq_forward(torch.from_numpy(z).to(DEV), torch.from_numpy(cb).to(DEV), precision=precision, want_q=False)
        got = idx.cpu().numpy()
        assert got[5] == want[5] == 0 and got[9] == want[9]
        clean = np.ones(W, bool); clean[[5, 9]] = False
        assert np.array_equal(got[clean], want[clean])
    cbn = cb.copy(); cbn[17, 4] = np.nan     # NaN codeword: every distance row has a NaN at 17 -> index 17 everywhere
    zc = seeded(3, (B, D, W))
    for precision in PRECISIONS:
        idx, _, _ = F.vq_forward(torch.from_numpy(zc).to(DEV), torch.from_numpy(cbn).to(DEV), precision=precision, want_q=False)
        assert (idx == 17).all()


def test_duplicate_codes_pick_lowest_index():
    K, D = 512, 64
    cb = seeded(4, (K, D)); cb[256:] = cb[:256]
    z = seeded(5, (2, D, 700))
    for precision in PRECISIONS:
        idx, _, _ = F.vq_forward(torch.from_numpy(z).to(DEV), torch.from_numpy(cb).to(DEV), precision=precision, want_q=False)
        assert int(idx.max()) < 256


def test_non_contiguous_input_no_grad_and_eval():
    vq = vq_b200.VectorQuantizer(256, 64, 0.25).to(DEV).eval()
    base = torch.randn(2, 500, 64, device=DEV)
    x = base.permute(0, 2, 1)                 # BCW view of BWC storage
    assert not x.is_contiguous()
    with torch.no_grad():
        a = vq(x)
        b = vq(x.contiguous())
    assert torch.equal(a[5], b[5]) and torch.equal(a[2], b[2]) and a[2].is_contiguous()
    assert not a[0].requires_grad and not a[2].requires_grad


def test_detect_anomaly_and_sparse_encodings():
    vq = vq_b200.VectorQuantizer(128, 32, 0.5, dense_encodings=False).to(DEV)
    x = torch.randn(2, 32, 300, device=DEV, requires_grad=True)
    with torch.autograd.detect_anomaly():          # configs/debug/default.yaml:26
        emb, com, q, ppl, enc, idx = vq(x)
        (emb + com + q.square().mean()).backward()
    assert enc.is_sparse and enc.shape == (600, 128)
    assert torch.equal(enc.to_dense().argmax(1), idx.reshape(-1))
    assert torch.isfinite(x.grad).all() and torch.isfinite(vq.codebook.weight.grad).all()


def test_cabi_argument_errors_on_device():
    z = torch.zeros(1, 24, 8, device=DEV)          # D % 16 != 0
    with pytest.raises(_lib.VqbError) as ei:
        F.vq_forward(z, torch.zeros(16, 24, device=DEV))
    assert ei.value.code == -2
    z = torch.zeros(1, 32, 8, device=DEV)
    ws = torch.empty(64, dtype=torch.uint8, device=DEV)
    with pytest.raises(_lib.VqbError) as ei:
        F.vq_forward(z, torch.zeros(16, 32, device=DEV), workspace=ws)
    assert ei.value.code == -4


# ----------------------------------------------------------------------------------------------- helpers around the path
def test_onehot_gather_window_match_oracle():
    K, D, B, Lq = 300, 48, 3, 1100
    cb = seeded(6, (K, D))
    idx = np.random.default_rng(7).integers(0, K, B * Lq)
    it = torch.from_numpy(idx).to(DEV)
    enc = F.onehot(it, K).cpu().numpy()
    want = np.zeros((B * Lq, K), np.float32); want[np.arange(B * Lq), idx] = 1
    assert np.array_equal(enc, want)
    deq = F.gather(torch.from_numpy(cb).to(DEV), it, B, Lq).cpu().numpy()
    assert np.array_equal(deq, O.rows_to_bcw(cb[idx], B, Lq))
    tok, mask = F.window_indices(it, B, window=512)
    wt, wm = O.window_indices(idx, B, 512)
    assert np.array_equal(tok.cpu().numpy(), wt) and np.array_equal(mask.cpu().numpy(), wm)


def test_out_of_range_indices_are_refused_and_never_dereferenced():
    """caller-supplied indices (BERT-predicted tokens) outside [0, K).  The wrappers raise IndexError (the
    reference's scatter_ / one-hot matmul raise a device assert); with validate=False the kernels stay memory-safe: the
    one-hot row is all-zero, the gathered codeword and the gradient of that frame are NaN, every other frame is untouched."""
    K, D, B, Lq = 50, 32, 2, 40
    cb = torch.from_numpy(seeded(6, (K, D))).to(DEV)
    idx = torch.from_numpy(np.random.default_rng(7).integers(0, K, B * Lq)).to(DEV)
    bad = idx.clone()
    bad[3], bad[57] = K, -1
    for fn in (lambda: F.onehot(bad, K), lambda: F.gather(cb, bad, B, Lq),
               lambda: vq_b200.VectorQuantizer(K, D, 0.25).to(DEV).decode(bad, B)):
        with pytest.raises(IndexError):
            fn()
    enc = F.onehot(bad, K, validate=False)
    assert float(enc[3].sum()) == 0 and float(enc[57].sum()) == 0 and float(enc.sum()) == B * Lq - 2
    deq = F.gather(cb, bad, B, Lq, validate=False)
    good = F.gather(cb, idx, B, Lq)
    rows, rows_good = deq.permute(0, 2, 1).reshape(-1, D), good.permute(0, 2, 1).reshape(-1, D)
    assert torch.isnan(rows[3]).all() and torch.isnan(rows[57]).all()
    keep = torch.ones(B * Lq, dtype=torch.bool, device=DEV); keep[3] = keep[57] = False
    assert torch.equal(rows[keep], rows_good[keep])
    z = torch.from_numpy(seeded(8, (B, D, Lq))).to(DEV)
    stats = torch.zeros(_lib.stats_len(K, D), device=DEV); stats[-1] = B * Lq
    one = torch.ones((), device=DEV)
    for W_ in (Lq,):
        dX, _ = F.vq_backward(z, cb, bad, stats, None, one, one, 0.25)
        dXg, _ = F.vq_backward(z, cb, idx, stats, None, one, one, 0.25)
        r, rg = dX.permute(0, 2, 1).reshape(-1, D), dXg.permute(0, 2, 1).reshape(-1, D)
        assert torch.isnan(r[3]).all() and torch.isnan(r[57]).all() and torch.equal(r[keep], rg[keep])


@pytest.mark.parametrize("want_q", [False, True])
def test_host_buffer_path_matches_device_path(want_q):
    B, D, W, K = 6, 64, 1500, 512
    z = torch.from_numpy(seeded(8, (B, D, W))).pin_memory()
    cb = torch.from_numpy(seeded(9, (K, D))).pin_memory()
    idx_d, q_d, stats_d = F.vq_forward(z.to(DEV), cb.to(DEV), precision="bf16", want_q=want_q, want_resid=True)
    out = F.vq_forward_host(z, cb, precision="bf16", want_resid=True, chunk_batches=4, want_q=want_q)   # 2 chunks: 4 + 2
    idx_h, stats_h = out[0], out[1]
    assert torch.equal(idx_h, idx_d.cpu())
    if want_q:
        assert torch.equal(out[2], q_d.cpu()), "straight-through output of the host-buffer path"
    sd, sh = stats_d.cpu().numpy(), stats_h.numpy()
    assert np.array_equal(sd[:K], sh[:K])
    np.testing.assert_allclose(sh[K:], sd[K:], rtol=1e-4, atol=1e-3)
    _lib.check("vqb_host_release", _lib.lib().vqb_host_release())


# ----------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("B,D,W,K", [(64, 64, 16384, 1024), (16, 256, 16384, 8192)])
def test_full_size_properties(B, D, W, K):
    """BASELINE.json config 2 (N = 2^20) and a 2^18-frame slice of config 3: size-independent properties.
    (a) bf16 shortlist path and exact fp32 path give identical indices; (b) histogram sums to N;
    (c) idempotence: quantising the gathered codewords returns the same indices with zero loss;
    (d) a random sample of frames agrees with the oracle."""
    g = torch.Generator(device=DEV).manual_seed(42)
    z = torch.randn(B, D, W, device=DEV, generator=g)
    cb = torch.randn(K, D, device=DEV, generator=g)
    N = B * W
    idx_b, _, st_b = F.vq_forward(z, cb, precision="bf16", want_q=False)
    counters = F.debug_counters()
    idx_f, _, st_f = F.vq_forward(z, cb, precision="fp32", want_q=False)
    # Both paths score in fp32 in the reference's op order but sum the dot product in different orders, so they may
    # resolve a near-tie differently: any differing frame must be a near-tie under the stated epsilon.
    diff = torch.nonzero(idx_b != idx_f).reshape(-1)
    assert diff.numel() <= max(2, N // 100000), f"{diff.numel()} frames differ between bf16-shortlist and fp32 paths"
    if diff.numel():
        rows_d = z.permute(0, 2, 1).reshape(N, D)[diff].cpu().numpy()
        cbn = cb.cpu().numpy()
        dd = O.distances(rows_d, cbn)
        eps_d = O.near_tie_eps((rows_d ** 2).sum(1), float((cbn ** 2).sum(1).max()))
        ar = np.arange(diff.numel())
        assert np.all(np.abs(dd[ar, idx_b[diff].cpu().numpy()] - dd[ar, idx_f[diff].cpu().numpy()]) <= eps_d)
    assert float(st_b[:K].sum()) == N and float(st_f[:K].sum()) == N
    assert counters["fallback"] < 0.01 * N, counters
    codes = F.gather(cb, idx_b, B, W)
    idx_2, q2, st_2 = F.vq_forward(codes, cb, precision="bf16", want_q=True)
    assert torch.equal(idx_2, idx_b) and float(st_2[K + K * D]) == 0.0 and torch.equal(q2, codes)
    pick = np.random.default_rng(0).choice(N, 2048, replace=False)
    rows = z.permute(0, 2, 1).reshape(N, D)[torch.from_numpy(pick).to(DEV)].cpu().numpy()
    sub = np.ascontiguousarray(rows.T[None])           # [1, D, 2048]
    ref = O.vq_forward(sub, cb.cpu().numpy(), 0.25)
    assert_index_parity(idx_b.cpu().numpy()[pick], sub, cb.cpu().numpy(), ref.indices, ref.margin, ref.eps)


def _reference_indices_and_margins(z: torch.Tensor, cb: torch.Tensor, chunk: int = 32768):
    """Indices from the UNMODIFIED reference module run on the host cores in chunks (oracle/make_ref.py; the pinned torch port
    if no reference source is on this machine), plus the fp32 top-2 margins of the reference's own distance expression
    (vector_quantizer.py:32-33) and the stated near-tie tolerance."""
    from oracle import make_ref
    from oracle.ref_port_torch import vq_forward_chunked
    B, D, W = z.shape
    K = cb.shape[0]
    rows = z.permute(0, 2, 1).reshape(-1, D)
    N = rows.shape[0]
    VQ, where = make_ref.load_reference_class()
    idx = torch.empty(N, dtype=torch.int64)
    margin = torch.empty(N)
    w2 = torch.sum(cb ** 2, dim=1)
    vq = None
    if VQ is not None:
        vq = VQ(num_embedding=K, embedding_dim=D, commitment_cost=0.25)
        with torch.no_grad():
            vq.codebook.weight.copy_(cb)
    with torch.no_grad():
        for s0 in range(0, N, chunk):
            x = rows[s0:s0 + chunk]
            if vq is not None:                                        # [1, D, n] BCW item -> the module's own forward
                idx[s0:s0 + chunk] = vq(x.t().unsqueeze(0).contiguous())[5].reshape(-1)
            else:
                idx[s0:s0 + chunk] = vq_forward_chunked(x.t().unsqueeze(0).contiguous(), cb, 0.25, chunk=chunk)[4]
            d = torch.sum(x ** 2, dim=1, keepdim=True) + (w2 - 2 * torch.matmul(x, cb.t()))
            top2 = torch.topk(d, 2, dim=1, largest=False).values
            margin[s0:s0 + chunk] = top2[:, 1] - top2[:, 0]
    x2 = (rows ** 2).sum(1).numpy()
    eps = O.near_tie_eps(x2, float(w2.max()))
    return idx.numpy(), margin.numpy(), eps, where or "port"


@pytest.mark.parametrize("name,B,D,W,K,default_init", [("cfg3-data", 64, 256, 16384, 8192, False), ("cfg2-tie-heavy", 64, 64, 16384, 1024, True)])
def test_million_frames_against_the_reference(name, B, D, W, K, default_init):
    """VERDICT r01 'Next' #4: 2^20 frames of BASELINE config 3 data, and BASELINE config 2 with the reference's default-init
    codebook U(-1/K, 1/K) (tie-heavy), compared frame by frame with the reference quantiser run on the box's host cores."""
    g = torch.Generator().manual_seed(42)
    z = torch.randn(B, D, W, generator=g)
    cb = (torch.rand(K, D, generator=g) * 2 - 1) / K if default_init else torch.randn(K, D, generator=g)
    ref_idx, margin, eps, where = _reference_indices_and_margins(z, cb)
    zd, cbd = z.to(DEV), cb.to(DEV)
    for precision in PRECISIONS:
        idx, _, st = F.vq_forward(zd, cbd, precision=precision, want_q=False)
        got = idx.cpu().numpy()
        n_bad = assert_index_parity(got, z.numpy(), cb.numpy(), ref_idx, margin, eps)
        assert n_bad <= max(4, (margin <= eps).sum()), (name, precision, n_bad)
        assert float(st[:K].sum()) == B * W
    if default_init:
        assert int((margin == 0).sum()) > 100, "expected hundreds of exact fp32 ties in the default-init regime"


@pytest.mark.parametrize("precision", PRECISIONS)
def test_large_tie_heavy_reference_fixture(precision):
    """tests/golden/large: N = 176 000 default-init frames produced by the unmodified reference (77 exact ties, 46 frames whose
    fp32 and fp64 argmin differ)."""
    import hashlib
    from conftest import load_large_golden
    g = load_large_golden()
    z, cb, beta = g["z"], g["codebook"], float(g["beta"])
    vq, zt, (emb, com, q, ppl, enc, idx) = run_module(z, cb, beta, precision, Gq=None, dense_encodings=False)
    (emb + com).backward()
    got = idx.reshape(-1).cpu().numpy()
    x2 = (O.bcw_to_rows(z) ** 2).sum(1)
    eps = O.near_tie_eps(x2, float((cb ** 2).sum(1).max()))
    n_bad = assert_index_parity(got, z, cb, g["indices"].astype(np.int64), g["margin"], eps)
    np.testing.assert_allclose(emb.item(), g["embedding_loss"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(ppl.item(), g["perplexity"], rtol=1e-4 if n_bad else LOSS_RTOL)
    if n_bad == 0:
        assert hashlib.sha256(np.ascontiguousarray(q.detach().cpu().numpy()).tobytes()).hexdigest() == str(g["quantized_sha"])
    dE = vq.codebook.weight.grad.cpu().numpy()
    np.testing.assert_allclose(dE, g["dE"], rtol=1e-4, atol=1e-6 * np.abs(g["dE"]).max())


def test_forward_backward_are_cuda_graph_capturable():
    """The C ABI only enqueues work on the caller's stream (no host sync, no hidden streams): a forward + backward captured
    in a CUDA graph replays to the same results on new data."""
    B, D, W, K = 4, 64, 2048, 512
    g = torch.Generator(device=DEV).manual_seed(5)
    cb = torch.randn(K, D, device=DEV, generator=g)
    z = torch.randn(B, D, W, device=DEV, generator=g)
    Gq = torch.randn(B, D, W, device=DEV, generator=g)
    one = torch.ones((), device=DEV)
    stats = torch.empty(_lib.stats_len(K, D), device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                                   # warm-up on the capture stream (workspace + attribute setup)
            idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
            F.vq_backward(z, cb, idx, st, Gq, one, one, 0.25)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
        losses = F.vq_finalize(st, K, D, 0.25)
        dX, dE = F.vq_backward(z, cb, idx, st, Gq, one, one, 0.25)
    z2 = torch.randn(B, D, W, device=DEV, generator=g)
    z.copy_(z2)                                              # new data in the captured input buffer
    graph.replay()
    torch.cuda.synchronize()
    got = (idx.clone(), q.clone(), losses.clone(), dX.clone(), dE.clone())
    idx_e, q_e, st_e = F.vq_forward(z2, cb, precision="bf16", want_q=True, want_resid=True)
    losses_e = F.vq_finalize(st_e, K, D, 0.25)
    dX_e, dE_e = F.vq_backward(z2, cb, idx_e, st_e, Gq, one, one, 0.25)
    assert torch.equal(got[0], idx_e) and torch.equal(got[1], q_e)
    torch.testing.assert_close(got[2], losses_e, rtol=1e-6, atol=0)
    assert torch.equal(got[3], dX_e)
    torch.testing.assert_close(got[4], dE_e, rtol=1e-4, atol=1e-7)
