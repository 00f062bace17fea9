"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/vqb.h declares, validates its
arguments before touching a device, and fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest
import torch

import vq_b200
from vq_b200 import _lib, functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vqb.h")).read()
    return sorted(set(re.findall(r"VQB_API\s+(?:const\s+char\*|int|long long)\s+(vqb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 17
    assert sorted(_lib.EXPORTS) == names, "ctypes table and include/vqb.h disagree"
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in vqb.h but not exported by libvqb_b200.so"


def test_version_and_workspace_sizing():
    lib = _lib.lib()
    assert lib.vqb_version() == 200
    small = F.workspace_bytes(1000, 512, 64, _lib.PREC_FP32)
    big = F.workspace_bytes(1 << 20, 1024, 64, _lib.PREC_BF16)
    assert 0 < small < big
    # bf16 path: bf16 latent copy (2 B/elem) + ~50 B/frame of shortlist state + a fixed ~130 MB event scratch; never anything like N*K
    assert big < (1 << 20) * (64 * 2 + 64) + (200 << 20)


@pytest.mark.parametrize("N,K,D", [(100, 0, 64), (100, 70000, 64), (100, 512, 24), (100, 512, 1024), (0, 512, 64)])
def test_argument_errors_are_reported_not_swallowed(N, K, D):
    with pytest.raises(_lib.VqbError) as ei:
        F.workspace_bytes(N, K, D, _lib.PREC_BF16)
    assert ei.value.code == -2 and "unsupported" in str(ei.value)


def test_unknown_precision_flag():
    with pytest.raises(_lib.VqbError) as ei:
        F.workspace_bytes(100, 512, 64, 0x07)
    assert ei.value.code == -6


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    vq = vq_b200.VectorQuantizer(16, 16, 0.25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(torch.zeros(1, 16, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.vq_forward(torch.zeros(1, 16, 8), torch.zeros(16, 16))
    # calling the C ABI directly without a device is an error code, never a silent CPU result
    z = torch.zeros(1, 16, 8)
    cb = torch.zeros(16, 16)
    idx = torch.zeros(8, dtype=torch.int64)
    stats = torch.zeros(_lib.stats_len(16, 16))
    ws = torch.zeros(1 << 16, dtype=torch.uint8)
    rc = _lib.lib().vqb_forward(z.data_ptr(), cb.data_ptr(), 1, 16, 8, 16, 0, idx.data_ptr(), None, stats.data_ptr(), ws.data_ptr(),
                                ws.numel(), None)
    assert rc != 0 and _lib.lib().vqb_last_error()


def test_module_interface_matches_reference():
    """ctor / attributes / state_dict key of vector_quantizer.py:11-21 and the init range U(-1/K, 1/K)."""
    vq = vq_b200.VectorQuantizer(num_embedding=128, embedding_dim=32, commitment_cost=0.25)
    assert (vq.num_embedding, vq.embedding_dim, vq.commitment_cost) == (128, 32, 0.25)
    assert isinstance(vq.codebook, torch.nn.Embedding) and list(vq.state_dict()) == ["codebook.weight"]
    w = vq.codebook.weight
    assert w.shape == (128, 32) and w.requires_grad and float(w.abs().max()) <= 1 / 128


def test_experiment_switches_need_the_opt_in(monkeypatch):
    """A stray VQB_* variable must not change production behaviour: without VQB_EXPERIMENTS=1 the switches are ignored
    (here: VQB_TAIL_TMA=0 would select the register-staged tail, which needs no TMA-fed launch... observable through the
    workspace plan of VQB_TC_FUSE=0, which reserves the bf16 latent copy)."""
    flags = _lib.PREC_BF16
    base = F.workspace_bytes_bw(4, 64, 4096, 512, flags)
    monkeypatch.setenv("VQB_TC_FUSE", "0")
    _lib.lib().vqb_debug_reload_env()
    assert F.workspace_bytes_bw(4, 64, 4096, 512, flags) == base, "switch honoured without VQB_EXPERIMENTS=1"
    monkeypatch.setenv("VQB_EXPERIMENTS", "1")
    _lib.lib().vqb_debug_reload_env()
    assert F.workspace_bytes_bw(4, 64, 4096, 512, flags) == base + 4 * 4096 * 64 * 2
    monkeypatch.delenv("VQB_EXPERIMENTS")
    monkeypatch.delenv("VQB_TC_FUSE")
    _lib.lib().vqb_debug_reload_env()
    assert F.workspace_bytes_bw(4, 64, 4096, 512, flags) == base


def test_exact_workspace_drops_the_bf16_latent_copy():
    """the exact per-shape size omits the bf16 latent copy when the tensor-core kernel reads the fp32 latents
    itself (8.6 GB at BASELINE config 3) and sizes the event scratch by the grid; the N-only query stays an upper bound."""
    flags = _lib.PREC_BF16 | _lib.WANT_Q | _lib.WANT_RESID
    N, K, D = 1 << 24, 8192, 256
    exact = F.workspace_bytes_bw(1024, D, 16384, K, flags)
    upper = F.workspace_bytes(N, K, D, flags)
    assert upper - exact >= N * D * 2 and exact < 1.2e9, (exact, upper)
    ragged = F.workspace_bytes_bw(4, 64, 333, 512, _lib.PREC_BF16)           # W % 4 != 0: unfused, keeps the copy
    assert ragged <= F.workspace_bytes(4 * 333, 512, 64, _lib.PREC_BF16)
    tiny = F.workspace_bytes_bw(1, 64, 256, 512, _lib.PREC_BF16)             # two frame tiles: event scratch for two CTAs only
    assert tiny < (4 << 20), tiny


def test_tail3_permutation_is_the_layout_its_kernel_reads():
    """tail3_kernel gathers from a permuted copy of the codebook (and accumulates the residual sums in the same order): inside a row,
    lane sl of a frame's lpf lanes owns the dims whose d % 8 lies in [S sl, S sl + S), S = 8 / lpf, and its m-th dim sits at
    (m / 4) * 4 lpf + 4 sl + m % 4.  Host-only check of the C++ function codebook_prep_kernel and fold_resid_perm_kernel share:
    a bijection of [0, D), every aligned group of four positions belongs to ONE lane, and for a fixed float4 index the lanes of a
    frame are 16 bytes apart (one contiguous line of 16 lpf bytes)."""
    lib = _lib.lib()
    for D in (32, 64, 96, 128, 192, 256):
        assert lib.vqb_debug_tail3_lanes(D) in (4, 8)
        for lpf in (2, 4, 8):
            if D // lpf > 32 or (D // lpf) % 4:
                continue
            S = 8 // lpf
            pos = [lib.vqb_debug_tail3_perm_pos(d, lpf) for d in range(D)]
            assert sorted(pos) == list(range(D))
            owner = {p: (d % 8) // S for d, p in enumerate(pos)}
            for p0 in range(0, D, 4):
                assert len({owner[p0 + j] for j in range(4)}) == 1
                assert owner[p0] == (p0 // 4) % lpf                      # lanes interleave float4 by float4
            for d, p in enumerate(pos):                                  # the closed form of the header comment
                sl, m = (d % 8) // S, (d // 8) * S + (d % 8) % S
                assert p == (m // 4) * 4 * lpf + 4 * sl + m % 4
    assert lib.vqb_debug_tail3_perm_pos(-1, 8) < 0 and lib.vqb_debug_tail3_perm_pos(3, 3) < 0 and lib.vqb_debug_tail3_lanes(48) < 0
