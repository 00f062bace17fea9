"""Randomised shape stress for the CUDA path against the CPU oracle (manual run on a B200, not part of the test suite):
many small random (B, D, W, K) incl. ragged / tiny / non-multiple shapes, bf16 path, forward + backward.
usage: python scripts/stress_shapes.py [n_cases] [seed] [bf16|fp32]   (environment switches such as VQB_TC_TAIL=1 apply)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vq_b200
from vq_b200 import functional as F
from oracle import vq_oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda:0")
bad = 0
t0 = time.time()
for case in range(n_cases):
    D = int(rng.choice([16, 32, 48, 64, 96, 128, 192, 256, 320, 512]))
    K = int(rng.choice([1, 2, 7, 40, 255, 256, 257, 512, 700, 1024, 2050, 4096]))
    kind = rng.integers(0, 4)
    W = int({0: rng.integers(1, 40), 1: 4 * rng.integers(1, 80), 2: 128 * rng.integers(1, 12), 3: 1024 + rng.integers(0, 3000)}[int(kind)])
    B = int(rng.integers(1, 4))
    if B * W * K * 4 > 3e8:      # keep the oracle's N x K distances small
        K = max(1, int(3e8 / (B * W * 4)))
    z = rng.standard_normal((B, D, W), dtype=np.float32) * float(rng.choice([0.1, 1.0, 3.0]))
    cb = rng.standard_normal((K, D), dtype=np.float32) * float(rng.choice([0.05, 1.0]))
    if rng.random() < 0.3 and K > 4:   # a hot code
        hot = rng.random((B, W)) < 0.3
        z[np.nonzero(hot)[0], :, np.nonzero(hot)[1]] = cb[K // 2] + 0.01 * rng.standard_normal((int(hot.sum()), D), dtype=np.float32)
    ref = O.vq_forward(z, cb, 0.25)
    vq = vq_b200.VectorQuantizer(K, D, 0.25, precision=precision).to(dev)
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(cb))
    zt = torch.from_numpy(z).to(dev).requires_grad_(True)
    emb, com, q, ppl, enc, idx = vq(zt)
    (emb + com + q.sum() * 1e-3).backward()
    got = idx.reshape(-1).cpu().numpy()
    clear = ref.margin > ref.eps
    ok = np.array_equal(got[clear], ref.indices[clear])
    ok &= abs(emb.item() - ref.embedding_loss) <= 2e-5 * abs(ref.embedding_loss) + 1e-12
    dX, dE = O.vq_backward(z, cb, got, 0.25, 1.0, 1.0, np.full_like(z, 1e-3))
    ok &= np.allclose(zt.grad.cpu().numpy(), dX, rtol=1e-5, atol=1e-7 * np.abs(dX).max())
    ok &= np.allclose(vq.codebook.weight.grad.cpu().numpy(), dE, rtol=1e-4, atol=1e-6 * max(np.abs(dE).max(), 1e-30))
    if np.array_equal(got, ref.indices):
        ok &= np.array_equal(q.detach().cpu().numpy(), ref.quantized)
    # index export (no `quantized`, no residual sums) and the host-buffer entry point must give the same indices / statistics
    cbt = vq.codebook.weight.detach()
    idx2, _, st2 = F.vq_forward(zt.detach(), cbt, precision=precision, want_q=False, want_resid=False)
    ok &= bool(torch.equal(idx2, idx.reshape(-1)))
    if case % 4 == 0:
        zh = torch.from_numpy(z).pin_memory()
        idx_h = torch.empty(B * W, dtype=torch.int64).pin_memory()
        st_h = torch.empty(K * (D + 1) + 2, dtype=torch.float32).pin_memory()
        F.vq_forward_host(zh, torch.from_numpy(cb).pin_memory(), precision=precision, want_resid=True, idx_out=idx_h, stats_out=st_h)
        ok &= np.array_equal(idx_h.numpy(), got)
        ok &= np.allclose(st_h.numpy()[:K], st2.cpu().numpy()[:K])
        ok &= np.allclose(st_h.numpy()[K + K * D], st2.cpu().numpy()[K + K * D], rtol=1e-5)
    if not ok:
        bad += 1
        print("MISMATCH", dict(B=B, D=D, W=W, K=K), flush=True)
print(f"{n_cases - bad}/{n_cases} cases ok in {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
