"""Like profile_train.py but with clustered ("trained-like") latents: x = codeword + 0.1 * noise, so shortlists are single codes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200
from vq_b200 import functional as F
B, D, W, K, iters = 64, 256, 16384, 8192, 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
cb = torch.randn(K, D, device=dev, generator=g)
pick = torch.randint(0, K, (B * W,), device=dev, generator=g)
z = (cb[pick] + 0.1 * torch.randn(B * W, D, device=dev, generator=g)).reshape(B, W, D).permute(0, 2, 1).contiguous()
for _ in range(iters):
    idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True)
torch.cuda.synchronize()
print("ok", F.debug_counters(), bool((idx == pick).all()))
