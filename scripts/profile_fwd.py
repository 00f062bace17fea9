"""Small fixed workload for ncu captures: a few forwards of the bottleneck at a BASELINE shape class.
usage: python scripts/profile_fwd.py [B] [D] [W] [K] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200
from vq_b200 import functional as F

B, D, W, K, iters = (int(a) for a in (sys.argv[1:6] + ["32", "256", "16384", "8192", "3"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
z = torch.randn(B, D, W, device=dev, generator=g)
cb = torch.randn(K, D, device=dev, generator=g)
for _ in range(iters):
    idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True)
torch.cuda.synchronize()
print("ok", F.debug_counters())
