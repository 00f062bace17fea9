"""Small fixed workload for ncu captures of the non-tensor kernels: forward + backward at a BASELINE shape class.
usage: python scripts/profile_train.py [B] [D] [W] [K] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200
from vq_b200 import functional as F

B, D, W, K, iters = (int(a) for a in (sys.argv[1:6] + ["64", "256", "16384", "8192", "3"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
z = torch.randn(B, D, W, device=dev, generator=g)
cb = torch.randn(K, D, device=dev, generator=g)
Gq = torch.randn(B, D, W, device=dev, generator=g)
one = torch.ones((), device=dev)
for _ in range(iters):
    idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True)
    dX, dE = F.vq_backward(z, cb, idx, st, Gq, one, one, 0.25)
torch.cuda.synchronize()
print("ok", F.debug_counters())
