"""Small fixed workload for ncu captures with a precision choice: python scripts/profile_prec.py B D W K iters precision"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import functional as F
B, D, W, K, iters = (int(a) for a in sys.argv[1:6])
prec = sys.argv[6] if len(sys.argv) > 6 else "bf16"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
z = torch.randn(B, D, W, device=dev, generator=g)
cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
for _ in range(iters):
    idx, q, st = F.vq_forward(z, cb, precision=prec, want_q=True, want_resid=True)
torch.cuda.synchronize()
print("ok", F.debug_counters())
