"""Smallest end-to-end forward of the default path (for compute-sanitizer / debugging): python scripts/tiny_forward.py [D] [W] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import functional as F
D, W, K = (int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((1, 64), (2, 256), (3, 128)))
g = torch.Generator().manual_seed(0)
z = torch.randn(2, D, W, generator=g).cuda()
cb = torch.randn(K, D, generator=g).cuda()
idx, q, st = F.vq_forward(z, cb, precision=os.environ.get("PREC", "bf16"), want_q=True, want_resid=True)
torch.cuda.synchronize()
rows = z.permute(0, 2, 1).reshape(-1, D)
d = (rows ** 2).sum(1, keepdim=True) + ((cb ** 2).sum(1) - 2 * rows @ cb.t())
print("ok", int((idx != d.argmin(1)).sum()), "mismatches of", idx.numel(), "sum counts", float(st[:K].sum()))
