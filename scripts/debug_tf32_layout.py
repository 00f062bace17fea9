"""Diagnostic for the kind::tf32 tile: one-hot latents reveal which (frame, dim) of the A operand and which (code, dim) of the B
operand each tensor-core product actually used.  usage: python scripts/debug_tf32_layout.py [D] [W] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import functional as F
D, W, K = (int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((1, 64), (2, 128), (3, 256)))
g = torch.Generator().manual_seed(0)
cb = torch.randn(K, D, generator=g).cuda()
e2 = (cb.double() ** 2).sum(1)
for f0, d0 in ((5, 3), (37, 3), (5, 11), (5, 35), (70, 20), (127, 63), (0, 0), (1, 0), (4, 0), (0, 1), (0, 8)):
    if f0 >= W or d0 >= D:
        continue
    z = torch.zeros(1, D, W).cuda()
    z[0, d0, f0] = 1.0
    got = F.debug_tc_scores(z, cb, precision=os.environ.get("PREC", "tf32")).double()
    dot = (e2[None, :] - got) / 2                      # [N, K]
    rows = torch.nonzero(dot.abs().max(1).values > 1e-3).reshape(-1).tolist()
    desc = []
    for r in rows[:6]:
        # which codebook column does this row reproduce?
        corr = (dot[r][:, None] - cb.double()).abs().max(0).values      # [D]: error vs cb[:, d]
        dbest = int(corr.argmin())
        desc.append(f"row {r}: matches cb[:, {dbest}] (err {corr[dbest]:.2e}), |dot|max {dot[r].abs().max():.3f}")
    print(f"one-hot x[f={f0}, d={d0}] -> nonzero frames {rows[:12]}{'...' if len(rows) > 12 else ''}; " + "; ".join(desc), flush=True)
