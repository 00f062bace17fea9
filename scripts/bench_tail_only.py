"""Tail-dominated step (tiny codebook: the search is ~3% of the step) repeated back to back: how fast is the tail kernel
under sustained load, without the tensor-core kernel in front of it?  usage: python scripts/bench_tail_only.py [K] [steps] [B] [D] [W]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200
from vq_b200 import _lib, functional as F

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, D, W = (int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((3, 1024), (4, 256), (5, 16384)))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
z = torch.randn(B, D, W, device=dev, generator=g)
if os.environ.get("BENCH_SEEDS"):      # bench.py draws the codebook from its own generator
    g = torch.Generator(device=dev).manual_seed(4242)
cb = torch.randn(K, D, device=dev, generator=g)
stats = torch.empty(_lib.stats_len(K, D), device=dev)
lib = _lib.lib()
for _ in range(3):
    F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
torch.cuda.synchronize()
lib.vqb_debug_kernel_timing(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
e1.record()
torch.cuda.synchronize()
out = {"K": K, "D": D, "N": B * W, "step_ms": e0.elapsed_time(e1) / steps}
for sid, name in enumerate(("search", "prep", "fallback", "tail", "pack")):
    ms, n = C.c_double(0), C.c_int(0)
    lib.vqb_debug_stage_time_ms(sid, C.byref(ms), C.byref(n))
    out[name] = ms.value / max(1, n.value)
out["losses"] = F.vq_finalize(stats, K, D, 0.25).tolist()
print(out)
