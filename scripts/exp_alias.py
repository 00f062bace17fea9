"""Does the tail's time depend on where `quantized` lies relative to the latents?  tail3_kernel loads a tile and stores the same
tile's `quantized` at the same offset of another tensor; with power-of-two tensor sizes the two streams can alias in the DRAM
channel / bank mapping.  z and q are carved out of ONE buffer with a chosen gap.  usage: python scripts/exp_alias.py"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import _lib, functional as F

dev = torch.device("cuda:0")
lib = _lib.lib()
os.environ["VQB_EXPERIMENTS"] = "1"
for name, (B, D, W, K) in {"cfg2": (64, 64, 16384, 1024), "cfg3s": (128, 256, 16384, 8192)}.items():
    n = B * D * W
    cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
    stats = torch.empty(_lib.stats_len(K, D), device=dev)
    for gap in (0, 256, 4096, 65536, 1 << 20, (1 << 20) + 4096 + 256, 3 << 20):
        buf = torch.empty(2 * n + gap // 4 + 1024, device=dev)
        z = buf[:n].view(B, D, W)
        z.copy_(torch.randn(B, D, W, device=dev, generator=torch.Generator(device=dev).manual_seed(42)))
        q = buf[n + gap // 4: 2 * n + gap // 4].view(B, D, W)
        orig = torch.empty_like
        torch.empty_like = lambda t, *a, **k: q if t is z else orig(t, *a, **k)
        try:
            for form in ("3", "2"):
                os.environ["VQB_TAIL_FORM"] = form
                lib.vqb_debug_reload_env()
                best = 1e9
                for rep in range(4):
                    for _ in range(2):
                        F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
                    torch.cuda.synchronize()
                    lib.vqb_debug_kernel_timing(1)
                    for _ in range(8):
                        F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
                    torch.cuda.synchronize()
                    ms, cnt = C.c_double(0), C.c_int(0)
                    lib.vqb_debug_stage_time_ms(3, C.byref(ms), C.byref(cnt))
                    lib.vqb_debug_kernel_timing(0)
                    best = min(best, ms.value / max(1, cnt.value))
                print(json.dumps({"case": name, "gap_bytes": gap, "q_minus_z_mod_2MiB": (q.data_ptr() - z.data_ptr()) % (2 << 20), "tail_form": form,
                                  "tail_ms": round(best, 4)}), flush=True)
        finally:
            torch.empty_like = orig
        del buf, z, q
        torch.cuda.empty_cache()
