"""bf16 vs tf32 shortlist (and the exact fp32 search) at the BASELINE shapes: stage times, shortlist statistics, index agreement.
usage: python scripts/exp_precisions.py [--with-fp32]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import _lib, functional as F

dev = torch.device("cuda:0")
lib = _lib.lib()
CASES = [("cfg2", 64, 64, 16384, 1024), ("cfg5/cfg4 shape", 64, 64, 11000, 512), ("cfg1", 2, 64, 11000, 512), ("mid D128", 64, 128, 16384, 2048),
         ("cfg3 slice", 128, 256, 16384, 8192)]
precs = ["bf16", "tf32"] + (["fp32"] if "--with-fp32" in sys.argv else [])
for name, B, D, W, K in CASES:
    g = torch.Generator(device=dev).manual_seed(42)
    z = torch.randn(B, D, W, device=dev, generator=g)
    cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
    stats = torch.empty(_lib.stats_len(K, D), device=dev)
    ref = None
    for prec in precs:
        for _ in range(3):
            idx, q, st = F.vq_forward(z, cb, precision=prec, want_q=True, want_resid=True, stats=stats)
        torch.cuda.synchronize()
        lib.vqb_debug_kernel_timing(1)
        steps = 10 if prec != "fp32" else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            idx, q, st = F.vq_forward(z, cb, precision=prec, want_q=True, want_resid=True, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        out = {"case": name, "N": B * W, "K": K, "D": D, "precision": prec, "step_ms": round(e0.elapsed_time(e1) / steps, 4)}
        for sid, sname in enumerate(("search", "prep", "fallback", "tail", "pack")):
            ms, n = C.c_double(0), C.c_int(0)
            lib.vqb_debug_stage_time_ms(sid, C.byref(ms), C.byref(n))
            out[sname] = round(ms.value / max(1, n.value), 4)
        lib.vqb_debug_kernel_timing(0)
        c = F.debug_counters(dev)
        out.update(rescored=c["rescored"], fallback_frames=c["fallback"], mean_candidates=round(c["shortlisted"] / (B * W), 4))
        peak = 1652.1 if prec == "bf16" else 826.0
        out["search_TFs"] = round(2.0 * K * D * B * W / (out["search"] * 1e-3) / 1e12, 1)
        out["frac_of_burst_peak"] = round(out["search_TFs"] / peak, 3) if prec != "fp32" else None
        if ref is None:
            ref = idx.clone()
        else:
            out["idx_differ_from_bf16"] = int((idx != ref).sum())
        print(json.dumps(out), flush=True)
    del z, q, idx
    torch.cuda.empty_cache()
