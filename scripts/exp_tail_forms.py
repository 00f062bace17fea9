"""A/B of the tail kernels (round 2): stage times of the training-mode forward for VQB_TAIL_FORM = 0 (round-1 tail_tma_kernel),
2 (tail2_kernel, default tile), 216 / 232 (tail2_kernel with 16- / 32-frame tiles) on cfg-3-like, cfg-2, a collapsed-codebook
cfg-4-like case and a mid-size case.  usage: python scripts/exp_tail_forms.py [--quick]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vq_b200  # noqa: F401
from vq_b200 import _lib, functional as F

dev = torch.device("cuda:0")
lib = _lib.lib()
quick = "--quick" in sys.argv
CASES = [("cfg3-like", 256 if not quick else 64, 256, 16384, 8192, "randn"), ("cfg2", 64, 64, 16384, 1024, "randn"),
         ("cfg4-collapsed", 64, 64, 11000, 512, "collapsed"), ("mid", 64, 128, 16384, 2048, "randn"), ("cfg5-like", 64, 64, 11000, 512, "randn")]
os.environ["VQB_EXPERIMENTS"] = "1"
for name, B, D, W, K, kind in CASES:
    g = torch.Generator(device=dev).manual_seed(42)
    cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
    z = torch.randn(B, D, W, device=dev, generator=g)
    if kind == "collapsed":
        mu = torch.randn(D, device=dev, generator=g) * 3
        z = z * 0.1 + mu[None, :, None]
        cb[5] = mu
    stats = torch.empty(_lib.stats_len(K, D), device=dev)
    ref = None
    for form in ("0", "2", "216", "232"):
        os.environ["VQB_TAIL_FORM"] = form
        lib.vqb_debug_reload_env()
        for _ in range(3):
            idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
        torch.cuda.synchronize()
        lib.vqb_debug_kernel_timing(1)
        steps = 10
        for _ in range(steps):
            idx, q, st = F.vq_forward(z, cb, precision="bf16", want_q=True, want_resid=True, stats=stats)
        torch.cuda.synchronize()
        out = {"case": name, "form": form, "N": B * W, "K": K, "D": D}
        for sid, sname in enumerate(("search", "prep", "fallback", "tail", "pack")):
            ms, n = C.c_double(0), C.c_int(0)
            lib.vqb_debug_stage_time_ms(sid, C.byref(ms), C.byref(n))
            out[sname] = round(ms.value / max(1, n.value), 4)
        lib.vqb_debug_kernel_timing(0)
        out["tail_GBs"] = round((8 * D + 8) * B * W / (out["tail"] * 1e-3) / 1e9, 1)
        losses = F.vq_finalize(st, K, D, 0.25).tolist()
        cur = (idx.clone(), q.clone(), st.clone())
        if ref is None:
            ref = cur
            out["check"] = "reference form"
        else:
            same = torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and torch.equal(cur[2][:K], ref[2][:K])
            err = ((cur[2][K:] - ref[2][K:]).abs().max() / ref[2][K:].abs().max().clamp_min(1e-30)).item()
            out["check"] = f"idx/q/counts identical={same}, max stats err {err:.2e}"
        out["ppl"] = round(losses[2], 2)
        print(json.dumps(out), flush=True)
    del z, q, idx
    torch.cuda.empty_cache()
