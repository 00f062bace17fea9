"""Hand-off timeline of tc_search_kernel (debug build, scripts/build_trace.py): SM-clock timestamps of the MMA issuer, the converter
warps and two epilogue warps of CTA 0 / CTA 1, printed as per-tile intervals.
usage: python scripts/trace_tc.py [case] [out.json]      (cases as in scripts/exp_env_sweep.py; experiment switches from the environment)"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vq_b200  # noqa: F401
from vq_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "scripts", "_trace", "libvqb_b200_trace.so")
from vq_b200 import functional as F

CASES = {"cfg2": (64, 64, 16384, 1024), "cfg5": (64, 64, 11000, 512), "mid": (64, 128, 16384, 2048), "cfg3s": (32, 256, 16384, 8192)}
case = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
out = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/trace_{case}.json"
B, D, W, K = CASES[case]
dev = torch.device("cuda:0")
lib = _lib.lib()
lib.vqb_debug_set_trace.restype = C.c_int
lib.vqb_debug_set_trace.argtypes = [C.c_void_p, C.c_uint]
z = torch.randn(B, D, W, device=dev, generator=torch.Generator(device=dev).manual_seed(42))
cb = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(4242))
prec = os.environ.get("PREC", "bf16")
for _ in range(3):
    F.vq_forward(z, cb, precision=prec, want_q=True, want_resid=True)
torch.cuda.synchronize()
cap = 1 << 16
buf = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
assert lib.vqb_debug_set_trace(buf.data_ptr(), cap) == 0
F.vq_forward(z, cb, precision=prec, want_q=True, want_resid=True)
torch.cuda.synchronize()
lib.vqb_debug_set_trace(None, 0)
recs = buf.cpu().numpy().astype("uint64").reshape(2, 8, 8, 1024)
res = {}
for cta in range(2):
    ev = {}
    for role in range(8):
        for event in range(8):
            col = recs[cta, role, event]
            hit = {int(i): int(col[i]) & 0xFFFFFFFF for i in range(1024) if int(col[i]) >> 63}
            if hit:
                ev[f"{role}.{event}"] = hit
    res[cta] = ev
json.dump(res, open(out, "w"))
# summary for CTA 0: steady-state tiles 40..79
names = {"1.0": "mma wait_tempty", "1.1": "mma stage_free", "1.2": "mma A ready", "1.3": "mma B ready", "1.4": "mma bias ready",
         "4.0": "epi0 wait_tfull", "4.1": "epi0 acc visible", "4.2": "epi0 released", "4.3": "epi0 resolve start", "4.4": "epi0 resolved",
         "4.5": "epi0 past bar2", "4.6": "epi0 past bar1", "7.0": "epi0 cutoff known", "7.1": "epi0 own slots done", "7.2": "epi0 global done", "6.0": "epi0 slab0 in regs", "6.1": "epi0 slab0 scanned", "6.2": "epi0 slab1 in regs", "6.3": "epi0 slab1 scanned", "5.0": "epi13 wait_tfull", "5.1": "epi13 acc visible", "5.2": "epi13 released"}
for cta in range(2):
    e = res[cta]
    if "1.4" not in e:
        continue
    print(f"--- CTA {cta}: tile = one 128 x 256 accumulator; clocks relative to the MMA commit of tile 40")
    t0 = e["1.4"].get(40)
    for t in range(40, 52):
        row = [f"tile {t:3d}"]
        for key in ("1.0", "1.1", "1.2", "1.3", "1.4", "4.0", "4.1", "6.0", "6.1", "6.2", "4.2", "6.3", "5.1", "5.2", "4.3", "4.6", "7.0", "7.1", "7.2", "4.4", "4.5"):
            v = e.get(key, {}).get(t)
            dv = ((v - t0) & 0xFFFFFFFF) if v is not None and t0 is not None else -1
            if dv > 0x7FFFFFFF:
                dv -= 1 << 32
            row.append(f"{names[key].split(' ', 1)[1][:10]}={dv}")
        print(" ".join(row))
    c = sorted(e["1.4"].items())
    if len(c) > 60:
        print("mean cycles per tile (MMA commit to commit, tiles 20..):", (((c[-1][1] - c[20][1]) & 0xFFFFFFFF) / (c[-1][0] - c[20][0])))
