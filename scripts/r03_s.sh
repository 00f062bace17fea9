#!/bin/bash
export VQB_EXPERIMENTS=1
python - <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import vq_b200
from vq_b200 import functional as F
rng = np.random.default_rng(31)
for K in (512, 1024, 4096):
    B, D, W = 2, 64, 2048
    u = rng.standard_normal(D).astype(np.float32); u /= np.linalg.norm(u)
    s_k = np.linspace(0.0, 4.0, K, dtype=np.float32)
    cb = (s_k[:, None] * u[None, :] + 0.01 * rng.standard_normal((K, D))).astype(np.float32)
    a = rng.uniform(3.0, 6.0, (B, 1, W)).astype(np.float32)
    z = (a * u[None, :, None] + 0.05 * rng.standard_normal((B, D, W))).astype(np.float32)
    for env in ({}, {"VQB_TC_EPI": "1"}, {"VQB_TC_EVSM": "-2"}):
        for k in ("VQB_TC_EPI", "VQB_TC_EVSM"): os.environ.pop(k, None)
        os.environ.update(env)
        from vq_b200 import _lib; _lib.lib().vqb_debug_reload_env()
        F.vq_forward(torch.from_numpy(z).cuda(), torch.from_numpy(cb).cuda(), precision="bf16", want_q=True, want_resid=True)
        torch.cuda.synchronize()
        print("adversarial K", K, env, F.debug_counters(), "of", B * W, "frames")
PY
for e in "VQB_TC_EPI=1" "VQB_TAIL_LPF=2" "VQB_TAIL_LPF=8 VQB_TAIL_AHEAD=1" "VQB_TAIL_FORM=300" "VQB_TC_EVSM=-2"; do env $e timeout 300 python scripts/stress_shapes.py 200 7 | tail -1 | sed "s/^/$e: /"; done
