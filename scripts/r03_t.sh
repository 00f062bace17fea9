#!/bin/bash
export VQB_EXPERIMENTS=1
python - <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import vq_b200
from vq_b200 import functional as F, _lib
rng = np.random.default_rng(31)
for K in (512, 1024, 4096):
    B, D, W = 2, 64, 2048
    u = rng.standard_normal(D).astype(np.float32); u /= np.linalg.norm(u)
    cb = np.random.default_rng(30).standard_normal((K, D), dtype=np.float32)
    cb = np.ascontiguousarray(cb[np.argsort(cb @ u)])
    a = rng.uniform(3.0, 6.0, (B, 1, W)).astype(np.float32)
    z = (a * u[None, :, None] + 0.05 * rng.standard_normal((B, D, W))).astype(np.float32)
    for env in ({}, {"VQB_TC_EPI": "1"}, {"VQB_TC_EVSM": "-2"}, {"VQB_TC_EVSM": "0"}):
        for k in ("VQB_TC_EPI", "VQB_TC_EVSM"): os.environ.pop(k, None)
        os.environ.update(env)
        _lib.lib().vqb_debug_reload_env()
        F.vq_forward(torch.from_numpy(z).cuda(), torch.from_numpy(cb).cuda(), precision="bf16", want_q=True, want_resid=True)
        torch.cuda.synchronize()
        print("adversarial K", K, env, F.debug_counters(), "of", B * W, "frames")
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adversarial" 2>&1 | tail -3
