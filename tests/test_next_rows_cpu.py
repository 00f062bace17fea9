"""CPU tests of the rows next to the hot path: codebook on-disk formats (f3) and checkpoint-key compatibility of the
VQ-VAE step harness (f1)."""
import os

import numpy as np
import torch

import vq_b200
from vq_b200 import codebook_io
from conftest import GOLDEN_DIR, load_golden


def test_codebook_csv_roundtrip_both_layouts(tmp_path):
    g = load_golden("trained_codebook_csv")          # the reference's committed logs/best_checkpoint/codebook.csv
    cb = torch.from_numpy(g["codebook"])
    assert np.array_equal(g["csv_header_row"].reshape(-1), np.arange(64))       # that file carries a pandas header row
    for header in (False, True):                     # vqvae.py:241-243 writes none; the committed file has one
        p = os.path.join(tmp_path, f"cb_{header}.csv")
        codebook_io.save_codebook_csv(cb, p, header=header)
        back = codebook_io.load_codebook_csv(p, expect_rows=512)
        assert back.shape == (512, 64) and torch.equal(back, cb), "no codeword may be lost or altered"
    # a header-less file whose first codeword happens to be 0..D-1 is only disambiguated by expect_rows
    odd = torch.arange(8.0).repeat(3, 1)
    p = os.path.join(tmp_path, "odd.csv")
    codebook_io.save_codebook_csv(odd, p, header=False)
    assert codebook_io.load_codebook_csv(p, expect_rows=3).shape == (3, 8)


def test_state_dict_keys_match_reference_checkpoint():
    g = np.load(os.path.join(GOLDEN_DIR, "vqvae_step", "vqvae_step_b2_t4096.npz"))
    ref_sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    model = vq_b200.VQVAEStep()
    missing, unexpected = model.load_state_dict(ref_sd, strict=True)     # strict loads at main.py:66,117,197
    assert not missing and not unexpected
    assert codebook_io.STATE_DICT_KEY in ref_sd
    assert torch.equal(codebook_io.codebook_from_state_dict(ref_sd), model.vector_quantizer.codebook.weight.detach())


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU port of the reference on the host cores): exactly one JSON line on stdout with
    the contract's keys; under torchrun only rank 0 prints."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0"})
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert out.returncode == 0 and len(lines) == 1, (out.returncode, out.stdout, out.stderr[-500:])
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    # "reference": the unmodified reference module (from /root/reference, baseline/_ref or oracle/_ref); "port" only if none is there
    from oracle import make_ref
    want_kind = "reference" if make_ref.find_reference_file() else "port"
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
