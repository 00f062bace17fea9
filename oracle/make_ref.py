"""Recipe for oracle/_ref/: the UNMODIFIED reference quantiser, made available to the GPU box.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_ref.py          (run in the build container; __graft_entry__.build() runs it when /root/reference exists)

The reference is a pure-Python torch module with nothing to compile and no installable package (`pip install /root/reference`
fails: neither setup.py nor pyproject.toml), and /root/reference does not exist on the GPU box.  Its quantiser
(src/model/components/vector_quantizer.py:6-54) imports only torch, so this recipe places a byte-identical copy of that ONE
file under oracle/_ref/ (git-ignored output directory - never committed - that travels to the GPU box with the snapshot,
like a built .so) together with its sha256.  bench.py's CPU arm (`--impl reference`, `cpu_baseline`) and the parity tests
import `VectorQuantizer` from /root/reference when it is there, else from oracle/_ref/, and only then fall back to the
torch port (oracle/ref_port_torch.py), reporting which one ran (`kind`: "reference" | "port").

Nothing in the product package imports anything from here.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys

REF_ROOT = "/root/reference"
REL = os.path.join("src", "model", "components", "vector_quantizer.py")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def make() -> str | None:
    src = os.path.join(REF_ROOT, REL)
    if not os.path.isfile(src):
        return None
    dst = os.path.join(OUT, REL)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    digest = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "SHA256"), "w") as f:
        f.write(f"{digest}  {REL}\n")
    return dst


def find_reference_file() -> tuple[str, str] | None:
    """(path, where) of the reference quantiser source: the live reference, a pip-style install, or oracle/_ref."""
    root = os.path.dirname(HERE)
    for base, where in ((REF_ROOT, "/root/reference"), (os.path.join(root, "baseline", "_ref"), "baseline/_ref"), (OUT, "oracle/_ref")):
        p = os.path.join(base, REL)
        if os.path.isfile(p):
            return p, where
    return None


def load_reference_class():
    """The reference's VectorQuantizer class (unmodified source, executed as-is) and where it came from, or (None, None)."""
    found = find_reference_file()
    if found is None:
        return None, None
    path, where = found
    spec = importlib.util.spec_from_file_location("_reference_vector_quantizer", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.VectorQuantizer, where


if __name__ == "__main__":
    out = make()
    print(out if out else f"{REF_ROOT} not present: nothing to do", file=sys.stderr if out is None else sys.stdout)
