"""pytest configuration: the `gpu` marker and shared fixtures (golden vectors, oracle import)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def _seeded(seed, shape, scale=1.0):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * np.float32(scale)


# name -> how to rebuild inputs that were too large to commit (oracle/make_golden.py case (v))
REGENERATED = {
    "randn_k8192_d256": lambda: (_seeded(7, (2, 256, 320)), _seeded(8, (8192, 256))),
}
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz")) if os.path.isdir(GOLDEN_DIR) else []


def load_golden(name):
    import hashlib
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    if "z" not in g:
        z, cb = REGENERATED[name]()
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        assert sha(z) == str(g["z_sha"]) and sha(cb) == str(g["codebook_sha"]), \
            "numpy generator drift: regenerated inputs do not match the fixture's sha256"
        g["z"], g["codebook"] = z, cb
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


@pytest.fixture
def experiment_env():
    """Set VQB_* experiment switches for one test: they are honoured only with VQB_EXPERIMENTS=1 and are read once, so the
    library is told to re-read them after every change (vqb_debug_reload_env) and again when the test is over."""
    from vq_b200 import _lib
    saved = {}

    def set_env(**kv):
        for k, v in {"VQB_EXPERIMENTS": "1", **kv}.items():
            saved.setdefault(k, os.environ.get(k))
            os.environ[k] = str(v)
        _lib.lib().vqb_debug_reload_env()

    yield set_env
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    _lib.lib().vqb_debug_reload_env()


def load_large_golden(name="large_default_init_k512_d64_n176000"):
    """The N = 176 000 tie-heavy fixture (oracle/make_golden.py::run_large_case): latents regenerated from the numpy seed."""
    import hashlib
    g = dict(np.load(os.path.join(GOLDEN_DIR, "large", name + ".npz"), allow_pickle=False))
    z = _seeded(21, tuple(int(v) for v in g["shape"]))
    assert hashlib.sha256(np.ascontiguousarray(z).tobytes()).hexdigest() == str(g["z_sha"]), "numpy generator drift"
    g["z"] = z
    return g
