"""Build libvqb_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python multi-source-lms-for-audio_b200/build.py [--force] [--verbose]

The shared library lands next to this file; it is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libvqb_b200.so")
SOURCES = ["vqb_api.cu", "vqb_kernels.cu", "vqb_tail2.cu", "vqb_tail3.cu", "vqb_tc.cu", "vqb_comm.cu", "vqb_host.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libvqb_b200.so cannot be built (there is no non-CUDA fallback)")
    return exe


def _stamp() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/vqb.h"]:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p):
            h.update(name.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + NVCC_FLAGS + ["cudart-shared"]).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    exe = nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [exe, *ARCH, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ, src[:-3] + ".log"), "w") as f:   # ptxas -v: registers / spills / smem per kernel
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # --cudart shared: reuse the CUDA runtime the host process (torch) already carries instead of embedding a second, static
    # copy with its whole export table; _lib.lib() imports torch first, so libcudart.so.12 is resolved by soname
    cmd = [exe, *ARCH, "-shared", "--cudart", "shared", "-o", LIB, *objs, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
