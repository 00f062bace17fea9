"""Debug build of the library with the hand-off trace of tc_search_kernel compiled in (-DVQB_TC_TRACE) -> scripts/_trace/libvqb_b200_trace.so.
Never shipped, never loaded by the package; scripts/trace_tc.py points the ctypes loader at it."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-source-lms-for-audio_b200"))
import build as B  # noqa: E402

out_dir = os.path.join(ROOT, "scripts", "_trace")
os.makedirs(out_dir, exist_ok=True)
objs = []
procs = []
for src in B.SOURCES:
    obj = os.path.join(out_dir, src[:-3] + ".o")
    flags = [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    procs.append((src, subprocess.Popen([B.nvcc(), *B.ARCH, *flags, "-DVQB_TC_TRACE", *sys.argv[1:], "-c", os.path.join(B.CSRC, src), "-o", obj])))
    objs.append(obj)
for src, p in procs:
    if p.wait() != 0:
        raise SystemExit(f"nvcc failed on {src}")
lib = os.path.join(out_dir, "libvqb_b200_trace.so")
subprocess.check_call([B.nvcc(), *B.ARCH, "-shared", "--cudart", "shared", "-o", lib, *objs, "-ldl"])
print(lib)
