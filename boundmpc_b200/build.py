"""In-tree build of the CUDA library (nvcc, sm_100a).  `python -m boundmpc_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libboundmpc_b200.so")
SOURCES = ["bmpc_kernels.cu"]
import glob


def _headers():
    """Everything the translation unit can include, plus this file (the flags are part of the build)."""
    return (sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")))
            + [os.path.join(HERE, "..", "include", "boundmpc_b200.h"), os.path.abspath(__file__)])
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-diag-suppress", "128"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, s) for s in SOURCES] + _headers())


TIMING_LIB = os.path.join(HERE, "libboundmpc_b200_timing.so")   # development build with per-phase cycle counters


def build(force=False, verbose=False, timing=False, defines=(), out=None):
    if timing or out:
        nvcc = os.environ.get("NVCC", "nvcc")
        target = out or TIMING_LIB
        cmd = [nvcc] + NVCC_FLAGS + (["-DBMPC_TIMING"] if timing else []) + ["-D" + d for d in defines] + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        return target
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, timing="--timing" in sys.argv, defines=defs, out=outs[0] if outs else None))
