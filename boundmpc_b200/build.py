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
            + [os.path.join(HERE, "..", "include", "boundmpc_b200.h")])
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-diag-suppress", "128"]


def _digest():
    from ._buildutil import content_hash
    return content_hash([os.path.join(CSRC, s) for s in SOURCES] + _headers(), " ".join(NVCC_FLAGS))


def needs_build():
    from ._buildutil import is_current
    return not is_current(LIB, _digest())


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
    from ._buildutil import build_lock, is_current, mark_current
    digest = _digest()
    if not force and is_current(LIB, digest):
        return LIB
    with build_lock(LIB):
        if not force and is_current(LIB, digest):          # another process built it while this one waited
            return LIB
        nvcc = os.environ.get("NVCC", "nvcc")
        tmp = f"{LIB}.tmp.{os.getpid()}"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIB)
        mark_current(LIB, digest)
        if verbose:
            print(r.stderr)
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, timing="--timing" in sys.argv, defines=defs, out=outs[0] if outs else None))
