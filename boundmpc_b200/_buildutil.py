"""Shared helpers of the in-tree builds (CUDA library, CPU oracle, host emulation): staleness by CONTENT hash (file times
do not survive the copy onto a GPU box), one builder at a time (file lock: eight ranks of a torchrun launch import at
once), and atomic replacement of the output (a reader never sees a half-written library)."""
import contextlib
import fcntl
import hashlib
import os


def content_hash(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_current(lib, digest):
    try:
        return os.path.exists(lib) and open(lib + ".hash").read().strip() == digest
    except OSError:
        return False


def mark_current(lib, digest):
    tmp = f"{lib}.hash.{os.getpid()}"
    with open(tmp, "w") as f:
        f.write(digest)
    os.replace(tmp, lib + ".hash")


@contextlib.contextmanager
def build_lock(lib):
    with open(lib + ".lock", "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)
