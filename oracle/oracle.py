"""ctypes wrapper of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see boundmpc_oracle.cpp)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libboundmpc_oracle.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """gcc build of the oracle (content hash, lock, atomic replace: see boundmpc_b200/_buildutil.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(_HERE))
    from boundmpc_b200._buildutil import content_hash, is_current, mark_current, build_lock
    srcs = [os.path.join(_HERE, f) for f in ("boundmpc_oracle.cpp", "ocp_model.hpp", "ad.hpp", "Makefile")]
    # (-march=native: the library is specific to the host CPU, so its feature flags are part of the key and a box with another
    # CPU rebuilds instead of loading code it cannot execute)
    try:
        cpu = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except (OSError, StopIteration):
        cpu = ""
    digest = content_hash(srcs, cpu)
    if not force and is_current(_LIB, digest):
        return _LIB
    with build_lock(_LIB):
        if force or not is_current(_LIB, digest):
            tmp = f"libboundmpc_oracle.tmp.{os.getpid()}.so"
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", f"OUT={tmp}"])
            os.replace(os.path.join(_HERE, tmp), _LIB)
            mark_current(_LIB, digest)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def dims(N=10, S=4):
    n, m, np_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().orc_dims(N, S, ctypes.byref(n), ctypes.byref(m), ctypes.byref(np_))
    return n.value, m.value, np_.value


def bounds(N=10, S=4, dt=0.1):
    n, m, _ = dims(N, S)
    lbx, ubx, lbg, ubg = np.empty(n), np.empty(n), np.empty(m), np.empty(m)
    lib().orc_bounds(N, S, ctypes.c_double(dt), _p(lbx), _p(ubx), _p(lbg), _p(ubg))
    return lbx, ubx, lbg, ubg


def eval_fg(x, p, N=10, S=4, dt=0.1):
    n, m, _ = dims(N, S)
    x = np.ascontiguousarray(x, float)
    p = np.ascontiguousarray(p, float)
    f = ctypes.c_double()
    g = np.empty(m)
    lib().orc_eval(N, S, ctypes.c_double(dt), _p(x), _p(p), ctypes.byref(f), _p(g))
    return f.value, g


def derivs(x, p, lam, N=10, S=4, dt=0.1):
    n, m, _ = dims(N, S)
    x = np.ascontiguousarray(x, float)
    p = np.ascontiguousarray(p, float)
    lam = np.ascontiguousarray(lam, float)
    grad, jac, hess = np.empty(n), np.empty((m, n)), np.empty((n, n))
    lib().orc_derivs(N, S, ctypes.c_double(dt), _p(x), _p(p), _p(lam), _p(grad), _p(jac), _p(hess))
    return grad, jac, hess


def solve(x0, p, N=10, S=4, dt=0.1, tol=1e-8, max_iter=500, mu_init=1e-3, bound_push=1e-3, verbose=0, mu_strategy=-1, max_soc=-1):
    n, m, _ = dims(N, S)
    x0 = np.ascontiguousarray(x0, float)
    p = np.ascontiguousarray(p, float)
    opts = np.array([tol, max_iter, mu_init, bound_push, verbose, mu_strategy, max_soc], float)
    x, g, lam_g, lam_x = np.empty(n), np.empty(m), np.empty(m), np.empty(n)
    f, it, kkt = ctypes.c_double(), ctypes.c_int(), ctypes.c_double()
    st = lib().orc_solve(N, S, ctypes.c_double(dt), _p(x0), _p(p), _p(opts), _p(x), _p(g), _p(lam_g), _p(lam_x),
                         ctypes.byref(f), ctypes.byref(it), ctypes.byref(kkt))
    return dict(x=x, g=g, lam_g=lam_g, lam_x=lam_x, f=f.value, iters=it.value, kkt=kkt.value, status=st)


def derivs_interval(x, p, lam, N=10, S=4, dt=0.1):
    """Interval-form rows: d [12N], grad f [n], jac [48N, n], hess of f + lam.(c, d) [n, n]."""
    n, m, _ = dims(N, S)
    x = np.ascontiguousarray(x, float)
    p = np.ascontiguousarray(p, float)
    lam = np.ascontiguousarray(lam, float)
    d, grad, jac, hess = np.empty(12 * N), np.empty(n), np.empty((48 * N, n)), np.empty((n, n))
    lib().orc_derivs_interval(N, S, ctypes.c_double(dt), _p(x), _p(p), _p(lam), _p(d), _p(grad), _p(jac), _p(hess))
    return d, grad, jac, hess


class OracleSolver:
    """The solver call surface (`solver(x0=, p=)`, `stats()`, `bounds()`, `eval_batch(...)['d']`) on top of the CPU oracle.
    Used by bench.py's reference arm to generate its inputs without touching the CUDA library (boundmpc_b200.batches
    needs a solver for the nominal closed loops and for the bound-excess probe of its repaired generator)."""

    def __init__(self, N=10, nr_segs=4, dt=0.1, tol=1e-9):
        self.N, self.nr_segs, self.dt, self.tol = N, nr_segs, dt, tol
        self.n, self.m, self.np = dims(N, nr_segs)
        self._stats = {}

    def bounds(self):
        return bounds(self.N, self.nr_segs, self.dt)

    def __call__(self, x0=None, p=None, **kw):
        r = solve(np.asarray(x0, float).ravel(), np.asarray(p, float).ravel(), self.N, self.nr_segs, self.dt, tol=self.tol)
        self._stats = dict(iter_count=int(r["iters"]), success=r["status"] == 0, return_status=str(r["status"]))
        return dict(x=r["x"], g=r["g"], lam_g=r["lam_g"], lam_x=r["lam_x"], f=r["f"])

    def stats(self):
        return dict(self._stats)

    def generate_dependencies(self, *a, **k):
        return None

    def eval_batch(self, x, p, lam=None, want_jac=False, want_hess=False):
        """Interval-form inequality rows d [B, 12 N] (all the batch generator reads)."""
        x, p = np.atleast_2d(x), np.atleast_2d(p)
        d = np.empty((len(x), 12 * self.N))
        g, f = np.empty(self.m), ctypes.c_double()
        for i in range(len(x)):
            xi, pi = np.ascontiguousarray(x[i], float), np.ascontiguousarray(p[i], float)
            lib().orc_eval_d(self.N, self.nr_segs, ctypes.c_double(self.dt), _p(xi), _p(pi), ctypes.byref(f), _p(g), _p(d[i]))
        return {"d": d}
