"""ctypes wrapper of the host-emulation build of the kernel source.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libbmpc_emu.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


class Cfg(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("nr_segs", ctypes.c_int32), ("dt", ctypes.c_double),
                ("u_min", ctypes.c_double), ("u_max", ctypes.c_double), ("ut_min", ctypes.c_double), ("ut_max", ctypes.c_double),
                ("q_lim_lower", ctypes.c_double * 7), ("q_lim_upper", ctypes.c_double * 7),
                ("dq_lim_lower", ctypes.c_double * 7), ("dq_lim_upper", ctypes.c_double * 7),
                ("tol", ctypes.c_double), ("max_iter", ctypes.c_int32), ("mu_init", ctypes.c_double),
                ("bound_push", ctypes.c_double), ("device", ctypes.c_int32), ("threads", ctypes.c_int32)]


def make_cfg(N=10, S=4, dt=0.1, tol=1e-9, max_iter=500):
    from boundmpc_b200 import robot_model as rm
    c = Cfg()
    c.N, c.nr_segs, c.dt = N, S, dt
    c.u_min, c.u_max, c.ut_min, c.ut_max = rm.U_MIN, rm.U_MAX, rm.U_MIN, rm.U_MAX
    for i in range(7):
        c.q_lim_lower[i], c.q_lim_upper[i] = rm.Q_LIM_LOWER[i], rm.Q_LIM_UPPER[i]
        c.dq_lim_lower[i], c.dq_lim_upper[i] = rm.DQ_LIM_LOWER[i], rm.DQ_LIM_UPPER[i]
    c.tol, c.max_iter, c.mu_init, c.bound_push, c.device, c.threads = tol, max_iter, 0.0, 0.0, -1, 0
    return c


def build(force=False):
    from boundmpc_b200._buildutil import content_hash, is_current, mark_current, build_lock
    src = [os.path.join(_HERE, "bmpc_emu.cpp")] + [os.path.join(_ROOT, "boundmpc_b200", "csrc", f) for f in
           ("bmpc_common.h", "bmpc_model.cuh", "bmpc_riccati.cuh", "bmpc_ipm.cuh", "bmpc_eval.cuh", "bmpc_host.h", "bmpc_prepare.cuh", "bmpc_post.cuh")]
    src.append(os.path.join(_ROOT, "include", "boundmpc_b200.h"))
    digest = content_hash(src)
    if not force and is_current(_LIB, digest):
        return _LIB
    with build_lock(_LIB):
        if force or not is_current(_LIB, digest):
            tmp = f"{_LIB}.tmp.{os.getpid()}"
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", tmp, src[0]])
            os.replace(tmp, _LIB)
            mark_current(_LIB, digest)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def solve(x0, p, N=10, S=4, dt=0.1, tol=1e-9, max_iter=500, sliced=False):
    """sliced=True: the park / resume path of the two-pass scheduling (adds the key 'hard')."""
    x0 = np.ascontiguousarray(np.atleast_2d(x0), float)
    p = np.ascontiguousarray(np.atleast_2d(p), float)
    B, n = x0.shape
    m = 43 * N
    cfg = make_cfg(N, S, dt, tol, max_iter)
    x, g, lg, lx = np.empty((B, n)), np.empty((B, m)), np.empty((B, m)), np.empty((B, n))
    f, kkt = np.empty(B), np.empty(B)
    it, st = np.empty(B, np.int32), np.empty(B, np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    if sliced:
        hard = np.empty(B, np.int32)
        rc = lib().emu_solve_sliced(ctypes.byref(cfg), B, _p(x0), _p(p), _p(x), _p(g), _p(lg), _p(lx), _p(f),
                                    it.ctypes.data_as(i32p), st.ctypes.data_as(i32p), _p(kkt), hard.ctypes.data_as(i32p))
        assert rc == 0
        return dict(x=x, g=g, lam_g=lg, lam_x=lx, f=f, iters=it, status=st, kkt=kkt, hard=hard)
    rc = lib().emu_solve(ctypes.byref(cfg), B, _p(x0), _p(p), _p(x), _p(g), _p(lg), _p(lx), _p(f),
                         it.ctypes.data_as(i32p), st.ctypes.data_as(i32p), _p(kkt))
    assert rc == 0
    return dict(x=x, g=g, lam_g=lg, lam_x=lx, f=f, iters=it, status=st, kkt=kkt)


def evaluate(x, p, lam=None, N=10, S=4, dt=0.1, want_jac=True, want_hess=True):
    x = np.ascontiguousarray(np.atleast_2d(x), float)
    p = np.ascontiguousarray(np.atleast_2d(p), float)
    B, n = x.shape
    m = 43 * N
    cfg = make_cfg(N, S, dt)
    lam = None if lam is None else np.ascontiguousarray(np.atleast_2d(lam), float)
    f, g, d, grad = np.empty(B), np.empty((B, m)), np.empty((B, 12 * N)), np.empty((B, n))
    jac = np.empty((B, 48 * N, n)) if want_jac else None
    hess = np.empty((B, n, n)) if want_hess else None
    rc = lib().emu_eval(ctypes.byref(cfg), B, _p(x), _p(p), _p(lam), _p(f), _p(g), _p(d), _p(grad), _p(jac), _p(hess))
    assert rc == 0
    return dict(f=f, g=g, d=d, grad=grad, jac=jac, hess=hess)


def prepare(tabs, path_id, sector, state, prev, N=10, S=4):
    """Serial form of the CUDA parameter builder.  tabs [P, J, 41]; returns x0, p, new sector."""
    tabs = np.ascontiguousarray(tabs, float)
    state = np.ascontiguousarray(np.atleast_2d(state), float)
    prev = np.ascontiguousarray(np.atleast_2d(prev), float)
    B = state.shape[0]
    pid = np.ascontiguousarray(path_id, np.int32)
    sec = np.ascontiguousarray(sector, np.int32).copy()
    x0, p = np.empty((B, 44 * N)), np.empty((B, 141 + 91 * S))
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib().emu_prepare(N, S, B, _p(tabs), tabs.shape[1], pid.ctypes.data_as(i32p), sec.ctypes.data_as(i32p), _p(state), _p(prev),
                           _p(x0), _p(p))
    assert rc == 0
    return x0, p, sec


def post(tabs, path_id, sector, state, w, ec, N=10, S=4, dt=0.1):
    """Serial form of the CUDA post-processing.  Returns traj [B, N, 42] and the next-step state [B, 76]."""
    tabs = np.ascontiguousarray(tabs, float)
    state = np.ascontiguousarray(np.atleast_2d(state), float)
    w = np.ascontiguousarray(np.atleast_2d(w), float)
    B = state.shape[0]
    i32p = ctypes.POINTER(ctypes.c_int32)
    pid = np.ascontiguousarray(path_id, np.int32); sec = np.ascontiguousarray(sector, np.int32); e = np.ascontiguousarray(ec, np.int32)
    traj, so = np.empty((B, N, 42)), np.empty((B, 76))
    cfg = make_cfg(N, S, dt)
    rc = lib().emu_post(ctypes.byref(cfg), B, _p(tabs), tabs.shape[1], pid.ctypes.data_as(i32p), sec.ctypes.data_as(i32p), _p(state), _p(w),
                        e.ctypes.data_as(i32p), _p(traj), _p(so))
    assert rc == 0
    return traj, so


def finish(tabs, path_id, sector, state, x, g, status, prev, ec, advance=True, N=10, S=4, dt=0.1):
    """Serial form of k_finish.  Returns traj, next state, updated prev and error counts."""
    tabs = np.ascontiguousarray(tabs, float)
    state = np.ascontiguousarray(np.atleast_2d(state), float)
    x = np.ascontiguousarray(np.atleast_2d(x), float); g = np.ascontiguousarray(np.atleast_2d(g), float)
    prev = np.ascontiguousarray(np.atleast_2d(prev), float).copy()
    B = state.shape[0]
    i32p = ctypes.POINTER(ctypes.c_int32)
    pid = np.ascontiguousarray(path_id, np.int32); sec = np.ascontiguousarray(sector, np.int32)
    st = np.ascontiguousarray(status, np.int32); e = np.ascontiguousarray(ec, np.int32).copy()
    traj, so = np.empty((B, N, 42)), np.empty((B, 76))
    cfg = make_cfg(N, S, dt)
    rc = lib().emu_finish(ctypes.byref(cfg), B, _p(tabs), tabs.shape[1], pid.ctypes.data_as(i32p), sec.ctypes.data_as(i32p), _p(state), _p(x),
                          _p(g), st.ctypes.data_as(i32p), _p(prev), e.ctypes.data_as(i32p), _p(traj), _p(so), int(advance))
    assert rc == 0
    return traj, so, prev, e


def post_log(tabs, path_id, sector, state, p, w, ec, N=10, S=4, dt=0.1):
    """Serial form of k_post with the logging branch.  Returns traj, next state, ref [B, N, 55], err [B, N, 33]."""
    tabs = np.ascontiguousarray(tabs, float)
    state = np.ascontiguousarray(np.atleast_2d(state), float)
    p = np.ascontiguousarray(np.atleast_2d(p), float); w = np.ascontiguousarray(np.atleast_2d(w), float)
    B = state.shape[0]
    i32p = ctypes.POINTER(ctypes.c_int32)
    pid = np.ascontiguousarray(path_id, np.int32); sec = np.ascontiguousarray(sector, np.int32); e = np.ascontiguousarray(ec, np.int32)
    traj, so, ref, err = np.empty((B, N, 42)), np.empty((B, 76)), np.empty((B, N, 55)), np.empty((B, N, 33))
    cfg = make_cfg(N, S, dt)
    rc = lib().emu_post_log(ctypes.byref(cfg), B, _p(tabs), tabs.shape[1], pid.ctypes.data_as(i32p), sec.ctypes.data_as(i32p), _p(state), _p(p),
                            _p(w), e.ctypes.data_as(i32p), _p(traj), _p(so), _p(ref), _p(err))
    assert rc == 0
    return traj, so, ref, err


def update(tabs, phimax, new_path, cart, state, sector, path_id):
    """Serial form of k_update.  Returns updated copies of state, sector, path_id."""
    tabs = np.ascontiguousarray(tabs, float); phimax = np.ascontiguousarray(phimax, float)
    cart = np.ascontiguousarray(np.atleast_2d(cart), float)
    state = np.ascontiguousarray(np.atleast_2d(state), float).copy()
    i32p = ctypes.POINTER(ctypes.c_int32)
    npth = np.ascontiguousarray(new_path, np.int32); sec = np.ascontiguousarray(sector, np.int32).copy(); pid = np.ascontiguousarray(path_id, np.int32).copy()
    rc = lib().emu_update(state.shape[0], _p(tabs), tabs.shape[1], _p(phimax), npth.ctypes.data_as(i32p), _p(cart), _p(state),
                          sec.ctypes.data_as(i32p), pid.ctypes.data_as(i32p))
    assert rc == 0
    return state, sec, pid


def kkt_step(x, y, s, zs, zL, zU, p, mu, delta_w=0.0, N=10, S=4, dt=0.1):
    """Host build of the kernel's Newton step (Riccati sweep) at given primal-dual points."""
    p = np.ascontiguousarray(np.atleast_2d(p), float)
    v = np.ascontiguousarray(np.concatenate([np.atleast_2d(a) for a in (x, y, s, zs, zL, zU)], axis=1), float)
    B = v.shape[0]
    mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, float), (B,)))
    dw = np.ascontiguousarray(np.broadcast_to(np.asarray(delta_w, float), (B,)))
    dx, ynew, ok = np.empty((B, 44 * N)), np.empty((B, 36 * N)), np.empty(B, np.int32)
    cfg = make_cfg(N, S, dt)
    rc = lib().emu_kkt_step(ctypes.byref(cfg), B, _p(v), _p(p), _p(mu), _p(dw), _p(dx), _p(ynew), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert rc == 0
    return dict(dx=dx, ynew=ynew, ok=ok)
