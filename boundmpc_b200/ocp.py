"""Drop-in replacement of the reference's `casadi_ocp_formulation.setup_optimization_problem`.

Same signature and return tuple as bound_mpc/bound_mpc/BoundMPC/casadi_ocp_formulation.py:9-11,391;
the returned handle has the call surface `BoundMPC` uses on the CasADi/Ipopt solver object
(BoundMPC.py:150-161, 446-457): `solver(x0=, lbx=, ubx=, lbg=, ubg=, p=) -> {'x','g','lam_g','lam_x','f'}`
with 1-D numpy arrays, `solver.stats() -> {'iter_count','success','return_status'}` and
`solver.generate_dependencies(...)` (a no-op: nothing is generated at run time).  Behind it
sits the CUDA library (include/boundmpc_b200.h) — no CasADi, no Ipopt, no CPU path.
`solve_batch` is the batched entry the reference does not have.
"""
import ctypes
import numpy as np
from . import _cabi

RETURN_STATUS = {0: "Solve_Succeeded", 1: "Maximum_Iterations_Exceeded", 2: "Restoration_Failed",
                 3: "Error_In_Step_Computation", 4: "Invalid_Number_Detected", 5: "Infeasible_Problem_Detected"}


def _g_names(N):
    """g_names list of the reference (casadi_ocp_formulation.py:271-349)."""
    per = (["Dynamical System"] * 21 + ["Dynamical System"] * 2 + ["Dynamical System"] * 3 +
           ["Path parameter", "Path velocity", "Tangential orientation error"] +
           ["Orthogonal position error"] * 2 + ["Orthogonal orientation error"] * 2)
    return per * N


def resolve_tol(solver_opts):
    """Termination tolerance of a handle.  `solver_opts['ipopt']['tol']` is honoured like the reference's Ipopt honours it
    (BoundMPC.py:121 passes 10e-6); `solver_opts['b200']['tol']` overrides it; without either the handle converges to 1e-9.
    Why a tighter default: Ipopt at 1e-5 stops 1e-4 (relative) away from the KKT point (SURVEY App. D.5,
    tests/test_emu_parity.py), so comparing a primal trajectory to 1e-6 relative needs the tight setting (about 2.6 more
    iterations per solve on the bench workload: 10.1 against 7.4)."""
    so = solver_opts or {}
    if "tol" in so.get("b200", {}):
        return float(so["b200"]["tol"])
    if "tol" in so.get("ipopt", {}):
        return float(so["ipopt"]["tol"])
    return 1e-9


_DUAL_NOTE = [False]


def note_dual_start(lam_g0, lam_x0):
    """Warn (once per process) when a caller passes initial multipliers: the solver ignores them.  Measured on the CPU
    oracle along the experiment1 / experiment2 closed loops a dual warm start would save 17 - 21 % of the iterations at the
    reference's tolerance; it is not implemented because the reference does not use it either."""
    import warnings
    given = any(v is not None and np.any(np.asarray(v, float) != 0.0) for v in (lam_g0, lam_x0))
    if given and not _DUAL_NOTE[0]:
        _DUAL_NOTE[0] = True
        warnings.warn("boundmpc_b200: lam_g0 / lam_x0 are ignored (every solve starts from zero equality multipliers, as the "
                      "reference's own call does, BoundMPC.py:451-452)", RuntimeWarning, stacklevel=3)
    return given


class BatchSolver:
    """Handle on the CUDA solver for one OCP shape (N, nr_segs, dt, limits)."""

    def __init__(self, N, nr_segs, dt, u_min, u_max, ut_min, ut_max, q_lim_lower, q_lim_upper, dq_lim_lower,
                 dq_lim_upper, solver_opts=None, device=-1):
        opts = dict((solver_opts or {}).get("ipopt", {}))
        opts.update((solver_opts or {}).get("b200", {}))
        cfg = _cabi.BmpcConfig()
        cfg.N, cfg.nr_segs, cfg.dt = int(N), int(nr_segs), float(dt)
        cfg.u_min, cfg.u_max, cfg.ut_min, cfg.ut_max = float(u_min), float(u_max), float(ut_min), float(ut_max)
        for i in range(7):
            cfg.q_lim_lower[i], cfg.q_lim_upper[i] = float(q_lim_lower[i]), float(q_lim_upper[i])
            cfg.dq_lim_lower[i], cfg.dq_lim_upper[i] = float(dq_lim_lower[i]), float(dq_lim_upper[i])
        cfg.tol = resolve_tol(solver_opts)
        cfg.max_iter = int(opts.get("max_iter", 500))
        cfg.mu_init = float(opts.get("mu_init", 0.0))
        cfg.bound_push = float(opts.get("warm_start_bound_push", 0.0))
        cfg.device = int(device)
        self._device = int(device)
        cfg.threads = int((solver_opts or {}).get("b200", {}).get("threads", 0))
        self._lib = _cabi.lib()
        h = ctypes.c_void_p()
        _cabi.check(self._lib.bmpc_create(ctypes.byref(cfg), ctypes.byref(h)), "bmpc_create")
        self._h = h
        n, m, np_ = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _cabi.check(self._lib.bmpc_dims(h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(np_)), "bmpc_dims")
        self.N, self.nr_segs, self.dt = int(N), int(nr_segs), float(dt)
        self.n, self.m, self.np = n.value, m.value, np_.value
        self.tol = cfg.tol
        self._stats = {"iter_count": 0, "success": False, "return_status": "unset"}
        self._ws = None
        self._bounds = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.bmpc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- constant bound vectors (casadi_ocp_formulation.py:384-391)
    def bounds(self):
        lbx, ubx, lbg, ubg = np.empty(self.n), np.empty(self.n), np.empty(self.m), np.empty(self.m)
        dp = _cabi.c_double_p
        _cabi.check(self._lib.bmpc_bounds(self._h, lbx.ctypes.data_as(dp), ubx.ctypes.data_as(dp),
                                          lbg.ctypes.data_as(dp), ubg.ctypes.data_as(dp)), "bmpc_bounds")
        return lbx, ubx, lbg, ubg

    # ---- CasADi-like single solve (BoundMPC.py:446-453)
    def set_stats(self, iters, status, kkt=None):
        self._stats = {"iter_count": int(iters), "success": int(status) == 0, "return_status": RETURN_STATUS.get(int(status), "Internal_Error")}
        if kkt is not None:
            self._stats["kkt_error"] = float(kkt)

    def _check_bounds(self, **given):
        """The bound vectors are constants of the handle (bmpc_bounds = what setup_optimization_problem returned); a caller
        that passes different ones would silently get the wrong problem solved, so that is an error."""
        if self._bounds is None:
            self._bounds = dict(zip(("lbx", "ubx", "lbg", "ubg"), self.bounds()))
        for k, v in given.items():
            if v is None:
                continue
            a = np.asarray(v, float).ravel()
            if a.shape != self._bounds[k].shape or not np.array_equal(a, self._bounds[k]):
                raise ValueError(f"{k} differs from the bounds this solver was built with (setup_optimization_problem); "
                                 "the B200 solver has them compiled into the handle")

    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, p=None, lam_g0=None, lam_x0=None):
        """The CasADi call of BoundMPC.py:446-453.  `lam_g0` / `lam_x0` are accepted for signature compatibility and NOT used:
        like the reference's own call (which leaves them commented out, :451-452) every solve starts from zero equality
        multipliers and centred bound multipliers; passing non-zero values warns once (see note_dual_start)."""
        self._check_bounds(lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
        note_dual_start(lam_g0, lam_x0)
        out = self.solve_batch(np.asarray(x0, float).reshape(1, -1), np.asarray(p, float).reshape(1, -1))
        st = int(out["status"][0])
        self._stats = {"iter_count": int(out["iters"][0]), "success": st == 0,
                       "return_status": RETURN_STATUS.get(st, "Internal_Error"), "kkt_error": float(out["kkt"][0])}
        return {"x": out["x"][0], "g": out["g"][0], "lam_g": out["lam_g"][0], "lam_x": out["lam_x"][0],
                "f": float(out["f"][0])}

    def stats(self):
        return dict(self._stats)

    def generate_dependencies(self, *args, **kwargs):
        """BoundMPC.py:155-157 calls this when params.build is True; nothing to generate."""
        return None

    def launch_count(self):
        return int(self._lib.bmpc_launch_count(self._h))

    def launch_shape(self):
        """threads per CTA, resident CTAs per SM, shared memory per CTA [B], SM count of the solver kernel."""
        v = [ctypes.c_int32() for _ in range(4)]
        _cabi.check(self._lib.bmpc_launch_shape(self._h, *[ctypes.byref(x) for x in v]), "bmpc_launch_shape")
        return {"threads": v[0].value, "ctas_per_sm": v[1].value, "smem_bytes": v[2].value, "sms": v[3].value}

    def fp64_peak(self, kind=0):
        """Measured FP64 rate of the device in FLOP/s: kind 0 = DFMA loop, 1 = DMMA (mma.sync m8n8k4) loop."""
        v = ctypes.c_double()
        _cabi.check(self._lib.bmpc_fp64_peak(self._h, int(kind), ctypes.byref(v)), "bmpc_fp64_peak")
        return v.value

    # ---- batched entry
    def solve_batch(self, x0, p, out=None):
        """x0 [B, n], p [B, np] -> dict of x, g, lam_g, lam_x [B, .], f, kkt [B], iters, status [B].
        numpy inputs take the host-pointer entry (copies inside); torch CUDA tensors are solved in
        place on the current stream with no host synchronisation."""
        if isinstance(x0, np.ndarray):
            return self._solve_host(x0, p, out)
        return self._solve_device(x0, p, out)

    def _solve_host(self, x0, p, out=None):
        x0 = np.ascontiguousarray(x0, np.float64)
        p = np.ascontiguousarray(p, np.float64)
        B = x0.shape[0]
        if x0.shape != (B, self.n) or p.shape != (B, self.np):
            raise ValueError(f"expected x0 [{B},{self.n}] and p [{B},{self.np}], got {x0.shape}, {p.shape}")
        o = out or {}
        x = o.get("x") if "x" in o else np.empty((B, self.n))
        g = o.get("g") if "g" in o else np.empty((B, self.m))
        lg = o.get("lam_g") if "lam_g" in o else np.empty((B, self.m))
        lx = o.get("lam_x") if "lam_x" in o else np.empty((B, self.n))
        f = o.get("f") if "f" in o else np.empty(B)
        kkt = o.get("kkt") if "kkt" in o else np.empty(B)
        it = o.get("iters") if "iters" in o else np.empty(B, np.int32)
        st = o.get("status") if "status" in o else np.empty(B, np.int32)
        for name, a, shape, dt in (("x", x, (B, self.n), np.float64), ("g", g, (B, self.m), np.float64), ("lam_g", lg, (B, self.m), np.float64),
                                   ("lam_x", lx, (B, self.n), np.float64), ("f", f, (B,), np.float64), ("kkt", kkt, (B,), np.float64),
                                   ("iters", it, (B,), np.int32), ("status", st, (B,), np.int32)):
            if not isinstance(a, np.ndarray) or a.dtype != dt or a.shape != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"solve_batch: out['{name}'] must be a C-contiguous {np.dtype(dt).name} array of shape {shape}")
        P = _cabi.ptr
        _cabi.check(self._lib.bmpc_solve_batch_host(self._h, B, P(x0), P(p), P(x), P(g), P(lg), P(lx), P(f), P(it), P(st), P(kkt)),
                    "bmpc_solve_batch_host")
        return {"x": x, "g": g, "lam_g": lg, "lam_x": lx, "f": f, "kkt": kkt, "iters": it, "status": st}

    def workspace_bytes(self, B):
        sz = ctypes.c_size_t()
        _cabi.check(self._lib.bmpc_workspace_bytes(self._h, int(B), ctypes.byref(sz)), "bmpc_workspace_bytes")
        return sz.value

    def _solve_device(self, x0, p, out=None):
        import torch
        if not (x0.is_cuda and p.is_cuda and x0.dtype == torch.float64 and p.dtype == torch.float64):
            raise ValueError("solve_batch: tensors must be float64 CUDA tensors")
        x0, p = x0.contiguous(), p.contiguous()
        B = x0.shape[0]
        if tuple(x0.shape) != (B, self.n) or tuple(p.shape) != (B, self.np):
            raise ValueError(f"expected x0 [{B},{self.n}] and p [{B},{self.np}]")
        dev = x0.device
        o = out or {}
        mk = lambda k, shape, dt=torch.float64: o[k] if k in o else torch.empty(shape, dtype=dt, device=dev)
        x, g, lg, lx = mk("x", (B, self.n)), mk("g", (B, self.m)), mk("lam_g", (B, self.m)), mk("lam_x", (B, self.n))
        f, kkt = mk("f", (B,)), mk("kkt", (B,))
        it, st = mk("iters", (B,), torch.int32), mk("status", (B,), torch.int32)
        need = self.workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        V = ctypes.c_void_p
        _cabi.check(self._lib.bmpc_solve_batch(self._h, B, V(x0.data_ptr()), V(p.data_ptr()), V(x.data_ptr()), V(g.data_ptr()),
                                               V(lg.data_ptr()), V(lx.data_ptr()), V(f.data_ptr()), V(it.data_ptr()), V(st.data_ptr()),
                                               V(kkt.data_ptr()), V(self._ws.data_ptr()), V(stream)), "bmpc_solve_batch")
        return {"x": x, "g": g, "lam_g": lg, "lam_x": lx, "f": f, "kkt": kkt, "iters": it, "status": st}

    # ---- batched parameter builder (pre-solve half of BoundMPC.step)
    PS_SIZE, PT_ROW = 76, 41

    def prepare_batch(self, tables, path_id, sector, state, prev_x, out=None):
        """tables [P, J, 41], path_id [B] int32, sector [B] int32, state [B, 76], prev_x [B, n] ->
        dict x0 [B, n], p [B, np], sector [B] (advanced).  numpy inputs take the host-pointer entry; torch CUDA
        tensors are processed on the current stream (sector is updated in place)."""
        if isinstance(state, np.ndarray):
            tables = np.ascontiguousarray(tables, np.float64)
            state = np.ascontiguousarray(state, np.float64)
            prev_x = np.ascontiguousarray(prev_x, np.float64)
            pid = np.ascontiguousarray(path_id, np.int32)
            sec = np.ascontiguousarray(sector, np.int32).copy()
            B = state.shape[0]
            if tables.ndim != 3 or tables.shape[2] != self.PT_ROW or state.shape != (B, self.PS_SIZE) or prev_x.shape != (B, self.n):
                raise ValueError("prepare_batch: unexpected array shapes")
            o = out or {}
            x0 = o.get("x0") if "x0" in o else np.empty((B, self.n))
            p = o.get("p") if "p" in o else np.empty((B, self.np))
            P = _cabi.ptr
            _cabi.check(self._lib.bmpc_prepare_batch_host(self._h, B, P(tables), tables.shape[0], tables.shape[1], P(pid), P(sec),
                                                          P(state), P(prev_x), P(x0), P(p)), "bmpc_prepare_batch_host")
            return {"x0": x0, "p": p, "sector": sec}
        import torch
        B = state.shape[0]
        dev = state.device
        o = out or {}
        x0 = o["x0"] if "x0" in o else torch.empty((B, self.n), dtype=torch.float64, device=dev)
        p = o["p"] if "p" in o else torch.empty((B, self.np), dtype=torch.float64, device=dev)
        for t in (tables, state, prev_x, path_id, sector):
            if not (t.is_cuda and t.is_contiguous()):
                raise ValueError("prepare_batch: tensors must be contiguous CUDA tensors")
        V = ctypes.c_void_p
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self._lib.bmpc_prepare_batch(self._h, B, V(tables.data_ptr()), int(tables.shape[0]), int(tables.shape[1]),
                                                 V(path_id.data_ptr()), V(sector.data_ptr()), V(state.data_ptr()), V(prev_x.data_ptr()),
                                                 V(x0.data_ptr()), V(p.data_ptr()), V(stream)), "bmpc_prepare_batch")
        return {"x0": x0, "p": p, "sector": sector}

    # ---- batched post-processing (compute_return_data of BoundMPC.step)
    TR_ROW = 42
    TRAJ_KEYS = {"p": (0, 6), "v": (6, 12), "a": (12, 18), "q": (18, 25), "dq": (25, 32), "ddq": (32, 39), "phi": (39, 40),
                 "dphi": (40, 41), "ddphi": (41, 42)}

    def post_batch(self, tables, path_id, sector, state, w, error_count=None, out=None):
        """tables [P, J, 41], path_id / sector [B] int32 (sector: output of prepare_batch), state [B, 76] (of this step),
        w [B, n] (trajectory kept by the controller), error_count [B] int32 or None -> dict traj [B, N, 42] (columns
        TRAJ_KEYS) and state [B, 76] of the next step.  numpy -> host-pointer entry, torch CUDA tensors -> current stream."""
        if isinstance(state, np.ndarray):
            tables = np.ascontiguousarray(tables, np.float64)
            state = np.ascontiguousarray(state, np.float64)
            w = np.ascontiguousarray(w, np.float64)
            pid = np.ascontiguousarray(path_id, np.int32)
            sec = np.ascontiguousarray(sector, np.int32)
            ec = None if error_count is None else np.ascontiguousarray(error_count, np.int32)
            B = state.shape[0]
            if tables.ndim != 3 or tables.shape[2] != self.PT_ROW or state.shape != (B, self.PS_SIZE) or w.shape != (B, self.n):
                raise ValueError("post_batch: unexpected array shapes")
            o = out or {}
            traj = o.get("traj") if "traj" in o else np.empty((B, self.N, self.TR_ROW))
            so = o.get("state") if "state" in o else np.empty((B, self.PS_SIZE))
            P = _cabi.ptr
            _cabi.check(self._lib.bmpc_post_batch_host(self._h, B, P(tables), tables.shape[0], tables.shape[1], P(pid), P(sec), P(state),
                                                       P(w), P(ec), P(traj), P(so)), "bmpc_post_batch_host")
            return {"traj": traj, "state": so}
        import torch
        B, dev = state.shape[0], state.device
        o = out or {}
        traj = o["traj"] if "traj" in o else torch.empty((B, self.N, self.TR_ROW), dtype=torch.float64, device=dev)
        so = o["state"] if "state" in o else torch.empty((B, self.PS_SIZE), dtype=torch.float64, device=dev)
        for t in (tables, state, w, path_id, sector):
            if not (t.is_cuda and t.is_contiguous()):
                raise ValueError("post_batch: tensors must be contiguous CUDA tensors")
        V = ctypes.c_void_p
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self._lib.bmpc_post_batch(self._h, B, V(tables.data_ptr()), int(tables.shape[0]), int(tables.shape[1]),
                                              V(path_id.data_ptr()), V(sector.data_ptr()), V(state.data_ptr()), V(w.data_ptr()),
                                              V(error_count.data_ptr()) if error_count is not None else None, V(traj.data_ptr()),
                                              V(so.data_ptr()), V(stream)), "bmpc_post_batch")
        return {"traj": traj, "state": so}

    RF_ROW, ER_ROW = 55, 33
    ERR_KEYS = ("e_p", "de_p", "e_p_par", "e_p_orth", "de_p_par", "de_p_orth", "e_r", "de_r", "e_r_par", "e_r_orth1", "e_r_orth2")

    def post_log_batch(self, tables, path_id, sector, state, p, w, error_count=None):
        """post_batch plus the logging branch (ref_data / err_data of BoundMPC.step).  numpy or torch CUDA inputs; returns
        dict traj [B, N, 42], state [B, 76], ref [B, N, 55], err [B, N, 33] (err columns: ERR_KEYS, 3 each) of the same kind."""
        import torch
        host = isinstance(state, np.ndarray)
        dev = (torch.device("cuda", self._device) if self._device >= 0 else torch.device("cuda")) if host else state.device
        T = (lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)) if host else (lambda a, dt=None: a)
        tables, state, p, w = T(tables, np.float64), T(state, np.float64), T(p, np.float64), T(w, np.float64)
        path_id, sector = T(path_id, np.int32), T(sector, np.int32)
        ec = None if error_count is None else T(error_count, np.int32)
        B = state.shape[0]
        mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
        traj, so, ref, err = mk(B, self.N, self.TR_ROW), mk(B, self.PS_SIZE), mk(B, self.N, self.RF_ROW), mk(B, self.N, self.ER_ROW)
        V = ctypes.c_void_p
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self._lib.bmpc_post_log_batch(self._h, B, V(tables.data_ptr()), int(tables.shape[0]), int(tables.shape[1]),
                                                  V(path_id.data_ptr()), V(sector.data_ptr()), V(state.data_ptr()), V(p.data_ptr()),
                                                  V(w.data_ptr()), V(ec.data_ptr()) if ec is not None else None, V(traj.data_ptr()),
                                                  V(so.data_ptr()), V(ref.data_ptr()), V(err.data_ptr()), V(stream)), "bmpc_post_log_batch")
        out = {"traj": traj, "state": so, "ref": ref, "err": err}
        return {k: v.cpu().numpy() for k, v in out.items()} if host else out

    def update_batch(self, tables, path_phi_max, new_path, cart, state, sector, path_id):
        """BoundMPC.update for a batch on the device (torch CUDA tensors): controllers with new_path[b] >= 0 switch to that
        path; state [B, 76], sector [B], path_id [B] are updated in place.  cart [B, 24] = measured pose, velocity,
        acceleration, jerk (Cartesian)."""
        import torch
        V = ctypes.c_void_p
        stream = torch.cuda.current_stream(state.device).cuda_stream
        _cabi.check(self._lib.bmpc_update_batch(self._h, int(state.shape[0]), V(tables.data_ptr()), int(tables.shape[0]), int(tables.shape[1]),
                                                V(path_phi_max.data_ptr()), V(new_path.data_ptr()), V(cart.data_ptr()), V(state.data_ptr()),
                                                V(sector.data_ptr()), V(path_id.data_ptr()), V(stream)), "bmpc_update_batch")

    def finish_batch(self, tables, path_id, sector, state, sol, prev_x, error_count, advance=True, out=None):
        """Second half of BoundMPC.step for a batch on the device (torch CUDA tensors): accept / reject every solve of
        `sol` (dict of solve_batch), keep the previous solution where rejected, post-process and (advance=True) move the
        robot by one sample.  prev_x [B, n] and error_count [B] int32 are updated in place.  -> dict traj, state."""
        import torch
        B, dev = state.shape[0], state.device
        o = out or {}
        traj = o["traj"] if "traj" in o else torch.empty((B, self.N, self.TR_ROW), dtype=torch.float64, device=dev)
        so = o["state"] if "state" in o else torch.empty((B, self.PS_SIZE), dtype=torch.float64, device=dev)
        V = ctypes.c_void_p
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self._lib.bmpc_finish_batch(self._h, B, V(tables.data_ptr()), int(tables.shape[0]), int(tables.shape[1]),
                                                V(path_id.data_ptr()), V(sector.data_ptr()), V(state.data_ptr()), V(sol["x"].data_ptr()),
                                                V(sol["g"].data_ptr()), V(sol["status"].data_ptr()), V(prev_x.data_ptr()),
                                                V(error_count.data_ptr()), V(traj.data_ptr()), V(so.data_ptr()), int(bool(advance)),
                                                V(stream)), "bmpc_finish_batch")
        return {"traj": traj, "state": so}


    def mpc_step_host(self, tables, path_id, sector, state, prev_x, error_count, want_log=True):
        """One whole MPC step (`BoundMPC.step`: prepare -> solve -> accept / fallback -> post-processing [+ logging branch]) for B
        controllers whose state lives on the host, in ONE library call (`bmpc_mpc_step_batch_host`).  numpy arrays: tables
        [P, J, 41], path_id / sector / error_count [B] int32, state [B, 76], prev_x [B, n].  Returns dict x, traj [B, N, 42],
        state [B, 76] (next step), ref [B, N, 55] / err [B, N, 33] (or None), iters, status, and the updated sector, prev,
        error_count."""
        tables = np.ascontiguousarray(tables, np.float64)
        state = np.ascontiguousarray(state, np.float64)
        B = state.shape[0]
        pid = np.ascontiguousarray(path_id, np.int32)
        sec = np.array(sector, np.int32).reshape(B).copy()
        ec = np.array(error_count, np.int32).reshape(B).copy()
        prev = np.array(prev_x, np.float64).reshape(B, self.n).copy()
        if tables.ndim != 3 or tables.shape[2] != self.PT_ROW or state.shape != (B, self.PS_SIZE):
            raise ValueError("mpc_step_host: unexpected array shapes")
        x, traj, so = np.empty((B, self.n)), np.empty((B, self.N, self.TR_ROW)), np.empty((B, self.PS_SIZE))
        ref = np.empty((B, self.N, self.RF_ROW)) if want_log else None
        err = np.empty((B, self.N, self.ER_ROW)) if want_log else None
        it, st = np.empty(B, np.int32), np.empty(B, np.int32)
        P = _cabi.ptr
        _cabi.check(self._lib.bmpc_mpc_step_batch_host(self._h, B, P(tables), tables.shape[0], tables.shape[1], P(pid), P(sec), P(state),
                                                       P(prev), P(ec), P(x), P(traj), P(so), P(ref), P(err), P(it), P(st)),
                    "bmpc_mpc_step_batch_host")
        return {"x": x, "traj": traj, "state": so, "ref": ref, "err": err, "iters": it, "status": st, "sector": sec, "prev": prev,
                "error_count": ec}

    # ---- NLP function evaluation for parity tests (nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l)
    def eval_batch(self, x, p, lam=None, want_jac=True, want_hess=True):
        x = np.ascontiguousarray(np.atleast_2d(x), np.float64)
        p = np.ascontiguousarray(np.atleast_2d(p), np.float64)
        B = x.shape[0]
        lam = None if lam is None else np.ascontiguousarray(np.atleast_2d(lam), np.float64)
        f, g, d, grad = np.empty(B), np.empty((B, self.m)), np.empty((B, 12 * self.N)), np.empty((B, self.n))
        jac = np.empty((B, 48 * self.N, self.n)) if want_jac else None
        hess = np.empty((B, self.n, self.n)) if want_hess else None
        P = _cabi.ptr
        _cabi.check(self._lib.bmpc_eval_batch_host(self._h, B, P(x), P(p), P(lam), P(f), P(g), P(d), P(grad), P(jac), P(hess)),
                    "bmpc_eval_batch_host")
        return {"f": f, "g": g, "d": d, "grad": grad, "jac": jac, "hess": hess}

    def kkt_step_batch(self, x, y, s, zs, zL, zU, p, mu, delta_w=0.0):
        """One Newton step of the interior-point iteration at given primal-dual points (parity tests of the Riccati
        sweep against a dense KKT solve).  Returns dx [B, n], ynew [B, 36 N], ok [B]."""
        p = np.ascontiguousarray(np.atleast_2d(p), np.float64)
        v = np.ascontiguousarray(np.concatenate([np.atleast_2d(a) for a in (x, y, s, zs, zL, zU)], axis=1), np.float64)
        B = v.shape[0]
        assert v.shape[1] == 3 * self.n + 60 * self.N and p.shape == (B, self.np)
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, np.float64), (B,)))
        dw = np.ascontiguousarray(np.broadcast_to(np.asarray(delta_w, np.float64), (B,)))
        dx, ynew, ok = np.empty((B, self.n)), np.empty((B, 36 * self.N)), np.empty(B, np.int32)
        P = _cabi.ptr
        _cabi.check(self._lib.bmpc_kkt_step_batch_host(self._h, B, P(v), P(p), P(mu), P(dw), P(dx), P(ynew), P(ok)),
                    "bmpc_kkt_step_batch_host")
        return {"dx": dx, "ynew": ynew, "ok": ok}


def setup_optimization_problem(N, nr_joints, nr_segs, dt, u_min, u_max, ut_min, ut_max, q_lim_lower, q_lim_upper,
                               dq_lim_lower, dq_lim_upper, solver_opts):
    """Same signature / return tuple as casadi_ocp_formulation.py:9-11,391."""
    if nr_joints != 7:
        raise ValueError("the kinematic model is the 7-joint iiwa14 chain")
    solver = BatchSolver(N, nr_segs, dt, u_min, u_max, ut_min, ut_max, q_lim_lower, q_lim_upper, dq_lim_lower,
                         dq_lim_upper, solver_opts)
    lbx, ubx, lbg, ubg = solver.bounds()
    return solver, lbx.tolist(), ubx.tolist(), lbg.tolist(), ubg.tolist(), _g_names(N)


def default_solver(N=10, nr_segs=4, dt=0.1, solver_opts=None, device=-1):
    """Solver for the reference's robot limits (RobotModel.py:20-43)."""
    from . import robot_model as rm
    return BatchSolver(N, nr_segs, dt, rm.U_MIN, rm.U_MAX, rm.U_MIN, rm.U_MAX, rm.Q_LIM_LOWER, rm.Q_LIM_UPPER,
                       rm.DQ_LIM_LOWER, rm.DQ_LIM_UPPER, solver_opts, device)
