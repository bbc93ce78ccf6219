"""BatchSolver look-alike on top of the host-emulation build.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
from . import emu
from oracle import oracle as O


class EmuSolver:
    def __init__(self, N=10, S=4, dt=0.1, tol=1e-9):
        self.N, self.nr_segs, self.dt, self.tol = N, S, dt, tol
        self.n, self.m, self.np = 44 * N, 43 * N, 141 + 91 * S
        self._stats = {}

    def bounds(self):
        return O.bounds(self.N, self.nr_segs, self.dt)

    def solve_batch(self, x0, p, out=None):
        return emu.solve(x0, p, self.N, self.nr_segs, self.dt, self.tol)

    def eval_batch(self, x, p, lam=None, want_jac=True, want_hess=True):
        return emu.evaluate(x, p, lam, self.N, self.nr_segs, self.dt, want_jac, want_hess)

    def __call__(self, x0=None, p=None, **kw):
        r = self.solve_batch(np.asarray(x0, float).reshape(1, -1), np.asarray(p, float).reshape(1, -1))
        self._stats = dict(iter_count=int(r['iters'][0]), success=int(r['status'][0]) == 0, return_status=str(r['status'][0]))
        return dict(x=r['x'][0], g=r['g'][0], lam_g=r['lam_g'][0], lam_x=r['lam_x'][0], f=float(r['f'][0]))

    def stats(self):
        return dict(self._stats)

    def generate_dependencies(self, *a, **k):
        return None

    # serial form of bmpc_mpc_step_batch_host (k_prepare -> k_solve -> k_finish + logging branch), same return dict as
    # BatchSolver.mpc_step_host: lets the CPU suite drive BoundMPC's device step
    def mpc_step_host(self, tables, path_id, sector, state, prev_x, error_count, want_log=True):
        state = np.ascontiguousarray(np.atleast_2d(state), float)
        B = state.shape[0]
        prev = np.array(prev_x, float).reshape(B, self.n).copy()
        ec = np.array(error_count, np.int32).reshape(B).copy()
        x0, p, sec = emu.prepare(tables, path_id, sector, state, prev, self.N, self.nr_segs)
        r = emu.solve(x0, p, self.N, self.nr_segs, self.dt, self.tol)
        kept_prev = prev.copy()
        traj, so, prev, ec_out = emu.finish(tables, path_id, sec, state, r["x"], r["g"], r["status"], prev, ec, advance=False, N=self.N,
                                            S=self.nr_segs, dt=self.dt)
        ref = err = None
        if want_log:
            w = np.where((ec_out == 0)[:, None], r["x"], kept_prev)
            _, _, ref, err = emu.post_log(tables, path_id, sec, state, p, w, ec_out, self.N, self.nr_segs, self.dt)
        return {"x": r["x"], "traj": traj, "state": so, "ref": ref, "err": err, "iters": r["iters"], "status": r["status"], "sector": sec,
                "prev": prev, "error_count": ec_out}

    def set_stats(self, iters, status, kkt=None):
        self._stats = dict(iter_count=int(iters), success=int(status) == 0, return_status=str(int(status)))
