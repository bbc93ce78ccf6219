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
