"""Reference-executing harness.  TEST INFRASTRUCTURE ONLY (needs /root/reference).

Runs the reference's *own* Python (`casadi_ocp_formulation.setup_optimization_problem`,
`BoundMPC.__init__/.step`, `RobotModel`, `ReferencePath`) through the numeric
`casadi` stand-in in this directory.  Used to
  * evaluate the reference NLP functions f(w,p), g(w,p) at arbitrary points,
    with complex-step gradient / Jacobian,
  * drive the unmodified reference `BoundMPC` class with any solver backend,
  * generate the golden fixtures in tests/golden/ (see make_golden.py).
Nothing in the product imports this.
"""
import os
import sys
import types
import numpy as np

REF_ROOT = os.environ.get("BOUNDMPC_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "bound_mpc")
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REF_PKG, "bound_mpc", "BoundMPC"))


def _activate():
    """Put the stand-in modules ahead of everything and the reference on the path."""
    if not available():
        raise RuntimeError("reference tree not available")
    for p in (REF_PKG, HERE):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import casadi  # noqa: F401  (the stand-in)
    assert casadi.__file__.startswith(HERE), casadi.__file__
    return casadi


class _TagProvider:
    """First pass: every symbol entry gets a unique integer tag as its value."""

    def __init__(self):
        self.n = 0
        self.syms = []  # (name, shape, first_tag)

    def sym(self, name, shape):
        from casadi import M
        r, c = shape
        tags = np.arange(self.n, self.n + r * c, dtype=float)
        self.syms.append((name, shape, self.n))
        self.n += r * c
        # column-major fill, like CasADi's dense symbols
        return M(tags.reshape(c, r).T.reshape(r, c, 1).copy())


class _EvalProvider:
    """Later passes: symbol k (creation order) gets values gathered from (x, p)."""

    def __init__(self, src, syms):
        self.src = src  # [n_tags, B] values
        self.syms = syms
        self.k = 0

    def sym(self, name, shape):
        from casadi import M
        nm, shp, first = self.syms[self.k]
        assert nm == name and tuple(shp) == tuple(shape), (nm, name)
        self.k += 1
        r, c = shape
        vals = self.src[first:first + r * c]  # [r*c, B] column-major
        return M(np.swapaxes(vals.reshape(c, r, -1), 0, 1).copy())


class RefNLP:
    """f(w,p), g(w,p) of the reference NLP, evaluated by the reference's own code."""

    def __init__(self, N=10, nr_segs=4, dt=0.1):
        ca = _activate()
        from bound_mpc.BoundMPC.casadi_ocp_formulation import setup_optimization_problem
        from bound_mpc.RobotModel import RobotModel
        self._setup = setup_optimization_problem
        rm = RobotModel()
        lim = rm.get_robot_limits()
        self.args = (N, 7, nr_segs, dt, lim[7], lim[6], lim[7], lim[6],
                     lim[1], lim[0], lim[3], lim[2], {})
        self.N, self.nr_segs, self.dt = N, nr_segs, dt
        tp = _TagProvider()
        ca.set_provider(tp)
        solver, lbu, ubu, lbg, ubg, g_names = self._setup(*self.args)
        self.syms = tp.syms
        self.ntags = tp.n
        xt = np.rint(solver.prob['x'].a[:, 0, 0].real).astype(int)
        pt = np.rint(solver.prob['p'].a[:, 0, 0].real).astype(int)
        self.x_tags, self.p_tags = xt, pt
        self.n, self.np_, self.m = len(xt), len(pt), solver.prob['g'].a.shape[0]
        self.lbx, self.ubx = np.array(lbu, float), np.array(ubu, float)
        self.lbg, self.ubg = np.array(lbg, float), np.array(ubg, float)
        self.g_names = g_names
        assert len(set(xt.tolist()) | set(pt.tolist())) == self.ntags

    def eval(self, x, p):
        """x [B,n], p [B,np] (real or complex) -> f [B], g [B,m]."""
        ca = _activate()
        x = np.atleast_2d(x)
        p = np.atleast_2d(p)
        B = max(x.shape[0], p.shape[0])
        dt_ = complex if (np.iscomplexobj(x) or np.iscomplexobj(p)) else float
        src = np.zeros((self.ntags, B), dtype=dt_)
        src[self.x_tags] = np.broadcast_to(x, (B, self.n)).T
        src[self.p_tags] = np.broadcast_to(p, (B, self.np_)).T
        ca.set_provider(_EvalProvider(src, self.syms))
        solver, *_ = self._setup(*self.args)
        f = np.broadcast_to(solver.prob['f'].a[0, 0, :], (B,)).copy()
        g = np.broadcast_to(solver.prob['g'].a[:, 0, :], (self.m, B)).T.copy()
        return f, g

    def eval_derivs(self, x, p, h=1e-30):
        """Complex-step gradient of f [n] and Jacobian of g [m,n] at one point."""
        x = np.asarray(x, float).ravel()
        X = np.tile(x.astype(complex), (self.n + 1, 1))
        X[np.arange(self.n), np.arange(self.n)] += 1j * h
        f, g = self.eval(X, np.asarray(p, float).reshape(1, -1))
        grad = f[:self.n].imag / h
        jac = (g[:self.n].imag / h).T
        return f[self.n].real, g[self.n].real, grad, jac

    def lag_grad(self, x, p, lam, h=1e-30):
        """Complex-step gradient of f + lam.g (one batched pass)."""
        _, _, grad, jac = self.eval_derivs(x, p, h)
        return grad + jac.T @ lam

    def lag_hess(self, x, p, lam, eps=1e-6):
        """Central differences of the complex-step Lagrangian gradient, coloured with
        period 3*44 (block-tridiagonal structure)."""
        x = np.asarray(x, float).ravel()
        n = self.n
        per = 132
        H = np.zeros((n, n))
        for c in range(min(per, n)):
            d = np.zeros(n)
            d[c::per] = eps
            gp = self.lag_grad(x + d, p, lam)
            gm = self.lag_grad(x - d, p, lam)
            col = (gp - gm) / (2 * eps)
            for j in range(c, n, per):
                blk = j // 44
                lo, hi = max(0, (blk - 1) * 44), min(n, (blk + 2) * 44)
                H[lo:hi, j] = col[lo:hi]
        return 0.5 * (H + H.T)


class _Params:
    def __init__(self, n=10, nr_segs=4, dt=0.1, weights=None, real_time=True):
        self.n, self.nr_segs, self.dt = n, nr_segs, dt
        self.weights = list(weights)
        self.build = True
        self.simulate = False
        self.experiment = False
        self.use_acados = False
        self.learning_based = False
        self.real_time = real_time


def make_reference_mpc(scn, backend, n=10, real_time=True):
    """Construct the UNMODIFIED reference BoundMPC for a scenario dict (see
    boundmpc_b200.scenarios) with `backend` as the solver behind `casadi.nlpsol`."""
    import copy
    ca = _activate()
    ca.set_provider(_TagProvider())
    ca.set_solver_backend(backend)
    from bound_mpc.BoundMPC.BoundMPC import BoundMPC
    s = copy.deepcopy(scn)
    params = _Params(n=n, nr_segs=s['nr_segs'], dt=s['dt'], weights=s['weights'],
                     real_time=real_time)
    # The runner -> create_traj_msg -> node chain swaps upper/lower twice
    # (util_functions.py:58-70, bound_mpc_node.py:57-58, ReferencePath.py:52-55); the net effect
    # is pos_lim = [lower, upper] as ReferencePath reads it.
    mpc = BoundMPC(s['p_via'], s['r_via'], [s['p_lower'], s['p_upper']],
                   [s['r_lower'], s['r_upper']], s['bp1'], s['br1'], s['s'],
                   s['e_p_min'], s['e_r_min'], s['e_p_max'], s['e_r_max'],
                   p0=np.array(s['p0fk']), params=params)
    return mpc


def reference_robot_model():
    _activate()
    from bound_mpc.RobotModel import RobotModel
    return RobotModel()


def reference_integrate_joint():
    _activate()
    from bound_mpc.utils.util_functions import integrate_joint
    return integrate_joint
