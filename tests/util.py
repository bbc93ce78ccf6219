"""Shared helpers of the test-suite."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def active_set(sol, N=10, tol=1e-6):
    """Active inequality rows and variable bounds of a solution dict (x, g): a row is active
    when it sits on its bound within the 1e-6 slack the reference itself uses in its
    feasibility check (BoundMPC.py:461-463)."""
    from boundmpc_b200 import robot_model as rm
    g = np.asarray(sol["g"]).reshape(N, 43)[:, 36:]
    rows = {(k, 36 + i) for k in range(N) for i in range(7) if g[k, i] > -tol}
    x = np.asarray(sol["x"]).reshape(N, 44)
    lb = np.full(44, -np.inf)
    ub = np.full(44, np.inf)
    lb[:8], ub[:8] = rm.U_MIN, rm.U_MAX
    lb[8:15], ub[8:15] = rm.Q_LIM_LOWER, rm.Q_LIM_UPPER
    lb[15:22], ub[15:22] = rm.DQ_LIM_LOWER, rm.DQ_LIM_UPPER
    lb[41] = 0.0
    bnds = {(k, i, -1) for k in range(N) for i in range(44) if x[k, i] - lb[i] < tol}
    bnds |= {(k, i, 1) for k in range(N) for i in range(44) if ub[i] - x[k, i] < tol}
    return rows, bnds


def rel_q_error(x, xref, N=10):
    """max over the joint trajectory of |q - q_ref| / max|q_ref| (the north-star criterion)."""
    a = np.asarray(x).reshape(N, 44)[:, 8:15]
    b = np.asarray(xref).reshape(N, 44)[:, 8:15]
    return np.abs(a - b).max() / np.abs(b).max()


def lam48_from_lam43(lam_g, N=10):
    """Equality multipliers of a reference-form lam_g with zero interval-row multipliers."""
    out = np.zeros((N, 48))
    out[:, :36] = np.asarray(lam_g).reshape(N, 43)[:, :36]
    return out.ravel()


def dense_kkt_step(ev, x, s, zs, zL, zU, lbx, ubx, mu, delta_w, N=10, inertia=False):
    """Newton step of the interior-point iteration from a DENSE solve of the assembled KKT system (numpy.linalg.solve on
    the (n + 36 N)-square matrix; SURVEY 4 "Riccati step vs dense solve").  `ev` = outputs of eval_batch for ONE instance at
    (x, lam = (y, z_s)): d [12 N], grad [n], jac [48 N, n], hess [n, n] (Hessian of f + y.c + z_s.d).  Slacks and bound
    multipliers are eliminated the way Ipopt does it (Waechter & Biegler 2006, eq. (13)):
        [H + Sigma_x + delta_w I + Jd^T Sigma_s Jd   Jc^T] [dx  ]     [grad - mu/(x-l) + mu/(u-x) + Jd^T (mu/s + Sigma_s (d+s))]
        [Jc                                          0   ] [ynew] = - [c                                                         ]
    Returns dx, ynew."""
    n = 44 * N
    jac = ev["jac"].reshape(N, 48, n)
    Jc, Jd = jac[:, :36].reshape(36 * N, n), jac[:, 36:].reshape(12 * N, n)
    H = ev["hess"].copy()
    gh = ev["grad"].copy()
    fl, fu = np.isfinite(lbx), np.isfinite(ubx)
    sig = np.zeros(n)
    sig[fl] += zL[fl] / (x[fl] - lbx[fl]); gh[fl] -= mu / (x[fl] - lbx[fl])
    sig[fu] += zU[fu] / (ubx[fu] - x[fu]); gh[fu] += mu / (ubx[fu] - x[fu])
    Sg = zs / s
    H += np.diag(sig + delta_w) + Jd.T @ (Sg[:, None] * Jd)
    gh += Jd.T @ (mu / s + Sg * (ev["d"] + s))
    c = ev["c"]
    K = np.block([[H, Jc.T], [Jc, np.zeros((36 * N, 36 * N))]])
    sol = np.linalg.solve(K, -np.concatenate([gh, c]))
    if inertia:
        return sol[:n], sol[n:], int((np.linalg.eigvalsh(K) < 0).sum())   # negative eigenvalues (36 N iff the reduced Hessian is PD)
    return sol[:n], sol[n:]


def interior_point(x0, lbx, ubx, d_of_x, rng, N=10, push=1e-3, mu=1e-2):
    """A strictly interior primal-dual point around x0 for the KKT-step tests: pushed x, slacks from d(x), random
    positive multipliers scattered around mu / slack, random equality multipliers."""
    x = np.clip(x0, np.where(np.isfinite(lbx), lbx + push, -np.inf), np.where(np.isfinite(ubx), ubx - push, np.inf))
    d = d_of_x(x)
    s = np.maximum(-d, push)
    zs = mu / s * rng.uniform(0.3, 3.0, s.shape)
    zL = np.where(np.isfinite(lbx), mu / np.maximum(x - lbx, push) * rng.uniform(0.3, 3.0, x.shape), 0.0)
    zU = np.where(np.isfinite(ubx), mu / np.maximum(ubx - x, push) * rng.uniform(0.3, 3.0, x.shape), 0.0)
    y = rng.normal(0.0, 1.0, 36 * N)
    return x, y, s, zs, zL, zU
