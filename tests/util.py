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
