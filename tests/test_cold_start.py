"""Cold starts in the middle of a path (literal SURVEY 8d workload: every odd instance; BoundMPC.py:316-321 used away from
step 0).  The start has its path parameter at 0 while the robot is metres along the path: equality rows violated by 0.75.
Ipopt recovers from such starts in its restoration phase; this solver replaces the states of such a start by the
roll-out of its own inputs (`Config::rollout_thr`, bmpc_ipm.cuh) and then converges like from a warm start."""
import numpy as np
import pytest

from tests.util import load
from tests.emu import emu
from oracle import oracle as O

COLD, WARM, INFEASIBLE = (0, 1, 2, 3), (4, 5), (6, 7)      # rows of tests/golden/spec_cold.npz


def _kkt_point(r, p):
    """feasible, and stationary with the oracle's dense derivatives (independent of the iteration)"""
    g = r["g"].reshape(10, 43)
    assert np.abs(g[:, :36]).max() < 1e-8 and g[:, 36:].max() < 1e-8
    grad, jac, _ = O.derivs(r["x"], p, r["lam_g"])
    res = grad + jac.T @ r["lam_g"] + r["lam_x"]
    assert np.abs(res).max() < 1e-6 * max(1.0, np.abs(r["lam_g"]).max())


def test_cold_starts_converge_to_kkt_points():
    S = load("spec_cold.npz")
    for j in COLD + WARM:
        c0 = np.abs(O.eval_fg(S["x0"][j], S["p"][j])[1].reshape(10, 43)[:, :36]).max()
        assert (c0 > 0.5) == (j in COLD)                     # only the cold starts trip the repair
        r = O.solve(S["x0"][j], S["p"][j], tol=1e-9)
        assert r["status"] == 0 and r["iters"] <= 40
        _kkt_point(r, S["p"][j])
    for j in INFEASIBLE:
        assert O.solve(S["x0"][j], S["p"][j], tol=1e-9)["status"] == 5


def test_host_build_matches_oracle_on_cold_starts():
    S = load("spec_cold.npz")
    r = emu.solve(S["x0"], S["p"], tol=1e-9)
    for j in range(8):
        ro = O.solve(S["x0"][j], S["p"][j], tol=1e-9)
        assert ro["status"] == r["status"][j] and abs(ro["iters"] - int(r["iters"][j])) <= 2
        if ro["status"] == 0:
            assert np.abs(r["x"][j] - ro["x"]).max() < 1e-6


@pytest.mark.gpu
def test_gpu_matches_oracle_on_cold_starts():
    from boundmpc_b200.ocp import default_solver
    S = load("spec_cold.npz")
    r = default_solver().solve_batch(S["x0"], S["p"])
    for j in range(8):
        ro = O.solve(S["x0"][j], S["p"][j], tol=1e-9)
        assert ro["status"] == r["status"][j] and abs(ro["iters"] - int(r["iters"][j])) <= 2
        if ro["status"] == 0:
            assert np.abs(r["x"][j] - ro["x"]).max() < 1e-6
