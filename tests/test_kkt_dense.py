"""Riccati step vs a dense solve of the assembled KKT system (VERDICT r1 "missing 6").

The structure-exploiting KKT solve (csrc/bmpc_riccati.cuh, what MUMPS does for Ipopt) and the oracle's Riccati are the same
derivation, so agreement between them cannot reveal a shared mistake.  Here the Newton step comes from
numpy.linalg.solve on the full (n + 36 N)-square system assembled from the DENSE derivative exports (jac, hess), which
share nothing with the sweeps.  CPU: host build of the kernel source; GPU: through the C ABI."""
import numpy as np
import pytest

from tests.util import load, dense_kkt_step, interior_point
from tests.emu import emu
from oracle import oracle as O


def _cases(N=10):
    rng = np.random.default_rng(7)
    lbx, ubx, _, _ = O.bounds(N, 4, 0.1)
    out = []
    for scn in ("exp1", "exp2"):
        S = load(f"seq_{scn}.npz")
        for i in (0, 3, len(S["x0"]) - 1):
            # a point on the way (warm start perturbed) and the converged point with its active rows
            for base in (S["x0"][i] + rng.normal(0, 1e-3, 44 * N), S["x"][i]):
                out.append((base, S["p"][i], lbx, ubx, rng))
    return out


def _check(evaluate, kkt_step, tol):
    worst = 0.0
    seen = {True: 0, False: 0}
    for x0, p, lbx, ubx, rng in _cases():
        d_of_x = lambda x: evaluate(x, p, None, False, False)["d"][0]
        x, y, s, zs, zL, zU = interior_point(x0, lbx, ubx, d_of_x, rng)
        lam = np.concatenate([y.reshape(10, 36), zs.reshape(10, 12)], axis=1).ravel()
        ev = {k: (v[0] if v is not None else None) for k, v in evaluate(x, p, lam, True, True).items()}
        ev["c"] = ev["g"].reshape(10, 43)[:, :36].ravel()
        for mu, dw in ((1e-2, 0.0), (1e-5, 0.0), (1e-3, 1e-2), (1e-3, 10.0), (1e-3, 1e3)):
            r = kkt_step(x, y, s, zs, zL, zU, p, mu, dw)
            dx, yn, neg = dense_kkt_step(ev, x, s, zs, zL, zU, lbx, ubx, mu, dw, inertia=True)
            # inertia control: the sweep refuses the step exactly when the dense KKT matrix does not have the inertia
            # (n, 36 N, 0) Ipopt asks MUMPS for, i.e. when the reduced Hessian is not positive definite
            assert bool(r["ok"][0]) == (neg == 360), (mu, dw, neg)
            seen[bool(r["ok"][0])] += 1
            if not r["ok"][0]:
                continue
            e1 = np.abs(r["dx"][0] - dx).max() / max(1.0, np.abs(dx).max())
            e2 = np.abs(r["ynew"][0] - yn).max() / max(1.0, np.abs(yn).max())
            worst = max(worst, e1, e2)
            assert e1 < tol and e2 < tol, (mu, dw, e1, e2)
    assert seen[True] >= 20 and seen[False] >= 4, seen      # both branches of the inertia test were exercised
    return worst


def test_riccati_step_vs_dense_kkt_host_build():
    worst = _check(lambda x, p, lam, wj, wh: emu.evaluate(x, p, lam, want_jac=wj, want_hess=wh),
                   lambda *a: emu.kkt_step(*a), 1e-9)
    print("host build: worst relative deviation from the dense KKT solve", worst)


def test_dense_step_vanishes_at_the_kkt_point():
    """At a converged point the dense Newton step (numpy only, no sweep involved) is ~0: pins the sign conventions of the
    dense assembly itself against the golden KKT points (tests/golden, certified with reference-executed derivatives)."""
    lbx, ubx, _, _ = O.bounds(10, 4, 0.1)
    S = load("seq_exp1.npz")
    x, p = S["x"][0], S["p"][0]
    r = O.solve(S["x0"][0], p, tol=1e-10)
    assert r["status"] == 0 and np.abs(r["x"] - x).max() < 1e-5
    x = r["x"]
    mu = 1e-10
    d = emu.evaluate(x, p, None, want_jac=False, want_hess=False)["d"][0]
    lam_g = r["lam_g"].reshape(10, 43)
    y = lam_g[:, :36].ravel()
    dd = d.reshape(10, 12)
    s = np.maximum(-dd, 1e-9)
    zs = mu / s
    for k in range(10):      # active rows carry their multiplier (reference form -> interval form: z = 2 h lam on the active side)
        for q in range(2):
            if lam_g[k, 36 + q] > 1e-6: zs[k, q] = lam_g[k, 36 + q]; s[k, q] = mu / zs[k, q]
        for j in range(5):
            h = -0.5 * (dd[k, 2 + 2 * j] + dd[k, 3 + 2 * j])
            if lam_g[k, 38 + j] > 1e-6:
                side = int(dd[k, 3 + 2 * j] > dd[k, 2 + 2 * j])
                zs[k, 2 + 2 * j + side] = 2 * h * lam_g[k, 38 + j]; s[k, 2 + 2 * j + side] = mu / zs[k, 2 + 2 * j + side]
    s, zs = s.ravel(), zs.ravel()
    fl, fu = np.isfinite(lbx), np.isfinite(ubx)
    zL = np.where(fl, np.maximum(np.maximum(-r["lam_x"], 0), mu / np.maximum(x - lbx, 1e-9)), 0.0)
    zU = np.where(fu, np.maximum(np.maximum(r["lam_x"], 0), mu / np.maximum(ubx - x, 1e-9)), 0.0)
    lam = np.concatenate([y.reshape(10, 36), zs.reshape(10, 12)], axis=1).ravel()
    ev = {k: (v[0] if v is not None else None) for k, v in emu.evaluate(x, p, lam).items()}
    ev["c"] = ev["g"].reshape(10, 43)[:, :36].ravel()
    dx, yn = dense_kkt_step(ev, x, s, zs, zL, zU, lbx, ubx, mu, 0.0)
    assert np.abs(dx).max() < 1e-5 * max(1.0, np.abs(x).max()), np.abs(dx).max()
    assert np.abs(yn - y).max() < 1e-4 * max(1.0, np.abs(y).max()), np.abs(yn - y).max()


@pytest.mark.gpu
def test_riccati_step_vs_dense_kkt_gpu():
    from boundmpc_b200.ocp import default_solver
    solver = default_solver()
    worst = _check(lambda x, p, lam, wj, wh: solver.eval_batch(x, p, lam, want_jac=wj, want_hess=wh),
                   lambda *a: solver.kkt_step_batch(*a), 1e-9)
    print("GPU: worst relative deviation from the dense KKT solve", worst)
