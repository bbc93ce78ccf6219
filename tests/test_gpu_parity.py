"""Parity of the CUDA path (through the C ABI) with the oracle / golden fixtures.  GPU only."""
import numpy as np
import pytest
import torch
from tests.util import load, active_set, rel_q_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    from boundmpc_b200.ocp import default_solver
    return default_solver(N=10, nr_segs=4, dt=0.1)


@pytest.mark.parametrize("scn", ["exp1", "exp2"])
def test_eval_against_reference_executed_values(solver, scn):
    from oracle import oracle as O
    G = load(f"nlp_{scn}.npz")
    rng = np.random.default_rng(5)
    lam = rng.normal(size=(len(G["x"]), 480))
    lam.reshape(-1, 10, 48)[:, :, 36:] = np.abs(lam.reshape(-1, 10, 48)[:, :, 36:])
    e = solver.eval_batch(G["x"], G["p"], lam)
    for i in range(len(G["x"])):
        assert abs(e["f"][i] - G["f"][i]) <= 1e-12 * abs(G["f"][i])
        assert (np.abs(e["g"][i] - G["g"][i]) <= 1e-11 * np.maximum(1.0, np.abs(G["g"][i]))).all()
        assert np.abs(e["grad"][i] - G["grad"][i]).max() <= 1e-10 * max(1.0, np.abs(G["grad"][i]).max())
        je = e["jac"][i].reshape(10, 48, 440)[:, :36].reshape(360, 440)
        jr = G["jac"][i].reshape(10, 43, 440)[:, :36].reshape(360, 440)
        assert np.abs(je - jr).max() <= 1e-9 * max(1.0, np.abs(jr).max())
        d, grad, jac, hess = O.derivs_interval(G["x"][i], G["p"][i], lam[i])
        assert (np.abs(e["d"][i] - d) <= 1e-11 * np.maximum(1.0, np.abs(d))).all()
        assert np.abs(e["jac"][i] - jac).max() <= 1e-10 * max(1.0, np.abs(jac).max())
        assert np.abs(e["hess"][i] - hess).max() <= 1e-10 * np.abs(hess).max()


@pytest.mark.parametrize("scn", ["exp1", "exp2"])
def test_solve_against_golden_kkt_points(solver, scn):
    S = load(f"seq_{scn}.npz")
    r = solver.solve_batch(S["x0"], S["p"])
    assert (r["status"] == 0).all()
    assert (r["kkt"] <= solver.tol).all()
    for i in range(len(S["step"])):
        assert rel_q_error(r["x"][i], S["x"][i]) < 1e-6           # north-star: primal joint trajectory <= 1e-6 relative
        assert np.abs(r["x"][i] - S["x"][i]).max() < 1e-5
        assert abs(r["f"][i] - S["f"][i]) < 1e-7 * abs(S["f"][i])
        assert active_set({"x": r["x"][i], "g": r["g"][i]}) == active_set({"x": S["x"][i], "g": S["g"][i]})
        g = r["g"][i].reshape(10, 43)
        assert np.abs(g[:, :36]).max() < 1e-7 and g[:, 36:].max() < 1e-7


def test_same_iterates_as_oracle(solver):
    from oracle import oracle as O
    S = load("seq_exp1.npz")
    r = solver.solve_batch(S["x0"][:4], S["p"][:4])
    for i in range(4):
        ro = O.solve(S["x0"][i], S["p"][i], tol=solver.tol)
        assert ro["iters"] == r["iters"][i]
        assert np.abs(ro["x"] - r["x"][i]).max() < 1e-7


def test_casadi_like_call_surface(solver):
    S = load("seq_exp1.npz")
    lbx, ubx, lbg, ubg = solver.bounds()
    sol = solver(x0=S["x0"][0].tolist(), lbx=lbx.tolist(), ubx=ubx.tolist(), lbg=lbg.tolist(), ubg=ubg.tolist(), p=S["p"][0])
    st = solver.stats()
    assert st["success"] and st["return_status"] == "Solve_Succeeded" and st["iter_count"] > 5
    assert sol["x"].shape == (440,) and sol["g"].shape == (430,) and sol["lam_g"].shape == (430,) and sol["lam_x"].shape == (440,)
    # the reference's own feasibility check (BoundMPC.py:461-465)
    g = sol["g"]
    viol = -np.sum(g[np.where(g < lbg - 1e-6)[0]]) + np.sum(g[np.where(g > ubg + 1e-6)[0]])
    assert viol < 1e-4
    assert abs(sol["f"] - 1963.4557512307) < 1e-6


def test_batch_determinism_and_device_entry(solver):
    S = load("seq_exp2.npz")
    reps = 40
    x0 = np.tile(S["x0"], (reps, 1))
    p = np.tile(S["p"], (reps, 1))
    a = solver.solve_batch(x0, p)
    nb = len(S["step"])
    for k in ("x", "lam_g", "lam_x", "f"):
        ref = a[k][:nb]
        assert np.array_equal(a[k].reshape(reps, *ref.shape), np.broadcast_to(ref, (reps, *ref.shape)))   # bitwise
    xd = torch.from_numpy(x0).cuda()
    pd = torch.from_numpy(p).cuda()
    b = solver.solve_batch(xd, pd)
    torch.cuda.synchronize()
    assert np.array_equal(b["x"].cpu().numpy(), a["x"])
    assert np.array_equal(b["iters"].cpu().numpy(), a["iters"])


def test_edge_cases(solver):
    S = load("seq_exp1.npz")
    r0 = solver.solve_batch(np.empty((0, 440)), np.empty((0, 505)))
    assert r0["x"].shape == (0, 440)
    with pytest.raises(ValueError):
        solver.solve_batch(S["x0"][:2], S["p"][:3])
    # a start outside the bounds is pushed inside and still converges to the same point
    x0 = S["x0"][0].copy()
    x0[15:22] = 3.0            # dq guess beyond every velocity limit at stage 0
    x0[41::44] = -0.1          # negative path parameter
    r = solver.solve_batch(x0[None], S["p"][:1])
    assert r["status"][0] == 0
    assert rel_q_error(r["x"][0], S["x"][0]) < 1e-6


def test_two_pass_scheduling_is_bitwise_neutral(solver, monkeypatch):
    """Batches larger than the resident grid are solved in two passes (first slice of every instance, parked iterates
    resumed hard-first); BMPC_SINGLE_PASS switches to the plain work queue.  Same bits either way."""
    S1, S2 = load("seq_exp1.npz"), load("seq_exp2.npz")
    x0 = np.tile(np.concatenate([S1["x0"], S2["x0"]]), (20, 1))      # 620 instances > 444 resident CTAs
    p = np.tile(np.concatenate([S1["p"], S2["p"]]), (20, 1))
    a = solver.solve_batch(x0, p)
    monkeypatch.setenv("BMPC_SINGLE_PASS", "1")
    b = solver.solve_batch(x0, p)
    for k in ("x", "g", "lam_g", "lam_x", "f", "kkt", "iters", "status"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["status"] == 0).all()


def test_host_entry_with_page_locked_buffers(solver, monkeypatch):
    """bmpc_solve_batch_host with page-locked buffers (results written by the kernel, inputs fetched by it: no copies
    around the launch) returns bitwise the same as with pageable buffers (device buffers + copies), and as with
    BMPC_NO_ZERO_COPY=1."""
    from boundmpc_b200 import batches
    B = 300
    x0, p = batches.make_batch(solver, ("exp1", "exp2"), 0, B, bound_scale=True)
    ref = solver.solve_batch(x0, p)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        return t.numpy()
    out = {k: pinned(v) for k, v in ref.items()}
    for v in out.values():
        v.fill(0)
    res = solver.solve_batch(x0, p, out)
    for k in ref:
        assert res[k] is out[k]
        np.testing.assert_array_equal(res[k], ref[k], err_msg=k)
    # page-locked inputs too: fetched by the CTA that starts an instance instead of copied before the launch
    x0p, pp = pinned(x0), pinned(p)
    x0p[:], pp[:] = x0, p
    for v in out.values():
        v.fill(0)
    res = solver.solve_batch(x0p, pp, out)
    for k in ref:
        np.testing.assert_array_equal(res[k], ref[k], err_msg=k + " (pinned inputs)")
    res = solver.solve_batch(x0p, pp)          # pinned inputs, pageable results
    for k in ref:
        np.testing.assert_array_equal(res[k], ref[k], err_msg=k + " (pinned inputs only)")
    monkeypatch.setenv("BMPC_NO_ZERO_COPY", "1")
    for v in out.values():
        v.fill(0)
    res = solver.solve_batch(x0p, pp, out)
    for k in ref:
        np.testing.assert_array_equal(res[k], ref[k], err_msg=k)


def test_reference_tolerance_ipopt_tol_is_honoured():
    """`solver_opts['ipopt']['tol']` (the reference passes 10e-6, BoundMPC.py:121) reaches the kernel: every solve ends with
    Ipopt's scaled error <= 1e-5 (north star: "KKT residual <= Ipopt tol"), in the same number of iterations as the oracle at
    that tolerance, fewer than the tight solve needs, and at the distance from the KKT point SURVEY App. D.5 gives for a
    1e-5 iterate (1e-4 relative; the tight handle is the one compared to 1e-6)."""
    from oracle import oracle as O
    from boundmpc_b200.ocp import default_solver
    s5 = default_solver(N=10, nr_segs=4, dt=0.1, solver_opts={"ipopt": {"tol": 10e-6, "max_iter": 500}})
    assert s5.tol == 1e-5
    for scn in ("exp1", "exp2"):
        S = load(f"seq_{scn}.npz")
        r = s5.solve_batch(S["x0"], S["p"])
        assert (r["status"] == 0).all() and (r["kkt"] <= 1e-5).all()
        assert r["iters"].sum() < S["iters"].sum() and (r["iters"] <= S["iters"]).all()
        for i in range(len(S["step"])):
            assert rel_q_error(r["x"][i], S["x"][i]) < 2e-4
            g = r["g"][i].reshape(10, 43)
            assert np.abs(g[:, :36]).max() < 1e-5 and g[:, 36:].max() < 1e-5
            if i < 4:
                ro = O.solve(S["x0"][i], S["p"][i], tol=1e-5)
                assert abs(ro["iters"] - int(r["iters"][i])) <= 1      # (equal on every box so far; the oracle is built -march=native)
                if ro["iters"] == r["iters"][i]:
                    assert np.abs(ro["x"] - r["x"][i]).max() < 1e-6
