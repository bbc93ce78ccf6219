"""Replanning: `bmpc_update_batch` + the re-projected warm start of `bmpc_prepare_batch` against the host
mirror's `BoundMPC.update` / `prepare` (BoundMPC.py:163-217, 335-369).  CPU: host build of the kernel source."""
import numpy as np
import pytest
from tests.emu import emu
from tests.emu.emu_solver import EmuSolver
from boundmpc_b200 import batches, scenarios
from boundmpc_b200.bound_mpc import integrate_joint
from boundmpc_b200.lie import exp_so3
from boundmpc_b200.robot_model import RobotModel


def _new_path(scn, p_lie):
    """Replanning towards the same goal: the rest of the scenario's path, restarted at the current tool pose."""
    s = {k: v for k, v in scn.items()}
    s['p_via'] = [p_lie[:3].copy()] + [np.array(v, float) for v in scn['p_via'][1:]]
    s['r_via'] = [exp_so3(p_lie[3:])] + [np.array(v, float) for v in scn['r_via'][1:]]
    return s


def _replan_loop(s, update_fn, prepare_fn):
    """Every step: builder vs mirror `prepare` — before the update (shifted warm start) and after it (re-projected warm
    start incl. the fallback steps in which the previous solution is kept); `update_fn` / `prepare_fn` = the kernels
    under test (host build or GPU), compared with the numpy mirror directly."""
    scn = scenarios.experiment1(n=10)
    mpc = batches.make_mpc(scn, s)
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    tabs = [mpc.ref_path.path_table()]
    path_id, cases, errs = np.array([0], np.int32), set(), []
    a_c = j_c = np.zeros(6)
    for step in range(26):
        p_lie = rm.fk(q)
        if step == 12:                             # replanning event: new path from the current pose
            n = _new_path(scn, p_lie)
            st_old, sec_old, _ = mpc.builder_state(q, dq, ddq, p_lie, v, x_phi_d, jerk)
            mpc.update(n['p_via'], n['r_via'], [n['p_lower'], n['p_upper']], [n['r_lower'], n['r_upper']], n['bp1'], n['br1'], n['s'],
                       n['e_p_min'], n['e_r_min'], n['e_p_max'], n['e_r_max'], p_lie, v, a_c, j_c, p0=p_lie, params=batches.Params(scn))
            x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
            t_new = mpc.ref_path.path_table()
            assert np.isfinite(t_new).all()
            J = max(len(tabs[0]), len(t_new))
            T = np.zeros((2, J, 41))
            for k, t in enumerate((tabs[0], t_new)):
                T[k, :len(t)] = t
                T[k, len(t):] = t[-1]
            tabs = T
            st_u, sec_u, pid_u = update_fn(tabs, [0.0, mpc.ref_path.phi_max], [1], np.concatenate((p_lie, v, a_c, j_c)), st_old, [sec_old], path_id)
            st_m, sec_m, _ = mpc.builder_state(q, dq, ddq, p_lie, v, x_phi_d, jerk)
            assert sec_u[0] == sec_m == 0 and pid_u[0] == 1
            keep = np.ones(76, bool); keep[50:53] = False            # (x_phi_d is the caller's)
            assert np.abs(st_u[0][keep] - st_m[keep]).max() < 1e-12
            path_id = pid_u
        st, sector, prev = mpc.builder_state(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        tab3 = tabs if isinstance(tabs, np.ndarray) else np.asarray(tabs)
        x0e, pe, sece = prepare_fn(tab3, path_id, [sector], st, prev)
        w0, params, aux = mpc.prepare(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        assert sece[0] == mpc.ref_path.sector
        assert (np.abs(pe[0] - params) / np.maximum(1.0, np.abs(params))).max() < 1e-12
        assert np.abs(x0e[0] - w0).max() < 1e-6                      # (central-difference ddJ: 1e-6 * rounding / eps)
        if step >= 12:
            ph = x0e[0].reshape(10, 44)[:, 41]
            cases |= {"clamp"} if (np.abs(ph - (aux['phi_switch'][1] - 0.01)) < 1e-15).any() else set()
            cases |= {"project"} if ((ph > 0) & (ph < aux['phi_switch'][1] - 0.011)).any() else set()
        sol = mpc.solver(x0=w0, lbx=mpc.lbu, ubx=mpc.ubu, lbg=mpc.lbg, ubg=mpc.ubg, p=params)
        traj, _, _, _, _ = mpc.finish(sol, mpc.solver.stats(), aux)
        errs.append(mpc.error_count)
        if traj is None:                              # (the restarted path is infeasible from this moving state: after N rejected
            break                                     #  solves the controller gives up, BoundMPC.py:504-506)
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, a_c, j_c = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
    assert "project" in cases and len(errs) >= 20       # at least 8 steps after the update were compared


def test_update_and_reprojected_warm_start_match_host_mirror():
    _replan_loop(EmuSolver(), emu.update, emu.prepare)


@pytest.mark.gpu
def test_gpu_update_and_reprojection_match_host_mirror():
    """k_update and k_prepare (re-projected warm start) on the GPU against the numpy mirror's `update` / `prepare` directly
    (not via the host build of the same source), along a closed loop with a replanning event."""
    import torch
    from boundmpc_b200.ocp import default_solver
    s = default_solver()
    dev = torch.device("cuda")
    T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a), dt)).to(dev)

    def update_fn(tabs, phimax, new_path, cart, state, sector, path_id):
        t = dict(tables=T(tabs, np.float64), phimax=T(phimax, np.float64), new_path=T(new_path, np.int32), cart=T(np.atleast_2d(cart), np.float64),
                 state=T(np.atleast_2d(state), np.float64), sector=T(sector, np.int32), path_id=T(path_id, np.int32))
        s.update_batch(t["tables"], t["phimax"], t["new_path"], t["cart"], t["state"], t["sector"], t["path_id"])
        return t["state"].cpu().numpy(), t["sector"].cpu().numpy(), t["path_id"].cpu().numpy()

    def prepare_fn(tabs, path_id, sector, st, prev):
        r = s.prepare_batch(np.asarray(tabs), np.asarray(path_id, np.int32), np.asarray(sector, np.int32), np.atleast_2d(st), np.atleast_2d(prev))
        return r["x0"], r["p"], r["sector"]

    _replan_loop(s, update_fn, prepare_fn)


@pytest.mark.gpu
def test_gpu_update_and_reprojection_match_host_build():
    import torch
    from boundmpc_b200.ocp import default_solver
    s = default_solver()
    D = batches.make_builder_batch(s, ("exp1", "exp2"), 0, 256, bound_scale=True)
    rng = np.random.default_rng(3)
    B, dev = 256, torch.device("cuda")
    # k_update: half of the controllers move to the other path
    phimax = np.array([6.9, 1.6])
    new_path = np.where(rng.random(B) < 0.5, 1 - D["path_id"], -1).astype(np.int32)
    cart = 0.1 * rng.normal(size=(B, 24))
    first = D["tables"][np.maximum(new_path, 0), 0]                   # a replanned path starts at the measured tool pose
    cart[:, 0:3] = first[:, 0:3] + 0.01 * rng.normal(size=(B, 3))
    cart[:, 3:6] = first[:, 38:41] + 0.01 * rng.normal(size=(B, 3))
    state0 = D["state"].copy()
    moved = new_path >= 0
    state0[moved, 21:27] = cart[moved, 0:6]
    st_e, sec_e, pid_e = emu.update(D["tables"], phimax, new_path, cart, state0, D["sector"], D["path_id"])
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(tables=D["tables"], phimax=phimax, new_path=new_path, cart=cart, state=state0, sector=D["sector"], path_id=D["path_id"]).items()}
    s.update_batch(t["tables"], t["phimax"], t["new_path"], t["cart"], t["state"], t["sector"], t["path_id"])
    assert np.abs(t["state"].cpu().numpy() - st_e).max() < 1e-12
    assert np.array_equal(t["sector"].cpu().numpy(), sec_e) and np.array_equal(t["path_id"].cpu().numpy(), pid_e)
    assert (st_e[new_path >= 0, 74] == 1).all() and (st_e[new_path < 0, 74] == 0).all()
    # k_prepare with the updated flag: re-projected warm start
    x0_e, p_e, sec2_e = emu.prepare(D["tables"], pid_e, sec_e, st_e, D["prev"])
    prev = torch.from_numpy(D["prev"]).to(dev)
    r = s.prepare_batch(t["tables"], t["path_id"], t["sector"], t["state"], prev)
    assert np.array_equal(r["sector"].cpu().numpy(), sec2_e)
    assert (np.abs(r["p"].cpu().numpy() - p_e) / np.maximum(1.0, np.abs(p_e))).max() < 1e-12
    assert np.abs(r["x0"].cpu().numpy() - x0_e).max() < 1e-6         # (central-difference ddJ)
