"""Batched post-processing (csrc/bmpc_post.cuh) against the host mirror of `compute_return_data`
(boundmpc_b200/bound_mpc.py, itself checked against the reference class in tests/test_host_mirror.py).
CPU: the kernel source compiled for the host (tests/emu); GPU: the C ABI against the same host build."""
import copy
import numpy as np
import pytest
from tests.emu import emu
from tests.emu.emu_solver import EmuSolver
from boundmpc_b200 import batches, scenarios
from boundmpc_b200.bound_mpc import integrate_joint
from boundmpc_b200.robot_model import RobotModel

RF = {'p': (0, 6), 'dp': (6, 12), 'ddp': (12, 18), 'dp_normed': (18, 21), 'r_par_bound': (21, 22), 'bound_lower': (22, 26),
      'bound_upper': (26, 30), 'e_p_off': (30, 32), 'e_r_off': (32, 34), 'bp1': (34, 37), 'bp2': (37, 40), 'br1': (40, 43), 'br2': (43, 46),
      'v1': (46, 49), 'v2': (49, 52), 'v3': (52, 55)}
ER = ("e_p", "de_p", "e_p_par", "e_p_orth", "de_p_par", "de_p_orth", "e_r", "de_r", "e_r_par", "e_r_orth1", "e_r_orth2")


def _compare_log(ref, err, ref_data, err_data):
    """ref [N, 55], err [N, 33] of the kernel against the mirror's ref_data / err_data (lists per node)."""
    from boundmpc_b200.lie import exp_so3
    worst = 0.0
    for key, (a, b) in RF.items():
        m = np.array([np.ravel(x) for x in ref_data[key]])
        k = ref[:len(m), a:b]
        if key == 'p':       # rotation part: a rotation vector of angle pi has two representations; compare the rotations
            worst = max(worst, float(np.abs(k[:, :3] - m[:, :3]).max()))
            worst = max(worst, max(float(np.abs(exp_so3(k[i, 3:]) - exp_so3(m[i, 3:])).max()) for i in range(len(m))))
        else:
            worst = max(worst, float(np.abs(k - m).max()))
    for j, key in enumerate(ER):
        m = np.array([np.ravel(x) for x in err_data[key]])
        worst = max(worst, float(np.abs(err[:len(m), 3 * j:3 * j + 3] - m).max()))
    return worst


COLS = {"p": slice(0, 6), "v": slice(6, 12), "a": slice(12, 18), "q": slice(18, 25), "dq": slice(25, 32), "ddq": slice(32, 39)}


def _compare(T, traj, M):
    worst = 0.0
    for key, sl in COLS.items():
        worst = max(worst, float(np.abs(T[:M, sl].T - traj[key]).max() / max(1.0, np.abs(traj[key]).max())))
    for k, key in enumerate(("phi", "dphi", "ddphi")):
        worst = max(worst, float(np.abs(T[:M, 39 + k] - traj[key]).max()))
    return worst


@pytest.mark.parametrize("name,steps", [("exp1", 60), ("exp2", 100)])
def test_post_matches_host_mirror_along_the_closed_loop(name, steps):
    s = EmuSolver()
    scn = scenarios.experiment1(n=10) if name == "exp1" else scenarios.experiment2(n=10)
    mpc = batches.make_mpc(scn, s, real_time=False)          # real_time False: step also returns ref_data / err_data
    tab = mpc.ref_path.path_table()[None]
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    sectors = set()
    for step in range(steps):
        p_lie = rm.fk(q)
        st, _, _ = mpc.builder_state(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        w0, params, aux = mpc.prepare(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        sector = mpc.ref_path.sector
        sectors.add(sector)
        sol = mpc.solver(x0=w0, lbx=mpc.lbu, ubx=mpc.ubu, lbg=mpc.lbg, ubg=mpc.ubg, p=params)
        if step == 5:        # a failed solve two steps ago: the controller keeps its previous solution, error_count = 2
            m2 = copy.copy(mpc)
            m2.error_count = 2
            m2.pr_ref, m2.iw_ref = mpc.pr_ref.copy(), mpc.iw_ref.copy()
            traj2, _, _ = m2.compute_return_data(np.asarray(sol['x']), True, aux)
            T2, _ = emu.post(tab, [0], [sector], st, np.asarray(sol['x']), [2])
            assert _compare(T2[0], traj2, 8) < 1e-12 and not T2[0][8:].any()
        traj, ref_data, err_data, _, _ = mpc.finish(sol, mpc.solver.stats(), aux)
        T, so = emu.post(tab, [0], [sector], st, np.asarray(sol['x']), [0])
        assert _compare(T[0], traj, 10) < 1e-12
        T2, so2, ref, err = emu.post_log(tab, [0], [sector], st, params, np.asarray(sol['x']), [0])      # with the logging branch
        assert np.array_equal(T2, T) and np.array_equal(so2, so)
        assert _compare_log(ref[0], err[0], ref_data, err_data) < 1e-11
        ref = np.concatenate(([mpc.phi_current[0], mpc.dphi_current[0], mpc.ddphi_current[0], mpc.dddphi_current[0]], mpc.pr_ref, mpc.iw_ref))
        assert np.abs(so[0][40:50] - ref).max() < 1e-12                      # path-parameter state, pr_ref, iw_ref of the next step
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        if mpc.phi_max[0] - mpc.phi_current[0] <= 0.01:
            break
    assert len(sectors) >= 2          # the rotation reference went through at least one segment switch


def test_closed_loop_on_builder_solver_finish_matches_host_mirror():
    """The three device stages of a roll-out step (host builds: emu.prepare -> solver -> emu.finish with advance) against
    the mirror's `BoundMPC.step` + `integrate_joint` loop, including a rejected solve (fallback on the previous solution,
    BoundMPC.py:467-496) and the steps after it."""
    s = EmuSolver()
    scn = scenarios.experiment2(n=10)
    mpc = batches.make_mpc(scn, s)
    tab = mpc.ref_path.path_table()[None]
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([mpc.phi_max[0], 0.0, 0.0])
    from boundmpc_b200.rollout import initial_state
    st, sector, prev = initial_state(mpc, scn['q0'])
    st, prev, sec, ec = st[None].copy(), prev[None].copy(), np.array([sector], np.int32), np.array([0], np.int32)
    for step in range(30):
        # device-form step
        x0, p, sec = emu.prepare(tab, [0], sec, st, prev)
        r = s.solve_batch(x0, p)
        status, g = r["status"].copy(), r["g"].copy()
        fail = step in (7, 8)
        if fail:                                 # pretend the solver failed and left a violated point
            status[:] = 1
            g[0, 3] = 0.5
        traj_d, st_next, prev, ec = emu.finish(tab, [0], sec, st, r["x"], g, status, prev, ec)
        # mirror step from the same measured state
        p_lie = rm.fk(q)
        w0, params, aux = mpc.prepare(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        assert np.abs(w0 - x0[0]).max() < 1e-5          # warm starts: shifted previous solutions of the two loops
        sol = mpc.solver(x0=w0, lbx=mpc.lbu, ubx=mpc.ubu, lbg=mpc.lbg, ubg=mpc.ubg, p=params)
        stats = mpc.solver.stats()
        if fail:
            stats = dict(stats, success=False)
            sol = dict(sol, g=g[0])
        traj, _, _, _, _ = mpc.finish(sol, stats, aux)
        M = 10 - mpc.error_count
        assert ec[0] == mpc.error_count
        assert _compare(traj_d[0], traj, M) < 1e-6
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        # the advanced device state is the mirror's next measured state and controller state
        # (the two loops solve separately from inputs that differ in the last bits; the jerk is the soft direction)
        assert np.abs(st_next[0, 0:7] - q).max() < 1e-8 and np.abs(st_next[0, 7:14] - dq).max() < 1e-7
        assert np.abs(st_next[0, 21:27] - p_lie).max() < 1e-8 and np.abs(st_next[0, 27:33] - v).max() < 1e-7
        assert np.abs(st_next[0, 33:40] - jerk).max() < 1e-5
        assert abs(st_next[0, 40] - mpc.phi_current[0]) < 1e-8 and np.abs(st_next[0, 44:47] - mpc.pr_ref).max() < 1e-8
        st = st_next
    assert mpc.error_count == 0 and step == 29


@pytest.mark.gpu
def test_gpu_post_matches_host_build():
    import torch
    from boundmpc_b200.ocp import default_solver
    s = default_solver()
    D = batches.make_builder_batch(s, ("exp1", "exp2"), 0, 512, bound_scale=True)
    rng = np.random.default_rng(11)
    w = D["x0"] + 1e-2 * rng.normal(size=D["x0"].shape)                     # any trajectory will do
    ec = rng.integers(0, 3, len(w)).astype(np.int32)
    Te, se = emu.post(D["tables"], D["path_id"], D["sector_out"], D["state"], w, ec)
    r = s.post_batch(D["tables"], D["path_id"], D["sector_out"], D["state"], w, ec)             # host-pointer entry
    assert np.abs(r["traj"] - Te).max() <= 1e-12 * max(1.0, np.abs(Te).max())
    assert np.abs(r["state"] - se).max() <= 1e-12
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(tables=D["tables"], path_id=D["path_id"], sector=D["sector_out"], state=D["state"], w=w, ec=ec).items()}
    rd = s.post_batch(t["tables"], t["path_id"], t["sector"], t["state"], t["w"], t["ec"])
    assert np.array_equal(rd["traj"].cpu().numpy(), r["traj"]) and np.array_equal(rd["state"].cpu().numpy(), r["state"])
    # logging branch: reference data and error terms
    _, _, ref_e, err_e = emu.post_log(D["tables"], D["path_id"], D["sector_out"], D["state"], D["p"], w, ec)
    rl = s.post_log_batch(D["tables"], D["path_id"], D["sector_out"], D["state"], D["p"], w, ec)
    assert np.array_equal(rl["traj"], r["traj"])
    assert np.abs(rl["err"] - err_e).max() < 1e-10
    rot = np.abs(rl["ref"][:, :, 3:6] - ref_e[:, :, 3:6]).max(axis=2)        # (sign of a rotation vector of angle pi)
    flip = np.abs(rl["ref"][:, :, 3:6] + ref_e[:, :, 3:6]).max(axis=2)
    assert (np.minimum(rot, flip) < 1e-10).all()
    mask = np.ones(55, bool); mask[3:6] = False
    assert np.abs(rl["ref"][:, :, mask] - ref_e[:, :, mask]).max() < 1e-10


@pytest.mark.gpu
def test_gpu_rollout_follows_the_host_mirror_closed_loop():
    """On-device roll-out (prepare -> solve -> finish on one stream) of both experiments against the
    nominal closed loop of the host mirror driven by the same CUDA solver."""
    import torch
    from boundmpc_b200.ocp import default_solver
    from boundmpc_b200.rollout import initial_state, rollout
    s = default_solver()
    steps, dev = 40, torch.device("cuda")
    tabs, states, sectors, refs = [], [], [], []
    for name in ("exp1", "exp2"):
        scn = scenarios.experiment1(n=10) if name == "exp1" else scenarios.experiment2(n=10)
        mpc = batches.make_mpc(scn, s)
        tabs.append(mpc.ref_path.path_table())
        st, sector, _ = initial_state(mpc, scn['q0'])
        states.append(st); sectors.append(sector)
        snaps, stats, _ = batches.nominal_sequence(scn, s, max_steps=steps + 1)
        refs.append((np.stack([sn[1]['q'] for sn in snaps[1:steps + 1]]), [st_[0] for st_ in stats[:steps]]))
    J = max(t.shape[0] for t in tabs)
    T = np.zeros((2, J, 41))
    for k, t in enumerate(tabs):
        T[k, :t.shape[0]] = t
        T[k, t.shape[0]:] = t[-1]
    out = rollout(s, torch.from_numpy(T).to(dev), torch.tensor([0, 1], dtype=torch.int32, device=dev),
                  torch.from_numpy(np.stack(states)).to(dev), torch.tensor(sectors, dtype=torch.int32, device=dev), steps)
    assert int(out["status"].abs().sum()) == 0 and int(out["error_count"].sum()) == 0
    q = out["q"].cpu().numpy()                    # [steps, 2, 7]: joint position after each step
    for k in range(2):
        assert np.abs(q[:, k] - refs[k][0]).max() < 1e-6
        it = np.array(out["iters"][:, k].cpu().tolist()) - np.array(refs[k][1])      # (inputs differ in the last bits)
        assert (it == 0).mean() >= 0.9 and np.abs(it).max() <= 2


def test_finish_gives_up_after_n_rejected_solves():
    """BoundMPC.py:504-506: with error_count reaching N nothing is returned; the kernel writes a zero trajectory, leaves the
    controller state alone and keeps the previous solution."""
    s = EmuSolver()
    D = batches.make_builder_batch(s, ("exp2",), 2, 1)
    assert D["state"][0, 73] == 1.0
    g = np.zeros((1, 430)); g[0, 5] = 1.0                      # a violated equality row
    traj, so, prev, ec = emu.finish(D["tables"], D["path_id"], D["sector_out"], D["state"], D["x0"], g, [1], D["prev"], [9])
    assert ec[0] == 10 and not traj.any()
    assert np.array_equal(so[0], D["state"][0]) and np.array_equal(prev, D["prev"])
    traj, so, prev, ec = emu.finish(D["tables"], D["path_id"], D["sector_out"], D["state"], D["x0"], g, [1], D["prev"], [3])
    assert ec[0] == 4 and traj[0, :6].any() and not traj[0, 6:].any()
    # without a previous solution the rejected point is used anyway and no previous solution is recorded (BoundMPC.py:480-486)
    st = D["state"].copy(); st[0, 73] = 0.0
    traj, so, prev, ec = emu.finish(D["tables"], D["path_id"], D["sector_out"], st, D["x0"], g, [1], D["prev"], [0])
    assert ec[0] == 0 and so[0, 73] == 0.0 and np.array_equal(prev, D["prev"]) and traj[0, 9].any()
