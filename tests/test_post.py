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
    mpc = batches.make_mpc(scn, s)
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
        traj, _, _, _, _ = mpc.finish(sol, mpc.solver.stats(), aux)
        T, so = emu.post(tab, [0], [sector], st, np.asarray(sol['x']), [0])
        assert _compare(T[0], traj, 10) < 1e-12
        ref = np.concatenate(([mpc.phi_current[0], mpc.dphi_current[0], mpc.ddphi_current[0], mpc.dddphi_current[0]], mpc.pr_ref, mpc.iw_ref))
        assert np.abs(so[0][40:50] - ref).max() < 1e-12                      # path-parameter state, pr_ref, iw_ref of the next step
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mpc.dt)
        jerk = traj['dddq'][:, 0].copy()
        if mpc.phi_max[0] - mpc.phi_current[0] <= 0.01:
            break
    assert len(sectors) >= 2          # the rotation reference went through at least one segment switch


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
