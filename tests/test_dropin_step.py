"""`BoundMPC.step` on the device (one library call: builder -> solve -> finish + logging branch) against the numpy mirror
of the reference's pre- / post-processing around the same solver.  CPU: the host build of the kernel
sources behind tests/emu/EmuSolver; GPU: `bmpc_mpc_step_batch_host` through the C ABI."""
import numpy as np
import pytest

from boundmpc_b200 import scenarios, batches
from boundmpc_b200.bound_mpc import integrate_joint
from boundmpc_b200.robot_model import RobotModel
from boundmpc_b200.lie import exp_so3


def _rot_close(a, b, tol):
    """rotation vectors compared as rotations (angle pi has two signs)"""
    return np.abs(exp_so3(np.asarray(a)) - exp_so3(np.asarray(b))).max() < tol


def _loop(solver, scn, steps, tol):
    dev = batches.make_mpc(scn, solver, real_time=False)      # logging branch on
    mir = batches.make_mpc(scn, solver, real_time=False)
    mir.device_step = False
    assert dev.device_step and hasattr(solver, "mpc_step_host")
    rm = RobotModel()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    x_phi_d = np.array([dev.phi_max[0], 0.0, 0.0])
    worst = 0.0
    for k in range(steps):
        p_lie = rm.fk(q)
        td, rd, ed, _, it_d = dev.step(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        tm, rmr, em, _, it_m = mir.step(q, dq, ddq, p_lie, v, x_phi_d, jerk)
        assert td is not None and tm is not None
        assert abs(it_d - it_m) <= 1, (k, it_d, it_m)
        for key in tm:
            a, b = np.asarray(td[key]), np.asarray(tm[key])
            assert a.shape == b.shape, (key, a.shape, b.shape)
            if key == "p":      # rows 3.. are rotation vectors
                worst = max(worst, np.abs(a[:3] - b[:3]).max())
                assert np.abs(a[:3] - b[:3]).max() < tol, (k, key)
                assert all(_rot_close(a[3:, i], b[3:, i], tol) for i in range(a.shape[1])), (k, key)
            else:
                e = np.abs(a - b).max() / max(1.0, np.abs(b).max())
                worst = max(worst, e) if key not in ("dddq", "dddphi") else worst
                # (the jerks sit in flat directions of the objective: two solves of inputs that differ by 1e-16 agree to 1e-6 there)
                assert e < (tol if key not in ("dddq", "dddphi") else 100 * tol), (k, key, e)
        for key in rmr:
            for i, (a, b) in enumerate(zip(rd[key], rmr[key])):
                a, b = np.ravel(a), np.ravel(b)
                if key == "p":
                    assert np.abs(a[:3] - b[:3]).max() < tol and _rot_close(a[3:], b[3:], tol), (k, key, i)
                else:
                    assert np.abs(a - b).max() < tol * max(1.0, np.abs(b).max()), (k, key, i)
        for key in em:
            for i, (a, b) in enumerate(zip(ed[key], em[key])):
                if key == "e_r":
                    assert _rot_close(a, b, 10 * tol), (k, key, i)
                else:
                    assert np.abs(np.ravel(a) - np.ravel(b)).max() < 10 * tol * max(1.0, np.abs(b).max()), (k, key, i)
        # controller state carried to the next step
        for attr in ("phi_current", "dphi_current", "ddphi_current", "dddphi_current", "iw_ref"):
            assert np.abs(np.ravel(getattr(dev, attr)) - np.ravel(getattr(mir, attr))).max() < tol, (k, attr)
        assert _rot_close(dev.pr_ref, mir.pr_ref, tol) and dev.error_count == mir.error_count == 0
        assert dev.ref_path.sector == mir.ref_path.sector
        assert np.abs(np.asarray(dev.prev_solution) - np.asarray(mir.prev_solution)).max() < 1e3 * tol   # (flat jerk directions)
        jm = np.concatenate((jerk[:, None], tm['dddq'][:, :2]), axis=1)
        q, dq, ddq, _, v, _, _ = integrate_joint(rm, jm, q, dq, ddq, mir.dt)
        jerk = tm['dddq'][:, 0].copy()
        # keep the two controllers on the same trajectory (teacher forcing: differences do not accumulate)
        dev.prev_solution = np.array(mir.prev_solution).copy()
        for attr in ("phi_current", "dphi_current", "ddphi_current", "dddphi_current", "iw_ref", "pr_ref"):
            setattr(dev, attr, np.array(getattr(mir, attr)).copy())
    return worst


@pytest.mark.parametrize("name,steps", [("exp1", 5), ("exp2", 4)])
def test_device_step_matches_mirror_host_build(name, steps):
    from tests.emu.emu_solver import EmuSolver
    scn = scenarios.experiment1(n=10) if name == "exp1" else scenarios.experiment2(n=10)
    worst = _loop(EmuSolver(), scn, steps, 1e-7)
    print(name, "host build: worst relative deviation of traj_data from the mirror", worst)


@pytest.mark.gpu
@pytest.mark.parametrize("name,steps", [("exp1", 40), ("exp2", 25)])
def test_device_step_matches_mirror_gpu(name, steps):
    from boundmpc_b200.ocp import default_solver
    scn = scenarios.experiment1(n=10) if name == "exp1" else scenarios.experiment2(n=10)
    worst = _loop(default_solver(), scn, steps, 1e-7)
    print(name, "GPU: worst relative deviation of traj_data from the mirror", worst)
