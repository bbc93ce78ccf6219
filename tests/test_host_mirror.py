"""Host-side mirror (ReferencePath, BoundMPC pre/post-processing, RobotModel) against the
reference's own classes executed through tests/golden/refexec.  Needs /root/reference (skipped
on the GPU box); the solver is replaced by a replay of recorded solutions, so no GPU is needed."""
import sys
import os
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from refexec import harness as H  # noqa: E402
from boundmpc_b200 import scenarios  # noqa: E402
from boundmpc_b200.bound_mpc import BoundMPC, integrate_joint  # noqa: E402
from boundmpc_b200.robot_model import RobotModel  # noqa: E402

pytestmark = pytest.mark.skipif(not H.available(), reason="reference tree not available")


class _Params:
    def __init__(self, scn, real_time):
        self.n, self.nr_segs, self.dt, self.weights = scn['n'], scn['nr_segs'], scn['dt'], list(scn['weights'])
        self.build, self.real_time = True, real_time


class _Replay:
    """Solver stand-in that replays recorded solutions and records its inputs."""
    def __init__(self, recs):
        self.recs, self.k, self.inputs = recs, 0, []

    def bounds(self):
        from oracle import oracle as O
        return O.bounds()

    def __call__(self, x0=None, p=None, **kw):
        self.inputs.append((np.array(x0, float), np.array(p, float)))
        r = self.recs[self.k]
        self.k += 1
        self._st = dict(iter_count=r['iters'], success=True, return_status='Solve_Succeeded')
        return dict(x=r['x'], g=r['g'], lam_g=r['lam_g'], lam_x=r['lam_x'], f=r['f'])

    def stats(self):
        return self._st

    def generate_dependencies(self, *a, **k):
        pass


@pytest.mark.parametrize("scn_name,steps", [("exp1", 60), ("exp2", 61)])
def test_step_preprocessing_and_outputs_match_reference(scn_name, steps):
    from oracle import oracle as O
    scn = scenarios.SCENARIOS[scn_name]()
    recs, ref_out = [], []

    def backend(x0, lbx, ubx, lbg, ubg, p):
        x0, p = np.array(x0, float), np.array(p, float)
        r = O.solve(x0, p, tol=1e-9)
        recs.append(dict(x0=x0, p=p, **r))
        return (dict(x=r['x'], g=r['g'], lam_g=r['lam_g'], lam_x=r['lam_x'], f=r['f']),
                dict(iter_count=r['iters'], success=r['status'] == 0, return_status='ok'))

    ref = H.make_reference_mpc(scn, backend)
    rm_ref = H.reference_robot_model()
    ij_ref = H.reference_integrate_joint()
    q, dq, ddq, jerk, v = scn['q0'].copy(), np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(6)
    xd = np.array([ref.phi_max[0], 0, 0])
    states = []
    for k in range(steps):
        p_lie = rm_ref.forward_kinematics(q, dq)[0]
        states.append((q.copy(), dq.copy(), ddq.copy(), p_lie.copy(), v.copy(), jerk.copy()))
        traj = ref.step(q, dq, ddq, p_lie, v, xd, jerk)[0]
        ref_out.append({kk: np.array(vv) for kk, vv in traj.items()})
        jm = np.concatenate((jerk[:, None], traj['dddq'][:, :2]), axis=1)
        q, dq, ddq, p_lie, v, a, j = ij_ref(rm_ref, jm, q, dq, ddq, ref.dt)
        jerk = traj['dddq'][:, 0].copy()
        if ref.phi_max[0] - ref.phi_current[0] <= 0.01:
            break
    # the mirror, fed with the same states and the same solver outputs
    import copy
    s2 = copy.deepcopy(scn)
    rep = _Replay(recs)
    mine = BoundMPC(s2['p_via'], s2['r_via'], [s2['p_lower'], s2['p_upper']], [s2['r_lower'], s2['r_upper']], s2['bp1'],
                    s2['br1'], s2['s'], s2['e_p_min'], s2['e_r_min'], s2['e_p_max'], s2['e_r_max'], p0=s2['p0fk'],
                    params=_Params(s2, True), solver=rep)
    assert abs(mine.phi_max[0] - ref.phi_max[0]) < 1e-14
    rm = RobotModel()
    for k, st in enumerate(states):
        q, dq, ddq, p_lie, v, jerk = st
        assert np.abs(rm.forward_kinematics(q, dq)[0] - p_lie).max() < 1e-12
        traj = mine.step(q, dq, ddq, p_lie, v, xd, jerk)[0]
        x0, p = rep.inputs[k]
        pr = recs[k]['p'].copy()
        # row nr_segs of the a-tables is uninitialised memory in the reference (np.empty)
        for base in (220, 265, 310, 355, 400):
            pr[base + 4:base + 45:5] = 0.0
        assert np.abs(x0 - recs[k]['x0']).max() < 1e-11, k
        assert (np.abs(p - pr) <= 1e-10 * np.maximum(1.0, np.abs(pr))).all(), (k, np.argmax(np.abs(p - pr)))
        for key in ('p', 'v', 'a', 'q', 'dq', 'ddq', 'dddq', 'phi', 'dphi', 'ddphi', 'dddphi'):
            assert np.abs(traj[key] - ref_out[k][key]).max() < 1e-9, (k, key)
    # closed-loop integration helper
    jm = np.random.default_rng(0).normal(size=(7, 3))
    a = integrate_joint(rm, jm, q, dq, ddq, 0.1)
    b = ij_ref(rm_ref, jm, q, dq, ddq, 0.1)
    for u, w in zip(a[:6], b[:6]):
        assert np.abs(u - w).max() < 1e-12
    assert np.abs(a[6] - b[6]).max() < 1e-6      # Cartesian jerk uses a finite-difference d2J/dt2


def test_kinematics_match_reference():
    rm, rr = RobotModel(), H.reference_robot_model()
    rng = np.random.default_rng(1)
    for _ in range(10):
        q, dq = rng.uniform(-2, 2, 7), rng.normal(size=7)
        assert np.abs(rm.fk(q) - rr.fk(q)).max() < 1e-13
        assert np.abs(rm.jacobian_fk(q) - rr.jacobian_fk(q)).max() < 1e-13
        assert np.abs(rm.djacobian_fk(q, dq) - rr.djacobian_fk(q, dq)).max() < 1e-12
    assert rm.get_robot_limits()[0] == rr.get_robot_limits()[0] and rm.get_robot_limits()[3] == rr.get_robot_limits()[3]
