"""The reference's two benchmark scenarios as plain data (no ROS).

exp1: nodes/experiment1_runner.py:22-77, exp2: nodes/experiment2_runner.py:22-118,
default path / weights: utils/path_utils.py:4-68.  The numbers are the synthetic-input
definition of BASELINE.json's configs; dt and weights are exact Python doubles
(the ROS float32 round trip is deliberately not replicated).
"""
import numpy as np
from .lie import exp_so3
from .robot_model import RobotModel


def default_weights():
    """path_utils.py:42-68 (index 4 doubles as dphi_max, BoundMPC.py:79)."""
    return np.array([1000.0, 1.0, 0.1, 0.1, 0.5, 0.05, 8.0, 5.0, 4.0, 0.5,
                     0.01, 0.01, 0.001, 0.0001, 10.0])


def _rot_euler_XYZ(a, b, c):
    """scipy R.from_euler('XYZ', [a,b,c]) (intrinsic): Rx(a) @ Ry(b) @ Rz(c)."""
    return exp_so3([a, 0, 0]) @ exp_so3([0, b, 0]) @ exp_so3([0, 0, c])


def _default_path(n):
    """path_utils.py:4-39 with nr_segs=n."""
    return dict(
        p_lower=[np.array([-1.0, -1.0]) for _ in range(n)],
        p_upper=[np.array([1.0, 1.0]) for _ in range(n)],
        r_lower=[np.array([-1.0, -1.0]) for _ in range(n)],
        r_upper=[np.array([1.0, 1.0]) for _ in range(n)],
        bp1=[np.array([0.0, 0.0, 1.0]) for _ in range(n)],
        br1=[np.array([0.0, 0.0, 1.0]) for _ in range(n)],
        s=[0.0] * n,
        e_p_min=[0.01] * n,
        e_r_min=[15 * np.pi / 180] * n,
        e_p_max=[0.20] * n,
        e_r_max=[45 * np.pi / 180] * n,
    )


def experiment1(n=10, bound_scale=None, tight=False):
    q0 = np.zeros(7)
    q0[1] = np.pi / 3.5
    q0[3] = -np.pi / 3.5
    q0[5] = -12.85714286 * np.pi / 180
    rm = RobotModel()
    p0fk = rm.fk(q0)
    p0 = p0fk[:3]
    r0 = exp_so3(p0fk[3:])
    scn = _default_path(5)
    scn['p_via'] = [p0.copy(),
                    p0 + np.array([-p0[0] * 2, 0.0, 0.0]),
                    p0 + np.array([-p0[0], p0[0], 0.0]),
                    p0 + np.array([-p0[0], -p0[0], 0.0]),
                    p0.copy()]
    r1 = _rot_euler_XYZ(0, 0, -np.pi) @ r0
    r2 = _rot_euler_XYZ(0, 0, -np.pi / 2) @ r1
    r3 = _rot_euler_XYZ(0, np.pi / 2, 0) @ _rot_euler_XYZ(np.pi / 1.001, 0, 0) @ r2
    scn['r_via'] = [r0.copy(), r1, r2, r3, r0.copy()]
    scn['e_p_max'] = [0.5] * 5
    scn['br1'][0] = np.array([0, 1.0, 0])
    scn['br1'][1] = np.array([0, 1.0, 0])
    if tight:  # BASELINE config 4
        scn['e_p_max'] = [0.1] * 5
        scn['e_r_max'] = [15 * np.pi / 180] * 5
        scn['e_p_min'] = [0.005] * 5
        scn['e_r_min'] = [5 * np.pi / 180] * 5
    return _finish(scn, 'exp1', q0, p0fk, n, bound_scale)


def experiment2(n=10, bound_scale=None):
    q0 = np.zeros(7)
    q0[3] = -np.pi / 1.8
    q0[5] = np.pi / 2 - np.pi / 1.8
    rm = RobotModel()
    p0fk = rm.fk(q0)
    p0 = p0fk[:3]
    r0 = exp_so3(p0fk[3:])
    scn = _default_path(5)
    r1 = _rot_euler_XYZ(np.pi / 2, 0, 0) @ r0
    r2 = _rot_euler_XYZ(0, 0, -np.pi / 3) @ r1
    r3 = (_rot_euler_XYZ(0, 0, np.pi / 2.01) @ _rot_euler_XYZ(np.pi / 2, 0, 0)
          @ _rot_euler_XYZ(0, 0, -np.pi / 2) @ r1)
    r4 = (_rot_euler_XYZ(0, 0, np.pi / 2) @ _rot_euler_XYZ(np.pi / 2, 0, 0)
          @ _rot_euler_XYZ(0, 0, -np.pi / 2) @ r1)
    scn['r_via'] = [r0.copy(), r1, r2, r3, r4]
    scn['p_via'] = [p0.copy(),
                    p0 + np.array([-0.2, -0.0, 0.1]),
                    p0 + np.array([-0.6, -0.6, 0.1]),
                    p0 + np.array([-0.8, -0.5, -0.2]),
                    p0 + np.array([-0.8, -0.5, -0.5])]
    scn['p_lower'] = [np.array(v) for v in ([-1.0, -1.0], [-0.01, -1.0], [-1.0, -1.0], [-0.1, -0.1], [-0.1, -0.1])]
    scn['p_upper'] = [np.array(v) for v in ([1.0, 1.0], [0.01, 1.0], [1.0, 1.0], [0.1, 0.1], [0.1, 0.1])]
    scn['r_lower'] = [np.array(v) for v in ([-1.0, -1.0], [-0.11, -0.11], [-1.0, -1.0], [-0.1, -0.1], [-0.1, -0.1])]
    scn['r_upper'] = [np.array(v) for v in ([1.0, 1.0], [0.11, 0.11], [1.0, 1.0], [0.1, 0.1], [0.1, 0.1])]
    scn['bp1'] = [np.array(v) for v in ([0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [0.0, 1.0, 0.0])]
    scn['br1'] = [np.array(v) for v in ([0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [0.0, 1.0, 0.0])]
    return _finish(scn, 'exp2', q0, p0fk, n, bound_scale)


def _finish(scn, name, q0, p0fk, n, bound_scale):
    scn.update(name=name, q0=q0, p0fk=p0fk, n=n, nr_segs=4, dt=0.1,
               weights=default_weights())
    if bound_scale is not None:  # BASELINE config 5: widths x U(0.75, 1.25)
        for k, f in zip(('e_p_min', 'e_p_max', 'e_r_min', 'e_r_max'), bound_scale):
            scn[k] = [v * f for v in scn[k]]
    return scn


SCENARIOS = {'exp1': experiment1, 'exp2': experiment2}
