"""Host-side kinematic model of the KUKA iiwa14 + tool (numpy, fp64).

Mirrors the call surface of the reference's `RobotModel` (RobotModel/RobotModel.py:6-60)
but not its Maple-generated closed forms: kinematics are the serial-chain product read
from the xacro joint origins (urdf/body/iiwa14.xacro:65,104,143,182,221,260,299,340),
base frame = link_0, tool point 0.2174 m along z of link_7, tool frame = link_7 frame.
This equals `RobotModel.fk_pos / fk / jacobian_fk / djacobian_fk / ddjacobian_fk` to
rounding (tests/test_host_mirror.py).  Limits are the reference's (RobotModel.py:20-43),
NOT the URDF <limit> tags.
"""
import numpy as np
from .lie import log_so3

PI = np.pi
# joint origins (xyz, rpy) parent->child, joint axis = local z
JOINT_XYZ = np.array([
    [0.0, 0.0, 0.1525],
    [0.0, 0.0, 0.2075],
    [0.0, 0.2325, 0.0],
    [0.0, 0.0, 0.1875],
    [0.0, 0.2125, 0.0],
    [0.0, 0.0, 0.1875],
    [0.0, 0.0796, 0.0],
])
JOINT_RPY = np.array([
    [0.0, 0.0, 0.0],
    [PI / 2, 0.0, PI],
    [PI / 2, 0.0, PI],
    [PI / 2, 0.0, 0.0],
    [-PI / 2, PI, 0.0],
    [PI / 2, 0.0, 0.0],
    [-PI / 2, PI, 0.0],
])
TOOL_Z = 0.2174

DEG = PI / 180.0
Q_LIM_UPPER = np.array([165, 115, 165, 115, 165, 115, 170]) * DEG
Q_LIM_LOWER = -Q_LIM_UPPER
DQ_LIM_UPPER = np.array([85, 85, 100, 75, 130, 135, 135]) * DEG
DQ_LIM_LOWER = -DQ_LIM_UPPER
TAU_LIM_UPPER = np.array([320, 320, 176, 176, 110, 40, 40], float)
TAU_LIM_LOWER = -TAU_LIM_UPPER
U_MAX = 35.0
U_MIN = -35.0


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


# constant part of each joint transform; entries are exactly 0/+-1 after rounding
JOINT_ROT = np.array([np.round(_rpy(*rpy)) for rpy in JOINT_RPY])


def chain(q):
    """-> z [7,3] joint axes, o [7,3] joint origins, p tool point, R tool rotation (world)."""
    R = np.eye(3)
    o = np.zeros(3)
    zs = np.empty((7, 3))
    os_ = np.empty((7, 3))
    for i in range(7):
        o = o + R @ JOINT_XYZ[i]
        R = R @ JOINT_ROT[i]
        zs[i] = R[:, 2]
        os_[i] = o
        c, s = np.cos(q[i]), np.sin(q[i])
        R = R @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    p = o + R[:, 2] * TOOL_Z
    return zs, os_, p, R


class RobotModel:
    """Same public surface as the reference class (limits + numpy kinematics)."""

    def __init__(self):
        self.q_lim_lower = Q_LIM_LOWER.tolist()
        self.q_lim_upper = Q_LIM_UPPER.tolist()
        self.dq_lim_lower = DQ_LIM_LOWER.tolist()
        self.dq_lim_upper = DQ_LIM_UPPER.tolist()
        self.tau_lim_lower = TAU_LIM_LOWER.tolist()
        self.tau_lim_upper = TAU_LIM_UPPER.tolist()
        self.u_max = U_MAX
        self.u_min = U_MIN

    def get_robot_limits(self):
        # order of the reference, RobotModel.py:45-48
        return (self.q_lim_upper, self.q_lim_lower, self.dq_lim_upper,
                self.dq_lim_lower, self.tau_lim_upper, self.tau_lim_lower,
                self.u_max, self.u_min)

    def fk_pos(self, q):
        return chain(q)[2]

    def hom_transform_endeffector(self, q):
        _, _, p, R = chain(q)
        H = np.eye(4)
        H[:3, :3] = R
        H[:3, 3] = p
        return H

    def fk(self, q):
        _, _, p, R = chain(q)
        return np.concatenate((p, log_so3(R)))

    def jacobian_fk(self, q):
        z, o, p, _ = chain(q)
        J = np.empty((6, 7))
        J[:3] = np.cross(z, p - o).T
        J[3:] = z.T
        return J

    def _rates(self, q, dq, ddq=None):
        z, o, p, _ = chain(q)
        r = p - o
        Om = np.zeros((8, 3))        # angular velocity of the link carrying axis i
        for i in range(7):
            Om[i + 1] = Om[i] + dq[i] * z[i]
        zd = np.cross(Om[:7], z)
        # velocity of the tool point relative to each joint origin: d/dt (p - o_i)
        # p - o_i is carried by joints i..6
        rd = np.empty((7, 3))
        tail = np.zeros(3)           # sum_{k>=i} dq_k z_k x r_k, accumulated from the tip
        for i in range(6, -1, -1):
            tail = tail + dq[i] * np.cross(z[i], r[i])
            rd[i] = tail + np.cross(Om[i], r[i])
        return z, o, p, r, Om, zd, rd

    def djacobian_fk(self, q, dq):
        z, o, p, r, Om, zd, rd = self._rates(q, dq)
        dJ = np.empty((6, 7))
        dJ[:3] = (np.cross(zd, r) + np.cross(z, rd)).T
        dJ[3:] = zd.T
        return dJ

    def ddjacobian_fk(self, q, dq, ddq, eps=1e-6):
        # second time derivative of J along (dq, ddq): d/dt [ sum_i dJ/dq_i dq_i ]
        # evaluated by central differences of the analytic dJ along the flow
        # (host-side post-processing only; the OCP never uses it)
        qp, qm = q + eps * dq + 0.5 * eps ** 2 * ddq, q - eps * dq + 0.5 * eps ** 2 * ddq
        dqp, dqm = dq + eps * ddq, dq - eps * ddq
        return (self.djacobian_fk(qp, dqp) - self.djacobian_fk(qm, dqm)) / (2 * eps)

    def forward_kinematics(self, q, dq):
        return self.fk(q), self.jacobian_fk(q), self.djacobian_fk(q, dq)

    def velocity_ee(self, q, dq):
        return self.jacobian_fk(q)[:3] @ dq

    def omega_ee(self, q, dq):
        return self.jacobian_fk(q)[3:] @ dq
