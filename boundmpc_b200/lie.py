"""Small SO(3) helpers used by the host-side pre/post-processing.

Mirrors (not copies) the reference's `bound_mpc/utils/lie_functions.py:5-64` and the
scipy `Rotation` calls sprinkled through `BoundMPC.py` / `ReferencePath.py`.
All fp64, numpy only (no scipy dependency on the hot host path).
"""
import numpy as np


def skew(a):
    return np.array([[0.0, -a[2], a[1]],
                     [a[2], 0.0, -a[0]],
                     [-a[1], a[0], 0.0]])


def exp_so3(rv):
    """Rotation vector -> rotation matrix (Rodrigues)."""
    rv = np.asarray(rv, float)
    th = np.linalg.norm(rv)
    K = skew(rv)
    if th < 1e-8:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + (np.sin(th) / th) * K + ((1.0 - np.cos(th)) / th ** 2) * K @ K


def log_so3(Rm):
    """Rotation matrix -> rotation vector, robust near 0 and pi (quaternion route, the
    same branch structure scipy's `Rotation.from_matrix(...).as_rotvec()` takes)."""
    Rm = np.asarray(Rm, float)
    # matrix -> quaternion (x, y, z, w), Shepperd's method
    d = np.array([Rm[0, 0], Rm[1, 1], Rm[2, 2], Rm[0, 0] + Rm[1, 1] + Rm[2, 2]])
    c = int(np.argmax(d))
    q = np.empty(4)
    if c != 3:
        i, j, k = c, (c + 1) % 3, (c + 2) % 3
        q[i] = 1 - d[3] + 2 * Rm[i, i]
        q[j] = Rm[j, i] + Rm[i, j]
        q[k] = Rm[k, i] + Rm[i, k]
        q[3] = Rm[k, j] - Rm[j, k]
    else:
        q[0] = Rm[2, 1] - Rm[1, 2]
        q[1] = Rm[0, 2] - Rm[2, 0]
        q[2] = Rm[1, 0] - Rm[0, 1]
        q[3] = 1 + d[3]
    q /= np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    s = np.linalg.norm(q[:3])
    ang = 2.0 * np.arctan2(s, q[3])
    if ang <= 1e-3:
        a2 = ang * ang
        scale = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    else:
        scale = ang / np.sin(ang / 2)
    return scale * q[:3]


def rodrigues(axis, ang):
    """Rotation about a unit axis (reference: lie_functions.py:19-38)."""
    K = skew(axis)
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _jac_so3_coeff(a):
    th = np.linalg.norm(a) + 1e-6  # the reference's regularised angle, lie_functions.py:43,56
    return 1.0 / th ** 2 - (1 + np.cos(th)) / (2 * th * np.sin(th))


def jac_so3_inv_right(a):
    """reference: lie_functions.py:41-51"""
    K = skew(a)
    return np.eye(3) + 0.5 * K + _jac_so3_coeff(a) * K @ K


def jac_so3_inv_left(a):
    """reference: lie_functions.py:54-64"""
    K = skew(a)
    return np.eye(3) - 0.5 * K + _jac_so3_coeff(a) * K @ K


def euler_zyx_intrinsic_from_matrix(Rm):
    """scipy `Rotation.from_matrix(R).as_euler('zyx')` (extrinsic z, then y, then x):
    R = Rx(c) Ry(b) Rz(a), returned as [a, b, c]."""
    Rm = np.asarray(Rm, float)
    # R = Rx(c) @ Ry(b) @ Rz(a)
    # R[0,2] = sin(b); R[0,0] = cos(b)cos(a); R[0,1] = -cos(b) sin(a)
    # R[1,2] = -sin(c)cos(b); R[2,2] = cos(c)cos(b)
    sb = np.clip(Rm[0, 2], -1.0, 1.0)
    b = np.arcsin(sb)
    if abs(sb) < 1 - 1e-12:
        a = np.arctan2(-Rm[0, 1], Rm[0, 0])
        c = np.arctan2(-Rm[1, 2], Rm[2, 2])
    else:  # gimbal lock: set third angle to zero like scipy
        c = 0.0
        a = np.arctan2(Rm[1, 0], Rm[1, 1])
    return np.array([a, b, c])
