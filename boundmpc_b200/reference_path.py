"""Piecewise-linear Cartesian reference path with a sliding window of `nr_segs` segments.

Host-side mirror of the reference's `ReferencePath` (ReferencePath/ReferencePath.py:5-257): same
constructor arguments, attributes and `get_parameters / get_limits / get_bound_params`, same
numbers (tests/test_host_mirror.py compares against the reference class), re-implemented on
arrays.  Quirks kept on purpose: every per-segment table is padded with
`nr_segs - 1` copies of its last entry, the padded angular rate is [1, 1, 1] and the padded
segment length 1; a zero-length position segment borrows its arc length from the rotation.
Unlike the reference the caller's lists are not mutated.
"""
import numpy as np
from .lie import log_so3


def _unit_or(v, fallback, eps):
    n = np.linalg.norm(v)
    return v / n if n > eps else np.array(fallback, float)


class ReferencePath:
    def __init__(self, p, r, p_limit, r_limit, bp1, br1, s, e_p_min, e_r_min, e_p_max, e_r_max, nr_segs=2, phi_bias=0):
        S = self.nr_segs = int(nr_segs)
        pad = S - 1
        L = len(p)                       # via points
        self.phi_bias = phi_bias
        self.switched = True
        self.sector = 0

        def padded(seq):
            seq = [np.array(v, float) if np.ndim(v) else float(v) for v in seq]
            return seq + [seq[-1]] * pad

        self.s, self.e_p_min, self.e_r_min = padded(s), padded(e_p_min), padded(e_r_min)
        self.e_p_max, self.e_r_max = padded(e_p_max), padded(e_r_max)
        self.p_lower, self.p_upper = padded(p_limit[0]), padded(p_limit[1])
        self.r_lower, self.r_upper = padded(r_limit[0]), padded(r_limit[1])

        p = [np.array(v, float) for v in p]
        r = [np.array(v, float) for v in r]
        # rotation increments and their running sum ("integrated omega" of the reference)
        dr = [log_so3(r[i] @ r[i - 1].T) for i in range(1, L)]
        iw = [np.zeros(3)]
        for d in dr:
            iw.append(iw[-1] + d)
        dr += [np.ones(3)] * pad
        iw += [iw[-1]] * pad
        # position increments; a (near) zero one borrows the previous direction
        dp = []
        for i in range(1, L):
            d = p[i] - p[i - 1]
            if np.linalg.norm(d) < 1e-3:
                d = dp[-1] if i > 1 else np.array([0.0, 1.0, 0.0])
            dp.append(d)
        dp += [dp[-1]] * pad
        # arc length per segment
        seg = []
        for i in range(1, L):
            li = np.linalg.norm(p[i] - p[i - 1])
            if li < 1e-3:
                li = np.linalg.norm(dr[i - 1]) / np.pi
            seg.append(li)
        self.phi = [0.0] + seg + [1.0] * pad
        self.phi_max = float(np.sum(seg)) + phi_bias
        # angular rate per unit path parameter (the reference rescales the first L entries)
        dr = [np.array(d, float) for d in dr]
        for i in range(L):
            dr[i] = dr[i] / self.phi[i + 1]
        self.p = p + [p[-1]] * pad
        self.r = r + [r[-1]] * pad
        self.dp, self.dr, self.iw = dp, dr, iw
        # orthonormal error bases: one Gram-Schmidt step against the tangent / rotation axis
        nb = len(bp1)
        self.bp1, self.bp2, self.br1, self.br2 = [], [], [], []
        for i in range(nb):
            t = self.dp[i] / np.linalg.norm(self.dp[i])
            b = np.array(bp1[i], float)
            b = b - (t @ b) * t
            b = b / np.linalg.norm(b)
            self.bp1.append(b)
            self.bp2.append(np.cross(t, b))
        for i in range(nb):
            w = _unit_or(self.dr[i], [0.0, 1.0, 0.0], 1e-4)
            b = np.array(br1[i], float)
            b = b - (w @ b) * w
            b = b / np.linalg.norm(b)
            self.br1.append(b)
            self.br2.append(np.cross(w, b))
        for lst in (self.bp1, self.bp2, self.br1, self.br2):
            lst += [lst[-1]] * pad
        # window tables
        self.pd = np.zeros((6, S))
        self.dpd = np.zeros((6, S))
        self.dpd_normed = np.zeros((3, S))
        self.ddpd = np.zeros((6, S))
        self.asymm_lower = np.zeros((4, S))
        self.asymm_upper = np.zeros((4, S))
        self.phi_switch = np.ones(S + 1) * phi_bias
        self._cum = np.cumsum(self.phi)
        for i in range(S):
            self.set_point(i)
        self.compute_normed_velocity()

    def compute_normed_velocity(self):
        for i in range(self.nr_segs):
            self.dpd_normed[:, i] = _unit_or(self.dpd[3:, i], [0.0, 1.0, 0.0], 1e-4)

    def set_point(self, idx):
        j = self.sector + idx
        self.pd[:3, idx] = self.p[j]
        self.pd[3:, idx] = self.iw[j]
        self.dpd[:3, idx] = self.dp[j] / np.linalg.norm(self.dp[j])
        self.dpd[3:, idx] = self.dr[j]
        self.asymm_lower[:2, idx] = self.p_lower[j]
        self.asymm_lower[2:, idx] = self.r_lower[j]
        self.asymm_upper[:2, idx] = self.p_upper[j]
        self.asymm_upper[2:, idx] = self.r_upper[j]
        self.phi_switch[idx + 1] = self._cum[j + 1] + self.phi_bias

    def update(self, phi_current):
        """Slide the window while the path parameter is past the first switching point."""
        if phi_current <= self.phi_switch[1]:
            self.switched = False
        while phi_current > self.phi_switch[1]:
            self.switched = True
            self.sector += 1
            S = self.nr_segs
            for tab in (self.pd, self.dpd, self.asymm_lower, self.asymm_upper):
                tab[:, :S - 1] = tab[:, 1:S].copy()
            self.phi_switch[:S - 1] = self.phi_switch[1:S].copy()
            self.phi_switch[S - 1] = self.phi_switch[S] + self.phi_bias
            self.set_point(S - 1)
            self.compute_normed_velocity()

    def get_parameters(self, phi_current):
        self.update(phi_current)
        return self.pd, self.dpd_normed, self.dpd, self.ddpd, self.phi_switch

    def _window(self, lst):
        return np.array(lst[self.sector:self.sector + self.nr_segs])

    def get_limits(self):
        return (self.asymm_lower, self.asymm_upper, self._window(self.bp1).T, self._window(self.bp2).T,
                self._window(self.br1).T, self._window(self.br2).T)

    # ------------------------------------------------------------------ batched parameter builder
    PT_ROW = 41

    def path_table(self):
        """[J, 41] table of the padded path segments for the CUDA parameter builder
        (`bmpc_prepare_batch`, csrc/bmpc_prepare.cuh: row layout PT_*).  Built once per path."""
        J = len(self.dp)
        T = np.zeros((J, self.PT_ROW))
        for j in range(J):
            T[j, 0:3] = self.p[j]
            T[j, 3:6] = self.iw[j]
            T[j, 6:9] = self.dp[j] / np.linalg.norm(self.dp[j])
            T[j, 9:12] = self.dr[j]
            T[j, 12] = self._cum[j + 1] + self.phi_bias
            T[j, 13:15], T[j, 15:17] = self.p_lower[j], self.p_upper[j]
            T[j, 17:19], T[j, 19:21] = self.r_lower[j], self.r_upper[j]
            T[j, 21:24], T[j, 24:27] = self.bp1[j], self.bp2[j]
            T[j, 27:30], T[j, 30:33] = self.br1[j], self.br2[j]
            T[j, 33:38] = self.e_p_min[j], self.e_r_min[j], self.e_p_max[j], self.e_r_max[j], self.s[j]
            T[j, 38:41] = log_so3(self.r[j])
        return T

    def get_bound_params(self):
        return (self._window(self.e_p_min), self._window(self.e_r_min), self._window(self.e_p_max),
                self._window(self.e_r_max), self._window(self.s))
