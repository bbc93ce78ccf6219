"""`BoundMPC`: the reference's controller object with the CUDA solver behind it.

Drop-in for bound_mpc/bound_mpc/BoundMPC/BoundMPC.py:19-770 — same constructor, `update(...)`
and `step(q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current, x_des=None) ->
(traj_data, ref_data, err_data, time_elapsed, iters)` — with the pre- and post-processing
re-implemented on numpy (no CasADi, no scipy) around `boundmpc_b200.ocp.setup_optimization_problem`.
`prepare()` / `finish()` expose the two halves of `step` so that batches of instances can be
built on the host and solved in one `solve_batch` call (bench.py, scenarios).

Reference behaviour kept on purpose: row `nr_segs` of the bound-coefficient
tables is never written by the reference (np.empty) — it is zero here, the solver never selects
it arithmetically; `weights[4]` doubles as dphi_max; only the first entry of e_p_min / e_r_min /
e_p_max / e_r_max / s is used; `updated` is never reset after `update()`.

Known deviation of the mirror: `prev_traj / prev_vel / prev_acc / prev_jerk` — the inputs of the re-projected
warm start after `update()` — come from the solver's own columns of `w_opt`; in the reference they are views that end up
holding the re-integrated trajectory for the columns >= error_count.  The two agree to solver tolerance when
error_count == 0 and differ after a fallback step followed by `update()`; the device path (`k_prepare`) follows the
mirror.  tests/test_replan.py covers update() with and without preceding fallback steps against the mirror, not against
the reference, for that combination.
"""
import copy
import time
from collections import defaultdict
import numpy as np

from .lie import exp_so3, log_so3, jac_so3_inv_left, jac_so3_inv_right, rodrigues, euler_zyx_intrinsic_from_matrix
from .reference_path import ReferencePath
from .robot_model import RobotModel


# ------------------------------------------------------------------------------------------ helpers
def compute_bound_params(phi1, e0, e1, s, e_max):
    """Quartic b(t) with b(0)=e0, b(phi1)=e1, b'(0)=s, b'(phi1)=-s, b(phi1/2)=e_max
    (mpc_utils_casadi.py:130-137 with phi0 = 0)."""
    a0 = e0
    a1 = s
    a2 = -(4 * phi1 * s + phi1 * s + 11 * e0 + 5 * e1 - 16 * e_max) / phi1 ** 2
    a3 = (5 * phi1 * s + 3 * phi1 * s + 18 * e0 + 14 * e1 - 32 * e_max) / phi1 ** 3
    a4 = -2 * (2 * phi1 * s + 4 * e0 + 4 * e1 - 8 * e_max) / phi1 ** 4
    return a4, a3, a2, a1, a0


def compute_initial_rot_errors(pr, pr_ref, dp_ref, br1, br2):
    """util_functions.py:11-31: rotation error and its split along (br2, axis, br1) by 'zyx' Euler angles."""
    dtau = log_so3(exp_so3(pr) @ exp_so3(pr_ref).T)
    n = np.linalg.norm(dp_ref)
    axis = dp_ref / n if n > 1e-4 else np.array([0.0, 1.0, 0.0])
    r01 = np.column_stack((br2, axis, br1))
    eul = euler_zyx_intrinsic_from_matrix(r01.T @ exp_so3(dtau) @ r01)
    return dtau, eul[1] * axis, eul[0] * br1, eul[2] * br2


def integrate_rotation_reference(pr_ref, omega, phi0, phi1):
    """util_functions.py:88-99"""
    r0 = exp_so3(pr_ref)
    n = np.linalg.norm(omega)
    if n > 1e-4:
        r0 = rodrigues(omega / n, float(np.ravel((phi1 - phi0) * n)[0])) @ r0
    return log_so3(r0)


def integrate_jerk(jm, q, dq, ddq, h):
    """State after one sample with piecewise-linear jerk from jm[:,0] to jm[:,1]
    (jerk_trajectory_casadi.py:78-175 at t = h)."""
    um, u = jm[:, 0], jm[:, 1]
    qn = q + h * dq + h * h / 2 * ddq + h ** 3 / 8 * um + h ** 3 / 24 * u
    dqn = dq + h * ddq + h * h / 3 * um + h * h / 6 * u
    ddqn = ddq + h / 2 * (um + u)
    return qn, dqn, ddqn


def integrate_joint(model, jerk_matrix, q, dq, ddq, dt):
    """util_functions.py:152-161 (the closed-loop 'simulation' step of bound_mpc_node.py:321-331)."""
    qn, dqn, ddqn = integrate_jerk(np.asarray(jerk_matrix)[:, :2], q, dq, ddq, dt)
    pn, jac, djac = model.forward_kinematics(qn, dqn)
    ddjac = model.ddjacobian_fk(q, dq, ddq)
    vn = jac @ dqn
    an = djac @ dqn + jac @ ddqn
    jn = ddjac @ dqn + 2 * djac @ ddqn + jac @ ddqn
    return qn, dqn, ddqn, pn, vn, an, jn


def _segment(phi, phi_switch, S):
    i = S - 1
    for q in range(S - 2, -1, -1):
        if phi < phi_switch[q + 1]:
            i = q
    return i


class BoundMPC:
    def __init__(self, pos_points, rot_points, pos_lim, rot_lim, bp1, br1, s, e_p_min, e_r_min, e_p_max, e_r_max,
                 p0=np.zeros(6), params=None, solver_opts=None, solver=None):
        self.N = params.n
        self.robot_model = RobotModel()
        self.updated = False
        self.updated_once = False
        self.build = params.build
        self.log = not params.real_time
        self.p0 = np.array(p0, float)
        self.error_count = 0
        self.dt = params.dt
        self.T = self.dt * self.N
        self.nr_segs = params.nr_segs
        self.ref_path = ReferencePath(pos_points, rot_points, pos_lim, rot_lim, bp1, br1, s, e_p_min, e_r_min,
                                      e_p_max, e_r_max, self.nr_segs)
        S = self.nr_segs
        self.dtau_init = np.zeros((3, S))
        self.dtau_init_par = np.zeros((3, S))
        self.dtau_init_orth1 = np.zeros((3, S))
        self.dtau_init_orth2 = np.zeros((3, S))
        self.phi_max = np.array([self.ref_path.phi_max - 0.0001])
        self.weights = np.array(params.weights, float)
        self.dphi_max = np.array([self.weights[4]])
        self.pr_ref = self.p0[3:].copy()
        self.iw_ref = np.zeros(3)
        lim = self.robot_model.get_robot_limits()
        (self.q_lim_upper, self.q_lim_lower, self.dq_lim_upper, self.dq_lim_lower, self.tau_lim_upper,
         self.tau_lim_lower, self.u_max, self.u_min) = lim
        self.ut_max, self.ut_min = self.u_max, self.u_min
        self.phi_current = np.array([0.0])
        self.phi_prev = np.array([0.0])
        self.dphi_current = np.array([0.0])
        self.ddphi_current = np.array([0.0])
        self.dddphi_current = np.array([0.0])
        self.nr_joints, self.nr_u, self.nr_x = 7, 8, 44
        self.prev_solution = None
        self.prev_infeasible_solution = None
        self.lam_g0 = 0
        self.lam_x0 = 0
        # (the reference builds this dictionary itself, BoundMPC.py:120-148; the handle reads tol and max_iter from it, the
        # other entries describe Ipopt's strategy, which csrc/bmpc_ipm.cuh restates)
        self.solver_opts = solver_opts if solver_opts is not None else {
            'verbose': False, 'verbose_init': False, 'print_time': False,
            'ipopt': {'tol': 10e-6, 'max_iter': 500, 'mu_strategy': 'adaptive', 'adaptive_mu_globalization': 'kkt-error',
                      'warm_start_init_point': 'yes', 'mu_oracle': 'loqo', 'line_search_method': 'filter',
                      'expect_infeasible_problem': 'no', 'print_level': 0}}
        if solver is not None:      # share one CUDA handle between many controller objects
            self.solver = solver
            lbx, ubx, lbg, ubg = solver.bounds()
            self.lbu, self.ubu, self.lbg, self.ubg = lbx.tolist(), ubx.tolist(), lbg.tolist(), ubg.tolist()
            self.g_names = None
        else:
            from .ocp import setup_optimization_problem
            self.solver, self.lbu, self.ubu, self.lbg, self.ubg, self.g_names = setup_optimization_problem(
                self.N, self.nr_joints, self.nr_segs, self.dt, self.u_min, self.u_max, self.ut_min, self.ut_max,
                self.q_lim_lower, self.q_lim_upper, self.dq_lim_lower, self.dq_lim_upper, self.solver_opts)
            if self.build:
                self.solver.generate_dependencies('gen_traj_opt_nlp_deps.cpp', {'cpp': True})
        self._lbg = np.array(self.lbg)
        self._ubg = np.array(self.ubg)
        self.device_step = True          # step() on the device when the solver offers it (see step)
        self._tab = self._tab_for = None
        self._zero_id = np.zeros(1, np.int32)

    # ------------------------------------------------------------------ replanning (BoundMPC.py:163-217)
    def update(self, pos_points, rot_points, pos_lim, rot_lim, bp1, br1, s, e_p_min, e_r_min, e_p_max, e_r_max,
               p, v, a, jerk, p0=np.zeros(6), params=None):
        self.updated = True
        self.updated_once = True
        self.p0 = np.array(p0, float)
        self.ref_path = ReferencePath(pos_points, rot_points, pos_lim, rot_lim, bp1, br1, s, e_p_min, e_r_min,
                                      e_p_max, e_r_max, self.nr_segs)
        self.phi_max = np.array([self.ref_path.phi_max - 0.0001])
        self.weights = np.array(params.weights, float)
        dp0 = self.ref_path.dp[0] / np.linalg.norm(self.ref_path.dp[0])
        self.phi_current = np.array([(self.p0[:3] - np.asarray(pos_points[0])) @ dp0])
        self.phi_prev = self.phi_current
        t = self.ref_path.dpd[:3, 0]
        self.dphi_current = np.array([v[:3] @ t])
        self.ddphi_current = np.array([a[:3] @ t])
        self.dddphi_current = np.array([jerk[:3] @ t])
        self.pr_ref = integrate_rotation_reference(log_so3(np.asarray(rot_points[0])), self.ref_path.dr[0], 0.0,
                                                   self.phi_current)
        self.iw_ref = self.ref_path.pd[3:, 0] + self.phi_current * self.ref_path.dpd[3:, 0]

    # ------------------------------------------------------------------ BoundMPC.py:219-265
    def compute_error_bounds(self, asymm_upper, asymm_lower, phi_switch, s, e_p_min, e_r_min, e_p_max, e_r_max):
        S = self.nr_segs
        asym = np.concatenate((asymm_upper[:2], -asymm_lower[:2], asymm_upper[2:], -asymm_lower[2:]))   # [8, S]
        ep, er, epm, erm, s0 = e_p_min[0], e_r_min[0], e_p_max[0], e_r_max[0], s[0]
        e0 = np.array([ep, ep, -ep, -ep, er, er, -er, -er, er])
        emax0 = np.array([epm, epm, -epm, -epm, erm, erm, -erm, -erm, erm])
        sv0 = np.array([s0, s0, -s0, -s0, s0, s0, -s0, -s0, s0])
        A = np.zeros((5, S + 1, 9))          # a4, a3, a2, a1, a0 ; row S stays zero (see module docstring)
        for i in range(S):
            scale = np.concatenate((asym[:, i], asym[-1:, i]))
            A[:, i, :] = compute_bound_params(phi_switch[i + 1] - phi_switch[i], e0, e0, sv0 * scale, emax0 * scale)
        return A[4], A[3], A[2], A[1], A[0]

    # ------------------------------------------------------------------ BoundMPC.py:267-304
    def compute_orientation_projection_vectors(self, br1, br2, dp_normed_ref):
        S = dp_normed_ref.shape[1]
        d0 = self.dtau_init[:, 0]
        jac_r, jac_l = jac_so3_inv_right(d0), jac_so3_inv_left(d0)
        R0 = exp_so3(d0)
        V = np.zeros((3, 3, S))
        for i in range(S):
            rest1 = R0 @ exp_so3(self.dtau_init_orth1[:, i]).T
            rest2 = rest1 @ exp_so3(self.dtau_init_par[:, i]).T
            a = jac_r @ br1[:, i]
            b = jac_so3_inv_right(log_so3(rest1)) @ dp_normed_ref[:, i]
            c = jac_so3_inv_right(log_so3(rest2)) @ br2[:, i]
            # dual basis of (a, b, c): rows of inv([a b c])
            V[:, :, i] = np.linalg.inv(np.column_stack((a, b, c)))
        return V[0], V[1], V[2], jac_l, jac_r

    # ------------------------------------------------------------------ first half of step (BoundMPC.py:310-443)
    def prepare(self, q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current):
        q0, dq0, ddq0, p0, v0 = (np.asarray(a, float) for a in (q0, dq0, ddq0, p0, v0))
        jerk_current = np.asarray(jerk_current, float)
        N, S = self.N, self.nr_segs
        p_ref, dp_normed_ref, dp_ref, ddp_ref, phi_switch = self.ref_path.get_parameters(self.phi_current)
        asymm_lower, asymm_upper, bp1, bp2, br1, br2 = self.ref_path.get_limits()
        e_p_min, e_r_min, e_p_max, e_r_max, s = self.ref_path.get_bound_params()
        # warm start
        if self.prev_solution is None:
            w0 = np.zeros((N, self.nr_x))
            w0[:, 8:15] = q0
            w0[:, 29:35] = p0
        else:
            w0 = np.array(self.prev_solution, float).reshape(N, -1).copy()
            if np.linalg.norm(p0[3:] - w0[0, 32:35]) > 1.5:     # "Reversing integrated omega", BoundMPC.py:325-333
                first = w0[0, 32:35].copy()
                w0[:-1, 32:35] = p0[3:] + (w0[1:, 32:35] - first)
                w0[-1, 32:35] = w0[-2, 32:35]
            if self.updated:                                     # re-projection after update(), BoundMPC.py:335-369
                t, pr = dp_ref[:3, 0], p_ref[:3, 0]
                for i in range(N):
                    phik = phi_switch[0] + (self.prev_traj[:3, i] - pr) @ t
                    if phik > phi_switch[1] - 0.01:
                        w0[i, 41:44] = [phi_switch[1] - 0.01, 0.0, 0.0]
                    elif phik < 0:
                        w0[i, 8:15] = q0
                        w0[i, 41:44] = 0.0
                        w0[i, 29:35] = p0
                        w0[i, 35:39] = 0.0
                    else:
                        w0[i, 41] = phik
                        w0[i, 42] = self.prev_vel[:3, i] @ t
                        w0[i, 43] = self.prev_acc[:3, i] @ t
                        w0[i, 7] = self.prev_jerk[:3, i] @ t
            else:                                                # shift by one stage, BoundMPC.py:373-375
                w0[:-1] = w0[1:].copy()
        w0 = w0.ravel()
        # orientation-error linearisation data (BoundMPC.py:379-389)
        for i in range(S):
            (self.dtau_init[:, i], self.dtau_init_par[:, i], self.dtau_init_orth1[:, i],
             self.dtau_init_orth2[:, i]) = compute_initial_rot_errors(p0[3:], self.pr_ref, dp_ref[3:, i], br1[:, i], br2[:, i])
        v_1, v_2, v_3, jac_l, jac_r = self.compute_orientation_projection_vectors(br1, br2, dp_normed_ref)
        a0, a1, a2, a3, a4 = self.compute_error_bounds(asymm_upper, asymm_lower, phi_switch, s, e_p_min, e_r_min,
                                                       e_p_max, e_r_max)
        x_phi_d_current = np.array(x_phi_d, float)
        weights_current = self.weights.copy()
        if x_phi_d[0] < 1:
            weights_current[6] *= min(1 / self.phi_max[0] ** 2, 2.0)
        phi_max = np.array([min(self.phi_current[0] + 5.0, self.phi_max[0])])
        x_phi_d_current[0] = min(self.phi_current[0] + 5.0, x_phi_d_current[0])
        qd = q0 if phi_max[0] - self.phi_current[0] < 0.05 else np.zeros(7)
        params = np.concatenate((
            q0, dq0, ddq0, self.phi_current, self.dphi_current, self.ddphi_current, p0, v0,
            self.iw_ref, self.dtau_init[:, 0], self.dtau_init_par.T.ravel(), self.dtau_init_orth1.T.ravel(),
            self.dtau_init_orth2.T.ravel(), x_phi_d_current, jerk_current, self.dddphi_current, phi_switch,
            jac_r.T.ravel(), jac_l.T.ravel(), p_ref.ravel(), dp_ref.ravel(), dp_normed_ref.ravel(),
            bp1.ravel(), bp2.ravel(), br1.ravel(), br2.ravel(),
            a4.T.ravel(), a3.T.ravel(), a2.T.ravel(), a1.T.ravel(), a0.T.ravel(),
            weights_current, phi_max, self.dphi_max, v_1.ravel(), v_2.ravel(), v_3.ravel(), qd))
        aux = dict(q0=q0, dq0=dq0, ddq0=ddq0, p0=p0, jerk_current=jerk_current, a=(a4, a3, a2, a1, a0),
                   jac_l=jac_l, jac_r=jac_r, p_ref=p_ref, dp_normed_ref=dp_normed_ref, dp_ref=dp_ref,
                   phi_switch=phi_switch.copy(), bp1=bp1, bp2=bp2, br1=br1, br2=br2, v=(v_1, v_2, v_3),
                   x_phi_d=x_phi_d_current, phi_max=phi_max)
        return w0, params, aux

    # ------------------------------------------------------------------ inputs of the CUDA parameter builder
    PS_SIZE = 76

    def builder_state(self, q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current, bound_scale=(1.0, 1.0, 1.0, 1.0)):
        """Per-instance state vector of `bmpc_prepare_batch` (csrc/bmpc_prepare.cuh: PS_*), the window position and
        the previous solution: everything `prepare()` reads from this object and its arguments."""
        st = np.zeros(self.PS_SIZE)
        st[0:7], st[7:14], st[14:21], st[21:27], st[27:33], st[33:40] = q0, dq0, ddq0, p0, v0, jerk_current
        st[40:44] = self.phi_current[0], self.dphi_current[0], self.ddphi_current[0], self.dddphi_current[0]
        st[44:47], st[47:50], st[50:53] = self.pr_ref, self.iw_ref, x_phi_d
        st[53:57] = bound_scale
        st[57] = self.phi_max[0]
        st[58:73] = self.weights
        st[73] = 0.0 if self.prev_solution is None else 1.0
        st[74] = 1.0 if self.updated else 0.0
        prev = np.zeros(self.N * self.nr_x) if self.prev_solution is None else np.asarray(self.prev_solution, float).ravel()
        return st, int(self.ref_path.sector), prev

    # ------------------------------------------------------------------ second half of step (BoundMPC.py:454-506)
    def finish(self, sol, stats, aux, time_elapsed=0.0):
        w_curr = np.array(sol['x'], float).ravel()
        iters = stats['iter_count']
        g = np.asarray(sol['g'], float).ravel()
        g_viol = -np.sum(g[g < self._lbg - 1e-6]) + np.sum(g[g > self._ubg + 1e-6])
        success = stats['success'] or g_viol < 1e-4
        using_previous = False
        if not success:
            self.error_count += 1
            print(f"[ERROR] Could not find feasible solution. Using previous solution. Error count: {self.error_count}")
            print(f"Constraint Violation Sum: {g_viol}")
            print(f"Solver status: {stats['return_status']}")
            using_previous = True
            if self.prev_solution is not None:
                self.prev_infeasible_solution = w_curr
                w_opt = np.copy(self.prev_solution)
            else:
                print("[WARNING] Previous solution not found, using infeasible solution.")
                self.error_count = 0
                w_opt = w_curr
                self.lam_g0, self.lam_x0 = sol['lam_g'], sol['lam_x']
                self.prev_infeasible_solution = self.prev_solution
        else:
            self.error_count = 0
            w_opt = w_curr
            self.prev_solution = copy.deepcopy(w_opt)
            self.lam_g0, self.lam_x0 = sol['lam_g'], sol['lam_x']
            self.prev_infeasible_solution = w_opt
        if self.error_count < self.N:
            traj_data, ref_data, err_data = self.compute_return_data(w_opt, using_previous, aux)
            return traj_data, ref_data, err_data, time_elapsed, iters
        return None, None, None, None, None

    def step(self, q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current, x_des=None):
        """One MPC step (BoundMPC.py:306-506).  With the CUDA solver behind it the whole step -- parameter builder, solve,
        accept / fallback, post-processing and logging branch -- is one library call on the device (`_step_device`);
        `device_step = False` (or a solver without that entry) takes the numpy mirror around `solver(x0=, p=)`."""
        if self.device_step and hasattr(self.solver, "mpc_step_host"):
            return self._step_device(q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current)
        return self.step_mirror(q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current)

    REF_COLS = (("p", 0, 6), ("dp", 6, 12), ("ddp", 12, 18), ("dp_normed", 18, 21), ("r_par_bound", 21, 22), ("bound_lower", 22, 26),
                ("bound_upper", 26, 30), ("e_p_off", 30, 32), ("e_r_off", 32, 34), ("bp1", 34, 37), ("bp2", 37, 40), ("br1", 40, 43),
                ("br2", 43, 46), ("v1", 46, 49), ("v2", 49, 52), ("v3", 52, 55))
    ERR_KEYS = ("e_p", "de_p", "e_p_par", "e_p_orth", "de_p_par", "de_p_orth", "e_r", "de_r", "e_r_par", "e_r_orth1", "e_r_orth2")

    def _step_device(self, q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current):
        """`step` through `bmpc_mpc_step_batch_host` (B = 1): this object only packs its state (76 doubles + the previous
        solution) and unpacks the results into the reference's dictionaries."""
        N = self.N
        self.ref_path.update(self.phi_current)       # window slide of get_parameters (the builder kernel then finds nothing to slide)
        st, sector, prev = self.builder_state(np.asarray(q0, float), np.asarray(dq0, float), np.asarray(ddq0, float),
                                              np.asarray(p0, float), np.asarray(v0, float), np.asarray(x_phi_d, float),
                                              np.asarray(jerk_current, float))
        if self._tab_for is not self.ref_path:
            self._tab, self._tab_for = np.ascontiguousarray(self.ref_path.path_table()[None]), self.ref_path
        ec_in = self.error_count
        t0 = time.perf_counter()
        r = self.solver.mpc_step_host(self._tab, self._zero_id, [sector], st[None], prev[None], [ec_in], self.log)
        time_elapsed = time.perf_counter() - t0
        status, iters, ec = int(r["status"][0]), int(r["iters"][0]), int(r["error_count"][0])
        self.solver.set_stats(iters, status)
        so = r["state"][0]
        accepted = ec == 0
        assert int(r["sector"][0]) == self.ref_path.sector
        self.error_count = ec
        if so[73] != 0.0:
            self.prev_solution = r["prev"][0]
        if accepted:
            self.prev_infeasible_solution = r["x"][0]
        else:
            print(f"[ERROR] Could not find feasible solution. Using previous solution. Error count: {ec}")
            self.prev_infeasible_solution = r["x"][0]
        if ec >= N:
            return None, None, None, None, None
        self.phi_prev = np.copy(self.phi_current)
        self.phi_current, self.dphi_current = np.array([so[40]]), np.array([so[41]])
        self.ddphi_current, self.dddphi_current = np.array([so[42]]), np.array([so[43]])
        self.pr_ref, self.iw_ref = so[44:47].copy(), so[47:50].copy()
        M = N - ec
        T = r["traj"][0][:M]
        W = (r["x"][0] if accepted else prev).reshape(N, -1)
        traj_data = dict(p=T[:, 0:6].T.copy(), v=T[:, 6:12].T.copy(), a=T[:, 12:18].T.copy(), q=T[:, 18:25].T.copy(),
                         dq=T[:, 25:32].T.copy(), ddq=T[:, 32:39].T.copy(), dddq=W[ec:, :7].T.copy(), phi=T[:, 39].copy(),
                         dphi=T[:, 40].copy(), ddphi=T[:, 41].copy(), dddphi=W[ec:, 7].copy())
        ref_data = err_data = None
        if self.log:
            R, E = r["ref"][0][:M], r["err"][0][:M]
            ref_data, err_data = defaultdict(list), defaultdict(list)
            for key, a, b in self.REF_COLS:
                ref_data[key] = [R[i, a:b].copy() for i in range(M)]
            for k_, key in enumerate(self.ERR_KEYS):
                err_data[key] = [E[i, 3 * k_:3 * k_ + 3].copy() for i in range(M)]
        return traj_data, ref_data, err_data, time_elapsed, iters

    def step_mirror(self, q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current, x_des=None):
        """One MPC step with the numpy mirror of the reference's pre- / post-processing around `solver(x0=, p=)`."""
        w0, params, aux = self.prepare(q0, dq0, ddq0, p0, v0, x_phi_d, jerk_current)
        t0 = time.perf_counter()
        sol = self.solver(x0=w0, lbx=self.lbu, ubx=self.ubu, lbg=self.lbg, ubg=self.ubg, p=params)
        w_curr = np.array(sol['x']).flatten()  # noqa: F841  (the reference times this conversion too)
        time_elapsed = time.perf_counter() - t0
        return self.finish(sol, self.solver.stats(), aux, time_elapsed)

    # ------------------------------------------------------------------ BoundMPC.py:508-770
    def compute_return_data(self, w_opt, using_previous, aux):
        N, h, ec = self.N, self.dt, self.error_count
        W = np.array(w_opt, float).reshape(N, -1).T          # [44, N]
        q0, dq0, ddq0, p0 = aux['q0'], aux['dq0'], aux['ddq0'], aux['p0']
        phi_switch, p_ref, dp_ref = aux['phi_switch'], aux['p_ref'], aux['dp_ref']
        optimal_jerk = W[:7, ec:]
        optimal_jerk_phi = W[7, ec:]
        M = optimal_jerk.shape[1]
        # re-integration of the joint / path-parameter state from the jerks (general hat-function
        # formulas of the reference == rolling the one-step recurrence)
        jm = np.concatenate((aux['jerk_current'][:, None], optimal_jerk), axis=1)
        jmp = np.concatenate((self.dddphi_current, optimal_jerk_phi))[None, :]
        oq, odq, oddq = np.empty((7, M)), np.empty((7, M)), np.empty((7, M))
        ophi, odphi, oddphi = np.empty(M), np.empty(M), np.empty(M)
        q, dq, ddq = q0, dq0, ddq0
        ph, dph, ddph = self.phi_current, self.dphi_current, self.ddphi_current
        self.phi_prev = np.copy(self.phi_current)
        for i in range(M):
            q, dq, ddq = integrate_jerk(jm[:, i:i + 2], q, dq, ddq, h)
            ph, dph, ddph = integrate_jerk(jmp[:, i:i + 2], ph, dph, ddph, h)
            oq[:, i], odq[:, i], oddq[:, i] = q, dq, ddq
            ophi[i], odphi[i], oddphi[i] = ph[0], dph[0], ddph[0]
        # Cartesian state of the solver's own joint trajectory (used by the re-planning warm start)
        self.prev_traj = W[29:35, :].copy()
        self.prev_vel = W[35:41, :].copy()
        self.prev_acc = np.empty_like(self.prev_vel)
        self.prev_jerk = np.empty_like(self.prev_vel)
        rm = self.robot_model
        for i in range(N):
            qi, dqi, ddqi = W[8:15, i], W[15:22, i], W[22:29, i]
            jac, djac = rm.jacobian_fk(qi), rm.djacobian_fk(qi, dqi)
            self.prev_acc[:, i] = jac @ ddqi + djac @ dqi
            self.prev_jerk[:, i] = jac @ W[:7, i] + djac @ ddqi + rm.ddjacobian_fk(qi, dqi, ddqi) @ dqi
        otraj, ovel, oacc = np.empty((6, M)), np.empty((6, M)), np.empty((6, M))
        oiw = np.empty((6, M))
        omega_prev = (rm.jacobian_fk(q0) @ dq0)[3:]
        for i in range(M):
            pc, jac, djac = rm.forward_kinematics(oq[:, i], odq[:, i])
            otraj[:, i] = pc
            ovel[:, i] = jac @ odq[:, i]
            oacc[:, i] = jac @ oddq[:, i] + djac @ odq[:, i]
            base = oiw[3:, i - 1] if i > 0 else p0[3:]
            oiw[3:, i] = base + 0.5 * h * (omega_prev + ovel[3:, i])
            omega_prev = ovel[3:, i].copy()
        oiw[:3] = otraj[:3]
        if ec > 0 and np.linalg.norm(p0[3:] - oiw[3:, 0]) > 3.1:
            oiw[3:] *= -1
        # rotation reference bookkeeping (BoundMPC.py:593-604)
        iw_ref_copy = self.iw_ref.copy()
        if ophi[0] > phi_switch[1]:
            pr = log_so3(self.ref_path.r[self.ref_path.sector + 1])
            self.pr_ref = integrate_rotation_reference(pr, dp_ref[3:, 1], phi_switch[1], ophi[0])
            self.iw_ref = p_ref[3:, 1] + (ophi[0] - phi_switch[1]) * dp_ref[3:, 1]
        else:
            self.pr_ref = integrate_rotation_reference(self.pr_ref, dp_ref[3:, 0], self.phi_current, ophi[0])
            self.iw_ref = p_ref[3:, 0] + (ophi[0] - phi_switch[0]) * dp_ref[3:, 0]
        self.phi_current = np.array([ophi[0]])
        self.dphi_current = np.array([odphi[0]])
        self.ddphi_current = np.array([oddphi[0]])
        self.dddphi_current = np.array([optimal_jerk_phi[0]])
        ref_data = err_data = None
        if self.log:
            ref_data, err_data = self._log_data(aux, ophi, odphi, oiw, ovel, otraj, iw_ref_copy)
        traj_data = dict(p=otraj, v=ovel, a=oacc, q=oq, dq=odq, ddq=oddq, dddq=optimal_jerk, phi=ophi, dphi=odphi,
                         ddphi=oddphi, dddphi=optimal_jerk_phi)
        return traj_data, ref_data, err_data

    # ------------------------------------------------------------------ logging branch (BoundMPC.py:614-755)
    def _reference(self, aux, phi):
        """reference_function (bound_mpc_functions.py:43-149) on numbers."""
        S = self.nr_segs
        ps = aux['phi_switch']
        i = _segment(phi, ps, S)
        jb = min(i, S - 2)
        r = i if phi < ps[S] else S
        tau = phi - ps[i]
        dpd = aux['dp_ref'][:, i]
        pd = aux['p_ref'][:, i] + dpd * tau
        a4, a3, a2, a1, a0 = aux['a']
        b = (((a4[r] * tau + a3[r]) * tau + a2[r]) * tau + a1[r]) * tau + a0[r]
        v1, v2, v3 = aux['v']
        return dict(i=i, p_d=pd, dp_d=dpd, ddp_d=0 * dpd, dp_normed_d=aux['dp_normed_ref'][:, i],
                    bp1=aux['bp1'][:, jb], bp2=aux['bp2'][:, jb], br1=aux['br1'][:, i], br2=aux['br2'][:, i],
                    v1=v1[:, i], v2=v2[:, i], v3=v3[:, i], bound_lower=np.array([b[2], b[3], b[6], b[7]]),
                    bound_upper=np.array([b[0], b[1], b[4], b[5]]), r_par_bound=np.array([b[8]]),
                    e_p_off=0.5 * np.array([b[0] + b[2], b[1] + b[3]]), e_r_off=0.5 * np.array([b[4] + b[6], b[5] + b[7]]))

    def _log_data(self, aux, ophi, odphi, oiw, ovel, otraj, iw_ref0):
        ref_data, err_data = defaultdict(list), defaultdict(list)
        p0 = aux['p0']
        jl, jr = aux['jac_l'], aux['jac_r']
        d0 = self.dtau_init[:, 0]
        for i in range(ophi.shape[0]):
            R = self._reference(aux, ophi[i])
            s = R['i']
            t, wr = R['dp_d'][:3], R['dp_d'][3:]
            for k_, key in (('p', 'p_d'), ('dp', 'dp_d'), ('ddp', 'ddp_d'), ('dp_normed', 'dp_normed_d')):
                ref_data[k_].append(np.array(R[key], float).copy())
            for key in ('r_par_bound', 'bound_lower', 'bound_upper', 'e_p_off', 'e_r_off', 'bp1', 'bp2', 'br1', 'br2', 'v1', 'v2', 'v3'):
                ref_data[key].append(R[key])
            # error_function (bound_mpc_functions.py:152-202)
            e_p = oiw[:3, i] - R['p_d'][:3]
            e_p_par = (t @ e_p) * t
            de_p = ovel[:3, i] - t * odphi[i]
            de_p_par = (t @ de_p) * t
            e_r = d0 + jl @ (oiw[3:, i] - p0[3:]) - jr @ (R['p_d'][3:] - iw_ref0)
            de_r = jl @ ovel[3:, i] - jr @ (wr * odphi[i])
            dd = e_r - d0
            err_data['e_p'].append(e_p)
            err_data['de_p'].append(de_p)
            err_data['e_p_par'].append(e_p_par)
            err_data['e_p_orth'].append(e_p - e_p_par)
            err_data['de_p_par'].append(de_p_par)
            err_data['de_p_orth'].append(de_p - de_p_par)
            err_data['e_r'].append(e_r.copy())
            err_data['de_r'].append(de_r)
            err_data['e_r_par'].append(self.dtau_init_par[:, s] + (dd @ R['v2']) * R['dp_normed_d'])
            err_data['e_r_orth1'].append(self.dtau_init_orth1[:, s] + (dd @ R['v1']) * R['br1'])
            err_data['e_r_orth2'].append(self.dtau_init_orth2[:, s] + (dd @ R['v3']) * R['br2'])
        # exact rotation reference / error along the horizon (BoundMPC.py:716-752)
        ps, dp_ref = aux['phi_switch'], aux['dp_ref']
        pr = self.pr_ref.copy()
        ref_data['p'][0][3:] = pr
        n_h = self.N - self.error_count
        for i in range(n_h - 1):
            ref_data['p'][i][3:] = pr
            err_data['e_r'][i] = log_so3(exp_so3(otraj[3:, i]) @ exp_so3(pr).T)
            phi, nxt = ophi[i], ophi[i + 1]
            if nxt > ps[1] and phi < ps[1]:
                pr = integrate_rotation_reference(log_so3(self.ref_path.r[self.ref_path.sector + 1]), dp_ref[3:, 1], ps[1], nxt)
            elif nxt > ps[2] and phi < ps[2]:
                pr = integrate_rotation_reference(log_so3(self.ref_path.r[self.ref_path.sector + 2]), dp_ref[3:, 2], ps[2], nxt)
            elif nxt > ps[2]:
                pr = integrate_rotation_reference(pr, dp_ref[3:, 2], phi, nxt)
            elif nxt > ps[1]:
                pr = integrate_rotation_reference(pr, dp_ref[3:, 1], phi, nxt)
            else:
                pr = integrate_rotation_reference(pr, dp_ref[3:, 0], phi, nxt)
        ref_data['p'][-1][3:] = pr
        err_data['e_r'][-1] = log_so3(exp_so3(otraj[3:, -1]) @ exp_so3(pr).T)
        return ref_data, err_data
