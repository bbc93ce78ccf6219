// boundmpc_b200 — batched post-processing: `compute_return_data` of `BoundMPC.step`
// (bound_mpc/bound_mpc/BoundMPC/BoundMPC.py:508-611,757-770), without the logging branch.
//
// From the trajectory the controller keeps (the solution x, or the previous one after a failed solve) it
// produces, per instance,
//   * the re-integrated joint and path-parameter trajectory q, dq, ddq, phi, dphi, ddphi of the remaining
//     horizon (general hat-function formulas of jerk_trajectory_casadi.py:46-175 == rolling the one-step
//     recurrence from the measured state with the applied jerk, BoundMPC.py:528-556),
//   * its Cartesian image p = fk(q) (position, rotation vector), v = J dq, a = J ddq + dJ dq with the geometric
//     Jacobian and its time derivative (RobotModel.py:254-373 and its d/dt; BoundMPC.py:563-579),
//   * the controller state of the next step: path-parameter state, rotation reference pr_ref and iw_ref
//     (BoundMPC.py:593-611, utils/util_functions.py:88-99),
// i.e. everything `traj_data` returns except the jerk columns, which are copies of x.  Kinematics are the serial
// chain of bmpc_model.cuh (same constants); one thread post-processes one instance.
#pragma once
#include "bmpc_model.cuh"
#include "bmpc_prepare.cuh"

namespace bmpc {

// trajectory record per horizon node (doubles)
enum { TR_P = 0, TR_V = 6, TR_A = 12, TR_Q = 18, TR_DQ = 25, TR_DDQ = 32, TR_PHI = 39, TR_ROW = 42 };

// serial chain: joint axes z_i, joint origins o_i, tool point p, tool rotation R (robot_model.py `chain`)
BMPC_DEV void rm_chain(const double* q, double (*z)[3], double (*o)[3], double* p, double* R) {
  double Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, oc[3] = {0, 0, 0};
  for (int i = 0; i < 7; i++) {
    double Rn[9];
    for (int a = 0; a < 3; a++) oc[a] += Rc[3 * a] * kJXYZ[i][0] + Rc[3 * a + 1] * kJXYZ[i][1] + Rc[3 * a + 2] * kJXYZ[i][2];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rn[3 * a + b] = Rc[3 * a] * kJROT[i][0][b] + Rc[3 * a + 1] * kJROT[i][1][b] + Rc[3 * a + 2] * kJROT[i][2][b];
    for (int a = 0; a < 3; a++) { z[i][a] = Rn[3 * a + 2]; o[i][a] = oc[a]; }
    double sn, cs;
    sincos(q[i], &sn, &cs);
    for (int a = 0; a < 3; a++) {
      Rc[3 * a] = Rn[3 * a] * cs + Rn[3 * a + 1] * sn;
      Rc[3 * a + 1] = Rn[3 * a + 1] * cs - Rn[3 * a] * sn;
      Rc[3 * a + 2] = Rn[3 * a + 2];
    }
  }
  for (int a = 0; a < 3; a++) { p[a] = oc[a] + Rc[3 * a + 2] * kTOOLZ; }
  for (int a = 0; a < 9; a++) R[a] = Rc[a];
}

// pose (position, rotation vector), J dq, J ddq + dJ dq of one node (RobotModel.forward_kinematics + the products of
// BoundMPC.py:566-571; dJ as in robot_model.py `_rates`: d z_i/dt = Om_i x z_i, d(p - o_i)/dt = tail_i + Om_i x r_i)
BMPC_DEV void rm_node(const double* q, const double* dq, const double* ddq, double* pose, double* vel, double* acc) {
  double z[7][3], o[7][3], p[3], R[9];
  rm_chain(q, z, o, p, R);
  for (int a = 0; a < 3; a++) pose[a] = p[a];
  so3_log(R, pose + 3);
  double Om[8][3] = {{0, 0, 0}};
  for (int i = 0; i < 7; i++) for (int a = 0; a < 3; a++) Om[i + 1][a] = Om[i][a] + dq[i] * z[i][a];
  for (int a = 0; a < 6; a++) { vel[a] = 0.0; acc[a] = 0.0; }
  double tail[3] = {0, 0, 0};
  for (int i = 6; i >= 0; i--) {
    double r[3], zr[3], zd[3], rd[3], c1[3], c2[3], t2[3];
    for (int a = 0; a < 3; a++) r[a] = p[a] - o[i][a];
    cross3(z[i], r, zr);                               // J_v column
    cross3(Om[i], z[i], zd);                           // d z_i / dt
    for (int a = 0; a < 3; a++) tail[a] += dq[i] * zr[a];
    cross3(Om[i], r, t2);
    for (int a = 0; a < 3; a++) rd[a] = tail[a] + t2[a];
    cross3(zd, r, c1);
    cross3(z[i], rd, c2);                              // dJ_v column = zd x r + z x rd
    for (int a = 0; a < 3; a++) {
      vel[a] += zr[a] * dq[i];
      vel[3 + a] += z[i][a] * dq[i];
      acc[a] += zr[a] * ddq[i] + (c1[a] + c2[a]) * dq[i];
      acc[3 + a] += z[i][a] * ddq[i] + zd[a] * dq[i];
    }
  }
}

// utils/util_functions.py:88-99 (mirror: bound_mpc.integrate_rotation_reference)
BMPC_DEV void integrate_rot_ref(const double* pr_ref, const double* omega, double phi0, double phi1, double* out) {
  double R0[9];
  so3_exp(pr_ref, R0);
  const double n = v3_norm(omega);
  if (n > 1e-4) {
    double ax[3] = {omega[0] / n, omega[1] / n, omega[2] / n}, K[9], K2[9], Rr[9], Rn[9];
    const double ang = (phi1 - phi0) * n;
    skew_sq(ax, K, K2);
    for (int i = 0; i < 9; i++) Rr[i] = ((i & 3) == 0 ? 1.0 : 0.0) + sin(ang) * K[i] + (1 - cos(ang)) * K2[i];
    m3_mul(Rr, R0, Rn);
    so3_log(Rn, out);
  } else {
    so3_log(R0, out);
  }
}

// One instance.  `w` = the trajectory the controller keeps [44 N], `ec` = error count (nodes already consumed from a
// previous solution, BoundMPC.py:467-496), `st` = controller state of this step, `tab` = path table [J][PT_ROW],
// `sector` = window position after the slide.  Writes the trajectory record traj [N][TR_ROW] (rows >= N - ec zero)
// and the state of the next step st_out [PS_SIZE].
BMPC_DEV void post_instance(const Config& C, const double* tab, int sector, const double* st, const double* w, int ec, double* traj,
                            double* st_out) {
  const int N = C.N, M = N - ec;
  const double h = C.dt;
  double q[7], dq[7], ddq[7], um[8], ph[3];
  for (int i = 0; i < 7; i++) { q[i] = st[PS_Q + i]; dq[i] = st[PS_DQ + i]; ddq[i] = st[PS_DDQ + i]; um[i] = st[PS_JERK + i]; }
  for (int i = 0; i < 3; i++) ph[i] = st[PS_PHI + i];
  um[7] = st[PS_PHI + 3];
  double phi_first[3] = {0, 0, 0}, ujp_first = 0.0;
  for (int i = 0; i < N; i++) {
    double* T = traj + (size_t)i * TR_ROW;
    if (i >= M) { for (int a = 0; a < TR_ROW; a++) T[a] = 0.0; continue; }
    const double* u = w + NX * (ec + i);
    // one sample with piecewise-linear jerk from um to u (App. A.4)
    for (int j = 0; j < 7; j++) {
      const double qn = q[j] + h * dq[j] + h * h / 2 * ddq[j] + h * h * h / 8 * um[j] + h * h * h / 24 * u[j];
      const double dqn = dq[j] + h * ddq[j] + h * h / 3 * um[j] + h * h / 6 * u[j];
      const double ddqn = ddq[j] + h / 2 * (um[j] + u[j]);
      q[j] = qn; dq[j] = dqn; ddq[j] = ddqn; um[j] = u[j];
    }
    {
      const double up = u[oUPHI];
      const double p0 = ph[0] + h * ph[1] + h * h / 2 * ph[2] + h * h * h / 8 * um[7] + h * h * h / 24 * up;
      const double p1 = ph[1] + h * ph[2] + h * h / 3 * um[7] + h * h / 6 * up;
      const double p2 = ph[2] + h / 2 * (um[7] + up);
      ph[0] = p0; ph[1] = p1; ph[2] = p2; um[7] = up;
    }
    if (i == 0) { phi_first[0] = ph[0]; phi_first[1] = ph[1]; phi_first[2] = ph[2]; ujp_first = u[oUPHI]; }
    rm_node(q, dq, ddq, T + TR_P, T + TR_V, T + TR_A);
    for (int j = 0; j < 7; j++) { T[TR_Q + j] = q[j]; T[TR_DQ + j] = dq[j]; T[TR_DDQ + j] = ddq[j]; }
    T[TR_PHI] = ph[0]; T[TR_PHI + 1] = ph[1]; T[TR_PHI + 2] = ph[2];
  }
  // ---- controller state of the next step (BoundMPC.py:593-611)
  for (int i = 0; i < PS_SIZE; i++) st_out[i] = st[i];
  const double* row0 = tab + (size_t)sector * PT_ROW;
  const double ps0 = sector == 0 ? 0.0 : row0[PT_CUM - PT_ROW], ps1 = row0[PT_CUM];
  double prn[3];
  if (phi_first[0] > ps1) {
    const double* row1 = row0 + PT_ROW;
    integrate_rot_ref(row1 + PT_LOGR, row1 + PT_DR, ps1, phi_first[0], prn);
    for (int a = 0; a < 3; a++) st_out[PS_IWREF + a] = row1[PT_IW + a] + (phi_first[0] - ps1) * row1[PT_DR + a];
  } else {
    integrate_rot_ref(st + PS_PRREF, row0 + PT_DR, st[PS_PHI], phi_first[0], prn);
    for (int a = 0; a < 3; a++) st_out[PS_IWREF + a] = row0[PT_IW + a] + (phi_first[0] - ps0) * row0[PT_DR + a];
  }
  for (int a = 0; a < 3; a++) { st_out[PS_PRREF + a] = prn[a]; st_out[PS_PHI + a] = phi_first[a]; }
  st_out[PS_PHI + 3] = ujp_first;
}

// ---- second half of `BoundMPC.step` as a whole (BoundMPC.py:454-506 + compute_return_data) and the closed-loop advance of
// bound_mpc_node.py:321-331,362.  Decision per instance:
//   success = solver success or summed constraint violation beyond 1e-6 below 1e-4 (BoundMPC.py:461-465)
//   success            -> keep x, it becomes the previous solution, error_count = 0
//   failure, has prev  -> keep the previous solution, error_count + 1 (its nodes error_count.. are used)
//   failure, no prev   -> keep x anyway, error_count = 0, no previous solution recorded
// returns 0 / 1 / 2 for these cases.
BMPC_DEV double constraint_violation(const Config& C, const double* g) {
  double v = 0.0;
  for (int i = 0; i < C.m; i++) {
    const int r = i % NG;
    const double gi = g[i];
    if (gi > 1e-6) v += gi;                       // above ubg = 0
    else if (r < NE && gi < -1e-6) v -= gi;       // below lbg = 0 (equality rows; lbg = -inf on the inequality rows)
  }
  return v;
}
BMPC_DEV int finish_decision(const Config& C, int status, const double* g, bool has_prev) {
  const bool success = status == ST_SUCCESS || constraint_violation(C, g) < 1e-4;
  return success ? 0 : (has_prev ? 1 : 2);
}
// The controller state after the robot has moved by one sample under the first jerk of the kept trajectory
// (integrate_joint, utils/util_functions.py:152-161, = node 0 of the post-processed trajectory): joint state, pose,
// Cartesian velocity and applied jerk replace the measured ones in the next-step state of post_instance.
BMPC_DEV void advance_state(const double* w, int ec, const double* traj, double* st_out) {
  for (int j = 0; j < 7; j++) {
    st_out[PS_Q + j] = traj[TR_Q + j]; st_out[PS_DQ + j] = traj[TR_DQ + j]; st_out[PS_DDQ + j] = traj[TR_DDQ + j];
    st_out[PS_JERK + j] = w[NX * ec + oU + j];
  }
  for (int a = 0; a < 6; a++) { st_out[PS_P0 + a] = traj[TR_P + a]; st_out[PS_V0 + a] = traj[TR_V + a]; }
}

}  // namespace bmpc
