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

// ---- logging branch of compute_return_data (BoundMPC.py:614-755): reference data and error terms per node.
// Everything it needs beyond the trajectory is in the parameter vector p the builder produced (reference window, bound
// coefficients, projection vectors, SO(3) Jacobians, initial rotation errors).
enum { RF_P = 0, RF_DP = 6, RF_DDP = 12, RF_DPN = 18, RF_RPAR = 21, RF_LO = 22, RF_UP = 26, RF_EPOFF = 30, RF_EROFF = 32,
       RF_BP1 = 34, RF_BP2 = 37, RF_BR1 = 40, RF_BR2 = 43, RF_V1 = 46, RF_V2 = 49, RF_V3 = 52, RF_ROW = 55 };
enum { ER_EP = 0, ER_DEP = 3, ER_EPPAR = 6, ER_EPORTH = 9, ER_DEPPAR = 12, ER_DEPORTH = 15, ER_ER = 18, ER_DER = 21, ER_ERPAR = 24,
       ER_ERO1 = 27, ER_ERO2 = 30, ER_ROW = 33 };

// `prn` = rotation reference after this step's update (state_out[PS_PRREF]); traj = output of post_instance
BMPC_DEV void log_instance(const Config& C, const double* tab, int sector, const double* st, const double* p, const double* traj, int ec,
                           const double* prn, double* ref, double* err) {
  const PLayout& L = C.L;
  const int N = C.N, M = N - ec, S = L.S;
  const double h = C.dt;
  const double* ps = p + L.phisw;
  double jl[9], jr[9], d0[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { jl[3 * r + c] = p[L.jacl + 3 * c + r]; jr[3 * r + c] = p[L.jacr + 3 * c + r]; }
  for (int a = 0; a < 3; a++) d0[a] = p[L.dtau + a];
  const double* p0 = st + PS_P0;
  // integrated angular velocity along the horizon (BoundMPC.py:573-585); omega before the first node = J(q0) dq0
  double om_prev[3], iw[3];
  {
    double pose[6], v0[6], a0[6];
    rm_node(st + PS_Q, st + PS_DQ, st + PS_DDQ, pose, v0, a0);
    for (int a = 0; a < 3; a++) { om_prev[a] = v0[3 + a]; iw[a] = p0[3 + a]; }
  }
  double sign = 1.0;
  for (int i = 0; i < N; i++) {
    double* R = ref + (size_t)i * RF_ROW;
    double* E = err + (size_t)i * ER_ROW;
    if (i >= M) { for (int a = 0; a < RF_ROW; a++) R[a] = 0.0; for (int a = 0; a < ER_ROW; a++) E[a] = 0.0; continue; }
    const double* T = traj + (size_t)i * TR_ROW;
    for (int a = 0; a < 3; a++) { iw[a] += 0.5 * h * (om_prev[a] + T[TR_V + 3 + a]); om_prev[a] = T[TR_V + 3 + a]; }
    if (i == 0 && ec > 0) {                                    // BoundMPC.py:586-588 (the flip acts on the whole horizon)
      double dd = 0.0;
      for (int a = 0; a < 3; a++) dd += (p0[3 + a] - iw[a]) * (p0[3 + a] - iw[a]);
      if (sqrt(dd) > 3.1) sign = -1.0;
    }
    const double oiw[3] = {sign * iw[0], sign * iw[1], sign * iw[2]};
    const double phi = T[TR_PHI], dphi = T[TR_PHI + 1];
    // reference_function on numbers (bound_mpc_functions.py:43-149; index rules SURVEY App. B.1-B.3)
    int si = S - 1;
    for (int q = S - 2; q >= 0; q--) if (phi < ps[q + 1]) si = q;
    const int jb = si < S - 2 ? si : S - 2, rr = phi < ps[S] ? si : S;
    const double tau = phi - ps[si];
    double b[9];
    for (int j = 0; j < 9; j++) {
      const int o = j * (S + 1) + rr;
      b[j] = (((p[L.a4 + o] * tau + p[L.a3 + o]) * tau + p[L.a2 + o]) * tau + p[L.a1 + o]) * tau + p[L.a0 + o];
    }
    double pd[6], dpd[6];
    for (int k = 0; k < 6; k++) { dpd[k] = p[L.dpref + k * S + si]; pd[k] = p[L.pref + k * S + si] + dpd[k] * tau; }
    for (int k = 0; k < 6; k++) { R[RF_P + k] = pd[k]; R[RF_DP + k] = dpd[k]; R[RF_DDP + k] = 0.0; }
    double dpn[3], v1[3], v2[3], v3[3], br1[3], br2[3];
    for (int k = 0; k < 3; k++) {
      dpn[k] = p[L.dpn + k * S + si]; v1[k] = p[L.v1 + k * S + si]; v2[k] = p[L.v2 + k * S + si]; v3[k] = p[L.v3 + k * S + si];
      br1[k] = p[L.br1 + k * S + si]; br2[k] = p[L.br2 + k * S + si];
      R[RF_DPN + k] = dpn[k]; R[RF_BP1 + k] = p[L.bp1 + k * S + jb]; R[RF_BP2 + k] = p[L.bp2 + k * S + jb];
      R[RF_BR1 + k] = br1[k]; R[RF_BR2 + k] = br2[k]; R[RF_V1 + k] = v1[k]; R[RF_V2 + k] = v2[k]; R[RF_V3 + k] = v3[k];
    }
    R[RF_RPAR] = b[8];
    R[RF_LO] = b[2]; R[RF_LO + 1] = b[3]; R[RF_LO + 2] = b[6]; R[RF_LO + 3] = b[7];
    R[RF_UP] = b[0]; R[RF_UP + 1] = b[1]; R[RF_UP + 2] = b[4]; R[RF_UP + 3] = b[5];
    R[RF_EPOFF] = 0.5 * (b[0] + b[2]); R[RF_EPOFF + 1] = 0.5 * (b[1] + b[3]);
    R[RF_EROFF] = 0.5 * (b[4] + b[6]); R[RF_EROFF + 1] = 0.5 * (b[5] + b[7]);
    // error_function (bound_mpc_functions.py:152-202)
    const double* t = dpd;
    const double* wr = dpd + 3;
    double ep[3], dep[3], x1[3], x2[3], er[3], der[3];
    for (int k = 0; k < 3; k++) { ep[k] = T[TR_P + k] - pd[k]; dep[k] = T[TR_V + k] - t[k] * dphi; }
    const double te = dot3(t, ep), tde = dot3(t, dep);
    for (int k = 0; k < 3; k++) { x1[k] = oiw[k] - p0[3 + k]; x2[k] = pd[3 + k] - st[PS_IWREF + k]; }
    double y1[3], y2[3], y3[3], y4[3], wd[3];
    m3_vec(jl, x1, y1); m3_vec(jr, x2, y2);
    for (int k = 0; k < 3; k++) wd[k] = wr[k] * dphi;
    m3_vec(jl, T + TR_V + 3, y3); m3_vec(jr, wd, y4);
    double dd[3];
    for (int k = 0; k < 3; k++) { er[k] = d0[k] + y1[k] - y2[k]; der[k] = y3[k] - y4[k]; dd[k] = er[k] - d0[k]; }
    const double s1 = dot3(dd, v1), s2 = dot3(dd, v2), s3 = dot3(dd, v3);
    for (int k = 0; k < 3; k++) {
      E[ER_EP + k] = ep[k]; E[ER_DEP + k] = dep[k];
      E[ER_EPPAR + k] = te * t[k]; E[ER_EPORTH + k] = ep[k] - te * t[k];
      E[ER_DEPPAR + k] = tde * t[k]; E[ER_DEPORTH + k] = dep[k] - tde * t[k];
      E[ER_ER + k] = er[k]; E[ER_DER + k] = der[k];
      E[ER_ERPAR + k] = p[L.par + 3 * si + k] + s2 * dpn[k];
      E[ER_ERO1 + k] = p[L.orth1 + 3 * si + k] + s1 * br1[k];
      E[ER_ERO2 + k] = p[L.orth2 + 3 * si + k] + s3 * br2[k];
    }
  }
  // exact rotation reference / error along the horizon (BoundMPC.py:716-752)
  const double* row0 = tab + (size_t)sector * PT_ROW;
  double pr[3] = {prn[0], prn[1], prn[2]};
  for (int i = 0; i < M; i++) {
    double* R = ref + (size_t)i * RF_ROW;
    double* E = err + (size_t)i * ER_ROW;
    const double* T = traj + (size_t)i * TR_ROW;
    double Ra[9], Rb[9], Rc[9];
    for (int a = 0; a < 3; a++) R[RF_P + 3 + a] = pr[a];
    so3_exp(T + TR_P + 3, Ra);
    so3_exp(pr, Rb);
    m3_mul_bt(Ra, Rb, Rc);
    so3_log(Rc, E + ER_ER);
    if (i + 1 < M) {
      const double phi = T[TR_PHI], nxt = T[TR_ROW + TR_PHI];
      double out[3];
      if (nxt > ps[1] && phi < ps[1]) integrate_rot_ref(row0 + PT_ROW + PT_LOGR, row0 + PT_ROW + PT_DR, ps[1], nxt, out);
      else if (nxt > ps[2] && phi < ps[2]) integrate_rot_ref(row0 + 2 * PT_ROW + PT_LOGR, row0 + 2 * PT_ROW + PT_DR, ps[2], nxt, out);
      else if (nxt > ps[2]) integrate_rot_ref(pr, row0 + 2 * PT_ROW + PT_DR, phi, nxt, out);
      else if (nxt > ps[1]) integrate_rot_ref(pr, row0 + PT_ROW + PT_DR, phi, nxt, out);
      else integrate_rot_ref(pr, row0 + PT_DR, phi, nxt, out);
      for (int a = 0; a < 3; a++) pr[a] = out[a];
    }
  }
}

// ---- replanning: `BoundMPC.update` (BoundMPC.py:163-217) and the re-projected warm start (:335-369)
// J_v(q) xa and dJ_v(q, dq) xb (linear rows of the geometric Jacobian and of its time derivative along dq)
BMPC_DEV void rm_lin(const double* q, const double* dq, const double* xa, const double* xb, double* Jxa, double* dJxb) {
  double z[7][3], o[7][3], p[3], R[9];
  rm_chain(q, z, o, p, R);
  double Om[8][3] = {{0, 0, 0}};
  for (int i = 0; i < 7; i++) for (int a = 0; a < 3; a++) Om[i + 1][a] = Om[i][a] + dq[i] * z[i][a];
  for (int a = 0; a < 3; a++) { Jxa[a] = 0.0; dJxb[a] = 0.0; }
  double tail[3] = {0, 0, 0};
  for (int i = 6; i >= 0; i--) {
    double r[3], zr[3], zd[3], rd[3], c1[3], c2[3], t2[3];
    for (int a = 0; a < 3; a++) r[a] = p[a] - o[i][a];
    cross3(z[i], r, zr);
    cross3(Om[i], z[i], zd);
    for (int a = 0; a < 3; a++) tail[a] += dq[i] * zr[a];
    cross3(Om[i], r, t2);
    for (int a = 0; a < 3; a++) rd[a] = tail[a] + t2[a];
    cross3(zd, r, c1);
    cross3(z[i], rd, c2);
    for (int a = 0; a < 3; a++) { Jxa[a] += zr[a] * xa[i]; dJxb[a] += (c1[a] + c2[a]) * xb[i]; }
  }
}

// Warm start after a path update: the previous solution (after the "reversing integrated omega" repair, not shifted) with
// its path-parameter entries re-projected on the first segment of the new path; Cartesian acceleration / jerk of the
// previous trajectory as in compute_return_data (BoundMPC.py:563-571: J ddq + dJ dq and J u + dJ ddq + ddJ dq with the
// central-difference ddJ of the host model).  `row0` = first window row of the new path, `ps0`, `ps1` = phi_switch[0], [1].
BMPC_DEV void warm_start_updated(const Config& C, const double* st, const double* prev, bool reverse, const double* row0, double ps0,
                                 double ps1, double* x0) {
  const int N = C.N;
  const double* t = row0 + PT_DPN;
  const double* pr = row0 + PT_P;
  for (int i = 0; i < N; i++) {
    const double* w = prev + NX * i;
    double* o = x0 + NX * i;
    for (int a = 0; a < NX; a++) o[a] = w[a];
    if (reverse) {
      const int r = i < N - 1 ? i : N - 2;
      for (int a = 0; a < 3; a++) o[oPROT + a] = st[PS_P0 + 3 + a] + (prev[NX * (r + 1) + oPROT + a] - prev[oPROT + a]);
    }
    const double d[3] = {w[oPPOS] - pr[0], w[oPPOS + 1] - pr[1], w[oPPOS + 2] - pr[2]};
    const double phik = ps0 + dot3(d, t);
    if (phik > ps1 - 0.01) { o[oPHI] = ps1 - 0.01; o[oDPHI] = 0.0; o[oDDPHI] = 0.0; }
    else if (phik < 0) {
      for (int j = 0; j < 7; j++) o[oQ + j] = st[PS_Q + j];
      for (int a = 0; a < 6; a++) o[oPPOS + a] = st[PS_P0 + a];
      for (int a = 0; a < 4; a++) o[oVLIN + a] = 0.0;                  // (entries 35..38, as the reference writes them)
      o[oPHI] = 0.0; o[oDPHI] = 0.0; o[oDDPHI] = 0.0;
    } else {
      const double* q = w + oQ;
      const double* dq = w + oDQ;
      const double* ddq = w + oDDQ;
      double Jddq[3], dJdq[3], Ju[3], dJddq[3], up[3], um[3], dummy[3];
      rm_lin(q, dq, ddq, dq, Jddq, dJdq);
      rm_lin(q, dq, w + oU, ddq, Ju, dJddq);
      const double eps = 1e-6;
      double qp[7], qm[7], dqp[7], dqm[7];
      for (int j = 0; j < 7; j++) {
        qp[j] = q[j] + eps * dq[j] + 0.5 * eps * eps * ddq[j]; qm[j] = q[j] - eps * dq[j] + 0.5 * eps * eps * ddq[j];
        dqp[j] = dq[j] + eps * ddq[j]; dqm[j] = dq[j] - eps * ddq[j];
      }
      rm_lin(qp, dqp, dq, dq, dummy, up);
      rm_lin(qm, dqm, dq, dq, dummy, um);
      double acc[3], jrk[3];
      for (int a = 0; a < 3; a++) { acc[a] = Jddq[a] + dJdq[a]; jrk[a] = Ju[a] + dJddq[a] + (up[a] - um[a]) / (2 * eps); }
      o[oPHI] = phik;
      o[oDPHI] = dot3(w + oVLIN, t);
      o[oDDPHI] = dot3(acc, t);
      o[oUPHI] = dot3(jrk, t);
    }
  }
}

// window rows / switching points the re-projection needs, for the (already slid) window position `sector`
BMPC_DEV void warm_start_updated_at(const Config& C, const double* tab, int sector, const double* st, const double* prev, double* x0) {
  const double* row0 = tab + (size_t)sector * PT_ROW;
  const double ps0 = sector == 0 ? 0.0 : row0[PT_CUM - PT_ROW], ps1 = row0[PT_CUM];
  warm_start_updated(C, st, prev, warm_reverse(st, prev), row0, ps0, ps1, x0);
}

// BoundMPC.update for one controller: the measured Cartesian state cart = (p0(6), v(6), a(6), jerk(6)) projected on the
// first segment of the new path `tab` (row 0), rotation reference restarted at the first via point.
BMPC_DEV void update_state(const double* tab, double path_phimax, const double* cart, double* st) {
  const double* row0 = tab;
  const double* t = row0 + PT_DPN;
  const double d[3] = {cart[0] - row0[PT_P], cart[1] - row0[PT_P + 1], cart[2] - row0[PT_P + 2]};
  const double phi = dot3(d, t);
  st[PS_PHI] = phi;
  st[PS_PHI + 1] = dot3(cart + 6, t);
  st[PS_PHI + 2] = dot3(cart + 12, t);
  st[PS_PHI + 3] = dot3(cart + 18, t);
  double prn[3];
  integrate_rot_ref(row0 + PT_LOGR, row0 + PT_DR, 0.0, phi, prn);
  for (int a = 0; a < 3; a++) { st[PS_PRREF + a] = prn[a]; st[PS_IWREF + a] = row0[PT_IW + a] + phi * row0[PT_DR + a]; }
  st[PS_PHIMAX] = path_phimax - 0.0001;
  st[PS_UPDATED] = 1.0;
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
