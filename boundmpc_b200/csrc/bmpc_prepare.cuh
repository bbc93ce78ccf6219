// boundmpc_b200 — batched parameter builder: the pre-solve half of `BoundMPC.step`
// (bound_mpc/bound_mpc/BoundMPC/BoundMPC.py:310-443).
//
// For every controller instance it produces the parameter vector p [141 + 91 S] and the warm start
// x0 [44 N] the solver call of BoundMPC.py:446-453 receives:
//   * sliding window over the reference path (ReferencePath.py:190-238: `update`, `get_parameters`,
//     `get_limits`, `get_bound_params`),
//   * warm start: cold start (BoundMPC.py:316-321), shifted previous solution (:373-375) with the
//     "reversing integrated omega" repair (:325-333),
//   * orientation-error linearisation data: `compute_initial_rot_errors` (utils/util_functions.py:11-31)
//     per window segment, `compute_orientation_projection_vectors` (BoundMPC.py:267-304) with the SO(3)
//     inverse Jacobians of utils/lie_functions.py:41-64,
//   * quartic error-bound coefficients: `compute_error_bounds` (BoundMPC.py:219-265) +
//     `compute_bound_params` (mpc_utils_casadi.py:130-137),
//   * weight / path-parameter clamps of BoundMPC.py:397-414 and the parameter order of :416-443.
// The re-projection branch after `update()` (BoundMPC.py:335-369, replanning) needs the kinematic model: bmpc_post.cuh
// (warm_start_updated), called by k_prepare for states with the `updated` flag.
// One thread builds one instance (the work is a few hundred scalar operations on 3-vectors); the functions
// are plain host/device code so that tests/emu runs the same source on the CPU.
#pragma once
#include "bmpc_common.h"

namespace bmpc {

// ---- path table: one row per (padded) path segment j, built once per path by the host mirror
// (boundmpc_b200/reference_path.py `path_table`); all entries fp64
enum {
  PT_P = 0,        // [3] via point p_j                      (ReferencePath.p)
  PT_IW = 3,       // [3] integrated rotation increments      (ReferencePath.iw)
  PT_DPN = 6,      // [3] unit direction dp_j / |dp_j|
  PT_DR = 9,       // [3] angular rate per unit path parameter (ReferencePath.dr)
  PT_CUM = 12,     // [1] path parameter at the end of segment j (cumulative arc length)
  PT_PLO = 13,     // [2] p_lower_j
  PT_PUP = 15,     // [2] p_upper_j
  PT_RLO = 17,     // [2] r_lower_j
  PT_RUP = 19,     // [2] r_upper_j
  PT_BP1 = 21, PT_BP2 = 24, PT_BR1 = 27, PT_BR2 = 30,   // [3] each: error bases
  PT_EB = 33,      // [5] e_p_min, e_r_min, e_p_max, e_r_max, s of segment j
  PT_LOGR = 38,    // [3] rotation vector of the via-point rotation R_j (post-processing: rotation reference at a segment switch)
  PT_ROW = 41
};
// ---- per-instance controller state (doubles)
enum {
  PS_Q = 0, PS_DQ = 7, PS_DDQ = 14, PS_P0 = 21, PS_V0 = 27, PS_JERK = 33,
  PS_PHI = 40,       // [4] phi, dphi, ddphi, dddphi of the controller (BoundMPC.phi_current ...)
  PS_PRREF = 44,     // [3] pr_ref
  PS_IWREF = 47,     // [3] iw_ref
  PS_XPHID = 50,     // [3] x_phi_d
  PS_BSCALE = 53,    // [4] per-instance factors on e_p_min, e_r_min, e_p_max, e_r_max (1 = the path's values)
  PS_PHIMAX = 57,    // [1] BoundMPC.phi_max = path length - 1e-4
  PS_W = 58,         // [15] weights
  PS_HASPREV = 73,   // [1] != 0: prev_x holds the previous solution (warm start), else cold start
  PS_UPDATED = 74,   // [1] != 0: the path has been replaced (BoundMPC.update): the warm start is re-projected on the new path
                     //     (BoundMPC.py:335-369) instead of shifted; like the reference's flag it is never reset
  PS_SIZE = 76
};

BMPC_HD void m3_mul(const double* A, const double* B, double* C) {      // C = A B (row-major 3 x 3)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
BMPC_HD void m3_mul_bt(const double* A, const double* B, double* C) {   // C = A B^T
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
BMPC_HD void m3_vec(const double* A, const double* v, double* r) {
  for (int i = 0; i < 3; i++) r[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
BMPC_HD double v3_norm(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
BMPC_HD void skew_sq(const double* a, double* K, double* K2) {
  K[0] = 0; K[1] = -a[2]; K[2] = a[1]; K[3] = a[2]; K[4] = 0; K[5] = -a[0]; K[6] = -a[1]; K[7] = a[0]; K[8] = 0;
  m3_mul(K, K, K2);
}
// rotation vector -> matrix (lie.py exp_so3)
BMPC_HD void so3_exp(const double* rv, double* R) {
  double K[9], K2[9];
  skew_sq(rv, K, K2);
  const double th = v3_norm(rv);
  double c1, c2;
  if (th < 1e-8) { c1 = 1.0; c2 = 0.5; }
  else { c1 = sin(th) / th; c2 = (1.0 - cos(th)) / (th * th); }
  for (int i = 0; i < 9; i++) R[i] = ((i & 3) == 0 ? 1.0 : 0.0) + c1 * K[i] + c2 * K2[i];
}
// rotation matrix -> rotation vector through the unit quaternion (lie.py log_so3; the branch structure of
// scipy's Rotation.from_matrix(...).as_rotvec() the reference calls)
BMPC_HD void so3_log(const double* R, double* rv) {
  const double d3 = R[0] + R[4] + R[8];
  const double d[4] = {R[0], R[4], R[8], d3};
  int c = 0;
  for (int i = 1; i < 4; i++) if (d[i] > d[c]) c = i;
  double q[4];
  if (c != 3) {
    const int i = c, j = (c + 1) % 3, k = (c + 2) % 3;
    q[i] = 1 - d3 + 2 * R[3 * i + i];
    q[j] = R[3 * j + i] + R[3 * i + j];
    q[k] = R[3 * k + i] + R[3 * i + k];
    q[3] = R[3 * k + j] - R[3 * j + k];
  } else {
    q[0] = R[7] - R[5]; q[1] = R[2] - R[6]; q[2] = R[3] - R[1]; q[3] = 1 + d3;
  }
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= nq;
  if (q[3] < 0) for (int i = 0; i < 4; i++) q[i] = -q[i];
  const double s = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  const double ang = 2.0 * atan2(s, q[3]);
  double scale;
  if (ang <= 1e-3) { const double a2 = ang * ang; scale = 2 + a2 / 12 + 7 * a2 * a2 / 2880; }
  else scale = ang / sin(ang / 2);
  for (int i = 0; i < 3; i++) rv[i] = scale * q[i];
}
// inverse left (sgn = -1) / right (sgn = +1) Jacobian of SO(3) with the reference's regularised angle
// (lie_functions.py:41-64)
BMPC_HD void so3_jac_inv(const double* a, double sgn, double* J) {
  double K[9], K2[9];
  skew_sq(a, K, K2);
  const double th = v3_norm(a) + 1e-6;
  const double cf = 1.0 / (th * th) - (1 + cos(th)) / (2 * th * sin(th));
  for (int i = 0; i < 9; i++) J[i] = ((i & 3) == 0 ? 1.0 : 0.0) + sgn * 0.5 * K[i] + cf * K2[i];
}
// scipy Rotation.from_matrix(R).as_euler('zyx') (lie.py euler_zyx_intrinsic_from_matrix)
BMPC_HD void euler_zyx(const double* R, double* e) {
  const double sb = fmin(fmax(R[2], -1.0), 1.0);
  e[1] = asin(sb);
  if (fabs(sb) < 1 - 1e-12) { e[0] = atan2(-R[1], R[0]); e[2] = atan2(-R[5], R[8]); }
  else { e[2] = 0.0; e[0] = atan2(R[3], R[4]); }
}
BMPC_HD void m3_inv(const double* A, double* Ai) {
  const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c0 + A[1] * c1 + A[2] * c2);
  Ai[0] = c0 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c1 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c2 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

constexpr int PREP_SMAX = 8;   // window segments supported by the builder

// Warm start, entry `a` of stage `k` (BoundMPC.py:316-333,373-375): cold start = zeros with q0 and p0 in every stage;
// otherwise the previous solution shifted by one stage, after the "reversing integrated omega" repair of p_rot when the
// measured orientation vector has jumped (|p0_rot - prev p_rot_0| > 1.5).
BMPC_HD bool warm_reverse(const double* st, const double* prev) {
  double dd = 0.0;
  for (int i = 0; i < 3; i++) { const double e = st[PS_P0 + 3 + i] - prev[oPROT + i]; dd += e * e; }
  return sqrt(dd) > 1.5;
}
BMPC_HD double warm_start_value(int N, const double* st, const double* prev, bool reverse, int k, int a) {
  if (st[PS_HASPREV] == 0.0) return (a >= oQ && a < oQ + 7) ? st[PS_Q + a - oQ] : ((a >= oPPOS && a < oPPOS + 6) ? st[PS_P0 + a - oPPOS] : 0.0);
  const int src = k + 1 < N ? k + 1 : N - 1;                               // shift by one stage
  if (reverse && a >= oPROT && a < oPROT + 3) {
    // the repair rewrites rows 0..N-2 from rows 1..N-1 and copies row N-2 to row N-1 BEFORE the shift
    const int r = src < N - 1 ? src : N - 2;
    return st[PS_P0 + 3 + a - oPROT] + (prev[NX * (r + 1) + a] - prev[a]);
  }
  return prev[NX * src + a];
}

// Parameter vector of one instance.  `tab` = path table [J][PT_ROW], `st` = state [PS_SIZE].  Writes p [141 + 91 S];
// returns the window position (sector) after the slide.
BMPC_HD int prepare_params(const PLayout& L, const double* tab, int J, int sector, const double* st, double* p) {
  const int S = L.S;
  const double phi = st[PS_PHI];
  // ---- slide the window while the path parameter is past its first switching point (ReferencePath.py:190-212)
  while (sector + S < J && phi > tab[(size_t)sector * PT_ROW + PT_CUM]) sector++;
  const double* row0 = tab + (size_t)sector * PT_ROW;
  double phisw[PREP_SMAX + 1];
  phisw[0] = sector == 0 ? 0.0 : row0[PT_CUM - PT_ROW];
  for (int i = 0; i < S; i++) phisw[i + 1] = row0[i * PT_ROW + PT_CUM];
  const double* q0 = st + PS_Q;
  const double* p0 = st + PS_P0;
  // ---- orientation-error linearisation data (BoundMPC.py:379-389, util_functions.py:11-31)
  double Rp[9], Rr[9], Rd[9], dtau[3];
  so3_exp(p0 + 3, Rp);
  so3_exp(st + PS_PRREF, Rr);
  m3_mul_bt(Rp, Rr, Rd);
  so3_log(Rd, dtau);
  double Rdt[9];
  so3_exp(dtau, Rdt);
  double par[PREP_SMAX][3], o1[PREP_SMAX][3], o2[PREP_SMAX][3], dpnr[PREP_SMAX][3];
  for (int i = 0; i < S; i++) {
    const double* row = row0 + i * PT_ROW;
    const double* dr = row + PT_DR;
    const double n = v3_norm(dr);
    double axis[3];
    for (int k = 0; k < 3; k++) axis[k] = n > 1e-4 ? dr[k] / n : (k == 1 ? 1.0 : 0.0);
    for (int k = 0; k < 3; k++) dpnr[i][k] = axis[k];                      // ReferencePath.compute_normed_velocity
    const double* br1 = row + PT_BR1;
    const double* br2 = row + PT_BR2;
    double r01[9], T[9], M[9], e[3];
    for (int k = 0; k < 3; k++) { r01[3 * k] = br2[k]; r01[3 * k + 1] = axis[k]; r01[3 * k + 2] = br1[k]; }
    m3_mul(Rdt, r01, T);
    for (int a = 0; a < 3; a++)                                            // M = r01^T T
      for (int b = 0; b < 3; b++) M[3 * a + b] = r01[a] * T[b] + r01[3 + a] * T[3 + b] + r01[6 + a] * T[6 + b];
    euler_zyx(M, e);
    for (int k = 0; k < 3; k++) { par[i][k] = e[1] * axis[k]; o1[i][k] = e[0] * br1[k]; o2[i][k] = e[2] * br2[k]; }
  }
  // ---- projection vectors (BoundMPC.py:267-304): dual basis of (Jr br1, Jr(log rest1) dp_normed, Jr(log rest2) br2)
  double jr[9], jl[9], v1[PREP_SMAX][3], v2[PREP_SMAX][3], v3[PREP_SMAX][3];
  so3_jac_inv(dtau, 1.0, jr);
  so3_jac_inv(dtau, -1.0, jl);
  for (int i = 0; i < S; i++) {
    const double* row = row0 + i * PT_ROW;
    double E1[9], E2[9], rest1[9], rest2[9], l1[3], l2[3], J1[9], J2[9], a[3], b[3], c[3], A[9], Ai[9];
    so3_exp(o1[i], E1);
    so3_exp(par[i], E2);
    m3_mul_bt(Rdt, E1, rest1);
    m3_mul_bt(rest1, E2, rest2);
    so3_log(rest1, l1);
    so3_log(rest2, l2);
    so3_jac_inv(l1, 1.0, J1);
    so3_jac_inv(l2, 1.0, J2);
    m3_vec(jr, row + PT_BR1, a);
    m3_vec(J1, dpnr[i], b);
    m3_vec(J2, row + PT_BR2, c);
    for (int k = 0; k < 3; k++) { A[3 * k] = a[k]; A[3 * k + 1] = b[k]; A[3 * k + 2] = c[k]; }
    m3_inv(A, Ai);
    for (int k = 0; k < 3; k++) { v1[i][k] = Ai[k]; v2[i][k] = Ai[3 + k]; v3[i][k] = Ai[6 + k]; }
  }
  // ---- assemble p in the order of BoundMPC.py:416-443 (layout: casadi_ocp_formulation.py:361-376)
  for (int i = 0; i < 7; i++) { p[L.q0 + i] = q0[i]; p[L.dq0 + i] = st[PS_DQ + i]; p[L.ddq0 + i] = st[PS_DDQ + i]; p[L.jerk + i] = st[PS_JERK + i]; }
  for (int i = 0; i < 3; i++) { p[L.phi0 + i] = st[PS_PHI + i]; p[L.iwref + i] = st[PS_IWREF + i]; p[L.dtau + i] = dtau[i]; }
  p[L.jerk + 7] = st[PS_PHI + 3];
  for (int i = 0; i < 6; i++) { p[L.p0 + i] = p0[i]; p[L.v0 + i] = st[PS_V0 + i]; }
  for (int i = 0; i < S; i++)
    for (int k = 0; k < 3; k++) { p[L.par + 3 * i + k] = par[i][k]; p[L.orth1 + 3 * i + k] = o1[i][k]; p[L.orth2 + 3 * i + k] = o2[i][k]; }
  for (int i = 0; i <= S; i++) p[L.phisw + i] = phisw[i];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { p[L.jacr + 3 * c + r] = jr[3 * r + c]; p[L.jacl + 3 * c + r] = jl[3 * r + c]; }   // (stored column-major)
  for (int i = 0; i < S; i++) {
    const double* row = row0 + i * PT_ROW;
    for (int k = 0; k < 3; k++) {
      p[L.pref + k * S + i] = row[PT_P + k]; p[L.pref + (3 + k) * S + i] = row[PT_IW + k];
      p[L.dpref + k * S + i] = row[PT_DPN + k]; p[L.dpref + (3 + k) * S + i] = row[PT_DR + k];
      p[L.dpn + k * S + i] = dpnr[i][k];
      p[L.bp1 + k * S + i] = row[PT_BP1 + k]; p[L.bp2 + k * S + i] = row[PT_BP2 + k];
      p[L.br1 + k * S + i] = row[PT_BR1 + k]; p[L.br2 + k * S + i] = row[PT_BR2 + k];
      p[L.v1 + k * S + i] = v1[i][k]; p[L.v2 + k * S + i] = v2[i][k]; p[L.v3 + k * S + i] = v3[i][k];
    }
  }
  // ---- quartic error bounds (BoundMPC.py:219-265, mpc_utils_casadi.py:130-137); row S of the tables stays zero
  {
    const double ep = row0[PT_EB] * st[PS_BSCALE], er = row0[PT_EB + 1] * st[PS_BSCALE + 1];
    const double epm = row0[PT_EB + 2] * st[PS_BSCALE + 2], erm = row0[PT_EB + 3] * st[PS_BSCALE + 3], s0 = row0[PT_EB + 4];
    for (int j = 0; j < 9; j++) {
      const bool rot = j >= 4;
      const double sg = (j == 2 || j == 3 || j == 6 || j == 7) ? -1.0 : 1.0;
      const double e0 = sg * (rot ? er : ep), em0 = sg * (rot ? erm : epm), sv0 = sg * s0;
      for (int i = 0; i <= S; i++) {
        double a4 = 0, a3 = 0, a2 = 0, a1 = 0, a0 = 0;
        if (i < S) {
          const double* row = row0 + i * PT_ROW;
          // asym rows: up0, up1, -lo0, -lo1 (position), up2, up3, -lo2, -lo3 (orientation); entry 8 repeats entry 7
          const int jj = j < 8 ? j : 7;
          const double scale = jj < 2 ? row[PT_PUP + jj] : (jj < 4 ? -row[PT_PLO + jj - 2] : (jj < 6 ? row[PT_RUP + jj - 4] : -row[PT_RLO + jj - 6]));
          const double ph = phisw[i + 1] - phisw[i], s = sv0 * scale, emax = em0 * scale;
          a0 = e0; a1 = s;
          a2 = -(4 * ph * s + ph * s + 11 * e0 + 5 * e0 - 16 * emax) / (ph * ph);
          a3 = (5 * ph * s + 3 * ph * s + 18 * e0 + 14 * e0 - 32 * emax) / (ph * ph * ph);
          a4 = -2 * (2 * ph * s + 4 * e0 + 4 * e0 - 8 * emax) / (ph * ph * ph * ph);
        }
        const int o = j * (S + 1) + i;
        p[L.a4 + o] = a4; p[L.a3 + o] = a3; p[L.a2 + o] = a2; p[L.a1 + o] = a1; p[L.a0 + o] = a0;
      }
    }
  }
  // ---- weights, path-parameter clamps, joint reference (BoundMPC.py:397-414)
  const double phimax_path = st[PS_PHIMAX];
  for (int i = 0; i < 15; i++) p[L.w + i] = st[PS_W + i];
  if (st[PS_XPHID] < 1) p[L.w + 6] = st[PS_W + 6] * fmin(1.0 / (phimax_path * phimax_path), 2.0);
  const double phimax = fmin(phi + 5.0, phimax_path);
  p[L.phimax] = phimax;
  p[L.dphimax] = st[PS_W + 4];
  p[L.xphid] = fmin(phi + 5.0, st[PS_XPHID]);
  p[L.xphid + 1] = st[PS_XPHID + 1];
  p[L.xphid + 2] = st[PS_XPHID + 2];
  const bool hold = phimax - phi < 0.05;
  for (int i = 0; i < 7; i++) p[L.qd + i] = hold ? q0[i] : 0.0;
  return sector;
}

// serial form (host emulation, tests): parameters and warm start of one instance
BMPC_HD int prepare_instance(const PLayout& L, int N, const double* tab, int J, int sector, const double* st, const double* prev,
                             double* x0, double* p) {
  const bool rev = st[PS_HASPREV] != 0.0 && warm_reverse(st, prev);
  for (int k = 0; k < N; k++)
    for (int a = 0; a < NX; a++) x0[NX * k + a] = warm_start_value(N, st, prev, rev, k, a);
  return prepare_params(L, tab, J, sector, st, p);
}

}  // namespace bmpc
