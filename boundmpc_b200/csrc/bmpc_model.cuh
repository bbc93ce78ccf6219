// boundmpc_b200 — per-horizon-node evaluation of the OCP functions (kernel family (a)).
//
// Replaces what CasADi generates from casadi_ocp_formulation.py:88-357 (nlp_f, nlp_g,
// nlp_grad_f, nlp_jac_g, nlp_hess_l).  All derivatives are
// hand-derived:
//   * kinematics of the serial chain (RobotModel.py:62-100,1055-1107,1270-1303) through the
//     identity  d^m p / dq_{i1}..dq_{im} = z_{i1} x (z_{i2} x ... (z_{im} x (p - o_{im})))
//     for sorted joint indices, which gives the geometric Jacobian, the derivative of
//     J(q) dq and the multiplier-weighted second/third-order contractions in O(1) cross
//     products per joint pair;
//   * reference / error / bound functions (bound_mpc_functions.py:43-202,
//     mpc_utils_casadi.py:6-165), which are affine in (p_pos, p_rot) and polynomial in phi
//     inside one path segment; the segment selectors have zero derivative
//     (bound_mpc_functions.py:13-40).
#pragma once
#include "bmpc_common.h"

namespace bmpc {

// chain constants (urdf/body/iiwa14.xacro:65,104,143,182,221,260,299,340); every joint frame
// rotation is a signed permutation matrix, stored as R[k][row][col]
#ifdef BMPC_HOST_EMU
#define BMPC_CONST static const
#else
#define BMPC_CONST __device__ __constant__
#endif
BMPC_CONST double kJXYZ[7][3] = {{0, 0, 0.1525}, {0, 0, 0.2075}, {0, 0.2325, 0}, {0, 0, 0.1875},
                                 {0, 0.2125, 0}, {0, 0, 0.1875}, {0, 0.0796, 0}};
BMPC_CONST double kJROT[7][3][3] = {
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}},  {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}}, {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}},
    {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}}, {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}}, {{1, 0, 0}, {0, 0, -1}, {0, 1, 0}},
    {{-1, 0, 0}, {0, 0, 1}, {0, 1, 0}}};
constexpr double kTOOLZ = 0.2174;

// per-CTA workspace (all pointers into one global allocation, see bmpc_kernels.cu)
struct Work {
  double *x, *xt, *dx, *gradf, *gh, *zL, *zU, *dzL, *dzU;  // [n]
  double *y, *ynew, *c, *ct;                                // [36 N]
  double *s, *st, *zs, *ds, *dzs, *d, *dtr;                 // [12 N]
  double *rec;                                              // [N][R_SIZE]
  double *fk;                                               // [2 N][F_SIZE] (shared memory when it fits, see work_attach_smem)
  double *prec; int prec_stride;                            // path part of the records: prec + k * prec_stride + R_x, R_x >= R_HY
  double *Kk;                                               // [N][8*44]
  double *kap;                                              // [N][8]
  double *cost;                                             // [N]
  double *sig;                                              // [n] bound part of the barrier Hessian: z_L / (x - l) + z_U / (u - x)
  double *wp0;                                              // [44] stage-0 "previous block" built from p
};

BMPC_HD size_t work_doubles(int N) {
  size_t n = (size_t)NX * N, ne = (size_t)NE * N, nd = (size_t)ND * N;
  return 9 * n + 4 * ne + 7 * nd + (size_t)N * R_SIZE + (size_t)2 * N * F_SIZE + (size_t)N * 8 * NX + (size_t)N * 8 + N + NX + n;
}
BMPC_DEV void work_carve(Work& W, double* base, int N) {
  size_t n = (size_t)NX * N, ne = (size_t)NE * N, nd = (size_t)ND * N;
  double* q = base;
  W.x = q; q += n; W.xt = q; q += n; W.dx = q; q += n; W.gradf = q; q += n; W.gh = q; q += n;
  W.zL = q; q += n; W.zU = q; q += n; W.dzL = q; q += n; W.dzU = q; q += n;
  W.y = q; q += ne; W.ynew = q; q += ne; W.c = q; q += ne; W.ct = q; q += ne;
  W.s = q; q += nd; W.st = q; q += nd; W.zs = q; q += nd; W.ds = q; q += nd; W.dzs = q; q += nd;
  W.d = q; q += nd; W.dtr = q; q += nd;
  W.rec = q; q += (size_t)N * R_SIZE;
  W.fk = q; q += (size_t)2 * N * F_SIZE;
  W.prec = W.rec; W.prec_stride = R_SIZE;
  W.Kk = q; q += (size_t)N * 8 * NX;
  W.kap = q; q += (size_t)N * 8;
  W.cost = q; q += N;
  W.sig = q; q += n;
  W.wp0 = q; q += NX;
}

// previous-stage block of stage 0: u_{-1}, x_0 come from the parameter vector (App. A.4)
BMPC_DEV void build_wp0(const Ctx cx, const Config& C, const double* p, double* wp) {
  const PLayout& L = C.L;
  PAR_FOR(i, NX) {
    double v = 0.0;
    if (i < 8) v = BMPC_LDG(p + L.jerk + i);
    else if (i < 15) v = BMPC_LDG(p + L.q0 + (i - 8));
    else if (i < 22) v = BMPC_LDG(p + L.dq0 + (i - 15));
    else if (i < 29) v = BMPC_LDG(p + L.ddq0 + (i - 22));
    else if (i < 35) v = BMPC_LDG(p + L.p0 + (i - 29));
    else if (i < 41) v = BMPC_LDG(p + L.v0 + (i - 35));
    else v = BMPC_LDG(p + L.phi0 + (i - 41));
    wp[i] = v;
  }
}

BMPC_DEV const double* prev_block(const Work& W, const double* x, int k) { return k == 0 ? W.wp0 : x + NX * (k - 1); }

// ---------------------------------------------------------------------------------------------
// Phase 1: one-step integration with piecewise-linear jerk (bound_mpc_functions.py:254-260,
// jerk_trajectory_casadi.py at t = h) -> linear residual rows and the chain inputs
BMPC_DEV void phase_integrate(const Ctx cx, const Config& C, const Work& W, const double* x, double* c) {
  PAR_FOR(it, C.N * 8) {
    const int k = it >> 3, j = it & 7;
    const double* wp = prev_block(W, x, k);
    const double* w = x + NX * k;
    if (j < 7) {
      const double q = wp[oQ + j], dq = wp[oDQ + j], ddq = wp[oDDQ + j], um = wp[oU + j], u = w[oU + j];
      const double qn = q + C.a_dq * dq + C.a_ddq * ddq + C.a_um * um + C.a_u * u;
      const double dqn = dq + C.b_ddq * ddq + C.b_um * um + C.b_u * u;
      const double ddqn = ddq + C.c_um * um + C.c_u * u;
      c[NE * k + j] = qn - w[oQ + j];
      c[NE * k + 7 + j] = dqn - w[oDQ + j];
      c[NE * k + 14 + j] = ddqn - w[oDDQ + j];
      double* f0 = W.fk + (size_t)(2 * k) * F_SIZE;
      double* f1 = f0 + F_SIZE;
#pragma unroll 1
      for (int h = 0; h < 2; h++) {   // (one copy of sincos)
        double sn, cs;
        sincos(h ? q : qn, &sn, &cs);
        double* f = h ? f1 : f0;
        f[F_SN + j] = sn; f[F_CS + j] = cs; f[F_DQ + j] = h ? dq : dqn;
      }
    } else {
      const double ph = wp[oPHI], dph = wp[oDPHI], ddph = wp[oDDPHI], um = wp[oUPHI], u = w[oUPHI];
      c[NE * k + 33] = ph + C.a_dq * dph + C.a_ddq * ddph + C.a_um * um + C.a_u * u - w[oPHI];
      c[NE * k + 34] = dph + C.b_ddq * ddph + C.b_um * um + C.b_u * u - w[oDPHI];
      c[NE * k + 35] = ddph + C.c_um * um + C.c_u * u - w[oDDPHI];
    }
  }
}

// Phase 2: forward kinematics of one chain evaluation (one thread per chain):
// joint axes z_i, lever arms r_i = p - o_i and the running sums needed by the derivative
// formulas.  `tails` = false skips W / OT (values-only evaluation of the omega(q_k) chain).
BMPC_NOINLINE void fk_chain(double* f) {
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double o[3] = {0, 0, 0};
  double org[7][3], z[7][3], dq[7];
#pragma unroll
  for (int k = 0; k < 7; k++) {
    double Rn[3][3];
    for (int i = 0; i < 3; i++) o[i] += R[i][0] * kJXYZ[k][0] + R[i][1] * kJXYZ[k][1] + R[i][2] * kJXYZ[k][2];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rn[i][j] = R[i][0] * kJROT[k][0][j] + R[i][1] * kJROT[k][1][j] + R[i][2] * kJROT[k][2][j];
    for (int i = 0; i < 3; i++) { z[k][i] = Rn[i][2]; org[k][i] = o[i]; }
    const double sn = f[F_SN + k], cs = f[F_CS + k];
    dq[k] = f[F_DQ + k];
    for (int i = 0; i < 3; i++) {
      R[i][0] = Rn[i][0] * cs + Rn[i][1] * sn;
      R[i][1] = Rn[i][1] * cs - Rn[i][0] * sn;
      R[i][2] = Rn[i][2];
    }
  }
  double pos[3];
  for (int i = 0; i < 3; i++) { pos[i] = o[i] + R[i][2] * kTOOLZ; f[F_POS + i] = pos[i]; }
  // head sums  OH_i = sum_{k<i} dq_k z_k  (OH_7 = omega)
  double acc[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 7; k++)
    for (int i = 0; i < 3; i++) { f[F_Z + 3 * k + i] = z[k][i]; f[F_OH + 3 * k + i] = acc[i]; acc[i] += dq[k] * z[k][i]; }
  for (int i = 0; i < 3; i++) f[F_OH + 21 + i] = acc[i];
  // tail sums  W_i = sum_{k>=i} dq_k z_k x r_k ,  OT_i = sum_{k>i} dq_k z_k
  double wv[3] = {0, 0, 0}, ot[3] = {0, 0, 0};
#pragma unroll
  for (int k = 6; k >= 0; k--) {
    double r[3], cr[3];
    for (int i = 0; i < 3; i++) { r[i] = pos[i] - org[k][i]; f[F_R + 3 * k + i] = r[i]; }
    cross3(z[k], r, cr);
    for (int i = 0; i < 3; i++) {
      f[F_OT + 3 * k + i] = ot[i];
      wv[i] += dq[k] * cr[i];
      f[F_W + 3 * k + i] = wv[i];
      ot[i] += dq[k] * z[k][i];
    }
  }
}

BMPC_DEV void phase_fk(const Ctx cx, const Config& C, const Work& W, int w = 0) {
  ROLE_FOR(it, 2 * C.N, w, w + 1) fk_chain(W.fk + (size_t)it * F_SIZE);
}

// Phase 3a: kinematic residual rows 21..32 (casadi_ocp_formulation.py:284-291,
// bound_mpc_functions.py:262-282: p_pos = fk_pos(q_n), v = [velocity_ee; omega_ee](q_n, dq_n),
// trapezoidal integration of omega into p_rot)
// (the kinematic phases take a warp range [w0, w1): eval_full runs them next to the path terms)
BMPC_DEV void phase_kin_residual(const Ctx cx, const Config& C, const Work& W, const double* x, double* c, int w0, int w1) {
  ROLE_FOR(it, C.N * 3, w0, w1) {
    const int k = it / 3, i = it - 3 * k;
    const double* wp = prev_block(W, x, k);
    const double* w = x + NX * k;
    const double* f0 = W.fk + (size_t)(2 * k) * F_SIZE;
    const double* f1 = f0 + F_SIZE;
    const double hh = 0.5 * C.dt;
    c[NE * k + 21 + i] = f0[F_POS + i] - w[oPPOS + i];
    c[NE * k + 24 + i] = wp[oPROT + i] + hh * (f1[F_OH + 21 + i] + f0[F_OH + 21 + i]) - w[oPROT + i];
    c[NE * k + 27 + i] = f0[F_W + i] - w[oVLIN + i];
    c[NE * k + 30 + i] = f0[F_OH + 21 + i] - w[oVANG + i];
  }
}

// ---------------------------------------------------------------------------------------------
// Path terms of one stage (one thread): reference pose, bound polynomials, error split, blended
// cost, interval-form inequality rows.  MODE 0: values (cost, d).  MODE 1: + derivative record.
// MODE 2: values of the reference-form rows 36..42 (written to gq[7]) for the final report.
struct PathIdx { int i, jb, r; };
BMPC_DEV PathIdx path_segments(double phi, const double* phisw, int S) {
  PathIdx s;
  s.i = S - 1;    // rule B.1: S-row tables, bound_mpc_functions.py:13-20
  for (int q = S - 2; q >= 0; q--) if (phi < BMPC_LDG(phisw + q + 1)) s.i = q;
  s.jb = S - 2;   // rule B.2: first row of "current and next", bound_mpc_functions.py:34-40,109-110
  for (int q = S - 3; q >= 0; q--) if (phi < BMPC_LDG(phisw + q + 1)) s.jb = q;
  s.r = S;        // rule B.3: (S+1)-row coefficient tables, bound_mpc_functions.py:84-88
  for (int q = S - 1; q >= 0; q--) if (phi < BMPC_LDG(phisw + q + 1)) s.r = q;
  return s;
}

// blended error  r = a + sigma * b  (casadi_ocp_formulation.py:237-243): adds w |r|^2, its gradient
// and Hessian w.r.t. y (7).  A, B: Jacobians of a, b w.r.t. y (3 x 7).
BMPC_DEV void blend_terms(double wgt, const double* a, const double (*A)[7], const double* b, const double (*B)[7],
                          double sg, double sg1, double sg2, double& cost, double* GY, double* HY) {
  double r[3], Jr[3][7];
  for (int m = 0; m < 3; m++) {
    r[m] = a[m] + sg * b[m];
    for (int i = 0; i < 7; i++) Jr[m][i] = A[m][i] + sg * B[m][i];
    Jr[m][6] += sg1 * b[m];
    cost += wgt * r[m] * r[m];
  }
  for (int i = 0; i < 7; i++) {
    GY[i] += 2 * wgt * (Jr[0][i] * r[0] + Jr[1][i] * r[1] + Jr[2][i] * r[2]);
    for (int j = 0; j < 7; j++) HY[i * 7 + j] += 2 * wgt * (Jr[0][i] * Jr[0][j] + Jr[1][i] * Jr[1][j] + Jr[2][i] * Jr[2][j]);
  }
  for (int m = 0; m < 3; m++) {
    const double rm = 2 * wgt * r[m];
    for (int i = 0; i < 6; i++) { HY[i * 7 + 6] += rm * sg1 * B[m][i]; HY[6 * 7 + i] += rm * sg1 * B[m][i]; }
    HY[48] += rm * (2 * sg1 * B[m][6] + sg2 * b[m]);
  }
}

BMPC_NOINLINE void path_stage(const int MODE, const Config& C, const double* p, const double* wp, const double* w, double* rec, double* dout,
                              double* cost_out, double* gq, const double* sk, const double* zk) {
  const PLayout& L = C.L;
  const int S = L.S;
  const double phi = w[oPHI], dphi = w[oDPHI], ddphi = w[oDDPHI];
  const PathIdx sg_ = path_segments(phi, p + L.phisw, S);
  const int si = sg_.i, jb = sg_.jb, rr = sg_.r;
  const double tau = phi - BMPC_LDG(p + L.phisw + si);
  double dpd[6], pd[6];
  for (int k = 0; k < 6; k++) { dpd[k] = BMPC_LDG(p + L.dpref + k * S + si); pd[k] = dpd[k] * tau + BMPC_LDG(p + L.pref + k * S + si); }
  // bound polynomials and derivatives (mpc_utils_casadi.py:140-165)
  double b[9], b1[9], b2[9];
  for (int j = 0; j < 9; j++) {
    const int o = j * (S + 1) + rr;
    const double c4 = BMPC_LDG(p + L.a4 + o), c3 = BMPC_LDG(p + L.a3 + o), c2 = BMPC_LDG(p + L.a2 + o), c1 = BMPC_LDG(p + L.a1 + o), c0 = BMPC_LDG(p + L.a0 + o);
    b[j] = (((c4 * tau + c3) * tau + c2) * tau + c1) * tau + c0;
    b1[j] = ((4 * c4 * tau + 3 * c3) * tau + 2 * c2) * tau + c1;
    b2[j] = (12 * c4 * tau + 6 * c3) * tau + 2 * c2;
  }
  double t[3], wr[3], dpn[3], bp1[3], bp2[3], br1[3], br2[3], v1[3], v2[3], v3[3], par0[3], o10[3], o20[3];
  for (int k = 0; k < 3; k++) {
    t[k] = dpd[k]; wr[k] = dpd[3 + k];
    dpn[k] = BMPC_LDG(p + L.dpn + k * S + si);
    bp1[k] = BMPC_LDG(p + L.bp1 + k * S + jb); bp2[k] = BMPC_LDG(p + L.bp2 + k * S + jb);
    br1[k] = BMPC_LDG(p + L.br1 + k * S + si); br2[k] = BMPC_LDG(p + L.br2 + k * S + si);
    v1[k] = BMPC_LDG(p + L.v1 + k * S + si); v2[k] = BMPC_LDG(p + L.v2 + k * S + si); v3[k] = BMPC_LDG(p + L.v3 + k * S + si);
    par0[k] = BMPC_LDG(p + L.par + 3 * si + k); o10[k] = BMPC_LDG(p + L.orth1 + 3 * si + k); o20[k] = BMPC_LDG(p + L.orth2 + 3 * si + k);
  }
  // position error (mpc_utils_casadi.py:19-67)
  double ep[3], eppar[3], eporth[3];
  for (int k = 0; k < 3; k++) ep[k] = w[oPPOS + k] - pd[k];
  const double tt = dot3(t, t), te = dot3(t, ep);
  for (int k = 0; k < 3; k++) { eppar[k] = te * t[k]; eporth[k] = ep[k] - eppar[k]; }
  // orientation error (mpc_utils_casadi.py:6-10); jac_dtau_* are stored column-major
  double Jl[3][3], Jr_[3][3], dl[3], dr[3], dv[3], dph[3];
  for (int r = 0; r < 3; r++)
    for (int c_ = 0; c_ < 3; c_++) { Jl[r][c_] = BMPC_LDG(p + L.jacl + c_ * 3 + r); Jr_[r][c_] = BMPC_LDG(p + L.jacr + c_ * 3 + r); }
  for (int k = 0; k < 3; k++) { dl[k] = w[oPROT + k] - BMPC_LDG(p + L.p0 + 3 + k); dr[k] = pd[3 + k] - BMPC_LDG(p + L.iwref + k); }
  for (int k = 0; k < 3; k++) {
    dv[k] = Jl[k][0] * dl[0] + Jl[k][1] * dl[1] + Jl[k][2] * dl[2] - (Jr_[k][0] * dr[0] + Jr_[k][1] * dr[1] + Jr_[k][2] * dr[2]);
    dph[k] = -(Jr_[k][0] * wr[0] + Jr_[k][1] * wr[1] + Jr_[k][2] * wr[2]);   // d(dv)/d phi
  }
  const double s1 = dot3(dv, v1), sp = dot3(dv, v2), s2 = dot3(dv, v3);
  double er[3], erpar[3], erorth[3];
  for (int k = 0; k < 3; k++) { er[k] = BMPC_LDG(p + L.dtau + k) + dv[k]; erpar[k] = par0[k] + sp * dpn[k]; erorth[k] = er[k] - erpar[k]; }
  // interval rows
  const double nn = dot3(dpn, dpn), nb1 = dot3(br1, br1), nb2 = dot3(br2, br2);
  const double proj = dot3(dpn, par0) + sp * nn;
  const double m0 = dot3(ep, bp1) - 0.5 * (b[0] + b[2]), h0 = 0.5 * (b[0] - b[2]);
  const double m1 = dot3(ep, bp2) - 0.5 * (b[1] + b[3]), h1 = 0.5 * (b[1] - b[3]);
  const double m2 = dot3(br1, o10) + s1 * nb1 - 0.5 * (b[4] + b[6]), h2 = 0.5 * (b[4] - b[6]);
  const double m3 = dot3(br2, o20) + s2 * nb2 - 0.5 * (b[5] + b[7]), h3 = 0.5 * (b[5] - b[7]);
  if (MODE == 2) {
    gq[0] = phi - BMPC_LDG(p + L.phimax);
    gq[1] = dphi - BMPC_LDG(p + L.dphimax);
    gq[2] = proj * proj - b[8] * b[8];
    gq[3] = m0 * m0 - h0 * h0; gq[4] = m1 * m1 - h1 * h1;
    gq[5] = m2 * m2 - h2 * h2; gq[6] = m3 * m3 - h3 * h3;
  }
  dout[0] = phi - BMPC_LDG(p + L.phimax);
  dout[1] = dphi - BMPC_LDG(p + L.dphimax);
  dout[2] = proj - b[8]; dout[3] = -proj - b[8];
  dout[4] = m0 - h0; dout[5] = -m0 - h0;
  dout[6] = m1 - h1; dout[7] = -m1 - h1;
  dout[8] = m2 - h2; dout[9] = -m2 - h2;
  dout[10] = m3 - h3; dout[11] = -m3 - h3;
  // objective (bound_mpc_functions.py:205-246 with the sigmoid blend casadi_ocp_formulation.py:237-243)
  const double* wt = p + L.w;
  const double w0 = BMPC_LDG(wt + 0), w1 = BMPC_LDG(wt + 1), w2 = BMPC_LDG(wt + 2), w5 = BMPC_LDG(wt + 5);
  const double w6 = BMPC_LDG(wt + 6), w7 = BMPC_LDG(wt + 7), w8 = BMPC_LDG(wt + 8), w9 = BMPC_LDG(wt + 9);
  const double w10 = BMPC_LDG(wt + 10), w11 = BMPC_LDG(wt + 11), w12 = BMPC_LDG(wt + 12), w13 = BMPC_LDG(wt + 13);
  const double arg = 100.0 * (phi - (BMPC_LDG(p + L.phimax) - 0.02));
  double sg;
  { const double e = bmpc_exp(-fabs(arg)); sg = arg >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e); }
  const double sg1 = 100.0 * sg * (1.0 - sg), sg2 = 100.0 * sg1 * (1.0 - 2.0 * sg);
  double cost = 0.0;
  double ev[6], fv[6];
  const double idt = 1.0 / C.dt;
  for (int m = 0; m < 6; m++) {
    ev[m] = w[oVLIN + m] - dphi * dpd[m];
    fv[m] = (w[oVLIN + m] - wp[oVLIN + m]) * idt - ddphi * dpd[m];
    cost += w2 * ev[m] * ev[m] + w5 * fv[m] * fv[m];
  }
  const double x0 = BMPC_LDG(p + L.xphid) - phi, x1 = BMPC_LDG(p + L.xphid + 1) - dphi, x2 = BMPC_LDG(p + L.xphid + 2) - ddphi;
  cost += w6 * x0 * x0 + w7 * x1 * x1 + w8 * x2 * x2;
  for (int j = 0; j < 7; j++) {
    const double dq_ = w[oQ + j] - BMPC_LDG(p + L.qd + j);
    cost += w10 * dq_ * dq_ + w11 * w[oDQ + j] * w[oDQ + j] + w12 * w[oDDQ + j] * w[oDDQ + j] + w13 * w[oU + j] * w[oU + j];
  }
  cost += w9 * w[oUPHI] * w[oUPHI];
  if (MODE != 1) {
    for (int m = 0; m < 3; m++) {
      const double rp = eppar[m] + sg * eporth[m], rq = erpar[m] + sg * erorth[m];
      cost += w0 * rp * rp + w1 * rq * rq;
    }
    *cost_out = cost;
    return;
  }
  // ---------------- derivative record
  double GY[7], HY[49];
  for (int i = 0; i < 7; i++) GY[i] = 0.0;
  for (int i = 0; i < 49; i++) HY[i] = 0.0;
  {
    double A[3][7], B[3][7];
    for (int m = 0; m < 3; m++) {
      for (int i = 0; i < 7; i++) { A[m][i] = 0.0; B[m][i] = 0.0; }
      for (int i = 0; i < 3; i++) { A[m][i] = t[m] * t[i]; B[m][i] = (m == i ? 1.0 : 0.0) - t[m] * t[i]; }
      A[m][6] = -tt * t[m];
      B[m][6] = -t[m] + tt * t[m];
    }
    blend_terms(w0, eppar, A, eporth, B, sg, sg1, sg2, cost, GY, HY);
    double jl1[3], jl2[3], jl3[3];   // Jl^T v1, Jl^T v2, Jl^T v3
    for (int i = 0; i < 3; i++) {
      jl1[i] = Jl[0][i] * v1[0] + Jl[1][i] * v1[1] + Jl[2][i] * v1[2];
      jl2[i] = Jl[0][i] * v2[0] + Jl[1][i] * v2[1] + Jl[2][i] * v2[2];
      jl3[i] = Jl[0][i] * v3[0] + Jl[1][i] * v3[1] + Jl[2][i] * v3[2];
    }
    const double dp1 = dot3(dph, v1), dp2 = dot3(dph, v2), dp3 = dot3(dph, v3);
    for (int m = 0; m < 3; m++) {
      for (int i = 0; i < 7; i++) { A[m][i] = 0.0; B[m][i] = 0.0; }
      for (int i = 0; i < 3; i++) { A[m][3 + i] = dpn[m] * jl2[i]; B[m][3 + i] = Jl[m][i] - A[m][3 + i]; }
      A[m][6] = dpn[m] * dp2;
      B[m][6] = dph[m] - A[m][6];
    }
    blend_terms(w1, erpar, A, erorth, B, sg, sg1, sg2, cost, GY, HY);
    // interval-row gradients: y = (p_pos 0..2, p_rot 3..5, phi 6), index 7 = dphi
    double* JD = rec + R_JD;
    double* HD = rec + R_HD;
    for (int i = 0; i < ND * 8; i++) JD[i] = 0.0;
    JD[0 * 8 + 6] = 1.0; HD[0] = 0.0;
    JD[1 * 8 + 7] = 1.0; HD[1] = 0.0;
    const double pg = nn * dp2;
    for (int i = 0; i < 3; i++) { JD[2 * 8 + 3 + i] = nn * jl2[i]; JD[3 * 8 + 3 + i] = -nn * jl2[i]; }
    JD[2 * 8 + 6] = pg - b1[8]; JD[3 * 8 + 6] = -pg - b1[8];
    HD[2] = -b2[8]; HD[3] = -b2[8];
    const double mg0 = -dot3(t, bp1) - 0.5 * (b1[0] + b1[2]), hg0 = 0.5 * (b1[0] - b1[2]);
    const double mg1 = -dot3(t, bp2) - 0.5 * (b1[1] + b1[3]), hg1 = 0.5 * (b1[1] - b1[3]);
    for (int i = 0; i < 3; i++) {
      JD[4 * 8 + i] = bp1[i]; JD[5 * 8 + i] = -bp1[i];
      JD[6 * 8 + i] = bp2[i]; JD[7 * 8 + i] = -bp2[i];
    }
    JD[4 * 8 + 6] = mg0 - hg0; JD[5 * 8 + 6] = -mg0 - hg0;
    JD[6 * 8 + 6] = mg1 - hg1; JD[7 * 8 + 6] = -mg1 - hg1;
    HD[4] = -0.5 * (b2[0] + b2[2]) - 0.5 * (b2[0] - b2[2]); HD[5] = 0.5 * (b2[0] + b2[2]) - 0.5 * (b2[0] - b2[2]);
    HD[6] = -0.5 * (b2[1] + b2[3]) - 0.5 * (b2[1] - b2[3]); HD[7] = 0.5 * (b2[1] + b2[3]) - 0.5 * (b2[1] - b2[3]);
    const double mg2 = nb1 * dp1 - 0.5 * (b1[4] + b1[6]), hg2 = 0.5 * (b1[4] - b1[6]);
    const double mg3 = nb2 * dp3 - 0.5 * (b1[5] + b1[7]), hg3 = 0.5 * (b1[5] - b1[7]);
    for (int i = 0; i < 3; i++) {
      JD[8 * 8 + 3 + i] = nb1 * jl1[i]; JD[9 * 8 + 3 + i] = -nb1 * jl1[i];
      JD[10 * 8 + 3 + i] = nb2 * jl3[i]; JD[11 * 8 + 3 + i] = -nb2 * jl3[i];
    }
    JD[8 * 8 + 6] = mg2 - hg2; JD[9 * 8 + 6] = -mg2 - hg2;
    JD[10 * 8 + 6] = mg3 - hg3; JD[11 * 8 + 6] = -mg3 - hg3;
    HD[8] = -0.5 * (b2[4] + b2[6]) - 0.5 * (b2[4] - b2[6]); HD[9] = 0.5 * (b2[4] + b2[6]) - 0.5 * (b2[4] - b2[6]);
    HD[10] = -0.5 * (b2[5] + b2[7]) - 0.5 * (b2[5] - b2[7]); HD[11] = 0.5 * (b2[5] + b2[7]) - 0.5 * (b2[5] - b2[7]);
  }
  GY[6] += -2 * w6 * x0;
  HY[48] += 2 * w6;
  double se = 0.0, sf = 0.0;
  for (int m = 0; m < 6; m++) {
    rec[R_DPD + m] = dpd[m];
    rec[R_GV + m] = 2 * w2 * ev[m] + 2 * w5 * fv[m] * idt;
    rec[R_GVP + m] = -2 * w5 * fv[m] * idt;
    se += ev[m] * dpd[m]; sf += fv[m] * dpd[m];
  }
  rec[R_GPH] = -2 * w2 * se - 2 * w7 * x1;
  rec[R_GPH + 1] = -2 * w5 * sf - 2 * w8 * x2;
  for (int i = 0; i < 7; i++) rec[R_GY + i] = GY[i];
  for (int i = 0; i < 49; i++) rec[R_HY + i] = HY[i];
  rec[R_COST] = cost;
  for (int r = 0; r < ND; r++) {   // slack terms for phase_path_blocks: Sigma_r, 1 / s_r, Sigma_r (d_r + s_r)
    const double sv = sk[r], sgm = zk[r] / sv;
    rec[R_SIG + r] = sgm; rec[R_SIG + ND + r] = 1.0 / sv; rec[R_SIG + 2 * ND + r] = sgm * (dout[r] + sv);
  }
  *cost_out = cost;
}

// One item per stage, dealt to the lanes of warp `w` (long serial items: they run next to the
// forward-kinematics chains of warp 0 instead of after them).
BMPC_DEV void phase_path(const Ctx cx, const Config& C, const Work& W, const double* p, const double* x, double* d, double* gq, int mode, int w = 0) {
  ROLE_FOR(k, C.N, w, w + 1) {
    path_stage(mode, C, p, prev_block(W, x, k), x + NX * k, W.prec + (size_t)k * W.prec_stride, d + ND * k, W.cost + k,
               mode == 2 ? gq + NQ * k : nullptr, W.s + ND * k, W.zs + ND * k);
  }
}

// y-block of the condensed Hessian, HYB = HY + z.HD + J_d^T Sigma_s J_d + dphi tracking term, and the
// slack part of g^ (Sigma_r = z_r / s_r); 80 independent items per stage, run after path_stage<1>.
BMPC_DEV void phase_path_blocks(const Ctx cx, const Config& C, const Work& W, const double* p) {
  const double w2 = BMPC_LDG(p + C.L.w + 2), w7 = BMPC_LDG(p + C.L.w + 7);
  PAR_FOR(it, C.N * 80) {
    const int k = it / 80, q = it - 80 * k;
    double* rec = W.rec + (size_t)k * R_SIZE;
    const double* pr = W.prec + (size_t)k * W.prec_stride;
    const double* JD = pr + R_JD;
    const double* sg = pr + R_SIG;
    if (q < 64) {
      const int a = q >> 3, b = q & 7;
      double v = (a < 7 && b < 7) ? pr[R_HY + a * 7 + b] : 0.0;
#pragma unroll
      for (int r = 0; r < ND; r++) v += sg[r] * JD[r * 8 + a] * JD[r * 8 + b];
      if (a == 6 && b == 6) for (int r = 0; r < ND; r++) v += W.zs[ND * k + r] * pr[R_HD + r];
      if (a == 7 && b == 7) {
        double nd = 0.0;
        for (int m = 0; m < 6; m++) nd += pr[R_DPD + m] * pr[R_DPD + m];
        v += 2 * w2 * nd + 2 * w7;
      }
      rec[R_HYB + q] = v;
    } else {
      const int a = (q - 64) & 7;
      const double* cf = sg + (q < 72 ? ND : 2 * ND);
      double g = 0.0;
#pragma unroll
      for (int r = 0; r < ND; r++) g += JD[r * 8 + a] * cf[r];
      rec[(q < 72 ? R_GJ1 : R_GJ2) + a] = g;
    }
  }
  // staged path records -> global records (coalesced); HYB / GJ1 / GJ2 were written above
  if (W.prec != W.rec) {
    PAR_FOR(it, C.N * (R_HYB - R_HY + 3 * ND)) {
      const int per = R_HYB - R_HY + 3 * ND, k = it / per, q = it - per * k;
      const int o = q < R_HYB - R_HY ? R_HY + q : R_SIG + (q - (R_HYB - R_HY));
      W.rec[(size_t)k * R_SIZE + o] = W.prec[(size_t)k * W.prec_stride + o];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 4: multiplier-weighted second derivatives of the kinematic rows.  For chain ch = 2k
// (integrated state) the scalar is
//   Phi = lam_p . fk_pos + lam_v . (Jv dq) + (lam_w + h/2 lam_rot) . (Jw dq)
// and for ch = 2k+1 (stage variables) it is  h/2 lam_rot . Jw(q_k) dq_k  (SURVEY App. A.7).
BMPC_DEV void phase_kin_hessian(const Ctx cx, const Config& C, const Work& W, int w0, int w1) {
  ROLE_FOR(it, 2 * C.N * 49, w0, w1) {
    const int ch = it / 49, ij = it - 49 * ch, i = ij / 7, j = ij - 7 * i;
    const int k = ch >> 1, which = ch & 1;
    const double* f = W.fk + (size_t)ch * F_SIZE;
    const double* yk = W.y + NE * k;
    double* rec = W.rec + (size_t)k * R_SIZE;
    const double hh = 0.5 * C.dt;
    double lp[3], mv[3], mw[3];
    for (int a = 0; a < 3; a++) {
      if (which == 0) { lp[a] = yk[21 + a]; mv[a] = yk[27 + a]; mw[a] = yk[30 + a] + hh * yk[24 + a]; }
      else { lp[a] = 0.0; mv[a] = 0.0; mw[a] = hh * yk[24 + a]; }
    }
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const double* zl = f + F_Z + 3 * lo;
    const double* zh = f + F_Z + 3 * hi;
    const double* rh = f + F_R + 3 * hi;
    double zr[3], t1[3];
    cross3(zh, rh, zr);      // z_hi x r_hi
    cross3(zl, zr, t1);      // z_lo x (z_hi x r_hi)  = d2 p / dq_lo dq_hi
    // mixed block  d2 Phi / dq_i d(dq_j)
    double hqd = dot3(mv, t1);
    if (i < j) { double zz[3]; cross3(f + F_Z + 3 * i, f + F_Z + 3 * j, zz); hqd += dot3(mw, zz); }
    (which == 0 ? rec[R_HQDN + ij] : rec[R_HQDK + ij]) = hqd;
    if (i <= j) {
      double hqq = dot3(lp, t1);
      double a1[3], a2[3], a3[3], om[3];
      cross3(zh, f + F_W + 3 * hi, a1); cross3(zl, a1, a2);                 // z_i x (z_j x W_j)
      double acc = dot3(mv, a2);
      for (int a = 0; a < 3; a++) om[a] = f[F_OH + 3 * hi + a] - f[F_OH + 3 * lo + a];   // sum_{i<=k<j} dq_k z_k
      cross3(om, zr, a1); cross3(zl, a1, a2);                               // z_i x (Om_{i:j} x (z_j x r_j))
      acc += dot3(mv, a2);
      cross3(f + F_OH + 3 * lo, t1, a3);                                    // Om_{<i} x (z_i x (z_j x r_j))
      acc += dot3(mv, a3);
      hqq += acc;
      cross3(zh, f + F_OT + 3 * hi, a1); cross3(zl, a1, a2);                // z_i x (z_j x Om_{>j})
      hqq += dot3(mw, a2);
      double* H = which == 0 ? rec + R_HQQN : rec + R_HQQK;
      H[i * 7 + j] = hqq;
      H[j * 7 + i] = hqq;
    }
  }
}

// Phase 5: kinematic rows of [A_hat | B] (first derivatives).  One item per (stage, joint).
//   d pos / dq_i = z_i x r_i ;  d(Jv dq)/dq_i = z_i x W_i + Om_{<i} x (z_i x r_i) ;  d(Jw dq)/dq_i = z_i x Om_{>i}
// Structural zeros and the identity of the p_rot columns of GK: the sparsity pattern never changes, so this
// runs once per CTA workspace (kernel start), not per evaluation.
BMPC_DEV void phase_kin_jacobian_init(const Ctx cx, const Config& C, const Work& W) {
  PAR_FOR(it, C.N * NZ) {
    const int k = it / NZ, col = it - NZ * k;
    double* GK = W.rec + (size_t)k * R_SIZE + R_GK;
    for (int r = 0; r < NK; r++) GK[r * NZ + col] = 0.0;
    if (col >= oPROT && col < oPROT + 3) GK[(3 + col - oPROT) * NZ + col] = 1.0;
  }
}
BMPC_DEV void phase_kin_jacobian(const Ctx cx, const Config& C, const Work& W, int w0, int w1) {
  ROLE_FOR(it, C.N * 7, w0, w1) {
    const int k = it / 7, j = it - 7 * k;
    const double* f0 = W.fk + (size_t)(2 * k) * F_SIZE;
    const double* f1 = f0 + F_SIZE;
    double* GK = W.rec + (size_t)k * R_SIZE + R_GK;
    const double hh = 0.5 * C.dt;
    double Kq[12], Kd[12], Kqk[3], Kdk[3];
    {
      const double* z = f0 + F_Z + 3 * j;
      double jv[3], a1[3], a2[3], dw[3];
      cross3(z, f0 + F_R + 3 * j, jv);
      cross3(z, f0 + F_W + 3 * j, a1);
      cross3(f0 + F_OH + 3 * j, jv, a2);
      cross3(z, f0 + F_OT + 3 * j, dw);
      for (int a = 0; a < 3; a++) {
        Kq[a] = jv[a]; Kq[3 + a] = hh * dw[a]; Kq[6 + a] = a1[a] + a2[a]; Kq[9 + a] = dw[a];
        Kd[a] = 0.0; Kd[3 + a] = hh * z[a]; Kd[6 + a] = jv[a]; Kd[9 + a] = z[a];
      }
      const double* zk = f1 + F_Z + 3 * j;
      double dwk[3];
      cross3(zk, f1 + F_OT + 3 * j, dwk);
      for (int a = 0; a < 3; a++) { Kqk[a] = hh * dwk[a]; Kdk[a] = hh * zk[a]; }
    }
    for (int r = 0; r < NK; r++) {
      const double kq = Kq[r], kd = Kd[r];
      const double eq = (r >= 3 && r < 6) ? Kqk[r - 3] : 0.0, ed = (r >= 3 && r < 6) ? Kdk[r - 3] : 0.0;
      GK[r * NZ + oU + j] = C.a_um * kq + C.b_um * kd;
      GK[r * NZ + oQ + j] = kq + eq;
      GK[r * NZ + oDQ + j] = C.a_dq * kq + kd + ed;
      GK[r * NZ + oDDQ + j] = C.a_ddq * kq + C.b_ddq * kd;
      GK[r * NZ + NX + oU + j] = C.a_u * kq + C.b_u * kd;
    }
  }
}

// Phase 6: gradient of the objective (nlp_grad_f)
BMPC_DEV void phase_grad_f(const Ctx cx, const Config& C, const Work& W, const double* p, const double* x, double* gradf) {
  const PLayout& L = C.L;
  const double* wt = p + L.w;
  PAR_FOR(it, C.n) {
    const int k = it / NX, i = it - NX * k;
    const double* rec = W.prec + (size_t)k * W.prec_stride;
    const double v = x[it];
    double g;
    if (i < 7) g = 2 * BMPC_LDG(wt + 13) * v;
    else if (i == 7) g = 2 * BMPC_LDG(wt + 9) * v;
    else if (i < 15) g = 2 * BMPC_LDG(wt + 10) * (v - BMPC_LDG(p + L.qd + (i - 8)));
    else if (i < 22) g = 2 * BMPC_LDG(wt + 11) * v;
    else if (i < 29) g = 2 * BMPC_LDG(wt + 12) * v;
    else if (i < 35) g = rec[R_GY + (i - 29)];
    else if (i < 41) {
      g = rec[R_GV + (i - 35)];
      if (k + 1 < C.N) g += rec[W.prec_stride + R_GVP + (i - 35)];
    } else if (i == 41) g = rec[R_GY + 6];
    else g = rec[R_GPH + (i - 42)];
    gradf[it] = g;
  }
}

// ---------------------------------------------------------------------------------------------
// Whole-horizon evaluation.  full = derivative records too (needs the current multipliers in W.y).
// The two long serial pieces — the kinematic chains (2 N lanes of warp 0) and the path terms (N lanes
// of warp 1) — run side by side; everything else is dealt to all threads.
BMPC_NOINLINE void eval_values(const Ctx cx, const Config& C, const Work& W, const double* p, const double* x, double* c, double* d) {
  phase_integrate(cx, C, W, x, c);
  BMPC_SYNC();
  BMPC_TMARK(3);
  // warp 0: the kinematic chains, then (warp-level barrier) the 3 N residual items that need them; warp 1: path terms
  phase_fk(cx, C, W);
  if (in_role(cx, 0, 1)) { BMPC_WSYNC(); }
  phase_kin_residual(cx, C, W, x, c, 0, 1);
  phase_path(cx, C, W, p, x, d, nullptr, 0, 1);
  BMPC_SYNC();
  BMPC_TMARK(4);
}

// The path terms with derivatives are by far the longest serial piece of an evaluation (one lane per stage, several
// times the kinematic chains).  Everything that needs the chains only - kinematic residual rows, curvature matrices,
// Jacobian rows - therefore runs NEXT to them: warp 0 has the path terms, warp 1 the chains, and warps 1 .. nw-1 go on
// with the kinematic phases after a barrier of their own (named barrier 1).
BMPC_DEV void role_barrier(const Ctx cx, int w0, int w1) {
#ifndef BMPC_HOST_EMU
  if (in_role(cx, w0, w1)) asm volatile("bar.sync 1, %0;" ::"r"(32 * (w1 - w0)) : "memory");
#else
  (void)cx; (void)w0; (void)w1;
#endif
}
BMPC_NOINLINE void eval_full(const Ctx cx, const Config& C, const Work& W, const double* p, const double* x) {
  const int nw = ctx_nwarps(cx);
  phase_integrate(cx, C, W, x, W.c);
  BMPC_SYNC();
  BMPC_TMARK(0);
#ifdef BMPC_HOST_EMU
  const int k0 = 0;
#else
  const int k0 = 1;
#endif
  phase_path(cx, C, W, p, x, W.d, nullptr, 1, 0);
  phase_fk(cx, C, W, k0);
  role_barrier(cx, k0, nw);
  phase_kin_residual(cx, C, W, x, W.c, k0, nw);
  phase_kin_hessian(cx, C, W, k0, nw);
  phase_kin_jacobian(cx, C, W, k0, nw);
  BMPC_SYNC();
  BMPC_TMARK(1);
  phase_path_blocks(cx, C, W, p);
  BMPC_TMARK(35);
  phase_grad_f(cx, C, W, p, x, W.gradf);
  BMPC_TMARK(36);
  BMPC_SYNC();
  BMPC_TMARK(2);
}

// trivial (non-kinematic) rows of G = [A_hat | B]: for a column of z = (s_k, u_k) the up-to-three
// x-rows with constant coefficients
BMPC_DEV int triv_col(const Config& C, int col, int* r, double* cf) {
  if (col < 7) { r[0] = col; r[1] = 7 + col; r[2] = 14 + col; cf[0] = C.a_um; cf[1] = C.b_um; cf[2] = C.c_um; return 3; }
  if (col == 7) { r[0] = 33; r[1] = 34; r[2] = 35; cf[0] = C.a_um; cf[1] = C.b_um; cf[2] = C.c_um; return 3; }
  if (col < 15) { r[0] = col - 8; cf[0] = 1.0; return 1; }
  if (col < 22) { r[0] = col - 15; r[1] = 7 + col - 15; cf[0] = C.a_dq; cf[1] = 1.0; return 2; }
  if (col < 29) { const int j = col - 22; r[0] = j; r[1] = 7 + j; r[2] = 14 + j; cf[0] = C.a_ddq; cf[1] = C.b_ddq; cf[2] = 1.0; return 3; }
  if (col < 41) return 0;
  if (col == 41) { r[0] = 33; cf[0] = 1.0; return 1; }
  if (col == 42) { r[0] = 33; r[1] = 34; cf[0] = C.a_dq; cf[1] = 1.0; return 2; }
  if (col == 43) { r[0] = 33; r[1] = 34; r[2] = 35; cf[0] = C.a_ddq; cf[1] = C.b_ddq; cf[2] = 1.0; return 3; }
  if (col < 51) { const int j = col - 44; r[0] = j; r[1] = 7 + j; r[2] = 14 + j; cf[0] = C.a_u; cf[1] = C.b_u; cf[2] = C.c_u; return 3; }
  r[0] = 33; r[1] = 34; r[2] = 35; cf[0] = C.a_u; cf[1] = C.b_u; cf[2] = C.c_u;
  return 3;
}

// (G_k^T v)[col] for v in R^36, G_k = [A_hat_k | B_k]
BMPC_DEV double GT_vec(const Config& C, const double* GK, const double* v, int col) {
  double a = 0.0;
  for (int r = 0; r < NK; r++) a += GK[r * NZ + col] * v[rKIN + r];
  int rr[3]; double cf[3];
  const int nt = triv_col(C, col, rr, cf);
  for (int t = 0; t < nt; t++) a += cf[t] * v[rr[t]];
  return a;
}

// (G_k z)[row] for z = (ds(44), du(8)); ds may be null (stage 0: the initial state is fixed)
BMPC_DEV double G_vec(const Config& C, const double* GK, const double* ds, const double* du, int row) {
  double a = 0.0;
  if (row >= rKIN && row < rKIN + NK) {
    const double* g = GK + (row - rKIN) * NZ;
    if (ds) for (int cidx = 0; cidx < NX; cidx++) a += g[cidx] * ds[cidx];
    for (int cidx = 0; cidx < 8; cidx++) a += g[NX + cidx] * du[cidx];
    return a;
  }
  if (row < 21) {
    const int typ = row / 7, j = row - 7 * typ;
    if (typ == 0) { a = C.a_u * du[j]; if (ds) a += ds[oQ + j] + C.a_dq * ds[oDQ + j] + C.a_ddq * ds[oDDQ + j] + C.a_um * ds[oU + j]; }
    else if (typ == 1) { a = C.b_u * du[j]; if (ds) a += ds[oDQ + j] + C.b_ddq * ds[oDDQ + j] + C.b_um * ds[oU + j]; }
    else { a = C.c_u * du[j]; if (ds) a += ds[oDDQ + j] + C.c_um * ds[oU + j]; }
    return a;
  }
  if (row == 33) { a = C.a_u * du[7]; if (ds) a += ds[oPHI] + C.a_dq * ds[oDPHI] + C.a_ddq * ds[oDDPHI] + C.a_um * ds[oUPHI]; }
  else if (row == 34) { a = C.b_u * du[7]; if (ds) a += ds[oDPHI] + C.b_ddq * ds[oDDPHI] + C.b_um * ds[oUPHI]; }
  else { a = C.c_u * du[7]; if (ds) a += ds[oDDPHI] + C.c_um * ds[oUPHI]; }
  return a;
}

}  // namespace bmpc
