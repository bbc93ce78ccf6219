// boundmpc_b200 — KKT step of the interior-point iteration (kernel family (b2)).
//
// Replaces the sparse symmetric-indefinite factorisation MUMPS performs for Ipopt on the
// ~940 x 940 augmented system of every iteration (BoundMPC.py:120-148 leaves the default
// linear solver).  The NLP of casadi_ocp_formulation.py:88-357 couples
// stage k only with stage k-1, so after eliminating slacks and bound multipliers the Newton
// system is an equality-constrained QP with the recursion
//     w_k = (u_k ; x_{k+1}),   x_{k+1} = A_hat_k w_{k-1} + B_k u_k + c_k
// and is solved by a Riccati sweep with state w_{k-1} (44) and input u_k (8).  The 36 x 52
// matrix G_k = [A_hat_k | B_k] is never formed: 12 kinematic rows are dense (record GK), the
// other 24 rows have at most three constant entries per column (triv_col), which cuts the
// block products from inner dimension 36 to 12.  The staged blocks (M, Y, Z, Q) live in
// shared memory.  Inertia control: the reduced Hessian is positive definite iff every 8 x 8
// block Q_uu has a Cholesky factor; otherwise the caller retries with a larger delta_w
// (Ipopt's inertia correction, Waechter & Biegler 2006, Sec. 3.1).
#pragma once
#include "bmpc_model.cuh"

namespace bmpc {

// shared-memory working set of one CTA (doubles)
struct Smem {
  double M[NX * NX];     // W~_kk + P_{k+1}, then reused for P_k
  double GK[NK * NZ];    // kinematic rows of G_k
  double Y[NE * NZ];     // M_xx G
  double Z[NU * NZ];     // M_ux G
  double Q[NZ * NZ];     // upper triangle of E^T M E + O-terms
  double OU[NU * NX];    // rows u_k of the off-diagonal Hessian block W~_{k,k-1}
  double pv[NX];         // p_{k+1} / p_k
  double mv[NX];         // g^_k + p_{k+1}
  double tv[NE];         // M_xx c + m_x
  double tu[NU];
  double qv[NZ];
  double Lc[NU * NU];    // Cholesky factor of Q_uu
  double odv[6];
  double red[8 * 32];    // block reductions
  double filt[2 * 64];   // filter entries (theta, phi)
  int flag[4];
};

struct KktCoef {         // constants of the velocity / acceleration tracking Hessian
  double w2, w5, w7, w8, w9, w10, w11, w12, w13, idt;
};
BMPC_DEV KktCoef kkt_coef(const Config& C, const double* p) {
  const double* wt = p + C.L.w;
  KktCoef k;
  k.w2 = BMPC_LDG(wt + 2); k.w5 = BMPC_LDG(wt + 5); k.w7 = BMPC_LDG(wt + 7); k.w8 = BMPC_LDG(wt + 8); k.w9 = BMPC_LDG(wt + 9);
  k.w10 = BMPC_LDG(wt + 10); k.w11 = BMPC_LDG(wt + 11); k.w12 = BMPC_LDG(wt + 12); k.w13 = BMPC_LDG(wt + 13);
  k.idt = 1.0 / C.dt;
  return k;
}

// expansion weights of (q_n, dq_n) w.r.t. the variable types  um, q, dq, ddq (previous block) and u
BMPC_DEV void type_coef(const Config& C, double* al, double* be, int* off) {
  al[0] = C.a_um; al[1] = 1.0; al[2] = C.a_dq; al[3] = C.a_ddq; al[4] = C.a_u;
  be[0] = C.b_um; be[1] = 0.0; be[2] = 1.0; be[3] = C.b_ddq; be[4] = C.b_u;
  off[0] = oU; off[1] = oQ; off[2] = oDQ; off[3] = oDDQ; off[4] = oU;
}

// M += W~_kk : Lagrangian Hessian diagonal block of stage k, bound / slack barrier terms, delta_w.
// Also writes g^_k (gradient of the barrier problem without the equality multipliers) into gh.
// M must hold P_{k+1} (or zero) on entry.  All threads; ends with a sync.
BMPC_DEV void assemble_diag(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, int k, double mu, double delta_w,
                            double* M, int mode) {
  // mode 0: barrier terms + write g^ ; 1: barrier terms ; 2: pure Lagrangian Hessian (tests)
  const bool with_gh = mode == 0, barrier = mode != 2;
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const double* recn = rec + R_SIZE;
  const bool has_next = k + 1 < C.N;
  double al[5], be[5]; int off[5];
  type_coef(C, al, be, off);
  double nd = 0.0;
  for (int m = 0; m < 6; m++) nd += rec[R_DPD + m] * rec[R_DPD + m];
  // (a) diagonal: separable cost, tracking terms, bounds, regularisation; gradient part of g^
  PAR_FOR(i, NX) {
    const int gi = NX * k + i;
    double dg = delta_w, gb = W.gradf[gi];
    if (i < 7) dg += 2 * kc.w13; else if (i == 7) dg += 2 * kc.w9; else if (i < 15) dg += 2 * kc.w10;
    else if (i < 22) dg += 2 * kc.w11; else if (i < 29) dg += 2 * kc.w12;
    else if (i >= oVLIN && i < oVLIN + 6) dg += 2 * kc.w2 + 2 * kc.w5 * kc.idt * kc.idt * (has_next ? 2.0 : 1.0);
    else if (i == oDDPHI) dg += 2 * kc.w5 * nd + 2 * kc.w8;
    const double lb = C.lb[i], ub = C.ub[i];
    if (barrier && lb > -1e300) { const double sl = W.x[gi] - lb; dg += W.zL[gi] / sl; gb -= mu / sl; }
    if (barrier && ub < 1e300) { const double su = ub - W.x[gi]; dg += W.zU[gi] / su; gb += mu / su; }
    M[i * NX + i] += dg;
    if (with_gh) W.gh[gi] = gb;
  }
  BMPC_SYNC();
  // (b) y-block (p_pos, p_rot, phi) + dphi: cost Hessian, inequality curvature, J_d^T Sigma J_d
  PAR_FOR(it, 64) {
    const int a = it >> 3, b = it & 7;
    const int ia = a < 6 ? oPPOS + a : (a == 6 ? oPHI : oDPHI);
    const int ib = b < 6 ? oPPOS + b : (b == 6 ? oPHI : oDPHI);
    double v = 0.0;
    if (a < 7 && b < 7) v = rec[R_HY + a * 7 + b];
    if (a == 7 && b == 7) v = 2 * kc.w2 * nd + 2 * kc.w7;
    const double* JD = rec + R_JD;
    for (int r = 0; r < ND; r++) {
      const double sv = W.s[ND * k + r], zv = W.zs[ND * k + r];
      if (barrier) v += (zv / sv) * JD[r * 8 + a] * JD[r * 8 + b];
      if (a == 6 && b == 6) v += zv * rec[R_HD + r];
    }
    M[ia * NX + ib] += v;
    if (with_gh && b == 0) {   // g^ += J_d^T (mu / s + Sigma_s (d + s))
      double gsum = 0.0;
      for (int r = 0; r < ND; r++) {
        const double sv = W.s[ND * k + r], zv = W.zs[ND * k + r];
        gsum += JD[r * 8 + a] * (mu / sv + (zv / sv) * (W.d[ND * k + r] + sv));
      }
      W.gh[NX * k + ia] += gsum;
    }
  }
  //     velocity / acceleration tracking cross terms (v(6) x dphi, v(6) x ddphi)
  PAR_FOR(m, 6) {
    const double dp = rec[R_DPD + m];
    M[(oVLIN + m) * NX + oDPHI] += -2 * kc.w2 * dp;
    M[oDPHI * NX + oVLIN + m] += -2 * kc.w2 * dp;
    M[(oVLIN + m) * NX + oDDPHI] += -2 * kc.w5 * dp * kc.idt;
    M[oDDPHI * NX + oVLIN + m] += -2 * kc.w5 * dp * kc.idt;
  }
  //     kinematic curvature u_k x u_k of this stage
  PAR_FOR(ij, 49) {
    const int i = ij / 7, j = ij - 7 * i;
    const double hqq = rec[R_HQQN + ij], hqd = rec[R_HQDN + ij], hdq = rec[R_HQDN + j * 7 + i];
    M[(oU + i) * NX + oU + j] += al[4] * al[4] * hqq + al[4] * be[4] * (hqd + hdq);
  }
  BMPC_SYNC();
  // (c) kinematic curvature (um,q,dq,ddq)^2 contributed by the next stage's constraints
  if (has_next) {
    PAR_FOR(it, 16 * 49) {
      const int st = it / 49, ij = it - 49 * st, s = st >> 2, t = st & 3, i = ij / 7, j = ij - 7 * i;
      const double hqq = recn[R_HQQN + ij], hqd = recn[R_HQDN + ij], hdq = recn[R_HQDN + j * 7 + i];
      double v = al[s] * al[t] * hqq + al[s] * be[t] * hqd + be[s] * al[t] * hdq;
      if (s == 1 && t == 1) v += recn[R_HQQK + ij];
      if (s == 1 && t == 2) v += recn[R_HQDK + ij];
      if (s == 2 && t == 1) v += recn[R_HQDK + j * 7 + i];
      M[(off[s] + i) * NX + off[t] + j] += v;
    }
  }
  BMPC_SYNC();
}

// OU = rows u_k of W~_{k,k-1} (kinematic coupling of u_k with (um,q,dq,ddq) of the previous block);
// odv[m] = d2/(d ddphi_{k+1} d v_k[m]).  The v_{k+1} x v_k entries are the constant -2 w5 / dt^2.
BMPC_DEV void build_offdiag(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, int k, double* OU, double* odv) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  double al[5], be[5]; int off[5];
  type_coef(C, al, be, off);
  PAR_FOR(i, NU * NX) OU[i] = 0.0;
  BMPC_SYNC();
  PAR_FOR(it, 4 * 49) {
    const int t = it / 49, ij = it - 49 * t, i = ij / 7, j = ij - 7 * i;
    const double hqq = rec[R_HQQN + ij], hqd = rec[R_HQDN + ij], hdq = rec[R_HQDN + j * 7 + i];
    OU[i * NX + off[t] + j] = al[4] * al[t] * hqq + al[4] * be[t] * hqd + be[4] * al[t] * hdq;
  }
  PAR_FOR(m, 6) odv[m] = 2 * kc.w5 * rec[R_DPD + m] * kc.idt;
  BMPC_SYNC();
}

// One backward Riccati step for stage k.  On entry S.M = P_{k+1} (zero for k = N-1), S.pv = p_{k+1}.
// On exit S.M = P_k, S.pv = p_k, gains stored in W.Kk / W.kap.  Returns false if Q_uu is not PD.
BMPC_DEV bool riccati_stage(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, Smem& S, int k, double mu, double delta_w) {
  assemble_diag(cx, C, W, kc, k, mu, delta_w, S.M, 0);
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const double* c = W.c + NE * k;
  PAR_FOR(i, NK * NZ) S.GK[i] = rec[R_GK + i];
  PAR_FOR(i, NX) S.mv[i] = W.gh[NX * k + i] + S.pv[i];
  if (k > 0) build_offdiag(cx, C, W, kc, k, S.OU, S.odv);
  else BMPC_SYNC();
  const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
  const int ncol = k > 0 ? NZ : NU;          // stage 0: the previous block is fixed -> only the u-columns
  const int c0 = k > 0 ? 0 : NX;
  // Y = M_xx G, Z = M_ux G, t = M_xx c + m_x, tu = M_ux c + m_u
  PAR_FOR(it, (NE + NU) * ncol) {
    const int row = it / ncol, col = c0 + it - ncol * row;
    const double* Mr = row < NE ? S.M + (8 + row) * NX + 8 : S.M + (row - NE) * NX + 8;
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < NK; r++) a += Mr[rKIN + r] * S.GK[r * NZ + col];
    int rr[3]; double cf[3];
    const int nt = triv_col(C, col, rr, cf);
    for (int t = 0; t < nt; t++) a += cf[t] * Mr[rr[t]];
    if (row < NE) S.Y[row * NZ + col] = a; else S.Z[(row - NE) * NZ + col] = a;
  }
  PAR_FOR(row, NE + NU) {
    const double* Mr = row < NE ? S.M + (8 + row) * NX + 8 : S.M + (row - NE) * NX + 8;
    double a = row < NE ? S.mv[8 + row] : S.mv[row - NE];
    for (int l = 0; l < NE; l++) a += Mr[l] * c[l];
    if (row < NE) S.tv[row] = a; else S.tu[row - NE] = a;
  }
  BMPC_SYNC();
  // Q = G^T Y + U^T Z + Z^T U + U^T M_uu U + (E^T O [I 0] + transpose),  upper triangle a <= b
  PAR_FOR(it, ncol * ncol) {
    const int ai = it / ncol, bi = it - ncol * ai;
    if (ai > bi) continue;
    const int a = c0 + ai, b = c0 + bi;
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < NK; r++) v += S.GK[r * NZ + a] * S.Y[(rKIN + r) * NZ + b];
    int rr[3]; double cf[3];
    const int nt = triv_col(C, a, rr, cf);
    for (int t = 0; t < nt; t++) v += cf[t] * S.Y[rr[t] * NZ + b];
    if (a >= NX) v += S.Z[(a - NX) * NZ + b];
    if (b >= NX) v += S.Z[(b - NX) * NZ + a];
    if (a >= NX && b >= NX) v += S.M[(a - NX) * NX + (b - NX)];
    if (k > 0) {
      // EO[a][b] = (E^T O)[a][b] for b < 44; contributes at (a,b) and, for a < 44, EO[b][a] at (a,b)
      // O_x: rows v_{k+1} (x-rows 27..32) <- cols v_k (35..40): ovv on the diagonal; row ddphi (35) <- odv
      if (b < NX) {
        if (a >= NX) v += S.OU[(a - NX) * NX + b];
        if (b >= oVLIN && b < oVLIN + 6) {
          const int m = b - oVLIN;
          double g35 = a == oUPHI ? C.c_um : (a == oDDPHI ? 1.0 : (a == NX + oUPHI ? C.c_u : 0.0));
          v += S.GK[(6 + m) * NZ + a] * ovv + g35 * S.odv[m];
        }
      }
      if (a < NX) {
        if (b >= NX) { /* EO[b][a] with b >= 44 handled above by symmetry of the (a>=NX) branch: here a < NX so add it */
          v += S.OU[(b - NX) * NX + a];
        }
        if (a >= oVLIN && a < oVLIN + 6) {
          const int m = a - oVLIN;
          double g35 = b == oUPHI ? C.c_um : (b == oDDPHI ? 1.0 : (b == NX + oUPHI ? C.c_u : 0.0));
          v += S.GK[(6 + m) * NZ + b] * ovv + g35 * S.odv[m];
        }
      }
    }
    S.Q[a * NZ + b] = v;
  }
  // q = G^T t + U^T tu + [O_x^T c ; 0]
  PAR_FOR(ai, ncol) {
    const int a = c0 + ai;
    double v = GT_vec(C, S.GK, S.tv, a);
    if (a >= NX) v += S.tu[a - NX];
    if (k > 0 && a >= oVLIN && a < oVLIN + 6) v += ovv * c[27 + (a - oVLIN)] + S.odv[a - oVLIN] * c[35];
    S.qv[a] = v;
  }
  BMPC_SYNC();
  // Cholesky of Q_uu (8 x 8) by one thread; failure flag broadcast through shared memory
  if (cx.tid == 0) {
    int ok = 1;
    for (int i = 0; i < NU && ok; i++)
      for (int j = 0; j <= i; j++) {
        double a = S.Q[(NX + j) * NZ + NX + i];
        for (int l = 0; l < j; l++) a -= S.Lc[i * NU + l] * S.Lc[j * NU + l];
        if (i == j) {
          if (!(a > 1e-14)) { ok = 0; break; }
          S.Lc[i * NU + i] = sqrt(a);
        } else S.Lc[i * NU + j] = a / S.Lc[j * NU + j];
      }
    S.flag[0] = ok;
  }
  BMPC_SYNC();
  if (!S.flag[0]) return false;
  // gains: K = -Q_uu^{-1} Q_us (8 x 44), kappa = -Q_uu^{-1} q_u
  double* K = W.Kk + (size_t)k * NU * NX;
  double* kap = W.kap + k * NU;
  PAR_FOR(col, (k > 0 ? NX : 0) + 1) {
    const bool isk = col == (k > 0 ? NX : 0);
    double bcol[NU];
    for (int i = 0; i < NU; i++) bcol[i] = isk ? -S.qv[NX + i] : -S.Q[col * NZ + NX + i];
    for (int i = 0; i < NU; i++) { double a = bcol[i]; for (int l = 0; l < i; l++) a -= S.Lc[i * NU + l] * bcol[l]; bcol[i] = a / S.Lc[i * NU + i]; }
    for (int i = NU - 1; i >= 0; i--) { double a = bcol[i]; for (int l = i + 1; l < NU; l++) a -= S.Lc[l * NU + i] * bcol[l]; bcol[i] = a / S.Lc[i * NU + i]; }
    for (int i = 0; i < NU; i++) { if (isk) kap[i] = bcol[i]; else K[i * NX + col] = bcol[i]; }
  }
  BMPC_SYNC();
  if (k == 0) return true;
  // P_k = Q_ss + Q_us^T K (symmetric), p_k = q_s + Q_us^T kappa
  PAR_FOR(it, NX * NX) {
    const int a = it / NX, b = it - NX * a;
    if (a > b) continue;
    double v = S.Q[a * NZ + b];
    for (int i = 0; i < NU; i++) v += S.Q[a * NZ + NX + i] * K[i * NX + b];
    S.M[a * NX + b] = v;
    S.M[b * NX + a] = v;
  }
  PAR_FOR(a, NX) {
    double v = S.qv[a];
    for (int i = 0; i < NU; i++) v += S.Q[a * NZ + NX + i] * kap[i];
    S.pv[a] = v;
  }
  BMPC_SYNC();
  return true;
}

// Full KKT solve: dx (primal step) and ynew (equality multipliers of the full step).
BMPC_DEV bool kkt_solve(const Ctx& cx, const Config& C, const Work& W, const double* p, Smem& S, double mu, double delta_w) {
  const KktCoef kc = kkt_coef(C, p);
  PAR_FOR(i, NX * NX) S.M[i] = 0.0;
  PAR_FOR(i, NX) S.pv[i] = 0.0;
  BMPC_SYNC();
  for (int k = C.N - 1; k >= 0; k--)
    if (!riccati_stage(cx, C, W, kc, S, k, mu, delta_w)) return false;
  // forward sweep
  for (int k = 0; k < C.N; k++) {
    const double* GK = W.rec + (size_t)k * R_SIZE + R_GK;
    const double* K = W.Kk + (size_t)k * NU * NX;
    const double* dsv = k > 0 ? W.dx + NX * (k - 1) : nullptr;
    double* dw = W.dx + NX * k;
    PAR_FOR(i, NU) {
      double a = W.kap[k * NU + i];
      if (dsv) for (int j = 0; j < NX; j++) a += K[i * NX + j] * dsv[j];
      dw[i] = a;
    }
    BMPC_SYNC();
    PAR_FOR(i, NE) dw[8 + i] = G_vec(C, GK, dsv, dw, i) + W.c[NE * k + i];
    BMPC_SYNC();
  }
  // adjoint sweep for the equality multipliers:
  //   y_k = [W~_kk dw_k + O_k dw_{k-1} + O_{k+1}^T dw_{k+1} + g^_k]_x + [A_hat_{k+1}^T y_{k+1}]_x
  const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
  for (int k = C.N - 1; k >= 0; k--) {
    PAR_FOR(i, NX * NX) S.M[i] = 0.0;
    BMPC_SYNC();
    assemble_diag(cx, C, W, kc, k, mu, delta_w, S.M, 1);
    const bool has_next = k + 1 < C.N;
    if (has_next) build_offdiag(cx, C, W, kc, k + 1, S.OU, S.odv);   // O_{k+1}
    const double* dw = W.dx + NX * k;
    const double* dwp = k > 0 ? W.dx + NX * (k - 1) : nullptr;
    const double* dwn = has_next ? W.dx + NX * (k + 1) : nullptr;
    const double* GKn = W.rec + (size_t)(k + 1) * R_SIZE + R_GK;
    const double* reck = W.rec + (size_t)k * R_SIZE;
    PAR_FOR(i, NE) {
      const int r = 8 + i;
      double a = W.gh[NX * k + r];
      for (int j = 0; j < NX; j++) a += S.M[r * NX + j] * dw[j];
      if (dwp) {   // O_k rows v, ddphi
        if (r >= oVLIN && r < oVLIN + 6) a += ovv * dwp[r];
        if (r == oDDPHI) for (int m = 0; m < 6; m++) a += 2 * kc.w5 * reck[R_DPD + m] * kc.idt * dwp[oVLIN + m];
      }
      if (has_next) {
        for (int q = 0; q < 7; q++) a += S.OU[q * NX + r] * dwn[q];                     // O_u,k+1^T du_{k+1}
        if (r >= oVLIN && r < oVLIN + 6) a += ovv * dwn[r] + S.odv[r - oVLIN] * dwn[oDDPHI];   // O_x,k+1^T dx_{k+2}
        a += GT_vec(C, GKn, W.ynew + NE * (k + 1), r);
      }
      W.ynew[NE * k + i] = a;
    }
    BMPC_SYNC();
  }
  return true;
}

}  // namespace bmpc
