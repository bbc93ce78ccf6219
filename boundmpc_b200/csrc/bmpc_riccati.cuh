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

// joint-variable classification of a stage index: type 0 = u (previous input when seen from the next
// stage), 1 = q, 2 = dq, 3 = ddq; -1 otherwise
BMPC_DEV int jtype(int a) { return a < 7 ? 0 : (a < 8 ? -1 : (a < 15 ? 1 : (a < 22 ? 2 : (a < 29 ? 3 : -1)))); }
BMPC_DEV int jidx(int a) { return a < 7 ? a : (a < 15 ? a - 8 : (a < 22 ? a - 15 : a - 22)); }
BMPC_DEV int yidx(int a) { return (a >= oPPOS && a < oPPOS + 6) ? a - oPPOS : (a == oPHI ? 6 : (a == oDPHI ? 7 : -1)); }

// One entry (a, b) of the diagonal block W~_kk (gather form: every entry is written by exactly one
// thread).  barrier = true adds the bound terms and uses the y-block with the slack terms.
BMPC_DEV double wd_entry(const Config& C, const Work& W, const KktCoef& kc, const double* al, const double* be, int k, int a, int b,
                         bool barrier) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const bool has_next = k + 1 < C.N;
  double v = 0.0;
  if (a == b) {
    if (a < 7) v = 2 * kc.w13; else if (a == 7) v = 2 * kc.w9; else if (a < 15) v = 2 * kc.w10;
    else if (a < 22) v = 2 * kc.w11; else if (a < 29) v = 2 * kc.w12;
    else if (a >= oVLIN && a < oVLIN + 6) v = 2 * kc.w2 + 2 * kc.w5 * kc.idt * kc.idt * (has_next ? 2.0 : 1.0);
    else if (a == oDDPHI) {
      double nd = 0.0;
      for (int m = 0; m < 6; m++) nd += rec[R_DPD + m] * rec[R_DPD + m];
      v = 2 * kc.w5 * nd + 2 * kc.w8;
    }
    if (barrier) {
      const int gi = NX * k + a;
      const double lb = C.lb[a], ub = C.ub[a];
      if (lb > -1e300) v += W.zL[gi] / (W.x[gi] - lb);
      if (ub < 1e300) v += W.zU[gi] / (ub - W.x[gi]);
    }
  }
  const int ya = yidx(a), yb = yidx(b);
  if (ya >= 0 && yb >= 0) {
    if (barrier) v += rec[R_HYB + ya * 8 + yb];
    else {
      if (ya < 7 && yb < 7) v += rec[R_HY + ya * 7 + yb];
      if (ya == 6 && yb == 6) for (int r = 0; r < ND; r++) v += W.zs[ND * k + r] * rec[R_HD + r];
      if (ya == 7 && yb == 7) {
        double nd = 0.0;
        for (int m = 0; m < 6; m++) nd += rec[R_DPD + m] * rec[R_DPD + m];
        v += 2 * kc.w2 * nd + 2 * kc.w7;
      }
    }
  }
  // velocity / acceleration tracking cross terms
  {
    const int va = (a >= oVLIN && a < oVLIN + 6) ? a - oVLIN : -1, vb = (b >= oVLIN && b < oVLIN + 6) ? b - oVLIN : -1;
    if (va >= 0 && b == oDPHI) v += -2 * kc.w2 * rec[R_DPD + va];
    if (vb >= 0 && a == oDPHI) v += -2 * kc.w2 * rec[R_DPD + vb];
    if (va >= 0 && b == oDDPHI) v += -2 * kc.w5 * rec[R_DPD + va] * kc.idt;
    if (vb >= 0 && a == oDDPHI) v += -2 * kc.w5 * rec[R_DPD + vb] * kc.idt;
  }
  // kinematic curvature
  const int sa = jtype(a), sb = jtype(b);
  if (sa >= 0 && sb >= 0) {
    const int i = jidx(a), j = jidx(b);
    if (sa == 0 && sb == 0) {
      const double hqq = rec[R_HQQN + i * 7 + j], hqd = rec[R_HQDN + i * 7 + j], hdq = rec[R_HQDN + j * 7 + i];
      v += al[4] * al[4] * hqq + al[4] * be[4] * (hqd + hdq);
    }
    if (has_next) {
      const double* rn = rec + R_SIZE;
      const double hqq = rn[R_HQQN + i * 7 + j], hqd = rn[R_HQDN + i * 7 + j], hdq = rn[R_HQDN + j * 7 + i];
      v += al[sa] * al[sb] * hqq + al[sa] * be[sb] * hqd + be[sa] * al[sb] * hdq;
      if (sa == 1 && sb == 1) v += rn[R_HQQK + i * 7 + j];
      if (sa == 1 && sb == 2) v += rn[R_HQDK + i * 7 + j];
      if (sa == 2 && sb == 1) v += rn[R_HQDK + j * 7 + i];
    }
  }
  return v;
}

// entry (i, col) of the rows u_k of W~_{k,k-1} (kinematic coupling of u_k with (um, q, dq, ddq) of the previous block)
BMPC_DEV double ou_entry(const Work& W, const double* al, const double* be, int k, int i, int col) {
  const int t = jtype(col);
  if (i >= 7 || t < 0) return 0.0;
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const int j = jidx(col);
  return al[4] * al[t] * rec[R_HQQN + i * 7 + j] + al[4] * be[t] * rec[R_HQDN + i * 7 + j] + be[4] * al[t] * rec[R_HQDN + j * 7 + i];
}

// All stages at once: W.Wd, W.OUa and g^ (gradient of the barrier problem without the equality multipliers)
BMPC_DEV void kkt_build(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, double mu, bool barrier) {
  double al[5], be[5]; int off[5];
  type_coef(C, al, be, off);
  const int N = C.N;
  PAR_FOR(it, N * NX * NX) {
    const int k = it / (NX * NX), ab = it - k * NX * NX, a = ab / NX, b = ab - NX * a;
    W.Wd[it] = wd_entry(C, W, kc, al, be, k, a, b, barrier);
  }
  PAR_FOR(it, N * NU * NX) {
    const int k = it / (NU * NX), ic = it - k * NU * NX, i = ic / NX, col = ic - NX * i;
    W.OUa[it] = k > 0 ? ou_entry(W, al, be, k, i, col) : 0.0;
  }
  if (barrier) {
    PAR_FOR(gi, C.n) {
      const int k = gi / NX, a = gi - NX * k;
      double gb = W.gradf[gi];
      const double lb = C.lb[a], ub = C.ub[a];
      if (lb > -1e300) gb -= mu / (W.x[gi] - lb);
      if (ub < 1e300) gb += mu / (ub - W.x[gi]);
      const int ya = yidx(a);
      if (ya >= 0) { const double* rec = W.rec + (size_t)k * R_SIZE; gb += mu * rec[R_GJ1 + ya] + rec[R_GJ2 + ya]; }
      W.gh[gi] = gb;
    }
  }
  BMPC_SYNC();
}

// One backward Riccati step for stage k.  On entry S.M = P_{k+1} (zero for k = N-1), S.pv = p_{k+1}.
// On exit S.M = P_k, S.pv = p_k, gains stored in W.Kk / W.kap.  Returns false if Q_uu is not PD.
BMPC_DEV bool riccati_stage(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, Smem& S, int k, double delta_w) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const double* c = W.c + NE * k;
  const double* Wd = W.Wd + (size_t)k * NX * NX;
  PAR_FOR(i, NX * NX) S.M[i] += Wd[i] + ((i % (NX + 1)) == 0 ? delta_w : 0.0);
  PAR_FOR(i, NK * NZ) S.GK[i] = rec[R_GK + i];
  PAR_FOR(i, NX) S.mv[i] = W.gh[NX * k + i] + S.pv[i];
  if (k > 0) {
    const double* OUa = W.OUa + (size_t)k * NU * NX;
    PAR_FOR(i, NU * NX) S.OU[i] = OUa[i];
    PAR_FOR(m, 6) S.odv[m] = 2 * kc.w5 * rec[R_DPD + m] * kc.idt;
  }
  BMPC_SYNC();
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
      // EO = E^T O (52 x 44): contributes EO[a][b] for b < 44 and EO[b][a] for a < 44.
      // O_x: rows v_{k+1} (x-rows 27..32) x cols v_k (35..40): ovv on the diagonal; row ddphi (x-row 35): odv
      if (b < NX && b >= oVLIN && b < oVLIN + 6) {
        const int m = b - oVLIN;
        const double g35 = a == oUPHI ? C.c_um : (a == oDDPHI ? 1.0 : (a == NX + oUPHI ? C.c_u : 0.0));
        v += S.GK[(6 + m) * NZ + a] * ovv + g35 * S.odv[m];
      }
      if (a < NX) {
        if (b >= NX) v += S.OU[(b - NX) * NX + a];
        if (a >= oVLIN && a < oVLIN + 6) {
          const int m = a - oVLIN;
          const double g35 = b == oUPHI ? C.c_um : (b == oDDPHI ? 1.0 : (b == NX + oUPHI ? C.c_u : 0.0));
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
  // Cholesky of Q_uu (8 x 8) in registers by one thread; Lc holds L with the RECIPROCAL diagonal
  if (cx.tid == 0) {
    double A[NU][NU];
#pragma unroll
    for (int i = 0; i < NU; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A[i][j] = S.Q[(NX + j) * NZ + NX + i];
    int ok = 1;
#pragma unroll
    for (int j = 0; j < NU; j++) {
      double d = A[j][j];
#pragma unroll
      for (int l = 0; l < j; l++) d -= A[j][l] * A[j][l];
      if (!(d > 1e-14)) ok = 0;
      const double inv = 1.0 / sqrt(d > 1e-14 ? d : 1.0);
      A[j][j] = inv;
#pragma unroll
      for (int i = j + 1; i < NU; i++) {
        double e = A[i][j];
#pragma unroll
        for (int l = 0; l < j; l++) e -= A[i][l] * A[j][l];
        A[i][j] = e * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < NU; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) S.Lc[i * NU + j] = A[i][j];
    S.flag[0] = ok;
  }
  BMPC_SYNC();
  if (!S.flag[0]) return false;
  // gains: K = -Q_uu^{-1} Q_us (8 x 44), kappa = -Q_uu^{-1} q_u
  double* K = W.Kk + (size_t)k * NU * NX;
  double* kap = W.kap + k * NU;
  PAR_FOR(col, (k > 0 ? NX : 0) + 1) {
    const bool isk = col == (k > 0 ? NX : 0);
    double L[NU][NU], bcol[NU];
#pragma unroll
    for (int i = 0; i < NU; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) L[i][j] = S.Lc[i * NU + j];
#pragma unroll
    for (int i = 0; i < NU; i++) bcol[i] = isk ? -S.qv[NX + i] : -S.Q[col * NZ + NX + i];
#pragma unroll
    for (int i = 0; i < NU; i++) {
      double a = bcol[i];
#pragma unroll
      for (int l = 0; l < i; l++) a -= L[i][l] * bcol[l];
      bcol[i] = a * L[i][i];
    }
#pragma unroll
    for (int i = NU - 1; i >= 0; i--) {
      double a = bcol[i];
#pragma unroll
      for (int l = i + 1; l < NU; l++) a -= L[l][i] * bcol[l];
      bcol[i] = a * L[i][i];
    }
#pragma unroll
    for (int i = 0; i < NU; i++) { if (isk) kap[i] = bcol[i]; else K[i * NX + col] = bcol[i]; }
  }
  BMPC_SYNC();
  if (k == 0) return true;
  // P_k = Q_ss + Q_us^T K (symmetric), p_k = q_s + Q_us^T kappa
  PAR_FOR(it, NX * NX) {
    const int a = it / NX, b = it - NX * a;
    if (a > b) continue;
    double v = S.Q[a * NZ + b];
#pragma unroll
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

// Backward + forward + adjoint sweeps on the blocks prepared by kkt_build:
// dx (primal step) and ynew (equality multipliers of the full step).
BMPC_DEV bool kkt_solve(const Ctx& cx, const Config& C, const Work& W, const double* p, Smem& S, double delta_w) {
  const KktCoef kc = kkt_coef(C, p);
  PAR_FOR(i, NX * NX) S.M[i] = 0.0;
  PAR_FOR(i, NX) S.pv[i] = 0.0;
  BMPC_SYNC();
  for (int k = C.N - 1; k >= 0; k--)
    if (!riccati_stage(cx, C, W, kc, S, k, delta_w)) return false;
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
    const bool has_next = k + 1 < C.N;
    const double* Wd = W.Wd + (size_t)k * NX * NX;
    const double* OUn = W.OUa + (size_t)(k + 1) * NU * NX;
    const double* dw = W.dx + NX * k;
    const double* dwp = k > 0 ? W.dx + NX * (k - 1) : nullptr;
    const double* dwn = has_next ? W.dx + NX * (k + 1) : nullptr;
    const double* GKn = W.rec + (size_t)(k + 1) * R_SIZE + R_GK;
    const double* reck = W.rec + (size_t)k * R_SIZE;
    PAR_FOR(i, NE) {
      const int r = 8 + i;
      double a = W.gh[NX * k + r] + delta_w * dw[r];
      for (int j = 0; j < NX; j++) a += Wd[r * NX + j] * dw[j];
      if (dwp) {   // O_k rows v, ddphi
        if (r >= oVLIN && r < oVLIN + 6) a += ovv * dwp[r];
        if (r == oDDPHI) for (int m = 0; m < 6; m++) a += 2 * kc.w5 * reck[R_DPD + m] * kc.idt * dwp[oVLIN + m];
      }
      if (has_next) {
        for (int q = 0; q < 7; q++) a += OUn[q * NX + r] * dwn[q];                                   // O_u,k+1^T du_{k+1}
        if (r >= oVLIN && r < oVLIN + 6) a += ovv * dwn[r] + 2 * kc.w5 * reck[R_SIZE + R_DPD + (r - oVLIN)] * kc.idt * dwn[oDDPHI];
        a += GT_vec(C, GKn, W.ynew + NE * (k + 1), r);
      }
      W.ynew[NE * k + i] = a;
    }
    BMPC_SYNC();
  }
  return true;
}

}  // namespace bmpc
