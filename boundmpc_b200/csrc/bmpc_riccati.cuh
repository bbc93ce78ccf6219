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

// shared-memory working set of one CTA (doubles).  Leading dimensions 44 / 52 / 12 are = 4 or 12
// (mod 16), which makes the DMMA fragment loads (8 rows x 4 columns of doubles per half-warp)
// bank-conflict free; arrays read as tile operands carry the few pad elements a partial edge
// tile touches (their products only reach tile elements that are never stored).
constexpr int LDM = NX;        // 44
constexpr int LDY = NZ;        // 52
constexpr int LDQ = 12;
struct Smem {
  double M[48 * LDM];          // W~_kk + P_{k+1}, overwritten by P_k
  double GK[NK * NZ + 4];      // kinematic rows of G_k = [A_hat | B]
  double YZ[NX * LDY];         // M[:, x] G  (rows 0..7 = u rows "Z", rows 8..43 = x rows "Y")
  double Qu[56 * LDQ];         // columns u_k of Q: rows 0..43 = Q_su, rows 44..51 = Q_uu
  double Ks[NU * NX + 4];      // feedback gain K_k
  double pv[NX];               // p_{k+1} / p_k
  double mv[NX];               // g^_k + p_{k+1}
  double tv[NX];               // M[:, x] c + m   (rows 0..7 = u part)
  double qv[NZ];
  double cv[NE];               // c_k
  double kapv[NU];
  double odv[6];
  double tcc[NZ * 3];          // constant rows of G per column: coefficients ...
  int tcr[NZ * 3];             // ... and x-row indices (triv_col as a table)
  double red[8 * 32];          // block reductions
  double filt[2 * 64];         // filter entries (theta, phi)
  int flag[4];
};

// table form of triv_col, built once per kernel
BMPC_DEV void build_tables(const Ctx& cx, const Config& C, Smem& S) {
  PAR_FOR(col, NZ) {
    int rr[3]; double cf[3];
    const int nt = triv_col(C, col, rr, cf);
    for (int t = 0; t < 3; t++) { S.tcr[3 * col + t] = t < nt ? rr[t] : 0; S.tcc[3 * col + t] = t < nt ? cf[t] : 0.0; }
  }
  BMPC_SYNC();
}

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

// 8 x 8 Cholesky factor of Q_uu in registers; Lr holds L with the RECIPROCAL diagonal.  Every thread
// that needs the factor computes it itself from shared memory (no broadcast, no extra barrier; the
// result is identical in all threads, so the inertia verdict is CTA-uniform).
BMPC_DEV bool chol8(const double* Qu, double (&A)[NU][NU]) {
#pragma unroll
  for (int i = 0; i < NU; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) A[i][j] = Qu[(NX + j) * LDQ + i];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NU; j++) {
    double d = A[j][j];
#pragma unroll
    for (int l = 0; l < j; l++) d -= A[j][l] * A[j][l];
    if (!(d > 1e-14)) ok = false;
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
  return ok;
}

// One backward Riccati step for stage k.  On entry S.M = P_{k+1} (zero for k = N-1), S.pv = p_{k+1}.
// On exit S.M = P_k, S.pv = p_k, gains stored in W.Kk / W.kap.  Returns false if Q_uu is not PD.
// The dense block products run as 8 x 8 DMMA tiles (mma_tile); the constant rows of G are applied
// from the tcr / tcc tables in the tile epilogues.
BMPC_DEV bool riccati_stage(const Ctx& cx, const Config& C, const Work& W, const KktCoef& kc, Smem& S, int k, double delta_w) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const double* Wd = W.Wd + (size_t)k * NX * NX;
  const double* OUa = W.OUa + (size_t)k * NU * NX;
  const bool first = k == 0;      // stage 0: the previous block is fixed -> only the u-columns
  PAR_FOR(i, NX * NX) S.M[i] += Wd[i] + ((i % (NX + 1)) == 0 ? delta_w : 0.0);
  PAR_FOR(i, NK * NZ) S.GK[i] = rec[R_GK + i];
  PAR_FOR(i, NX) S.mv[i] = W.gh[NX * k + i] + S.pv[i];
  PAR_FOR(i, NE) S.cv[i] = W.c[NE * k + i];
  PAR_FOR(m, 6) S.odv[m] = 2 * kc.w5 * rec[R_DPD + m] * kc.idt;
  BMPC_SYNC();
  const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
  // ---- YZ = M[:, x] G  (44 x 52): 6 x 7 tiles, inner dimension = the 12 kinematic rows
  {
    const int tj0 = first ? 5 : 0, ntj = 7 - tj0;
    TILE_FOR(t, 6 * ntj) {
      const int ti = t / ntj, tj = tj0 + t - ntj * ti;
      mma_tile(cx, 3,
               [&](int r, int kk) { return S.M[(8 * ti + r) * LDM + 8 + rKIN + kk]; },
               [&](int kk, int c) { return S.GK[kk * NZ + 8 * tj + c]; },
               [&](int r, int c, double v) {
                 const int i = 8 * ti + r, col = 8 * tj + c;
                 if (i < NX && col < NZ) {
                   const double* Mr = S.M + i * LDM + 8;
#pragma unroll
                   for (int q = 0; q < 3; q++) v += S.tcc[3 * col + q] * Mr[S.tcr[3 * col + q]];
                   S.YZ[i * LDY + col] = v;
                 }
               });
    }
    // t = M[:, x] c + m
    PAR_FOR(i, NX) {
      const double* Mr = S.M + i * LDM + 8;
      double a0 = S.mv[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int l = 0; l < NE; l += 4) { a0 += Mr[l] * S.cv[l]; a1 += Mr[l + 1] * S.cv[l + 1]; a2 += Mr[l + 2] * S.cv[l + 2]; a3 += Mr[l + 3] * S.cv[l + 3]; }
      S.tv[i] = (a0 + a1) + (a2 + a3);
    }
  }
  BMPC_SYNC();
  // ---- columns u_k of Q = E^T M E + O-terms  (52 x 8): 7 tiles
  {
    const int ti0 = first ? 5 : 0;
    TILE_FOR(t, 7 - ti0) {
      const int ti = ti0 + t;
      mma_tile(cx, 3,
               [&](int r, int kk) { return S.GK[kk * NZ + 8 * ti + r]; },
               [&](int kk, int c) { return S.YZ[(8 + rKIN + kk) * LDY + NX + c]; },
               [&](int r, int j, double v) {
                 const int a = 8 * ti + r;
                 if (a >= NZ || (first && a < NX)) return;
#pragma unroll
                 for (int q = 0; q < 3; q++) v += S.tcc[3 * a + q] * S.YZ[(8 + S.tcr[3 * a + q]) * LDY + NX + j];
                 if (a < NX) {
                   v += S.YZ[j * LDY + a] + OUa[j * NX + a];
                   if (a >= oVLIN && a < oVLIN + 6) v += S.GK[(6 + a - oVLIN) * NZ + NX + j] * ovv + (j == 7 ? C.c_u * S.odv[a - oVLIN] : 0.0);
                 } else {
                   const int i = a - NX;
                   v += S.YZ[i * LDY + NX + j] + S.YZ[j * LDY + NX + i] + S.M[i * LDM + j];
                 }
                 S.Qu[a * LDQ + j] = v;
               });
    }
    // q = G^T t + U^T t_u + [O_x^T c ; 0]
    PAR_FOR(ai, first ? NU : NZ) {
      const int a = first ? NX + ai : ai;
      double v = 0.0;
#pragma unroll
      for (int r = 0; r < NK; r++) v += S.GK[r * NZ + a] * S.tv[8 + rKIN + r];
#pragma unroll
      for (int q = 0; q < 3; q++) v += S.tcc[3 * a + q] * S.tv[8 + S.tcr[3 * a + q]];
      if (a >= NX) v += S.tv[a - NX];
      if (!first && a >= oVLIN && a < oVLIN + 6) v += ovv * S.cv[27 + (a - oVLIN)] + S.odv[a - oVLIN] * S.cv[35];
      S.qv[a] = v;
    }
  }
  BMPC_SYNC();
  // ---- gains: K = -Q_uu^{-1} Q_us (8 x 44), kappa = -Q_uu^{-1} q_u; the factor is recomputed per thread
  double* K = W.Kk + (size_t)k * NU * NX;
  double* kap = W.kap + k * NU;
  bool ok;
  {
    double L[NU][NU];
    ok = chol8(S.Qu, L);
    if (ok) {
      PAR_FOR(col, (first ? 0 : NX) + 1) {
        const bool isk = col == (first ? 0 : NX);
        double bcol[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) bcol[i] = isk ? -S.qv[NX + i] : -S.Qu[col * LDQ + i];
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
        for (int i = 0; i < NU; i++) {
          if (isk) { kap[i] = bcol[i]; S.kapv[i] = bcol[i]; }
          else { K[i * NX + col] = bcol[i]; S.Ks[i * NX + col] = bcol[i]; }
        }
      }
    }
  }
  if (!ok) return false;
  BMPC_SYNC();
  if (first) return true;
  // ---- P_k = Q_ss + Q_su K  (symmetric; upper tiles, mirrored), p_k = q_s + Q_su kappa
  TILE_FOR(t, 21) {
    int ti = 0, rem = t;
    while (rem >= 6 - ti) { rem -= 6 - ti; ti++; }
    const int tj = ti + rem;
    mma_tile(cx, 5,
             [&](int r, int kk) { return kk < NK ? S.GK[kk * NZ + 8 * ti + r] : S.Qu[(8 * ti + r) * LDQ + kk - NK]; },
             [&](int kk, int c) { return kk < NK ? S.YZ[(8 + rKIN + kk) * LDY + 8 * tj + c] : S.Ks[(kk - NK) * NX + 8 * tj + c]; },
             [&](int r, int c, double v) {
               const int a = 8 * ti + r, b = 8 * tj + c;
               if (a >= NX || b >= NX || a > b) return;
#pragma unroll
               for (int q = 0; q < 3; q++) v += S.tcc[3 * a + q] * S.YZ[(8 + S.tcr[3 * a + q]) * LDY + b];
               // E^T O [I 0] + transpose: rows v_{k+1} x cols v_k carry ovv, row ddphi_{k+1} carries odv
               if (b >= oVLIN && b < oVLIN + 6) {
                 const double g35 = a == oUPHI ? C.c_um : (a == oDDPHI ? 1.0 : 0.0);
                 v += S.GK[(6 + b - oVLIN) * NZ + a] * ovv + g35 * S.odv[b - oVLIN];
               }
               if (a >= oVLIN && a < oVLIN + 6) {
                 const double g35 = b == oUPHI ? C.c_um : (b == oDDPHI ? 1.0 : 0.0);
                 v += S.GK[(6 + a - oVLIN) * NZ + b] * ovv + g35 * S.odv[a - oVLIN];
               }
               S.M[a * LDM + b] = v;
               S.M[b * LDM + a] = v;
             });
  }
  PAR_FOR(a, NX) {
    double v = S.qv[a];
#pragma unroll
    for (int i = 0; i < NU; i++) v += S.Qu[a * LDQ + i] * S.kapv[i];
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
