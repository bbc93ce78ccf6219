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
constexpr int VEC_NMAX = 10, VEC_CAP = VEC_NMAX * (3 * NX + NE + 2 * ND);
constexpr int EV_STEP0 = NX * LDM + (NK * NZ + 4) + 2 * (4 * 49 + 4) + (64 + 8 + NX) + NX + 64 + 256;   // offset of YZ[256] in ev
constexpr int EV_CAP = NX * LDM + (NK * NZ + 4) + 2 * (4 * 49 + 4) + (64 + 8 + NX) + NX + 64 + NX * LDY + 56 * LDQ + (NU * NX + 4);
struct Smem {
  union {
  double ev[EV_CAP];           // evaluation scratch (kinematic chains, path records) while no Riccati sweep is running
  struct {
  double M[NX * LDM];          // W~_kk + P_{k+1}, overwritten by P_k
  double GK[NK * NZ + 4];      // kinematic rows of G_k = [A_hat | B] (loaded asynchronously during phases 1-2a)
  double Hn[2][4 * 49 + 4];    // HQQN, HQDN, HQQK, HQDK of stage k in buffer k & 1
  double Hc[64 + 8 + NX];      // current stage only: HYB [64], DPD [6], sig [44]
  double ghs[NX];              // g^_k
  double Muu[64];              // copy of M[0..7][0..7] (needed by Q_uu after S.M has been reused for Q_ss)
  double YZ[NX * LDY];         // M[:, x] G  (rows 0..7 = u rows "Z", rows 8..43 = x rows "Y")
  double Qu[56 * LDQ];         // columns u_k of Q: rows 0..43 = Q_su, rows 44..51 = Q_uu
  double Ks[NU * NX + 4];      // feedback gain K_k
  };
  };
  double vec[VEC_CAP];         // iterate of the interior-point method (x, y, s, z_s, z_L, z_U) for N <= 10
  Config cfg;                  // copies of the kernel parameter and of the workspace pointers: the large phase bodies are
  Work work;                   // real functions and take them by reference; these copies keep those reads on chip
  double pv[NX];               // p_{k+1} / p_k
  double mv[NX];               // g^_k + p_{k+1}
  double tv[NX];               // M[:, x] c + m   (rows 0..7 = u part)
  double qv[NZ];
  double cv[NE];               // c_k
  double kapv[NU];
  double odv[6];
  double tcc[NZ * 3];          // constant rows of G per column: coefficients ...
  short tcr[NZ * 3 + 4];       // ... and x-row indices (triv_col as a table)
  double alc[5], bec[5];       // d q_n / d(um, q, dq, ddq, u) and d dq_n / d(...) of the integrator (type_coef)
  double red[8 * RED_WARPS];   // block reductions (8 values x up to RED_WARPS warps)
  double filt[2 * 64];         // filter entries (theta, phi)
  int flag[4];
#ifdef BMPC_TIMING
  long long tm[64];
#endif
};

// The blocks of the Riccati sweep are idle during the evaluation phases: the forward-kinematics scratch and the
// path part of the stage records (both written lane by lane, i.e. as scattered 8-byte stores if they lived in
// global memory) are placed there when they fit (N = 10: both; N = 20: the kinematics scratch only).
// (vec_ext: dynamic shared memory behind the Smem struct for the iterate of horizons above VEC_NMAX, or null)
BMPC_DEV void work_attach_smem(Work& W, Smem& S, int N, double* vec_ext = nullptr, bool vec_in_ws = false) {
  int used = 0;
  if (2 * N * F_SIZE <= EV_CAP) { W.fk = S.ev; used = 2 * N * F_SIZE; }
  if (used + N * R_PATH <= EV_CAP) { W.prec = S.ev + used - R_HY; W.prec_stride = R_PATH; }
  if (N <= VEC_NMAX) {
    // the iterate lives in shared memory for the whole solve ...
    const int n = NX * N, ne = NE * N, nd = ND * N;
    double* q = S.vec;
    if (!vec_in_ws) { W.x = q; q += n; W.zL = q; q += n; W.zU = q; q += n; W.y = q; q += ne; W.s = q; q += nd; W.zs = q; q += nd; }
    // ... and the step / trial vectors sit behind the sweep scratch (YZ[0..256)) of the Riccati blocks: they are
    // written by the forward / adjoint sweeps after the backward pass and are dead before the next one starts
    q = S.ev + EV_STEP0;
    W.dx = q; q += n; W.dzL = q; q += n; W.dzU = q; q += n; W.xt = q; q += n; W.ynew = q; q += ne; W.ct = q; q += ne;
    W.ds = q; q += nd; W.dzs = q; q += nd; W.st = q; q += nd; W.dtr = q; q += nd;
  } else if (vec_ext) {
    // longer horizons (N = 20: 30 KB): the iterate alone, in the dynamic shared memory behind the struct (two CTAs per SM
    // instead of three); the step / trial vectors stay in the workspace
    const int n = NX * N, ne = NE * N, nd = ND * N;
    double* q = vec_ext;
    W.x = q; q += n; W.zL = q; q += n; W.zU = q; q += n; W.y = q; q += ne; W.s = q; q += nd; W.zs = q; q += nd;
  }
}

// table form of triv_col, built once per kernel
BMPC_DEV void build_tables(const Ctx cx, const Config& C, Smem& S) {
  PAR_FOR(col, NZ) {
    int rr[3]; double cf[3];
    const int nt = triv_col(C, col, rr, cf);
    for (int t = 0; t < 3; t++) { S.tcr[3 * col + t] = t < nt ? rr[t] : 0; S.tcc[3 * col + t] = t < nt ? cf[t] : 0.0; }
  }
  if (cx.tid == 0) {
    S.alc[0] = C.a_um; S.alc[1] = 1.0; S.alc[2] = C.a_dq; S.alc[3] = C.a_ddq; S.alc[4] = C.a_u;
    S.bec[0] = C.b_um; S.bec[1] = 0.0; S.bec[2] = 1.0; S.bec[3] = C.b_ddq; S.bec[4] = C.b_u;
  }
  BMPC_SYNC();
}

// (G_k^T v)[col] with the constant rows taken from the tables
BMPC_DEV double GT_tab(const Smem& S, const double* GK, const double* v, int col) {
  double a = 0.0;
#pragma unroll
  for (int r = 0; r < NK; r++) a += GK[r * NZ + col] * v[rKIN + r];
#pragma unroll
  for (int t = 0; t < 3; t++) a += S.tcc[3 * col + t] * v[S.tcr[3 * col + t]];
  return a;
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

// first index of the joint-variable types um, q, dq, ddq inside a stage block, and the stage
// index of the y-block entries (p_pos, p_rot, phi, dphi)
BMPC_DEV int joff(int t) { return t == 0 ? oU : (t == 1 ? oQ : (t == 2 ? oDQ : oDDQ)); }
BMPC_DEV int yrow(int y) { return y < 6 ? oPPOS + y : (y == 6 ? oPHI : oDPHI); }

// constant part of the diagonal of W~_kk (quadratic cost terms, bound_mpc_functions.py:205-246)
BMPC_DEV double wdiag_const(const KktCoef& kc, const double* dpd, int a, bool has_next) {
  if (a < 7) return 2 * kc.w13;
  if (a == 7) return 2 * kc.w9;
  if (a < 15) return 2 * kc.w10;
  if (a < 22) return 2 * kc.w11;
  if (a < 29) return 2 * kc.w12;
  if (a >= oVLIN && a < oVLIN + 6) return 2 * kc.w2 + 2 * kc.w5 * kc.idt * kc.idt * (has_next ? 2.0 : 1.0);
  if (a == oDDPHI) {
    double nd = 0.0;
    for (int m = 0; m < 6; m++) nd += dpd[m] * dpd[m];
    return 2 * kc.w5 * nd + 2 * kc.w8;
  }
  return 0.0;
}

// One entry (a, b) of the diagonal block of the Lagrangian Hessian (dense export for the parity
// tests only; the solver never forms these blocks, see add_W / adjoint_rhs).
BMPC_DEV double wd_entry(const Config& C, const Work& W, const KktCoef& kc, const double* al, const double* be, int k, int a, int b) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  const bool has_next = k + 1 < C.N;
  double v = a == b ? wdiag_const(kc, rec + R_DPD, a, has_next) : 0.0;
  const int ya = yidx(a), yb = yidx(b);
  if (ya >= 0 && yb >= 0) {
    if (ya < 7 && yb < 7) v += rec[R_HY + ya * 7 + yb];
    if (ya == 6 && yb == 6) for (int r = 0; r < ND; r++) v += W.zs[ND * k + r] * rec[R_HD + r];
    if (ya == 7 && yb == 7) {
      double nd = 0.0;
      for (int m = 0; m < 6; m++) nd += rec[R_DPD + m] * rec[R_DPD + m];
      v += 2 * kc.w2 * nd + 2 * kc.w7;
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

// entry (i, col) of the rows u_k of W~_{k,k-1} (kinematic coupling of u_k with (um, q, dq, ddq) of the
// previous block); H = HQQN (49) followed by HQDN (49) of stage k; al / be may point to the shared tables
BMPC_DEV double ou_entry(const double* H, const double* al, const double* be, int i, int col) {
  const int t = jtype(col);
  if (i >= 7 || t < 0) return 0.0;
  const int j = jidx(col);
  return al[4] * al[t] * H[i * 7 + j] + al[4] * be[t] * H[49 + i * 7 + j] + be[4] * al[t] * H[49 + j * 7 + i];
}

// Once per interior-point iteration: bound part of the barrier Hessian (W.sig) and the gradient g^ of
// the barrier problem without the equality multipliers (W.gh).
BMPC_DEV void kkt_prepare(const Ctx cx, const Config& C, const Work& W, double mu) {
  PAR_FOR(gi, C.n) {
    const int k = gi / NX, a = gi - NX * k;
    double gb = W.gradf[gi], sg = 0.0;
    const double lb = C.lb[a], ub = C.ub[a];
    // (one reciprocal per bound: the IEEE divide is ~30 instructions of dependent latency, and these loops are made of them)
    if (lb > -1e300) { const double isl = 1.0 / (W.x[gi] - lb); gb -= mu * isl; sg += W.zL[gi] * isl; }
    if (ub < 1e300) { const double isu = 1.0 / (ub - W.x[gi]); gb += mu * isu; sg += W.zU[gi] * isu; }
    const int ya = yidx(a);
    if (ya >= 0) { const double* rec = W.rec + (size_t)k * R_SIZE; gb += mu * rec[R_GJ1 + ya] + rec[R_GJ2 + ya]; }
    W.gh[gi] = gb;
    W.sig[gi] = sg;
  }
  BMPC_SYNC();
}

// Stage data of the backward sweep is staged in shared memory one stage ahead (stage_prefetch runs
// in the gain phase of stage k + 1, on the warps that have no column to solve).
// kinematic rows of stage k into S.GK (single buffer: issued after the last reader of stage k + 1's rows, the
// P update, has passed its barrier; first needed by phase 2b)
BMPC_DEV void gk_load(const Ctx cx, const Work& W, Smem& S, int k) {
  const double* src = W.rec + (size_t)k * R_SIZE + R_GK;
#pragma unroll 1
  PAR_FOR(i, NK * NZ) cp_async8(S.GK + i, src + i);
}

BMPC_DEV void stage_prefetch(const Ctx cx, const Config& C, const Work& W, Smem& S, int k, int w0, int w1) {
  const double* rec = W.rec + (size_t)k * R_SIZE;
  double* Hk = S.Hn[k & 1];
  // one rolled loop over the 390 doubles of the stage (code size: this runs once per stage inside the sweep)
  constexpr int n0 = 4 * 49, n1 = n0 + 64, n2 = n1 + 6, n3 = n2 + NX, n4 = n3 + NX, n5 = n4 + NE;
#pragma unroll 1
  ROLE_FOR(i, n5, w0, w1) {
    const double* src;
    double* dst;
    if (i < n0) { src = rec + R_HQQN + i; dst = Hk + i; }
    else if (i < n1) { src = rec + R_HYB + (i - n0); dst = S.Hc + (i - n0); }
    else if (i < n2) { src = rec + R_DPD + (i - n1); dst = S.Hc + 64 + (i - n1); }
    else if (i < n3) { src = W.sig + NX * k + (i - n2); dst = S.Hc + 72 + (i - n2); }
    else if (i < n4) { src = W.gh + NX * k + (i - n3); dst = S.ghs + (i - n3); }
    else { src = W.c + NE * k + (i - n4); dst = S.cv + (i - n4); }
    cp_async8(dst, src);
  }
}

// S.M += W~_kk + delta_w I, added block by block from the staged stage records (the 44 x 44 block is
// never stored).  Items: 49 index pairs (i, j) of the 7 x 7 kinematic curvature matrices, each expanded
// through the integrator into the 16 blocks of the joint-variable square (um, q, dq, ddq)^2; the 8 x 8
// y-block; the remaining diagonal; the velocity / acceleration tracking cross terms.  Every entry of
// S.M is touched by at most one item.
constexpr int W_ITEMS = 49 + 64 + 8 + 24;
BMPC_DEV void add_W(const Ctx cx, const Config& C, const KktCoef& kc, Smem& S, int k, double delta_w) {
  const double* Hk = S.Hn[k & 1];          // HQQN 0, HQDN 49, HQQK 98, HQDK 147
  const double* Hx = S.Hn[(k + 1) & 1];    // same of stage k + 1
  const double* dpd = S.Hc + 64;
  const double* sg = S.Hc + 72;
  const bool has_next = k + 1 < C.N;
  PAR_FOR(it, W_ITEMS) {
    if (it < 49) {
      const int i = it / 7, j = it - 7 * i, ij = it, ji = j * 7 + i;
      const double al[4] = {C.a_um, 1.0, C.a_dq, C.a_ddq}, be[4] = {C.b_um, 0.0, 1.0, C.b_ddq};
      const int off[4] = {oU, oQ, oDQ, oDDQ};
      double hqq = 0.0, hqd = 0.0, hdq = 0.0, kqq = 0.0, kqd = 0.0, kdq = 0.0;
      if (has_next) { hqq = Hx[ij]; hqd = Hx[49 + ij]; hdq = Hx[49 + ji]; kqq = Hx[98 + ij]; kqd = Hx[147 + ij]; kdq = Hx[147 + ji]; }
      const double own = C.a_u * C.a_u * Hk[ij] + C.a_u * C.b_u * (Hk[49 + ij] + Hk[49 + ji]);
#pragma unroll
      for (int ta = 0; ta < 4; ta++)
#pragma unroll
        for (int tb = 0; tb < 4; tb++) {
          double v = al[ta] * al[tb] * hqq + al[ta] * be[tb] * hqd + be[ta] * al[tb] * hdq;
          if (ta == 0 && tb == 0) v += own;
          if (ta == 1 && tb == 1) v += kqq;
          if (ta == 1 && tb == 2) v += kqd;
          if (ta == 2 && tb == 1) v += kdq;
          const int a = off[ta] + i;
          if (ta == tb && i == j) v += wdiag_const(kc, dpd, a, has_next) + sg[a] + delta_w;
          S.M[a * LDM + off[tb] + j] += v;
        }
    } else if (it < 113) {
      const int q = it - 49, ya = q >> 3, yb = q & 7, a = yrow(ya), b = yrow(yb);
      double v = S.Hc[q];
      if (a == b) v += sg[a] + delta_w;
      S.M[a * LDM + b] += v;
    } else if (it < 121) {
      const int q = it - 113, a = q == 0 ? oUPHI : (q == 7 ? oDDPHI : oVLIN + q - 1);
      S.M[a * LDM + a] += wdiag_const(kc, dpd, a, has_next) + sg[a] + delta_w;
    } else {
      const int q = it - 121, wh = q / 6, m = q - 6 * wh;
      const double v = (wh < 2 ? -2 * kc.w2 : -2 * kc.w5 * kc.idt) * dpd[m];
      const int a = oVLIN + m, b = wh < 2 ? oDPHI : oDDPHI;
      if (wh & 1) S.M[b * LDM + a] += v; else S.M[a * LDM + b] += v;
    }
  }
}

// Products with the constant rows of G (the integrator): they act on whole index blocks with scalar
// coefficients, so one item handles the three x-rows (q_j, dq_j, ddq_j) — or
// (phi, dphi, ddphi) for j = 7 — of one vector and emits the five columns (um_j, q_j, dq_j, ddq_j, u_j).
struct TrivOut { double um, q, dq, ddq, u; };
BMPC_DEV TrivOut triv_combine(const Config& C, double m1, double m2, double m3) {
  TrivOut o;
  o.um = C.a_um * m1 + C.b_um * m2 + C.c_um * m3;
  o.q = m1;
  o.dq = C.a_dq * m1 + m2;
  o.ddq = C.a_ddq * m1 + C.b_ddq * m2 + m3;
  o.u = C.a_u * m1 + C.b_u * m2 + C.c_u * m3;
  return o;
}
// x-rows (0..35) and z-columns (0..51) addressed by item index j (0..6 joints, 7 = path parameter)
BMPC_DEV int trow(int j, int t) { return j < 7 ? 7 * t + j : 33 + t; }              // t = 0, 1, 2
BMPC_DEV int tcol(int j, int t) { return j < 7 ? (t == 0 ? oU : (t == 1 ? oQ : (t == 2 ? oDQ : oDDQ))) + j : (t == 0 ? oUPHI : oPHI + t - 1); }   // t = 0..3

// b := Q_uu^{-1} b by an 8 x 8 Cholesky factorisation in registers (L kept with the RECIPROCAL diagonal) and two
// triangular solves.  Every thread that has a column to solve factorises Q_uu itself from shared memory (no
// broadcast, no extra barrier; the verdict is the same in all of them, so the inertia test is CTA-uniform).
// Factorisation and forward substitution run in ONE unrolled pass: row j of the substitution only needs pivot j and
// the columns < j of L, so its dependency chain runs in the shadow of the factorisation's instead of after it.
// (device: rsqrt, 1 ulp; eight inlined IEEE sqrt + divide pairs were 6.5 KB of the stage's instruction stream)
BMPC_DEV bool chol8_solve(const double* Qu, double (&b)[NU]) {
  double A[NU][NU];
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
#if defined(BMPC_HOST_EMU) || defined(BMPC_IEEE_CHOL)
    const double inv = 1.0 / sqrt(d > 1e-14 ? d : 1.0);
#else
    const double inv = rsqrt(d > 1e-14 ? d : 1.0);
#endif
    A[j][j] = inv;
#pragma unroll
    for (int i = j + 1; i < NU; i++) {
      double e = A[i][j];
#pragma unroll
      for (int l = 0; l < j; l++) e -= A[i][l] * A[j][l];
      A[i][j] = e * inv;
    }
    double a = b[j];
#pragma unroll
    for (int l = 0; l < j; l++) a -= A[j][l] * b[l];
    b[j] = a * inv;
  }
#pragma unroll
  for (int i = NU - 1; i >= 0; i--) {
    double a = b[i];
#pragma unroll
    for (int l = i + 1; l < NU; l++) a -= A[l][i] * b[l];
    b[i] = a * A[i][i];
  }
  return ok;
}

// One backward Riccati step for stage k.  On entry S.M = P_{k+1} (zero for k = N-1), S.pv = p_{k+1}.
// On exit S.M = P_k, S.pv = p_k, gains stored in W.Kk / W.kap.  Returns false if Q_uu is not PD.
// Every product with G = [A_hat | B] is split into the constant integrator rows (a cheap pass, see
// triv_combine) and the 12 dense kinematic rows, which run as 8 x 8 DMMA tiles on top of the result of
// the pass (mma_rowblock: A fragments shared by the tiles of a row block).
BMPC_DEV bool riccati_stage(const Ctx cx, const Config& C, const Work& W, const KktCoef& kc, Smem& S, int k, double delta_w) {
  const bool first = k == 0;      // stage 0: the previous block is fixed -> only the u-columns
  const double* GKs = S.GK;
  const double* Hk = S.Hn[k & 1];
  const int nw = ctx_nwarps(cx);
  // ---- phase 1: M = P_{k+1} + W~_kk + delta_w I
  gk_load(cx, W, S, k);
  add_W(cx, C, kc, S, k, delta_w);
  PAR_FOR(i, NX) S.mv[i] = S.ghs[i] + S.pv[i];
  PAR_FOR(m, 6) S.odv[m] = 2 * kc.w5 * S.Hc[64 + m] * kc.idt;
  BMPC_SYNC();
  BMPC_TMARK(8);
  const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
  // ---- phase 2a: integrator part of YZ = M[:, x] G (44 x 52) and t = M[:, x] c + m
  PAR_FOR(it, NX * 8) {
    const int i = it >> 3, j = it & 7;
    const double* Mr = S.M + i * LDM + 8;
    const TrivOut o = triv_combine(C, Mr[trow(j, 0)], Mr[trow(j, 1)], Mr[trow(j, 2)]);
    double* Yr = S.YZ + i * LDY;
    Yr[tcol(j, 0)] = o.um; Yr[tcol(j, 1)] = o.q; Yr[tcol(j, 2)] = o.dq; Yr[tcol(j, 3)] = o.ddq; Yr[NX + j] = o.u;
    Yr[oPPOS + j] = 0.0;                      // columns p_pos, p_rot, v have no integrator rows
    if (j < 4) Yr[oPPOS + 8 + j] = 0.0;
  }
  PAR_FOR(i, 64) S.Muu[i] = S.M[(i >> 3) * LDM + (i & 7)];
  PAR_FOR(i, NX) {
    const double* Mr = S.M + i * LDM + 8;
    double a0 = S.mv[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int l = 0; l < NE; l += 4) { a0 += Mr[l] * S.cv[l]; a1 += Mr[l + 1] * S.cv[l + 1]; a2 += Mr[l + 2] * S.cv[l + 2]; a3 += Mr[l + 3] * S.cv[l + 3]; }
    S.tv[i] = (a0 + a1) + (a2 + a3);
  }
  cp_async_wait();   // S.GK
  BMPC_SYNC();
  BMPC_TMARK(9);
  // ---- phase 2b: YZ += M[:, kin] GK: 6 row blocks x (4 + 3) column tiles, 3 k-steps
  for (int per = (12 + nw - 1) / nw, t = ctx_warp(cx) * per; t < 12 && t < (ctx_warp(cx) + 1) * per; t++) {
    const int ti = t >> 1, tj0 = (t & 1) ? 4 : 0, nt = (t & 1) ? 3 : 4;
    mma_rowblock<3, 4>(cx, nt,
        [&](int r, int kk) { const int i = 8 * ti + r; return S.M[(i < NX ? i : NX - 1) * LDM + 8 + rKIN + kk]; },   // (rows >= 44 of the last row block are discarded)
        [&](int tt, int kk, int c) { return GKs[kk * NZ + 8 * (tj0 + tt) + c]; },
        [&](int tt, int r, int c) { const int i = 8 * ti + r, col = 8 * (tj0 + tt) + c; return (i < NX && col < NZ) ? S.YZ[i * LDY + col] : 0.0; },
        [&](int tt, int r, int c, double v) { const int i = 8 * ti + r, col = 8 * (tj0 + tt) + c; if (i < NX && col < NZ) S.YZ[i * LDY + col] = v; });
  }
  BMPC_SYNC();
  BMPC_TMARK(10);
  // ---- phase 3: columns u_k of Q = E^T M E + O-terms  (52 x 8): 7 tiles
  {
    // row tiles ti0.., dealt round-robin; the (up to two) tiles of a warp run interleaved
    const int ti0 = first ? 5 : 0;
    for (int wti = ti0 + ctx_warp(cx); wti < 7; wti += 2 * nw)
    mma_colblock<3, 2>(cx, wti + nw < 7 ? 2 : 1,
        [&](int tt, int r, int kk) { return GKs[kk * NZ + 8 * (wti + tt * nw) + r]; },
        [&](int kk, int c) { return S.YZ[(8 + rKIN + kk) * LDY + NX + c]; },
        [&](int tt, int r, int j, double v) {
          const int a = 8 * (wti + tt * nw) + r;
          if (a >= NZ || (first && a < NX)) return;
#pragma unroll
          for (int q = 0; q < 3; q++) v += S.tcc[3 * a + q] * S.YZ[(8 + S.tcr[3 * a + q]) * LDY + NX + j];
          if (a < NX) {
            v += S.YZ[j * LDY + a] + ou_entry(Hk, S.alc, S.bec, j, a);
            if (a >= oVLIN && a < oVLIN + 6) v += GKs[(6 + a - oVLIN) * NZ + NX + j] * ovv + (j == 7 ? C.c_u * S.odv[a - oVLIN] : 0.0);
          } else {
            const int i = a - NX;
            v += S.YZ[i * LDY + NX + j] + S.YZ[j * LDY + NX + i] + S.Muu[i * 8 + j];
          }
          S.Qu[a * LDQ + j] = v;
        });
    // q = G^T t + U^T t_u + [O_x^T c ; 0]
    PAR_FOR(ai, first ? NU : NZ) {
      const int a = first ? NX + ai : ai;
      double v = 0.0;
#pragma unroll
      for (int r = 0; r < NK; r++) v += GKs[r * NZ + a] * S.tv[8 + rKIN + r];
#pragma unroll
      for (int q = 0; q < 3; q++) v += S.tcc[3 * a + q] * S.tv[8 + S.tcr[3 * a + q]];
      if (a >= NX) v += S.tv[a - NX];
      if (!first && a >= oVLIN && a < oVLIN + 6) v += ovv * S.cv[27 + (a - oVLIN)] + S.odv[a - oVLIN] * S.cv[35];
      S.qv[a] = v;
    }
  }
  // Q_ss pass (S.M is dead: phase 3 reads Q_uu's M block from S.Muu).  It needs Y only and is needed by phase 5 only, so it
  // can run in phase 3 on all warps or (C.qss_late) in phase 4 on the warps that have no gain column to solve.
  auto qss_pass = [&](int w0, int w1) {
  // Q_ss, integrator rows of G^T Y: item (j, b) -> rows um_j, q_j, dq_j, ddq_j of column b; O-terms:
  // rows v_{k+1} x cols v_k carry ovv, row ddphi_{k+1} carries odv (E^T O [I 0] + transpose)
  ROLE_FOR(it, 8 * NX, w0, w1) {
    const int j = it / NX, b = it - NX * j;
    const TrivOut o = triv_combine(C, S.YZ[(8 + trow(j, 0)) * LDY + b], S.YZ[(8 + trow(j, 1)) * LDY + b], S.YZ[(8 + trow(j, 2)) * LDY + b]);
    double vv[4] = {o.um, o.q, o.dq, o.ddq};
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int a = tcol(j, t);
      double v = vv[t];
      if (b >= oVLIN && b < oVLIN + 6) {
        const double g35 = a == oUPHI ? C.c_um : (a == oDDPHI ? 1.0 : 0.0);
        v += GKs[(6 + b - oVLIN) * NZ + a] * ovv + g35 * S.odv[b - oVLIN];
      }
      S.M[a * LDM + b] = v;
    }
  }
  // rows p_pos, p_rot, v: no integrator part.  Phase 5 reads Q_ss only through the upper tiles (tile row <= tile
  // column) and writes P_k symmetrically, so of these 12 rows only the columns from the first column of their tile
  // row on are needed: 3 x 20 + 8 x 12 + 4 = 160 entries instead of 528.
  ROLE_FOR(it, 160, w0, w1) {
    int a, b;
    if (it < 60) { a = 29 + it / 20; b = 24 + it - 20 * (it / 20); }
    else if (it < 156) { const int q = it - 60; a = 32 + q / 12; b = 32 + q - 12 * (q / 12); }
    else { a = 40; b = 40 + it - 156; }
    double v = 0.0;
    if (b >= oVLIN && b < oVLIN + 6) v += GKs[(6 + b - oVLIN) * NZ + a] * ovv;
    if (a >= oVLIN && a < oVLIN + 6) {
      const double g35 = b == oUPHI ? C.c_um : (b == oDDPHI ? 1.0 : 0.0);
      v += GKs[(6 + a - oVLIN) * NZ + b] * ovv + g35 * S.odv[a - oVLIN];
    }
    S.M[a * LDM + b] = v;
  }
  };
  const bool qss_late = C.qss_late && nw > 2;
  if (!first && !qss_late) qss_pass(0, nw);
  BMPC_SYNC();
  BMPC_TMARK(11);
  // ---- phase 4, warps 0-1: gains K = -Q_uu^{-1} Q_us (8 x 44), kappa = -Q_uu^{-1} q_u (one column per
  // thread, each with its own register copy of the 8 x 8 factor).  Other warps: stage the data of stage k - 1.
  double* K = W.Kk + (size_t)k * NU * NX;
  double* kap = W.kap + k * NU;
  // (factorisation and forward substitution fused, see chol8_solve; a thread without a column has nothing to do)
  ROLE_FOR(col, (first ? 0 : NX) + 1, 0, 2) {
    const bool isk = col == (first ? 0 : NX);
    double bcol[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) bcol[i] = isk ? -S.qv[NX + i] : -S.Qu[col * LDQ + i];
    if (!chol8_solve(S.Qu, bcol)) S.flag[3] = 1;
#pragma unroll
    for (int i = 0; i < NU; i++) {
      if (isk) { kap[i] = bcol[i]; S.kapv[i] = bcol[i]; }
      else { K[i * NX + col] = bcol[i]; S.Ks[i * NX + col] = bcol[i]; }
    }
  }
  BMPC_TMARK(42);
  BMPC_TMARK2(48);
  if (!first) {
    if (in_role(cx, 2, nw)) stage_prefetch(cx, C, W, S, k - 1, 2, nw);   // asynchronous, completed below
    BMPC_TMARK2(49);
    if (qss_late) qss_pass(2, nw);
    cp_async_wait();
    BMPC_TMARK2(52);
  }
  BMPC_SYNC();
  BMPC_TMARK(12);
  BMPC_TMARK2(44);
  if (S.flag[3]) { BMPC_SYNC(); return false; }   // (barrier: nobody re-arms the flag before everyone has read it)
  if (first) return true;
  BMPC_TMARK(40);
  // ---- phase 5: P_k = Q_ss + Q_su K (symmetric: upper tiles, mirrored), p_k = q_s + Q_su kappa
  for (int ti = ctx_warp(cx); ti < 6; ti = nw == 4 ? ((ti == 2 || ti == 3) ? 7 - ti : 6) : ti + nw) {   // 4 warps: 6, 5, 4 + 1, 3 + 2 tiles
    const int nt = 6 - ti;
    auto fa = [&](int r, int kk) { return kk < NK ? GKs[kk * NZ + 8 * ti + r] : S.Qu[(8 * ti + r) * LDQ + kk - NK]; };
    auto fb = [&](int tt, int kk, int c) { return kk < NK ? S.YZ[(8 + rKIN + kk) * LDY + 8 * (ti + tt) + c] : S.Ks[(kk - NK) * NX + 8 * (ti + tt) + c]; };
    auto fc = [&](int tt, int r, int c) { const int a = 8 * ti + r, b = 8 * (ti + tt) + c; return (a < NX && b < NX) ? S.M[a * LDM + b] : 0.0; };
    auto fe = [&](int tt, int r, int c, double v) {
      const int a = 8 * ti + r, b = 8 * (ti + tt) + c;
      if (a >= NX || b >= NX || a > b) return;
      S.M[a * LDM + b] = v;
      S.M[b * LDM + a] = v;
    };
    // (a pass walks through the code of all NT tile slots whatever nt is: the short row blocks, which are the second
    // pass of their warps, get narrower instantiations)
    if (nt > 4) mma_rowblock<5, 6>(cx, nt, fa, fb, fc, fe);
    else if (nt > 2) mma_rowblock<5, 4>(cx, nt, fa, fb, fc, fe);
    else mma_rowblock<5, 2>(cx, nt, fa, fb, fc, fe);
  }
  BMPC_TMARK(30);
  BMPC_TMARK2(45);
  PAR_FOR(a, NX) {
    double v = S.qv[a];
#pragma unroll
    for (int i = 0; i < NU; i++) v += S.Qu[a * LDQ + i] * S.kapv[i];
    S.pv[a] = v;
  }
  BMPC_TMARK(31);
  BMPC_TMARK2(46);
  BMPC_SYNC();
  BMPC_TMARK(13);
  BMPC_TMARK2(47);
  return true;
}

// Operands of the single-warp sweeps (gains, kinematic rows, residuals) live in the global workspace: read where they
// are used, every stage pays three to four L2 round trips on the critical path of the solve.  They are staged in the
// shared memory of the idle Riccati blocks instead (S.ev[0 .. 3 FS_SIZE), below the sweep scratch S.YZ[0 .. 256)):
// the warps that have no part in the sweep copy stage k + 2 with cp.async while warp 0 works on stage k.
constexpr int FS_K = 0, FS_KAP = NU * NX, FS_GK = FS_KAP + NU, FS_C = FS_GK + NK * NZ, FS_SIZE = FS_C + NE;   // 1020 doubles
static_assert(FS_KAP % 2 == 0 && FS_GK % 2 == 0 && FS_C % 2 == 0 && FS_SIZE % 2 == 0 && R_SIZE % 2 == 0 && NE % 2 == 0, "16-byte staging");
static_assert(3 * FS_SIZE <= NX * LDM + (NK * NZ + 4) + 2 * (4 * 49 + 4) + (64 + 8 + NX) + NX + 64, "sweep staging overlaps S.YZ");
BMPC_NOINLINE void fs_stage(const Ctx cx, const Work& W, double* buf, int k, int w0, int w1) {   // (one copy: three call sites)
  // 16-byte copies: every segment starts at an even offset of the 256-byte aligned workspace slice / of S.ev
  const double* Kg = W.Kk + (size_t)k * NU * NX;
  const double* kg = W.kap + k * NU;
  const double* rg = W.rec + (size_t)k * R_SIZE + R_GK;
  const double* cg = W.c + NE * k;
#pragma unroll 2
  ROLE_FOR(h, FS_SIZE / 2, w0, w1) {
    const int i = 2 * h;
    cp_async16(buf + i, i < FS_KAP ? Kg + i : (i < FS_GK ? kg + (i - FS_KAP) : (i < FS_C ? rg + (i - FS_GK) : cg + (i - FS_C))));
  }
}

// Forward sweep: du_k = kappa_k + K_k ds_k, dx_{k+1} = G_k (ds_k, du_k) + c_k.  Called by the whole CTA: warp 0 runs the
// recursion (warp-level barriers only; the running (ds, du) vector lives in shared memory), the other warps stage.
// Lane roles per stage: (A) 8 rows of K x 4 column quarters; (B) lanes 0-7 finish du; (C) lanes 0-23 = 12 kinematic
// rows x 2 column halves, lanes 24-31 = the three integrator rows of joint j = lane - 24 (7 = path parameter, same
// code for all eight lanes); (D) lanes 0-11 finish the kinematic rows.
BMPC_DEV void forward_sweep(const Ctx cx, const Config& C, const Work& W, Smem& S) {
  double* zb = S.YZ;            // [2][64]: z = (ds (44), du (8)) of the current / next stage
  double* part = S.YZ + 128;    // [32] partial sums
  double* ring = S.ev;          // [3][FS_SIZE]
  const int N = C.N, nw = ctx_nwarps(cx);
#ifdef BMPC_HOST_EMU
  const int p0 = 0;
#else
  const int p0 = 1;             // staging warps: [p0, nw)
#endif
  const bool lead = in_role(cx, 0, 1), prod = in_role(cx, p0, nw);
  if (lead) LANE_FOR(l, 128) zb[l] = 0.0;
  if (prod) {
    fs_stage(cx, W, ring, 0, p0, nw);
    cp_async_commit();
    if (N > 1) fs_stage(cx, W, ring + FS_SIZE, 1, p0, nw);
    cp_async_commit();
    cp_async_wait_group1();
  }
  BMPC_SYNC();
  for (int k = 0; k < N; k++) {
    if (prod) {
      if (k + 2 < N) fs_stage(cx, W, ring + FS_SIZE * ((k + 2) % 3), k + 2, p0, nw);   // (that buffer held stage k - 1)
      cp_async_commit();
    }
    if (lead) {
      const double* b = ring + FS_SIZE * (k % 3);
      const double* K = b + FS_K;
      const double* GK = b + FS_GK;
      double* z = zb + 64 * (k & 1);
      double* zn = zb + 64 * ((k + 1) & 1);
      double* dw = W.dx + NX * k;
      LANE_FOR(l, 32) {           // (A) lane -> (row i of K, quarter of the columns)
        const int i = l >> 2, c0 = 11 * (l & 3);
        double a = 0.0;
        if (k > 0) {
#ifdef BMPC_SPLIT_ACC
          double a1 = 0.0;
#pragma unroll
          for (int j = 0; j < 10; j += 2) { a += K[i * NX + c0 + j] * z[c0 + j]; a1 += K[i * NX + c0 + j + 1] * z[c0 + j + 1]; }
          a = (a + K[i * NX + c0 + 10] * z[c0 + 10]) + a1;
#else
#pragma unroll
          for (int j = 0; j < 11; j++) a += K[i * NX + c0 + j] * z[c0 + j];
#endif
        }
        part[l] = a;
      }
      BMPC_WSYNC();
      LANE_FOR(i, NU) {           // (B)
        const double du = b[FS_KAP + i] + ((part[4 * i] + part[4 * i + 1]) + (part[4 * i + 2] + part[4 * i + 3]));
        z[NX + i] = du; zn[i] = du; dw[i] = du;
      }
      BMPC_WSYNC();
      LANE_FOR(l, 32) {           // (C)
        if (l < 24) {
          const int r = l % 12, c0 = 26 * (l / 12);
#ifdef BMPC_SPLIT_ACC
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int j = 0; j < 24; j += 4) {
            a0 += GK[r * NZ + c0 + j] * z[c0 + j]; a1 += GK[r * NZ + c0 + j + 1] * z[c0 + j + 1];
            a2 += GK[r * NZ + c0 + j + 2] * z[c0 + j + 2]; a3 += GK[r * NZ + c0 + j + 3] * z[c0 + j + 3];
          }
          a0 += GK[r * NZ + c0 + 24] * z[c0 + 24]; a1 += GK[r * NZ + c0 + 25] * z[c0 + 25];
          part[l] = (a0 + a1) + (a2 + a3);
#else
          double a = 0.0;
#pragma unroll
          for (int j = 0; j < 26; j++) a += GK[r * NZ + c0 + j] * z[c0 + j];
          part[l] = a;
#endif
        } else {
          const int j = l - 24;
          const double um = z[tcol(j, 0)], q = z[tcol(j, 1)], dq = z[tcol(j, 2)], ddq = z[tcol(j, 3)], u = z[NX + j];
          const int r0 = trow(j, 0), r1 = trow(j, 1), r2 = trow(j, 2);
          // (same order of operations as G_vec)
          double g0 = C.a_u * u; g0 += q + C.a_dq * dq + C.a_ddq * ddq + C.a_um * um;
          double g1 = C.b_u * u; g1 += dq + C.b_ddq * ddq + C.b_um * um;
          double g2 = C.c_u * u; g2 += ddq + C.c_um * um;
          const double v0 = b[FS_C + r0] + g0, v1 = b[FS_C + r1] + g1, v2 = b[FS_C + r2] + g2;
          zn[8 + r0] = v0; dw[8 + r0] = v0; zn[8 + r1] = v1; dw[8 + r1] = v1; zn[8 + r2] = v2; dw[8 + r2] = v2;
        }
      }
      BMPC_WSYNC();
      LANE_FOR(l, NK) {           // (D)
        const double v = b[FS_C + rKIN + l] + (part[l] + part[l + 12]);
        zn[8 + rKIN + l] = v; dw[8 + rKIN + l] = v;
      }
      BMPC_WSYNC();
    }
    if (prod) cp_async_wait_group1();   // stage k + 1 has landed
    BMPC_SYNC();
  }
}

// Right-hand side of the adjoint recursion for the equality multipliers, all stages in parallel:
//   r_k = [W~_kk dw_k + O_k dw_{k-1} + O_{k+1}^T dw_{k+1} + g^_k]_x
// with the products taken block by block from the stage records (see add_W).  The curvature part
// uses d q_n = dw_{k+1}[q] - c_{k+1}[q rows] (the linearised change of the integrated state).
BMPC_DEV void adjoint_rhs(const Ctx cx, const Config& C, const Work& W, const KktCoef& kc, const Smem& S, double delta_w) {
  const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
  // Items in row-class order (joint rows q / dq / ddq, the y rows, the v rows, ddphi), stage-minor: the classes take
  // different branches below, each with its own reads from the global stage records, and a warp whose lanes sit in
  // one class pays one such round trip instead of one per class.
  const int N = C.N;
  PAR_FOR(it, N * NE) {
    int k, r;
    if (it < 21 * N) { const int t = it / (7 * N), q = it - 7 * N * t; k = q / 7; r = 8 + 7 * t + (q - 7 * k); }
    else if (it < 29 * N) { const int q = it - 21 * N; k = q >> 3; r = yrow(q & 7); }
    else if (it < 35 * N) { const int q = it - 29 * N; k = q / 6; r = oVLIN + (q - 6 * k); }
    else { k = it - 35 * N; r = oDDPHI; }
    const bool has_next = k + 1 < C.N;
    const double* rec = W.rec + (size_t)k * R_SIZE;
    const double* rn = rec + R_SIZE;
    const double* dw = W.dx + NX * k;
    const double* dwp = dw - NX;
    const double* dwn = dw + NX;
    double a = W.gh[NX * k + r] + (delta_w + W.sig[NX * k + r] + wdiag_const(kc, rec + R_DPD, r, has_next)) * dw[r];
    const int ya = yidx(r);
    if (ya >= 0) {
#pragma unroll
      for (int yb = 0; yb < 8; yb++) a += rec[R_HYB + ya * 8 + yb] * dw[yrow(yb)];
    }
    if (r >= oVLIN && r < oVLIN + 6) {
      const int m = r - oVLIN;
      a += -2 * kc.w2 * rec[R_DPD + m] * dw[oDPHI] - 2 * kc.w5 * kc.idt * rec[R_DPD + m] * dw[oDDPHI];
      if (k > 0) a += ovv * dwp[r];
      if (has_next) a += ovv * dwn[r] + 2 * kc.w5 * rn[R_DPD + m] * kc.idt * dwn[oDDPHI];
    }
    if (r == oDPHI) for (int m = 0; m < 6; m++) a += -2 * kc.w2 * rec[R_DPD + m] * dw[oVLIN + m];
    if (r == oDDPHI) {
      for (int m = 0; m < 6; m++) a += -2 * kc.w5 * kc.idt * rec[R_DPD + m] * dw[oVLIN + m];
      if (k > 0) for (int m = 0; m < 6; m++) a += 2 * kc.w5 * rec[R_DPD + m] * kc.idt * dwp[oVLIN + m];
    }
    const int t = jtype(r);
    if (t > 0 && has_next) {
      const int i = jidx(r);
      const double* cn = W.c + NE * (k + 1);
      double hq = 0.0, hd = 0.0, hk = 0.0;
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const double dqh = dwn[oQ + j] - cn[j], ddh = dwn[oDQ + j] - cn[7 + j];
        hq += rn[R_HQQN + i * 7 + j] * dqh + rn[R_HQDN + i * 7 + j] * ddh;
        hd += rn[R_HQDN + j * 7 + i] * dqh;
        if (t == 1) hk += rn[R_HQQK + i * 7 + j] * dw[oQ + j] + rn[R_HQDK + i * 7 + j] * dw[oDQ + j];
        if (t == 2) hk += rn[R_HQDK + j * 7 + i] * dw[oQ + j];
      }
      a += S.alc[t] * hq + S.bec[t] * hd + hk;
    }
    W.ynew[NE * k + r - 8] = a;
  }
}

// Adjoint recursion y_k = r_k + [A_hat_{k+1}^T y_{k+1}]_x, executed by ONE warp.  Of the columns 8..43 of the kinematic
// rows GK only 8..28 (q, dq, ddq of the previous block) are ever non-zero, plus the identity of the p_rot columns
// (phase_kin_jacobian_init / phase_kin_jacobian): those 12 x 21 entries of the stages N-1, N-2, ... are staged in shared
// memory by the whole CTA (adjoint_stage, issued before the right-hand side is built, complete after it) as far as the
// idle Riccati blocks have room (12 stages); earlier stages of longer horizons are read from the workspace.
constexpr int AS_COLS = 21, AS_SIZE = NK * AS_COLS, AS_MAX = 12;
static_assert(AS_MAX * AS_SIZE <= 3 * FS_SIZE, "adjoint staging overlaps S.YZ");
BMPC_DEV int adjoint_slots(int N) { return N - 1 < AS_MAX ? N - 1 : AS_MAX; }
BMPC_DEV void adjoint_stage(const Ctx cx, const Config& C, const Work& W, Smem& S) {
  const int ns = adjoint_slots(C.N);
#pragma unroll 1
  PAR_FOR(it, ns * AS_SIZE) {
    const int slot = it / AS_SIZE, q = it - AS_SIZE * slot, r = q / AS_COLS, cc = q - AS_COLS * r;
    cp_async8(S.ev + it, W.rec + (size_t)(C.N - 1 - slot) * R_SIZE + R_GK + r * NZ + 8 + cc);
  }
}
BMPC_DEV void adjoint_sweep(const Ctx cx, const Config& C, const Work& W, Smem& S) {
  double* yb = S.YZ;            // [2][40]
  const int N = C.N, ns = adjoint_slots(N);
  LANE_FOR(i, NE) yb[40 * ((N - 1) & 1) + i] = W.ynew[NE * (N - 1) + i];
  BMPC_WSYNC();
  for (int k = N - 2; k >= 0; k--) {
    const int slot = N - 2 - k;   // of stage k + 1
    const double* g = slot < ns ? S.ev + slot * AS_SIZE : W.rec + (size_t)(k + 1) * R_SIZE + R_GK + 8;
    const int gs = slot < ns ? AS_COLS : NZ;
    const double* yn = yb + 40 * ((k + 1) & 1);
    double* yc = yb + 40 * (k & 1);
    LANE_FOR(i, NE) {
      // (G^T y)[8 + i] with the terms in the order of GT_tab; the structural zeros contribute nothing
      double a = 0.0;
      if (i < AS_COLS) {
#pragma unroll
        for (int r = 0; r < NK; r++) a += g[r * gs + i] * yn[rKIN + r];
      } else if (i >= oPROT - 8 && i < oPROT - 8 + 3) a += yn[i];
#pragma unroll
      for (int t = 0; t < 3; t++) a += S.tcc[3 * (8 + i) + t] * yn[S.tcr[3 * (8 + i) + t]];
      const double v = W.ynew[NE * k + i] + a;
      W.ynew[NE * k + i] = v;
      yc[i] = v;
    }
    BMPC_WSYNC();
  }
}

// Backward + forward + adjoint sweeps: dx (primal step) and ynew (equality multipliers of the full
// step).  kkt_prepare must have run for the current iterate.
BMPC_NOINLINE bool kkt_solve(const Ctx cx, const Config& C, const Work& W, const double* p, Smem& S, double delta_w) {
  const KktCoef kc = kkt_coef(C, p);
  PAR_FOR(i, NX * NX) S.M[i] = 0.0;
  PAR_FOR(i, NX) S.pv[i] = 0.0;
  if (cx.tid == 0) S.flag[3] = 0;
  stage_prefetch(cx, C, W, S, C.N - 1, 0, ctx_nwarps(cx));
  cp_async_wait();
  BMPC_SYNC();
  BMPC_TMARK(7);
  for (int k = C.N - 1; k >= 0; k--)
    if (!riccati_stage(cx, C, W, kc, S, k, delta_w)) return false;
  forward_sweep(cx, C, W, S);
  BMPC_TMARK(14);
  adjoint_stage(cx, C, W, S);
  adjoint_rhs(cx, C, W, kc, S, delta_w);
  cp_async_wait();
  BMPC_SYNC();
  BMPC_TMARK(15);
  if (ctx_warp(cx) == 0) adjoint_sweep(cx, C, W, S);
  BMPC_SYNC();
  BMPC_TMARK(16);
  return true;
}

}  // namespace bmpc
