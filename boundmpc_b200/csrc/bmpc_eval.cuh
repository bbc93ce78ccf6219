// boundmpc_b200 — dense export of the NLP derivatives of one instance (parity tests only;
// what nlp_grad_f / nlp_jac_g / nlp_hess_l return, casadi_ocp_formulation.py:389).
#pragma once
#include "bmpc_ipm.cuh"

namespace bmpc {

struct EvalIO {
  const double* x;     // [n]
  const double* p;     // [np]
  const double* lam;   // [48 N]
  double *f, *g, *d, *grad, *jac, *hess;   // may be null
};

BMPC_DEV void eval_instance(const Ctx cx, const Config& C, const Work& W, Smem& S, const EvalIO& io) {
  const int N = C.N, n = C.n;
  build_wp0(cx, C, io.p, W.wp0);
  PAR_FOR(i, n) { W.x[i] = io.x[i]; W.zL[i] = 0.0; W.zU[i] = 0.0; }
  PAR_FOR(i, NE * N) { const int k = i / NE, r = i - NE * k; W.y[i] = io.lam ? io.lam[(NE + ND) * k + r] : 0.0; }
  PAR_FOR(i, ND * N) { const int k = i / ND, r = i - ND * k; W.zs[i] = io.lam ? io.lam[(NE + ND) * k + NE + r] : 0.0; W.s[i] = 1.0; }
  BMPC_SYNC();
  eval_full(cx, C, W, io.p, W.x);
  phase_path(cx, C, W, io.p, W.x, W.dtr, W.st, 2);
  BMPC_SYNC();
  if (io.f && cx.tid == 0) { double f = 0; for (int k = 0; k < N; k++) f += W.cost[k]; *io.f = f; }
  if (io.g) PAR_FOR(i, NG * N) { const int k = i / NG, r = i - NG * k; io.g[i] = r < NE ? W.c[NE * k + r] : W.st[NQ * k + r - NE]; }
  if (io.d) PAR_FOR(i, ND * N) io.d[i] = W.d[i];
  if (io.grad) PAR_FOR(i, n) io.grad[i] = W.gradf[i];
  if (io.jac) {
    const int rows = (NE + ND) * N;
    PAR_FOR(it, rows * n) {
      const int row = it / n, col = it - n * row;
      const int k = row / (NE + ND), r = row - (NE + ND) * k, kc_ = col / NX, a = col - NX * kc_;
      const double* rec = W.rec + (size_t)k * R_SIZE;
      double v = 0.0;
      if (r < NE) {
        if (kc_ == k) {
          if (a < 8) { double e[NE]; for (int q = 0; q < NE; q++) e[q] = q == r ? 1.0 : 0.0; v = GT_vec(C, rec + R_GK, e, NX + a); }
          else v = (a - 8 == r) ? -1.0 : 0.0;
        } else if (kc_ == k - 1) { double e[NE]; for (int q = 0; q < NE; q++) e[q] = q == r ? 1.0 : 0.0; v = GT_vec(C, rec + R_GK, e, a); }
      } else if (kc_ == k) {
        const int ya = (a >= oPPOS && a < oPPOS + 6) ? a - oPPOS : (a == oPHI ? 6 : (a == oDPHI ? 7 : -1));
        if (ya >= 0) v = rec[R_JD + (r - NE) * 8 + ya];
      }
      io.jac[it] = v;
    }
  }
  if (io.hess) {
    const KktCoef kc = kkt_coef(C, io.p);
    PAR_FOR(i, n * n) io.hess[i] = 0.0;
    BMPC_SYNC();
    double al[5], be[5]; int off[5];
    type_coef(C, al, be, off);
    const double ovv = -2 * kc.w5 * kc.idt * kc.idt;
    for (int k = 0; k < N; k++) {
      const double* rec = W.rec + (size_t)k * R_SIZE;
      PAR_FOR(i, NX * NX) {
        const int a = i / NX, b = i - NX * a;
        io.hess[(size_t)(NX * k + a) * n + NX * k + b] = wd_entry(C, W, kc, al, be, k, a, b);
      }
      if (k > 0) {
        PAR_FOR(i, NX * NX) {
          const int a = i / NX, b = i - NX * a;   // row in w_k, col in w_{k-1}
          double v = 0.0;
          if (a < 8) v = ou_entry(rec + R_HQQN, al, be, a, b);
          else if (a >= oVLIN && a < oVLIN + 6 && b == a) v = ovv;
          else if (a == oDDPHI && b >= oVLIN && b < oVLIN + 6) v = 2 * kc.w5 * rec[R_DPD + b - oVLIN] * kc.idt;
          io.hess[(size_t)(NX * k + a) * n + NX * (k - 1) + b] = v;
          io.hess[(size_t)(NX * (k - 1) + b) * n + NX * k + a] = v;
        }
      }
    }
  }
  BMPC_SYNC();
}


// One Newton step of the interior-point iteration at a given primal-dual point (parity tests only: the Riccati sweep
// against a dense solve of the assembled KKT system).  v [n + 36 N + 12 N + 12 N + n + n] = x, y, s, z_s, z_L,
// z_U; out: dx [n], ynew [36 N] (the equality multipliers of the full step), ok (0: some Q_uu was not positive definite).
struct KktIO { const double* v; const double* p; double mu, delta_w; double* dx; double* ynew; int32_t* ok; };
BMPC_DEV void kkt_step_instance(const Ctx cx, const Config& C, const Work& W, Smem& S, const KktIO& io) {
  const int N = C.N, n = C.n, ne = NE * N, nd = ND * N;
  build_wp0(cx, C, io.p, W.wp0);
  const double* q = io.v;
  PAR_FOR(i, n) { W.x[i] = q[i]; W.zL[i] = q[n + ne + 2 * nd + i]; W.zU[i] = q[2 * n + ne + 2 * nd + i]; }
  PAR_FOR(i, ne) W.y[i] = q[n + i];
  PAR_FOR(i, nd) { W.s[i] = q[n + ne + i]; W.zs[i] = q[n + ne + nd + i]; }
  BMPC_SYNC();
  eval_full(cx, C, W, io.p, W.x);
  kkt_prepare(cx, C, W, io.mu);
  const bool ok = kkt_solve(cx, C, W, io.p, S, io.delta_w);
  PAR_FOR(i, n) io.dx[i] = W.dx[i];
  PAR_FOR(i, ne) io.ynew[i] = W.ynew[i];
  if (cx.tid == 0) *io.ok = ok ? 1 : 0;
  BMPC_SYNC();
}

}  // namespace bmpc
