// TEST INFRASTRUCTURE — CPU oracle for BoundMPC's per-step OCP.  Never linked into the
// product (boundmpc_b200/csrc); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` legs build and call it.
//
// What it restates:
//   * the NLP of casadi_ocp_formulation.py:9-391 (ocp_model.hpp), with exact derivatives by
//     forward AD (ad.hpp) — what CasADi's nlp_grad_f / nlp_jac_g / nlp_hess_l provide
//     (casadi_ocp_formulation.py:389);
//   * the solve behind `self.solver(x0, lbx, ubx, lbg, ubg, p)` (BoundMPC.py:446-457):
//     a primal-dual interior-point method in Ipopt's formulation (slacks for the inequality
//     rows, log barriers with multipliers for bounds, fraction-to-boundary, filter line
//     search with second-order correction, barrier decrease globalised by the kkt-error
//     progress test, inertia correction by Hessian perturbation; Wächter & Biegler 2006 — the
//     published algorithm of Ipopt 3.x, pulled in unpinned through `casadi`,
//     bound_mpc/requirements.txt:1).  The symmetric indefinite KKT system MUMPS factorises
//     is solved here by a stage-wise Riccati recursion (same Newton step).
// Parity status: function values / derivatives are pinned against the reference's own
// Python executed through tests/golden/refexec; converged points are certified by KKT
// residuals computed from those reference-executed derivatives and by SURVEY App. D.4.
// No Ipopt output exists anywhere (casadi is not installable here): "parity unpinned"
// with respect to Ipopt's iterate path; the converged KKT point is what is compared.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <atomic>
static std::atomic<long> g_soc_solves{0}, g_soc_accepted{0};
#include "ocp_model.hpp"

using namespace orc;

namespace {

struct Prob {
  int N, S, n, m;
  double dt;
  Layout L;
  IntCoef ic;
  double lb[NX], ub[NX];
  Prob(int N_, int S_, double dt_) : N(N_), S(S_), n(NX * N_), m(NG * N_), dt(dt_), L(S_), ic(dt_) {
    const double DEG = M_PI / 180.0;
    const double ql[7] = {165, 115, 165, 115, 165, 115, 170};
    const double dql[7] = {85, 85, 100, 75, 130, 135, 135};
    for (int i = 0; i < NX; i++) { lb[i] = -INFINITY; ub[i] = INFINITY; }
    for (int i = 0; i < 8; i++) { lb[i] = -35; ub[i] = 35; }
    for (int i = 0; i < 7; i++) { lb[oQ + i] = -ql[i] * DEG; ub[oQ + i] = ql[i] * DEG; lb[oDQ + i] = -dql[i] * DEG; ub[oDQ + i] = dql[i] * DEG; }
    lb[oPHI] = 0.0;
  }
};

// previous-stage block for stage 0 from the parameter vector
void wprev0(const Prob& P, const double* p, double* wp) {
  const Layout& L = P.L;
  for (int i = 0; i < 8; i++) wp[oU + i] = p[L.jerk + i];
  for (int i = 0; i < 7; i++) { wp[oQ + i] = p[L.q0 + i]; wp[oDQ + i] = p[L.dq0 + i]; wp[oDDQ + i] = p[L.ddq0 + i]; }
  for (int i = 0; i < 6; i++) { wp[oPPOS + i] = p[L.p0 + i]; wp[oVLIN + i] = p[L.v0 + i]; }
  for (int i = 0; i < 3; i++) wp[oPHI + i] = p[L.phi0 + i];
}

struct StageD {          // derivatives of one stage w.r.t. zeta = (wprev(44), w(44))
  double g[NG], cost;
  double d[ND];            // interval-form inequality rows
  double Jdd[ND][2 * NX];
  double gradf[2 * NX];
  double Jg[NG][2 * NX];
  double HL[2 * NX][2 * NX];
};

// values only
void stage_values(const Prob& P, const double* p, const double* wp, const double* w, double* g, double& cost, double* din = nullptr) {
  const IntCoef& ic = P.ic;
  double qn[7], dqn[7], ddqn[7];
  for (int j = 0; j < 7; j++) {
    double q = wp[oQ + j], dq = wp[oDQ + j], ddq = wp[oDDQ + j], um = wp[oU + j], u = w[oU + j];
    qn[j] = q + ic.a_dq * dq + ic.a_ddq * ddq + ic.a_um * um + ic.a_u * u;
    dqn[j] = dq + ic.b_ddq * ddq + ic.b_um * um + ic.b_u * u;
    ddqn[j] = ddq + ic.c_um * um + ic.c_u * u;
    g[j] = qn[j] - w[oQ + j]; g[7 + j] = dqn[j] - w[oDQ + j]; g[14 + j] = ddqn[j] - w[oDDQ + j];
  }
  double pos[3], vl[3], va[3], pk[3], vlk[3], vak[3];
  kinematics<double>(qn, dqn, pos, vl, va);
  kinematics<double>(wp + oQ, wp + oDQ, pk, vlk, vak);
  for (int i = 0; i < 3; i++) {
    g[21 + i] = pos[i] - w[oPPOS + i];
    g[24 + i] = wp[oPROT + i] + 0.5 * ic.h * (vak[i] + va[i]) - w[oPROT + i];
    g[27 + i] = vl[i] - w[oVLIN + i];
    g[30 + i] = va[i] - w[oVANG + i];
  }
  {
    double ph = wp[oPHI], dph = wp[oDPHI], ddph = wp[oDDPHI], um = wp[oUPHI], u = w[oUPHI];
    g[33] = ph + ic.a_dq * dph + ic.a_ddq * ddph + ic.a_um * um + ic.a_u * u - w[oPHI];
    g[34] = dph + ic.b_ddq * ddph + ic.b_um * um + ic.b_u * u - w[oDPHI];
    g[35] = ddph + ic.c_um * um + ic.c_u * u - w[oDDPHI];
  }
  double c[NC];
  for (int i = 0; i < 3; i++) { c[cPPOS + i] = w[oPPOS + i]; c[cPROT + i] = w[oPROT + i]; }
  for (int i = 0; i < 6; i++) { c[cV + i] = w[oVLIN + i]; c[cVPREV + i] = wp[oVLIN + i]; }
  c[cPHI] = w[oPHI]; c[cDPHI] = w[oDPHI]; c[cDDPHI] = w[oDDPHI];
  double dloc[ND];
  path_terms<double>(P.L, p, P.dt, c, cost, g + 36, din ? din : dloc);
  const double* wt = p + P.L.w;
  for (int j = 0; j < 7; j++) {
    double dq_ = w[oQ + j] - p[P.L.qd + j];
    cost += wt[10] * dq_ * dq_ + wt[11] * w[oDQ + j] * w[oDQ + j] + wt[12] * w[oDDQ + j] * w[oDDQ + j] + wt[13] * w[oU + j] * w[oU + j];
  }
  cost += wt[9] * w[oUPHI] * w[oUPHI];
}

// values + all first/second derivatives.
// interval == false: lam = multipliers of the 43 reference rows (Hessian of f + lam.g, as CasADi's nlp_hess_l)
// interval == true : lam = [y(36), z(12)] with z the multipliers of the interval-form rows
void stage_derivs(const Prob& P, const double* p, const double* wp, const double* w, const double* lam, StageD& D, bool interval) {
  const IntCoef& ic = P.ic;
  memset(&D, 0, sizeof(StageD));
  stage_values(P, p, wp, w, D.g, D.cost, D.d);
  // --- linear integrator rows
  for (int j = 0; j < 7; j++) {
    double* r = D.Jg[j];
    r[oQ + j] = 1; r[oDQ + j] = ic.a_dq; r[oDDQ + j] = ic.a_ddq; r[oU + j] = ic.a_um; r[NX + oU + j] = ic.a_u; r[NX + oQ + j] = -1;
    r = D.Jg[7 + j];
    r[oDQ + j] = 1; r[oDDQ + j] = ic.b_ddq; r[oU + j] = ic.b_um; r[NX + oU + j] = ic.b_u; r[NX + oDQ + j] = -1;
    r = D.Jg[14 + j];
    r[oDDQ + j] = 1; r[oU + j] = ic.c_um; r[NX + oU + j] = ic.c_u; r[NX + oDDQ + j] = -1;
  }
  {
    double* r = D.Jg[33];
    r[oPHI] = 1; r[oDPHI] = ic.a_dq; r[oDDPHI] = ic.a_ddq; r[oUPHI] = ic.a_um; r[NX + oUPHI] = ic.a_u; r[NX + oPHI] = -1;
    r = D.Jg[34];
    r[oDPHI] = 1; r[oDDPHI] = ic.b_ddq; r[oUPHI] = ic.b_um; r[NX + oUPHI] = ic.b_u; r[NX + oDPHI] = -1;
    r = D.Jg[35];
    r[oDDPHI] = 1; r[oUPHI] = ic.c_um; r[NX + oUPHI] = ic.c_u; r[NX + oDDPHI] = -1;
  }
  // --- kinematic rows via AD over (q, dq) in R^14
  typedef D2<14> A;
  // map of the 14 AD inputs (q_n, dq_n) to zeta: list of (index, coefficient)
  struct Nz { int idx[5]; double cf[5]; int n; };
  Nz Tn[14], Tk[14];
  for (int j = 0; j < 7; j++) {
    Tn[j] = {{oQ + j, oDQ + j, oDDQ + j, oU + j, NX + oU + j}, {1, ic.a_dq, ic.a_ddq, ic.a_um, ic.a_u}, 5};
    Tn[7 + j] = {{oDQ + j, oDDQ + j, oU + j, NX + oU + j, 0}, {1, ic.b_ddq, ic.b_um, ic.b_u, 0}, 4};
    Tk[j] = {{oQ + j, 0, 0, 0, 0}, {1, 0, 0, 0, 0}, 1};
    Tk[7 + j] = {{oDQ + j, 0, 0, 0, 0}, {1, 0, 0, 0, 0}, 1};
  }
  for (int pass = 0; pass < 2; pass++) {
    A q[7], dq[7], pos[3], vl[3], va[3];
    const Nz* T = pass == 0 ? Tn : Tk;
    for (int j = 0; j < 7; j++) {
      double qv, dqv;
      if (pass == 0) {
        qv = wp[oQ + j] + ic.a_dq * wp[oDQ + j] + ic.a_ddq * wp[oDDQ + j] + ic.a_um * wp[oU + j] + ic.a_u * w[oU + j];
        dqv = wp[oDQ + j] + ic.b_ddq * wp[oDDQ + j] + ic.b_um * wp[oU + j] + ic.b_u * w[oU + j];
      } else { qv = wp[oQ + j]; dqv = wp[oDQ + j]; }
      q[j] = A::var(qv, j); dq[j] = A::var(dqv, 7 + j);
    }
    kinematics<A>(q, dq, pos, vl, va);
    // outputs, their rows and weights
    const A* outs[12]; int rows[12]; double scale[12]; int no = 0;
    if (pass == 0) {
      for (int i = 0; i < 3; i++) { outs[no] = &pos[i]; rows[no] = 21 + i; scale[no] = 1; no++; }
      for (int i = 0; i < 3; i++) { outs[no] = &va[i]; rows[no] = 24 + i; scale[no] = 0.5 * ic.h; no++; }
      for (int i = 0; i < 3; i++) { outs[no] = &vl[i]; rows[no] = 27 + i; scale[no] = 1; no++; }
      for (int i = 0; i < 3; i++) { outs[no] = &va[i]; rows[no] = 30 + i; scale[no] = 1; no++; }
    } else {
      for (int i = 0; i < 3; i++) { outs[no] = &va[i]; rows[no] = 24 + i; scale[no] = 0.5 * ic.h; no++; }
    }
    double Hs[14][14];
    memset(Hs, 0, sizeof(Hs));
    for (int o = 0; o < no; o++) {
      const A& f = *outs[o];
      for (int a = 0; a < 14; a++)
        for (int t = 0; t < T[a].n; t++) D.Jg[rows[o]][T[a].idx[t]] += scale[o] * f.g[a] * T[a].cf[t];
      double l = lam[rows[o]] * scale[o];
      for (int a = 0; a < 14; a++)
        for (int b = 0; b < 14; b++) Hs[a][b] += l * f.hess(a, b);
    }
    for (int a = 0; a < 14; a++)
      for (int b = 0; b < 14; b++) {
        if (Hs[a][b] == 0.0) continue;
        for (int t = 0; t < T[a].n; t++)
          for (int u = 0; u < T[b].n; u++) D.HL[T[a].idx[t]][T[b].idx[u]] += Hs[a][b] * T[a].cf[t] * T[b].cf[u];
      }
  }
  for (int i = 0; i < 3; i++) {
    D.Jg[21 + i][NX + oPPOS + i] -= 1; D.Jg[24 + i][NX + oPROT + i] -= 1; D.Jg[24 + i][oPROT + i] += 1;
    D.Jg[27 + i][NX + oVLIN + i] -= 1; D.Jg[30 + i][NX + oVANG + i] -= 1;
  }
  // --- path cost + inequalities via AD over c in R^21
  typedef D2<NC> C;
  int cmap[NC];
  for (int i = 0; i < 3; i++) { cmap[cPPOS + i] = NX + oPPOS + i; cmap[cPROT + i] = NX + oPROT + i; }
  for (int i = 0; i < 6; i++) { cmap[cV + i] = NX + oVLIN + i; cmap[cVPREV + i] = oVLIN + i; }
  cmap[cPHI] = NX + oPHI; cmap[cDPHI] = NX + oDPHI; cmap[cDDPHI] = NX + oDDPHI;
  C c[NC], cost, ineq[7], din[ND];
  for (int i = 0; i < 3; i++) { c[cPPOS + i] = C::var(w[oPPOS + i], cPPOS + i); c[cPROT + i] = C::var(w[oPROT + i], cPROT + i); }
  for (int i = 0; i < 6; i++) { c[cV + i] = C::var(w[oVLIN + i], cV + i); c[cVPREV + i] = C::var(wp[oVLIN + i], cVPREV + i); }
  c[cPHI] = C::var(w[oPHI], cPHI); c[cDPHI] = C::var(w[oDPHI], cDPHI); c[cDDPHI] = C::var(w[oDDPHI], cDDPHI);
  path_terms<C>(P.L, p, P.dt, c, cost, ineq, din);
  for (int a = 0; a < NC; a++) {
    D.gradf[cmap[a]] += cost.g[a];
    for (int b = 0; b < NC; b++) D.HL[cmap[a]][cmap[b]] += cost.hess(a, b);
  }
  for (int r = 0; r < 7; r++) {
    for (int a = 0; a < NC; a++) {
      D.Jg[36 + r][cmap[a]] += ineq[r].g[a];
      if (!interval && lam[36 + r] != 0.0)
        for (int b = 0; b < NC; b++) D.HL[cmap[a]][cmap[b]] += lam[36 + r] * ineq[r].hess(a, b);
    }
  }
  for (int r = 0; r < ND; r++) {
    for (int a = 0; a < NC; a++) {
      D.Jdd[r][cmap[a]] += din[r].g[a];
      if (interval && lam[NE + r] != 0.0)
        for (int b = 0; b < NC; b++) D.HL[cmap[a]][cmap[b]] += lam[NE + r] * din[r].hess(a, b);
    }
  }
  // --- separable quadratic cost
  const double* wt = p + P.L.w;
  for (int j = 0; j < 7; j++) {
    D.gradf[NX + oQ + j] += 2 * wt[10] * (w[oQ + j] - p[P.L.qd + j]); D.HL[NX + oQ + j][NX + oQ + j] += 2 * wt[10];
    D.gradf[NX + oDQ + j] += 2 * wt[11] * w[oDQ + j]; D.HL[NX + oDQ + j][NX + oDQ + j] += 2 * wt[11];
    D.gradf[NX + oDDQ + j] += 2 * wt[12] * w[oDDQ + j]; D.HL[NX + oDDQ + j][NX + oDDQ + j] += 2 * wt[12];
    D.gradf[NX + oU + j] += 2 * wt[13] * w[oU + j]; D.HL[NX + oU + j][NX + oU + j] += 2 * wt[13];
  }
  D.gradf[NX + oUPHI] += 2 * wt[9] * w[oUPHI]; D.HL[NX + oUPHI][NX + oUPHI] += 2 * wt[9];
}

// ------------------------------------------------------------------ whole-horizon evaluation
struct Eval {
  int N;
  double f;
  std::vector<double> g;       // [43N]
  std::vector<double> gradf;   // [44N]
  std::vector<double> Wd;      // [N][44][44]   diagonal Hessian blocks (Lagrangian)
  std::vector<double> Wo;      // [N][44][44]   block (k, k-1): rows w_k, cols w_{k-1}
  std::vector<double> A;       // [N][36][44]   d c_k / d w_{k-1}
  std::vector<double> B;       // [N][36][8]    d c_k / d u_k
  std::vector<double> Jd;      // [N][12][44]   d d_k / d w_k (interval rows)
  std::vector<double> d;       // [N][12]       interval rows
  std::vector<double> Jq;      // [N][7][44]    Jacobian of the reference-form rows 36..42
  explicit Eval(int N_) : N(N_), g(NG * N_), gradf(NX * N_), Wd(N_ * NX * NX), Wo(N_ * NX * NX),
                          A(N_ * NE * NX), B(N_ * NE * 8), Jd(N_ * ND * NX), d(N_ * ND), Jq(N_ * NI * NX) {}
};

void eval_values(const Prob& P, const double* x, const double* p, double& f, double* g, double* d = nullptr) {
  double wp0[NX];
  wprev0(P, p, wp0);
  f = 0;
  for (int k = 0; k < P.N; k++) {
    double c;
    stage_values(P, p, k == 0 ? wp0 : x + NX * (k - 1), x + NX * k, g + NG * k, c, d ? d + ND * k : nullptr);
    f += c;
  }
}

// lam: per stage NG entries (interval == false) or NE + ND entries (interval == true)
void eval_full(const Prob& P, const double* x, const double* p, const double* lam, Eval& E, bool interval) {
  double wp0[NX];
  wprev0(P, p, wp0);
  std::fill(E.gradf.begin(), E.gradf.end(), 0.0);
  std::fill(E.Wd.begin(), E.Wd.end(), 0.0);
  std::fill(E.Wo.begin(), E.Wo.end(), 0.0);
  E.f = 0;
  static thread_local StageD* D = nullptr;
  if (!D) D = new StageD;
  for (int k = 0; k < P.N; k++) {
    stage_derivs(P, p, k == 0 ? wp0 : x + NX * (k - 1), x + NX * k, lam + (interval ? NE + ND : NG) * k, *D, interval);
    E.f += D->cost;
    for (int i = 0; i < NG; i++) E.g[NG * k + i] = D->g[i];
    for (int i = 0; i < NX; i++) {
      E.gradf[NX * k + i] += D->gradf[NX + i];
      if (k > 0) E.gradf[NX * (k - 1) + i] += D->gradf[i];
    }
    double* Wd = &E.Wd[k * NX * NX];
    double* Wo = &E.Wo[k * NX * NX];
    for (int i = 0; i < NX; i++)
      for (int j = 0; j < NX; j++) {
        Wd[i * NX + j] += D->HL[NX + i][NX + j];
        if (k > 0) {
          E.Wd[(k - 1) * NX * NX + i * NX + j] += D->HL[i][j];
          Wo[i * NX + j] += D->HL[NX + i][j];
        }
      }
    for (int i = 0; i < NE; i++) {
      for (int j = 0; j < NX; j++) E.A[(k * NE + i) * NX + j] = D->Jg[i][j];
      for (int j = 0; j < 8; j++) E.B[(k * NE + i) * 8 + j] = D->Jg[i][NX + j];
    }
    for (int i = 0; i < NI; i++)
      for (int j = 0; j < NX; j++) E.Jq[(k * NI + i) * NX + j] = D->Jg[36 + i][NX + j];
    for (int i = 0; i < ND; i++) {
      E.d[ND * k + i] = D->d[i];
      for (int j = 0; j < NX; j++) E.Jd[(k * ND + i) * NX + j] = D->Jdd[i][NX + j];
    }
  }
}

// ------------------------------------------------------------------ interior-point method
struct Opts {
  double tol = 1e-8;
  int max_iter = 500;
  double mu_init = 1e-3;      // level of Ipopt's warm_start_mult_bound_push (see bmpc_host.h make_config)
  double bound_push = 1e-3;   // Ipopt warm_start_bound_push / warm_start_slack_bound_push
  double kappa_eps = 1000, kappa_mu = 0.1, theta_mu = 2.0, tau_min = 0.99, s_max = 100;   // fast monotone schedule (bmpc_host.h make_config)
  double gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8, s_phi = 2.3, s_theta = 1.1, delta_sw = 1.0;
  int verbose = 0;
  double diverge_tol = 1e7;   // dual infeasibility beyond which the multipliers are taken to diverge (locally infeasible instance)
  int mu_strategy = 3;        // 3 (default, = the CUDA path): monotone decrease + kkt-error progress test with re-centring; 0: monotone only
  int red_iters = 3;          // progress test: no improvement on any of the last red_iters accepted iterates ...
  double boost_fac = 10, boost_cap = 1.0;   // ... -> mu <- min(cap, fac * mu)
  int max_soc = 1;            // second-order corrections per iteration
  int soc_budget = 2;         // ... until this many corrections in a row have been rejected
  double rollout_thr = 0.5;   // start whose equality rows are violated by more than this (a cold start in the middle of a path): states
                              // replaced by the roll-out of the start's inputs (0: never)
  int boost_budget = 6;       // re-centrings per solve before it is stopped as locally infeasible
  int stall_stop = 3;         // third time without progress with mu at its cap: stop as locally infeasible
};

struct Ipm {
  const Prob& P;
  const double* p;
  Opts o;
  int N, n, ne, ni;
  std::vector<double> x, s, y, zs, zL, zU;           // primal, slacks, eq multipliers, ineq/bound multipliers
  std::vector<double> lam;                           // [43N] packed (y, zs) per stage for eval
  std::vector<double> dx, ds, ynew, dzs, dzL, dzU;
  std::vector<double> lbx, ubx;
  Eval E;
  double mu, delta_w_last = 0;
  const double* soc_c = nullptr; const double* soc_d = nullptr;   // second-order correction: residuals replacing c [NE N] and d + s [ND N]
  double* trace = nullptr; int trace_cap = 0;   // optional per-iteration log: e0, mu, alpha_pr_max, alpha, theta, dw
  std::vector<std::pair<double, double>> filter;
  // Riccati storage
  std::vector<double> Wt, gh, Kk, kk, Pn, pn;
  Ipm(const Prob& P_, const double* p_, const Opts& o_)
      : P(P_), p(p_), o(o_), N(P_.N), n(NX * P_.N), ne(NE * P_.N), ni(ND * P_.N), x(n), s(ni), y(ne), zs(ni),
        zL(n), zU(n), lam((NE + ND) * P_.N), dx(n), ds(ni), ynew(ne), dzs(ni), dzL(n), dzU(n), lbx(n), ubx(n), E(P_.N),
        Wt(N * NX * NX), gh(n), Kk(N * 8 * NX), kk(N * 8), Pn(NX * NX), pn(NX) {
    for (int k = 0; k < N; k++)
      for (int i = 0; i < NX; i++) { lbx[NX * k + i] = P.lb[i]; ubx[NX * k + i] = P.ub[i]; }
  }
  void pack_lam() {
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < NE; i++) lam[(NE + ND) * k + i] = y[NE * k + i];
      for (int i = 0; i < ND; i++) lam[(NE + ND) * k + NE + i] = zs[ND * k + i];
    }
  }
  const double* c_(int k) const { return &E.g[NG * k]; }
  const double* d_(int k) const { return &E.d[ND * k]; }

  // barrier objective and constraint violation at (x, s) given values (f, g)
  double barrier_phi(double f, const std::vector<double>& xx, const std::vector<double>& ss) const {
    double v = f;
    for (int i = 0; i < ni; i++) v -= mu * std::log(ss[i]);
    for (int i = 0; i < n; i++) {
      if (std::isfinite(lbx[i])) v -= mu * std::log(xx[i] - lbx[i]);
      if (std::isfinite(ubx[i])) v -= mu * std::log(ubx[i] - xx[i]);
    }
    return v;
  }
  double theta(const double* g, const double* d, const std::vector<double>& ss) const {
    double t = 0;
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < NE; i++) t += std::fabs(g[NG * k + i]);
      for (int i = 0; i < ND; i++) t += std::fabs(d[ND * k + i] + ss[ND * k + i]);
    }
    return t;
  }
  // dual residual  grad f + Jc^T y + Jd^T zs - zL + zU
  void dual_residual(const std::vector<double>& yy, const std::vector<double>& zz, std::vector<double>& r) const {
    r.assign(n, 0.0);
    for (int i = 0; i < n; i++) r[i] = E.gradf[i] - zL[i] + zU[i];
    for (int k = 0; k < N; k++) {
      const double* A = &E.A[k * NE * NX];
      const double* B = &E.B[k * NE * 8];
      const double* Jd = &E.Jd[k * ND * NX];
      for (int i = 0; i < NE; i++) {
        double yi = yy[NE * k + i];
        if (k > 0) for (int j = 0; j < NX; j++) r[NX * (k - 1) + j] += A[i * NX + j] * yi;
        for (int j = 0; j < 8; j++) r[NX * k + j] += B[i * 8 + j] * yi;
        r[NX * k + 8 + i] -= yi;
      }
      for (int i = 0; i < ND; i++)
        for (int j = 0; j < NX; j++) r[NX * k + j] += Jd[i * NX + j] * zz[ND * k + i];
    }
  }
  // Ipopt error measure E_mu (scaled)
  double kkt_error(double muv, double* parts = nullptr) const {
    std::vector<double> r;
    dual_residual(y, zs, r);
    double dinf = 0, pinf = 0, cinf = 0, ysum = 0, zsum = 0;
    int nz = 0;
    for (int i = 0; i < n; i++) dinf = std::max(dinf, std::fabs(r[i]));
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < NE; i++) { pinf = std::max(pinf, std::fabs(E.g[NG * k + i])); ysum += std::fabs(y[NE * k + i]); }
      for (int i = 0; i < ND; i++) {
        pinf = std::max(pinf, std::fabs(E.d[ND * k + i] + s[ND * k + i]));
        cinf = std::max(cinf, std::fabs(s[ND * k + i] * zs[ND * k + i] - muv));
        zsum += std::fabs(zs[ND * k + i]); nz++;
      }
    }
    for (int i = 0; i < n; i++) {
      if (std::isfinite(lbx[i])) { cinf = std::max(cinf, std::fabs((x[i] - lbx[i]) * zL[i] - muv)); zsum += zL[i]; nz++; }
      if (std::isfinite(ubx[i])) { cinf = std::max(cinf, std::fabs((ubx[i] - x[i]) * zU[i] - muv)); zsum += zU[i]; nz++; }
    }
    double sd = std::max(o.s_max, (ysum + zsum) / (ne + nz)) / o.s_max;
    double sc = std::max(o.s_max, zsum / nz) / o.s_max;
    if (parts) { parts[0] = dinf; parts[1] = pinf; parts[2] = cinf; }
    return std::max(dinf / sd, std::max(pinf, cinf / sc));
  }

  // Solve the condensed KKT system by Riccati; returns false if some Q_uu is not PD.
  // right-hand side: gh = ga * (mu-free part) + gc * (coefficient of mu), equality residual scaled by cs
  // (full step: ga = 1, gc = mu, cs = 1; affine-scaling step: 1, 0, 1; centring step: 0, 1, 0)
  bool riccati(double delta_w, double ga = 1.0, double gc = -1.0, double cs = 1.0) {
    if (gc < 0) gc = mu;
    std::vector<double> csc(NG * N);
    for (int i = 0; i < NG * N; i++) csc[i] = cs * E.g[i];
    if (soc_c) for (int k = 0; k < N; k++) for (int i = 0; i < NE; i++) csc[NG * k + i] = cs * soc_c[NE * k + i];
    auto c_ = [&](int k) { return &csc[NG * k]; };
    // W~ and g^ -------------------------------------------------------
    for (int k = 0; k < N; k++) {
      double* W = &Wt[k * NX * NX];
      memcpy(W, &E.Wd[k * NX * NX], sizeof(double) * NX * NX);
      const double* Jd = &E.Jd[k * ND * NX];
      for (int i = 0; i < NX; i++) {
        int gi = NX * k + i;
        double sig = delta_w, gb = ga * E.gradf[gi];
        if (std::isfinite(lbx[gi])) { double sl = x[gi] - lbx[gi]; sig += zL[gi] / sl; gb -= gc / sl; }
        if (std::isfinite(ubx[gi])) { double su = ubx[gi] - x[gi]; sig += zU[gi] / su; gb += gc / su; }
        W[i * NX + i] += sig;
        gh[gi] = gb;
      }
      for (int r = 0; r < ND; r++) {
        double sv = s[ND * k + r], zv = zs[ND * k + r];
        double Sig = zv / sv, rd = soc_d ? soc_d[ND * k + r] : d_(k)[r] + sv;
        double coef = gc / sv + ga * Sig * rd;
        for (int i = 0; i < NX; i++) {
          double ji = Jd[r * NX + i];
          if (ji == 0.0) continue;
          gh[NX * k + i] += ji * coef;
          for (int j = 0; j < NX; j++) W[i * NX + j] += Sig * ji * Jd[r * NX + j];
        }
      }
    }
    // backward ---------------------------------------------------------
    std::fill(Pn.begin(), Pn.end(), 0.0);
    std::fill(pn.begin(), pn.end(), 0.0);
    std::vector<double> M(NX * NX), mv(NX), MxxA(NE * NX), MxxB(NE * 8), Qus(8 * NX), Qss(NX * NX), Quu(64), t(NE), qu(8), qs(NX), Lc(64);
    for (int k = N - 1; k >= 0; k--) {
      const double* W = &Wt[k * NX * NX];
      const double* A = &E.A[k * NE * NX];
      const double* B = &E.B[k * NE * 8];
      const double* O = &E.Wo[k * NX * NX];
      const double* c = c_(k);
      for (int i = 0; i < NX * NX; i++) M[i] = W[i] + Pn[i];
      for (int i = 0; i < NX; i++) mv[i] = gh[NX * k + i] + pn[i];
      auto Mxx = [&](int i, int j) { return M[(8 + i) * NX + 8 + j]; };
      auto Mux = [&](int i, int j) { return M[i * NX + 8 + j]; };
      for (int i = 0; i < NE; i++) {
        for (int j = 0; j < 8; j++) { double a = 0; for (int l = 0; l < NE; l++) a += Mxx(i, l) * B[l * 8 + j]; MxxB[i * 8 + j] = a; }
        double a = mv[8 + i];
        for (int l = 0; l < NE; l++) a += Mxx(i, l) * c[l];
        t[i] = a;
      }
      for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) {
          double a = M[i * NX + j];
          for (int l = 0; l < NE; l++) a += Mux(i, l) * B[l * 8 + j] + B[l * 8 + i] * Mux(j, l) + B[l * 8 + i] * MxxB[l * 8 + j];
          Quu[i * 8 + j] = a;
        }
        double a = mv[i];
        for (int l = 0; l < NE; l++) a += Mux(i, l) * c[l] + B[l * 8 + i] * t[l];
        qu[i] = a;
      }
      // Cholesky of Quu
      bool ok = true;
      for (int i = 0; i < 8 && ok; i++)
        for (int j = 0; j <= i; j++) {
          double a = 0.5 * (Quu[i * 8 + j] + Quu[j * 8 + i]);
          for (int l = 0; l < j; l++) a -= Lc[i * 8 + l] * Lc[j * 8 + l];
          if (i == j) { if (!(a > 1e-14)) { ok = false; break; } Lc[i * 8 + i] = std::sqrt(a); }
          else Lc[i * 8 + j] = a / Lc[j * 8 + j];
        }
      if (!ok) return false;
      auto chol_solve = [&](double* b) {  // in place, 8-vector
        for (int i = 0; i < 8; i++) { double a = b[i]; for (int l = 0; l < i; l++) a -= Lc[i * 8 + l] * b[l]; b[i] = a / Lc[i * 8 + i]; }
        for (int i = 7; i >= 0; i--) { double a = b[i]; for (int l = i + 1; l < 8; l++) a -= Lc[l * 8 + i] * b[l]; b[i] = a / Lc[i * 8 + i]; }
      };
      double* kap = &kk[k * 8];
      for (int i = 0; i < 8; i++) kap[i] = -qu[i];
      chol_solve(kap);
      if (k == 0) break;
      for (int i = 0; i < NE; i++)
        for (int j = 0; j < NX; j++) { double a = 0; for (int l = 0; l < NE; l++) a += Mxx(i, l) * A[l * NX + j]; MxxA[i * NX + j] = a; }
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < NX; j++) {
          double a = O[i * NX + j];
          for (int l = 0; l < NE; l++) a += Mux(i, l) * A[l * NX + j] + B[l * 8 + i] * (MxxA[l * NX + j] + O[(8 + l) * NX + j]);
          Qus[i * NX + j] = a;
        }
      for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) {
          double a = 0;
          for (int l = 0; l < NE; l++) a += A[l * NX + i] * (MxxA[l * NX + j] + O[(8 + l) * NX + j]) + O[(8 + l) * NX + i] * A[l * NX + j];
          Qss[i * NX + j] = a;
        }
        double a = 0;
        for (int l = 0; l < NE; l++) a += A[l * NX + i] * t[l] + O[(8 + l) * NX + i] * c[l];
        qs[i] = a;
      }
      double* K = &Kk[k * 8 * NX];
      for (int j = 0; j < NX; j++) {
        double col[8];
        for (int i = 0; i < 8; i++) col[i] = -Qus[i * NX + j];
        chol_solve(col);
        for (int i = 0; i < 8; i++) K[i * NX + j] = col[i];
      }
      for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) { double a = Qss[i * NX + j]; for (int l = 0; l < 8; l++) a += Qus[l * NX + i] * K[l * NX + j]; Pn[i * NX + j] = a; }
        double a = qs[i];
        for (int l = 0; l < 8; l++) a += Qus[l * NX + i] * kap[l];
        pn[i] = a;
      }
      for (int i = 0; i < NX; i++)
        for (int j = i + 1; j < NX; j++) { double a = 0.5 * (Pn[i * NX + j] + Pn[j * NX + i]); Pn[i * NX + j] = Pn[j * NX + i] = a; }
    }
    // forward ----------------------------------------------------------
    for (int k = 0; k < N; k++) {
      const double* A = &E.A[k * NE * NX];
      const double* B = &E.B[k * NE * 8];
      const double* c = c_(k);
      double* dw = &dx[NX * k];
      const double* dsv = k > 0 ? &dx[NX * (k - 1)] : nullptr;
      for (int i = 0; i < 8; i++) {
        double a = kk[k * 8 + i];
        if (k > 0) for (int j = 0; j < NX; j++) a += Kk[(k * 8 + i) * NX + j] * dsv[j];
        dw[i] = a;
      }
      for (int i = 0; i < NE; i++) {
        double a = c[i];
        if (k > 0) for (int j = 0; j < NX; j++) a += A[i * NX + j] * dsv[j];
        for (int j = 0; j < 8; j++) a += B[i * 8 + j] * dw[j];
        dw[8 + i] = a;
      }
    }
    // equality multipliers by the adjoint recursion (independent of the Riccati internals)
    for (int k = N - 1; k >= 0; k--) {
      const double* W = &Wt[k * NX * NX];
      for (int i = 0; i < NE; i++) {
        int r = 8 + i;
        double a = gh[NX * k + r];
        for (int j = 0; j < NX; j++) a += W[r * NX + j] * dx[NX * k + j];
        if (k > 0) for (int j = 0; j < NX; j++) a += E.Wo[k * NX * NX + r * NX + j] * dx[NX * (k - 1) + j];
        if (k < N - 1) {
          const double* On = &E.Wo[(k + 1) * NX * NX];
          const double* An = &E.A[(k + 1) * NE * NX];
          for (int j = 0; j < NX; j++) a += On[j * NX + r] * dx[NX * (k + 1) + j];
          for (int j = 0; j < NE; j++) a += An[j * NX + r] * ynew[NE * (k + 1) + j];
        }
        ynew[NE * k + i] = a;
      }
    }
    return true;
  }

  // residual of the condensed stationarity rows for the u-blocks (diagnostic)
  double lin_residual() const {
    double worst = 0;
    for (int k = 0; k < N; k++) {
      const double* W = &Wt[k * NX * NX];
      for (int r = 0; r < 8; r++) {
        double a = gh[NX * k + r];
        for (int j = 0; j < NX; j++) a += W[r * NX + j] * dx[NX * k + j];
        if (k > 0) for (int j = 0; j < NX; j++) a += E.Wo[k * NX * NX + r * NX + j] * dx[NX * (k - 1) + j];
        for (int j = 0; j < NE; j++) a += E.B[(k * NE + j) * 8 + r] * ynew[NE * k + j];
        if (k < N - 1) {
          const double* On = &E.Wo[(k + 1) * NX * NX];
          const double* An = &E.A[(k + 1) * NE * NX];
          for (int j = 0; j < NX; j++) a += On[j * NX + r] * dx[NX * (k + 1) + j];
          for (int j = 0; j < NE; j++) a += An[j * NX + r] * ynew[NE * (k + 1) + j];
        }
        worst = std::max(worst, std::fabs(a));
      }
    }
    return worst;
  }

  int solve(const double* x0, int& iters, double& kkt_final) {
    // initial point: push into the bounds (Ipopt warm-start push), slacks from d(x0)
    for (int i = 0; i < n; i++) {
      double v = x0[i], l = lbx[i], u = ubx[i];
      if (std::isfinite(l) && std::isfinite(u)) {
        double pl = std::min(o.bound_push * std::max(1.0, std::fabs(l)), o.bound_push * (u - l));
        double pu = std::min(o.bound_push * std::max(1.0, std::fabs(u)), o.bound_push * (u - l));
        v = std::min(std::max(v, l + pl), u - pu);
      } else if (std::isfinite(l)) v = std::max(v, l + o.bound_push * std::max(1.0, std::fabs(l)));
      else if (std::isfinite(u)) v = std::min(v, u - o.bound_push * std::max(1.0, std::fabs(u)));
      x[i] = v;
    }
    if (o.rollout_thr > 0) {
      // Cold-start repair.  The reference's cold start (BoundMPC.py:316-321: zeros, q0, p0) used away from the start of the
      // path has its path parameter 0.75 m from where the robot is; Ipopt recovers from such starts in its restoration phase,
      // this iteration does not have one.  A start whose equality rows are violated by more than rollout_thr gets its state
      // entries replaced by the roll-out of its own inputs from the initial state (the dynamics rows are explicit:
      // x_k = F(w_{k-1}, u_k)), kept inside the variable bounds.  Warm starts and the step-0 cold start are not touched.
      double wp0[NX]; wprev0(P, p, wp0);
      double viol = 0;
      { double f; std::vector<double> g(NG * N); eval_values(P, x.data(), p, f, g.data()); for (int k = 0; k < N; k++) for (int i = 0; i < NE; i++) viol = std::max(viol, std::fabs(g[NG * k + i])); }
      if (viol > o.rollout_thr) {
        for (int k = 0; k < N; k++) {
          double g[NG], c;
          stage_values(P, p, k == 0 ? wp0 : &x[NX * (k - 1)], &x[NX * k], g, c);
          for (int i = 0; i < NE; i++) {
            double v = x[NX * k + 8 + i] + g[i];
            const double l = lbx[NX * k + 8 + i], u = ubx[NX * k + 8 + i];
            if (std::isfinite(l)) v = std::max(v, l + o.bound_push);
            if (std::isfinite(u)) v = std::min(v, u - o.bound_push);
            x[NX * k + 8 + i] = v;
          }
        }
      }
    }
    mu = o.mu_init;
    std::fill(y.begin(), y.end(), 0.0);
    {
      double f;
      std::vector<double> g(NG * N), d(ND * N);
      eval_values(P, x.data(), p, f, g.data(), d.data());
      for (int i = 0; i < ni; i++) s[i] = std::max(-d[i], o.bound_push);
    }
    for (int i = 0; i < ni; i++) zs[i] = mu / s[i];
    for (int i = 0; i < n; i++) {
      zL[i] = std::isfinite(lbx[i]) ? mu / (x[i] - lbx[i]) : 0.0;
      zU[i] = std::isfinite(ubx[i]) ? mu / (ubx[i] - x[i]) : 0.0;
    }
    filter.clear();
    double theta0 = -1, theta_max = 0, theta_min = 0;
    std::vector<double> xt(n), st(ni), gt(NG * N), dt_(ni);
    int status = 1;  // 0 = success, 1 = max_iter, 2 = line-search failure, 3 = regularisation failure, 5 = diverging multipliers
    int it = 0, ls_fail = 0;
    int n_soc = 0, stalls = 0, soc_fails = 0, boosts = 0;
    std::vector<double> refs;
    for (;; it++) {
      pack_lam();
      eval_full(P, x.data(), p, lam.data(), E, true);
      double parts[3];
      double e0 = kkt_error(0.0, parts);
      if (o.verbose) printf("it %3d f %.10g  E0 %.3e (d %.2e p %.2e c %.2e) mu %.2e\n", it, E.f, e0, parts[0], parts[1], parts[2], mu);
      kkt_final = e0;
      if (e0 <= o.tol) { status = 0; break; }
      if (it >= o.max_iter) { status = 1; break; }
      if (parts[0] > o.diverge_tol) { status = 5; break; }
      // barrier parameter
      {
        // monotone Fiacco-McCormick decrease (Waechter & Biegler 2006, eq. (7)), one level per iteration
        bool mu_changed = false;
        if (mu > o.tol / 10 && kkt_error(mu) <= o.kappa_eps * mu) {
          mu = std::max(o.tol / 10, std::min(o.kappa_mu * mu, std::pow(mu, o.theta_mu)));
          mu_changed = true;
        }
        if (mu_changed) { filter.clear(); refs.clear(); }
        if (o.mu_strategy == 3) {
          // progress test of Ipopt's adaptive strategy (adaptive_mu_globalization = kkt-error, the reference's setting,
          // BoundMPC.py:130-131; here with red_iters = 3): when the optimality error has not improved on any of the last
          // red_iters accepted iterates the iteration is crawling along the boundary (a slack pinned at zero by a grown
          // multiplier, every step cut by the fraction-to-the-boundary rule); re-centre at a larger barrier parameter
          double parts2[3];
          const double qf = kkt_error(0.0, parts2);
          bool suff = true;
          if ((int)refs.size() >= o.red_iters) { suff = false; for (double r : refs) if (qf <= 0.9999 * r) suff = true; }
          if (suff) { if ((int)refs.size() >= o.red_iters) refs.erase(refs.begin()); refs.push_back(qf); }
          else if (mu < o.boost_cap) {
            // (a solve that keeps cycling down and up the barrier ladder is not converging either: successful solves of the
            // bench / config-4 workloads re-centre at most four times)
            if (++boosts > o.boost_budget) { status = 5; break; }
            mu = std::min(o.boost_cap, o.boost_fac * mu);
            filter.clear(); refs.clear();
            if (o.verbose) printf("      no progress: mu -> %.3e\n", mu);
          } else if (o.stall_stop && ++stalls >= o.stall_stop) {
            // re-centring exhausted (mu at its cap) and still no progress: the multipliers of rows that cannot be satisfied
            // keep growing -- the instance is locally infeasible (Ipopt would end its restoration phase at a stationary point
            // of the constraint violation: "Converged to a point of local infeasibility")
            status = 5; break;
          } else refs.clear();
        }
      }
      double th_cur = theta(E.g.data(), E.d.data(), s);
      if (theta0 < 0) { theta0 = th_cur; theta_max = 1e4 * std::max(1.0, theta0); theta_min = 1e-4 * std::max(1.0, theta0); }
      // search direction with inertia correction
      double dw = 0;
      auto factor_solve = [&](double ga, double gc, double cs) -> bool {
        bool ok = riccati(dw, ga, gc, cs);
        if (!ok) {
          dw = delta_w_last == 0 ? 1e-4 : std::max(1e-20, delta_w_last / 3);
          for (int tries = 0; tries < 60; tries++) {
            ok = riccati(dw, ga, gc, cs);
            if (ok) break;
            dw *= (delta_w_last == 0 ? 100 : 8);
            if (dw > 1e40) break;
          }
          if (ok) delta_w_last = dw;
        }
        return ok;
      };
      auto rest_of_step = [&](double ga, double gc) {   // ds, dzs, dzL, dzU from dx for the right-hand side (ga, gc)
        for (int k = 0; k < N; k++) {
          const double* Jd = &E.Jd[k * ND * NX];
          for (int r = 0; r < ND; r++) {
            double jd = 0;
            for (int j = 0; j < NX; j++) jd += Jd[r * NX + j] * dx[NX * k + j];
            int i = ND * k + r;
            ds[i] = -ga * (soc_d ? soc_d[i] : d_(k)[r] + s[i]) - jd;
            dzs[i] = gc / s[i] - ga * zs[i] - zs[i] / s[i] * ds[i];
          }
        }
        for (int i = 0; i < n; i++) {
          dzL[i] = dzU[i] = 0;
          if (std::isfinite(lbx[i])) { double sl = x[i] - lbx[i]; dzL[i] = gc / sl - ga * zL[i] - zL[i] / sl * dx[i]; }
          if (std::isfinite(ubx[i])) { double su = ubx[i] - x[i]; dzU[i] = gc / su - ga * zU[i] + zU[i] / su * dx[i]; }
        }
      };
      if (!factor_solve(1, mu, 1)) { status = 3; break; }
      rest_of_step(1, mu);
      if (o.verbose > 1) printf("      delta_w %.1e lin_res %.2e\n", dw, lin_residual());
      // fraction to the boundary
      double tau = std::max(o.tau_min, 1 - mu), apr = 1, adu = 1;
      for (int i = 0; i < ni; i++) {
        if (ds[i] < 0) apr = std::min(apr, -tau * s[i] / ds[i]);
        if (dzs[i] < 0) adu = std::min(adu, -tau * zs[i] / dzs[i]);
      }
      for (int i = 0; i < n; i++) {
        if (std::isfinite(lbx[i])) {
          if (dx[i] < 0) apr = std::min(apr, -tau * (x[i] - lbx[i]) / dx[i]);
          if (dzL[i] < 0) adu = std::min(adu, -tau * zL[i] / dzL[i]);
        }
        if (std::isfinite(ubx[i])) {
          if (dx[i] > 0) apr = std::min(apr, tau * (ubx[i] - x[i]) / dx[i]);
          if (dzU[i] < 0) adu = std::min(adu, -tau * zU[i] / dzU[i]);
        }
      }
      // filter line search
      double phi_cur = barrier_phi(E.f, x, s);
      double dphi = 0;
      for (int i = 0; i < n; i++) {
        dphi += E.gradf[i] * dx[i];
        if (std::isfinite(lbx[i])) dphi -= mu * dx[i] / (x[i] - lbx[i]);
        if (std::isfinite(ubx[i])) dphi += mu * dx[i] / (ubx[i] - x[i]);
      }
      for (int i = 0; i < ni; i++) dphi -= mu * ds[i] / s[i];
      double alpha = apr;
      bool accepted = false, ftype = false;
      double th_t = 0, ph_t = 0;
      // flat merit functions: with the constraint violation at rounding level and no measurable predicted change of
      // the barrier objective the decrease tests below only see noise (they then cut the step to ~1e-9 for dozens
      // of iterations while the dual infeasibility stays above tol); Newton's full step is taken instead, as Ipopt
      // does for its "tiny steps"
      const bool flat = th_cur <= 1e-10 && std::fabs(dphi) <= 1e-10 * std::max(1.0, std::fabs(phi_cur));
      // acceptance of a trial point (theta, phi) for a step of length a_sw along the search direction
      // (Waechter & Biegler 2006, Alg. A, steps A-5.3 / A-5.4), with Ipopt's round-off allowance (Compare_le: lhs - rhs <=
      // 10 eps |reference value|): close to the solution the decrease conditions are decided by rounding noise and a
      // strict test sends the iteration into dozens of useless backtracking steps
      auto acceptable = [&](double th, double ph, double a_sw, bool& ft_out) -> bool {
        if (!std::isfinite(th) || !std::isfinite(ph) || th > theta_max) return false;
        if (flat) { ft_out = true; return true; }
        for (auto& fe : filter)
          if (!(th < fe.first || ph < fe.second)) return false;
        const bool sw = dphi < 0 && a_sw * std::pow(-dphi, o.s_phi) > o.delta_sw * std::pow(th_cur, o.s_theta);
        const double ro = 10 * 2.220446049250313e-16;
        if (th_cur <= theta_min && sw) {
          if (ph - phi_cur - o.eta_phi * a_sw * dphi <= ro * std::fabs(phi_cur)) { ft_out = true; return true; }
        } else {
          if (th - (1 - o.gamma_theta) * th_cur <= ro * std::fabs(th_cur) ||
              ph - phi_cur + o.gamma_phi * th_cur <= ro * std::fabs(phi_cur)) { ft_out = false; return true; }
        }
        return false;
      };
      for (int ls = 0; ls < 40; ls++, alpha *= 0.5) {
        for (int i = 0; i < n; i++) xt[i] = x[i] + alpha * dx[i];
        for (int i = 0; i < ni; i++) st[i] = s[i] + alpha * ds[i];
        double ft;
        eval_values(P, xt.data(), p, ft, gt.data(), dt_.data());
        th_t = theta(gt.data(), dt_.data(), st);
        ph_t = barrier_phi(ft, xt, st);
        if (acceptable(th_t, ph_t, alpha, ftype)) { accepted = true; break; }
        if (ls == 0 && o.max_soc > 0 && soc_fails < o.soc_budget && std::isfinite(th_t) && th_t >= th_cur) {
          // ---- second-order correction (Waechter & Biegler 2006, Sec. 2.4; Ipopt max_soc = 4, kappa_soc = 0.99): the
          // rejected full step has not reduced the constraint violation; re-solve with the residuals
          // c_soc = alpha c(x_k) + c(x_k + alpha dx) and the factorisation of this iteration
          std::vector<double> cs_c(ne), cs_d(ni), dx0 = dx, ds0 = ds, y0 = ynew, dzs0 = dzs, dzL0 = dzL, dzU0 = dzU;
          for (int k = 0; k < N; k++) {
            for (int i = 0; i < NE; i++) cs_c[NE * k + i] = alpha * E.g[NG * k + i] + gt[NG * k + i];
            for (int i = 0; i < ND; i++) cs_d[ND * k + i] = alpha * (E.d[ND * k + i] + s[ND * k + i]) + dt_[ND * k + i] + st[ND * k + i];
          }
          double th_old = th_t;
          bool soc_ok = false;
          for (int q = 0; q < o.max_soc; q++) {
            soc_c = cs_c.data(); soc_d = cs_d.data();
            const bool okf = riccati(dw, 1, mu, 1);
            g_soc_solves++;
            if (okf) rest_of_step(1, mu);
            soc_c = soc_d = nullptr;
            if (!okf) break;
            double a_soc = 1;
            for (int i = 0; i < ni; i++) if (ds[i] < 0) a_soc = std::min(a_soc, -tau * s[i] / ds[i]);
            for (int i = 0; i < n; i++) {
              if (std::isfinite(lbx[i]) && dx[i] < 0) a_soc = std::min(a_soc, -tau * (x[i] - lbx[i]) / dx[i]);
              if (std::isfinite(ubx[i]) && dx[i] > 0) a_soc = std::min(a_soc, tau * (ubx[i] - x[i]) / dx[i]);
            }
            for (int i = 0; i < n; i++) xt[i] = x[i] + a_soc * dx[i];
            for (int i = 0; i < ni; i++) st[i] = s[i] + a_soc * ds[i];
            eval_values(P, xt.data(), p, ft, gt.data(), dt_.data());
            const double th_s = theta(gt.data(), dt_.data(), st), ph_s = barrier_phi(ft, xt, st);
            if (o.verbose) printf("      soc %d: alpha %.3e theta %.3e -> %.3e (first trial %.3e)\n", q, a_soc, th_cur, th_s, th_t);
            if (acceptable(th_s, ph_s, alpha, ftype)) { soc_ok = true; alpha = a_soc; th_t = th_s; ph_t = ph_s; break; }
            if (!(th_s <= 0.99 * th_old)) break;
            th_old = th_s;
            for (int k = 0; k < N; k++) {
              for (int i = 0; i < NE; i++) cs_c[NE * k + i] = a_soc * cs_c[NE * k + i] + gt[NG * k + i];
              for (int i = 0; i < ND; i++) cs_d[ND * k + i] = a_soc * cs_d[ND * k + i] + dt_[ND * k + i] + st[ND * k + i];
            }
          }
          if (soc_ok) { accepted = true; n_soc++; g_soc_accepted++; soc_fails = 0; break; }
          soc_fails++;      // (soc_budget corrections rejected in a row: no more attempts in this solve)
          dx = dx0; ds = ds0; ynew = y0; dzs = dzs0; dzL = dzL0; dzU = dzU0;
        }
      }
      if (!accepted) {
        // no restoration phase: clear the filter and take the damped step that keeps the iterate interior
        filter.clear();
        alpha = apr * std::pow(0.5, 6);
        if (o.verbose) printf("      line search failed; damped step\n");
        if (++ls_fail > 8) { status = 2; break; }
      } else if (!ftype) {
        filter.push_back({(1 - o.gamma_theta) * th_cur, phi_cur - o.gamma_phi * th_cur});
      }
      if (o.verbose > 1) {
        double zmax = 0, smin = 1e300; int iz = 0, is = 0;
        for (int i = 0; i < ni; i++) { if (zs[i] > zmax) { zmax = zs[i]; iz = i; } if (s[i] < smin) { smin = s[i]; is = i; } }
        double dxm = 0; int idx = 0; for (int i = 0; i < n; i++) if (std::fabs(dx[i]) > dxm) { dxm = std::fabs(dx[i]); idx = i; }
        printf("      alpha_pr %.3e (max %.3e) alpha_du %.3e theta %.3e -> %.3e  phi %.6e -> %.6e dphi %.3e  zmax %.3e@%d,%d smin %.3e@%d,%d |dx| %.3e@%d,%d ftype %d filt %zu\n", alpha, apr, adu, th_cur, th_t, phi_cur, ph_t, dphi, zmax, iz / ND, iz % ND, smin, is / ND, is % ND, dxm, idx / 44, idx % 44, (int)ftype, filter.size());
      }
      if (trace && it < trace_cap) { double* T = trace + 6 * it; T[0] = e0; T[1] = mu; T[2] = apr; T[3] = alpha; T[4] = th_cur; T[5] = dw; }
      for (int i = 0; i < n; i++) x[i] += alpha * dx[i];
      for (int i = 0; i < ni; i++) s[i] += alpha * ds[i];
      for (int i = 0; i < ne; i++) y[i] += alpha * (ynew[i] - y[i]);
      const double ks = 1e10;
      for (int i = 0; i < ni; i++) {
        zs[i] += adu * dzs[i];
        zs[i] = std::max(std::min(zs[i], ks * mu / s[i]), mu / (ks * s[i]));
      }
      for (int i = 0; i < n; i++) {
        if (std::isfinite(lbx[i])) { double sl = x[i] - lbx[i]; zL[i] += adu * dzL[i]; zL[i] = std::max(std::min(zL[i], ks * mu / sl), mu / (ks * sl)); }
        if (std::isfinite(ubx[i])) { double su = ubx[i] - x[i]; zU[i] += adu * dzU[i]; zU[i] = std::max(std::min(zU[i], ks * mu / su), mu / (ks * su)); }
      }
    }
    iters = it;
    return status;
  }
};

}  // namespace

// ---------------------------------------------------------------------- C interface (ctypes)
extern "C" {

int orc_dims(int N, int S, int* n, int* m, int* np) {
  *n = NX * N; *m = NG * N; *np = Layout(S).np;
  return 0;
}

int orc_bounds(int N, int S, double dt, double* lbx, double* ubx, double* lbg, double* ubg) {
  Prob P(N, S, dt);
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < NX; i++) { lbx[NX * k + i] = P.lb[i]; ubx[NX * k + i] = P.ub[i]; }
    for (int i = 0; i < NG; i++) { lbg[NG * k + i] = i < NE ? 0.0 : -INFINITY; ubg[NG * k + i] = 0.0; }
  }
  return 0;
}

int orc_eval(int N, int S, double dt, const double* x, const double* p, double* f, double* g) {
  Prob P(N, S, dt);
  eval_values(P, x, p, *f, g);
  return 0;
}

// values with the interval-form inequality rows d [12 N] (the form the iteration works on)
int orc_eval_d(int N, int S, double dt, const double* x, const double* p, double* f, double* g, double* d) {
  Prob P(N, S, dt);
  eval_values(P, x, p, *f, g, d);
  return 0;
}

// dense gradient [n], Jacobian [m x n] row-major, Hessian of f + lam.g [n x n]
int orc_derivs(int N, int S, double dt, const double* x, const double* p, const double* lam,
               double* gradf, double* jac, double* hess) {
  Prob P(N, S, dt);
  Eval E(N);
  eval_full(P, x, p, lam, E, false);
  int n = P.n, m = P.m;
  for (int i = 0; i < n; i++) gradf[i] = E.gradf[i];
  if (jac) {
    memset(jac, 0, sizeof(double) * m * n);
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < NE; i++) {
        double* row = jac + (size_t)(NG * k + i) * n;
        if (k > 0) for (int j = 0; j < NX; j++) row[NX * (k - 1) + j] = E.A[(k * NE + i) * NX + j];
        for (int j = 0; j < 8; j++) row[NX * k + j] = E.B[(k * NE + i) * 8 + j];
        row[NX * k + 8 + i] = -1.0;
      }
      for (int i = 0; i < NI; i++) {
        double* row = jac + (size_t)(NG * k + 36 + i) * n;
        for (int j = 0; j < NX; j++) row[NX * k + j] = E.Jq[(k * NI + i) * NX + j];
      }
    }
  }
  if (hess) {
    memset(hess, 0, sizeof(double) * n * n);
    for (int k = 0; k < N; k++)
      for (int i = 0; i < NX; i++)
        for (int j = 0; j < NX; j++) {
          hess[(size_t)(NX * k + i) * n + NX * k + j] = E.Wd[k * NX * NX + i * NX + j];
          if (k > 0) {
            double v = E.Wo[k * NX * NX + i * NX + j];
            hess[(size_t)(NX * k + i) * n + NX * (k - 1) + j] = v;
            hess[(size_t)(NX * (k - 1) + j) * n + NX * k + i] = v;
          }
        }
  }
  return 0;
}

// interval-form derivatives (the form the interior-point iteration works on):
// lam [48 N] per stage 36 equality + 12 interval multipliers; jac [48 N x n]; hess of f + lam.(c,d)
int orc_derivs_interval(int N, int S, double dt, const double* x, const double* p, const double* lam,
                        double* d, double* gradf, double* jac, double* hess) {
  Prob P(N, S, dt);
  Eval E(N);
  eval_full(P, x, p, lam, E, true);
  int n = P.n;
  const int NL = NE + ND;
  for (int i = 0; i < ND * N; i++) d[i] = E.d[i];
  for (int i = 0; i < n; i++) gradf[i] = E.gradf[i];
  memset(jac, 0, sizeof(double) * NL * N * n);
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < NE; i++) {
      double* row = jac + (size_t)(NL * k + i) * n;
      if (k > 0) for (int j = 0; j < NX; j++) row[NX * (k - 1) + j] = E.A[(k * NE + i) * NX + j];
      for (int j = 0; j < 8; j++) row[NX * k + j] = E.B[(k * NE + i) * 8 + j];
      row[NX * k + 8 + i] = -1.0;
    }
    for (int i = 0; i < ND; i++) {
      double* row = jac + (size_t)(NL * k + NE + i) * n;
      for (int j = 0; j < NX; j++) row[NX * k + j] = E.Jd[(k * ND + i) * NX + j];
    }
  }
  memset(hess, 0, sizeof(double) * n * n);
  for (int k = 0; k < N; k++)
    for (int i = 0; i < NX; i++)
      for (int j = 0; j < NX; j++) {
        hess[(size_t)(NX * k + i) * n + NX * k + j] = E.Wd[k * NX * NX + i * NX + j];
        if (k > 0) {
          double v = E.Wo[k * NX * NX + i * NX + j];
          hess[(size_t)(NX * k + i) * n + NX * (k - 1) + j] = v;
          hess[(size_t)(NX * (k - 1) + j) * n + NX * k + i] = v;
        }
      }
  return 0;
}

static thread_local double* g_trace; static thread_local int g_trace_cap;
// opts: [tol, max_iter, mu_init, bound_push, verbose, mu_strategy (< 0: default), max_soc (< 0: default)]
int orc_solve(int N, int S, double dt, const double* x0, const double* p, const double* opts,
              double* x, double* g, double* lam_g, double* lam_x, double* f, int* iters, double* kkt) {
  Prob P(N, S, dt);
  Opts o;
  if (opts) {
    if (opts[0] > 0) o.tol = opts[0];
    if (opts[1] > 0) o.max_iter = (int)opts[1];
    if (opts[2] > 0) o.mu_init = opts[2];
    if (opts[3] > 0) o.bound_push = opts[3];
    o.verbose = (int)opts[4];
  }
  if (opts && opts[5] >= 0) o.mu_strategy = (int)opts[5];
  if (opts && opts[6] >= 0) o.max_soc = (int)opts[6];
  if (const char* e = getenv("ORC_ROLLTHR")) o.rollout_thr = atof(e);
  if (const char* e = getenv("ORC_STALL")) o.stall_stop = atoi(e);
  if (const char* e = getenv("ORC_SOCB")) o.soc_budget = atoi(e);
  Ipm ipm(P, p, o);
  ipm.trace = g_trace; ipm.trace_cap = g_trace_cap;
  int it = 0;
  double kk = 0;
  int status = ipm.solve(x0, it, kk);
  for (int i = 0; i < P.n; i++) { x[i] = ipm.x[i]; lam_x[i] = ipm.zU[i] - ipm.zL[i]; }
  std::vector<double> dd(ND * N);
  eval_values(P, x, p, *f, g, dd.data());
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < NE; i++) lam_g[NG * k + i] = ipm.y[NE * k + i];
    const double* z = &ipm.zs[ND * k];
    const double* d = &dd[ND * k];
    lam_g[NG * k + 36] = z[0];
    lam_g[NG * k + 37] = z[1];
    // (m)^2 - h^2 <= 0 vs (m - h <= 0, -m - h <= 0):  lam = (z+ + z-) / (2 h),  h = -(d+ + d-)/2
    for (int j = 0; j < 5; j++) {
      double h = -0.5 * (d[2 + 2 * j] + d[3 + 2 * j]);
      lam_g[NG * k + 38 + j] = (z[2 + 2 * j] + z[3 + 2 * j]) / (2 * h);
    }
  }
  *iters = it;
  *kkt = kk;
  return status;
}

long orc_soc_count(int which) { return which ? g_soc_accepted.load() : g_soc_solves.load(); }
void orc_set_trace(double* buf, int cap) { g_trace = buf; g_trace_cap = cap; }

}  // extern "C"
