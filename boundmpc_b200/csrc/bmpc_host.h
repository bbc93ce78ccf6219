// boundmpc_b200 — host-side construction of the solver configuration from the C-ABI struct.
#pragma once
#include <math.h>
#include "../../include/boundmpc_b200.h"
#include "bmpc_common.h"

namespace bmpc {

// variable bounds and options exactly as setup_optimization_problem receives / builds them
// (casadi_ocp_formulation.py:92-153) and the Ipopt options of BoundMPC.py:120-141
inline int make_config(const bmpc_config& in, Config& C) {
  if (in.N < 1 || in.N > 64 || in.nr_segs < 2 || in.nr_segs > 16 || !(in.dt > 0)) return -1;
  C.N = in.N; C.S = in.nr_segs; C.n = NX * in.N; C.m = NG * in.N; C.dt = in.dt;
  C.L = make_layout(in.nr_segs);
  C.np = C.L.np;
  for (int i = 0; i < NX; i++) { C.lb[i] = -INFINITY; C.ub[i] = INFINITY; }
  for (int i = 0; i < 7; i++) {
    C.lb[oU + i] = in.u_min; C.ub[oU + i] = in.u_max;
    C.lb[oQ + i] = in.q_lim_lower[i]; C.ub[oQ + i] = in.q_lim_upper[i];
    C.lb[oDQ + i] = in.dq_lim_lower[i]; C.ub[oDQ + i] = in.dq_lim_upper[i];
  }
  C.lb[oUPHI] = in.ut_min; C.ub[oUPHI] = in.ut_max;
  C.lb[oPHI] = 0.0;
  C.tol = in.tol > 0 ? in.tol : 1e-9;
  C.max_iter = in.max_iter > 0 ? in.max_iter : 500;
  // Without a restoration phase a locally infeasible instance (a start outside the error bounds that the jerk limit
  // cannot bring back within the first nodes) drifts for 100+ iterations with growing multipliers before the line
  // search gives up; the iteration stops when the dual infeasibility passes this level (converging instances of the
  // bench workload stay below 3e5).
  C.diverge_tol = 1e7;
  // initial barrier parameter / initial bound multipliers z = mu / slack.  The reference runs Ipopt's adaptive strategy
  // with warm_start_init_point (BoundMPC.py:120-141), which starts from the complementarity of the pushed start
  // (warm_start_mult_bound_push = 1e-3) instead of the monotone mode's mu_init = 0.1; 1e-3 reproduces that level
  // and saves about four iterations per solve on warm-started instances.
  C.mu_init = in.mu_init > 0 ? in.mu_init : 1e-3;
  C.bound_push = in.bound_push > 0 ? in.bound_push : 1e-3;
  // Barrier decrease: Ipopt's monotone rule with a fast schedule (kappa_eps 10 -> 1000, kappa_mu 0.2 -> 0.1, theta_mu
  // 1.5 -> 2: three barrier levels 1e-3, 1e-6, tol / 10 instead of five; 10.1 instead of 11.4 iterations per solve on the
  // bench workload, same converged points), made safe by the progress test below.
  C.kappa_eps = 1000.0; C.kappa_mu = 0.1; C.theta_mu = 2.0; C.tau_min = 0.99; C.s_max = 100.0;
  // Re-centring (see bmpc_ipm.cuh): Ipopt's kkt-error progress test with adaptive_mu_kkterror_red_iters = 3; a crawling
  // iteration gets mu <- min(1, 10 mu).  Longest solve of the 65,536-instance bench workload 125 -> 55 iterations.
  C.red_iters = 3; C.boost_fac = 10.0; C.boost_cap = 1.0;
  // ... and a solve that fails the progress test three more times at mu = 1 is stopped as locally infeasible (config 4:
  // 27 instead of 42 iterations per infeasible instance; the bench workload loses no converging instance).
  C.stall_stop = 3; C.boost_budget = 6;
  C.rollout_thr = 0.5;
  C.qss_late = 0;     // (Q_ss pass on the idle warps of the gain phase: 57.9 vs 55.4 ms, the two warps then outlast the Cholesky chains)
  // Pass A of the two-pass scheduling runs three iterations (was six).  A solve at the reference's tolerance takes 7.4
  // iterations on average, so a six-iteration slice parks most instances one iteration before their end.  All eight shards
  // of the 65,536-instance workload on one GPU at tol 1e-5: 42.9 ... 48.4 ms with six, 42.1 ... 43.5 with four
  // (profiles/r2b_shard_slice.json, four priority lists); with the fifth list 41.7 ... 43.1 ms with three against 42.1 ...
  // 44.0 with four (r2b_shard_slice_5lists.json); at tol 1e-9 all of them within noise (55.3 ... 56.2 ms).
  C.slice_iters = 3; C.hard_continue = 0;   // (continuing hard instances instead of parking them: slower on every shard measured)
  // One second-order correction per iteration (Ipopt: max_soc = 4; on the bench workload a second correction is never
  // accepted when the first is not): 0.9 % extra KKT solves, 7 % fewer iterations along the experiment1 closed loop.
  C.max_soc = 1; C.soc_budget = 2;
  C.gamma_theta = 1e-5; C.gamma_phi = 1e-5; C.eta_phi = 1e-8; C.s_phi = 2.3; C.s_theta = 1.1;
  const double h = in.dt;
  C.a_dq = h; C.a_ddq = h * h / 2; C.a_um = h * h * h / 8; C.a_u = h * h * h / 24;
  C.b_ddq = h; C.b_um = h * h / 3; C.b_u = h * h / 6; C.c_um = h / 2; C.c_u = h / 2;
  return 0;
}

inline void fill_bounds(const Config& C, double* lbx, double* ubx, double* lbg, double* ubg) {
  for (int k = 0; k < C.N; k++) {
    for (int i = 0; i < NX; i++) { lbx[NX * k + i] = C.lb[i]; ubx[NX * k + i] = C.ub[i]; }
    for (int i = 0; i < NG; i++) { lbg[NG * k + i] = i < NE ? 0.0 : -INFINITY; ubg[NG * k + i] = 0.0; }
  }
}

}  // namespace bmpc
