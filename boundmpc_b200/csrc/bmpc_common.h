// boundmpc_b200 — shared definitions of the CUDA solver.
//
// Execution model: one CTA solves one OCP instance ("instance" = one call of the
// reference's `self.solver(x0, lbx, ubx, lbg, ubg, p)`, BoundMPC.py:446-453).  The code is
// written as a sequence of *parallel phases*: `PAR_FOR(i, n)` distributes the items of a
// phase over the threads of the CTA, `BMPC_SYNC()` separates phases, and block-wide
// reductions hand their result to every thread so that control flow stays CTA-uniform.
// With BMPC_HOST_EMU defined the same source compiles for the host with a single
// "thread" (tid 0 of 1); tests/emu uses that build to check the kernel source on CPU
// machines.  The product library never contains the host build.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef BMPC_HOST_EMU
#define BMPC_DEV inline
#define BMPC_NOINLINE inline
#define BMPC_HD inline
#define BMPC_SYNC() ((void)0)
#define BMPC_LDG(ptr) (*(ptr))
#define BMPC_LDCG(ptr) (*(ptr))
#else
#define BMPC_DEV __device__ __forceinline__
// large phase bodies exist once in the kernel image (the interior-point loop has to stay resident in
// the instruction cache; everything force-inlined was 435 KB of SASS)
#define BMPC_NOINLINE __device__ __noinline__
#define BMPC_HD __host__ __device__ inline
#define BMPC_SYNC() __syncthreads()
#define BMPC_LDG(ptr) __ldg(ptr)
#define BMPC_LDCG(ptr) __ldcg(ptr)   // data another SM wrote during this launch: read at L2
#endif

// optional per-phase cycle accounting of thread 0 (development build, -DBMPC_TIMING)
#ifdef BMPC_TIMING
#define BMPC_TMARK(id) do { if (cx.tid == 0) { long long t_ = clock64(); cx.tm[id] += t_ - cx.tm[63]; cx.tm[63] = t_; } } while (0)
#define BMPC_TMARK2(id) do { if (cx.tid == 64) { long long t_ = clock64(); cx.tm[id] += t_ - cx.tm[62]; cx.tm[62] = t_; } } while (0)
#else
#define BMPC_TMARK(id) ((void)0)
#define BMPC_TMARK2(id) ((void)0)
#endif

#define PAR_FOR(i, n) for (int i = cx.tid; i < (n); i += cx.nt)
// items of a phase executed by a single warp (no CTA barrier needed between such phases)
#ifdef BMPC_HOST_EMU
#define LANE_FOR(l, n) for (int l = 0; l < (n); l++)
#define BMPC_WSYNC() ((void)0)
// items of a phase dealt to the threads of warps [w0, w1) only (other warps skip the loop)
#define ROLE_FOR(i, n, w0, w1) for (int i = 0; i < (n); i++)
#else
#define LANE_FOR(l, n) for (int l = (cx.tid & 31); l < (n); l += 32)
#define BMPC_WSYNC() __syncwarp()
#define ROLE_FOR(i, n, w0, w1) \
  for (int i = ((cx.tid >> 5) >= (w0) && (cx.tid >> 5) < (w1)) ? cx.tid - 32 * (w0) : (n); i < (n); i += 32 * ((w1) - (w0)))
#endif

namespace bmpc {

struct Ctx {
  int tid, nt;
  double* red;  // shared scratch for block reductions (8 values x 8 warps)
#ifdef BMPC_TIMING
  long long* tm;  // [64] per-phase cycles, [63] = last time stamp
#endif
};
BMPC_DEV int ctx_warp(const Ctx cx) { return cx.tid >> 5; }
BMPC_DEV int ctx_nwarps(const Ctx cx) { return (cx.nt + 31) >> 5; }
// is this thread in warps [w0, w1)?  (host emulation: the single thread plays every role)
BMPC_DEV bool in_role(const Ctx cx, int w0, int w1) {
#ifdef BMPC_HOST_EMU
  (void)cx; (void)w0; (void)w1;
  return true;
#else
  return (cx.tid >> 5) >= w0 && (cx.tid >> 5) < w1;
#endif
}

// One 8 x 8 output tile of a matrix product on the FP64 tensor-core path, executed by one warp:
//   C(r, c) = sum_{kk < 4 ksteps} a(r, kk) * b(kk, c),   r, c in 0..7,
// with `mma.sync.aligned.m8n8k4.f64` (SASS DMMA).  `a` and `b` are element accessors into shared
// memory; `epi(r, c, value)` receives every element of the tile exactly once (two per lane).
// Fragment layout (PTX ISA, m8n8k4 .f64): lane l holds A(l / 4, l % 4), B(l % 4, l / 4) and
// C(l / 4, 2 (l % 4) + {0, 1}).  The host-emulation build evaluates the same tile with loops.
template <class FA, class FB, class FE>
BMPC_DEV void mma_tile(const Ctx cx, int ksteps, FA a, FB b, FE epi) {
#ifdef BMPC_HOST_EMU
  (void)cx;
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) {
      double v = 0.0;
      for (int kk = 0; kk < 4 * ksteps; kk++) v += a(r, kk) * b(kk, c);
      epi(r, c, v);
    }
#else
  const int lane = cx.tid & 31, r = lane >> 2, q = lane & 3;
  double c0 = 0.0, c1 = 0.0;
  for (int ks = 0; ks < ksteps; ks++) {
    const double av = a(r, 4 * ks + q), bv = b(4 * ks + q, r);
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
  }
  epi(r, 2 * q, c0);
  epi(r, 2 * q + 1, c1);
#endif
}
// Row block of tiles that share their A fragments: for t < nt (<= NT)
//   C_t(r, c) = cin(t, r, c) + sum_{kk < 4 KS} a(r, kk) * b(t, kk, c),
// executed by one warp; the A fragments are loaded once, C starts from cin (a previous pass) and
// every element is handed to epi(t, r, c, value) exactly once.
template <int KS, int NT, class FA, class FB, class FC, class FE>
BMPC_DEV void mma_rowblock(const Ctx cx, int nt, FA a, FB b, FC cin, FE epi) {
#ifdef BMPC_HOST_EMU
  (void)cx;
  for (int t = 0; t < nt; t++)
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) {
        double v = cin(t, r, c);
        for (int kk = 0; kk < 4 * KS; kk++) v += a(r, kk) * b(t, kk, c);
        epi(t, r, c, v);
      }
#else
  const int lane = cx.tid & 31, r = lane >> 2, q = lane & 3;
  // k-step outer, tile inner: the NT accumulator pairs are independent, so the (long-latency) DMMAs of
  // one k-step are all in flight together instead of forming one serial chain per tile
  double av[KS], acc[NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) av[ks] = a(r, 4 * ks + q);
#pragma unroll
  for (int t = 0; t < NT; t++) {
    acc[t][0] = t < nt ? cin(t, r, 2 * q) : 0.0;
    acc[t][1] = t < nt ? cin(t, r, 2 * q + 1) : 0.0;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
#pragma unroll
    for (int t = 0; t < NT; t++) {
      if (t < nt) {
        const double bv = b(t, 4 * ks + q, r);
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(av[ks]), "d"(bv));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NT; t++) {
    if (t < nt) {
      epi(t, r, 2 * q, acc[t][0]);
      epi(t, r, 2 * q + 1, acc[t][1]);
    }
  }
#endif
}
// Column block of tiles that share their B fragments: C_t(r, c) = sum_{kk < 4 KS} a(t, r, kk) * b(kk, c) for
// t < nt (<= NT); same interleaving of the independent accumulators as mma_rowblock.
template <int KS, int NT, class FA, class FB, class FE>
BMPC_DEV void mma_colblock(const Ctx cx, int nt, FA a, FB b, FE epi) {
#ifdef BMPC_HOST_EMU
  (void)cx;
  for (int t = 0; t < nt; t++)
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) {
        double v = 0.0;
        for (int kk = 0; kk < 4 * KS; kk++) v += a(t, r, kk) * b(kk, c);
        epi(t, r, c, v);
      }
#else
  const int lane = cx.tid & 31, r = lane >> 2, q = lane & 3;
  double bv[KS], acc[NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) bv[ks] = b(4 * ks + q, r);
#pragma unroll
  for (int t = 0; t < NT; t++) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
#pragma unroll
    for (int t = 0; t < NT; t++) {
      if (t < nt) {
        const double av = a(t, r, 4 * ks + q);
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[t][0]), "+d"(acc[t][1]) : "d"(av), "d"(bv[ks]));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NT; t++) {
    if (t < nt) {
      epi(t, r, 2 * q, acc[t][0]);
      epi(t, r, 2 * q + 1, acc[t][1]);
    }
  }
#endif
}
// tiles of a phase are dealt to the warps of the CTA round-robin
#define TILE_FOR(t, n) for (int t = ctx_warp(cx); t < (n); t += ctx_nwarps(cx))

// ---- stage layout (casadi_ocp_formulation.py:90-153) ----------------------
constexpr int NX = 44;   // variables per stage: [u(7) u_phi q(7) dq(7) ddq(7) p_pos(3) p_rot(3) v_lin(3) v_ang(3) phi dphi ddphi]
constexpr int NU = 8;
constexpr int NE = 36;   // equality rows per stage (casadi_ocp_formulation.py:271-303)
constexpr int NQ = 7;    // inequality rows per stage in the reference's form (casadi_ocp_formulation.py:305-349)
constexpr int NG = 43;   // rows of g per stage
constexpr int ND = 12;   // inequality rows in interval form (see DESIGN.md "interval form")
constexpr int NZ = 52;   // (s_k, u_k): previous stage block + current input
enum { oU = 0, oUPHI = 7, oQ = 8, oDQ = 15, oDDQ = 22, oPPOS = 29, oPROT = 32, oVLIN = 35,
       oVANG = 38, oPHI = 41, oDPHI = 42, oDDPHI = 43 };
// rows of the x-part (stage offset minus 8); the 12 kinematic rows are contiguous
constexpr int rKIN = 21;  // p_pos(3) p_rot(3) v_lin(3) v_ang(3)
constexpr int NK = 12;

// ---- parameter vector layout for nr_segs = S (casadi_ocp_formulation.py:361-376, App. A.3)
struct PLayout {
  int S, np;
  int q0, dq0, ddq0, phi0, p0, v0, iwref, dtau, par, orth1, orth2, xphid, jerk, phisw, jacr, jacl;
  int pref, dpref, dpn, bp1, bp2, br1, br2, a4, a3, a2, a1, a0, w, phimax, dphimax, v1, v2, v3, qd;
};
BMPC_HD PLayout make_layout(int S) {
  PLayout L;
  L.S = S;
  L.q0 = 0; L.dq0 = 7; L.ddq0 = 14; L.phi0 = 21; L.p0 = 24; L.v0 = 30; L.iwref = 36; L.dtau = 39; L.par = 42;
  L.orth1 = 42 + 3 * S; L.orth2 = 42 + 6 * S; L.xphid = 42 + 9 * S; L.jerk = 45 + 9 * S;
  L.phisw = 53 + 9 * S; L.jacr = 54 + 10 * S; L.jacl = 63 + 10 * S; L.pref = 72 + 10 * S;
  L.dpref = 72 + 16 * S; L.dpn = 72 + 22 * S; L.bp1 = 72 + 25 * S; L.bp2 = 72 + 28 * S;
  L.br1 = 72 + 31 * S; L.br2 = 72 + 34 * S; L.a4 = 72 + 37 * S; L.a3 = 81 + 46 * S;
  L.a2 = 90 + 55 * S; L.a1 = 99 + 64 * S; L.a0 = 108 + 73 * S; L.w = 117 + 82 * S;
  L.phimax = 132 + 82 * S; L.dphimax = 133 + 82 * S; L.v1 = 134 + 82 * S; L.v2 = 134 + 85 * S;
  L.v3 = 134 + 88 * S; L.qd = 134 + 91 * S; L.np = 141 + 91 * S;
  return L;
}

// ---- problem description shared by all instances of a handle -------------------------------
struct Config {
  int N, S, n, m, np;
  double dt;
  double lb[NX], ub[NX];   // per-stage variable bounds (+-inf where free)
  // interior-point options
  double tol, diverge_tol;
  int max_iter;
  double mu_init, bound_push;
  double kappa_eps, kappa_mu, theta_mu, tau_min, s_max;
  double rollout_thr;            // cold-start repair (bmpc_ipm.cuh): equality violation of the start above which its states are rolled out
  double boost_fac, boost_cap;   // re-centring of a crawling iteration (bmpc_ipm.cuh): mu <- min(cap, fac * mu)
  int boost_budget;              // re-centrings per solve before it is stopped as locally infeasible
  int stall_stop;                // stop as locally infeasible at the stall_stop-th failed progress test with mu at its cap (0: never)
  int soc_budget;                // corrections rejected in a row after which none is tried any more in a solve
  int max_soc;                   // second-order corrections per iteration (0 or 1)
  int qss_late;                  // Riccati stage: Q_ss pass in the gain phase (idle warps) instead of the Q_u phase
  int hard_continue;             // != 0: instances flagged hard at the end of their slice continue instead of being parked
  int slice_iters;               // iterations of pass A of the two-pass scheduling (bmpc_ipm.cuh)
  int red_iters;                 // ... when the optimality error has not improved on any of the last red_iters iterates (<= 4)
  double gamma_theta, gamma_phi, eta_phi, s_phi, s_theta;
  // integration coefficients of the piecewise-linear jerk at t = h (App. A.4)
  double a_dq, a_ddq, a_um, a_u, b_ddq, b_um, b_u, c_um, c_u;
  PLayout L;
};

// ---- per-stage record produced by the evaluation phases (doubles) --------------------------
// GK    [12][52]  kinematic rows of [A_hat | B]  (columns: previous stage block (44), u_k (8))
// HQQn  [7][7]    multiplier-weighted d2 Phi / dq_n dq_n      (integrated state)
// HQDn  [7][7]    d2 Phi / dq_n d(dq_n)   (row: q, col: dq)
// HQQk, HQDk      same for the omega(q_k, dq_k) term of the p_rot rows
// HY    [7][7]    Hessian of the path cost w.r.t. y = (p_pos, p_rot, phi)
// GY    [7]       gradient of the path cost w.r.t. y
// JD    [12][8]   gradients of the interval rows: y(7) then dphi
// HD    [12]      second derivative of each interval row w.r.t. phi
// DPD   [6]       dp_d of the active segment (velocity / acceleration tracking terms)
// MISC            see offsets
enum {
  R_GK = 0,
  R_HQQN = R_GK + NK * NZ,
  R_HQDN = R_HQQN + 49,
  R_HQQK = R_HQDN + 49,
  R_HQDK = R_HQQK + 49,
  R_HY = R_HQDK + 49,
  R_GY = R_HY + 49,
  R_JD = R_GY + 7,
  R_HD = R_JD + ND * 8,
  R_DPD = R_HD + ND,
  R_GV = R_DPD + 6,        // [6] gradient w.r.t. v (velocity + acceleration tracking)
  R_GVP = R_GV + 6,        // [6] gradient w.r.t. v_prev
  R_GPH = R_GVP + 6,       // [2] gradient w.r.t. dphi, ddphi (tracking + path-state cost); phi part is in GY
  R_COST = R_GPH + 2,      // [1] stage cost
  R_HYB = R_COST + 2,      // [8][8] y-block (p_pos, p_rot, phi, dphi) of W~: HY + z.HD + J_d^T Sigma_s J_d + dphi tracking term
  R_GJ1 = R_HYB + 64,      // [8] sum_r JD_r / s_r              (multiplied by mu in g^)
  R_GJ2 = R_GJ1 + 8,       // [8] sum_r JD_r Sigma_r (d_r + s_r)
  R_SIG = R_GJ2 + 8,       // [3][12] Sigma_r = z_r / s_r, 1 / s_r, Sigma_r (d_r + s_r)
  R_SIZE = R_SIG + 3 * ND
};
// the part [R_HY, R_SIZE) of a record is produced by the (per-lane serial) path terms; it is staged in
// shared memory with stride R_PATH and copied out coalesced (see work_attach_smem)
constexpr int R_PATH = R_SIZE - R_HY;
// forward-kinematics scratch of one chain evaluation
enum {
  F_Z = 0,                 // [7][3] joint axes
  F_R = F_Z + 21,          // [7][3] tool point minus joint origin
  F_W = F_R + 21,          // [7][3] sum_{k>=i} dq_k z_k x r_k
  F_OT = F_W + 21,         // [7][3] sum_{k>i} dq_k z_k
  F_OH = F_OT + 21,        // [8][3] sum_{k<i} dq_k z_k   (entry 7 = total angular velocity)
  F_POS = F_OH + 24,       // [3]
  F_SN = F_POS + 3,        // [7] sin / cos of the joint angles of this evaluation
  F_CS = F_SN + 7,         // [7]
  F_DQ = F_CS + 7,         // [7]
  F_SIZE = F_DQ + 7 + 1    // (odd stride: the chains of a warp hit different shared-memory banks)
};

// asynchronous 8-byte copy global -> shared (LDGSTS): issued back to back, completed by cp_async_wait()
// of the issuing thread plus a barrier for the readers
BMPC_DEV void cp_async8(double* dst_smem, const double* src) {
#ifdef BMPC_HOST_EMU
  *dst_smem = *src;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
#endif
}
BMPC_DEV void cp_async16(double* dst_smem, const double* src) {   // two doubles, both addresses 16-byte aligned
#ifdef BMPC_HOST_EMU
  dst_smem[0] = src[0]; dst_smem[1] = src[1];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
#endif
}
BMPC_DEV void cp_async_wait() {
#ifndef BMPC_HOST_EMU
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// cp.async groups: copies issued since the last commit form a group; wait_group1 returns when all groups but the
// most recent one have landed
BMPC_DEV void cp_async_commit() {
#ifndef BMPC_HOST_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
BMPC_DEV void cp_async_wait_group1() {
#ifndef BMPC_HOST_EMU
  asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}

// one copy each of the transcendental routines
BMPC_NOINLINE double bmpc_log(double v) { return log(v); }
BMPC_NOINLINE double bmpc_exp(double v) { return exp(v); }
BMPC_DEV double bmpc_pow(double v, double e) { return bmpc_exp(e * bmpc_log(v)); }   // v > 0

BMPC_DEV void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
BMPC_DEV double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---- block reductions: result is returned to every thread ------------------------------------
enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
constexpr int RED_WARPS = 16;   // warps per CTA the reduction scratch is sized for (largest launch shape: 512 threads)

// (one copy in the kernel image: K and the operations are run-time arguments)
BMPC_NOINLINE void block_reduce_n(const Ctx cx, double* v, const int* op, int K) {
#ifndef BMPC_HOST_EMU
  const int lane = cx.tid & 31, warp = cx.tid >> 5, nw = (cx.nt + 31) >> 5;
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    double a = v[k];
    const int o_ = op[k];
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      double b = __shfl_xor_sync(0xffffffffu, a, o);
      a = o_ == RED_SUM ? a + b : (o_ == RED_MAX ? fmax(a, b) : fmin(a, b));
    }
    if (lane == 0) cx.red[k * RED_WARPS + warp] = a;
  }
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    double a = cx.red[k * RED_WARPS];
    const int o_ = op[k];
    for (int w = 1; w < nw; w++) {
      double b = cx.red[k * RED_WARPS + w];
      a = o_ == RED_SUM ? a + b : (o_ == RED_MAX ? fmax(a, b) : fmin(a, b));
    }
    v[k] = a;
  }
  __syncthreads();
#else
  (void)cx; (void)v; (void)op; (void)K;
#endif
}
template <int K>
BMPC_DEV void block_reduce(const Ctx cx, double (&v)[K], const int (&op)[K]) { block_reduce_n(cx, v, op, K); }

// status codes returned per instance
enum { ST_SUCCESS = 0, ST_MAXITER = 1, ST_LINESEARCH = 2, ST_REGULARIZATION = 3, ST_NUMERIC = 4, ST_DIVERGING = 5 };

}  // namespace bmpc
