// TEST INFRASTRUCTURE — host emulation build of the CUDA solver source.
// Compiles boundmpc_b200/csrc/*.cuh with BMPC_HOST_EMU (one "thread" per CTA, phases run
// sequentially) so that the kernel source can be checked against the oracle on machines
// without a GPU.  Never part of the product library and never loaded by boundmpc_b200.
#define BMPC_HOST_EMU 1
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../boundmpc_b200/csrc/bmpc_host.h"
#include "../../boundmpc_b200/csrc/bmpc_eval.cuh"
#include "../../boundmpc_b200/csrc/bmpc_prepare.cuh"
#include "../../boundmpc_b200/csrc/bmpc_post.cuh"

using namespace bmpc;

extern "C" {

int emu_solve(const bmpc_config* cfg, int batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
              double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  std::vector<double> ws(work_doubles(C.N));
  Smem* S = new Smem;
  Work W;
  work_carve(W, ws.data(), C.N);
  work_attach_smem(W, *S, C.N);
  Ctx cx{0, 1, S->red};
  build_tables(cx, C, *S);
  phase_kin_jacobian_init(cx, C, W);
  for (int b = 0; b < batch; b++) {
    InstanceIO io{x0 + (size_t)b * C.n, p + (size_t)b * C.np, x + (size_t)b * C.n, g + (size_t)b * C.m,
                  lam_g + (size_t)b * C.m, lam_x + (size_t)b * C.n, f + b, kkt + b, iters + b, status + b};
    solve_instance(cx, C, W, *S, io);
  }
  delete S;
  return 0;
}

// the two-pass path of k_solve: every instance is parked after its first slice (pass A), then all are resumed in
// reverse order (pass B) on the same scratch; hard[b] receives the "hard" flag of the parking decision (-1: finished in pass A)
int emu_solve_sliced(const bmpc_config* cfg, int batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
                     double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt, int32_t* hard) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  C.hard_continue = 0;          // park every instance, the hard ones too: the save / restore path is what this entry tests
  std::vector<double> ws(work_doubles(C.N));
  const size_t stride = save_doubles(C.N);
  std::vector<double> save(stride * (size_t)batch);
  Smem* S = new Smem;
  Work W;
  work_carve(W, ws.data(), C.N);
  work_attach_smem(W, *S, C.N);
  Ctx cx{0, 1, S->red};
  build_tables(cx, C, *S);
  phase_kin_jacobian_init(cx, C, W);
  auto make_io = [&](int b) {
    return InstanceIO{x0 + (size_t)b * C.n, p + (size_t)b * C.np, x + (size_t)b * C.n, g + (size_t)b * C.m,
                      lam_g + (size_t)b * C.m, lam_x + (size_t)b * C.n, f + b, kkt + b, iters + b, status + b};
  };
  for (int b = 0; b < batch; b++) {
    const int rc = solve_instance(cx, C, W, *S, make_io(b), RUN_SLICE, save.data() + stride * b);
    hard[b] = rc == DONE ? -1 : (rc == PARKED ? 1 : 0);   // (priority list 0 = "hard")
  }
  for (int b = batch - 1; b >= 0; b--)
    if (hard[b] >= 0) solve_instance(cx, C, W, *S, make_io(b), RUN_RESUME, save.data() + stride * b);
  delete S;
  return 0;
}

// parameter builder (csrc/bmpc_prepare.cuh), serial form
int emu_prepare(int N, int S, int batch, const double* tabs, int J, const int32_t* path_id, int32_t* sector, const double* state,
                const double* prev, double* x0, double* p) {
  const PLayout L = make_layout(S);
  if (S > PREP_SMAX) return -1;
  Config C;                      // (only N is read by the re-projected warm start)
  C.N = N;
  for (int b = 0; b < batch; b++) {
    const double* tab = tabs + (size_t)path_id[b] * J * PT_ROW;
    const double* st = state + (size_t)b * PS_SIZE;
    sector[b] = prepare_instance(L, N, tab, J, sector[b], st, prev + (size_t)b * NX * N, x0 + (size_t)b * NX * N, p + (size_t)b * L.np);
    if (st[PS_HASPREV] != 0.0 && st[PS_UPDATED] != 0.0)          // as in k_prepare
      warm_start_updated_at(C, tab, sector[b], st, prev + (size_t)b * NX * N, x0 + (size_t)b * NX * N);
  }
  return 0;
}

// BoundMPC.update (k_update)
int emu_update(int batch, const double* tabs, int J, const double* phimax, const int32_t* new_path, const double* cart, double* state,
               int32_t* sector, int32_t* path_id) {
  for (int b = 0; b < batch; b++) {
    if (new_path[b] < 0) continue;
    update_state(tabs + (size_t)new_path[b] * J * PT_ROW, phimax[new_path[b]], cart + (size_t)b * 24, state + (size_t)b * PS_SIZE);
    sector[b] = 0;
    path_id[b] = new_path[b];
  }
  return 0;
}

// post-processing (csrc/bmpc_post.cuh), serial form
int emu_post(const bmpc_config* cfg, int batch, const double* tabs, int J, const int32_t* path_id, const int32_t* sector, const double* state,
             const double* w, const int32_t* ec, double* traj, double* state_out) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  for (int b = 0; b < batch; b++)
    post_instance(C, tabs + (size_t)path_id[b] * J * PT_ROW, sector[b], state + (size_t)b * PS_SIZE, w + (size_t)b * C.n, ec[b],
                  traj + (size_t)b * C.N * TR_ROW, state_out + (size_t)b * PS_SIZE);
  return 0;
}

// post-processing with the logging branch (serial form of k_post with ref / err)
int emu_post_log(const bmpc_config* cfg, int batch, const double* tabs, int J, const int32_t* path_id, const int32_t* sector, const double* state,
                 const double* p, const double* w, const int32_t* ec, double* traj, double* state_out, double* ref, double* err) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  for (int b = 0; b < batch; b++) {
    const double* tab = tabs + (size_t)path_id[b] * J * PT_ROW;
    double* T = traj + (size_t)b * C.N * TR_ROW;
    double* so = state_out + (size_t)b * PS_SIZE;
    post_instance(C, tab, sector[b], state + (size_t)b * PS_SIZE, w + (size_t)b * C.n, ec[b], T, so);
    log_instance(C, tab, sector[b], state + (size_t)b * PS_SIZE, p + (size_t)b * C.np, T, ec[b], so + PS_PRREF, ref + (size_t)b * C.N * RF_ROW,
                 err + (size_t)b * C.N * ER_ROW);
  }
  return 0;
}

// second half of BoundMPC.step + closed-loop advance (k_finish of bmpc_kernels.cu, serial form)
int emu_finish(const bmpc_config* cfg, int batch, const double* tabs, int J, const int32_t* path_id, const int32_t* sector, const double* state,
               const double* x, const double* g, const int32_t* status, double* prev, int32_t* ec, double* traj, double* state_out, int advance) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  for (int b = 0; b < batch; b++) {
    const double* st = state + (size_t)b * PS_SIZE;
    const int d = finish_decision(C, status[b], g + (size_t)b * C.m, st[PS_HASPREV] != 0.0);
    const int e = d == 1 ? ec[b] + 1 : 0;
    const double* w = d == 1 ? prev + (size_t)b * C.n : x + (size_t)b * C.n;
    double* T = traj + (size_t)b * C.N * TR_ROW;
    double* so = state_out + (size_t)b * PS_SIZE;
    if (e < C.N) {
      post_instance(C, tabs + (size_t)path_id[b] * J * PT_ROW, sector[b], st, w, e, T, so);
      if (advance) advance_state(w, e, T, so);
    } else {
      for (int i = 0; i < C.N * TR_ROW; i++) T[i] = 0.0;
      for (int i = 0; i < PS_SIZE; i++) so[i] = st[i];
    }
    if (d == 0) { so[PS_HASPREV] = 1.0; memcpy(prev + (size_t)b * C.n, x + (size_t)b * C.n, sizeof(double) * C.n); }
    ec[b] = e;
  }
  return 0;
}

int emu_eval(const bmpc_config* cfg, int batch, const double* x, const double* p, const double* lam, double* f, double* g,
             double* d, double* grad, double* jac, double* hess) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  std::vector<double> ws(work_doubles(C.N));
  Smem* S = new Smem;
  Work W;
  work_carve(W, ws.data(), C.N);
  work_attach_smem(W, *S, C.N);
  Ctx cx{0, 1, S->red};
  build_tables(cx, C, *S);
  phase_kin_jacobian_init(cx, C, W);
  const size_t n = C.n, nl = (size_t)(NE + ND) * C.N;
  for (int b = 0; b < batch; b++) {
    EvalIO io{x + b * n, p + (size_t)b * C.np, lam ? lam + b * nl : nullptr, f ? f + b : nullptr, g ? g + (size_t)b * C.m : nullptr,
              d ? d + (size_t)b * ND * C.N : nullptr, grad ? grad + b * n : nullptr, jac ? jac + b * nl * n : nullptr,
              hess ? hess + b * n * n : nullptr};
    eval_instance(cx, C, W, *S, io);
  }
  delete S;
  return 0;
}
int emu_kkt_step(const bmpc_config* cfg, int batch, const double* v, const double* p, const double* mu, const double* dw, double* dx,
                 double* ynew, int32_t* ok) {
  Config C;
  if (make_config(*cfg, C)) return -1;
  std::vector<double> ws(work_doubles(C.N));
  Smem* S = new Smem;
  Work W;
  work_carve(W, ws.data(), C.N);
  work_attach_smem(W, *S, C.N);
  Ctx cx{0, 1, S->red};
  build_tables(cx, C, *S);
  phase_kin_jacobian_init(cx, C, W);
  const size_t n = C.n, ne = (size_t)NE * C.N, nv = 3 * n + ne + 2 * (size_t)ND * C.N;
  for (int b = 0; b < batch; b++) {
    KktIO io{v + b * nv, p + (size_t)b * C.np, mu[b], dw[b], dx + b * n, ynew + b * ne, ok + b};
    kkt_step_instance(cx, C, W, *S, io);
  }
  delete S;
  return 0;
}
}
