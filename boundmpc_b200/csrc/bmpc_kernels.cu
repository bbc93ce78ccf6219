// boundmpc_b200 — CUDA kernels (sm_100a) and the C ABI declared in include/boundmpc_b200.h.
//
// k_solve: persistent CTAs, one OCP instance per CTA at a time, instances pulled from an
// atomic work queue so that instances with few interior-point iterations free their SM early
// (SURVEY 7.1 K0-K5).  Per-CTA state lives in a global workspace slice that stays L2-resident;
// the Riccati blocks are staged in shared memory (struct Smem).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "bmpc_host.h"
#include "bmpc_eval.cuh"
#include "bmpc_prepare.cuh"
#include "bmpc_post.cuh"

using namespace bmpc;

#ifndef BMPC_MAX_THREADS
#define BMPC_MAX_THREADS 256
#endif
#ifndef BMPC_DEFAULT_THREADS
#define BMPC_DEFAULT_THREADS 128
#define BMPC_DEFAULT_CTAS 3
#endif

struct BatchIO {
  const double* x0; const double* p;
  double *x, *g, *lam_g, *lam_x, *f, *kkt;
  int32_t *iters, *status;
  // host entry with page-locked inputs: device-addressable host copies of x0 / p; the CTA that starts an instance
  // fetches its 7.5 KB into the device buffers x0 / p (which it owns until then) instead of a copy before the launch
  const double* x0_src; const double* p_src;
};

#ifdef BMPC_TRACE
// development build: per-instance time stamps (globaltimer, ns) of the scheduler: [b][0..1] first run, [b][2..3] resumed run
__device__ unsigned long long g_trace[65536 * 4];
extern "C" int bmpc_trace(unsigned long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 4 * n) == cudaSuccess ? 0 : -2;
}
extern "C" int bmpc_itlog(double* out, int enable) {
  if (enable >= 0) return cudaMemcpyToSymbol(bmpc::g_itlog_on, &enable, sizeof(int)) == cudaSuccess ? 0 : -2;
  return cudaMemcpyFromSymbol(out, bmpc::g_itlog, sizeof(double) * 12 * 500) == cudaSuccess ? 0 : -2;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
#ifdef BMPC_TIMING
__device__ unsigned long long g_phase_cycles[64];
extern "C" int bmpc_phase_cycles(unsigned long long* out, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
  if (e == cudaSuccess && reset) { unsigned long long z[64] = {0}; e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -2;
}
#endif

// Work queue of a launch (first 256 bytes of the workspace).  Single pass (batch <= grid): `next` hands out instances.
// Two passes (see SCHED_LISTS in bmpc_ipm.cuh): `next` hands out the pass-A slices; a CTA that has parked an instance
// appends it to its priority list (tail, then the slot, then `sliced`); once pass A is handed out, CTAs pop the lists in
// priority order and leave when every slice has been parked or finished and all lists are drained.
struct Sched { unsigned int next, sliced, tail[SCHED_LISTS], head[SCHED_LISTS]; };
static_assert(sizeof(Sched) <= 256, "queue header");
struct SchedMem { Sched* sc; int* list; int batch; double* save; size_t save_stride; int two_pass; };   // list [SCHED_LISTS][batch]

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) { return *(const volatile unsigned int*)p; }

// thread 0: next piece of work of this CTA; returns the instance (or -1: nothing left) and the mode
__device__ int sched_next(const SchedMem& M, int batch, int* mode) {
  Sched* sc = M.sc;
  const unsigned int t = atomicAdd(&sc->next, 1u);
  if (t < (unsigned int)batch) { *mode = M.two_pass ? RUN_SLICE : RUN_FULL; return (int)t; }
  if (!M.two_pass) return -1;
  *mode = RUN_RESUME;
  for (;;) {
    for (int which = 0; which < SCHED_LISTS; which++) {
      unsigned int* head = &sc->head[which];
      const unsigned int* tail = &sc->tail[which];
      const int* list = M.list + (size_t)which * M.batch;
      for (;;) {
        const unsigned int h = ld_volatile_u32(head);
        if (h >= ld_volatile_u32(tail)) break;
        if (atomicCAS(head, h, h + 1) != h) continue;
        int b;
        while ((b = *(const volatile int*)(list + h)) < 0) __nanosleep(100);   // (slot reserved, index not yet written)
        __threadfence();
        return b;
      }
    }
    if (ld_volatile_u32(&sc->sliced) >= (unsigned int)batch) {
      __threadfence();
      bool empty = true;
      for (int which = 0; which < SCHED_LISTS; which++) empty = empty && ld_volatile_u32(&sc->head[which]) >= ld_volatile_u32(&sc->tail[which]);
      if (empty) return -1;
    } else {
      __nanosleep(1000);
    }
  }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_solve(const __grid_constant__ Config C0, int batch, BatchIO io, double* ws,
                                                         size_t ws_stride, SchedMem M, int vec_ext) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
#ifdef BMPC_TIMING
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red, S.tm};
  if (threadIdx.x < 64) S.tm[threadIdx.x] = threadIdx.x >= 62 ? clock64() : 0;
  __syncthreads();
#else
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red};
#endif
  if (threadIdx.x == 0) {
    S.cfg = C0;
    work_carve(S.work, ws + (size_t)blockIdx.x * ws_stride, C0.N);
    work_attach_smem(S.work, S, C0.N, vec_ext > 0 ? reinterpret_cast<double*>(smem_raw + sizeof(Smem)) : nullptr, vec_ext < 0);
  }
  __syncthreads();
#ifdef BMPC_PROBE_SLOTS   // development aid (scripts/probe_slots.py): SM and hardware warp slots of a few CTAs
  if ((threadIdx.x & 31) == 0 && (blockIdx.x < 3 || (blockIdx.x >= 148 && blockIdx.x < 151) || (blockIdx.x >= 296 && blockIdx.x < 299))) {
    unsigned wslot, smid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("cta %d warp %d smid %u warpid %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), smid, wslot);
  }
#endif
  const Config& C = S.cfg;
  const Work& W = S.work;
  build_tables(cx, C, S);
  phase_kin_jacobian_init(cx, C, W);
  for (;;) {
    if (threadIdx.x == 0) { int mode = RUN_FULL; S.flag[2] = sched_next(M, batch, &mode); S.flag[0] = mode; }
    __syncthreads();
    const int b = S.flag[2], mode = S.flag[0];
    __syncthreads();
    if (b < 0) break;
    InstanceIO ii{io.x0 + (size_t)b * C.n, io.p + (size_t)b * C.np, io.x + (size_t)b * C.n, io.g + (size_t)b * C.m,
                  io.lam_g + (size_t)b * C.m, io.lam_x + (size_t)b * C.n, io.f + b, io.kkt + b, io.iters + b, io.status + b};
#ifdef BMPC_TRACE
    if (threadIdx.x == 0 && b < 65536) g_trace[4 * b + (mode == RUN_RESUME ? 2 : 0)] = gtime();
#endif
    if (io.x0_src && mode != RUN_RESUME) {
      // (loads in batches of four per thread: the reads cross PCIe, one round trip per batch)
      double* dst[2] = {const_cast<double*>(ii.x0), const_cast<double*>(ii.p)};
      const double* src[2] = {io.x0_src + (size_t)b * C.n, io.p_src + (size_t)b * C.np};
      const int cnt[2] = {C.n, C.np};
#pragma unroll 1
      for (int a = 0; a < 2; a++)
#pragma unroll 1
        for (int i0 = threadIdx.x; i0 < cnt[a]; i0 += 4 * blockDim.x) {
          double v[4];
#pragma unroll
          for (int q = 0; q < 4; q++) { const int i = i0 + q * blockDim.x; v[q] = i < cnt[a] ? src[a][i] : 0.0; }
#pragma unroll
          for (int q = 0; q < 4; q++) { const int i = i0 + q * blockDim.x; if (i < cnt[a]) dst[a][i] = v[q]; }
        }
      __syncthreads();
    }
    const int rc = solve_instance(cx, C, W, S, ii, mode, M.save ? M.save + (size_t)b * M.save_stride : nullptr);
#ifdef BMPC_TRACE
    if (threadIdx.x == 0 && b < 65536) g_trace[4 * b + (mode == RUN_RESUME ? 3 : 1)] = gtime() | (rc == PARKED ? 1ull : 0ull);
#endif
    if (mode == RUN_SLICE) {
      __threadfence();          // the parked iterate, written by all threads, before the list entry that publishes it
      __syncthreads();
      if (threadIdx.x == 0) {
        if (rc != DONE) {
          const int which = rc - PARKED;
          const unsigned int slot = atomicAdd(&M.sc->tail[which], 1u);
          *(volatile int*)(M.list + (size_t)which * M.batch + slot) = b;
          __threadfence();
        }
        atomicAdd(&M.sc->sliced, 1u);
      }
    }
  }
#ifdef BMPC_TIMING
  if (threadIdx.x < 62) atomicAdd(&g_phase_cycles[threadIdx.x], (unsigned long long)S.tm[threadIdx.x]);
#endif
}

// launch variants: (threads per CTA, resident CTAs per SM the register budget is compiled for)
typedef void (*solve_fn)(const Config, int, BatchIO, double*, size_t, SchedMem, int);
struct SolveVariant { int threads, minb; solve_fn fn; };
static const SolveVariant kVariants[] = {
#ifdef BMPC_TIMING
    {128, 3, k_solve<128, 3>},
#else
    {128, 3, k_solve<128, 3>}, {160, 3, k_solve<160, 3>}, {192, 3, k_solve<192, 3>}, {256, 2, k_solve<256, 2>},
    {384, 1, k_solve<384, 1>}, {512, 1, k_solve<512, 1>},
#endif
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

struct EvalBatchIO {
  const double *x, *p, *lam;
  double *f, *g, *d, *grad, *jac, *hess;
};

__global__ void __launch_bounds__(BMPC_MAX_THREADS) k_eval(const __grid_constant__ Config C, int batch, EvalBatchIO io, double* ws,
                                                           size_t ws_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
#ifdef BMPC_TIMING
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red, S.tm};
#else
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red};
#endif
  Work W;
  work_carve(W, ws + (size_t)blockIdx.x * ws_stride, C.N);
  work_attach_smem(W, S, C.N);
  build_tables(cx, C, S);
  phase_kin_jacobian_init(cx, C, W);
  const size_t n = C.n, nl = (size_t)(NE + ND) * C.N;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    EvalIO e{io.x + b * n, io.p + (size_t)b * C.np, io.lam ? io.lam + b * nl : nullptr, io.f ? io.f + b : nullptr,
             io.g ? io.g + (size_t)b * C.m : nullptr, io.d ? io.d + (size_t)b * ND * C.N : nullptr, io.grad ? io.grad + b * n : nullptr,
             io.jac ? io.jac + b * nl * n : nullptr, io.hess ? io.hess + b * n * n : nullptr};
    eval_instance(cx, C, W, S, e);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ parameter builder
// k_prepare: the pre-solve half of BoundMPC.step for a batch (bmpc_prepare.cuh).  Phase 1: one thread per instance builds
// its parameter vector (a few hundred scalar operations; rows of p are written by their owner thread).  Phase 2: the CTA
// writes the warm starts of its instances together, consecutive threads on consecutive entries (coalesced shift copy of
// the previous solutions).
constexpr int PREP_THREADS = 128;
struct PrepIO {
  const double* tabs; int J;
  const int32_t* path_id; int32_t* sector;
  const double* state; const double* prev;
  double* x0; double* p;
};
__global__ void __launch_bounds__(PREP_THREADS) k_prepare(const __grid_constant__ Config C, int batch, PrepIO io) {
  __shared__ unsigned char rev[PREP_THREADS];
  const int base = blockIdx.x * PREP_THREADS, b = base + threadIdx.x;
  const size_t n = C.n;
  if (b < batch) {
    const double* st = io.state + (size_t)b * PS_SIZE;
    const double* tab = io.tabs + (size_t)io.path_id[b] * io.J * PT_ROW;
    const int sec = prepare_params(C.L, tab, io.J, io.sector[b], st, io.p + (size_t)b * C.np);
    io.sector[b] = sec;
    rev[threadIdx.x] = st[PS_HASPREV] != 0.0 && warm_reverse(st, io.prev + b * n);
    if (st[PS_HASPREV] != 0.0 && st[PS_UPDATED] != 0.0) {          // after a path update: re-projected warm start, written by its owner
      warm_start_updated_at(C, tab, sec, st, io.prev + b * n, io.x0 + b * n);
      rev[threadIdx.x] = 2;
    }
  }
  __syncthreads();
  const int cnt = batch - base < PREP_THREADS ? batch - base : PREP_THREADS;
  for (size_t idx = threadIdx.x; idx < (size_t)cnt * n; idx += PREP_THREADS) {
    const int i = (int)(idx / n), e = (int)(idx - (size_t)i * n), k = e / NX, a = e - NX * k;
    const size_t bi = (size_t)(base + i);
    if (rev[i] != 2) io.x0[bi * n + e] = warm_start_value(C.N, io.state + bi * PS_SIZE, io.prev + bi * n, rev[i] != 0, k, a);
  }
}

// ------------------------------------------------------------------------------------------ post-processing
// k_post: compute_return_data of BoundMPC.step for a batch (bmpc_post.cuh), one thread per instance.
struct PostIO {
  const double* tabs; int J;
  const int32_t* path_id; const int32_t* sector;
  const double* state; const double* w; const int32_t* ec;
  double* traj; double* state_out;
  const double* p; double* ref; double* err;      // logging branch (reference data, error terms): all three or none
};
__global__ void __launch_bounds__(PREP_THREADS) k_post(const __grid_constant__ Config C, int batch, PostIO io) {
  const int b = blockIdx.x * PREP_THREADS + threadIdx.x;
  if (b >= batch) return;
  const double* tab = io.tabs + (size_t)io.path_id[b] * io.J * PT_ROW;
  const double* st = io.state + (size_t)b * PS_SIZE;
  const int ec = io.ec ? io.ec[b] : 0;
  double* traj = io.traj + (size_t)b * C.N * TR_ROW;
  double* so = io.state_out + (size_t)b * PS_SIZE;
  post_instance(C, tab, io.sector[b], st, io.w + (size_t)b * C.n, ec, traj, so);
  if (io.ref)
    log_instance(C, tab, io.sector[b], st, io.p + (size_t)b * C.np, traj, ec, so + PS_PRREF, io.ref + (size_t)b * C.N * RF_ROW,
                 io.err + (size_t)b * C.N * ER_ROW);
}

// k_update: BoundMPC.update for a batch (bmpc_post.cuh update_state): new path, projected path-parameter state, restarted
// rotation reference, window back at the start of the path.
struct UpdateIO { const double* tabs; int J; const double* phimax; const int32_t* new_path; const double* cart; double* state; int32_t* sector; int32_t* path_id; };
__global__ void __launch_bounds__(PREP_THREADS) k_update(int batch, UpdateIO io) {
  const int b = blockIdx.x * PREP_THREADS + threadIdx.x;
  if (b >= batch) return;
  const int pth = io.new_path[b];
  if (pth < 0) return;                                            // this controller keeps its path
  update_state(io.tabs + (size_t)pth * io.J * PT_ROW, io.phimax[pth], io.cart + (size_t)b * 24, io.state + (size_t)b * PS_SIZE);
  io.sector[b] = 0;
  io.path_id[b] = pth;
}

// k_finish: second half of BoundMPC.step for a batch + closed-loop advance (bmpc_post.cuh).  Phase 1: one thread per instance
// decides which trajectory the controller keeps, post-processes it and writes the next-step state.  Phase 2: the CTA
// updates the previous solutions of its instances together (coalesced copy of x where the solve was accepted).
struct FinishIO {
  PostIO post;                 // post.w = solution x, post.ec = error count (in)
  const double* g; const int32_t* status;
  double* prev;                // [B, n] in/out
  int32_t* ec_out;             // may alias post.ec
  int advance;
};
__global__ void __launch_bounds__(PREP_THREADS) k_finish(const __grid_constant__ Config C, int batch, FinishIO io) {
  __shared__ unsigned char dec[PREP_THREADS];
  const int base = blockIdx.x * PREP_THREADS, b = base + threadIdx.x;
  const size_t n = C.n;
  if (b < batch) {
    const double* st = io.post.state + (size_t)b * PS_SIZE;
    const bool has_prev = st[PS_HASPREV] != 0.0;
    const int d = finish_decision(C, io.status[b], io.g + (size_t)b * C.m, has_prev);
    const int ec = d == 1 ? io.post.ec[b] + 1 : 0;
    const double* w = d == 1 ? io.prev + b * n : io.post.w + b * n;
    double* traj = io.post.traj + (size_t)b * C.N * TR_ROW;
    double* so = io.post.state_out + (size_t)b * PS_SIZE;
    if (ec < C.N) {
      const double* tab = io.post.tabs + (size_t)io.post.path_id[b] * io.post.J * PT_ROW;
      post_instance(C, tab, io.post.sector[b], st, w, ec, traj, so);
      if (io.post.ref)      // logging branch (ref_data / err_data of BoundMPC.step), before the robot advance rewrites the joint state
        log_instance(C, tab, io.post.sector[b], st, io.post.p + (size_t)b * C.np, traj, ec, so + PS_PRREF,
                     io.post.ref + (size_t)b * C.N * RF_ROW, io.post.err + (size_t)b * C.N * ER_ROW);
      if (io.advance) advance_state(w, ec, traj, so);
    } else {                                       // the controller gives up (BoundMPC.py:504-506): nothing to return
      for (int i = 0; i < C.N * TR_ROW; i++) traj[i] = 0.0;
      for (int i = 0; i < PS_SIZE; i++) so[i] = st[i];
    }
    if (d == 0) so[PS_HASPREV] = 1.0;
    io.ec_out[b] = ec;
    dec[threadIdx.x] = (unsigned char)d;
  }
  __syncthreads();
  const int cnt = batch - base < PREP_THREADS ? batch - base : PREP_THREADS;
  for (size_t idx = threadIdx.x; idx < (size_t)cnt * n; idx += PREP_THREADS) {
    const int i = (int)(idx / n);
    const size_t e = idx - (size_t)i * n, bi = (size_t)(base + i);
    if (dec[i] == 0) io.prev[bi * n + e] = io.post.w[bi * n + e];
  }
}

// k_kkt: one Newton step per instance at caller-supplied primal-dual points (parity tests, bmpc_eval.cuh)
struct KktBatchIO { const double* v; const double* p; const double* mu; const double* dw; double* dx; double* ynew; int32_t* ok; };
__global__ void __launch_bounds__(BMPC_MAX_THREADS) k_kkt(const __grid_constant__ Config C, int batch, KktBatchIO io, double* ws,
                                                          size_t ws_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
#ifdef BMPC_TIMING
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red, S.tm};
#else
  Ctx cx{(int)threadIdx.x, (int)blockDim.x, S.red};
#endif
  Work W;
  work_carve(W, ws + (size_t)blockIdx.x * ws_stride, C.N);
  work_attach_smem(W, S, C.N);
  build_tables(cx, C, S);
  phase_kin_jacobian_init(cx, C, W);
  const size_t n = C.n, ne = (size_t)NE * C.N, nv = 3 * n + ne + 2 * (size_t)ND * C.N;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    KktIO k{io.v + b * nv, io.p + (size_t)b * C.np, io.mu[b], io.dw[b], io.dx + b * n, io.ynew + b * ne, io.ok + b};
    kkt_step_instance(cx, C, W, S, k);
    __syncthreads();
  }
}

// FP64 pipe peak probes (roofline denominator).  Each thread runs 8 independent
// dependency chains so the DFMA / DMMA pipe is the only limiter.
__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0 - 1e-12, c = 1e-13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters) {
  double acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < 4; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[0] = s;
}

// ------------------------------------------------------------------------------------------ host side
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
#define CU(call)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) return fail(BMPC_E_CUDA, #call ": %s", cudaGetErrorString(e_)); \
  } while (0)

// Entry points run on the handle's device and put the caller's current device back when they return (a handle created for
// another device than the caller's current one must not change what torch / the caller sees as current).
struct DevGuard {
  int prev = -1;
  bool ok = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    if (prev == dev) { prev = -1; ok = true; return; }
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct bmpc_handle {
  Config C;
  int device, threads, sms, ctas_per_sm, variant;
  size_t l2_window_max, l2_persist_bytes;   // L2 access-policy window for the workspace (0: off)
  size_t smem_solve;   // dynamic shared memory of k_solve: sizeof(Smem) [+ the iterate of horizons above VEC_NMAX]
  int vec_in_ws;       // development switch BMPC_VEC_WS=1: iterate in the workspace even when it fits in shared memory
  int single_pass, no_zero_copy;   // development switches, read from the environment once in bmpc_create
  int variant_lat;     // launch shape for batches of at most one instance per SM (-1: none): more threads per instance
  size_t ws_stride;    // doubles per CTA slot
  int64_t launches;
  // cached device buffers of the host-pointer entry points
  void* dbuf; size_t dbuf_bytes;
  void* pin; size_t pin_bytes;      // page-locked staging area of the host entry for small pageable batches
  cudaStream_t stream;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int BMPC_STAGE_MAX = 64;   // largest pageable batch the host entry stages through its page-locked area

static int ensure_dbuf(bmpc_handle* h, size_t bytes) {
  if (h->dbuf_bytes >= bytes) return BMPC_OK;
  if (h->dbuf) { cudaFree(h->dbuf); h->dbuf = nullptr; h->dbuf_bytes = 0; }
  CU(cudaMalloc(&h->dbuf, bytes));
  h->dbuf_bytes = bytes;
  return BMPC_OK;
}

static int ensure_pin(bmpc_handle* h, size_t bytes) {
  if (h->pin_bytes >= bytes) return BMPC_OK;
  if (h->pin) { cudaFreeHost(h->pin); h->pin = nullptr; h->pin_bytes = 0; }
  CU(cudaHostAlloc(&h->pin, bytes, cudaHostAllocDefault));
  h->pin_bytes = bytes;
  return BMPC_OK;
}

static int grid_for(const bmpc_handle* h, int batch) {
  int g = h->sms * h->ctas_per_sm;
  return batch < g ? batch : g;
}

extern "C" {

const char* bmpc_last_error(void) { return g_err; }

int bmpc_create(const bmpc_config* cfg, bmpc_handle** out) {
  if (!cfg || !out) return fail(BMPC_E_INVALID, "bmpc_create: null argument");
  bmpc_handle* h = new (std::nothrow) bmpc_handle();
  if (!h) return fail(BMPC_E_NOMEM, "bmpc_create: out of memory");
  if (make_config(*cfg, h->C)) { delete h; return fail(BMPC_E_INVALID, "bmpc_create: invalid N / nr_segs / dt"); }
  int dev = cfg->device;
  if (dev < 0) { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) { delete h; return fail(BMPC_E_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); } }
  h->device = dev;
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) { delete h; return fail(BMPC_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { delete h; return fail(BMPC_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  h->sms = prop.multiProcessorCount;
  // launch shape: cfg->threads (or BMPC_THREADS) and BMPC_CTAS_PER_SM select the compiled variant
  const char* envt = getenv("BMPC_THREADS");
  const char* envc = getenv("BMPC_CTAS_PER_SM");
  int want_t = cfg->threads > 0 ? cfg->threads : (envt ? atoi(envt) : BMPC_DEFAULT_THREADS);
  int want_c = envc ? atoi(envc) : BMPC_DEFAULT_CTAS;
  h->variant = -1;
  for (int v = 0; v < kNumVariants; v++)
    if (kVariants[v].threads == want_t && kVariants[v].minb == want_c) h->variant = v;
  // fewer resident CTAs than a variant was compiled for: same kernel, smaller grid (occupancy experiments)
  for (int v = 0; v < kNumVariants && h->variant < 0; v++)
    if (kVariants[v].threads == want_t && kVariants[v].minb > want_c && want_c >= 1) h->variant = v;
  if (h->variant < 0) { delete h; return fail(BMPC_E_INVALID, "bmpc_create: no kernel variant for this threads / CTAs-per-SM combination"); }
  h->threads = want_t;
  // Latency shape: when every instance has an SM to itself (batch <= SM count) the phases with a few hundred parallel
  // items finish sooner with 384 threads (150 instead of 196 us per interior-point iteration); an explicit threads /
  // BMPC_THREADS request switches this off.
  h->variant_lat = -1;
  if (cfg->threads <= 0 && !envt && !getenv("BMPC_NO_LATENCY_SHAPE"))
    for (int v = 0; v < kNumVariants; v++)
      if (kVariants[v].threads == 384 && kVariants[v].minb == 1) h->variant_lat = v;
  e = cudaSuccess;
  // horizons above VEC_NMAX: the iterate (N x 192 doubles) stays in the workspace (L2) and three CTAs per SM run; with
  // BMPC_VEC_SMEM=1 it goes behind the struct (two CTAs per SM).  Measured on config 4 (N = 20, 8,192 instances): 9,977
  // solves/s from the workspace against 9,031 on chip -- the third resident CTA is worth more than the L2 round trips.
  h->smem_solve = sizeof(Smem);
  {
    const size_t ext = (size_t)h->C.N * (3 * NX + NE + 2 * ND) * sizeof(double);
    if (h->C.N > VEC_NMAX && getenv("BMPC_VEC_SMEM") && 2 * (sizeof(Smem) + ext + 1024) <= (size_t)prop.sharedMemPerMultiprocessor) h->smem_solve += ext;
  }
  if (h->variant_lat >= 0) e = cudaFuncSetAttribute(kVariants[h->variant_lat].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_solve);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kVariants[h->variant].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_solve);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kkt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  if (e != cudaSuccess) { delete h; return fail(BMPC_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kVariants[h->variant].fn, h->threads, h->smem_solve);
  if (e != cudaSuccess || occ < 1) { delete h; return fail(BMPC_E_CUDA, "occupancy query failed: %s", cudaGetErrorString(e)); }
  h->ctas_per_sm = want_c < occ ? want_c : occ;
  h->ws_stride = align_up(work_doubles(h->C.N), 32);
  h->launches = 0;
  // persisting-L2 carve-out for the workspace window (see solve_batch_impl): opt-in with BMPC_L2_WINDOW=1.  Measured on the
  // bench shard: 54.9 vs 55.6 ms in an interleaved A/B, 55.4 vs 55.0 ms under ncu, and MORE DRAM write-back (17.1 vs 12.3 GB
  // per launch: the 67 MB of slices exceed the persisting carve-out and everything else is demoted to streaming).
  h->l2_window_max = h->l2_persist_bytes = 0;
  if (getenv("BMPC_L2_WINDOW") && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
    size_t want = (size_t)h->sms * h->ctas_per_sm * h->ws_stride * sizeof(double);
    if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
      h->l2_persist_bytes = want;
      h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    } else cudaGetLastError();
  }
  h->vec_in_ws = getenv("BMPC_VEC_WS") != nullptr;
  h->single_pass = getenv("BMPC_SINGLE_PASS") != nullptr;
  h->no_zero_copy = getenv("BMPC_NO_ZERO_COPY") != nullptr;
  if (const char* e_ = getenv("BMPC_SLICE_ITERS")) { const int v_ = atoi(e_); if (v_ >= 1) h->C.slice_iters = v_; }
  if (const char* e_ = getenv("BMPC_MAX_SOC")) h->C.max_soc = atoi(e_) > 0 ? 1 : 0;
  if (const char* e_ = getenv("BMPC_QSS_LATE")) h->C.qss_late = atoi(e_) > 0 ? 1 : 0;
  if (const char* e_ = getenv("BMPC_HARD_CONTINUE")) h->C.hard_continue = atoi(e_) > 0 ? 1 : 0;
  h->dbuf = nullptr; h->dbuf_bytes = 0;
  h->pin = nullptr; h->pin_bytes = 0;
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return fail(BMPC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  *out = h;
  return BMPC_OK;
}

void bmpc_destroy(bmpc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->dbuf) cudaFree(h->dbuf);
  if (h->pin) cudaFreeHost(h->pin);
  cudaStreamDestroy(h->stream);
  delete h;
}

int bmpc_dims(const bmpc_handle* h, int32_t* n, int32_t* m, int32_t* np) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_dims: null handle");
  if (n) *n = h->C.n;
  if (m) *m = h->C.m;
  if (np) *np = h->C.np;
  return BMPC_OK;
}

int bmpc_bounds(const bmpc_handle* h, double* lbx, double* ubx, double* lbg, double* ubg) {
  if (!h || !lbx || !ubx || !lbg || !ubg) return fail(BMPC_E_INVALID, "bmpc_bounds: null argument");
  fill_bounds(h->C, lbx, ubx, lbg, ubg);
  return BMPC_OK;
}

int bmpc_workspace_bytes(const bmpc_handle* h, int32_t batch, size_t* bytes) {
  if (!h || !bytes || batch < 0) return fail(BMPC_E_INVALID, "bmpc_workspace_bytes: invalid argument");
  const int b = batch > 0 ? batch : 1, grid = grid_for(h, b);
  *bytes = 256 + (size_t)grid * h->ws_stride * sizeof(double);
  if (b > grid) *bytes += align_up((size_t)SCHED_LISTS * b * sizeof(int), 256) + (size_t)b * align_up(save_doubles(h->C.N), 32) * sizeof(double);   // two-pass scheduling
  return BMPC_OK;
}

int64_t bmpc_launch_count(const bmpc_handle* h) { return h ? h->launches : 0; }

int bmpc_launch_shape(const bmpc_handle* h, int32_t* threads, int32_t* ctas_per_sm, int32_t* smem_bytes, int32_t* sms) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_launch_shape: null handle");
  if (threads) *threads = h->threads;
  if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
  if (smem_bytes) *smem_bytes = (int32_t)h->smem_solve;
  if (sms) *sms = h->sms;
  return BMPC_OK;
}

int bmpc_fp64_peak(bmpc_handle* h, int32_t kind, double* flops_per_s) {
  if (!h || !flops_per_s || kind < 0 || kind > 1) return fail(BMPC_E_INVALID, "bmpc_fp64_peak: invalid argument");
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  int rc = ensure_dbuf(h, 256);
  if (rc) return rc;
  const int iters = 4096, grid = h->sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(e0, h->stream));
    if (kind == 0) k_peak_dfma<<<grid, threads, 0, h->stream>>>((double*)h->dbuf, iters);
    else k_peak_dmma<<<grid, threads, 0, h->stream>>>((double*)h->dbuf, iters);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    // DFMA: 2 flop x 64 per loop trip per thread; DMMA m8n8k4: 2*8*8*4 = 512 flop per warp instruction, 32 per trip per warp
    const double fl = kind == 0 ? 2.0 * 64 * iters * (double)grid * threads : 512.0 * 32 * iters * (double)grid * (threads / 32);
    if (rep > 0 && fl / (ms * 1e-3) > best) best = fl / (ms * 1e-3);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *flops_per_s = best;
  return BMPC_OK;
}

static int solve_batch_impl(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
                            double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err, void* workspace, void* cuda_stream,
                            const double* x0_src, const double* p_src);
int bmpc_solve_batch(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
                     double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err, void* workspace, void* cuda_stream) {
  return solve_batch_impl(h, batch, x0, p, x, g, lam_g, lam_x, f, iters, status, kkt_err, workspace, cuda_stream, nullptr, nullptr);
}
static int solve_batch_impl(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
                            double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err, void* workspace, void* cuda_stream,
                            const double* x0_src, const double* p_src) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_solve_batch: null handle");
  if (batch < 0 || !x0 || !p || !x || !g || !lam_g || !lam_x || !f || !iters || !status || !kkt_err || !workspace)
    return fail(BMPC_E_INVALID, "bmpc_solve_batch: null buffer");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int grid = grid_for(h, batch);
  // workspace: [queue 256 B][per-CTA slices][parked lists 2 x batch int][parked iterates batch x save_stride]
  char* wsb = (char*)workspace;
  double* ws = (double*)(wsb + 256);
  SchedMem M{(Sched*)wsb, nullptr, batch, nullptr, 0, 0};
  CU(cudaMemsetAsync(wsb, 0, 256, st));
  if (batch > grid && !h->single_pass) {
    const size_t o_list = 256 + (size_t)grid * h->ws_stride * sizeof(double), list_bytes = align_up((size_t)SCHED_LISTS * batch * sizeof(int), 256);
    M.list = (int*)(wsb + o_list);
    M.save = (double*)(wsb + o_list + list_bytes);
    M.save_stride = align_up(save_doubles(h->C.N), 32);
    M.two_pass = 1;
    CU(cudaMemsetAsync(M.list, 0xFF, (size_t)SCHED_LISTS * batch * sizeof(int), st));
  }
  BatchIO io{x0, p, x, g, lam_g, lam_x, f, kkt_err, iters, status, x0_src, p_src};
  const int v = (h->variant_lat >= 0 && batch <= h->sms) ? h->variant_lat : h->variant;
  // The per-CTA workspace slices (150 KB each, 67 MB for a full grid) are the only data the kernel re-reads; inputs and
  // results stream through once.  An access-policy window marks the slices as persisting in L2 and everything else as
  // streaming, so the 176 MB of inputs / outputs of a large batch no longer evict (and write back) workspace lines.
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(kVariants[v].threads); lc.dynamicSmemBytes = h->smem_solve; lc.stream = st;
  cudaLaunchAttribute at[1];
  int nat = 0;
  if (h->l2_window_max > 0 && batch > grid) {
    size_t wbytes = (size_t)grid * h->ws_stride * sizeof(double);
    if (wbytes > h->l2_window_max) wbytes = h->l2_window_max;
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = ws;
    at[0].val.accessPolicyWindow.num_bytes = wbytes;
    at[0].val.accessPolicyWindow.hitRatio = h->l2_persist_bytes >= wbytes ? 1.0f : (float)h->l2_persist_bytes / (float)wbytes;
    at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    nat = 1;
  }
  lc.attrs = at; lc.numAttrs = nat;
  const int vec_ext = h->smem_solve > sizeof(Smem) ? 1 : (h->vec_in_ws ? -1 : 0);
  CU(cudaLaunchKernelEx(&lc, kVariants[v].fn, h->C, batch, io, ws, h->ws_stride, M, vec_ext));
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}


static bool is_mapped_host(const void* ptr) {
  cudaPointerAttributes a;
  if (!ptr) return false;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int bmpc_solve_batch_host(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g, double* lam_g,
                          double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_solve_batch_host: null handle");
  if (batch < 0 || !x0 || !p || !x || !iters || !status) return fail(BMPC_E_INVALID, "bmpc_solve_batch_host: null buffer");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  // Small batches in pageable memory (the single MPC step of a Python / ROS caller): ten small copies around the launch
  // cost more than the transfers themselves.  They are staged through a page-locked area of the handle, which the
  // kernel reads and writes itself (see below): two host memcpys and one launch.
  if (batch <= BMPC_STAGE_MAX && !h->no_zero_copy && !(is_mapped_host(x0) && is_mapped_host(x))) {
    const size_t B = batch, n = h->C.n, m = h->C.m, np = h->C.np;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_x0 = take(B * n * 8), o_p = take(B * np * 8), o_x = take(B * n * 8), o_g = take(B * m * 8), o_lg = take(B * m * 8),
                 o_lx = take(B * n * 8), o_f = take(B * 8), o_k = take(B * 8), o_it = take(B * 4), o_st = take(B * 4);
    int rc = ensure_pin(h, off);
    if (rc) return rc;
    char* q = (char*)h->pin;
    memcpy(q + o_x0, x0, B * n * 8);
    memcpy(q + o_p, p, B * np * 8);
    rc = bmpc_solve_batch_host(h, batch, (const double*)(q + o_x0), (const double*)(q + o_p), (double*)(q + o_x), (double*)(q + o_g),
                               (double*)(q + o_lg), (double*)(q + o_lx), (double*)(q + o_f), (int32_t*)(q + o_it), (int32_t*)(q + o_st),
                               (double*)(q + o_k));
    if (rc) return rc;
    memcpy(x, q + o_x, B * n * 8);
    if (g) memcpy(g, q + o_g, B * m * 8);
    if (lam_g) memcpy(lam_g, q + o_lg, B * m * 8);
    if (lam_x) memcpy(lam_x, q + o_lx, B * n * 8);
    if (f) memcpy(f, q + o_f, B * 8);
    if (kkt_err) memcpy(kkt_err, q + o_k, B * 8);
    memcpy(iters, q + o_it, B * 4);
    memcpy(status, q + o_st, B * 4);
    return BMPC_OK;
  }
  const size_t B = batch, n = h->C.n, m = h->C.m, np = h->C.np;
  size_t wsb = 0;
  bmpc_workspace_bytes(h, batch, &wsb);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_x0 = take(B * n * 8), o_p = take(B * np * 8), o_x = take(B * n * 8), o_g = take(B * m * 8), o_lg = take(B * m * 8),
               o_lx = take(B * n * 8), o_f = take(B * 8), o_k = take(B * 8), o_it = take(B * 4), o_st = take(B * 4), o_ws = take(wsb);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  cudaStream_t st = h->stream;
  // Results.  A result buffer in page-locked host memory the device can address (cudaHostAlloc / cudaHostRegister,
  // e.g. a pinned torch tensor) is written by the kernel itself: every instance stores its 14 KB of results when its
  // solve ends, so the transfer runs under the launch instead of after it (114 MB per 8,192 instances otherwise).
  // Anything else gets a device buffer and a copy.  Page-locked inputs are handled the same way from the other side:
  // the CTA that starts an instance fetches its x0 / p (7.5 KB) into the device buffers, which it then reads throughout
  // the solve; pageable inputs are copied before the launch.
  auto mapped = [&](const void* host) -> void* {
    if (!host || h->no_zero_copy) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
  };
  void *m_x = mapped(x), *m_g = mapped(g), *m_lg = mapped(lam_g), *m_lx = mapped(lam_x), *m_f = mapped(f), *m_k = mapped(kkt_err),
       *m_it = mapped(iters), *m_st = mapped(status);
  const void *s_x0 = mapped(x0), *s_p = mapped(p);
  if (!s_x0 || !s_p) {
    s_x0 = s_p = nullptr;
    CU(cudaMemcpyAsync(d + o_x0, x0, B * n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_p, p, B * np * 8, cudaMemcpyHostToDevice, st));
  }
  rc = solve_batch_impl(h, batch, (double*)(d + o_x0), (double*)(d + o_p), m_x ? (double*)m_x : (double*)(d + o_x),
                        m_g ? (double*)m_g : (double*)(d + o_g), m_lg ? (double*)m_lg : (double*)(d + o_lg),
                        m_lx ? (double*)m_lx : (double*)(d + o_lx), m_f ? (double*)m_f : (double*)(d + o_f),
                        m_it ? (int32_t*)m_it : (int32_t*)(d + o_it), m_st ? (int32_t*)m_st : (int32_t*)(d + o_st),
                        m_k ? (double*)m_k : (double*)(d + o_k), d + o_ws, st, (const double*)s_x0, (const double*)s_p);
  if (rc) return rc;
  if (!m_x) CU(cudaMemcpyAsync(x, d + o_x, B * n * 8, cudaMemcpyDeviceToHost, st));
  if (g && !m_g) CU(cudaMemcpyAsync(g, d + o_g, B * m * 8, cudaMemcpyDeviceToHost, st));
  if (lam_g && !m_lg) CU(cudaMemcpyAsync(lam_g, d + o_lg, B * m * 8, cudaMemcpyDeviceToHost, st));
  if (lam_x && !m_lx) CU(cudaMemcpyAsync(lam_x, d + o_lx, B * n * 8, cudaMemcpyDeviceToHost, st));
  if (f && !m_f) CU(cudaMemcpyAsync(f, d + o_f, B * 8, cudaMemcpyDeviceToHost, st));
  if (kkt_err && !m_k) CU(cudaMemcpyAsync(kkt_err, d + o_k, B * 8, cudaMemcpyDeviceToHost, st));
  if (!m_it) CU(cudaMemcpyAsync(iters, d + o_it, B * 4, cudaMemcpyDeviceToHost, st));
  if (!m_st) CU(cudaMemcpyAsync(status, d + o_st, B * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return BMPC_OK;
}

int bmpc_prepare_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                       int32_t* sector, const double* state, const double* prev_x, double* x0, double* p, void* cuda_stream) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_prepare_batch: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < h->C.S || !path_tables || !path_id || !sector || !state || !prev_x || !x0 || !p)
    return fail(BMPC_E_INVALID, "bmpc_prepare_batch: invalid argument");
  if (h->C.S > PREP_SMAX) return fail(BMPC_E_INVALID, "bmpc_prepare_batch: nr_segs above the builder's limit");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  PrepIO io{path_tables, path_rows, path_id, sector, state, prev_x, x0, p};
  k_prepare<<<(batch + PREP_THREADS - 1) / PREP_THREADS, PREP_THREADS, 0, (cudaStream_t)cuda_stream>>>(h->C, batch, io);
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_prepare_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                            const int32_t* path_id, int32_t* sector, const double* state, const double* prev_x, double* x0, double* p) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_prepare_batch_host: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < h->C.S || !path_tables || !path_id || !sector || !state || !prev_x || !x0 || !p)
    return fail(BMPC_E_INVALID, "bmpc_prepare_batch_host: invalid argument");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  const size_t B = batch, n = h->C.n, np = h->C.np, tb = (size_t)n_paths * path_rows * PT_ROW * 8;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_t = take(tb), o_id = take(B * 4), o_sec = take(B * 4), o_st = take(B * PS_SIZE * 8), o_prev = take(B * n * 8),
               o_x0 = take(B * n * 8), o_p = take(B * np * 8);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  cudaStream_t st = h->stream;
  CU(cudaMemcpyAsync(d + o_t, path_tables, tb, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_id, path_id, B * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_sec, sector, B * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_st, state, B * PS_SIZE * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_prev, prev_x, B * n * 8, cudaMemcpyHostToDevice, st));
  rc = bmpc_prepare_batch(h, batch, (double*)(d + o_t), n_paths, path_rows, (int32_t*)(d + o_id), (int32_t*)(d + o_sec), (double*)(d + o_st),
                          (double*)(d + o_prev), (double*)(d + o_x0), (double*)(d + o_p), st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(sector, d + o_sec, B * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(x0, d + o_x0, B * n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(p, d + o_p, B * np * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return BMPC_OK;
}

int bmpc_post_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                    const int32_t* sector, const double* state, const double* w, const int32_t* error_count, double* traj,
                    double* state_out, void* cuda_stream) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_post_batch: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 2 || !path_tables || !path_id || !sector || !state || !w || !traj || !state_out)
    return fail(BMPC_E_INVALID, "bmpc_post_batch: invalid argument");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  PostIO io{path_tables, path_rows, path_id, sector, state, w, error_count, traj, state_out, nullptr, nullptr, nullptr};
  k_post<<<(batch + PREP_THREADS - 1) / PREP_THREADS, PREP_THREADS, 0, (cudaStream_t)cuda_stream>>>(h->C, batch, io);
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_update_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const double* path_phi_max,
                      const int32_t* new_path, const double* cart, double* state, int32_t* sector, int32_t* path_id, void* cuda_stream) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_update_batch: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 1 || !path_tables || !path_phi_max || !new_path || !cart || !state || !sector || !path_id)
    return fail(BMPC_E_INVALID, "bmpc_update_batch: invalid argument");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  UpdateIO io{path_tables, path_rows, path_phi_max, new_path, cart, state, sector, path_id};
  k_update<<<(batch + PREP_THREADS - 1) / PREP_THREADS, PREP_THREADS, 0, (cudaStream_t)cuda_stream>>>(batch, io);
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_post_log_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                        const int32_t* sector, const double* state, const double* p, const double* w, const int32_t* error_count, double* traj,
                        double* state_out, double* ref, double* err, void* cuda_stream) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_post_log_batch: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 3 || !path_tables || !path_id || !sector || !state || !p || !w || !traj || !state_out || !ref || !err)
    return fail(BMPC_E_INVALID, "bmpc_post_log_batch: invalid argument");
  if (h->C.S < 3) return fail(BMPC_E_INVALID, "bmpc_post_log_batch: the rotation reference of the logging branch needs nr_segs >= 3");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  PostIO io{path_tables, path_rows, path_id, sector, state, w, error_count, traj, state_out, p, ref, err};
  k_post<<<(batch + PREP_THREADS - 1) / PREP_THREADS, PREP_THREADS, 0, (cudaStream_t)cuda_stream>>>(h->C, batch, io);
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

static int finish_impl(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                       const int32_t* sector, const double* state, const double* x, const double* g, const int32_t* status, double* prev_x,
                       int32_t* error_count, double* traj, double* state_out, int32_t advance, const double* p, double* ref, double* err,
                       void* cuda_stream) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_finish_batch: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 2 || !path_tables || !path_id || !sector || !state || !x || !g || !status || !prev_x ||
      !error_count || !traj || !state_out)
    return fail(BMPC_E_INVALID, "bmpc_finish_batch: invalid argument");
  if (ref && (h->C.S < 3 || path_rows < 3 || !p || !err)) return fail(BMPC_E_INVALID, "bmpc_finish_batch: the logging branch needs p, err and nr_segs >= 3");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  FinishIO io{{path_tables, path_rows, path_id, sector, state, x, error_count, traj, state_out, ref ? p : nullptr, ref, ref ? err : nullptr}, g, status,
              prev_x, error_count, advance};
  k_finish<<<(batch + PREP_THREADS - 1) / PREP_THREADS, PREP_THREADS, 0, (cudaStream_t)cuda_stream>>>(h->C, batch, io);
  CU(cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_finish_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                      const int32_t* sector, const double* state, const double* x, const double* g, const int32_t* status, double* prev_x,
                      int32_t* error_count, double* traj, double* state_out, int32_t advance, void* cuda_stream) {
  return finish_impl(h, batch, path_tables, n_paths, path_rows, path_id, sector, state, x, g, status, prev_x, error_count, traj, state_out, advance,
                     nullptr, nullptr, nullptr, cuda_stream);
}

// One whole MPC step for `batch` controllers held on the host: k_prepare -> k_solve -> k_finish (+ logging branch) on the
// handle's stream, inputs and outputs staged through ONE page-locked block each way.
int bmpc_mpc_step_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                             const int32_t* path_id, int32_t* sector, const double* state, double* prev_x, int32_t* error_count,
                             double* x, double* traj, double* state_out, double* ref, double* err, int32_t* iters, int32_t* status) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_mpc_step_batch_host: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 2 || !path_tables || !path_id || !sector || !state || !prev_x || !error_count || !x || !traj ||
      !state_out || !iters || !status || (ref && !err))
    return fail(BMPC_E_INVALID, "bmpc_mpc_step_batch_host: invalid argument");
  if (h->C.S > PREP_SMAX) return fail(BMPC_E_INVALID, "bmpc_mpc_step_batch_host: nr_segs above the builder's limit");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  const size_t B = batch, n = h->C.n, m = h->C.m, np = h->C.np, N = h->C.N, tb = (size_t)n_paths * path_rows * PT_ROW * 8;
  size_t wsb = 0;
  bmpc_workspace_bytes(h, batch, &wsb);
  // block layout, identical in the page-locked staging area and on the device: [inputs | outputs | device-only scratch]
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_t = take(tb), o_id = take(B * 4), o_st = take(B * PS_SIZE * 8);
  const size_t o_sec = take(B * 4), o_ec = take(B * 4), o_prev = take(B * n * 8);           // in / out
  const size_t in_end = off;
  const size_t o_x = take(B * n * 8), o_tr = take(B * N * TR_ROW * 8), o_so = take(B * PS_SIZE * 8), o_rf = take(ref ? B * N * RF_ROW * 8 : 8),
               o_er = take(ref ? B * N * ER_ROW * 8 : 8), o_it = take(B * 4), o_stt = take(B * 4);
  const size_t io_end = off;
  const size_t o_x0 = take(B * n * 8), o_p = take(B * np * 8), o_g = take(B * m * 8), o_lg = take(B * m * 8), o_lx = take(B * n * 8), o_f = take(B * 8),
               o_k = take(B * 8), o_ws = take(wsb);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  rc = ensure_pin(h, io_end);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  char* q = (char*)h->pin;
  cudaStream_t st = h->stream;
  memcpy(q + o_t, path_tables, tb);
  memcpy(q + o_id, path_id, B * 4);
  memcpy(q + o_st, state, B * PS_SIZE * 8);
  memcpy(q + o_sec, sector, B * 4);
  memcpy(q + o_ec, error_count, B * 4);
  memcpy(q + o_prev, prev_x, B * n * 8);
  CU(cudaMemcpyAsync(d, q, in_end, cudaMemcpyHostToDevice, st));
  rc = bmpc_prepare_batch(h, batch, (double*)(d + o_t), n_paths, path_rows, (int32_t*)(d + o_id), (int32_t*)(d + o_sec), (double*)(d + o_st),
                          (double*)(d + o_prev), (double*)(d + o_x0), (double*)(d + o_p), st);
  if (rc) return rc;
  rc = solve_batch_impl(h, batch, (double*)(d + o_x0), (double*)(d + o_p), (double*)(d + o_x), (double*)(d + o_g), (double*)(d + o_lg),
                        (double*)(d + o_lx), (double*)(d + o_f), (int32_t*)(d + o_it), (int32_t*)(d + o_stt), (double*)(d + o_k), d + o_ws, st,
                        nullptr, nullptr);
  if (rc) return rc;
  rc = finish_impl(h, batch, (double*)(d + o_t), n_paths, path_rows, (int32_t*)(d + o_id), (int32_t*)(d + o_sec), (double*)(d + o_st),
                   (double*)(d + o_x), (double*)(d + o_g), (int32_t*)(d + o_stt), (double*)(d + o_prev), (int32_t*)(d + o_ec), (double*)(d + o_tr),
                   (double*)(d + o_so), 0, (double*)(d + o_p), ref ? (double*)(d + o_rf) : nullptr, ref ? (double*)(d + o_er) : nullptr, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(q + o_sec, d + o_sec, io_end - o_sec, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(sector, q + o_sec, B * 4);
  memcpy(error_count, q + o_ec, B * 4);
  memcpy(prev_x, q + o_prev, B * n * 8);
  memcpy(x, q + o_x, B * n * 8);
  memcpy(traj, q + o_tr, B * N * TR_ROW * 8);
  memcpy(state_out, q + o_so, B * PS_SIZE * 8);
  if (ref) { memcpy(ref, q + o_rf, B * N * RF_ROW * 8); memcpy(err, q + o_er, B * N * ER_ROW * 8); }
  memcpy(iters, q + o_it, B * 4);
  memcpy(status, q + o_stt, B * 4);
  return BMPC_OK;
}

int bmpc_post_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows, const int32_t* path_id,
                         const int32_t* sector, const double* state, const double* w, const int32_t* error_count, double* traj,
                         double* state_out) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_post_batch_host: null handle");
  if (batch < 0 || n_paths < 1 || path_rows < 2 || !path_tables || !path_id || !sector || !state || !w || !traj || !state_out)
    return fail(BMPC_E_INVALID, "bmpc_post_batch_host: invalid argument");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  const size_t B = batch, n = h->C.n, tr = (size_t)h->C.N * TR_ROW, tb = (size_t)n_paths * path_rows * PT_ROW * 8;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_t = take(tb), o_id = take(B * 4), o_sec = take(B * 4), o_ec = take(B * 4), o_st = take(B * PS_SIZE * 8), o_w = take(B * n * 8),
               o_tr = take(B * tr * 8), o_so = take(B * PS_SIZE * 8);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  cudaStream_t st = h->stream;
  CU(cudaMemcpyAsync(d + o_t, path_tables, tb, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_id, path_id, B * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_sec, sector, B * 4, cudaMemcpyHostToDevice, st));
  if (error_count) CU(cudaMemcpyAsync(d + o_ec, error_count, B * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_st, state, B * PS_SIZE * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_w, w, B * n * 8, cudaMemcpyHostToDevice, st));
  rc = bmpc_post_batch(h, batch, (double*)(d + o_t), n_paths, path_rows, (int32_t*)(d + o_id), (int32_t*)(d + o_sec), (double*)(d + o_st),
                       (double*)(d + o_w), error_count ? (int32_t*)(d + o_ec) : nullptr, (double*)(d + o_tr), (double*)(d + o_so), st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(traj, d + o_tr, B * tr * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(state_out, d + o_so, B * PS_SIZE * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return BMPC_OK;
}

int bmpc_eval_batch_host(bmpc_handle* h, int32_t batch, const double* x, const double* p, const double* lam, double* f, double* g,
                         double* d_out, double* grad, double* jac, double* hess) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_eval_batch_host: null handle");
  if (batch < 0 || !x || !p) return fail(BMPC_E_INVALID, "bmpc_eval_batch_host: null buffer");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  const size_t B = batch, n = h->C.n, m = h->C.m, np = h->C.np, nl = (size_t)(NE + ND) * h->C.N, ndd = (size_t)ND * h->C.N;
  const int grid = grid_for(h, batch);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_x = take(B * n * 8), o_p = take(B * np * 8), o_l = take(B * nl * 8), o_f = take(B * 8), o_g = take(B * m * 8),
               o_d = take(B * ndd * 8), o_gr = take(B * n * 8), o_j = take(jac ? B * nl * n * 8 : 8), o_h = take(hess ? B * n * n * 8 : 8),
               o_ws = take((size_t)grid * h->ws_stride * 8);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  cudaStream_t st = h->stream;
  CU(cudaMemcpyAsync(d + o_x, x, B * n * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_p, p, B * np * 8, cudaMemcpyHostToDevice, st));
  if (lam) CU(cudaMemcpyAsync(d + o_l, lam, B * nl * 8, cudaMemcpyHostToDevice, st));
  EvalBatchIO io{(double*)(d + o_x), (double*)(d + o_p), lam ? (double*)(d + o_l) : nullptr, (double*)(d + o_f), (double*)(d + o_g),
                 (double*)(d + o_d), (double*)(d + o_gr), jac ? (double*)(d + o_j) : nullptr, hess ? (double*)(d + o_h) : nullptr};
  k_eval<<<grid, h->threads < BMPC_MAX_THREADS ? h->threads : BMPC_MAX_THREADS, sizeof(Smem), st>>>(h->C, batch, io, (double*)(d + o_ws), h->ws_stride);
  CU(cudaGetLastError());
  h->launches += 1;
  if (f) CU(cudaMemcpyAsync(f, d + o_f, B * 8, cudaMemcpyDeviceToHost, st));
  if (g) CU(cudaMemcpyAsync(g, d + o_g, B * m * 8, cudaMemcpyDeviceToHost, st));
  if (d_out) CU(cudaMemcpyAsync(d_out, d + o_d, B * ndd * 8, cudaMemcpyDeviceToHost, st));
  if (grad) CU(cudaMemcpyAsync(grad, d + o_gr, B * n * 8, cudaMemcpyDeviceToHost, st));
  if (jac) CU(cudaMemcpyAsync(jac, d + o_j, B * nl * n * 8, cudaMemcpyDeviceToHost, st));
  if (hess) CU(cudaMemcpyAsync(hess, d + o_h, B * n * n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return BMPC_OK;
}

int bmpc_kkt_step_batch_host(bmpc_handle* h, int32_t batch, const double* v, const double* p, const double* mu, const double* delta_w,
                             double* dx, double* ynew, int32_t* ok) {
  if (!h) return fail(BMPC_E_INVALID, "bmpc_kkt_step_batch_host: null handle");
  if (batch < 0 || !v || !p || !mu || !delta_w || !dx || !ynew || !ok) return fail(BMPC_E_INVALID, "bmpc_kkt_step_batch_host: null buffer");
  if (batch == 0) return BMPC_OK;
  DevGuard dg_(h->device);
  if (!dg_.ok) return fail(BMPC_E_CUDA, "cudaSetDevice failed");
  const size_t B = batch, n = h->C.n, ne = (size_t)NE * h->C.N, nv = 3 * n + ne + 2 * (size_t)ND * h->C.N, np = h->C.np;
  const int grid = grid_for(h, batch);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_v = take(B * nv * 8), o_p = take(B * np * 8), o_mu = take(B * 8), o_dw = take(B * 8), o_dx = take(B * n * 8),
               o_y = take(B * ne * 8), o_ok = take(B * 4), o_ws = take((size_t)grid * h->ws_stride * 8);
  int rc = ensure_dbuf(h, off);
  if (rc) return rc;
  char* d = (char*)h->dbuf;
  cudaStream_t st = h->stream;
  CU(cudaMemcpyAsync(d + o_v, v, B * nv * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_p, p, B * np * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_mu, mu, B * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d + o_dw, delta_w, B * 8, cudaMemcpyHostToDevice, st));
  KktBatchIO io{(double*)(d + o_v), (double*)(d + o_p), (double*)(d + o_mu), (double*)(d + o_dw), (double*)(d + o_dx), (double*)(d + o_y),
                (int32_t*)(d + o_ok)};
  const int thr = h->threads < BMPC_MAX_THREADS ? h->threads : BMPC_MAX_THREADS;
  k_kkt<<<grid, thr, sizeof(Smem), st>>>(h->C, batch, io, (double*)(d + o_ws), h->ws_stride);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaMemcpyAsync(dx, d + o_dx, B * n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ynew, d + o_y, B * ne * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ok, d + o_ok, B * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return BMPC_OK;
}

}  // extern "C"
