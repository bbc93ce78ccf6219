/* boundmpc_b200 — C ABI of the B200-native batched solver for BoundMPC's per-step OCP.
 *
 * This is the drop-in boundary for the hot path of Thieso/BoundMPC: the handle replaces the
 * object returned by `casadi.nlpsol('solver', 'ipopt', prob, opts)`
 * (bound_mpc/bound_mpc/BoundMPC/casadi_ocp_formulation.py:389, obtained in BoundMPC.py:150-161)
 * and `bmpc_solve_batch*` replaces the call
 *     sol = self.solver(x0=w0, lbx=self.lbu, ubx=self.ubu, lbg=self.lbg, ubg=self.ubg, p=params)
 *     stats = self.solver.stats()
 * (BoundMPC.py:446-457) for a batch of independent instances.  Plain C types only; the
 * caller owns every buffer; nothing is allocated per call by the device-pointer entry.
 * All functions return 0 on success and a negative code otherwise (never throw);
 * `bmpc_last_error()` describes the last failure of the calling thread.
 *
 * Layouts (row-major, fp64): x0, x, lam_x: [batch, n], n = 44 N   (casadi_ocp_formulation.py:90-153)
 *                            p: [batch, np], np = 141 + 91 nr_segs (casadi_ocp_formulation.py:361-376)
 *                            g, lam_g: [batch, m], m = 43 N        (casadi_ocp_formulation.py:271-349)
 * Sign conventions follow CasADi: L = f + lam_g . g + lam_x . x.
 */
#ifndef BOUNDMPC_B200_H
#define BOUNDMPC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bmpc_handle bmpc_handle;

/* Arguments of `setup_optimization_problem(N, nr_joints, nr_segs, dt, u_min, u_max, ut_min, ut_max,
 * q_lim_lower, q_lim_upper, dq_lim_lower, dq_lim_upper, solver_opts)` (casadi_ocp_formulation.py:9-11)
 * plus the Ipopt options the reference sets (BoundMPC.py:120-141).  nr_joints is fixed to 7. */
typedef struct bmpc_config {
  int32_t N;            /* horizon length (params.n) */
  int32_t nr_segs;      /* path segments in the window */
  double dt;            /* sample time */
  double u_min, u_max;  /* joint jerk bounds */
  double ut_min, ut_max;/* path-parameter jerk bounds */
  double q_lim_lower[7], q_lim_upper[7];
  double dq_lim_lower[7], dq_lim_upper[7];
  double tol;           /* termination tolerance on Ipopt's scaled error (<= 0: 1e-9) */
  int32_t max_iter;     /* <= 0: 500 (BoundMPC.py:122) */
  double mu_init;       /* initial barrier parameter, <= 0: 1e-3 (Ipopt warm_start_mult_bound_push) */
  double bound_push;    /* <= 0: 1e-3 (Ipopt warm_start_bound_push) */
  int32_t device;       /* CUDA device ordinal, -1: current device */
  int32_t threads;      /* threads per CTA, <= 0: default */
} bmpc_config;

/* status codes written to `status[]` (stats()['success'] == (status == 0), BoundMPC.py:456-465) */
#define BMPC_STATUS_SUCCESS 0
#define BMPC_STATUS_MAXITER 1
#define BMPC_STATUS_LINESEARCH 2
#define BMPC_STATUS_REGULARIZATION 3
#define BMPC_STATUS_NUMERIC 4
#define BMPC_STATUS_DIVERGING 5   /* multipliers diverge (dual infeasibility > 1e7): locally infeasible instance */

/* error codes */
#define BMPC_OK 0
#define BMPC_E_INVALID (-1)
#define BMPC_E_CUDA (-2)
#define BMPC_E_NOMEM (-3)

/* replaces: ca.nlpsol(...) at casadi_ocp_formulation.py:389 */
int bmpc_create(const bmpc_config* cfg, bmpc_handle** out);
void bmpc_destroy(bmpc_handle* h);

/* n, m, np of the NLP (len(lbu), len(lbg), params.shape) */
int bmpc_dims(const bmpc_handle* h, int32_t* n, int32_t* m, int32_t* np);

/* replaces: the lbu, ubu, lbg, ubg lists returned by setup_optimization_problem
 * (casadi_ocp_formulation.py:384-391).  Host arrays of length n, n, m, m. */
int bmpc_bounds(const bmpc_handle* h, double* lbx, double* ubx, double* lbg, double* ubg);

/* bytes of device scratch `bmpc_solve_batch` needs for `batch` instances */
int bmpc_workspace_bytes(const bmpc_handle* h, int32_t batch, size_t* bytes);

/* replaces: self.solver(x0=..., p=...) + self.solver.stats() (BoundMPC.py:446-457), batched.
 * All data pointers are DEVICE pointers; `workspace` is device memory of at least
 * bmpc_workspace_bytes(); `cuda_stream` is a cudaStream_t (NULL: default stream).
 * Asynchronous with respect to the host.  f, iters, status, kkt_err may not be NULL. */
int bmpc_solve_batch(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g,
                     double* lam_g, double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err,
                     void* workspace, void* cuda_stream);

/* Same call with HOST pointers (what a Python / ROS caller holds): copies x0 and p to the
 * device, solves, copies the results back and synchronises.  Device buffers are owned and
 * cached by the handle.  Any output pointer except x, iters, status may be NULL.
 * A result buffer in page-locked host memory the device can address (cudaHostAlloc /
 * cudaHostRegister, a pinned torch tensor) is written by the kernel directly, instance by
 * instance as the solves end, instead of being copied after the launch; page-locked x0 / p
 * are fetched by the kernel when an instance starts instead of being copied before it; batches
 * of at most 64 instances in pageable memory are staged through a page-locked area owned by
 * the handle (same results; BMPC_NO_ZERO_COPY=1 in the environment forces the copies). */
int bmpc_solve_batch_host(bmpc_handle* h, int32_t batch, const double* x0, const double* p, double* x, double* g,
                          double* lam_g, double* lam_x, double* f, int32_t* iters, int32_t* status, double* kkt_err);

/* Batched parameter builder: replaces the pre-solve half of `BoundMPC.step` (BoundMPC.py:310-443 — the window of
 * `ReferencePath.get_parameters / get_limits / get_bound_params` (ReferencePath.py:190-238), `compute_initial_rot_errors`
 * (utils/util_functions.py:11-31), `compute_orientation_projection_vectors` (BoundMPC.py:267-304), `compute_error_bounds`
 * (BoundMPC.py:219-265), the warm-start shift (:316-333,373-375) and the parameter order of :416-443) for `batch`
 * controller instances; its outputs x0, p are the inputs of bmpc_solve_batch.  DEVICE pointers, asynchronous.
 *   path_tables [n_paths, path_rows, 41]  one row per padded path segment (layout PT_* in csrc/bmpc_prepare.cuh; built once
 *                                         per path by boundmpc_b200.reference_path.ReferencePath.path_table())
 *   path_id     [batch]                   path of each instance
 *   sector      [batch]  in/out           window position (ReferencePath.sector); advanced like ReferencePath.update
 *   state       [batch, 76]               controller state (layout PS_*: q0, dq0, ddq0, p0, v0, jerk, phi state, pr_ref,
 *                                         iw_ref, x_phi_d, bound factors, phi_max, weights, has_prev)
 *   prev_x      [batch, n]                previous solution (read when has_prev != 0)
 *   x0 [batch, n], p [batch, np]          outputs
 * States with state[74] != 0 (set by bmpc_update_batch) get the re-projected warm start of BoundMPC.py:335-369. */
int bmpc_prepare_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                       const int32_t* path_id, int32_t* sector, const double* state, const double* prev_x,
                       double* x0, double* p, void* cuda_stream);
/* Same with HOST pointers (copies inside, synchronises). */
int bmpc_prepare_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                            const int32_t* path_id, int32_t* sector, const double* state, const double* prev_x,
                            double* x0, double* p);

/* Batched post-processing: replaces `compute_return_data` of `BoundMPC.step` (BoundMPC.py:508-611,757-770, without the
 * logging branch :614-755) for `batch` controller instances: re-integration of the joint / path-parameter trajectory from
 * the jerks (jerk_trajectory_casadi.py:46-175), Cartesian pose / velocity / acceleration of every remaining horizon node
 * (RobotModel.forward_kinematics, BoundMPC.py:563-579) and the controller state of the next step (path-parameter state,
 * pr_ref, iw_ref: BoundMPC.py:593-611, utils/util_functions.py:88-99).  DEVICE pointers, asynchronous.
 *   path_tables, path_id, state           as for bmpc_prepare_batch (state of THIS step); sector = its output
 *   w           [batch, n]                the trajectory the controller keeps (solution x, or the previous solution
 *                                         after a failed solve, BoundMPC.py:467-496)
 *   error_count [batch] or NULL (= 0)     nodes of w already consumed (BoundMPC.error_count)
 *   traj        [batch, N, 42]            per node: p(6) v(6) a(6) q(7) dq(7) ddq(7) phi dphi ddphi; rows >= N - error_count zero
 *                                         (the jerk entries of traj_data are the columns u, u_phi of w)
 *   state_out   [batch, 76]               state with phi .. dddphi, pr_ref, iw_ref advanced; the joint state is the caller's */
int bmpc_post_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                    const int32_t* path_id, const int32_t* sector, const double* state, const double* w,
                    const int32_t* error_count, double* traj, double* state_out, void* cuda_stream);
/* Replanning: replaces `BoundMPC.update` (BoundMPC.py:163-217) for a batch.  Controller b with new_path[b] >= 0 moves to
 * that path (tables as for bmpc_prepare_batch, path_phi_max [n_paths] = ReferencePath.phi_max): its path-parameter state is
 * the projection of the measured Cartesian state cart [batch, 24] = (pose(6), velocity(6), acceleration(6), jerk(6)) on the
 * first segment, the rotation reference restarts at the first via point, the window at the start of the path, and
 * state[74] (updated) is set: from then on bmpc_prepare_batch re-projects the previous solution on the new path
 * (BoundMPC.py:335-369) instead of shifting it.  new_path[b] < 0: controller b is left alone.  DEVICE pointers. */
int bmpc_update_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                      const double* path_phi_max, const int32_t* new_path, const double* cart, double* state,
                      int32_t* sector, int32_t* path_id, void* cuda_stream);

/* bmpc_post_batch plus the logging branch of compute_return_data (BoundMPC.py:614-755, what `step` returns as ref_data and
 * err_data when params.real_time is False): per node the reference pose / bounds / bases (`reference_function`,
 * bound_mpc_functions.py:43-149) and the error terms (`error_function`, :152-202, with the exact rotation error along the
 * horizon of BoundMPC.py:716-752).  p [batch, np] = the parameter vectors of this step (output of bmpc_prepare_batch).
 *   ref [batch, N, 55]: p_d(6) dp_d(6) ddp_d(6) dp_normed(3) r_par_bound bound_lower(4) bound_upper(4) e_p_off(2) e_r_off(2)
 *                       bp1 bp2 br1 br2 v1 v2 v3 (3 each)
 *   err [batch, N, 33]: e_p de_p e_p_par e_p_orth de_p_par de_p_orth e_r de_r e_r_par e_r_orth1 e_r_orth2 (3 each)  */
int bmpc_post_log_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                        const int32_t* path_id, const int32_t* sector, const double* state, const double* p, const double* w,
                        const int32_t* error_count, double* traj, double* state_out, double* ref, double* err, void* cuda_stream);

/* Second half of `BoundMPC.step` as a whole (BoundMPC.py:454-506) for a batch, optionally with the closed-loop advance of
 * bound_mpc_node.py:321-331,362 — the step of an on-device roll-out (with bmpc_prepare_batch and bmpc_solve_batch, no host
 * round trip).  Per instance: the solve is accepted if status == 0 or the summed constraint violation beyond 1e-6 is below
 * 1e-4 (BoundMPC.py:461-465); an accepted x becomes the previous solution (prev_x, error_count = 0); after a rejected solve
 * the previous solution is kept and error_count incremented (its nodes error_count.. are post-processed); traj / state_out as
 * for bmpc_post_batch.  advance != 0: state_out also carries the joint state, pose, Cartesian velocity and applied jerk after
 * one sample under the first kept jerk (integrate_joint, utils/util_functions.py:152-161), i.e. it is the input state of the
 * next bmpc_prepare_batch.  DEVICE pointers, asynchronous.  prev_x [batch, n] and error_count [batch] are updated in place. */
int bmpc_finish_batch(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                      const int32_t* path_id, const int32_t* sector, const double* state, const double* x, const double* g,
                      const int32_t* status, double* prev_x, int32_t* error_count, double* traj, double* state_out,
                      int32_t advance, void* cuda_stream);
/* One whole MPC step — `BoundMPC.step` (BoundMPC.py:306-506,508-770) — for `batch` controllers whose state the caller holds on
 * the HOST (the single controller of bound_mpc_node.py is batch = 1): bmpc_prepare_batch, bmpc_solve_batch and bmpc_finish_batch
 * (advance = 0) back to back on the handle's stream, one staged copy each way, one synchronisation.
 *   path_tables, path_id, state          as for bmpc_prepare_batch
 *   sector, prev_x, error_count          in / out (window position, previous solution, BoundMPC.error_count)
 *   x        [batch, n]                  the solver's solution of this step (kept or not, see status / error_count)
 *   traj     [batch, N, 42], state_out [batch, 76]   as for bmpc_post_batch, for the trajectory the controller keeps
 *   ref [batch, N, 55], err [batch, N, 33] or NULL, NULL   logging branch, as for bmpc_post_log_batch
 *   iters, status [batch]                solver statistics (stats()['iter_count'], status == 0 <=> stats()['success']) */
int bmpc_mpc_step_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                             const int32_t* path_id, int32_t* sector, const double* state, double* prev_x, int32_t* error_count,
                             double* x, double* traj, double* state_out, double* ref, double* err, int32_t* iters, int32_t* status);
/* bmpc_post_batch with HOST pointers (copies inside, synchronises). */
int bmpc_post_batch_host(bmpc_handle* h, int32_t batch, const double* path_tables, int32_t n_paths, int32_t path_rows,
                         const int32_t* path_id, const int32_t* sector, const double* state, const double* w,
                         const int32_t* error_count, double* traj, double* state_out);

/* Evaluation of the NLP functions for parity tests (what CasADi's nlp_f, nlp_g, nlp_grad_f,
 * nlp_jac_g, nlp_hess_l provide, casadi_ocp_formulation.py:389), HOST pointers, small batches.
 *   lam   [batch, 48 N]: per stage 36 equality multipliers then 12 interval-row multipliers
 *   f     [batch]         g      [batch, m]   (reference form)
 *   d     [batch, 12 N]   interval-form inequality rows
 *   grad  [batch, n]      jac    [batch, 48 N, n]  rows per stage: 36 equality rows, 12 interval rows
 *   hess  [batch, n, n]   Hessian of f + lam . (c, d)
 * Output pointers may be NULL. */
int bmpc_eval_batch_host(bmpc_handle* h, int32_t batch, const double* x, const double* p, const double* lam,
                         double* f, double* g, double* d, double* grad, double* jac, double* hess);

/* One Newton step of the interior-point iteration at caller-supplied primal-dual points, for parity tests of the
 * structure-exploiting KKT solve (what MUMPS does for Ipopt: the factorisation of the augmented system)
 * against a dense solve of the assembled system.  HOST pointers, small batches.
 *   v      [batch, 3 n + 60 N]: x (n), y (36 N), s (12 N), z_s (12 N), z_L (n), z_U (n)   (interval form, see bmpc_eval_batch_host)
 *   mu, delta_w [batch]         barrier parameter and Hessian perturbation (inertia correction)
 *   dx     [batch, n]           primal step;   ynew [batch, 36 N]  equality multipliers of the full step
 *   ok     [batch]              0: the reduced Hessian is not positive definite for this delta_w (no step returned) */
int bmpc_kkt_step_batch_host(bmpc_handle* h, int32_t batch, const double* v, const double* p, const double* mu,
                             const double* delta_w, double* dx, double* ynew, int32_t* ok);

/* name and duration of the kernels launched by the last solve on this handle (diagnostics):
 * number of kernel launches issued by this library since the handle was created */
int64_t bmpc_launch_count(const bmpc_handle* h);

/* launch shape of the solver kernel chosen by bmpc_create (diagnostics / bench config): threads per CTA,
 * resident CTAs per SM (occupancy-checked), dynamic shared memory per CTA in bytes, number of SMs */
int bmpc_launch_shape(const bmpc_handle* h, int32_t* threads, int32_t* ctas_per_sm, int32_t* smem_bytes, int32_t* sms);

/* Diagnostics for bench.py's roofline denominator (MEASURED_PEAKS.json has no FP64
 * entry): runs a register-resident DFMA loop (kind 0) or an mma.sync.m8n8k4.f64 DMMA loop
 * (kind 1) on every SM of the handle's device and returns the sustained rate in FLOP/s. */
int bmpc_fp64_peak(bmpc_handle* h, int32_t kind, double* flops_per_s);

const char* bmpc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
