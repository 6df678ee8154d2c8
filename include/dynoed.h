/* dynoed.h — C ABI of libdynoed_cuda.so: the B200-native batched evaluator for DynamicOED.jl's
 * hot path (sensitivity-augmented integration over the measurement grid -> OED criterion and
 * its gradient, for many candidate designs at once).
 *
 * There is no FFI in the reference (pure Julia).  Each entry point below names the reference
 * interface it stands in for; INTEGRATION.md shows the Julia `ccall` binding a maintainer adds.
 *
 * Conventions
 *   - plain C, no torch/CUDA types in signatures; every function returns 0 on success or a
 *     negative DYNOED_E* code; the message is available from dynoed_last_error().
 *   - the caller owns every buffer it passes.  `designs`, `crit`, `grad`, `fim` may be host
 *     pointers (pageable or pinned) or device pointers; all pointers of one call must be of the
 *     same kind (detected with cudaPointerGetAttributes).
 *   - NaN/Inf in outputs are data, not errors.
 *   - a ctx is bound to one device and is not re-entrant: one ctx per host thread / stream.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DYNOED_ECUDA.
 *   - stream order is kept: whatever is enqueued on the ctx stream after an evaluation sees its
 *     results, and an evaluation sees whatever was enqueued before it.  The one exception is
 *     opt-in: a dynoed_eval_batch_async call that carries DYNOED_OVERLAP_OK may start while the
 *     previous evaluation launch of the same ctx (same stream, same kernel, disjoint buffers) still
 *     drains.  Calls that share a buffer, or come from different contexts, never overlap.
 */
#ifndef DYNOED_H
#define DYNOED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNOED_ABI_VERSION 1

#define DYNOED_OK 0
#define DYNOED_EINVAL (-1)  /* bad argument / inconsistent descriptor */
#define DYNOED_ECUDA (-2)   /* CUDA runtime error (incl. "no device") */
#define DYNOED_ENOMODEL (-3)
#define DYNOED_EUNSUPPORTED (-4)

/* criteria — /root/reference/src/fisher.jl:26-107, evaluated on Symmetric(F + tau I) (:12-15) */
#define DYNOED_FISHER_A 0 /* -tr F          fisher.jl:28-30  */
#define DYNOED_FISHER_D 1 /* -det F         fisher.jl:43-45  */
#define DYNOED_FISHER_E 2 /* -lambda_min F  fisher.jl:58-60  */
#define DYNOED_A 3        /* tr F^-1        fisher.jl:73-76  */
#define DYNOED_D 4        /* 1/det F        fisher.jl:89-91  */
#define DYNOED_E 5        /* max 1/lambda   fisher.jl:104-107 */
#define DYNOED_LOG_D 6    /* -log det F: the log form of the D criterion (BASELINE north_star's "FisherD logdet F");
                           * an EXTENSION, not in the reference (its FisherD is -det F) */

/* fixed-step integrators (the reference passes `alg` to OrdinaryDiffEq.solve,
 * /root/reference/src/problem.jl:25, /root/reference/src/discretize.jl:314-317) */
#define DYNOED_RK4 0
#define DYNOED_TSIT5 1 /* Tsit5 tableau, fixed step */
/* fixed-grid Rosenbrock methods for the stiff / DAE configuration (the reference uses Rodas5,
 * /root/reference/test/references/Rober.jl:29); autonomous models only.  Gradients: weights and tau
 * from per-interval quadratures; unknown initial conditions and control values by one forward-mode
 * dual pass each (there is no adjoint sweep for these schemes) */
#define DYNOED_ROS2 2  /* 2 stages, order 2, L-stable */
#define DYNOED_ROS3P 3 /* 3 stages, order 3, A-stable */
#define DYNOED_RODAS3 4 /* 4 stages, order 3, stiffly accurate, L-stable */

/* flags of dynoed_eval_batch[_async] */
#define DYNOED_GRAD_NO_CONTROLS 1u /* gradient w.r.t. weights, tau and unknown initial conditions only:
                                    * forward sweep with per-interval quadratures, no adjoint sweep
                                    * (explicit tableaus) / no control passes (Rosenbrock);
                                    * control entries of grad are NaN ("not computed") */
#define DYNOED_OVERLAP_OK 2u       /* dynoed_eval_batch_async: this launch may start while the previous launch
                                    * of the same ctx drains.  The caller asserts that NO work enqueued on the
                                    * stream after that previous launch produces this launch's inputs (the
                                    * kernels do not wait for the preceding grid's memory to be flushed).
                                    * Without the flag every launch is fully ordered after its predecessor. */

typedef struct dynoed_ctx dynoed_ctx;

/* Static facts about a compiled-in model (the emitted augmented system of
 * /root/reference/src/augment.jl:128-244 after structural_simplify). */
typedef struct dynoed_model_info {
    int32_t nx, np, n_obs, n_ctrl, n_ic, n_theta, nz, nF;
    int32_t autonomous;        /* 1 if the right-hand side does not depend on t explicitly */
    int32_t ic_state[8];       /* 0-based state index of each unknown initial condition */
    int32_t flops_primal, flops_primal_z, flops_adjoint, flops_consts, flops_ros_f; /* per call, FMA = 2 */
    double x0[8];
    double G0[64];             /* column-major nx x np */
    double theta[48];          /* default parameter values (non-control, non-weight) */
} dynoed_model_info;

/* One OED problem = model + merged time grid + design-vector layout + criterion + integrator.
 * Replaces what OEDProblem / Timegrid / OEDRemake hold on the Julia side
 * (/root/reference/src/problem.jl:10-30, /root/reference/src/discretize.jl:32-83, :229-265). */
typedef struct dynoed_problem_desc {
    uint32_t struct_size;      /* = sizeof(dynoed_problem_desc) */
    const char* model;         /* name of a compiled-in model */
    int32_t criterion;         /* DYNOED_FISHER_A ... */
    int32_t method;            /* DYNOED_RK4 | DYNOED_TSIT5 | DYNOED_ROS2 | DYNOED_ROS3P | DYNOED_RODAS3 */
    int32_t n_intervals;       /* N: merged grid intervals (Timegrid.timespans) */
    const double* tgrid;       /* [N+1] merged grid points */
    int32_t substeps;          /* uniform substeps per merged interval (>0), or 0 with step_edges */
    const int32_t* edge_count; /* [N] number of substep edges per interval (graded grids) or NULL */
    const double* step_edges;  /* concatenated relative edges in [0,1] per interval, or NULL */
    const int32_t* indicators; /* [(n_ctrl+n_obs)][N], 0-based (Timegrid.indicators - 1) */
    const int32_t* ctrl_offsets; /* [n_ctrl] offset of each control's block in a design row */
    const int32_t* w_offsets;    /* [n_obs]  offset of each weight block */
    int32_t ic_offset;         /* offset of the initial-condition block (n_ic scalars) */
    int32_t tau_offset;        /* offset of the regularisation variable */
    int32_t n_var;             /* design-vector length */
    const double* theta;       /* [n_theta] parameter values or NULL for the model defaults */
    const double* x0;          /* [nx] or NULL */
    const double* G0;          /* [nx*np] or NULL */
    int32_t tile;              /* designs per CTA (0 = default 128; multiple of 32, <= 128) */
    int32_t max_batch_hint;    /* 0 or expected designs per call (sizes scratch eagerly) */
} dynoed_problem_desc;

typedef struct dynoed_kernel_stats {
    int32_t regs_per_thread, static_smem, dynamic_smem, blocks_per_sm, grid, block, sm_count;
    int64_t launches;          /* kernels launched by this ctx so far */
    int64_t scratch_bytes;
    int32_t n_steps, stages;
    double flops_per_design_eval, flops_per_design_grad; /* algorithmic, FMA = 2 */
    double bytes_per_design;   /* algorithmic HBM bytes: n_var in + (1 + n_var + nF) out */
    int64_t overlapped_launches; /* launches allowed to start while their predecessor drained */
    int64_t small_batch_launches; /* launches served by the small-batch kernels (direction- or time-parallel;
                                     objective-only calls of the time-parallel kernel included) */
    int32_t small_batch_max;   /* largest gradient call (designs) the small-batch kernels take; 0 = never */
    int32_t time_parallel_max; /* largest call (designs) that runs one CTA per design (tp_eval_kernel; gradient
                                  calls, and objective-only calls unless its tables spill); 0 = never */
} dynoed_kernel_stats;

int dynoed_abi_version(void);
int dynoed_model_count(void);
const char* dynoed_model_name(int i);
int dynoed_model_query(const char* name, dynoed_model_info* out);

/* OEDProblem(sys, criterion; alg, tspan) + build_predictor (problem.jl:23-30, :80-90) */
int dynoed_create(const dynoed_problem_desc* desc, int device, dynoed_ctx** out);
void dynoed_destroy(dynoed_ctx* ctx);
int dynoed_set_criterion(dynoed_ctx* ctx, int criterion);
/* run on a caller-owned CUDA stream (cudaStream_t as void*; NULL = the legacy default stream);
 * dynoed_own_stream() goes back to the stream the ctx created for itself */
int dynoed_set_stream(dynoed_ctx* ctx, void* cuda_stream);
int dynoed_own_stream(dynoed_ctx* ctx);

/* The objective closure and its ForwardDiff gradient (problem.jl:112-121, :154), batched:
 *   crit[k]  = c(F_k(t_f) + tau_k I) + tau_k
 *   grad[k]  = d crit[k] / d designs[k]   (same layout as designs; NULL = objective only)
 *   fim[k]   = packed upper triangle of F_k(t_f) (fisher.jl:3-5 order; NULL = skip)
 * designs: [n][n_var] row-major in the layout of discretize.jl:110-137.  Blocking. */
int dynoed_eval_batch(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit, double* grad,
                      double* fim, uint32_t flags);
/* device pointers only; enqueues on the ctx stream and returns */
int dynoed_eval_batch_async(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit,
                            double* grad, double* fim, uint32_t flags);
int dynoed_sync(dynoed_ctx* ctx);
/* Pipelined HOST-buffer evaluation for a single host thread: enqueues the chunked H2D / kernel / D2H
 * pipeline and returns at once with a call id; dynoed_wait(ctx, id) blocks until that call's results
 * are in the caller's arrays.  Consecutive calls overlap (one call's D2H under the next call's H2D),
 * which a sequence of blocking dynoed_eval_batch calls cannot do.  Use page-locked arrays
 * (dynoed_host_alloc) that stay valid and untouched until the wait; pageable arrays are accepted but
 * are staged synchronously (the call has completed when it returns).  At most two calls in flight. */
int dynoed_eval_batch_host_async(dynoed_ctx* ctx, const double* designs, int64_t n, double* crit,
                                 double* grad, double* fim, uint32_t flags, int64_t* call_id);
int dynoed_wait(dynoed_ctx* ctx, int64_t call_id);
/* arg-min of the criterion over the most recent batch (multistart selection) */
int dynoed_argmin(dynoed_ctx* ctx, int64_t* idx, double* val);

/* same record copied into caller-owned DEVICE memory on the ctx stream (no host sync): feeds the
 * NCCL all-gather of per-rank (min, index) records in the multi-GPU sweep */
int dynoed_argmin_async(dynoed_ctx* ctx, double* d_val, int64_t* d_idx);

/* the same record as ONE 16-byte device-to-device copy: {double value, int64 local index} */
int dynoed_argmin_record_async(dynoed_ctx* ctx, void* d_record);
/* the records of the last `count` batches (oldest first, count <= 64) as [count][2] x 8 bytes,
 * copied on the ctx stream: lets a multi-GPU sweep gather several batches with ONE collective, so
 * that consecutive evaluation launches stay back to back (and overlap their tails) */
int dynoed_argmin_history_async(dynoed_ctx* ctx, int count, void* d_records);
/* arg-min over gathered records ([world][groups][2] x 8 bytes, rank-major) on the ctx stream:
 * d_out[groups][3] x 8 bytes = {double value, int64 global index = rank * shard_size + local,
 * int64 owner rank}.  NaN never wins, ties go to the smallest global index.  (BASELINE config 5.) */
int dynoed_argmin_gathered(dynoed_ctx* ctx, const void* d_records, int world, int groups,
                           int64_t shard_size, void* d_out);

/* Selection helpers normally run on the ctx stream, where any node between two evaluation launches
 * ends their overlap.  With an auxiliary stream set (cudaStream_t as void*, NULL = unset) they run
 * there instead: the ring readers (dynoed_argmin_history_async, dynoed_pack_winner) are ordered after
 * every evaluation enqueued so far by an event; the reducers over gathered data
 * (dynoed_argmin_gathered, dynoed_select_winner) are simply enqueued on it — issue the collective
 * that feeds them on the same stream.  The caller orders its own stream after the aux stream. */
int dynoed_set_aux_stream(dynoed_ctx* ctx, void* cuda_stream);

/* Multistart sweep selection on the device (BASELINE config 5; replaces N sequential calls of the
 * objective closure /root/reference/src/problem.jl:112-121 plus a host-side arg-min).
 * dynoed_pack_winner: the shard was evaluated as the last `nbatches` (<= 64) device-path batches of
 * this ctx, batch b covering rows [b * batch_stride, ...) of designs / grad ([n_local][n_var], device).
 * Writes d_payload[2 + 2 n_var] = {best value, index_base + best row, design row, gradient row}
 * ({+inf, -1, 0...} for an empty or all-NaN shard; grad may be NULL -> zeros).
 * dynoed_select_winner: d_gathered = the payloads of all ranks, [world][2 + 2 n_var] (e.g. from ONE
 * all-gather); writes d_out[2 + 2 n_var + 1] = {winner's payload, owner rank} (owner -1 if none).
 * NaN never wins, ties go to the smallest global index.  One small kernel each, no host sync. */
int dynoed_pack_winner(dynoed_ctx* ctx, int nbatches, const double* designs, const double* grad,
                       int64_t n_local, int64_t batch_stride, int64_t index_base, double* d_payload);
int dynoed_select_winner(dynoed_ctx* ctx, const double* d_gathered, int world, double* d_out);

/* scratch-set guard counters (synchronises the device): CTAs that had to wait for a scratch set still
 * held by the launch two back, and waits abandoned after 20 s (must stay 0) */
int dynoed_guard_stats(dynoed_ctx* ctx, int64_t* waits, int64_t* timeouts);

/* the predictor p -> (X, t) of build_predictor (problem.jl:80-90) at the merged-grid points:
 * X[k] is [N+1][nz+nF] row-major (state order [x; vec(G); F_triu], augment.jl:240) */
int dynoed_trajectory(dynoed_ctx* ctx, const double* designs, int64_t n, double* X);

/* compute_information_gain (/root/reference/src/utilities.jl:47-70) at the merged-grid points:
 *   F  [n][nF]                       packed F(t_f)
 *   P  [n][N+1][n_obs][np][np]       local information gain  P_i(t_j) = Q_i^T Q_i,  Q = h_x G
 *   Pi [n][N+1][n_obs][np][np]       F^-1 P_i(t_j) F^-1
 * any of F / P / Pi may be NULL.  Blocking; host or device pointers (all of one kind). */
int dynoed_information_gain(dynoed_ctx* ctx, const double* designs, int64_t n, double* F, double* P,
                            double* Pi);

/* Hessians of the objective at n points, hess[n][n_var][n_var] (symmetric, row-major), by central
 * differences of the exact gradient: the 2 n_var + 1 perturbed designs of every point are evaluated
 * as ordinary batches of the gradient kernel (one launch for up to 131072 / (2 n_var + 1) points),
 * step h_i = rel_step * max(1, |x_i|) (rel_step = 0: cbrt(eps) = 6.06e-6; error ~ 1e-10 relative to
 * the gradient's scale).  crit[n] / grad[n][n_var] (either may be NULL) receive objective and
 * gradient at the points themselves.  Replaces the nested-ForwardDiff Hessian Optimization.jl
 * builds when Ipopt runs WITHOUT hessian_approximation = "limited-memory"
 * (/root/reference/README.md:63, /root/reference/src/problem.jl:154).  Blocking; host or device
 * pointers (all of one kind); overwrites the arg-min record of the previous batch. */
int dynoed_hessian_batch(dynoed_ctx* ctx, const double* designs, int64_t n, double rel_step, double* crit,
                         double* grad, double* hess);

/* Fixed-controls shortcut for batches of MEASUREMENT designs (integer / branch-and-bound mode,
 * /root/reference/src/problem.jl:137-146; F is linear in w, /root/reference/src/augment.jl:223-226):
 * dynoed_fim_basis integrates once per weight column with the controls / initial conditions of
 * `design` (host, one row) and returns basis[n_var][nF] (zero rows for non-weight columns);
 * dynoed_eval_fixed_controls then evaluates  F = designs x basis, the objective and its gradient
 * w.r.t. weights and tau (control / IC entries of grad are 0) with one HBM-streaming kernel. */
int dynoed_fim_basis(dynoed_ctx* ctx, const double* design, double* basis);
int dynoed_eval_fixed_controls(dynoed_ctx* ctx, const double* basis, const double* designs, int64_t n,
                               double* crit, double* grad, double* fim);

/* criterion(F, tau) on caller-supplied packed symmetric matrices (fisher.jl:12-15):
 * value[k] = c(Symmetric(F_k + tau_k I)); dcdF (optional) = packed d c / d F.  np <= 24. */
int dynoed_criterion_batch(int device, int criterion, int np, const double* F_packed,
                           const double* tau, int64_t batch, double* value, double* dcdF);

/* register-resident DFMA-chain microbenchmark: the FP64 roofline denominator (TFLOP/s, FMA=2) */
int dynoed_fp64_peak(int device, void* cuda_stream, double* tflops, double* ms);

/* FP64 instruction throughput (thread-level Ginstr/s) for operand patterns: 0 DFMA reg*const+const,
 * 1 DFMA three registers, 2 DFMA two registers, 3 DMUL, 4 DADD, 5 DFMA reg*uniform+reg,
 * 6 DFMA with a shared multiplicand (operand reuse), 7 DFMA second factor from two registers,
 * 8 DMMA m8n8k4 chains, 9 DMMA + DFMA mix (8, 9 in thread-level FMA equivalents) */
int dynoed_fp64_microbench(int device, int variant, double* ginstr_per_s);

/* CUDA-event timing of the evaluation kernel itself (device-pointer path), for the roofline */
int dynoed_set_profiling(dynoed_ctx* ctx, int on);
int dynoed_kernel_time(dynoed_ctx* ctx, double* mean_ms, int64_t* count);
/* with_grad: 0 objective only, 1 gradient, 2 gradient with DYNOED_GRAD_NO_CONTROLS */
int dynoed_kernel_stats_query(dynoed_ctx* ctx, int with_grad, dynoed_kernel_stats* out);
int dynoed_host_alloc(void** p, size_t bytes); /* pinned host memory for the e2e path */
int dynoed_host_free(void* p);
const char* dynoed_last_error(const dynoed_ctx* ctx); /* ctx may be NULL: last global error */

#ifdef __cplusplus
}
#endif
#endif /* DYNOED_H */
