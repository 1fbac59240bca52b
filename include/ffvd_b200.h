/*
 * ffvd_b200.h -- C ABI of libffvd_b200.so: the B200 (sm_100a) implementation of FFVD's
 * per-iteration GPSSM log-joint + gradient evaluation and SG-HMC update.
 *
 * The reference (xuhuifan/FFVD) has no FFI: its hot path is a Python operator API on top of
 * the TensorFlow runtime.  Each entry point below replaces the TF sub-graph that the cited
 * reference function builds; the Python package `ffvd_b200` (same module / function names as
 * `vfegpssm`) binds them with ctypes.  INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - every function returns an int status: 0 ok; <0 error (see FFVD_E_*); >0 = 1-based index
 *     of the failing Cholesky pivot (which matrix failed is in ffvd_last_error()).
 *   - tensors cross the ABI as DLPack `DLManagedTensor*` (borrowed for the duration of the
 *     call; the library never calls the deleter).  float64, C-contiguous, byte_offset honoured.
 *     kDLCUDA tensors are used in place; kDLCPU / kDLCUDAHost tensors are staged with an
 *     explicit cudaMemcpy in and out (a transfer, never CPU compute -- there is no CPU path).
 *   - outputs are caller-allocated tensors.  NULL output pointers are skipped.
 *   - work is enqueued on the context's stream; calls that return host scalars or touch host
 *     tensors synchronise that stream before returning, pure-device calls do not.
 */
#ifndef FFVD_B200_H_
#define FFVD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- minimal DLPack (v0.x "dltensor" capsule ABI) ------------------------------------- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;   /* code 2 = float */
typedef struct {
  void* data; DLDevice device; int32_t ndim; DLDataType dtype;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

/* ---- status codes ---------------------------------------------------------------------- */
#define FFVD_OK               0
#define FFVD_E_BADARG        -1   /* null pointer, bad flag                                 */
#define FFVD_E_DTYPE         -2   /* not float64                                            */
#define FFVD_E_SHAPE         -3   /* rank / extent mismatch, or not C-contiguous            */
#define FFVD_E_DEVICE        -4   /* tensor lives on another GPU than the context           */
#define FFVD_E_CUDA          -5   /* CUDA runtime error (message in ffvd_last_error)        */
#define FFVD_E_UNSUPPORTED   -6   /* valid reference option that this build does not cover  */
#define FFVD_E_LIMIT         -7   /* size beyond this build's limits (M > 2048, Din > 31)   */
#define FFVD_E_STALE         -8   /* FFVD_FLAG_REUSE_KZZ although Z / logv / logl changed (results of the call are NaN) */

#define FFVD_KERNEL_SE      0     /* kernels_multi_output.py:240-247 SquaredExponential (ARD) */
#define FFVD_KERNEL_LINEAR  1     /* kernels.py:250-281 LinearK (scalar variance)              */

/* flags for ffvd_nll_grads_* */
#define FFVD_FLAG_PRIOR_Z_NORMAL  1   /* dgp_model.py:108-109 (prior_type="normal", the CLI default) */
#define FFVD_FLAG_PRIOR_ONCE      2   /* count shared-parameter priors once instead of S times  */
#define FFVD_FLAG_NO_GRADS        4   /* forward only: nll + terms                               */
#define FFVD_FLAG_ASYNC           8   /* do not synchronise to read back the Cholesky status     */
/* The caller asserts that Z, logv, logl, the kernel kind, the jitter and every shape are unchanged since the previous
 * ffvd_nll_grads_* call on this context: K(Z,Z)'s Cholesky factor, L^{-1}, L^{-T} and the scaled inducing inputs are
 * reused instead of recomputed (SG-HMC chains that sample only X / U evaluate 21 times per outer iteration with fixed Z
 * and hyper-parameters, base_model.py:915-933).  Ignored when the context holds no factors for these shapes.
 * The assertion is CHECKED on the device: a 64-bit content hash of Z / logv / logl is recorded when the factors are built
 * and compared on every reusing call; on a mismatch the call's nll / terms (conditional: mean / var) are NaN and the next
 * call that synchronises returns FFVD_E_STALE.  The factors of a failed factorisation are never reused. */
#define FFVD_FLAG_REUSE_KZZ       64
/* time-sharded evaluation of ONE trajectory (ffvd_b200/distributed.py: a block of consecutive transitions per GPU):
 * every block is an ordinary problem whose nll / gradients are then rescaled by T_block / T_total by the caller */
#define FFVD_FLAG_NO_SHARED_PRIORS 16 /* drop the priors on Z, U, kernel hypers, logQ, C, d, logR (blocks other than the first) */
#define FFVD_FLAG_NO_X0_PRIOR      32 /* drop -1/2 |X_0|^2 (dgp_model.py:252): the block does not start at t = 0 */

/* Collapsed bound under TIME sharding (one trajectory split over ranks; conditionals_multi_output.py:246-254 needs
 * F^T F and F^T delta summed over ALL transitions before H^{-1} can be formed):
 *   1. every rank: ffvd_nll_grads_collapsed(block, flags | FFVD_FLAG_COLLAPSED_P1_ONLY)      -- pass 1 over its block
 *   2. every rank: ffvd_collapsed_stats_allreduce(ctx)   (or _get / sum / _set by other means)
 *   3. every rank: ffvd_nll_grads_collapsed(same tensors, flags | FFVD_FLAG_COLLAPSED_RESUME [| FFVD_FLAG_NO_REPLICATED])
 * FFVD_FLAG_NO_REPLICATED (every rank but one) drops what is a function of the global statistics only and therefore
 * identical on every rank: the log det H and quadratic terms, their dJ/dlogQ part, and the Cholesky backward of
 * G = Mat' S + c b^T.  Results are then rescaled by T_block / T and summed like the uncollapsed blocks. */
#define FFVD_FLAG_COLLAPSED_P1_ONLY 128
#define FFVD_FLAG_COLLAPSED_RESUME  256
#define FFVD_FLAG_NO_REPLICATED     512

/* Bitwise-repeatable results.  By default the shared sums (S = sum a a^T, u-bar, Z-bar, hyper-parameter gradients, x-bar
 * rows that several work items touch) are accumulated with FP64 REDs whose arrival order changes from run to run, so results
 * agree only to ~1e-12.  With this flag every accumulator element has ONE writer thread per private copy (a copy of the S
 * region per CTA, of the small accumulators per (CTA, warp), separate x-bar planes per kind of contribution, stored partials in
 * the Kzz backward) and the copies are summed in index order: two evaluations of the same inputs on the same device are
 * bit-identical.  Costs memory (C3: +2.7 GB) and a few percent of time; not capturable in CUDA graphs; the split collapsed
 * evaluation (P1_ONLY / RESUME) is not covered. */
#define FFVD_FLAG_DETERMINISTIC 1024

typedef struct ffvd_ctx ffvd_ctx;

/* One GPSSM problem (SURVEY section 8 batch-axis contract).  Shapes:
 *   X (S,T+1,D) or (T+1,D); Z (M,Din); U (M,D); logv (D); logl (D,Din) [SE only, else NULL];
 *   logQ (D); C (D,Dy); d (Dy); logR (Dy,Dy); Y (T,Dy); ctrl (T,Din-D) [NULL if Din==D].
 * Replaces the tf.Variables of dgp_model.py:56-69,176-185 and likelihoods.py:12-24,45-55. */
typedef struct {
  DLManagedTensor *X, *Z, *U, *logv, *logl, *logQ, *C, *d, *logR, *Y, *ctrl;
} ffvd_problem;

/* Outputs of one nll+gradient evaluation; any pointer may be NULL.
 *   nll (S) [or () for 2-d X]; terms (S,6) = {prior, loglik, xq, trace, term1, term2}
 *   (dgp_model.py:286-297: nll = sum of the six); g_* = d nll / d *, same shape as the
 *   parameter; g_X per sample, all others summed over the S samples. */
typedef struct {
  DLManagedTensor *nll, *terms, *g_X, *g_Z, *g_U, *g_logv, *g_logl, *g_logQ, *g_C, *g_d, *g_logR;
} ffvd_outputs;

int  ffvd_version(void);
const char* ffvd_status_string(int status);
const char* ffvd_last_error(void);

/* stream: a cudaStream_t (as void*) or NULL for a stream owned by the context.  To run on the legacy
 * default stream pass cudaStreamLegacy ((void*)0x1), not 0. */
int ffvd_ctx_create(int device, void* stream, ffvd_ctx** out);
int ffvd_ctx_destroy(ffvd_ctx* ctx);
int ffvd_ctx_synchronize(ffvd_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t ffvd_ctx_launch_count(ffvd_ctx* ctx);
/* Sum of the device times (ms, CUDA events on the context's stream) and the number of fused
 * tile-kernel launches recorded since the last reset (ring of 256); synchronises the stream.
 * bench.py derives the roofline figure from it. */
int ffvd_ctx_fused_time(ffvd_ctx* ctx, int reset, double* total_ms, int64_t* count);
/* Diagnostic: per-phase SM clock totals of the fused tile kernel (16 slots, summed over CTAs).
 * Only in builds made with -DFFVD_PHASE_TIMING (tools/phase_timing.py); FFVD_E_UNSUPPORTED otherwise. */
int ffvd_debug_phase_clocks(ffvd_ctx* ctx, int reset, uint64_t* out16);

/* Diagnostic: out[i] = the fused kernels' branch-free exp routine at x[i] <= 0 (kernels_multi_output.py:246-247 K_r2 is
 * exp(-r2/2); accuracy test against libm). */
int ffvd_debug_exp(ffvd_ctx* ctx, DLManagedTensor* x, DLManagedTensor* out);

/* kernels_multi_output.py:202-214,246-247 / kernels.py:270-276  K(X, X2) -> out (N,N2).
 * X2 may be NULL (K(X,X)).  logv: () ; logl: (Din) for SE, NULL for Linear. */
int ffvd_kernel_K(ffvd_ctx*, int kind, DLManagedTensor* X, DLManagedTensor* X2,
                  DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* out);
/* kernels_multi_output.py:199-200 / kernels.py:278-281  Kdiag(X) -> out (N). */
int ffvd_kernel_Kdiag(ffvd_ctx*, int kind, DLManagedTensor* X, DLManagedTensor* logv,
                      DLManagedTensor* logl, DLManagedTensor* out);

/* conditionals_multi_output.py:124-169 kernel_pre_cal: per output dim d,
 * LinvT_out[d] = chol(K_d(Z) + jitter I)^{-T}  (D,M,M).  logv (D), logl (D,Din). */
int ffvd_kernel_pre_cal(ffvd_ctx*, int kind, DLManagedTensor* Z, DLManagedTensor* logv,
                        DLManagedTensor* logl, double jitter, DLManagedTensor* LinvT_out);

/* conditionals_multi_output.py:73-120 (list of D kernels, f (M,D)) and, with
 * shared_kernel=1, conditionals.py:69-107 (one kernel for all R columns of f; logv (), logl (Din)).
 * Also the prediction form conditional_after_kernel_precalculation (cmo:306-387): the factors are
 * recomputed on the device, so Lm_inverse_seq is not an input.
 * mean_out, var_out (N,R).  q_sqrt (nullable): (M,R) per-point scales (cmo:51-52) or (R,M,M) /
 * (1,M,M) factors (cmo:53-62; (1,M,M) = one factor for every output, which is what the
 * reference's [:, :, 0] indexing computes); requires white=1.  full_cov=1, q_sqrt with white=0 and return_Lm:
 * ffvd_conditional_dense. */
int ffvd_conditional(ffvd_ctx*, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                     DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f,
                     DLManagedTensor* q_sqrt, int white, int full_cov, double jitter,
                     DLManagedTensor* mean_out, DLManagedTensor* var_out);

/* The same with flags: FFVD_FLAG_REUSE_KZZ -- the caller asserts that Z, logv, logl (same tensors), kind, jitter and
 * shapes are those of the previous ffvd_conditional* / ffvd_kernel_pre_cal call on this context, whose factors are then
 * reused: this is what passing `Lm_inverse_seq` to conditional_after_kernel_precalculation means in the reference
 * (kernel_pre_cal once, then one conditional per time step: base_model.py:36,49,210,289); FFVD_FLAG_ASYNC skips the
 * read-back of the Cholesky status. */
int ffvd_conditional_ex(ffvd_ctx*, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                        DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f,
                        DLManagedTensor* q_sqrt, int white, int full_cov, double jitter, int flags,
                        DLManagedTensor* mean_out, DLManagedTensor* var_out);

/* Every option of base_conditional (conditionals.py:6-66, conditionals_multi_output.py:6-70) on explicit matrices, op for
 * op: full_cov (var_out (R,N,N)), q_sqrt 2-d / 3-d with white or not, and return_Lm (Lm_out (kernels,M,M) = the Cholesky
 * factor of K(Z,Z) + jitter I, nullable).  For a handful of prediction points: O(M N (M + N)) work in plain FP64 kernels;
 * the hot-path branch (white, diagonal variances, any N) is ffvd_conditional_ex. */
int ffvd_conditional_dense(ffvd_ctx*, int kind, int shared_kernel, DLManagedTensor* Xnew, DLManagedTensor* Z,
                           DLManagedTensor* logv, DLManagedTensor* logl, DLManagedTensor* f, DLManagedTensor* q_sqrt,
                           int white, int full_cov, double jitter, DLManagedTensor* mean_out, DLManagedTensor* var_out,
                           DLManagedTensor* Lm_out);

/* conditionals_multi_output.py:206-227 collapse_u_mean_after_kernel_precalculation: the optimal
 * collapsed q(u).  Per sample s and output d: F = K(Xc,Z) L^{-T}, H = F^T F / Q_d + I,
 * U_mean_out[s,:,d] = H^{-1} F^T (x_{1:T,d} - x_{0:T-1,d}) / Q_d   (S,M,D) [(M,D) for 2-d X],
 * LHinvT_out[s,d] = chol(H)^{-T} (upper)   (S,D,M,M), nullable.
 * Only X, Z, logv, logl, logQ, ctrl of the problem matter (the others must still be valid tensors). */
int ffvd_collapse_u_mean(ffvd_ctx*, int kind, const ffvd_problem* p, double jitter,
                         DLManagedTensor* U_mean_out, DLManagedTensor* LHinvT_out);

/* likelihoods.py:96-111 (vec=1 -> out (N)) and :89-93 (vec=0 -> out (N,Dy)); no -0.5 log 2pi. */
int ffvd_logdensity_norm_diag(ffvd_ctx*, DLManagedTensor* y, DLManagedTensor* ymean,
                              DLManagedTensor* Rchols, int vec, DLManagedTensor* out);

/* likelihoods.py:114-127 logdensity_norm with a full lower-triangular factor Rchols (Dy,Dy): out[n] =
 * -1/2 |L^{-1}(y_n - ymean_n)|^2 - sum log L_ii (no -0.5 log 2pi).  y is (N,Dy) or ONE row (Dy) broadcast over the N rows
 * of ymean (the particle-Gibbs weights, base_model.py:62-66).  Dy <= 64. */
int ffvd_logdensity_norm(ffvd_ctx*, DLManagedTensor* y, DLManagedTensor* ymean, DLManagedTensor* Rchols,
                         DLManagedTensor* out);

/* dgp_model.py:248-297 + tf.gradients (base_model.py:148): uncollapsed q(u) (cases 1,2,3,6,7;
 * regularizer :337-359) and collapsed-u bound (cases 4,5; conditionals_multi_output.py:230-257). */
int ffvd_nll_grads_uncollapsed(ffvd_ctx*, int kind, const ffvd_problem* p, int flags, double jitter,
                               const ffvd_outputs* o);
int ffvd_nll_grads_collapsed(ffvd_ctx*, int kind, const ffvd_problem* p, int flags, double jitter,
                             const ffvd_outputs* o);
/* B independent problems (own Z,U,hypers,Y,T; equal M,D,Din,Dy) in one launch sequence --
 * BASELINE config 4 (many small chains). */
int ffvd_nll_grads_batched(ffvd_ctx*, int kind, int collapsed, int nprob, const ffvd_problem* p,
                           int flags, double jitter, const ffvd_outputs* o);

/* The statistics FFVD_FLAG_COLLAPSED_P1_ONLY left in the context: S = F^T F as (nb, Mp, Mp) and b = F^T delta as (nb, Mp),
 * nb = S*D matrices, zero padded to Mp (ffvd_collapsed_stats_shape).  _allreduce sums them over the ranks in place (one
 * ncclAllReduce); _get / _set copy them out / in for callers with their own transport.  conditionals_multi_output.py:246-251. */
int ffvd_collapsed_stats_shape(ffvd_ctx*, int* nb, int* Mp);
int ffvd_collapsed_stats_allreduce(ffvd_ctx*);
int ffvd_collapsed_stats_get(ffvd_ctx*, DLManagedTensor* S_out, DLManagedTensor* b_out);
int ffvd_collapsed_stats_set(ffvd_ctx*, DLManagedTensor* S_in, DLManagedTensor* b_in);

/* ---- CUDA graphs.  The reference issues 22 session.run calls per outer iteration (base_model.py:915-933, 944-950); a
 * captured sequence of ffvd_nll_grads_* / ffvd_sghmc_update / ffvd_adam_update calls replays as ONE launch, which is what
 * matters on the bundled data (an evaluation there is ~10 launches of a few microseconds).  Rules while capturing: device
 * tensors only, FFVD_FLAG_ASYNC on every evaluation, g_X present, and the same call made once before the capture (so the
 * workspace exists).  Tensors are captured by ADDRESS: update their contents in place between launches. */
typedef struct ffvd_graph ffvd_graph;
int ffvd_graph_capture_begin(ffvd_ctx*);
int ffvd_graph_capture_end(ffvd_ctx*, ffvd_graph** out);
int ffvd_graph_launch(ffvd_ctx*, ffvd_graph*);            /* into the context's stream */
int64_t ffvd_graph_kernel_count(ffvd_graph*);             /* kernels one launch replays */
int ffvd_graph_destroy(ffvd_ctx*, ffvd_graph*);

/* ---- multi-GPU (the reference is single-process, so these replace nothing in it: they are what a sharded
 * caller of dgp_model.py:248-297 needs).  One NCCL communicator per context; NCCL is loaded at run time (libnccl.so.2, or
 * FFVD_NCCL_LIB) and shared with whatever already loaded it (PyTorch).  All collectives run on the context's stream. */
/* rank 0: fill a 128-byte NCCL unique id; the caller ships it to the other ranks (MPI, torch.distributed, a file) */
int ffvd_comm_unique_id(void* id128);
int ffvd_comm_init(ffvd_ctx*, const void* id128, int rank, int nranks);
int ffvd_comm_destroy(ffvd_ctx*);
int ffvd_comm_info(ffvd_ctx*, int* rank, int* nranks, int* nccl_version);
/* In-place sum over the ranks of the shared-parameter gradients of one evaluation (g_Z, g_U, g_logv, g_logl, g_logQ, g_C,
 * g_d, g_logR; with_scalars != 0: also nll and terms, for time-sharded evaluations) in ONE ncclAllReduce on a packed
 * buffer; g_X is never communicated.  Device tensors only.  A context without a communicator (or nranks == 1) returns
 * immediately. */
int ffvd_allreduce_shared(ffvd_ctx*, const ffvd_outputs* o, int with_scalars);
/* In-place sum over the ranks of one device tensor (the collapsed bound's F^T F / F^T delta statistics under time
 * sharding, conditionals_multi_output.py:246-254). */
int ffvd_allreduce(ffvd_ctx*, DLManagedTensor* t);

/* base_model.py:143-179 generate_update_step: one adaptive SG-HMC update, in place, Jacobi
 * semantics.  All tensors same shape.  noise ~ N(0,1) supplied by the caller (base_model.py:171).
 * burn_in=1 also updates xi,g,g2 (burn_in_op :179); burn_in=0 is sample_op (:178). */
int ffvd_sghmc_update(ffvd_ctx*, DLManagedTensor* theta, DLManagedTensor* grad, DLManagedTensor* noise,
                      DLManagedTensor* xi, DLManagedTensor* g, DLManagedTensor* g2, DLManagedTensor* p,
                      double epsilon, double mdecay, double X_N, int burn_in);

/* dgp_model.py:303-305 tf.compat.v1.train.AdamOptimizer apply (step is 1-based). */
int ffvd_adam_update(ffvd_ctx*, DLManagedTensor* theta, DLManagedTensor* grad, DLManagedTensor* m,
                     DLManagedTensor* v, double lr, double beta1, double beta2, double eps, int64_t step);

#ifdef __cplusplus
}
#endif
#endif /* FFVD_B200_H_ */
