/* ddb.h — C ABI of the B200-native sparse-identification backend for DataDrivenDiffEq.jl.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own — it is
 * pure Julia — so each entry point below names the reference code it REPLACES (paths relative to the
 * reference checkout).  A Julia maintainer binds these with `ccall` from two methods
 * (`DataDrivenDiffEq.get_fit_targets`, src/commonsolve.jl:74-81 and `CommonSolve.solve!`,
 * lib/DataDrivenSparse/src/commonsolve.jl:1-24); see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - All matrices are column-major (Julia layout).  States are n x m (one sample per column, so one
 *     sample's values are contiguous), targets n_t x m, coefficients n_t x k.
 *   - Every function returns DDB_OK (0) or a negative ddb_status; the message of the last error of
 *     the calling thread is available from ddb_last_error().  No C++ exception crosses the boundary.
 *   - The caller owns every host buffer and every device buffer it passes in and keeps it alive for
 *     the duration of the call; the library neither frees nor retains them after returning, except
 *     ddb_bind_device (retains the two data pointers until the next bind/upload/destroy).
 *   - A ddb_ctx is bound to ONE GPU and is not thread-safe; distinct contexts are independent (one host thread
 *     per context is fine).  Several GPUs: either one process per GPU with ddb_comm_init_rank, or one process
 *     with a ddb_group (N contexts + NCCL communicators, driven by one call).  Calls are synchronous.
 *   - There is NO CPU fallback: without a CUDA device every entry point that computes fails with
 *     DDB_CUDA_ERROR.
 */
#ifndef DDB_H
#define DDB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddb_ctx ddb_ctx;

typedef enum ddb_status {
    DDB_OK = 0,
    DDB_INVALID_ARG = -1,
    DDB_UNSUPPORTED = -2,
    DDB_CUDA_ERROR = -3,
    DDB_NCCL_ERROR = -4,
    DDB_NOT_POSITIVE_DEFINITE = -5,
    DDB_NONFINITE_INPUT = -6,
    DDB_OUT_OF_MEMORY = -7,
    DDB_BAD_STATE = -8
} ddb_status;

/* Which sparse-regression algorithm the sweep runs (lib/DataDrivenSparse/src/algorithms/). */
typedef enum ddb_algorithm {
    DDB_ALG_STLSQ = 0, /* STLSQ.jl:48-140 */
    DDB_ALG_ADMM = 1,  /* ADMM.jl:32-139  */
    DDB_ALG_SR3 = 2    /* SR3.jl:45-141   */
} ddb_algorithm;

/* Proximal operator of SR3 (algorithms/proximals.jl). STLSQ and ADMM always use SOFT's mask. */
typedef enum ddb_proximal {
    DDB_PROX_SOFT = 0, /* SoftThreshold            proximals.jl:36-65   */
    DDB_PROX_HARD = 1, /* HardThreshold            proximals.jl:88-117  */
    DDB_PROX_CAD = 2   /* ClippedAbsoluteDeviation proximals.jl:156-195 */
} ddb_proximal;

/* Model-selection score evaluated after every step (options.selector, solver.jl:100). */
typedef enum ddb_selector {
    DDB_SEL_BIC = 0, /* default, src/utils/common_options.jl:48 */
    DDB_SEL_AIC = 1,
    DDB_SEL_AICC = 2,
    DDB_SEL_RSS = 3
} ddb_selector;

/* Mirrors the fields of DataDrivenCommonOptions (src/utils/common_options.jl:22-54) and of the
 * algorithm structs that the sweep reads. */
typedef struct ddb_sweep_params {
    int32_t algorithm; /* ddb_algorithm */
    int32_t proximal;  /* ddb_proximal (SR3 only) */
    int32_t selector;  /* ddb_selector */
    int32_t maxiters;  /* per threshold (solver.jl:95); reference default 100 */
    double abstol;     /* reference default sqrt(eps) */
    double reltol;     /* reference default sqrt(eps) */
    double rho;        /* STLSQ ridge (0 = plain least squares) / ADMM rho / SR3 nu */
    double cad_rho;    /* ClippedAbsoluteDeviation upper cut; NaN = 5*lambda (proximals.jl:169) */
} ddb_sweep_params;

/* Stage timings of the last ddb_solve_sparse / stage calls, milliseconds, CUDA events. */
typedef struct ddb_timings {
    double h2d_ms;
    double gram_ms;
    double factor_ms;
    double sweep_ms;
    double rss_ms;
    double select_ms;
    double d2h_ms;
    int32_t gram_launches;  /* kernels launched by the Gram stage */
    int32_t sweep_launches; /* kernels launched by the sweep + rss + select stages */
    int32_t big_steps;      /* sweep steps that needed the blocked (global-memory) Cholesky */
    int32_t candidates;     /* unique (equation, support, coefficients) candidates scored */
    double gram_dmma_flops; /* FP64 tensor-core flops the last Gram stage issued PER SAMPLE (512 per DMMA.8x8x4);
                               below the algorithmic k(k+1) + 2 k n_t + 2 n_t when the moment schedule ran */
    int32_t gram_schedule;  /* schedule of the last Gram stage: 0 self-building, 1 ring, 2 moment, 3 single-pair libraries (k + n_t <= 64)
                               of at most 3 states and degree >= 3: four balanced warps, builder reads a per-stage
                               table of powers, 4 other single-pair libraries: four balanced warps */
    int32_t allreduce_calls; /* NCCL all-reduces queued by the last sharded solve / since the last reset */
    double allreduce_ms;    /* their duration on the compute stream (includes waiting for the slowest rank) */
    double allreduce_bytes; /* payload per rank */
} ddb_timings;

/* ---- context ---------------------------------------------------------------------------------- */

/* Create a context on CUDA device `device`.  Owns its stream, workspaces and tables. */
int ddb_ctx_create(int device, ddb_ctx** out);
int ddb_ctx_destroy(ddb_ctx* ctx);
const char* ddb_last_error(void);
/* Library version string, e.g. "ddb200 0.2 (sm_100a)". */
const char* ddb_version(void);
/* Debugging aid: with DDB_GUARD=1 in the environment every device allocation of the library is bracketed by guard
 * zones; returns how many live allocations had theirs overwritten (0 = clean; always 0 without DDB_GUARD). */
int ddb_debug_check_guards(void);

/* ---- candidate library (replaces Basis code generation + evaluation) ------------------------- */

/* Install a monomial library: k terms over n states, exps is n x k column-major,
 * exps[i + n*j] = power of state i in term j, in the reference's term order
 * (polynomial_basis / monomial_basis, src/utils/basis_generators.jl:84-149; evaluated by
 * (b::Basis)(prob), src/problem/type.jl:404-407 -> src/basis/build_function.jl:186-205).
 * Total degree of a term must be <= 8. */
int ddb_set_library_monomial(ddb_ctx* ctx, int n, int k, const uint8_t* exps);

/* ---- data ------------------------------------------------------------------------------------- */

/* Copy host data to the device (replaces nothing in the reference; it is the host->device edge).
 * X: n x m, Y: n_t x m, column-major.  m is this context's LOCAL sample count (its row shard). */
int ddb_upload(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m);

/* Use data already resident on the device (device pointers, 16-byte aligned). */
int ddb_bind_device(ddb_ctx* ctx, const double* dX, const double* dY, int n_t, int64_t m);

/* Generate synthetic trajectories on the device and bind them.
 * system_id: 0 = Lorenz (n=3), 1 = Lorenz-96 (n = n_states).  Y = exact right-hand side.
 * noise_std > 0 adds N(0, noise_std^2) to Y.  Sample index offset `first_sample` lets each rank
 * generate its own shard of one global data set. */
int ddb_generate(ddb_ctx* ctx, int system_id, int n_states, int64_t m, int64_t first_sample,
                 uint64_t seed, double noise_std);

/* ---- feature map: libraries of sin / cos / exp / Chebyshev terms and their products with monomials ----
 * Replaces the non-polynomial generators `sin_basis`, `cos_basis`, `fourier_basis`, `chebyshev_basis`
 * (src/utils/basis_generators.jl:30-82) and libraries mixing them with polynomials (e.g.
 * `[polynomial_basis(u, 5); sin.(u); cos.(u)]`, lib/DataDrivenSparse/test/Core/pendulum.jl:13).
 * Feature f is  g_f(coef[f] * x[source[f]])  of the n_raw raw states; the monomial library
 * (ddb_set_library_monomial with n = n_feat) is then a table of exponents over the FEATURES.  After this call
 * ddb_upload / ddb_upload_gram / ddb_bind_device take RAW states (n_raw x m) and expand them once on the
 * device (n_feat x m doubles stay resident; Theta itself is still never built).  n_feat = 0 removes the map.
 * Call before binding data. */
typedef enum ddb_feature_kind {
    DDB_FEAT_IDENTITY = 0,  /* coef * x            */
    DDB_FEAT_SIN = 1,       /* sin(coef * x)       sin_basis, odd fourier_basis terms (coef = t/2) */
    DDB_FEAT_COS = 2,       /* cos(coef * x)       cos_basis, even fourier_basis terms             */
    DDB_FEAT_EXP = 3,       /* exp(coef * x)                                                        */
    DDB_FEAT_CHEBYSHEV = 4  /* cos(coef * acos(x)) chebyshev_basis                                  */
} ddb_feature_kind;
int ddb_set_features(ddb_ctx* ctx, int n_raw, int n_feat, const int32_t* kind, const int32_t* source, const double* coef);

/* Arbitrary features through NVRTC: `body` is CUDA C that assigns f[0] .. f[n_feat - 1] from the raw variables
 * x[0] .. x[n_raw - 1] of one sample ([states | controls | independent variable], as the caller stacks them) and the
 * parameter values p[0] .. p[n_params - 1], e.g. "f[0] = x[0]; f[1] = sin(x[0]) * x[1]; f[2] = exp(-p[0] * x[1] * x[1]);".
 * One kernel is compiled for sm_100a and evaluated once per upload / bind; the installed monomial library is a table
 * of exponents over these features (identity table = one term per expression).  Replaces the per-term code generation
 * of `DataDrivenFunction` (src/basis/build_function.jl:6-38, argument order implicits, states, parameters,
 * independent variable, controls) for terms the fixed kinds above do not cover; the Julia side can emit `body` with
 * `Symbolics.build_function(...; target = Symbolics.CTarget())`.  Call before binding data; ddb_set_features replaces it.
 * ddb_features_source_check only compiles (no GPU needed) and returns the compiler log. */
int ddb_set_features_source(ddb_ctx* ctx, int n_raw, int n_feat, const char* body, const double* params, int n_params);
int ddb_features_source_check(const char* body, char* log, int log_cap);

/* Copy the bound device data back to the host (for tests and the end-to-end bench). */
int ddb_download_data(ddb_ctx* ctx, double* X, double* Y);

/* ---- stage 1+2: fused library evaluation + normal-equation blocks ----------------------------- */

/* Number of doubles in the packed normal-equation buffer [G (k*k) | R (k*n_t) | yy (n_t)]. */
int64_t ddb_gram_count(const ddb_ctx* ctx);

/* Compute this context's partial blocks from its bound samples:
 *   G = Theta Theta' (k x k), R = Theta Y' (k x n_t), yy[i] = sum_s Y[i,s]^2,
 * Theta never written to memory.  Replaces the per-target `A*A'`, `B*A'` of
 * STLSQ.jl:90-91 / ADMM.jl:77-79 / SR3.jl:108-109 and the library evaluation feeding them.
 * dPacked: device buffer of ddb_gram_count() doubles, or NULL to use the context's own buffer.
 * G is exactly symmetric and the result is bit-reproducible from call to call.  For monomial libraries the
 * entries are contracted moment by moment (see ddb_moment_plan_describe): G[t,u] is the sum over the samples of
 * the product of two monomials whose product is theta_t theta_u, not necessarily theta_t and theta_u themselves,
 * so it differs from the plain contraction in the last bits (ddb_timings.gram_schedule says which ran). */
int ddb_gram(ddb_ctx* ctx, double* dPacked);

/* ddb_upload + ddb_gram in one call with the host->device copies OVERLAPPED by the Gram kernels: the samples
 * are cut into chunks (<= 16, >= ~64 MB each), chunk c + 1 travels on a copy stream while the kernels run on
 * chunk c, and the per-chunk blocks are added in chunk order (deterministic).  The data stay resident as
 * after ddb_upload.  Page-locked host buffers are copied by the DMA engine directly; PAGEABLE buffers (a Julia Matrix,
 * a NumPy array) go through the library's own page-locked bounce ring (3 x 32 MB, filled by a helper thread and a
 * small pool -- DDB_STAGE_THREADS, default half the cores / ranks, at most 8), which keeps the transfer behind the
 * Gram kernels as well (C3, 10 GB: 727 ms per solve against 718 ms from pinned memory).  This is what ddb_solve_sparse
 * uses. */
int ddb_upload_gram(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m, double* dPacked);

/* Device pointer of the context's own packed buffer (for an all-reduce by the caller:
 * torch.distributed / NCCL over NVLink; one sum all-reduce of ddb_gram_count() doubles). */
int ddb_gram_buffer(ddb_ctx* ctx, double** dPacked);

/* Copy the packed blocks to the host; any output may be NULL.  G k x k, R k x n_t, yy n_t. */
int ddb_gram_download(ddb_ctx* ctx, double* G, double* R, double* yy);
/* Install blocks computed elsewhere (restart, tests, or the all-reduced sum of the ranks' partial
 * blocks).  G, R, yy may be host or device pointers. */
int ddb_gram_upload(ddb_ctx* ctx, const double* G, const double* R, const double* yy, int n_t);

/* ---- pre-/post-processing options of the reference on the device ---------------------------------
 * ddb_set_sample_range: the following ddb_gram / ddb_rss / ddb_residual calls see samples [first, last) of the
 *   bound data only (last < 0 = to the end; first == last is an EMPTY view -- a rank of a sharded run that holds no
 *   sample of a batch or of the test part: every pass then contributes zeros to the sums over the ranks; any first /
 *   last otherwise: a view that does not start on a 16-byte boundary --
 *   odd first with an odd n or n_t -- is staged once into an aligned buffer for the bulk copies of the Gram
 *   kernels).  Replaces `splitobs(data, at = split)` and the
 *   batches of the `DataLoader` (src/utils/data_processing.jl:29-39): train part, test part and every batch are
 *   views of the resident data.  A new upload / bind resets the view.
 * ddb_permute_samples: reorders the resident samples, new sample s = old sample perm[s] (host array of m
 *   indices); the `shuffle = true` of the DataLoader (data_processing.jl:38).
 * ddb_gram_scale_terms: divides the blocks in the context by per-term scales (G_ij / (s_i s_j), R_ie / s_i):
 *   the regression then runs on Theta_j / s_j as after `apply_transform!(ZScoreTransform(center = false))`
 *   (data_processing.jl:64-80, src/commonsolve.jl:128-130) -- thresholds act on the scaled coefficients, the
 *   exact-rss pass divides the candidates by the scales, ddb_select returns SCALED coefficients (the caller
 *   maps them back, lib/DataDrivenSparse/src/commonsolve.jl:19-21).  Reset by a new Gram.
 * ddb_residual: rss[e] = sum_s (Y[e,s] - sum_j coef[e,j] theta_j(x(s)))^2 over the current sample view for an
 *   arbitrary coefficient matrix (n_t x k column-major, host): test error (commonsolve.jl:42) and the rss that
 *   `DataDrivenSolution` recomputes on the CPU (src/solution.jl:37-56). */
int ddb_set_sample_range(ddb_ctx* ctx, int64_t first, int64_t last);
int ddb_permute_samples(ddb_ctx* ctx, const int64_t* perm);
int ddb_gram_scale_terms(ddb_ctx* ctx, const double* scale);
int ddb_residual(ddb_ctx* ctx, const double* coef, double* rss);
/* `UnitRangeTransform` normalisation of the library rows (src/utils/data_processing.jl:68-73, 119-126):
 * ddb_term_minmax: minimum and maximum of every library term over the current sample view (host outputs, k each) --
 *   what `fit(UnitRangeTransform, Theta, dims = 2)` needs; one pass over the states, Theta is not built.
 * ddb_gram_affine_terms: turns the blocks in the context into those of theta~_j = (theta_j - shift_j) / scale_j
 *   (G~ = S (G - mu t' - t mu' + m mu mu') S, R~ = S (R - mu ysum'); tsum = Theta 1 (k), ysum = Y 1 (n_t), m_total
 *   samples behind the blocks).  The sweep then runs in the transformed space, the exact-rss pass evaluates every
 *   candidate on the raw data with the equivalent coefficients and constant.  Reset by a new Gram.
 * ddb_residual_affine: ddb_residual with a constant per target added to the fit (offset, n_t): residuals of a model
 *   given in the transformed space. */
int ddb_term_minmax(ddb_ctx* ctx, double* mn, double* mx);
/* css[j] = sum over the current sample view of (theta_j - center[j])^2 for every library term (host arrays, k each):
 * the second pass of the two-pass standard deviation `fit(ZScoreTransform, Theta, dims = 2)` takes on the materialised
 * Theta (src/utils/data_processing.jl:75-80); center = (sum theta_j) / m from the Gram blocks.  One HBM pass over the
 * states, Theta is not built; per-CTA partial sums added in a fixed order (bit-reproducible); sum over the ranks of a
 * sharded run. */
int ddb_term_css(ddb_ctx* ctx, const double* center, double* css);
int ddb_gram_affine_terms(ddb_ctx* ctx, const double* shift, const double* scale, const double* tsum, const double* ysum,
                          int64_t m_total);
int ddb_residual_affine(ddb_ctx* ctx, const double* coef, const double* offset, double* rss);
/* Per-target mean and centred sum of squares of the bound targets over the current sample view:
 *   mean[e] = (sum over ALL ranks of sum_s Y[e,s]) / m_total,   css[e] = sum_s (Y[e,s] - mean[e])^2  (summed over ranks),
 * two HBM passes over Y; with a communicator both sums are all-reduced inside (collective).  What
 * `nullloglikelihood` / `r2` of a solution need besides rss (src/solution.jl:139-157:
 * `sum(abs2, Y .- mean(Y, dims = 2))` = sum_e css[e]); replaces their CPU pass over Y.  mean, css: host, n_t each. */
int ddb_target_stats(ddb_ctx* ctx, int64_t m_total, double* mean, double* css);

/* ---- stage 3: thresholded sweep ---------------------------------------------------------------- */

/* ---- implicit sparse regression (ImplicitOptimizer) ----------------------------------------------
 * Replaces the per-candidate loop of `(x::ImplicitOptimizer)(X, Y)` (lib/DataDrivenSparse/src/algorithms/
 * Implicit.jl:57-108): `solver(X[inds, :], X[i, :])` for every library row i.  Every one of those regressions
 * is a sub-block of the SAME Gram matrix, so no further pass over the data is needed.
 *
 * ddb_implicit_select restricts the following ddb_sweep / ddb_rss / ddb_select calls to the library rows
 * idx[0..kk) (host array; indices into the installed library, distinct) and makes EACH of them a target that
 * is regressed on the other kk - 1 selected rows: afterwards "k" = "n_t" = kk for the outputs of ddb_select
 * (coef is kk x kk column-major, coef[e + kk*j] = coefficient of selected row j in the model of row e, zero on
 * the diagonal; rss[e] = sum_s (theta_e(s) - sum_j coef[e,j] theta_j(s))^2 by the exact streaming pass).
 * Needs the normal-equation blocks of the full library in the context (ddb_gram / ddb_gram_upload).
 * kk = 0 returns to the explicit mode; a new Gram, data set or library does so as well.  The choice of the
 * best left-hand side (dof >= 2, necessary terms, smallest AICc; Implicit.jl:87-98) is host logic. */
int ddb_implicit_select(ddb_ctx* ctx, const int32_t* idx, int kk);

/* Run the warm-started threshold sweep of SparseLinearSolver (solver.jl:72-118) for all n_t targets
 * on the normal-equation blocks currently in the context (after ddb_gram [+ all-reduce]).
 * lambdas: L thresholds exactly as the algorithm struct stores them (for SR3+HardThreshold the
 * caller passes thr^2/2, SR3.jl:63), sorted ascending by the library if they are not.
 * m_total: number of observations over ALL ranks (nobs of the statistics).
 * Produces the device-resident candidate list; nothing is returned to the host yet. */
int ddb_sweep(ddb_ctx* ctx, const ddb_sweep_params* params, const double* lambdas, int L,
              int64_t m_total);

/* Number of doubles of the candidate-rss table (one per unique candidate + n_t for the zero model). */
int64_t ddb_rss_count(const ddb_ctx* ctx);
/* Exact residual sums of squares of every candidate over this context's bound samples (second
 * streaming pass; replaces rss(cache), DataDrivenSparse.jl:173-176, called by the selector at
 * solver.jl:100).  dRss: device buffer of ddb_rss_count() doubles or NULL for the context's own.
 * The table holds the STREAMED part of every candidate's rss, a plain sum over samples (so the tables of the
 * ranks of a sharded run are combined with one sum all-reduce before ddb_select): the value of the candidate's
 * representative -- candidates that share a (target, support) are represented by the least-squares point c0 of
 * that support, zero models by nothing (0).  The remainder, rss(c) - rss(c0) = (c - c0)' G_SS (c - c0) -
 * 2 (c - c0)' (r_S - G_SS c0) resp. yy, is exact Gram arithmetic and is added by ddb_select. */
int ddb_rss(ddb_ctx* ctx, double* dRss);
int ddb_rss_buffer(ddb_ctx* ctx, double** dRss);

/* Apply the selector along each target's step sequence (solver.jl:100-107 semantics incl. the
 * `<=` tie-break and the `j == 1` rule) and copy the results to the host.  Outputs (any may be
 * NULL): coef n_t x k column-major = SparseRegressionResult.coefficients (un-rounded);
 * lambda_opt n_t; iters n_t (total iteration counter, solver.jl:117); rss n_t; dof n_t;
 * support_path n_t x L x ceil(k/32) uint32 words (converged support at every threshold,
 * index ((i*L + j)*W + w)); coef_path n_t x L x k doubles (coefficients of the cache when the
 * inner loop of threshold j stopped, index ((i*L + j)*k + t)); iters_path n_t x L;
 * flags n_t (bit 0: a factorisation broke down for that target and no fallback applied -- ADMM / SR3, non-finite
 * blocks; bit 1: STLSQ took minimum-norm steps for it, i.e. some masked system was rank deficient -- fewer samples than
 * active terms or collinear terms -- and was solved as `A[p, :]' \ b` solves it in the reference, STLSQ.jl:93-99:
 * Jacobi pseudo-inverse in shared memory up to 64 active terms, symmetric eigen-decomposition in global memory beyond). */
int ddb_select(ddb_ctx* ctx, double* coef, double* lambda_opt, int32_t* iters, double* rss,
               int32_t* dof, uint32_t* support_path, double* coef_path, int32_t* iters_path,
               int32_t* flags);

/* ---- one-call host path (what `solve!` ccalls on a single GPU) -------------------------------- */

/* upload + gram + sweep + rss + select, host buffers in, host buffers out.  Replaces
 * `alg(X, Y; options)` (DataDrivenSparse.jl:258-277) together with the library evaluation that
 * `init` performs before it (src/commonsolve.jl:92-139).  trainerror (n_t... summed to a scalar by
 * the caller) is rss of the selected model per target = sum(abs2, Y - C*Theta) rows
 * (lib/DataDrivenSparse/src/commonsolve.jl:37).  flags (n_t, may be NULL): bit 0 = a factorisation broke down for
 * that target (the caller maps it to DDReturnCode(2), src/DataDrivenDiffEq.jl:60-67). */
int ddb_solve_sparse(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m,
                     const ddb_sweep_params* params, const double* lambdas, int L, double* coef,
                     double* lambda_opt, int32_t* iters, double* rss, int32_t* dof,
                     uint32_t* support_path, int32_t* flags);

/* ---- multi-GPU: the row-sharded path ------------------------------------------------
 * Rows (samples) shard across the GPUs; the path has exactly two exchange steps, each ONE
 * `ncclAllReduce(sum, ncclDouble)` in place on the context's compute stream (no host synchronisation between the
 * producing kernel, the collective and its consumer): the packed [G | R | yy] after the Gram stage and the
 * candidate-rss table after the streaming rss pass.  The sweep runs replicated (deterministic, identical blocks
 * on every rank; candidates are numbered canonically so that entry c of the table is the same model everywhere).
 * Replaces nothing in the reference (it is single-process CPU code); the hook a maintainer binds is the same
 * `CommonSolve.solve!` method as on one GPU (lib/DataDrivenSparse/src/commonsolve.jl:1-24), with
 * `B200Sparse(inner; ngpus = N)` holding a ddb_group.  NCCL is loaded at run time (libnccl.so.2) on first use. */
#define DDB_UNIQUE_ID_BYTES 128
/* One process (or thread) per GPU: rank 0 creates the id, the host program carries its 128 bytes to the other
 * ranks (MPI, torch.distributed, a file ...), then every rank attaches a communicator to its context. */
int ddb_comm_unique_id(void* id128);
int ddb_comm_init_rank(ddb_ctx* ctx, int nranks, int rank, const void* id128);
/* nranks / rank of the context's communicator (1 / 0 without one) and the version of the loaded NCCL (0 if none). */
int ddb_comm_info(const ddb_ctx* ctx, int* nranks, int* rank, int* nccl_version);
/* In-place sum over the ranks, queued on the compute stream; no-ops without a communicator.
 * ddb_allreduce_gram: the context's own packed blocks (after ddb_gram / ddb_upload_gram with dPacked = NULL);
 * ddb_allreduce_rss: the context's own candidate-rss table (after ddb_rss with dRss = NULL);
 * ddb_allreduce_sum: any device buffer of `count` doubles (e.g. the [P | Q] blocks of the DMD path). */
int ddb_allreduce_gram(ddb_ctx* ctx);
int ddb_allreduce_rss(ddb_ctx* ctx);
int ddb_allreduce_sum(ddb_ctx* ctx, double* dBuf, int64_t count);
/* ddb_solve_sparse for ONE RANK of a sharded run (collective: every rank calls it with its own host shard
 * X n x m_local, Y n_t x m_local and the same library / parameters / thresholds; m_total = samples of all ranks;
 * every rank receives the full result).  flags (n_t, may be NULL) as in ddb_select. */
int ddb_solve_sparse_sharded(ddb_ctx* ctx, const double* X, const double* Y, int n_t, int64_t m_local, int64_t m_total,
                             const ddb_sweep_params* params, const double* lambdas, int L, double* coef,
                             double* lambda_opt, int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path,
                             int32_t* flags);

/* One process driving N GPUs of a node: N contexts + `ncclCommInitAll` (devices = NULL means 0 .. ngpus-1).
 * ddb_group_solve_sparse is the N-GPU form of ddb_solve_sparse: it cuts the host X / Y into N contiguous sample
 * ranges (interior boundaries on multiples of 16 samples, ddb_group_shard), runs one host thread per GPU through
 * ddb_solve_sparse_sharded and returns rank 0's (= every rank's) result.  m >= 16 * ngpus. */
typedef struct ddb_group ddb_group;
int ddb_group_create(int ngpus, const int* devices, ddb_group** out);
int ddb_group_destroy(ddb_group* g);
int ddb_group_size(const ddb_group* g);
ddb_ctx* ddb_group_ctx(ddb_group* g, int rank); /* borrowed; owned by the group */
int ddb_group_set_library_monomial(ddb_group* g, int n, int k, const uint8_t* exps);
int ddb_group_set_features(ddb_group* g, int n_raw, int n_feat, const int32_t* kind, const int32_t* source, const double* coef);
int ddb_group_shard(const ddb_group* g, int64_t m, int rank, int64_t* first, int64_t* last);
int ddb_group_solve_sparse(ddb_group* g, const double* X, const double* Y, int n_t, int64_t m,
                           const ddb_sweep_params* params, const double* lambdas, int L, double* coef,
                           double* lambda_opt, int32_t* iters, double* rss, int32_t* dof, uint32_t* support_path,
                           int32_t* flags);
/* ddb_target_stats over the shards the last ddb_group_solve_sparse left resident (one host thread per GPU). */
int ddb_group_target_stats(ddb_group* g, double* mean, double* css);
/* Stage timings of the last group solve: rank 0's, with the per-rank stages replaced by the slowest rank's. */
int ddb_group_get_timings(const ddb_group* g, ddb_timings* out);

/* ---- DataDrivenDMD Gram path ------------------------------------------------------------------- */

/* P = X X' and Q = Y X' (lib/DataDrivenDMD/src/solve.jl:128-129) for snapshot matrices X, Y
 * (n x m, device pointers; for a discrete problem Y = X shifted by one column, pass dY = dX + n).
 * dP, dQ: n x n device buffers (column-major), partial sums of this context's samples. */
int ddb_dmd_gram(ddb_ctx* ctx, const double* dX, const double* dY, int n, int64_t m, double* dP,
                 double* dQ);

/* K = Q P^-1 (DMDPINV, lib/DataDrivenDMD/src/algorithms.jl:80-83, K = Y / X through the normal
 * equations) from reduced device blocks dP, dQ; K n x n written to dK (device). */
int ddb_dmd_operator(ddb_ctx* ctx, const double* dP, const double* dQ, int n, double* dK);

/* Host-buffer form of the two calls above for a discrete problem with the identity lifting: snapshots
 * x (n x (m + 1), column-major, one snapshot per column; X = x[:, 1:m], Y = x[:, 2:m+1] as in
 * lib/DataDrivenDMD/src/solve.jl:36-65) -> P = X X', Q = Y X' and K = Q P^-1 (DMDPINV).  P, Q, K are n x n
 * column-major host buffers, each may be NULL.  n must be even.  This is the entry point the Julia glue
 * binds (`B200DMD`); DMDSVD / TOTALDMD / FBDMD follow from P, Q (and S = Y Y') on the host side. */
int ddb_dmd_solve(ddb_ctx* ctx, const double* x, int n, int64_t m_plus_1, double* P, double* Q, double* K);

/* Exact residual of a fitted linear model over resident snapshots: rss = || Z - M X ||_F^2 with X (n_in x m) and
 * Z (n_out x m) device, column-major (one snapshot per column) and M (n_out x n_in) device, column-major.  Replaces
 * `sum(abs2, Z .- C * (K * X + B * U))` behind `rss(::KoopmanResult)` (lib/DataDrivenDMD/src/result.jl:46-59; stack
 * [X; U] and [K B] for a controlled system); the prediction is formed tile by tile on the tensor pipe and never
 * written to memory.  rss: host scalar; sum it over the ranks of a sharded run. */
int ddb_dmd_residual(ddb_ctx* ctx, const double* dX, const double* dZ, const double* dM, int n_in, int n_out, int64_t m,
                     double* rss);

/* Dense eigen-decompositions of the DMD algorithms (cuSOLVER Xsyevd / Xgeev, bound at run time; device buffers,
 * column-major):
 * ddb_dmd_eig_sym: symmetric dA (n x n, lower triangle read) -> eigenvalues ascending in dW (n), eigenvectors
 *   overwrite dA.  `truncated_svd` through the Gram identity P = X X' = U S^2 U' (algorithms.jl:2-20).
 * ddb_dmd_eig: general real dA (n x n, destroyed) -> dW = n complex eigenvalues (re, im interleaved: 2 n doubles) and,
 *   unless dVR is NULL, the right eigenvectors in LAPACK's packed real form (n x n: a real eigenvalue's vector is one
 *   column, a complex pair j, j + 1 is stored as columns Re v, Im v with v_j = Re + i Im, v_{j+1} = conj).
 *   `eigen(K)` / `eigen(Atilde)` (algorithms.jl:81, 157). */
int ddb_dmd_eig_sym(ddb_ctx* ctx, double* dA, int n, double* dW);
int ddb_dmd_eig(ddb_ctx* ctx, double* dA, int n, double* dW, double* dVR);

/* D (rows x cols) = A (rows x K) * B (cols x K)' for ROW-major device matrices with the given leading dimensions --
 * the FP64 DMMA product the library's own substitution and operator steps use, exposed for the small dense
 * compositions of the DMD variants (U' Q U S^-2, TOTALDMD projections; algorithms.jl:147-158, 215-248).
 * A column-major n x m matrix is a row-major m x n one, so  C = A B  in Julia's layout is this call with the roles
 * of the operands exchanged. */
int ddb_gemm_nt(ddb_ctx* ctx, double* dD, int64_t ldd, const double* dA, int64_t lda, const double* dB, int64_t ldb, int rows,
                int cols, int K);

/* ---- introspection ----------------------------------------------------------------------------- */

/* Host-only: the work plan the Gram stage would use for an augmented library of kaug = k + n_t rows,
 * m local samples and nctas persistent CTAs (0 = choose as the library would for 148 SMs).  Writes up
 * to cap segments as 6 int64 each: {cta, block_row, block_col, slot, stage_begin, stage_end}; returns
 * the number of segments (or a negative status).  info (may be NULL) receives {block_tile,
 * samples_per_stage, blocks_per_dim, block_pairs, stages, slots}.  Needs no GPU. */
int64_t ddb_gram_plan_describe(int kaug, int64_t m, int nctas, int64_t* segments, int64_t cap, int64_t* info);

/* Host-only: the moment schedule of the Gram stage for a monomial library (exps as in ddb_set_library_monomial)
 * with n_t targets.  The Gram matrix of monomials is a moment matrix: G[t,u] only depends on the product
 * theta_t theta_u, so every distinct moment is contracted once, as (virtual row term) x (virtual column term),
 * and gathered into the packed [G (k x k) | R (k x n_t) | yy (n_t)].  info (8 int64, may be NULL) receives
 * {used, virtual_rows, virtual_columns, row_blocks, column_blocks, block_pairs, virtual_terms, degree}; `used`
 * = 0 means the library keeps the plain schedule (nothing else is written, returns 0).  Otherwise returns the
 * number of packed entries and writes, each up to its cap and each optional:
 *   src     per packed entry (pair << 14) | (row_in_tile * 128 + column_in_tile);
 *   factors (row_blocks + column_blocks) * 128 virtual terms x 8 factor codes: state index, 0x8000 | target
 *           index, 0xFFFF = no factor; block B holds the virtual terms [128 B, 128 B + 128); blocks
 *           [0, row_blocks) are the virtual rows, the others the virtual columns;
 *   pairs   block_pairs x {first_block, second_block, rows}: the pair's tile is (terms of first_block) x (terms
 *           of second_block); rows = 128, or 64 / 32 for a pair at the edge of the staircase whose needed entries
 *           all lie in the first 64 / 32 terms of first_block (such a pair may have a column block first).  Whole
 *           pairs come first, then the 64-row, then the 32-row ones.
 * Needs no GPU. */
int64_t ddb_moment_plan_describe(int n, int k, const uint8_t* exps, int n_t, int32_t* src, int64_t src_cap, uint16_t* factors,
                                 int64_t factors_cap, int32_t* pairs, int64_t pairs_cap, int64_t* info);

/* Host-only: the tensor-pipe builder of the moment schedule (csrc/gram.cu: gram_mb_kernel).  For libraries whose
 * virtual terms have at most two factors (total degree <= 2) the library values are not multiplied out on the FP64
 * pipe: one DMMA with a one-hot first operand forms 4 samples x 2 second factors x 8 first factors = 64 products,
 * each a single correctly rounded multiplication (same bits as the multiply it replaces).  Which products an op
 * forms is decided on the host: the terms of every virtual block are covered by (2 x 8)-factor rectangles.  Writes,
 * per block pair of ddb_moment_plan_describe and in the same order, optab: 8 warps x 12 op slots x 32 lanes x 2
 * words {A-operand offset | B-operand offset << 16, destination 0 | destination 1 << 16} (byte offsets inside a
 * raw-sample slot [ones(16) | 0.0 | pad | X tile 16 x n | Y tile 16 x n_t] resp. inside a stage of two 16 x 132
 * tiles; slot index = 3 * k-step + {0, 1, 2}), opcnt: 8 bytes per pair, bit kk set = the warp runs the third slot of
 * k-step kk, block_ops: ops per virtual block.  Returns the number of block pairs, 0 if the library keeps the
 * FP64-pipe builder (degree > 2, a pair that needs more than 12 ops per warp, or samples too wide).  Needs no GPU. */
int64_t ddb_moment_builder_describe(int n, int k, const uint8_t* exps, int n_t, uint32_t* optab, int64_t optab_cap,
                                    uint8_t* opcnt, int64_t opcnt_cap, int32_t* block_ops, int64_t block_cap);

/* Host-only: the work plan of the moment schedule for m local samples on nctas persistent CTAs (0 = 148): the block
 * pairs of ddb_moment_plan_describe run in lock-step waves, every pair over a number of equal sample ranges.
 * Writes up to cap segments as 7 int64 each {cta, first_block, second_block, slot, stage_begin, stage_end, rows}
 * (a stage is 16 samples; segments are listed CTA by CTA in execution order) and returns their number, 0 if the
 * library keeps the plain schedule.  info (4 doubles, may be NULL): {planned time in pair streams per CTA, the same
 * for a perfect split = pair units / nctas, waves, slots}.  Needs no GPU. */
int64_t ddb_moment_wave_describe(int n, int k, const uint8_t* exps, int n_t, int64_t m, int nctas, int64_t* segments,
                                 int64_t cap, double* info);

/* Host-only: the builder of the single-pair Gram kernel for libraries of at most 3 states and degree >= 3
 * (csrc/gram.cu: gram_pow_kernel).  Every 16-sample stage gets a table of powers x_i^e (e = 1 .. degree) and a library
 * value is the product of ONE table entry per state, x^a y^b z^c, instead of one staged state per unit of degree.
 * Writes 64 rows x 8 factor codes (library terms, then the targets): state * degree + exponent - 1 for every state with
 * a non-zero exponent, 0x8000 | target for a target row, 0xFFFF = no factor; *degree receives the library's total degree.
 * Returns the number of factors per term the kernel is instantiated for (1 .. 3), 0 when the builder does not apply
 * (more than 3 states, degree < 3, k + n_t > 64: nothing is written), negative on invalid arguments.  Needs no GPU. */
int ddb_pow_factors_describe(int n, int k, const uint8_t* exps, int n_t, uint16_t* factors, int* degree);

int ddb_get_timings(const ddb_ctx* ctx, ddb_timings* out);
/* The context's CUDA stream (cudaStream_t) so a caller can order its own work after ours. */
int ddb_get_stream(ddb_ctx* ctx, void** stream);
/* Block the host until everything queued on the context's stream has finished. */
int ddb_synchronize(ddb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DDB_H */
