/*
 * mgvi_b200.h -- C ABI of libmgvib200.so: the MGVI Fisher-CG / Newton-CG hot path on NVIDIA B200
 * (sm_100a), built to sit behind the Julia API of bat/MGVI.jl v0.4.5.
 *
 * The reference has no FFI of its own; its extension seams are Julia multiple-dispatch methods
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout) and is what a `ccall` from those methods binds to
 * (julia/MGVIB200.jl).
 *
 * Conventions
 *  - return value 0 = success; non-zero = error, message from mgvib200_last_error(ctx).
 *  - all arrays are float64 ("double") or int64_t, caller-owned HOST memory unless the function
 *    name ends in _dev (then: device memory of the context's GPU); the library never keeps a
 *    caller pointer after the call returns, never calls back into the host language.
 *  - matrices are column-major with the leading dimension equal to the row count (Julia
 *    layout): residual i is the contiguous column R[:, i].
 *  - field arrays are passed flat with `dims` in C order (dims[0] slowest).  A Julia caller
 *    passes reverse(size(field)).
 *  - every call is synchronous at return; one context is used from one host thread at a time
 *    (the reference's Threads.@threads over residuals, src/residual_samplers.jl:88, is replaced
 *    by in-library batching).
 *  - handles are opaque and owned by the library until the matching *_destroy.
 *  - the standard-normal draws are INPUTS, in the order the reference consumes them
 *    (src/residual_samplers.jl:75-76): per residual i, n1 = Z1[:, i] (n_lambda) then
 *    eta = Z2[:, i] (n_theta); dense sampler: one (n_theta x n) matrix Z (:121).
 */
#ifndef MGVI_B200_H
#define MGVI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgvib200_ctx mgvib200_ctx;
typedef struct mgvib200_model mgvib200_model;

#define MGVIB200_OK 0
#define MGVIB200_ERR_INVALID 1
#define MGVIB200_ERR_CUDA 2
#define MGVIB200_ERR_UNSUPPORTED 3
#define MGVIB200_ERR_NCCL 4

/* forward-model families (declarative: AD of arbitrary closures cannot cross a C ABI) */
#define MGVIB200_MODEL_FIELD_DHT 0     /* test/test_models/model_fft_gp.jl, docs/src/advanced_tutorial_lit.jl:253-307 */
#define MGVIB200_MODEL_POLY 1          /* test/test_models/model_polyfit.jl */
#define MGVIB200_MODEL_DENSE_LINEAR 2  /* dense Jacobian; operator type of MatrixInversion, src/residual_samplers.jl:56 */
#define MGVIB200_MODEL_FIELD_HYPER 3   /* docs/src/advanced_tutorial_lit.jl:194-307: hyperparameter field + bin aggregation */

#define MGVIB200_LINK_IDENTITY 0
#define MGVIB200_LINK_EXP 1

#define MGVIB200_LIKE_NORMAL 0         /* src/information.jl:12-18, 60-65 */
#define MGVIB200_LIKE_POISSON 1        /* src/information.jl:74-78 */

/* order of the flat distribution parameters lambda (src/shapes.jl:4-14,40-54; SURVEY.md note N1) */
#define MGVIB200_LAYOUT_MU_ONLY 0      /* lambda = [mu]                      (n_lambda = N)  */
#define MGVIB200_LAYOUT_BLOCKED 1      /* lambda = [mu(N); var(N)] per group (n_lambda = 2N) */
#define MGVIB200_LAYOUT_INTERLEAVED 2  /* lambda = [mu1, s1, mu2, s2, ...]   (n_lambda = 2N) */

typedef struct {
    int32_t kind;
    /* ---- FIELD_DHT: s = scale * H(amplitude . xi) + offset, lambda = link(s) ---- */
    int32_t ndim;               /* 1..3 */
    int64_t dims[3];            /* C order.  Any axis length up to 4096: power-of-two axes (>= 8) run the fused
                                   fast passes, other lengths a Bluestein chirp-z pass on the same FFT core (axes of
                                   <= 32 points a direct pass); power-of-two axes of 2-D / 3-D fields may be as long
                                   as 8192, a 1-D power-of-two line as long as 2^24 (four-step split).  Longer axes:
                                   MGVIB200_ERR_UNSUPPORTED. */
    const double* amplitude;    /* N */
    double scale;
    double offset;
    int32_t link;
    /* ---- likelihood (all kinds) ---- */
    int32_t likelihood;
    double sigma;               /* Normal: constant noise level (FIELD_DHT, DENSE_LINEAR) */
    int32_t lambda_layout;
    const double* data;         /* N observations (counts as float64 for Poisson) */
    int64_t n_data;
    /* ---- POLY: mean_i = sum_k coef_scale[k] p[k] x_i^k ; sigma = noise_scale * p[n_coef]^2 ---- */
    int32_t n_groups;
    const int64_t* group_sizes; /* n_groups */
    const double* x;            /* n_data abscissae, groups concatenated */
    int32_t n_coef;
    const double* coef_scale;   /* n_coef */
    double noise_scale;
    /* ---- DENSE_LINEAR: mu = J xi, J is (m x k) column-major ---- */
    const double* J;
    int64_t m, k;
    const double* sigma_vec;    /* optional per-datum noise level (m values); null: the constant `sigma`.
                                   F = diag(1/sigma_i^2) is folded into the operand load of the J'FJ GEMM */
    /* ---- FIELD_HYPER (docs/src/advanced_tutorial_lit.jl:194-307): a 1-D field (ndim = 1, dims[0] = N) whose
     *      amplitude spectrum depends on two hyperparameters, exponentiated, integrated over data bins:
     *        theta = [p1, p2, xi(N)]                          (PARIDX :153, hyperparameters first)
     *        A[k]  = ampl sqrt(2 pi cl) exp(-pi^2 kdist[k]^2 cl^2),  ampl = hyper[0] exp(hyper[1] p1),
     *                cl = hyper[2] exp(hyper[3] p2)           (amplitude_spectrum / sqrt_kernel :194-203)
     *        s = scale * H(A . xi),  fine = exp(s)            (gp_sample :253-256, poisson_gp_link :275-277)
     *        Lambda_b = binsize * sum_{t < grain} fine[bin_start + grain b + t],  b < n_data   (agg_lambdas :286-294)
     *        data_b ~ Poisson(Lambda_b)                       (model :301-307)
     *      n_theta = N + 2, n_lambda = n_data. ---- */
    const double* kdist;        /* N harmonic distances (dist_array :157-172) */
    double hyper[4];            /* a0, a1, l0, l1 */
    int64_t bin_start;          /* 0-based index of the first fine pixel of data bin 0 */
    int32_t grain;              /* fine pixels per data bin (GP_GRAIN_FACTOR) */
    double binsize;             /* width of a fine pixel (_GP_BINSIZE) */
} mgvib200_model_desc;

/* NewtonCG (src/newtoncg.jl:14-31) */
typedef struct {
    double alpha;      /* 0.1 */
    int64_t steps;     /* 4   */
    int64_t i0;        /* 5   */
    int64_t i1;        /* 50  */
} mgvib200_newtoncg_opts;

#define MGVIB200_TRACE_MAX 64
/* NewtonCGResults (src/newtoncg.jl:34-46) plus the discrete branch trace */
typedef struct {
    int64_t steps;
    double f_history[MGVIB200_TRACE_MAX + 1];     /* steps + 1 entries */
    int64_t n_cg_iterations;
    int64_t cg_iterations[MGVIB200_TRACE_MAX + 1]; /* as pushed by the reference (:108, :135) */
    int64_t cg_updates[MGVIB200_TRACE_MAX];        /* CG updates actually performed per Newton step */
    double step_lengths[MGVIB200_TRACE_MAX];       /* accepted line-search step per Newton step */
    int64_t f_calls, g_calls;
    double initial_f, minimum;
} mgvib200_newtoncg_trace;

/* sampler CG controls: LinearSolve.KrylovJL_CG() defaults are abstol = reltol = sqrt(eps(Float64)),
 * maxiters = n_theta (call site src/residual_samplers.jl:79-80).  maxiters <= 0 selects n_theta. */
typedef struct {
    double abstol;
    double reltol;
    int64_t maxiters;
} mgvib200_cg_opts;

/* ---- context (replaces MGVIContext's compute-unit role, src/mgvi_context.jl:9-15) ---- */
int mgvib200_init(int device_id, mgvib200_ctx** out);
void mgvib200_destroy(mgvib200_ctx* ctx);
const char* mgvib200_last_error(mgvib200_ctx* ctx);
const char* mgvib200_version(void);

/* ---- multi-GPU: one process per GPU; residual samples are sharded by the caller, the library
 *      all-reduces KL value / gradient input / mean metric weight over NCCL ---- */
int mgvib200_comm_unique_id(uint8_t out_id[128]);
int mgvib200_comm_init(mgvib200_ctx* ctx, int rank, int world_size, const uint8_t id[128]);
/* Host-supplied collective instead of NCCL (an MPI communicator held by the Julia host, gloo in the
 * CPU tests).  The only collective the library needs is an all-gather of doubles:
 * recv[r*count .. (r+1)*count) = rank r's send[0 .. count); pointers are memory of the context's
 * device, send may alias recv + rank*count; return 0 on success.  The library synchronises its
 * stream before the call and expects the result to be complete when the callback returns.
 * Sums over residual samples are formed in a fixed order from the gathered parts, so results do
 * not depend on the number of ranks (1, 2, 4 or 8). */
typedef int (*mgvib200_allgather_fn)(void* user, const double* send, double* recv, int64_t count);
int mgvib200_comm_init_external(mgvib200_ctx* ctx, int rank, int world_size, mgvib200_allgather_fn allgather,
                                void* user);

/* ---- model (the `forward_model` argument of mgvi_step, src/mgvi_impl.jl:129-132) ---- */
int mgvib200_model_create(mgvib200_ctx* ctx, const mgvib200_model_desc* desc, mgvib200_model** out);
void mgvib200_model_destroy(mgvib200_model* m);
int64_t mgvib200_model_n_theta(const mgvib200_model* m);
int64_t mgvib200_model_n_lambda(const mgvib200_model* m);

/* a4  _fisher_information_and_jac(model, center, OP, ctx)  src/residual_samplers.jl:4-8
 *     (ResidualSampler constructor :59-63): linearise at `center`; the metric weights stay on the
 *     device inside the model handle. */
int mgvib200_linearize(mgvib200_model* m, const double* center /* n_theta */);
/* diagnostics: metric weight w = F g'^2 and q = g' sqrt(F) on the mean rows (field / dense kinds) */
int mgvib200_get_weights(mgvib200_model* m, double* w /* N */, double* q /* N */);

/* a5  (J' F J + I) * V   src/residual_samplers.jl:72, src/mgvi_impl.jl:147-150 */
int mgvib200_metric_apply(mgvib200_model* m, const double* V /* n_theta x k */, int64_t k, double* MV);

/* a6/a7  sample_residuals(s::ResidualSampler, n) with the CG solver  src/residual_samplers.jl:66-92 */
int mgvib200_sample_residuals_cg(mgvib200_model* m, const double* Z1 /* n_lambda x n */,
                                 const double* Z2 /* n_theta x n */, int64_t n, const mgvib200_cg_opts* opts,
                                 double* R /* n_theta x n */, int64_t* iters /* n or NULL */,
                                 double* final_rnorm /* n or NULL */);

/* a8  residual_pushfwd_operator + sample_residuals(::MatrixInversion)  src/residual_samplers.jl:95-123 */
int mgvib200_sample_residuals_dense(mgvib200_model* m, const double* Z /* n_theta x n */, int64_t n,
                                    double* R /* n_theta x n */, double* L_sigma_or_null /* n_theta x n_theta */);

/* a9/a10  _mean_neg_log_pstr (src/mgvi_impl.jl:19-28) and its gradient (src/newtoncg.jl:100) */
int mgvib200_kl_value_grad(mgvib200_model* m, const double* x /* n_theta */, const double* R /* n_theta x n */,
                           int64_t n, double* f, double* grad_or_null /* n_theta */);

/* a11  mean metric  u -> mean_i (J_i' F_i J_i + I) u at the points x + r_i  src/mgvi_impl.jl:137-138 */
int mgvib200_mean_metric_apply(mgvib200_model* m, const double* x, const double* R, int64_t n, const double* u,
                               double* out);

/* a12  MGVI._optimize(f, adsel, Sigma_bar_inv, x0, ::NewtonCG, opts)  src/newtoncg.jl:85-167 */
int mgvib200_newtoncg(mgvib200_model* m, const double* x0, const double* R, int64_t n,
                      const mgvib200_newtoncg_opts* opts, double* x, double* f, mgvib200_newtoncg_trace* trace);

/* (f3)  The seam MGVI._optimize dispatches on for every other optimizer (ext/MGVIOptimExt.jl:12-23:
 *      OnceDifferentiable(f, x0) handed to Optim; ext/MGVIOptimizationBaseExt.jl:19-36): the KL objective
 *      f(x) = _mean_neg_log_pstr(model, data, R, x) and its gradient with the residual samples R held on
 *      the device between evaluations.  bind copies R (n_theta x n) once; eval moves only x and grad. */
int mgvib200_objective_bind(mgvib200_model* m, const double* R /* n_theta x n */, int64_t n);
int mgvib200_objective_eval(mgvib200_model* m, const double* x /* n_theta */, double* f, double* grad_or_null);

/* Parity probe of ONE Newton step (the loop body at src/newtoncg.jl:112-146): the caller injects the
 * state the reference holds at the top of step `step` -- x_n, f_prev = f(x_n) as carried, df_n -- and
 * gets back every intermediate of that step: grad f(x_n) (:115), the mean metric weight behind
 * Sigma_bar_inv(x_n) (:120, mgvi_impl.jl:137-138), every inner-CG iterate dx_k with its residual norm
 * and (steps >= 2) f(x_n - dx_k) (:121-139), the line-search evaluations in call order (:142), the
 * accepted step beta and x_{n+1} (:143).  force_k > 0 performs exactly that many CG updates
 * (teacher forcing); k_own_exit reports where the step's own exit rule (:124, :134) fired first.
 * Test instrumentation for rows a12/a14; the same code path as mgvib200_newtoncg. */
#define MGVIB200_PROBE_LS_MAX 96
typedef struct {
    /* inputs */
    int64_t step;            /* 1-based Newton step number */
    double f_prev, df_n;
    int64_t force_k;         /* 0 = the step's own exit rule */
    int64_t max_rec;         /* capacity of dx_iters (columns), f_k, cg_resid */
    /* vector outputs, host memory, each may be NULL */
    double* grad;            /* n_theta */
    double* wbar;            /* FIELD_DHT: N */
    double* dx_iters;        /* n_theta x max_rec */
    double* f_k;             /* max_rec (NaN at step 1) */
    double* cg_resid;        /* max_rec */
    double* x_new;           /* n_theta */
    /* scalar outputs */
    int64_t k_done, k_own_exit;
    double dphi0, beta, f_new;
    int64_t n_ls;
    int32_t ls_kind[MGVIB200_PROBE_LS_MAX];   /* 0 = phi(a), 1 = phi'(a) */
    double ls_a[MGVIB200_PROBE_LS_MAX], ls_val[MGVIB200_PROBE_LS_MAX];
} mgvib200_probe_rec;
int mgvib200_newton_probe(mgvib200_model* m, const double* x_n, const double* R, int64_t n,
                          const mgvib200_newtoncg_opts* opts, mgvib200_probe_rec* probe);

/* Parity probe of single updates of the Newton inner CG (IterativeSolvers cg_iterator!, call site
 * src/newtoncg.jl:120-139) on Sigma_bar_inv(x_n): for each of n_states injected iterator states
 * (x, r, u, residual, prev_residual -- as they stand BEFORE an update; columns of X, Rr, U) performs
 * exactly that update through the device CG kernels and returns the new iterate and residual norm.
 * The k-th iterate of a truncated CG amplifies rounding like cond^(k/2), so iterates are compared one
 * update at a time from the reference state.  FIELD_DHT models. */
int mgvib200_ncg_probe(mgvib200_model* m, const double* x_n, const double* R, int64_t n, int64_t n_states,
                       const double* X, const double* Rr, const double* U, const double* res, const double* prev,
                       double* X_out /* n_theta x n_states */, double* res_out /* n_states */);

/* a13  _build_samples  src/mgvi_impl.jl:152-156 : columns [c + r1, c - r1, c + r2, c - r2, ...] */
int mgvib200_build_samples(mgvib200_model* m, const double* center, const double* R, int64_t n,
                           double* samples /* n_theta x 2n */);

/* a14  mgvi_step  src/mgvi_impl.jl:129-144.  dense != 0 selects the MatrixInversion sampler
 *      (Z1 = the (n_theta x n) draw matrix, Z2 ignored). */
int mgvib200_step(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n,
                  int dense, const mgvib200_cg_opts* cg, const mgvib200_newtoncg_opts* opts,
                  double* samples /* n_theta x 2n */, double* center_out /* n_theta */, double* mnlp,
                  int64_t* sampler_iters /* n or NULL */, mgvib200_newtoncg_trace* trace /* or NULL */);

/*      mgvi_sample  src/mgvi_impl.jl:166-174 */
int mgvib200_sample(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n,
                    int dense, const mgvib200_cg_opts* cg, double* samples /* n_theta x 2n */,
                    int64_t* sampler_iters /* n or NULL */);

/* ---- device-resident variants (benchmark driver: inputs already in HBM) ---- */
void* mgvib200_dev_alloc(mgvib200_ctx* ctx, int64_t bytes);
void mgvib200_dev_free(mgvib200_ctx* ctx, void* p);
int mgvib200_memcpy_h2d(mgvib200_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
int mgvib200_memcpy_d2h(mgvib200_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);
int mgvib200_linearize_dev(mgvib200_model* m, const double* center_dev);
int mgvib200_metric_apply_dev(mgvib200_model* m, const double* V_dev, int64_t k, double* MV_dev);
int mgvib200_sample_residuals_cg_dev(mgvib200_model* m, const double* Z1_dev, const double* Z2_dev, int64_t n,
                                     const mgvib200_cg_opts* opts, double* R_dev, int64_t* iters_host,
                                     double* final_rnorm_host);
int mgvib200_step_dev(mgvib200_model* m, const double* center_dev, const double* Z1_dev, const double* Z2_dev,
                      int64_t n, int dense, const mgvib200_cg_opts* cg, const mgvib200_newtoncg_opts* opts,
                      double* samples_dev, double* center_out_dev, double* mnlp, int64_t* sampler_iters_host,
                      mgvib200_newtoncg_trace* trace);

/* ---- instrumentation ---- */
/* number of kernels this library launched on the context since the last reset */
int64_t mgvib200_launch_count(mgvib200_ctx* ctx, int reset);
/* device time (ms) of the most recent sampler CG loop / Newton phase, measured with CUDA events on
 * the library's own stream, and the lock-step iteration count of the most recent sampler call */
int mgvib200_last_timing(mgvib200_ctx* ctx, double* sampler_ms, double* newton_ms, int64_t* lockstep_iters);
/* per-kernel device timing: CUDA events around every launch between begin and end.  `end`
 * fills up to max_tags entries (tag, total ms, launch count) and returns how many; -1 on error.
 * tags: field pass = 100*prologue + 10*epilogue + fused_double ; vector kernels = 1000 + op. */
int mgvib200_profile_begin(mgvib200_ctx* ctx);
int mgvib200_profile_end(mgvib200_ctx* ctx, int max_tags, int32_t* tags, double* ms, int64_t* counts);
/* the CUDA stream the library launches on (cudaStream_t as void*) */
void* mgvib200_stream(mgvib200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MGVI_B200_H */
