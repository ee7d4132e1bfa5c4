// Host-side state of libmgvib200: context, model handles and the per-kind engines.
#pragma once
#include <math.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/mgvi_b200.h"
#include "rt.h"
#include "vec_ops.h"

namespace mgvi {

struct Comm;   // NCCL communicator wrapper (comm.cu); null = single GPU

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void ensure(size_t b) {
        if (b <= bytes) return;
        rt_free(p);
        p = nullptr; bytes = 0;
        p = rt_malloc(b);
        bytes = b;
    }
    void release() { rt_free(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// reduction scratch shared by all kernels of a model
struct ReduceWs {
    DevBuf partials;   // doubles
    DevBuf counters;   // unsigned, zero-initialised, self-resetting
    DevBuf red;        // doubles: small result slots
    size_t ncounters = 0;
    void release() { partials.release(); counters.release(); red.release(); ncounters = 0; }
};

}  // namespace mgvi

struct mgvib200_ctx {
    int device = 0;
    mgvi::rt_stream stream{};
    std::string err;
    int64_t launches = 0;
    mgvi::Comm* comm = nullptr;
    int rank = 0, world = 1;
    mgvi::rt_event ev0{}, ev1{}, ev_poll[2]{};
    double sampler_ms = 0.0, newton_ms = 0.0;
    int64_t lockstep_iters = 0;
    int* h_pinned = nullptr;        // small pinned staging area (64 KB)
    int sm_count = 148;
    // per-context device scratch of the model-independent drivers (Newton-CG, mgvi_step): owned by the
    // context so that several contexts (devices) can live in one process
    mgvi::DevBuf nw_x, nw_g, nw_trial, nw_gt, nw_zero, step_R, count_scalar;
    // multi-GPU bookkeeping (comm.cu): cached global residual count, models that cannot be sharded
    int64_t n_cache_local = -1, n_cache_glob = -1;
    int n_models_single_rank = 0;
    mgvi::ReduceWs drv_rws;
    // per-kernel timing (bench.py roofline): CUDA events around every launch while enabled
    bool profile = false;
    struct ProfRec { int tag; mgvi::rt_event a, b; };
    std::vector<ProfRec> prof;
    std::vector<mgvi::rt_event> ev_pool;
    size_t ev_used = 0;
    mgvi::rt_event prof_event() {
        if (ev_used == ev_pool.size()) ev_pool.push_back(mgvi::rt_event_create());
        return ev_pool[ev_used++];
    }
    void prof_pre(int tag) {
        if (!profile) return;
        ProfRec r; r.tag = tag; r.a = prof_event(); r.b = prof_event();
        mgvi::rt_event_record(r.a, stream);
        prof.push_back(r);
    }
    void prof_post() {
        if (!profile) return;
        mgvi::rt_event_record(prof.back().b, stream);
    }
};

namespace mgvi {

struct CgWs {
    int nvec = 0;
    DevBuf gamma, alpha, beta, eps, rnorm, pAp, iters, active, n_active, x_pending, x_count;
    ~CgWs() {
        DevBuf* all[] = {&gamma, &alpha, &beta, &eps, &rnorm, &pAp, &iters, &active, &n_active, &x_pending, &x_count};
        for (auto* b : all) b->release();
    }
    bool defer = false;     // hand out the deferred-solution-update flags (sampler CG of 2-D / 3-D fields only)
    CgScalars scalars(double atol, double rtol, long long itmax, int flavor) const {
        CgScalars c;
        c.gamma = gamma.as<double>(); c.alpha = alpha.as<double>(); c.beta = beta.as<double>();
        c.eps = eps.as<double>(); c.rnorm = rnorm.as<double>(); c.pAp = pAp.as<double>();
        c.iters = iters.as<int>(); c.active = active.as<int>(); c.n_active = n_active.as<int>();
        c.x_pending = defer ? x_pending.as<int>() : nullptr; c.x_count = defer ? x_count.as<unsigned>() : nullptr;
        c.atol = atol; c.rtol = rtol; c.itmax = itmax; c.flavor = flavor;
        return c;
    }
    void ensure(int n) {
        if (n <= nvec) return;
        size_t d = sizeof(double) * n, i = sizeof(int) * n;
        gamma.ensure(d); alpha.ensure(d); beta.ensure(d); eps.ensure(d); rnorm.ensure(d); pAp.ensure(d);
        iters.ensure(i); active.ensure(i); n_active.ensure(sizeof(int) * 4); x_pending.ensure(i); x_count.ensure(i);
        nvec = n;
    }
};

// Interface every model family implements; all pointers are DEVICE pointers.
struct Engine {
    mgvib200_ctx* ctx = nullptr;
    int64_t n_theta = 0, n_lambda = 0;
    virtual ~Engine() {}
    virtual void linearize(const double* center) = 0;
    virtual void get_weights(double* w_host, double* q_host) = 0;
    virtual void metric_apply(const double* V, int64_t k, double* MV) = 0;
    virtual void sample_cg(const double* Z1, const double* Z2, int64_t n, const mgvib200_cg_opts& o, double* R,
                           int64_t* iters_host, double* rnorm_host) = 0;
    virtual void sample_dense(const double* Z, int64_t n, double* R, double* Lsigma_or_null) = 0;
    // f = mean negative log posterior over x +- r_i ; grad may be null
    virtual double kl_value_grad(const double* x, const double* R, int64_t n, double* grad) = 0;
    // prepares the mean metric at x (weights over the '+' points) for subsequent apply calls
    virtual void mean_metric_prepare(const double* x, const double* R, int64_t n) = 0;
    virtual void mean_metric_apply(const double* u, double* out) = 0;
    // Newton inner CG on the prepared mean metric: init with rhs g, then single updates
    virtual void ncg_init(const double* g) = 0;
    virtual bool ncg_active() = 0;          // iterator would yield another update
    virtual void ncg_update() = 0;          // one CG update (dx += alpha u ...)
    virtual const double* ncg_dx() = 0;     // current dx (device)
    // parity probe (mgvib200_newton_probe): residual norm of the inner CG, mean metric weight (host copy)
    virtual double ncg_residual() { return NAN; }
    // inject the state of the IterativeSolvers iterator (x, r, u, residual, prev_residual) as it stands before
    // an update (mgvib200_ncg_probe); the next ncg_update() performs exactly that update
    virtual void ncg_set_state(const double*, const double*, const double*, double, double) {
        throw RtError{"the Newton-CG state probe is provided for FIELD_DHT models", 3};
    }
    virtual void get_mean_weights(double*) { throw RtError{"the mean metric weight is only exposed by FIELD_DHT models", 3}; }
};

Engine* make_field_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d);
Engine* make_poly_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d);
Engine* make_dense_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d);
Engine* make_hyper_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d);

// MatrixInversion sampler (src/residual_samplers.jl:95-123) on a materialised symmetric metric M (k x k,
// device memory; dense_engine.cu): blocked Cholesky of the index-reversed M on the FP64 tensor pipe,
// residuals R = L_Sigma Z and optionally L_Sigma itself.  Used by the field engines for small models.
struct DenseChol;
DenseChol* dense_chol_create(mgvib200_ctx* ctx);
void dense_chol_destroy(DenseChol* c);
void dense_chol_identity(DenseChol* c, double* eye, long long k);      // eye = I_k
void dense_chol_sample(DenseChol* c, const double* M, long long k, const double* Z, int64_t n, double* R, double* L);

// communicator (comm.cu)
void comm_allgather(mgvib200_ctx* ctx, const double* send_dev, double* recv_dev, long long count_per_rank);
int64_t comm_global_count(mgvib200_ctx* ctx, int64_t local);

// Canonical summation order over residual samples, independent of the number of ranks: the n_glob
// samples are cut into `chunks` = gcd(n_glob, 8) contiguous chunks of `per_chunk` samples; a sum
// over samples is  sum_c ( sum_{i in chunk c} term_i ), both levels sequential in index order.
// Rank g of W (W in {1, 2, 4, 8}) owns the contiguous chunks [g chunks/W, (g+1) chunks/W).
struct CanonShape {
    int64_t n_glob = 0, chunks = 1, per_chunk = 0, chunks_local = 1;
};
CanonShape canon_shape(mgvib200_ctx* ctx, int64_t n_local);

// generic vector helpers (vec_host.cu)
void vec_launch(mgvib200_ctx* ctx, ReduceWs& ws, VecParams P);
void reduce_ws_ensure(ReduceWs& ws, mgvib200_ctx* ctx, size_t npartials, size_t ncounters);
double vec_dot(mgvib200_ctx* ctx, ReduceWs& ws, const double* a, const double* b, long long N);
void vec_axpy(mgvib200_ctx* ctx, ReduceWs& ws, double* out, const double* x, double a, const double* y, long long N);
void vec_axpy_inplace(mgvib200_ctx* ctx, ReduceWs& ws, double* x, double a, const double* y, long long N);
void vec_build_samples(mgvib200_ctx* ctx, ReduceWs& ws, const double* c, const double* R, long long N, int n,
                       double* samples);
void vec_sum_chunks(mgvib200_ctx* ctx, ReduceWs& ws, const double* in, long long ld, int nsum, double scale,
                    double* out, long long N);
// out = scale * (canonical sum over all ranks' terms) ; in = [terms_local][N] with `terms_per_sample`
// consecutive terms per residual sample (2 for the +/- points, 1 for the '+' points)
void vec_canon_sum(mgvib200_ctx* ctx, ReduceWs& ws, const CanonShape& cs, const double* in, long long ld,
                   int terms_per_sample, double scale, double* out, long long N, DevBuf& gather);
// canonical sum of per-term scalars: `groups` arrays of `count` doubles each, laid out back to back at
// dev (local values); returns the sum over all ranks, group after group, each in global term order
double canon_scalar_sum(mgvib200_ctx* ctx, DevBuf& gather, const double* dev_local, int groups, long long count);
void vec_cg_init(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, const double* b, long long b_ld, double* x,
                 double* r, double* p, long long N, int nvec);
void vec_cg_update(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* x, double* r, const double* p,
                   const double* Ap, long long N, int nvec);
// deferred solution update (field_pass.h: defer_x): the residual half of the update; the flush of what is still pending
void vec_cg_update_r(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* r, const double* Ap, long long N, int nvec);
void vec_x_flush(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* x, const double* p, long long N, int nvec);

}  // namespace mgvi

struct mgvib200_model {
    mgvib200_ctx* ctx = nullptr;
    mgvi::Engine* eng = nullptr;
    mgvi::DevBuf stage[8];      // host-pointer API staging buffers
    bool linearized = false;
    mgvi::DevBuf obj_R;         // residual samples bound by mgvib200_objective_bind
    int64_t obj_n = 0;
};
