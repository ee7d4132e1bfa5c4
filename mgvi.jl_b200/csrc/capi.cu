// extern "C" surface of libmgvib200 (include/mgvi_b200.h) and the model-independent drivers:
// Newton-CG with strong-Wolfe line search (src/newtoncg.jl:85-167) and mgvi_step
// (src/mgvi_impl.jl:129-144).  Scalar control flow runs on the host, every vector operation is a
// kernel launch on the context's stream.
#include <math.h>
#include <functional>
#include "engine.h"

namespace mgvi {
void comm_unique_id(uint8_t out[128]);
void comm_init(mgvib200_ctx* ctx, int rank, int world, const uint8_t id[128]);
void comm_init_external(mgvib200_ctx* ctx, int rank, int world, mgvib200_allgather_fn fn, void* user);
void comm_destroy(mgvib200_ctx* ctx);
}  // namespace mgvi

using namespace mgvi;

namespace {

thread_local std::string g_last_global_error;

template <typename F>
int guarded(mgvib200_ctx* ctx, F&& f) {
    try {
        f();
        return MGVIB200_OK;
    } catch (const RtError& e) {
        if (ctx) ctx->err = e.msg; else g_last_global_error = e.msg;
        return e.code ? e.code : MGVIB200_ERR_INVALID;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what(); else g_last_global_error = e.what();
        return MGVIB200_ERR_INVALID;
    }
}

void need(bool cond, const char* msg) { if (!cond) throw RtError{msg, MGVIB200_ERR_INVALID}; }

double* stage_in(mgvib200_model* m, int slot, const double* host, size_t count) {
    m->stage[slot].ensure(sizeof(double) * count);
    rt_h2d(m->stage[slot].p, host, sizeof(double) * count, m->ctx->stream);
    return m->stage[slot].as<double>();
}
double* stage_out(mgvib200_model* m, int slot, size_t count) {
    m->stage[slot].ensure(sizeof(double) * count);
    return m->stage[slot].as<double>();
}

mgvib200_cg_opts cg_defaults(const mgvib200_cg_opts* o) {
    mgvib200_cg_opts d;
    d.abstol = 1.4901161193847656e-08;   // LinearSolve default_tol(Float64) = sqrt(eps)
    d.reltol = 1.4901161193847656e-08;
    d.maxiters = 0;
    if (o) d = *o;
    return d;
}

mgvib200_newtoncg_opts ncg_defaults(const mgvib200_newtoncg_opts* o) {
    mgvib200_newtoncg_opts d;
    d.alpha = 0.1; d.steps = 4; d.i0 = 5; d.i1 = 50;     // src/newtoncg.jl:14-31
    if (o) d = *o;
    return d;
}

// ----------------------------------------------------------------------------------------------
// LineSearches.StrongWolfe{Float64}() -- src/newtoncg.jl:30, call :142
// ----------------------------------------------------------------------------------------------
typedef std::function<double(double)> Fn1;

double sw_interpolate(double a1, double a2, double f1, double f2, double g1, double g2) {
    const double d1 = g1 + g2 - 3.0 * (f1 - f2) / (a1 - a2);
    const double d2 = sqrt(d1 * d1 - g1 * g2);
    return a2 - (a2 - a1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2));
}

double sw_zoom(double a_lo, double a_hi, double dphi0, double phi0, const Fn1& phi, const Fn1& dphi, double c1,
               double c2) {
    double a_j = NAN;
    for (int it = 0; it < 10; ++it) {
        const double f_lo = phi(a_lo), g_lo = dphi(a_lo);
        const double f_hi = phi(a_hi), g_hi = dphi(a_hi);
        if (a_lo < a_hi) a_j = sw_interpolate(a_lo, a_hi, f_lo, f_hi, g_lo, g_hi);
        else a_j = sw_interpolate(a_hi, a_lo, f_hi, f_lo, g_hi, g_lo);
        const double f_j = phi(a_j);
        if (f_j > phi0 + c1 * a_j * dphi0 || f_j > f_lo) {
            a_hi = a_j;
        } else {
            const double g_j = dphi(a_j);
            if (fabs(g_j) <= -c2 * dphi0) return a_j;
            if (g_j * (a_hi - a_lo) >= 0.0) a_hi = a_lo;
            a_lo = a_j;
        }
    }
    return a_j;
}

void strong_wolfe(const Fn1& phi, const Fn1& dphi, double alpha0, double phi0, double dphi0, double& a_out,
                  double& f_out) {
    const double c1 = 1e-4, c2 = 0.9, rho = 2.0, a_max = 65536.0;
    double a_prev = 0.0, a = alpha0, f_prev = phi0;
    int i = 1;
    while (a < a_max) {
        const double f_a = phi(a);
        if (f_a > phi0 + c1 * a * dphi0 || (f_a >= f_prev && i > 1)) {
            a_out = sw_zoom(a_prev, a, dphi0, phi0, phi, dphi, c1, c2);
            f_out = phi(a_out);
            return;
        }
        const double g_a = dphi(a);
        if (fabs(g_a) <= -c2 * dphi0) { a_out = a; f_out = f_a; return; }
        if (g_a >= 0.0) {
            a_out = sw_zoom(a, a_prev, dphi0, phi0, phi, dphi, c1, c2);
            f_out = phi(a_out);
            return;
        }
        a_prev = a; a *= rho; f_prev = f_a; ++i;
    }
    a_out = a_max;
    f_out = phi(a_max);
}

// ----------------------------------------------------------------------------------------------
// MGVI._optimize(f, adsel, Sigma_bar_inv, x0, ::NewtonCG, opts) -- src/newtoncg.jl:85-167
// All pointers device.  x_out may alias nothing else.
// ----------------------------------------------------------------------------------------------
struct NewtonScratch {
    double *g, *trial, *gt;
};

NewtonScratch newton_scratch(mgvib200_ctx* ctx, long long N) {
    const size_t b = sizeof(double) * N;
    ctx->nw_g.ensure(b); ctx->nw_trial.ensure(b); ctx->nw_gt.ensure(b);
    return NewtonScratch{ctx->nw_g.as<double>(), ctx->nw_trial.as<double>(), ctx->nw_gt.as<double>()};
}

void d2h_now(mgvib200_ctx* ctx, double* host, const double* dev, long long count) {
    rt_d2h(host, dev, sizeof(double) * count, ctx->stream);
    rt_sync(ctx->stream);
}

// One Newton step (the body of the loop at src/newtoncg.jl:112-146): x is updated in place, f_prev / df_n
// are the carried scalars.  `rec` (parity probe, mgvib200_newton_probe) records every intermediate and may
// force the number of inner CG updates; the product path passes null.
void newton_one_step(mgvib200_model* m, double* x, const double* R, int64_t n, const mgvib200_newtoncg_opts& o,
                     int64_t step, double& f_prev, double& df_n, double& f_n, mgvib200_newtoncg_trace& tr,
                     mgvib200_probe_rec* rec) {
    Engine* E = m->eng;
    mgvib200_ctx* ctx = m->ctx;
    const long long N = E->n_theta;
    const size_t b = sizeof(double) * N;
    NewtonScratch S = newton_scratch(ctx, N);
    double* g = S.g; double* trial = S.trial; double* gt = S.gt;
    ReduceWs& drws = ctx->drv_rws;
    auto f = [&](const double* xp) { tr.f_calls++; return E->kl_value_grad(xp, R, n, nullptr); };
    auto gradf = [&](const double* xp, double* gout) { tr.g_calls++; E->kl_value_grad(xp, R, n, gout); };
    const int64_t force_k = rec ? rec->force_k : 0;
    auto rec_iter = [&](int64_t k, double f_k) {
        if (!rec || k > rec->max_rec) return;
        if (rec->dx_iters) d2h_now(ctx, rec->dx_iters + (size_t)(k - 1) * N, E->ncg_dx(), N);
        if (rec->f_k) rec->f_k[k - 1] = f_k;
        if (rec->cg_resid) rec->cg_resid[k - 1] = E->ncg_residual();
    };

    gradf(x, g);                                           // :115
    if (rec && rec->grad) d2h_now(ctx, rec->grad, g, N);
    E->mean_metric_prepare(x, R, n);                       // Sigma_bar_inv(x_n), :120 / mgvi_impl.jl:137-138
    if (rec && rec->wbar) E->get_mean_weights(rec->wbar);
    E->ncg_init(g);
    int64_t k = 0, own_exit = 0;
    if (step == 1) {                                       // :121-127
        while (E->ncg_active()) {
            E->ncg_update(); ++k;
            rec_iter(k, NAN);
            if (k >= o.i0 && !own_exit) own_exit = k;
            if (force_k > 0 ? k >= force_k : own_exit != 0) break;
        }
    } else {                                               // :129-139
        double f_km1 = f_prev;
        while (E->ncg_active()) {
            E->ncg_update(); ++k;
            vec_axpy(ctx, drws, trial, x, -1.0, E->ncg_dx(), N);
            const double f_k = f(trial);
            rec_iter(k, f_k);
            if ((k >= o.i1 || fabs(f_k - f_km1) < o.alpha * df_n) && !own_exit) own_exit = k;
            if (force_k > 0 ? k >= force_k : own_exit != 0) {
                tr.cg_iterations[tr.n_cg_iterations++] = k;
                break;
            }
            f_km1 = f_k;
        }
    }
    tr.cg_updates[step - 1] = k;
    const double* dx = E->ncg_dx();
    if (k == 0) {
        // the iterator yielded nothing: dx = 0 (cg_iterator! zeroes it)
        ctx->nw_zero.ensure(b);
        rt_memset(ctx->nw_zero.p, 0, b, ctx->stream);
        dx = ctx->nw_zero.as<double>();
    }
    // line search along d = -dx (linesearch_args, :64-71): phi(a) = f(x + a d), phi'(a) = grad f(x + a d) . d
    const double dphi0 = -vec_dot(ctx, drws, g, dx, N);
    auto rec_ls = [&](int kind, double a, double v) {
        if (!rec || rec->n_ls >= MGVIB200_PROBE_LS_MAX) return;
        rec->ls_kind[rec->n_ls] = kind; rec->ls_a[rec->n_ls] = a; rec->ls_val[rec->n_ls] = v;
        rec->n_ls++;
    };
    Fn1 phi = [&](double a) {
        vec_axpy(ctx, drws, trial, x, -a, dx, N);
        const double v = f(trial);
        rec_ls(0, a, v);
        return v;
    };
    Fn1 dphi = [&](double a) {
        vec_axpy(ctx, drws, trial, x, -a, dx, N);
        gradf(trial, gt);
        const double v = -vec_dot(ctx, drws, gt, dx, N);
        rec_ls(1, a, v);
        return v;
    };
    double beta = 0.0;
    strong_wolfe(phi, dphi, 1.0, f_prev, dphi0, beta, f_n);          // :142
    tr.step_lengths[step - 1] = beta;
    vec_axpy_inplace(ctx, drws, x, -beta, dx, N);                  // :143
    df_n = fabs(f_n - f_prev);                                       // :144
    f_prev = f_n;
    tr.f_history[step] = f_n;                                        // :146
    if (rec) {
        rec->k_done = k; rec->k_own_exit = own_exit; rec->dphi0 = dphi0; rec->beta = beta; rec->f_new = f_n;
        if (rec->x_new) d2h_now(ctx, rec->x_new, x, N);
    }
}

void newtoncg_dev(mgvib200_model* m, const double* x0, const double* R, int64_t n,
                  const mgvib200_newtoncg_opts& o, double* x_out, double* f_out, mgvib200_newtoncg_trace* tr_out) {
    Engine* E = m->eng;
    mgvib200_ctx* ctx = m->ctx;
    const long long N = E->n_theta;
    need(o.steps >= 0 && o.steps <= MGVIB200_TRACE_MAX, "NewtonCG: steps must be in [0, 64]");
    const size_t b = sizeof(double) * N;
    ctx->nw_x.ensure(b);
    double* x = ctx->nw_x.as<double>();
    mgvib200_newtoncg_trace tr;
    memset(&tr, 0, sizeof tr);
    rt_event_record(ctx->ev0, ctx->stream);

    rt_d2d(x, x0, b, ctx->stream);
    tr.f_calls++;
    double f_prev = E->kl_value_grad(x, R, n, nullptr);   // :106
    tr.initial_f = f_prev;
    tr.f_history[0] = f_prev;
    tr.cg_iterations[tr.n_cg_iterations++] = o.i0;         // :108
    double f_n = 0.0, df_n = 0.0;
    for (int64_t step = 1; step <= o.steps; ++step)        // :112
        newton_one_step(m, x, R, n, o, step, f_prev, df_n, f_n, tr, nullptr);
    if (o.steps == 0) f_n = f_prev;
    tr.steps = o.steps;
    tr.minimum = f_n;
    rt_d2d(x_out, x, b, ctx->stream);
    rt_event_record(ctx->ev1, ctx->stream);
    rt_sync(ctx->stream);
    ctx->newton_ms = rt_event_ms(ctx->ev0, ctx->ev1);
    if (f_out) *f_out = f_n;
    if (tr_out) *tr_out = tr;
}

void step_dev(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n, int dense,
              const mgvib200_cg_opts& cg, const mgvib200_newtoncg_opts& o, double* samples, double* center_out,
              double* mnlp, int64_t* iters_host, mgvib200_newtoncg_trace* tr, bool optimize) {
    Engine* E = m->eng;
    const long long N = E->n_theta;
    E->linearize(center);                                            // ResidualSampler ctor, :133
    m->linearized = true;
    ReduceWs& rws = m->ctx->drv_rws;
    m->ctx->step_R.ensure(sizeof(double) * N * n);
    double* R = m->ctx->step_R.as<double>();
    if (dense) {
        E->sample_dense(Z1, n, R, nullptr);                          // :134 with MatrixInversion
        if (iters_host) for (int64_t i = 0; i < n; ++i) iters_host[i] = 0;
    } else {
        E->sample_cg(Z1, Z2, n, cg, R, iters_host, nullptr);         // :134
    }
    const double sampler_ms = m->ctx->sampler_ms;
    const double* c_final = center;
    if (optimize) {
        newtoncg_dev(m, center, R, n, o, center_out, mnlp, tr);      // :135-139
        c_final = center_out;
    }
    m->ctx->sampler_ms = sampler_ms;
    vec_build_samples(m->ctx, rws, c_final, R, N, (int)n, samples);  // :140
    rt_sync(m->ctx->stream);
}

}  // namespace

// ==============================================================================================
extern "C" {

const char* mgvib200_version(void) { return "mgvib200 0.1 (sm_100a)"; }

int mgvib200_init(int device_id, mgvib200_ctx** out) {
    if (!out) return MGVIB200_ERR_INVALID;
    *out = nullptr;
    mgvib200_ctx* ctx = new mgvib200_ctx();
    int rc = guarded(nullptr, [&] {
        ctx->device = device_id;
        rt_set_device(device_id);
        ctx->stream = rt_stream_create();
        ctx->ev0 = rt_event_create();
        ctx->ev1 = rt_event_create();
        ctx->ev_poll[0] = rt_event_create();
        ctx->ev_poll[1] = rt_event_create();
        ctx->h_pinned = (int*)rt_host_alloc(65536);
#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
        cudaDeviceProp prop;
        RT_CUDA(cudaGetDeviceProperties(&prop, device_id));
        ctx->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) throw RtError{"libmgvib200 is built for sm_100a (B200) only", MGVIB200_ERR_UNSUPPORTED};
#endif
    });
    if (rc != MGVIB200_OK) { delete ctx; return rc; }
    *out = ctx;
    return MGVIB200_OK;
}

void mgvib200_destroy(mgvib200_ctx* ctx) {
    if (!ctx) return;
    comm_destroy(ctx);
    try { rt_set_device(ctx->device); } catch (...) {}
    for (mgvi::DevBuf* b : {&ctx->nw_x, &ctx->nw_g, &ctx->nw_trial, &ctx->nw_gt, &ctx->nw_zero, &ctx->step_R,
                            &ctx->count_scalar})
        b->release();
    ctx->drv_rws.release();
    for (auto e : ctx->ev_pool) rt_event_destroy(e);
    rt_host_free(ctx->h_pinned);
    rt_event_destroy(ctx->ev0);
    rt_event_destroy(ctx->ev1);
    rt_event_destroy(ctx->ev_poll[0]);
    rt_event_destroy(ctx->ev_poll[1]);
    rt_stream_destroy(ctx->stream);
    delete ctx;
}

const char* mgvib200_last_error(mgvib200_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_last_global_error.c_str();
}

int mgvib200_comm_unique_id(uint8_t out_id[128]) {
    return guarded(nullptr, [&] { comm_unique_id(out_id); });
}

int mgvib200_comm_init(mgvib200_ctx* ctx, int rank, int world_size, const uint8_t id[128]) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    return guarded(ctx, [&] { rt_set_device(ctx->device); comm_init(ctx, rank, world_size, id); });
}

int mgvib200_comm_init_external(mgvib200_ctx* ctx, int rank, int world_size, mgvib200_allgather_fn allgather,
                                void* user) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    return guarded(ctx, [&] { rt_set_device(ctx->device); comm_init_external(ctx, rank, world_size, allgather, user); });
}

int mgvib200_model_create(mgvib200_ctx* ctx, const mgvib200_model_desc* d, mgvib200_model** out) {
    if (!ctx || !d || !out) return MGVIB200_ERR_INVALID;
    *out = nullptr;
    return guarded(ctx, [&] {
        rt_set_device(ctx->device);
        Engine* e = nullptr;
        switch (d->kind) {
            case MGVIB200_MODEL_FIELD_DHT: e = make_field_engine(ctx, *d); break;
            case MGVIB200_MODEL_POLY: e = make_poly_engine(ctx, *d); break;
            case MGVIB200_MODEL_DENSE_LINEAR: e = make_dense_engine(ctx, *d); break;
            case MGVIB200_MODEL_FIELD_HYPER: e = make_hyper_engine(ctx, *d); break;
            default: throw RtError{"unknown model kind", MGVIB200_ERR_INVALID};
        }
        mgvib200_model* m = new mgvib200_model();
        m->ctx = ctx;
        m->eng = e;
        *out = m;
    });
}

void mgvib200_model_destroy(mgvib200_model* m) {
    if (!m) return;
    delete m->eng;
    for (auto& s : m->stage) s.release();
    m->obj_R.release();
    delete m;
}

int64_t mgvib200_model_n_theta(const mgvib200_model* m) { return m ? m->eng->n_theta : -1; }
int64_t mgvib200_model_n_lambda(const mgvib200_model* m) { return m ? m->eng->n_lambda : -1; }

// ---- device-pointer entry points ---------------------------------------------------------------
int mgvib200_linearize_dev(mgvib200_model* m, const double* center_dev) {
    if (!m) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        rt_set_device(m->ctx->device);
        m->eng->linearize(center_dev);
        m->linearized = true;
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_metric_apply_dev(mgvib200_model* m, const double* V, int64_t k, double* MV) {
    if (!m) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "metric_apply: call mgvib200_linearize first");
        need(k >= 0, "metric_apply: k must be >= 0");
        if (k == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        m->eng->metric_apply(V, k, MV);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_sample_residuals_cg_dev(mgvib200_model* m, const double* Z1, const double* Z2, int64_t n,
                                     const mgvib200_cg_opts* opts, double* R, int64_t* iters_host,
                                     double* rnorm_host) {
    if (!m) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "sample_residuals_cg: call mgvib200_linearize first");
        need(n >= 0, "sample_residuals_cg: n must be >= 0");
        if (n == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        m->eng->sample_cg(Z1, Z2, n, cg_defaults(opts), R, iters_host, rnorm_host);
    });
}

int mgvib200_step_dev(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n,
                      int dense, const mgvib200_cg_opts* cg, const mgvib200_newtoncg_opts* opts, double* samples,
                      double* center_out, double* mnlp, int64_t* iters_host, mgvib200_newtoncg_trace* trace) {
    if (!m) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "step: n must be >= 1");
        rt_set_device(m->ctx->device);
        step_dev(m, center, Z1, Z2, n, dense, cg_defaults(cg), ncg_defaults(opts), samples, center_out, mnlp,
                 iters_host, trace, true);
    });
}

// ---- host-pointer entry points (copy in, run, copy out) ------------------------------------------
int mgvib200_linearize(mgvib200_model* m, const double* center) {
    if (!m || !center) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        rt_set_device(m->ctx->device);
        double* c = stage_in(m, 0, center, m->eng->n_theta);
        m->eng->linearize(c);
        m->linearized = true;
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_get_weights(mgvib200_model* m, double* w, double* q) {
    if (!m) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "get_weights: call mgvib200_linearize first");
        rt_set_device(m->ctx->device);
        m->eng->get_weights(w, q);
    });
}

int mgvib200_metric_apply(mgvib200_model* m, const double* V, int64_t k, double* MV) {
    if (!m || !V || !MV) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "metric_apply: call mgvib200_linearize first");
        need(k >= 0, "metric_apply: k must be >= 0");
        if (k == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        const size_t cnt = (size_t)m->eng->n_theta * k;
        double* v = stage_in(m, 1, V, cnt);
        double* o = stage_out(m, 2, cnt);
        m->eng->metric_apply(v, k, o);
        rt_d2h(MV, o, sizeof(double) * cnt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_sample_residuals_cg(mgvib200_model* m, const double* Z1, const double* Z2, int64_t n,
                                 const mgvib200_cg_opts* opts, double* R, int64_t* iters, double* final_rnorm) {
    if (!m || !Z1 || !Z2 || !R) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "sample_residuals_cg: call mgvib200_linearize first");
        need(n >= 0, "sample_residuals_cg: n must be >= 0");
        if (n == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        double* z1 = stage_in(m, 1, Z1, (size_t)m->eng->n_lambda * n);
        double* z2 = stage_in(m, 2, Z2, (size_t)m->eng->n_theta * n);
        double* r = stage_out(m, 3, (size_t)m->eng->n_theta * n);
        m->eng->sample_cg(z1, z2, n, cg_defaults(opts), r, iters, final_rnorm);
        rt_d2h(R, r, sizeof(double) * m->eng->n_theta * n, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_sample_residuals_dense(mgvib200_model* m, const double* Z, int64_t n, double* R, double* L) {
    if (!m || !Z || !R) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->linearized, "sample_residuals_dense: call mgvib200_linearize first");
        need(n >= 0, "sample_residuals_dense: n must be >= 0");
        if (n == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* z = stage_in(m, 1, Z, nt * n);
        double* r = stage_out(m, 3, nt * n);
        double* l = L ? stage_out(m, 4, nt * nt) : nullptr;
        m->eng->sample_dense(z, n, r, l);
        rt_d2h(R, r, sizeof(double) * nt * n, m->ctx->stream);
        if (L) rt_d2h(L, l, sizeof(double) * nt * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_kl_value_grad(mgvib200_model* m, const double* x, const double* R, int64_t n, double* f, double* grad) {
    if (!m || !x || !R || !f) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "kl_value_grad: n must be >= 1");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        double* gd = grad ? stage_out(m, 5, nt) : nullptr;
        *f = m->eng->kl_value_grad(xd, rd, n, gd);
        if (grad) rt_d2h(grad, gd, sizeof(double) * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_mean_metric_apply(mgvib200_model* m, const double* x, const double* R, int64_t n, const double* u,
                               double* out) {
    if (!m || !x || !R || !u || !out) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "mean_metric_apply: n must be >= 1");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        double* ud = stage_in(m, 5, u, nt);
        double* od = stage_out(m, 6, nt);
        m->eng->mean_metric_prepare(xd, rd, n);
        m->eng->mean_metric_apply(ud, od);
        rt_d2h(out, od, sizeof(double) * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_newtoncg(mgvib200_model* m, const double* x0, const double* R, int64_t n,
                      const mgvib200_newtoncg_opts* opts, double* x, double* f, mgvib200_newtoncg_trace* trace) {
    if (!m || !x0 || !R || !x) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "newtoncg: n must be >= 1");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x0, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        double* od = stage_out(m, 6, nt);
        newtoncg_dev(m, xd, rd, n, ncg_defaults(opts), od, f, trace);
        rt_d2h(x, od, sizeof(double) * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_objective_bind(mgvib200_model* m, const double* R, int64_t n) {
    if (!m || !R) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "objective_bind: n must be >= 1");
        rt_set_device(m->ctx->device);
        const size_t cnt = (size_t)m->eng->n_theta * n;
        m->obj_R.ensure(sizeof(double) * cnt);
        rt_h2d(m->obj_R.p, R, sizeof(double) * cnt, m->ctx->stream);
        rt_sync(m->ctx->stream);
        m->obj_n = n;
    });
}

int mgvib200_objective_eval(mgvib200_model* m, const double* x, double* f, double* grad) {
    if (!m || !x || !f) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(m->obj_n >= 1, "objective_eval: call mgvib200_objective_bind first");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x, nt);
        double* gd = grad ? stage_out(m, 5, nt) : nullptr;
        *f = m->eng->kl_value_grad(xd, m->obj_R.as<double>(), m->obj_n, gd);
        if (grad) rt_d2h(grad, gd, sizeof(double) * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_newton_probe(mgvib200_model* m, const double* x_n, const double* R, int64_t n,
                          const mgvib200_newtoncg_opts* opts, mgvib200_probe_rec* probe) {
    if (!m || !x_n || !R || !probe) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "newton_probe: n must be >= 1");
        need(probe->step >= 1 && probe->step <= MGVIB200_TRACE_MAX, "newton_probe: step must be in [1, 64]");
        need(probe->max_rec >= 0 && probe->force_k >= 0, "newton_probe: max_rec and force_k must be >= 0");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x_n, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        m->ctx->nw_x.ensure(sizeof(double) * nt);
        double* x = m->ctx->nw_x.as<double>();
        rt_d2d(x, xd, sizeof(double) * nt, m->ctx->stream);
        mgvib200_newtoncg_trace tr;
        memset(&tr, 0, sizeof tr);
        probe->n_ls = 0;
        double f_prev = probe->f_prev, df_n = probe->df_n, f_n = 0.0;
        newton_one_step(m, x, rd, n, ncg_defaults(opts), probe->step, f_prev, df_n, f_n, tr, probe);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_ncg_probe(mgvib200_model* m, const double* x_n, const double* R, int64_t n, int64_t n_states,
                       const double* X, const double* Rr, const double* U, const double* res, const double* prev,
                       double* X_out, double* res_out) {
    if (!m || !x_n || !R || !X || !Rr || !U || !res || !prev || !X_out || !res_out) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1 && n_states >= 0, "ncg_probe: n must be >= 1, n_states >= 0");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* xd = stage_in(m, 0, x_n, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        m->eng->mean_metric_prepare(xd, rd, n);                 // Sigma_bar_inv(x_n), src/newtoncg.jl:120
        for (int64_t j = 0; j < n_states; ++j) {
            double* a = stage_in(m, 1, X + j * nt, nt);
            double* b = stage_in(m, 2, Rr + j * nt, nt);
            double* c = stage_in(m, 5, U + j * nt, nt);
            m->eng->ncg_set_state(a, b, c, res[j], prev[j]);
            m->eng->ncg_update();
            rt_d2h(X_out + j * nt, m->eng->ncg_dx(), sizeof(double) * nt, m->ctx->stream);
            res_out[j] = m->eng->ncg_residual();
        }
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_build_samples(mgvib200_model* m, const double* center, const double* R, int64_t n, double* samples) {
    if (!m || !center || !R || !samples) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 0, "build_samples: n must be >= 0");
        if (n == 0) return;     // empty batch: nothing to do (the reference returns an n_theta x 0 matrix)
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta;
        double* cd = stage_in(m, 0, center, nt);
        double* rd = stage_in(m, 3, R, nt * n);
        double* sd = stage_out(m, 7, nt * 2 * n);
        vec_build_samples(m->ctx, m->ctx->drv_rws, cd, rd, (long long)nt, (int)n, sd);
        rt_d2h(samples, sd, sizeof(double) * nt * 2 * n, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

static int step_host(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n,
                     int dense, const mgvib200_cg_opts* cg, const mgvib200_newtoncg_opts* opts, double* samples,
                     double* center_out, double* mnlp, int64_t* iters, mgvib200_newtoncg_trace* trace,
                     bool optimize) {
    if (!m || !center || !Z1 || !samples) return MGVIB200_ERR_INVALID;
    if (!dense && !Z2) return MGVIB200_ERR_INVALID;
    return guarded(m->ctx, [&] {
        need(n >= 1, "step: n must be >= 1");
        rt_set_device(m->ctx->device);
        const size_t nt = (size_t)m->eng->n_theta, nl = (size_t)m->eng->n_lambda;
        double* cd = stage_in(m, 0, center, nt);
        double* z1 = stage_in(m, 1, Z1, dense ? nt * n : nl * n);
        double* z2 = dense ? nullptr : stage_in(m, 2, Z2, nt * n);
        double* sd = stage_out(m, 7, nt * 2 * n);
        double* od = stage_out(m, 6, nt);
        step_dev(m, cd, z1, z2, n, dense, cg_defaults(cg), ncg_defaults(opts), sd, od, mnlp, iters, trace,
                 optimize);
        rt_d2h(samples, sd, sizeof(double) * nt * 2 * n, m->ctx->stream);
        if (center_out && optimize) rt_d2h(center_out, od, sizeof(double) * nt, m->ctx->stream);
        rt_sync(m->ctx->stream);
    });
}

int mgvib200_step(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n, int dense,
                  const mgvib200_cg_opts* cg, const mgvib200_newtoncg_opts* opts, double* samples,
                  double* center_out, double* mnlp, int64_t* sampler_iters, mgvib200_newtoncg_trace* trace) {
    if (!center_out || !mnlp) return MGVIB200_ERR_INVALID;
    return step_host(m, center, Z1, Z2, n, dense, cg, opts, samples, center_out, mnlp, sampler_iters, trace, true);
}

int mgvib200_sample(mgvib200_model* m, const double* center, const double* Z1, const double* Z2, int64_t n,
                    int dense, const mgvib200_cg_opts* cg, double* samples, int64_t* sampler_iters) {
    return step_host(m, center, Z1, Z2, n, dense, cg, nullptr, samples, nullptr, nullptr, sampler_iters, nullptr,
                     false);
}

// ---- raw device memory for the benchmark driver --------------------------------------------------
void* mgvib200_dev_alloc(mgvib200_ctx* ctx, int64_t bytes) {
    if (!ctx || bytes < 0) return nullptr;
    void* p = nullptr;
    guarded(ctx, [&] { rt_set_device(ctx->device); p = rt_malloc((size_t)bytes); });
    return p;
}
void mgvib200_dev_free(mgvib200_ctx* ctx, void* p) {
    if (!ctx) return;
    guarded(ctx, [&] { rt_set_device(ctx->device); rt_free(p); });
}
int mgvib200_memcpy_h2d(mgvib200_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    return guarded(ctx, [&] { rt_set_device(ctx->device); rt_h2d(dst, src, (size_t)bytes, ctx->stream); rt_sync(ctx->stream); });
}
int mgvib200_memcpy_d2h(mgvib200_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    return guarded(ctx, [&] { rt_set_device(ctx->device); rt_d2h(dst, src, (size_t)bytes, ctx->stream); rt_sync(ctx->stream); });
}

int64_t mgvib200_launch_count(mgvib200_ctx* ctx, int reset) {
    if (!ctx) return -1;
    const int64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int mgvib200_last_timing(mgvib200_ctx* ctx, double* sampler_ms, double* newton_ms, int64_t* lockstep_iters) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    if (sampler_ms) *sampler_ms = ctx->sampler_ms;
    if (newton_ms) *newton_ms = ctx->newton_ms;
    if (lockstep_iters) *lockstep_iters = ctx->lockstep_iters;
    return MGVIB200_OK;
}

int mgvib200_profile_begin(mgvib200_ctx* ctx) {
    if (!ctx) return MGVIB200_ERR_INVALID;
    ctx->prof.clear();
    ctx->ev_used = 0;
    ctx->profile = true;
    return MGVIB200_OK;
}

int mgvib200_profile_end(mgvib200_ctx* ctx, int max_tags, int32_t* tags, double* ms, int64_t* counts) {
    if (!ctx) return -1;
    ctx->profile = false;
    int ntags = 0;
    int rc = guarded(ctx, [&] {
        rt_sync(ctx->stream);
        std::map<int, std::pair<double, int64_t>> acc;
        for (auto& r : ctx->prof) {
            auto& e = acc[r.tag];
            e.first += rt_event_ms(r.a, r.b);
            e.second += 1;
        }
        for (auto& kv : acc) {
            if (ntags >= max_tags) break;
            tags[ntags] = kv.first; ms[ntags] = kv.second.first; counts[ntags] = kv.second.second;
            ++ntags;
        }
        ctx->prof.clear();
        ctx->ev_used = 0;
    });
    return rc == MGVIB200_OK ? ntags : -1;
}

void* mgvib200_stream(mgvib200_ctx* ctx) { return ctx ? (void*)(intptr_t)ctx->stream : nullptr; }

}  // extern "C"
