// Hyperparameter-field engine -- the model of the reference's tutorial
// (docs/src/advanced_tutorial_lit.jl:194-307):
//
//   theta = [p1, p2, xi (N)]                                  (PARIDX :153, hyperparameters first)
//   ampl  = a0 exp(a1 p1),  cl = l0 exp(l1 p2)                (sqrt_kernel :198-203)
//   A[k]  = ampl sqrt(2 pi cl) exp(-pi^2 d_k^2 cl^2)          (amplitude_spectrum :194-196)
//   s     = c H(A . xi)                                       (gp_sample :253-256; H = FFTW r2r DHT)
//   fine  = exp(s)                                            (poisson_gp_link :275-277)
//   Lambda_b = binsize sum_{t < grain} fine[start + grain b + t]   (agg_lambdas :286-294)
//   data_b ~ Poisson(Lambda_b), Fisher 1 / Lambda_b           (model :301-307, src/information.jl:74-78)
//
// Unlike the fixed-amplitude field the Jacobian depends on the point through A(p1, p2), so the mean
// metric does NOT collapse to one product with an averaged weight: every Newton-CG iteration applies
// the n per-sample metrics (batched, one block per sample) and sums them in the canonical order.
// The tutorial's field has ~800 pixels (not a power of two): one thread block owns one vector, the
// Hartley transform is a direct O(N^2) sum against a cos+sin table that stays in L2, and only the
// pixels under data bins enter the second half of a product.  Jacobian, per point (with xi, A, fine,
// Lambda of that point):
//   J v    = agg( fine . c H( A . v_xi + (v1 dA/dp1 + v2 dA/dp2) . xi ) )
//   J' l   = [ sum dA/dp1 . xi . g ;  sum dA/dp2 . xi . g ;  A . g ],   g = c H( scatter(binsize l_b fine) )
//   dA/dp1 = a1 A,   dA/dp2 = A l1 (1/2 - 2 pi^2 d^2 cl^2)
#include <math.h>
#include <algorithm>
#include "engine.h"
#include "kernels.h"

namespace mgvi {

struct HyperPointState {
    double* theta;     // [npts][n_theta]
    double* A;         // [npts][N]
    double* fine;      // [npts][N]
    double* lam;       // [npts][nb]
};

struct HyperParams {
    long long N, nb, start;
    int grain, nvec, mode;
    double c, binsize, a0, a1, l0, l1;
    const double* kd;        // N
    const double* cas;       // N x N, cas[k N + j] = cos + sin (2 pi j k / N)   (symmetric)
    const double* data;      // nb
    HyperPointState st;
    int st_shared;           // 1: every vector uses point 0 of st; 0: vector s uses point s
    // HM_POINT: the point is x + sign * R[:, i]
    const double* x; const double* R; int point_mode;
    int want_grad;
    double* red;             // [2][nvec]: prior term, variable part of -loglike
    // HM_OP / HM_RHS operands
    const double* V; long long v_ld;      // input vectors (v_ld = 0: one shared vector)
    const double* Z1; const double* Z2;   // HM_RHS draws: nb x nvec, n_theta x nvec
    double* out; long long out_ld;
};

enum { HM_POINT = 0, HM_OP = 1, HM_RHS = 2 };

// fixed-order block sum (every thread gets the result)
SIMT_DEVICE double hyper_block_sum(double v, double* red, double* bc) {
    const double t = seg_reduce(v, SIMT_NTID, red);
    if (SIMT_TID == 0) *bc = t;
    SIMT_SYNC();
    const double r = *bc;
    SIMT_SYNC();
    return r;
}

SIMT_DEVICE double hyper_dA2_factor(const HyperParams& P, double kd, double cl) {
    return P.l1 * (0.5 - 2.0 * 9.869604401089358 * kd * kd * cl * cl);     // pi^2
}

// J' l for the point (theta, A, fine): lb[b] in shared memory; bufA / bufB are N doubles of scratch.
// Adds nothing: writes out[0], out[1], out[2 + k] = add[.] + result (add may be null).
SIMT_DEVICE void hyper_adjoint(const HyperParams& P, const double* theta, const double* A, const double* fine,
                               const double* lb, double* bufA, double* bufB, const double* add, double* out,
                               double* red, double* bc) {
    const int tid = SIMT_TID, nt = SIMT_NTID;
    const long long N = P.N, j0 = P.start, j1 = P.start + P.nb * P.grain;
    const double cl = P.l0 * exp(P.l1 * theta[1]);
    // z[j] = binsize l_b fine[j] under the data bins (zero elsewhere: those pixels are skipped in the sum)
    for (long long j = j0 + tid; j < j1; j += nt) bufA[j] = (P.binsize * lb[(j - j0) / P.grain]) * fine[j];
    SIMT_SYNC();
    double h1 = 0.0, h2 = 0.0;
    for (long long k = tid; k < N; k += nt) {
        double g = 0.0;
        for (long long j = j0; j < j1; ++j) g += SIMT_LDG(P.cas + j * N + k) * bufA[j];     // cas is symmetric: coalesced over k
        g *= P.c;
        const double a = A[k], xi = theta[2 + k];
        out[2 + k] = (add ? add[2 + k] : 0.0) + a * g;
        const double t = a * xi * g;
        h1 += P.a1 * t;
        h2 += hyper_dA2_factor(P, P.kd[k], cl) * t;
    }
    (void)bufB;
    const double s1 = hyper_block_sum(h1, red, bc);
    const double s2 = hyper_block_sum(h2, red, bc);
    if (tid == 0) {
        out[0] = (add ? add[0] : 0.0) + s1;
        out[1] = (add ? add[1] : 0.0) + s2;
    }
}

// One block per vector / point.
SIMT_DEVICE void hyper_body(const HyperParams& P, unsigned char* smem) {
    double* red = reinterpret_cast<double*>(smem);            // 32 doubles of reduction scratch
    double* bc = red + 40;                                    // broadcast slot
    double* bufA = reinterpret_cast<double*>(smem + kSmemHeader);
    double* bufB = bufA + P.N;
    double* lb = bufB + P.N;                                  // nb
    const int tid = SIMT_TID, nt = SIMT_NTID;
    const long long s = SIMT_BID, N = P.N, nth = N + 2;
    const long long j0 = P.start, j1 = P.start + P.nb * P.grain;
    const long long pi = P.st_shared ? 0 : s;
    double* theta = P.st.theta + pi * nth;
    double* A = P.st.A + pi * N;
    double* fine = P.st.fine + pi * N;
    double* lam = P.st.lam + pi * P.nb;

    if (P.mode == HM_POINT) {
        // ---- state of one point: theta, A, fine, Lambda (+ objective terms, + its gradient) ----------------
        int i; double sg;
        point_decode(P.point_mode, (int)s, i, sg);
        double pr = 0.0;
        for (long long a = tid; a < nth; a += nt) {
            double v = P.x[a];
            if (P.point_mode != POINT_CENTER) v += sg * P.R[(long long)i * nth + a];
            theta[a] = v;
            pr += 0.5 * v * v;                                // prior 1/2 |p|^2 (src/mgvi_impl.jl:3-7)
        }
        SIMT_SYNC();
        const double ampl = P.a0 * exp(P.a1 * theta[0]);
        const double cl = P.l0 * exp(P.l1 * theta[1]);
        const double pref = ampl * sqrt(2.0 * M_PI * cl);
        for (long long k = tid; k < N; k += nt) {
            const double kd = P.kd[k];
            const double a = pref * exp(-9.869604401089358 * kd * kd * cl * cl);
            A[k] = a;
            bufA[k] = a * theta[2 + k];
        }
        SIMT_SYNC();
        for (long long j = tid; j < N; j += nt) {
            double acc = 0.0;
            for (long long k = 0; k < N; ++k) acc += SIMT_LDG(P.cas + k * N + j) * bufA[k];
            const double f = exp(P.c * acc);
            fine[j] = f;
            bufB[j] = f;
        }
        SIMT_SYNC();
        double nl = 0.0;
        for (long long b = tid; b < P.nb; b += nt) {
            double t = 0.0;
            for (int g = 0; g < P.grain; ++g) t += bufB[j0 + b * P.grain + g];
            const double L = P.binsize * t;
            lam[b] = L;
            const double d = P.data[b];
            nl += L - d * log(L);                             // variable part of -logpdf(Poisson(L), d)
            lb[b] = d / L - 1.0;                              // dloglike / dLambda
        }
        if (P.red) {
            const double tp = hyper_block_sum(pr, red, bc);
            const double tn = hyper_block_sum(nl, red, bc);
            if (tid == 0) { P.red[s] = tp; P.red[P.nvec + s] = tn; }
        }
        SIMT_SYNC();
        if (P.want_grad) {
            // grad of -log posterior at this point: theta - J' dloglike/dLambda
            double* out = P.out + s * P.out_ld;
            for (long long b = tid; b < P.nb; b += nt) lb[b] = -lb[b];
            SIMT_SYNC();
            hyper_adjoint(P, theta, A, fine, lb, bufA, bufB, theta, out, red, bc);
        }
        return;
    }
    if (P.mode == HM_RHS) {
        // b = J'(sqrt(F) . n1) + eta   (src/residual_samplers.jl:75-77)
        for (long long b = tid; b < P.nb; b += nt) lb[b] = sqrt(1.0 / lam[b]) * P.Z1[s * P.nb + b];
        SIMT_SYNC();
        hyper_adjoint(P, theta, A, fine, lb, bufA, bufB, P.Z2 + s * nth, P.out + s * P.out_ld, red, bc);
        return;
    }
    // ---- HM_OP: out = v + J' F J v -----------------------------------------------------------------------
    const double* v = P.V + s * P.v_ld;
    const double cl = P.l0 * exp(P.l1 * theta[1]);
    const double v0 = v[0], v1 = v[1];
    for (long long k = tid; k < N; k += nt) {
        const double a = A[k];
        const double dA = a * (v0 * P.a1 + v1 * hyper_dA2_factor(P, P.kd[k], cl));
        bufA[k] = a * v[2 + k] + dA * theta[2 + k];
    }
    SIMT_SYNC();
    // only the pixels under data bins are observed
    for (long long j = j0 + tid; j < j1; j += nt) {
        double acc = 0.0;
        for (long long k = 0; k < N; ++k) acc += SIMT_LDG(P.cas + k * N + j) * bufA[k];
        bufB[j] = fine[j] * (P.c * acc);
    }
    SIMT_SYNC();
    for (long long b = tid; b < P.nb; b += nt) {
        double t = 0.0;
        for (int g = 0; g < P.grain; ++g) t += bufB[j0 + b * P.grain + g];
        lb[b] = (P.binsize * t) * (1.0 / lam[b]);             // F . (J v)
    }
    SIMT_SYNC();
    hyper_adjoint(P, theta, A, fine, lb, bufA, bufB, v, P.out + s * P.out_ld, red, bc);
}

SIMT_GLOBAL(256) k_hyper(SIMT_KARGS(HyperParams)) {
    SIMT_SMEM_DECL
    hyper_body(P, simt_smem);
}

namespace {

const long long kHyperMaxN = 4096;

struct HyperEngine : Engine {
    long long N = 0, nb = 0, start = 0;
    int grain = 1;
    double c = 1, binsize = 1, hy[4] = {1, 1, 1, 1};
    double const_term = 0.0;
    DevBuf kd, cas, data;
    struct State {
        DevBuf theta, A, fine, lam;
        long long npts = 0;
        void ensure(long long n, long long N, long long nb) {
            theta.ensure(8 * n * (N + 2)); A.ensure(8 * n * N); fine.ensure(8 * n * N); lam.ensure(8 * n * nb);
            npts = n;
        }
        HyperPointState view() const {
            return HyperPointState{theta.as<double>(), A.as<double>(), fine.as<double>(), lam.as<double>()};
        }
        void release() { theta.release(); A.release(); fine.release(); lam.release(); }
    } st_lin, st_mm, st_kl;
    long long n_mm = 0;
    DevBuf kl_red, pt_g, sc_gather, gather, mm_out, Mdense, eye;
    DevBuf ws_r, ws_p, ws_t, n_x, n_r, n_p, n_t;
    ReduceWs rws;
    CgWs cgws, ncgws;
    CgScalars ncg_sc;
    bool ncg_first = true;
    int h_act = 0;
    DenseChol* chol = nullptr;

    ~HyperEngine() override {
        DevBuf* all[] = {&kd, &cas, &data, &kl_red, &pt_g, &sc_gather, &gather, &mm_out, &Mdense, &eye, &ws_r, &ws_p, &ws_t,
                         &n_x, &n_r, &n_p, &n_t, &rws.partials, &rws.counters, &rws.red};
        for (auto* b : all) b->release();
        st_lin.release(); st_mm.release(); st_kl.release();
        DevBuf* cg[] = {&cgws.gamma, &cgws.alpha, &cgws.beta, &cgws.eps, &cgws.rnorm, &cgws.pAp, &cgws.iters,
                        &cgws.active, &cgws.n_active, &ncgws.gamma, &ncgws.alpha, &ncgws.beta, &ncgws.eps,
                        &ncgws.rnorm, &ncgws.pAp, &ncgws.iters, &ncgws.active, &ncgws.n_active};
        for (auto* b : cg) b->release();
        dense_chol_destroy(chol);
    }

    void setup(const mgvib200_model_desc& d) {
        if (d.ndim != 1) throw RtError{"FIELD_HYPER: one-dimensional fields (ndim = 1)", 1};
        N = d.dims[0]; nb = d.n_data; start = d.bin_start; grain = d.grain;
        if (N < 2 || N > kHyperMaxN) throw RtError{"FIELD_HYPER: 2 <= N <= 4096", 3};
        if (d.likelihood != MGVIB200_LIKE_POISSON || d.link != MGVIB200_LINK_EXP)
            throw RtError{"FIELD_HYPER: Poisson counts with the exponential link (the tutorial's model)", 3};
        if (!d.kdist || !d.data) throw RtError{"FIELD_HYPER: kdist and data are required", 1};
        if (grain < 1 || nb < 1 || start < 0 || start + nb * (long long)grain > N)
            throw RtError{"FIELD_HYPER: the data bins must lie inside the field", 1};
        c = d.scale; binsize = d.binsize;
        for (int i = 0; i < 4; ++i) hy[i] = d.hyper[i];
        if (!(hy[0] > 0) || !(hy[2] > 0) || !(binsize > 0)) throw RtError{"FIELD_HYPER: a0, l0 and binsize must be positive", 1};
        n_theta = N + 2;
        n_lambda = nb;
        kd.ensure(8 * N); data.ensure(8 * nb); cas.ensure(8 * N * N);
        rt_h2d(kd.p, d.kdist, 8 * N, ctx->stream);
        rt_h2d(data.p, d.data, 8 * nb, ctx->stream);
        std::vector<double> t((size_t)N * N);
        for (long long k = 0; k < N; ++k)
            for (long long j = 0; j < N; ++j) {
                const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((j * k) % N) / (long double)N;
                t[(size_t)k * N + j] = (double)(cosl(a) + sinl(a));
            }
        rt_h2d(cas.p, t.data(), 8 * N * N, ctx->stream);
        double cl = 0.0;
        for (long long b = 0; b < nb; ++b) cl += lgamma(d.data[b] + 1.0);
        const_term = cl + 0.5 * (double)n_theta * log(2.0 * M_PI);
        rt_sync(ctx->stream);
    }

    HyperParams base() const {
        HyperParams P;
        memset(&P, 0, sizeof P);
        P.N = N; P.nb = nb; P.start = start; P.grain = grain;
        P.c = c; P.binsize = binsize; P.a0 = hy[0]; P.a1 = hy[1]; P.l0 = hy[2]; P.l1 = hy[3];
        P.kd = kd.as<double>(); P.cas = cas.as<double>(); P.data = data.as<double>();
        return P;
    }

    void launch(const HyperParams& P) {
        if (P.nvec <= 0) return;
        const size_t smem = kSmemHeader + 8 * (size_t)(2 * N + nb);
        ctx->prof_pre(6000 + P.mode);
        rt_launch(k_hyper, P.nvec, 256, smem, ctx->stream, P);
        ctx->prof_post();
        ctx->launches++;
    }

    void points(State& st, const double* x, const double* R, int npts, int point_mode, double* red, double* grads) {
        st.ensure(npts, N, nb);
        HyperParams P = base();
        P.mode = HM_POINT; P.nvec = npts; P.st = st.view(); P.st_shared = 0;
        P.x = x; P.R = R; P.point_mode = point_mode; P.red = red;
        P.want_grad = grads ? 1 : 0; P.out = grads; P.out_ld = n_theta;
        launch(P);
    }

    // out[s] = M(point of s) v_s for nvec vectors
    void op(const State& st, bool shared_state, const double* V, long long v_ld, int nvec, double* out) {
        HyperParams P = base();
        P.mode = HM_OP; P.nvec = nvec; P.st = st.view(); P.st_shared = shared_state ? 1 : 0;
        P.V = V; P.v_ld = v_ld; P.out = out; P.out_ld = n_theta;
        launch(P);
    }

    void linearize(const double* x) override { points(st_lin, x, nullptr, 1, POINT_CENTER, nullptr, nullptr); }

    void get_weights(double* w_host, double* q_host) override {
        std::vector<double> L(nb);
        rt_d2h(L.data(), st_lin.lam.p, 8 * nb, ctx->stream);
        rt_sync(ctx->stream);
        for (long long b = 0; b < nb; ++b) {
            if (w_host) w_host[b] = 1.0 / L[b];               // src/information.jl:74-78
            if (q_host) q_host[b] = sqrt(1.0 / L[b]);
        }
    }

    void metric_apply(const double* V, int64_t k, double* MV) override { op(st_lin, true, V, n_theta, (int)k, MV); }

    VecParams vp() const {
        VecParams P;
        memset(&P, 0, sizeof P);
        P.nvec = 1; P.N = n_theta;
        return P;
    }

    void sample_cg(const double* Z1, const double* Z2, int64_t n, const mgvib200_cg_opts& o, double* R,
                   int64_t* iters_host, double* rnorm_host) override {
        const int nvec = (int)n;
        const long long nt = n_theta;
        const size_t vb = 8 * (size_t)nt * nvec;
        ws_r.ensure(vb); ws_p.ensure(vb); ws_t.ensure(vb);
        cgws.ensure(nvec);
        const long long itmax = o.maxiters > 0 ? o.maxiters : nt;
        CgScalars sc = cgws.scalars(o.abstol, o.reltol, itmax, CG_KRYLOV);
        rt_memset(cgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        rt_event_record(ctx->ev0, ctx->stream);
        {
            HyperParams P = base();
            P.mode = HM_RHS; P.nvec = nvec; P.st = st_lin.view(); P.st_shared = 1;
            P.Z1 = Z1; P.Z2 = Z2; P.out = ws_t.as<double>(); P.out_ld = nt;
            launch(P);
        }
        vec_cg_init(ctx, rws, sc, ws_t.as<double>(), nt, R, ws_r.as<double>(), ws_p.as<double>(), nt, nvec);
        int* hp = ctx->h_pinned;
        rt_d2h(hp, cgws.n_active.p, sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        long long it = 0;
        int poll = 1;
        // Krylov.jl cg!, lock-step over the residual samples; inactive samples are skipped by
        // the vector kernels and their (discarded) products are simply recomputed
        while (hp[0] > 0 && it < itmax) {
            for (int q = 0; q < poll && it < itmax; ++q, ++it) {
                if (it > 0) {
                    VecParams P = vp();
                    P.op = VOP_CG_PUPD; P.nvec = nvec; P.r = ws_r.as<double>(); P.p = ws_p.as<double>(); P.cg = sc;
                    vec_launch(ctx, rws, P);
                }
                op(st_lin, true, ws_p.as<double>(), nt, nvec, ws_t.as<double>());
                {
                    VecParams P = vp();
                    P.op = VOP_CG_PAP; P.nvec = nvec; P.p = ws_p.as<double>(); P.Ap = ws_t.as<double>(); P.cg = sc;
                    P.fin = FIN_CG_PAP;
                    vec_launch(ctx, rws, P);
                }
                vec_cg_update(ctx, rws, sc, R, ws_r.as<double>(), ws_p.as<double>(), ws_t.as<double>(), nt, nvec);
            }
            rt_d2h(hp, cgws.n_active.p, sizeof(int), ctx->stream);
            rt_sync(ctx->stream);
            if (poll < 8) poll *= 2;
        }
        rt_event_record(ctx->ev1, ctx->stream);
        std::vector<int> hi(nvec);
        std::vector<double> hr(nvec);
        rt_d2h(hi.data(), cgws.iters.p, sizeof(int) * nvec, ctx->stream);
        rt_d2h(hr.data(), cgws.rnorm.p, 8 * nvec, ctx->stream);
        rt_sync(ctx->stream);
        ctx->sampler_ms = rt_event_ms(ctx->ev0, ctx->ev1);
        long long mx = 0;
        for (int i = 0; i < nvec; ++i) {
            if (iters_host) iters_host[i] = hi[i];
            if (rnorm_host) rnorm_host[i] = hr[i];
            mx = std::max<long long>(mx, hi[i]);
        }
        ctx->lockstep_iters = mx;
    }

    // MatrixInversion (src/residual_samplers.jl:95-123): M materialised by n_theta metric products on the
    // identity (:102-104), then the shared Cholesky sampler
    void sample_dense(const double* Z, int64_t n, double* R, double* L) override {
        const long long nt = n_theta;
        Mdense.ensure(8 * nt * nt); eye.ensure(8 * nt * nt);
        if (!chol) chol = dense_chol_create(ctx);
        dense_chol_identity(chol, eye.as<double>(), nt);
        op(st_lin, true, eye.as<double>(), nt, (int)nt, Mdense.as<double>());
        dense_chol_sample(chol, Mdense.as<double>(), nt, Z, n, R, L);
    }

    double kl_value_grad(const double* x, const double* R, int64_t n, double* grad) override {
        const int npoints = (int)(2 * n);
        const CanonShape cs = canon_shape(ctx, n);
        kl_red.ensure(8 * 2 * (size_t)npoints);
        double* red = kl_red.as<double>();
        if (grad) pt_g.ensure(8 * (size_t)npoints * n_theta);
        points(st_kl, x, R, npoints, POINT_PM, red, grad ? pt_g.as<double>() : nullptr);
        const double ftot = canon_scalar_sum(ctx, sc_gather, red, 2, npoints);
        if (grad)
            vec_canon_sum(ctx, rws, cs, pt_g.as<double>(), n_theta, 2, 1.0 / (2.0 * (double)cs.n_glob), grad, n_theta, gather);
        return ftot / (2.0 * (double)cs.n_glob) + const_term;
    }

    CanonShape mm_cs;
    void mean_metric_prepare(const double* x, const double* R, int64_t n) override {
        mm_cs = canon_shape(ctx, n);
        n_mm = n;
        points(st_mm, x, R, (int)n, POINT_PLUS, nullptr, nullptr);     // linearisations at x + r_i (mgvi_impl.jl:137-138)
    }

    void mean_metric_apply(const double* u, double* out) override {
        mm_out.ensure(8 * (size_t)n_mm * n_theta);
        op(st_mm, false, u, 0, (int)n_mm, mm_out.as<double>());
        vec_canon_sum(ctx, rws, mm_cs, mm_out.as<double>(), n_theta, 1, 1.0 / (double)mm_cs.n_glob, out, n_theta, gather);
    }

    void ncg_alloc() {
        const size_t b = 8 * n_theta;
        n_x.ensure(b); n_r.ensure(b); n_p.ensure(b); n_t.ensure(b);
        ncgws.ensure(1);
        // IterativeSolvers cg_iterator!(dx, A, b; initially_zero = true): reltol = sqrt(eps), abstol = 0,
        // maxiter = size(A, 2)  (src/newtoncg.jl:120)
        ncg_sc = ncgws.scalars(0.0, 1.4901161193847656e-08, n_theta, CG_ITSOLVERS);
    }
    void ncg_init(const double* g) override {
        ncg_alloc();
        rt_memset(ncgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        vec_cg_init(ctx, rws, ncg_sc, g, n_theta, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), n_theta, 1);
        ncg_first = true;
    }
    void ncg_set_state(const double* x, const double* r, const double* u, double res, double prev) override {
        ncg_alloc();
        const size_t b = 8 * n_theta;
        rt_d2d(n_x.p, x, b, ctx->stream);
        rt_d2d(n_r.p, r, b, ctx->stream);
        rt_d2d(n_p.p, u, b, ctx->stream);
        const double hs[4] = {res * res, (res * res) / (prev * prev), res, 0.0};
        const int hi[2] = {0, 1};
        rt_h2d(ncgws.gamma.p, &hs[0], sizeof(double), ctx->stream);
        rt_h2d(ncgws.beta.p, &hs[1], sizeof(double), ctx->stream);
        rt_h2d(ncgws.rnorm.p, &hs[2], sizeof(double), ctx->stream);
        rt_h2d(ncgws.eps.p, &hs[3], sizeof(double), ctx->stream);
        rt_h2d(ncgws.iters.p, &hi[0], sizeof(int), ctx->stream);
        rt_h2d(ncgws.active.p, &hi[1], sizeof(int), ctx->stream);
        rt_h2d(ncgws.n_active.p, &hi[1], sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        ncg_first = false;
    }
    bool ncg_active() override {
        rt_d2h(&h_act, ncgws.active.p, sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        return h_act != 0;
    }
    void ncg_update() override {
        if (!ncg_first) {
            VecParams P = vp();
            P.op = VOP_CG_PUPD; P.r = n_r.as<double>(); P.p = n_p.as<double>(); P.cg = ncg_sc;
            vec_launch(ctx, rws, P);
        }
        mean_metric_apply(n_p.as<double>(), n_t.as<double>());
        VecParams P = vp();
        P.op = VOP_CG_PAP; P.p = n_p.as<double>(); P.Ap = n_t.as<double>(); P.cg = ncg_sc; P.fin = FIN_CG_PAP;
        vec_launch(ctx, rws, P);
        vec_cg_update(ctx, rws, ncg_sc, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), n_t.as<double>(), n_theta, 1);
        ncg_first = false;
    }
    const double* ncg_dx() override { return n_x.as<double>(); }
    double ncg_residual() override {
        double h = 0.0;
        rt_d2h(&h, ncgws.rnorm.p, sizeof(double), ctx->stream);
        rt_sync(ctx->stream);
        return h;
    }
};

}  // namespace

Engine* make_hyper_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d) {
    HyperEngine* e = new HyperEngine();
    e->ctx = ctx;
    try {
        e->setup(d);
    } catch (...) {
        delete e;
        throw;
    }
    return e;
}

}  // namespace mgvi
