// Polynomial-fit engine -- test/test_models/model_polyfit.jl:15-24 (BASELINE.json configs[0]):
//   mean_i = sum_k coef_scale[k] p[k] x_i^k,  sigma = noise_scale * p[P]^2,  Normal likelihood,
//   one product of Normals per grid inside a NamedTupleDist.
// n_theta = P + 1 is tiny, so the metric J'FJ (n_theta x n_theta) is assembled by block
// reductions over the data points and every solve runs in single-warp kernels (small_dense.h).
// lambda layout (SURVEY.md note N1): BLOCKED = [mu_g; var_g] per group, INTERLEAVED = [mu, sigma]
// pairs per group.  F = [1/sigma^2 ; 2/sigma^2] in both (src/information.jl:12-18,60-65).
#include <algorithm>
#include "engine.h"
#include "kernels.h"
#include "small_dense.h"

namespace mgvi {

struct PolyParams {
    const double* x; const double* data; const double* cs;
    const long long* mu_idx; const long long* noise_idx;
    long long N;
    int P, layout, nvec, mode;
    double ns;
    // operands
    const double* p0;        // centre / evaluation point (P+1)
    const double* R;         // residuals (P+1) x n
    const double* Z1; long long Z1_ld; const double* Z2;
    double* out;             // mode dependent
    double* S;               // P x P gram matrix of the design
    double* M;               // (P+1) x (P+1) metric (row-major)
    int npoints, point_mode;
    double inv_n;
};

enum { PM_GRAM = 0, PM_RHS = 1, PM_KL = 2 };

SIMT_DEVICE double poly_basis(const PolyParams& P, long long i, int k) {
    double xk = 1.0;
    const double xi = P.x[i];
    for (int j = 0; j < k; ++j) xk *= xi;
    return P.cs[k] * xk;
}

SIMT_DEVICE void poly_point(const PolyParams& P, int s, int a, double& val) {
    int i; double sg;
    point_decode(P.point_mode, s, i, sg);
    val = P.p0[a] + (P.point_mode == POINT_CENTER ? 0.0 : sg * P.R[(long long)i * (P.P + 1) + a]);
}

// One block per output scalar; the block reduces over the N data points.
//  PM_GRAM: block (a, b)            -> S[a][b] = sum_i X[i,a] X[i,b]
//  PM_RHS : block (vector s, comp a) -> out[s][a] = J'(sqrt(F) n1)[a] + eta[a]
//  PM_KL  : block (point s, q)       -> out[s][q]: q = 0 variable part of -loglike + prior,
//                                       q = 1..P  sum_i X[i,q-1] (d - mu)/sigma^2,
//                                       q = P+1   sum_i (d - mu)^2/sigma^3 - 1/sigma
SIMT_DEVICE void poly_reduce_body(const PolyParams& P, unsigned char* smem) {
    double* red = reinterpret_cast<double*>(smem);
    const int np1 = P.P + 1;
    double acc = 0.0;
    const long long bid = SIMT_BID;
    if (P.mode == PM_GRAM) {
        const int a = (int)(bid / P.P), b = (int)(bid % P.P);
        for (long long i = SIMT_TID; i < P.N; i += SIMT_NTID) acc += poly_basis(P, i, a) * poly_basis(P, i, b);
        const double tot = seg_reduce(acc, SIMT_NTID, red);
        if (SIMT_TID == 0) P.S[a * P.P + b] = tot;
    } else if (P.mode == PM_RHS) {
        const int s = (int)(bid / np1), a = (int)(bid % np1);
        const double sig = P.ns * P.p0[P.P] * P.p0[P.P];
        const double inv = 1.0 / (sig * sig);
        const double ds_dp = 2.0 * P.ns * P.p0[P.P];
        const double dn_dp = (P.layout == MGVIB200_LAYOUT_INTERLEAVED) ? ds_dp : 2.0 * sig * ds_dp;
        const double* z = P.Z1 + (long long)s * P.Z1_ld;
        if (a < P.P) {
            const double sf = sqrt(inv);
            for (long long i = SIMT_TID; i < P.N; i += SIMT_NTID) acc += poly_basis(P, i, a) * (sf * z[P.mu_idx[i]]);
        } else {
            const double sf = sqrt(2.0 * inv);
            for (long long i = SIMT_TID; i < P.N; i += SIMT_NTID) acc += dn_dp * (sf * z[P.noise_idx[i]]);
        }
        const double tot = seg_reduce(acc, SIMT_NTID, red);
        if (SIMT_TID == 0) P.out[(long long)s * np1 + a] = tot + P.Z2[(long long)s * np1 + a];
    } else {
        const int nq = P.P + 2;
        const int s = (int)(bid / nq), q = (int)(bid % nq);
        double pv[kSmallMax];
        for (int a = 0; a < np1; ++a) poly_point(P, s, a, pv[a]);
        const double sig = P.ns * pv[P.P] * pv[P.P];
        for (long long i = SIMT_TID; i < P.N; i += SIMT_NTID) {
            double mu = 0.0;
            for (int k = 0; k < P.P; ++k) mu += pv[k] * poly_basis(P, i, k);
            const double rres = P.data[i] - mu;
            if (q == 0) { const double zz = rres / sig; acc += 0.5 * zz * zz; }
            else if (q <= P.P) acc += poly_basis(P, i, q - 1) * (rres / (sig * sig));
            else acc += rres * rres / (sig * sig * sig) - 1.0 / sig;
        }
        double tot = seg_reduce(acc, SIMT_NTID, red);
        if (SIMT_TID == 0) {
            if (q == 0) {
                // + N (log|sigma| + 1/2 log 2pi) + 1/2 |p|^2 + n_theta/2 log 2pi   (src/mgvi_impl.jl:3-14)
                double pp = 0.0;
                for (int a = 0; a < np1; ++a) pp += pv[a] * pv[a];
                tot += (double)P.N * (log(fabs(sig)) + 0.9189385332046727418) + 0.5 * pp + 0.5 * np1 * 1.8378770664093454836;
            }
            P.out[(long long)s * nq + q] = tot;
        }
    }
}

// One warp: finishes KL value / gradient over the points, or builds the (mean) metric.
//  mode 0: out[0] = f = mean_s kl[s][0];  out[1 + a] = grad[a]
//  mode 1: M = I + mean_s G(point s),  G = S / sigma_s^2 (+) N F_n (dn/dp)^2
struct PolyFinalParams {
    PolyParams p;
    const double* kl;     // npoints x (P+2)
    int mode;
};

SIMT_DEVICE void poly_final_body(const PolyFinalParams& Q, unsigned char*) {
    const PolyParams& P = Q.p;
    const int lane = SIMT_TID & 31, np1 = P.P + 1, nq = P.P + 2;
    if (Q.mode == 0) {
        if (lane == 0) {
            double f = 0.0;
            for (int s = 0; s < P.npoints; ++s) f += Q.kl[(long long)s * nq];
            P.out[0] = f * P.inv_n;
        }
        if (lane < np1) {
            double g = 0.0;
            for (int s = 0; s < P.npoints; ++s) {
                double pa;
                poly_point(P, s, lane, pa);
                double pl;
                poly_point(P, s, P.P, pl);
                double jt;
                if (lane < P.P) jt = Q.kl[(long long)s * nq + 1 + lane];
                else jt = Q.kl[(long long)s * nq + 1 + P.P] * (2.0 * P.ns * pl);
                g += pa - jt;
            }
            P.out[1 + lane] = g * P.inv_n;
        }
    } else {
        double minv = 0.0, mnoise = 0.0;
        for (int s = 0; s < P.npoints; ++s) {
            double pl;
            poly_point(P, s, P.P, pl);
            const double sig = P.ns * pl * pl;
            const double inv = 1.0 / (sig * sig);
            const double ds_dp = 2.0 * P.ns * pl;
            const double dn_dp = (P.layout == MGVIB200_LAYOUT_INTERLEAVED) ? ds_dp : 2.0 * sig * ds_dp;
            minv += inv;
            mnoise += (double)P.N * (2.0 * inv) * dn_dp * dn_dp;
        }
        minv *= P.inv_n; mnoise *= P.inv_n;
        for (int i = lane; i < np1 * np1; i += 32) {
            const int a = i / np1, b = i % np1;
            double v = (a == b) ? 1.0 : 0.0;
            if (a < P.P && b < P.P) v += minv * P.S[a * P.P + b];
            if (a == P.P && b == P.P) v += mnoise;
            P.M[i] = v;
        }
    }
}

// y = M V for k columns (one warp per column)
struct SmallApplyParams { const double* M; const double* V; double* out; int n; };
SIMT_DEVICE void small_apply_body(const SmallApplyParams& P, unsigned char* smem) {
    double* Msh = reinterpret_cast<double*>(smem);
    double* vs = Msh + kSmallMax * kSmallMax;
    const int lane = SIMT_TID & 31, n = P.n;
    for (int i = lane; i < n * n; i += 32) Msh[i] = P.M[i];
    SIMT_SYNC();
    const long long s = SIMT_BID;
    const double v = lane < n ? P.V[s * n + lane] : 0.0;
    const double y = small_matvec(Msh, n, v, vs);
    if (lane < n) P.out[s * n + lane] = y;
}

SIMT_GLOBAL(256) k_poly_reduce(SIMT_KARGS(PolyParams)) { SIMT_SMEM_DECL poly_reduce_body(P, simt_smem); }
SIMT_GLOBAL(32) k_poly_final(SIMT_KARGS(PolyFinalParams)) { SIMT_SMEM_DECL poly_final_body(P, simt_smem); }
SIMT_GLOBAL(32) k_small_cg(SIMT_KARGS(SmallCgParams)) { SIMT_SMEM_DECL small_cg_body(P, simt_smem); }
SIMT_GLOBAL(32) k_small_ncg(SIMT_KARGS(SmallNcgParams)) { SIMT_SMEM_DECL small_ncg_body(P, simt_smem); }
SIMT_GLOBAL(32) k_small_dense(SIMT_KARGS(SmallDenseParams)) { SIMT_SMEM_DECL small_dense_body(P, simt_smem); }
SIMT_GLOBAL(32) k_small_apply(SIMT_KARGS(SmallApplyParams)) { SIMT_SMEM_DECL small_apply_body(P, simt_smem); }

namespace {

const size_t kSmallSmem = sizeof(double) * (2 * kSmallMax * kSmallMax + 64);

struct PolyEngine : Engine {
    long long N = 0;
    int P = 0, layout = 0;
    double ns = 60.0;
    DevBuf x, data, cs, mu_idx, noise_idx, S, M, Mbar, center, rhs, kl, fin, n_x, n_r, n_u, n_st, cg_it, cg_rn;
    double h_st[8];

    bool counted = false;

    ~PolyEngine() override {
        if (counted) ctx->n_models_single_rank--;
        DevBuf* all[] = {&x, &data, &cs, &mu_idx, &noise_idx, &S, &M, &Mbar, &center, &rhs, &kl, &fin, &n_x, &n_r,
                         &n_u, &n_st, &cg_it, &cg_rn};
        for (auto* b : all) b->release();
    }

    PolyParams base() const {
        PolyParams p;
        memset(&p, 0, sizeof p);
        p.x = x.as<double>(); p.data = data.as<double>(); p.cs = cs.as<double>();
        p.mu_idx = mu_idx.as<long long>(); p.noise_idx = noise_idx.as<long long>();
        p.N = N; p.P = P; p.layout = layout; p.ns = ns;
        p.S = S.as<double>();
        return p;
    }

    void launch_reduce(const PolyParams& p, long long grid) {
        ctx->prof_pre(2000 + p.mode);
        rt_launch(k_poly_reduce, grid, 256, 512, ctx->stream, p);
        ctx->prof_post();
        ctx->launches++;
    }

    void setup(const mgvib200_model_desc& d) {
        single_rank_only();
        N = d.n_data; P = d.n_coef; layout = d.lambda_layout; ns = d.noise_scale;
        if (d.likelihood != MGVIB200_LIKE_NORMAL) throw RtError{"POLY: Normal likelihood only", 1};
        if (P < 1 || P + 1 > kSmallMax) throw RtError{"POLY: 1 <= n_coef <= 31", 1};
        if (layout != MGVIB200_LAYOUT_BLOCKED && layout != MGVIB200_LAYOUT_INTERLEAVED)
            throw RtError{"POLY: lambda_layout must be BLOCKED or INTERLEAVED (the noise level is a parameter)", 1};
        if (!d.x || !d.data || !d.coef_scale || !d.group_sizes || d.n_groups < 1)
            throw RtError{"POLY: x, data, coef_scale and group_sizes are required", 1};
        long long tot = 0;
        for (int g = 0; g < d.n_groups; ++g) tot += d.group_sizes[g];
        if (tot != N) throw RtError{"POLY: group sizes must sum to n_data", 1};
        n_theta = P + 1;
        n_lambda = 2 * N;
        std::vector<long long> mi(N), ni(N);
        long long off = 0, i0 = 0;
        for (int g = 0; g < d.n_groups; ++g) {
            const long long n = d.group_sizes[g];
            for (long long j = 0; j < n; ++j) {
                if (layout == MGVIB200_LAYOUT_BLOCKED) { mi[i0 + j] = off + j; ni[i0 + j] = off + n + j; }
                else { mi[i0 + j] = off + 2 * j; ni[i0 + j] = off + 2 * j + 1; }
            }
            off += 2 * n; i0 += n;
        }
        x.ensure(8 * N); data.ensure(8 * N); cs.ensure(8 * P); mu_idx.ensure(8 * N); noise_idx.ensure(8 * N);
        rt_h2d(x.p, d.x, 8 * N, ctx->stream);
        rt_h2d(data.p, d.data, 8 * N, ctx->stream);
        rt_h2d(cs.p, d.coef_scale, 8 * P, ctx->stream);
        rt_h2d(mu_idx.p, mi.data(), 8 * N, ctx->stream);
        rt_h2d(noise_idx.p, ni.data(), 8 * N, ctx->stream);
        rt_sync(ctx->stream);
        S.ensure(8 * P * P); M.ensure(8 * n_theta * n_theta); Mbar.ensure(8 * n_theta * n_theta);
        center.ensure(8 * n_theta); fin.ensure(8 * (n_theta + 1));
        n_x.ensure(8 * n_theta); n_r.ensure(8 * n_theta); n_u.ensure(8 * n_theta); n_st.ensure(8 * 8);
        PolyParams p = base();
        p.mode = PM_GRAM;
        launch_reduce(p, (long long)P * P);
        rt_sync(ctx->stream);
    }

    void build_metric(const double* xc, const double* R, int npoints, int point_mode, double inv_n, double* Mout) {
        PolyFinalParams Q;
        Q.p = base();
        Q.p.p0 = xc; Q.p.R = R; Q.p.npoints = npoints; Q.p.point_mode = point_mode; Q.p.inv_n = inv_n;
        Q.p.M = Mout; Q.kl = nullptr; Q.mode = 1;
        rt_launch(k_poly_final, 1, 32, 0, ctx->stream, Q);
        ctx->launches++;
    }

    void linearize(const double* xc) override {
        rt_d2d(center.p, xc, 8 * n_theta, ctx->stream);
        build_metric(center.as<double>(), nullptr, 1, POINT_CENTER, 1.0, M.as<double>());
    }

    void get_weights(double* w_host, double* q_host) override {
        // Fisher weights on every lambda row: [1/sigma^2 on mean rows, 2/sigma^2 on noise rows]
        std::vector<double> c(n_theta);
        rt_d2h(c.data(), center.p, 8 * n_theta, ctx->stream);
        rt_sync(ctx->stream);
        const double sig = ns * c[P] * c[P], inv = 1.0 / (sig * sig);
        std::vector<long long> mi(N), ni(N);
        rt_d2h(mi.data(), mu_idx.p, 8 * N, ctx->stream);
        rt_d2h(ni.data(), noise_idx.p, 8 * N, ctx->stream);
        rt_sync(ctx->stream);
        for (long long i = 0; i < N; ++i) {
            if (w_host) { w_host[mi[i]] = inv; w_host[ni[i]] = 2.0 * inv; }
            if (q_host) { q_host[mi[i]] = sqrt(inv); q_host[ni[i]] = sqrt(2.0 * inv); }
        }
    }

    void apply(const double* Mdev, const double* V, int64_t k, double* out) {
        SmallApplyParams A;
        A.M = Mdev; A.V = V; A.out = out; A.n = (int)n_theta;
        rt_launch(k_small_apply, k, 32, kSmallSmem, ctx->stream, A);
        ctx->launches++;
    }

    void metric_apply(const double* V, int64_t k, double* MV) override { apply(M.as<double>(), V, k, MV); }

    void sample_cg(const double* Z1, const double* Z2, int64_t n, const mgvib200_cg_opts& o, double* R,
                   int64_t* iters_host, double* rnorm_host) override {
        rhs.ensure(8 * n_theta * n);
        rt_event_record(ctx->ev0, ctx->stream);
        PolyParams p = base();
        p.mode = PM_RHS; p.p0 = center.as<double>(); p.Z1 = Z1; p.Z1_ld = n_lambda; p.Z2 = Z2;
        p.out = rhs.as<double>();
        launch_reduce(p, n * n_theta);
        DevBuf& it = cg_it;
        DevBuf& rn = cg_rn;
        it.ensure(sizeof(int) * n); rn.ensure(8 * n);
        SmallCgParams c;
        c.M = M.as<double>(); c.B = rhs.as<double>(); c.X = R; c.iters = it.as<int>(); c.rnorm = rn.as<double>();
        c.n = (int)n_theta; c.atol = o.abstol; c.rtol = o.reltol; c.itmax = o.maxiters > 0 ? o.maxiters : n_theta;
        ctx->prof_pre(2100);
        rt_launch(k_small_cg, n, 32, kSmallSmem, ctx->stream, c);
        ctx->prof_post();
        ctx->launches++;
        rt_event_record(ctx->ev1, ctx->stream);
        std::vector<int> hi(n);
        std::vector<double> hr(n);
        rt_d2h(hi.data(), it.p, sizeof(int) * n, ctx->stream);
        rt_d2h(hr.data(), rn.p, 8 * n, ctx->stream);
        rt_sync(ctx->stream);
        ctx->sampler_ms = rt_event_ms(ctx->ev0, ctx->ev1);
        long long mx = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (iters_host) iters_host[i] = hi[i];
            if (rnorm_host) rnorm_host[i] = hr[i];
            mx = std::max<long long>(mx, hi[i]);
        }
        ctx->lockstep_iters = mx;
    }

    void sample_dense(const double* Z, int64_t n, double* R, double* L) override {
        SmallDenseParams d;
        d.M = M.as<double>(); d.Z = Z; d.R = R; d.Lout = L; d.n = (int)n_theta; d.nvec = (int)n;
        rt_launch(k_small_dense, 1, 32, kSmallSmem, ctx->stream, d);
        ctx->launches++;
    }

    // single-GPU model ("replicas only"): mgvib200_comm_init refuses contexts that hold one, and a model
    // created after comm_init is refused in setup(); the guard here covers every other ordering
    void single_rank_only() const {
        if (ctx->world > 1) throw RtError{"POLY: multi-GPU sharding is not provided for this latency-bound model", 3};
    }

    double kl_value_grad(const double* xc, const double* R, int64_t n, double* grad) override {
        single_rank_only();
        const int npoints = (int)(2 * n);
        const int64_t n_glob = n;
        const int nq = P + 2;
        kl.ensure(8 * (size_t)npoints * nq);
        PolyParams p = base();
        p.mode = PM_KL; p.p0 = xc; p.R = R; p.point_mode = POINT_PM; p.out = kl.as<double>();
        launch_reduce(p, (long long)npoints * nq);
        PolyFinalParams Q;
        Q.p = base();
        Q.p.p0 = xc; Q.p.R = R; Q.p.npoints = npoints; Q.p.point_mode = POINT_PM;
        Q.p.inv_n = 1.0 / (2.0 * (double)n_glob);
        Q.p.out = fin.as<double>(); Q.kl = kl.as<double>(); Q.mode = 0;
        rt_launch(k_poly_final, 1, 32, 0, ctx->stream, Q);
        ctx->launches++;
        double hf = 0.0;
        rt_d2h(&hf, fin.p, 8, ctx->stream);
        if (grad) rt_d2d(grad, fin.as<double>() + 1, 8 * n_theta, ctx->stream);
        rt_sync(ctx->stream);
        return hf;
    }

    void mean_metric_prepare(const double* xc, const double* R, int64_t n) override {
        single_rank_only();
        build_metric(xc, R, (int)n, POINT_PLUS, 1.0 / (double)n, Mbar.as<double>());
    }

    void mean_metric_apply(const double* u, double* out) override { apply(Mbar.as<double>(), u, 1, out); }

    void ncg_launch(const double* g, int init) {
        SmallNcgParams q;
        q.M = Mbar.as<double>(); q.g = g; q.x = n_x.as<double>(); q.r = n_r.as<double>(); q.u = n_u.as<double>();
        q.st = n_st.as<double>(); q.n = (int)n_theta; q.init = init; q.maxiter = n_theta;
        rt_launch(k_small_ncg, 1, 32, kSmallSmem, ctx->stream, q);
        ctx->launches++;
    }
    void ncg_init(const double* g) override { ncg_launch(g, 1); }
    bool ncg_active() override {
        rt_d2h(h_st, n_st.p, 8 * 5, ctx->stream);
        rt_sync(ctx->stream);
        return h_st[3] != 0.0;
    }
    void ncg_update() override { ncg_launch(nullptr, 0); }
    const double* ncg_dx() override { return n_x.as<double>(); }
    double ncg_residual() override {
        double h = 0.0;
        rt_d2h(&h, n_st.p, sizeof(double), ctx->stream);
        rt_sync(ctx->stream);
        return h;
    }
};

}  // namespace

Engine* make_poly_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d) {
    PolyEngine* e = new PolyEngine();
    e->ctx = ctx;
    try {
        e->setup(d);
    } catch (...) {
        delete e;
        throw;
    }
    ctx->n_models_single_rank++;
    e->counted = true;
    return e;
}

}  // namespace mgvi
