// Correlated-field engine: s = c * H(A . xi) + offset, lambda = g(s), Normal or Poisson data.
// Reference models: test/test_models/model_fft_gp.jl:60-72, docs/src/advanced_tutorial_lit.jl:253-307.
// Everything here is host orchestration of the kernels in field_pass.h / vec_ops.h; all pointers
// are device pointers.
#include <math.h>
#include <algorithm>
#include "engine.h"
#include "kernels.h"

namespace mgvi {

SIMT_GLOBAL(256) k_naive_axis(SIMT_KARGS(NaiveParams)) {
    SIMT_SMEM_DECL
    naive_axis_body(P, simt_smem);
}

template <int MAXT, int LGM>
SIMT_GLOBAL2(MAXT, (MAXT <= 256 ? 2 : 1)) k_bluestein_axis(SIMT_KARGS(BlueParams)) {      // 256: 128 registers, 512: 128, 1024: 64
    SIMT_SMEM_DECL
    bluestein_axis_body<LGM>(P, simt_smem);
}

namespace {

#ifndef MGVI_MINB
#define MGVI_MINB 3      // resident 256-thread blocks per SM the pass kernels are compiled for
#endif
const int kNaiveMaxAxis = 4096;
long long kBlueMinAxis = 32;           // (MGVIB200_BLUE_MIN overrides: measurements against the direct pass) generic path: axes longer than this take the O(m log m) Bluestein pass, shorter ones the direct O(n^2) one
const long long kFastMaxAxis = 4096;        // 1-D: longest line one block transforms (longer power-of-two lines: four-step split)
const long long kFastMaxAxisND = 8192;      // 2-D / 3-D: longest power-of-two axis (one complex line = 128 KB of shared memory, 1024 threads)
const long long kDenseFieldMax = 8192;

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2(long long v) { int l = 0; while ((1LL << l) < v) ++l; return l; }

struct PassGeo {
    int geo, n, log2n, L, G, threads, lgL, lg_tpo;
    long long inner, units_per_vec, ntb;
    size_t smem;
};

struct FieldEngine : Engine {
    int ndim = 1;
    long long dims[3] = {1, 1, 1};
    long long N = 0;
    bool fast = false;
    bool big = false;           // long 1-D field: four-step split n = n1 * n2 (big1d_pass.h)
    int lg_n1 = 0, lg_n2 = 0, lg_h = 0;
    DevBuf data_perm, tlo, thi, cs;
    double c = 1, offset = 0, sigma = 1, inv_sigma2 = 1;
    int link = LINK_EXP, like = LIKE_POISSON, layout = 0;
    long long z1_ld = 0, z1_stride = 1;
    double const_term = 0.0;    // per-point constants of -loglike and of the prior
    int max_L = 0;              // complex columns per strided tile; 0 = by axis length (strided_L), MGVIB200_STRIDED_L overrides
    int prefetch_waves = 0;     // software L2 prefetch distance in resident-block waves (0 = off: measured slower at C3, profiles/README.md)
    bool use_hot = true;        // compile-time-shape kernels for the benchmark sizes
    int prefetch_epi = 0;       // L1 hints for post-transform operands (MGVIB200_PREFETCH_EPI)
    bool use_paired = true;     // paired-arrangement transforms in the hot kernels (MGVIB200_PF=0: classic scheme)
    int stage_vin = 0;          // TMA bulk staging of the '+ v' rows in the hot last pass (needs -DMGVI_STAGE_VIN; MGVIB200_STAGE_VIN=1)

    DevBuf A, data, w, q, wbar, center, w_tiled, ws_tile;
    const double* w_tiled_src = nullptr;   // which weight array w_tiled mirrors, and at which generation
    long long w_gen = 0, w_tiled_gen = -1;
    int w_tiled_lg_wt = 0;
    int row_threads = 256;      // target block size of the row (contiguous) passes (MGVIB200_ROW_THREADS)
    int tiled_mode = -1;        // tile-major scratch for 2-D narrow column tiles: -1 auto, 0 off, 1 force (MGVIB200_TILED)
    std::map<long long, DevBuf> tw, twp, cas;     // twp: tables of the paired transforms (n = 4 * 8^s)
    std::map<long long, DevBuf> blue_chirp, blue_hf;   // any-length axes: Bluestein chirp and filter spectrum
    ReduceWs rws;
    CgWs cgws, ncgws;
    DevBuf ws_r, ws_p, ws_t;          // sampler CG vectors [nvec][N]
    DevBuf ws_p2;                     // second p buffer of the deferred solution update (field_pass.h: defer_x)
    // MGVIB200_DEFER_X=0: keep x += alpha p in the streaming update kernel.  Default on for 2-D / 3-D fields:
    // the 24 N bytes of the solution update ride in the fused axis-0 pass of the next iteration
    bool defer_force = false;
    int defer_mode = 1;               // 1: before the transforms of the block; 2: after them, operands prefetched into L2 at block start
    bool defer_on = true;
    DevBuf pt_t;                      // per-point scratch [npoints][N]
    DevBuf gather, sc_gather, kl_red; // canonical sums: gathered chunk partials [chunks][N], per-point scalars
    DevBuf qsum;                      // N
    DevBuf n_x, n_r, n_p, n_t;        // Newton CG state (single vector)
    DevBuf tmpA, tmpB;                // generic-path staging [nvec][N]
    DevBuf Mdense, eye;               // MatrixInversion sampler: materialised metric (small fields only)
    DenseChol* chol = nullptr;
    // replayable chunk of the sampler's lock-step CG iterations (sample_cg)
    struct CgGraphKey {
        // every buffer a captured kernel may point into: the CG vectors, the pass scratch, the staging buffers of
        // the generic and four-step paths and the reduction workspace (all of them can be re-allocated larger by a
        // later call with more vectors, e.g. a KL evaluation over 2n points)
        const void *R, *r, *p, *p2, *t, *tile, *wt, *ta, *tb, *cs, *part, *cnt;
        int nvec;
        double atol, rtol;
        long long itmax;
        bool operator==(const CgGraphKey& o) const {
            return R == o.R && r == o.r && p == o.p && p2 == o.p2 && t == o.t && tile == o.tile && wt == o.wt && ta == o.ta &&
                   tb == o.tb && cs == o.cs && part == o.part && cnt == o.cnt && nvec == o.nvec &&
                   atol == o.atol && rtol == o.rtol && itmax == o.itmax;
        }
    };
    static const int kCgChunk = 8;
    rt_graph cg_graph{};
    bool cg_graph_valid = false;
    CgGraphKey cg_graph_key{};
    int64_t cg_graph_launches = 0;
    bool graphs_on = true;          // MGVIB200_GRAPH=0 launches every kernel from the host
    // Two concurrent sample groups (sample_cg_grouped): with few samples per GPU (the 8-GPU share of C3 is 4) a
    // lock-step kernel is only ~9 waves of blocks and its ramp-up / tail is a visible fraction; two independent
    // half-size solves on two streams fill each other's tails (measured: 8 775 -> 9 149 sample-iterations/s at 4 samples)
    struct CgGroup {
        int off = 0, n = 0;
        rt_stream st{};
        ReduceWs rws;
        DevBuf tile, cs;                  // tile-major scratch (2-D) / complex scratch of the four-step path (long 1-D)
        rt_graph graph{};
        bool gvalid = false;
        CgGraphKey key{};
        int64_t glaunch = 0;
        rt_event ev[2]{}, ev_done{};
        int slot = 0;
        bool pending = false, done = false;
        long long it = 0;
        CgScalars sc{};
        int* h = nullptr;
    };
    CgGroup grp[2];
    bool groups_ready = false;
    int groups_mode = -1;             // MGVIB200_GROUPS: -1 auto (few waves per kernel), 1 never, 2 always (2-D / 3-D fast path)
    int big_block_threads = 64;     // (measured at C2: 64 -> 0.0475 ms per iteration, 256 -> 0.0490) MGVIB200_BIG_THREADS: target block size of the four-step kernels (units per block = this / unit size)
    int big_ap_lgL = 3;             // MGVIB200_BIG_AP_LGL: log2(lines per unit) of the column kernel A' (1 = one column pair)
    CgScalars ncg_sc;
    bool ncg_first = true;
    int h_ncg[2] = {0, 0};

    ~FieldEngine() override {
        if (cg_graph_valid) rt_graph_destroy(cg_graph);
        for (auto& G : grp) {
            if (G.gvalid) rt_graph_destroy(G.graph);
            G.rws.release(); G.tile.release(); G.cs.release();
            if (groups_ready) { rt_stream_destroy(G.st); rt_event_destroy(G.ev[0]); rt_event_destroy(G.ev[1]); rt_event_destroy(G.ev_done); }
        }
        dense_chol_destroy(chol);
        Mdense.release(); eye.release();
        DevBuf* all[] = {&A, &data, &w, &q, &wbar, &center, &w_tiled, &ws_tile, &ws_r, &ws_p, &ws_p2, &ws_t, &pt_t, &gather, &sc_gather, &kl_red, &qsum, &n_x, &n_r,
                         &n_p, &n_t, &tmpA, &tmpB, &cs, &tlo, &thi, &data_perm, &rws.partials, &rws.counters, &rws.red};
        for (auto* b : all) b->release();
        for (auto& kv : tw) kv.second.release();
        for (auto& kv : twp) kv.second.release();
        for (auto& kv : cas) kv.second.release();
        for (auto& kv : blue_chirp) kv.second.release();
        for (auto& kv : blue_hf) kv.second.release();
        DevBuf* cg[] = {&cgws.gamma, &cgws.alpha, &cgws.beta, &cgws.eps, &cgws.rnorm, &cgws.pAp, &cgws.iters,
                        &cgws.active, &cgws.n_active, &ncgws.gamma, &ncgws.alpha, &ncgws.beta, &ncgws.eps,
                        &ncgws.rnorm, &ncgws.pAp, &ncgws.iters, &ncgws.active, &ncgws.n_active};
        for (auto* b : cg) b->release();
    }

    // per-stage twiddle runs of the power-of-two FFT core (fft_core.h) for length n, built once per length
    void build_fft_tables(long long n) {
        if (tw.count(n)) return;
        // for each radix-8 stage with Ns > 1: W^j, W^2j, W^4j, j < Ns
        std::vector<double> t;
        const int lg = ilog2(n), lead = lg % 3;
        for (int lgNs = lead; lgNs < lg; lgNs += 3) {
            if (lgNs == 0) continue;
            const long long Ns = 1LL << lgNs;
            for (int e = 1; e <= 4; e *= 2)
                for (long long j = 0; j < Ns; ++j) {
                    long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)(e * j) /
                                    (long double)(8 * Ns);
                    t.push_back((double)cosl(a));
                    t.push_back((double)(-sinl(a)));
                }
        }
        if (t.empty()) { t.push_back(1.0); t.push_back(0.0); }
        tw[n].ensure(sizeof(double) * t.size());
        rt_h2d(tw[n].p, t.data(), sizeof(double) * t.size(), ctx->stream);
        rt_sync(ctx->stream);
        if (lead == 2 && lg >= 8) {
            // paired transforms (fft_core.h: fft_line_post): radix-8 stages Ns = 8, 64, ... (one run W^j each),
            // then the final radix-4 stage: exp(-2 pi i j / n), j < n/4
            std::vector<double> u;
            const long double tp = 2.0L * 3.14159265358979323846264338327950288L;
            for (long long Ns = 8; 8 * Ns <= n / 4; Ns *= 8)
                for (long long j = 0; j < Ns; ++j) {
                    const long double a = tp * (long double)j / (long double)(8 * Ns);
                    u.push_back((double)cosl(a)); u.push_back((double)(-sinl(a)));
                }
            for (long long j = 0; j < n / 4; ++j) {
                const long double a = tp * (long double)j / (long double)n;
                u.push_back((double)cosl(a)); u.push_back((double)(-sinl(a)));
            }
            twp[n].ensure(sizeof(double) * u.size());
            rt_h2d(twp[n].p, u.data(), sizeof(double) * u.size(), ctx->stream);
            rt_sync(ctx->stream);
        }
    }

    // Bluestein tables for an axis of any length n (field_pass.h: bluestein_axis_body): the chirp c_j =
    // exp(-i pi j^2 / n) and the length-m FFT of the wrapped conjugate chirp, scaled by 1/m, in long double
    static int blue_lg_m(long long n) { int lg = 8; while ((1LL << lg) < 2 * n - 1) ++lg; return lg; }
    void build_bluestein_tables(long long n) {
        if (blue_chirp.count(n)) return;
        const int lg_m = blue_lg_m(n);
        const long long m = 1LL << lg_m;
        build_fft_tables(m);
        const long double pi = 3.14159265358979323846264338327950288L;
        std::vector<long double> cr(n), ci(n);
        for (long long j = 0; j < n; ++j) {
            const long double a = pi * (long double)((j * j) % (2 * n)) / (long double)n;     // j^2 mod 2n keeps the argument small
            cr[j] = cosl(a); ci[j] = -sinl(a);
        }
        std::vector<long double> hr(m, 0.0L), hi(m, 0.0L);
        for (long long j = 0; j < n; ++j) {
            hr[j] = cr[j]; hi[j] = -ci[j];
            if (j) { hr[m - j] = cr[j]; hi[m - j] = -ci[j]; }
        }
        // iterative radix-2 FFT in long double (host, once per length)
        for (long long i = 1, j = 0; i < m; ++i) {
            long long bit = m >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { std::swap(hr[i], hr[j]); std::swap(hi[i], hi[j]); }
        }
        for (long long len = 2; len <= m; len <<= 1) {
            for (long long k = 0; k < len / 2; ++k) {
                const long double a = -2.0L * pi * (long double)k / (long double)len;
                const long double wr = cosl(a), wi = sinl(a);
                for (long long i = k; i < m; i += len) {
                    const long long j2 = i + len / 2;
                    const long double xr = hr[j2] * wr - hi[j2] * wi, xi = hr[j2] * wi + hi[j2] * wr;
                    hr[j2] = hr[i] - xr; hi[j2] = hi[i] - xi;
                    hr[i] += xr; hi[i] += xi;
                }
            }
        }
        std::vector<double> c(2 * n), h(2 * m);
        for (long long j = 0; j < n; ++j) { c[2 * j] = (double)cr[j]; c[2 * j + 1] = (double)ci[j]; }
        for (long long j = 0; j < m; ++j) { h[2 * j] = (double)(hr[j] / (long double)m); h[2 * j + 1] = (double)(hi[j] / (long double)m); }
        blue_chirp[n].ensure(16 * n); blue_hf[n].ensure(16 * m);
        rt_h2d(blue_chirp[n].p, c.data(), 16 * n, ctx->stream);
        rt_h2d(blue_hf[n].p, h.data(), 16 * m, ctx->stream);
        rt_sync(ctx->stream);
    }

    // ------------------------------------------------------------------------------------------
    void setup(const mgvib200_model_desc& d) {
        ndim = d.ndim;
        if (ndim < 1 || ndim > 3) throw RtError{"FIELD_DHT: ndim must be 1..3", 1};
        N = 1;
        for (int i = 0; i < ndim; ++i) {
            dims[i] = d.dims[i];
            if (dims[i] < 1) throw RtError{"FIELD_DHT: bad dims", 1};
            N *= dims[i];
        }
        if (d.n_data != N) throw RtError{"FIELD_DHT: n_data must equal prod(dims)", 1};
        if (!d.amplitude || !d.data) throw RtError{"FIELD_DHT: amplitude and data are required", 1};
        c = d.scale; offset = d.offset; link = d.link; like = d.likelihood; sigma = d.sigma;
        if (like == LIKE_NORMAL && !(sigma > 0)) throw RtError{"FIELD_DHT: sigma must be positive", 1};
        inv_sigma2 = (1.0 / sigma) * (1.0 / sigma);
        layout = (like == LIKE_POISSON) ? MGVIB200_LAYOUT_MU_ONLY : d.lambda_layout;
        n_theta = N;
        n_lambda = (layout == MGVIB200_LAYOUT_MU_ONLY) ? N : 2 * N;
        z1_ld = n_lambda;
        z1_stride = (layout == MGVIB200_LAYOUT_INTERLEAVED) ? 2 : 1;
        fast = true;
        for (int i = 0; i < ndim; ++i)
            if (!is_pow2(dims[i]) || dims[i] < 8 || dims[i] > (ndim >= 2 ? kFastMaxAxisND : kFastMaxAxis)) fast = false;
        long long big_min = kFastMaxAxis + 1;     // tests lower this to run the four-step path at small sizes
        if (const char* e = getenv("MGVIB200_BIG_MIN")) big_min = std::max<long long>(64, atoll(e));
        if (ndim == 1 && is_pow2(N) && N >= big_min && N <= kFastMaxAxis * kFastMaxAxis) {
            big = true; fast = true;
            const int lg = ilog2(N);
            lg_n1 = (lg + 1) / 2; lg_n2 = lg - lg_n1; lg_h = (lg + 1) / 2;
        }
        if (N >= (1LL << 31)) fast = false;      // the pass kernels use 32-bit element offsets
        if (const char* e = getenv("MGVIB200_FORCE_GENERIC")) if (atoi(e)) fast = false;
        if (const char* e = getenv("MGVIB200_STRIDED_L")) max_L = std::max(1, atoi(e));
        if (const char* e = getenv("MGVIB200_PREFETCH_WAVES")) prefetch_waves = std::max(0, atoi(e));
        if (const char* e = getenv("MGVIB200_NO_HOT")) use_hot = !atoi(e);
        if (const char* e = getenv("MGVIB200_PREFETCH_EPI")) prefetch_epi = atoi(e);
        if (const char* e = getenv("MGVIB200_STAGE_VIN")) stage_vin = atoi(e);
        if (const char* e = getenv("MGVIB200_TILED")) tiled_mode = atoi(e);
        if (const char* e = getenv("MGVIB200_GRAPH")) graphs_on = atoi(e) != 0;
        if (const char* e = getenv("MGVIB200_PF")) use_paired = atoi(e) != 0;
        if (const char* e = getenv("MGVIB200_GROUPS")) groups_mode = atoi(e);
        if (const char* e = getenv("MGVIB200_DEFER_X")) { defer_on = atoi(e) != 0; defer_force = defer_on; if (atoi(e) > 0) defer_mode = std::min(2, atoi(e)); }
        if (const char* e = getenv("MGVIB200_BLUE_MIN")) kBlueMinAxis = std::max<long long>(1, atoll(e));
        if (const char* e = getenv("MGVIB200_BIG_THREADS")) big_block_threads = std::max(32, atoi(e));
        if (const char* e = getenv("MGVIB200_BIG_AP_LGL")) big_ap_lgL = std::min(3, std::max(1, atoi(e)));
        if (const char* e = getenv("MGVIB200_ROW_THREADS")) row_threads = std::max(32, atoi(e));
        if (getenv("MGVIB200_FORCE_GENERIC") && atoi(getenv("MGVIB200_FORCE_GENERIC"))) big = false;
        if (!fast)
            for (int i = 0; i < ndim; ++i)
                if (dims[i] > kNaiveMaxAxis)
                    throw RtError{"FIELD_DHT: axis length must be <= 4096, or a power of two <= 8192 (2-D / 3-D) / <= 2^24 (1-D)", 3};
        A.ensure(sizeof(double) * N);
        data.ensure(sizeof(double) * N);
        rt_h2d(A.p, d.amplitude, sizeof(double) * N, ctx->stream);
        rt_h2d(data.p, d.data, sizeof(double) * N, ctx->stream);
        w.ensure(sizeof(double) * N);
        q.ensure(sizeof(double) * N);
        wbar.ensure(sizeof(double) * N);
        center.ensure(sizeof(double) * N);
        qsum.ensure(sizeof(double) * N);
        // constants of the objective: Distributions.jl logpdf normalisers + prior (src/mgvi_impl.jl:3-7)
        double cl = 0.0;
        if (like == LIKE_POISSON) {
            for (long long i = 0; i < N; ++i) cl += lgamma(d.data[i] + 1.0);
        } else {
            cl = (double)N * (log(sigma) + 0.5 * log(2.0 * M_PI));
        }
        const_term = cl + 0.5 * (double)N * log(2.0 * M_PI);
        std::vector<long long> lens;
        if (big) { lens.push_back(1LL << lg_n1); lens.push_back(1LL << lg_n2); }
        else for (int i = 0; i < ndim; ++i) lens.push_back(dims[i]);
        for (size_t i = 0; i < lens.size(); ++i) {
            const long long n = lens[i];
            if (fast) {
                build_fft_tables(n);
            } else if (n > kBlueMinAxis) {
                build_bluestein_tables(n);
            } else {
                if (cas.count(n)) continue;
                std::vector<double> t((size_t)n * n);
                for (long long k = 0; k < n; ++k)
                    for (long long j = 0; j < n; ++j) {
                        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((j * k) % n) /
                                        (long double)n;
                        t[(size_t)k * n + j] = (double)(cosl(a) + sinl(a));
                    }
                cas[n].ensure(sizeof(double) * n * n);
                rt_h2d(cas[n].p, t.data(), sizeof(double) * n * n, ctx->stream);
                rt_sync(ctx->stream);
            }
        }
        if (big) {
            // W_N^e split as W_N^(e_lo) * W_N^(e_hi 2^h); observations in the permuted data-space order
            const long long nlo = 1LL << lg_h, nhi = N >> lg_h;
            std::vector<double> a(2 * nlo), b(2 * nhi), dp(N);
            const long double tp = 2.0L * 3.14159265358979323846264338327950288L;
            for (long long e = 0; e < nlo; ++e) {
                long double ang = tp * (long double)e / (long double)N;
                a[2 * e] = (double)cosl(ang); a[2 * e + 1] = (double)(-sinl(ang));
            }
            for (long long e = 0; e < nhi; ++e) {
                long double ang = tp * (long double)(e << lg_h) / (long double)N;
                b[2 * e] = (double)cosl(ang); b[2 * e + 1] = (double)(-sinl(ang));
            }
            const long long n1 = 1LL << lg_n1, n2 = 1LL << lg_n2;
            for (long long r = 0; r < n1; ++r)
                for (long long cc = 0; cc < n2; ++cc) dp[r * n2 + cc] = d.data[r + n1 * cc];
            tlo.ensure(16 * nlo); thi.ensure(16 * nhi); data_perm.ensure(8 * N);
            rt_h2d(tlo.p, a.data(), 16 * nlo, ctx->stream);
            rt_h2d(thi.p, b.data(), 16 * nhi, ctx->stream);
            rt_h2d(data_perm.p, dp.data(), 8 * N, ctx->stream);
            rt_sync(ctx->stream);
        }
        rt_sync(ctx->stream);
    }

    // ------------------------------------------------------------------------------------------
    // long 1-D fields (big1d_pass.h)
    // ------------------------------------------------------------------------------------------
    struct BigLaunch { BigParams B; long long grid; int threads; size_t smem; };

    // geometry and operands of one kernel of the four-step path (no launch)
    template <int KIND>
    BigLaunch prep_big(FieldPassParams P, int nvec) {
        BigLaunch L;
        BigParams& B = L.B;
        memset(&B, 0, sizeof B);
        const int npairs = (nvec + 1) / 2;
        cs.ensure(16 * (size_t)N * npairs);
        B.lg_n1 = lg_n1; B.lg_n2 = lg_n2; B.lg_h = lg_h;
        B.cs = cs.as<double2>();
        B.tw1 = tw.at(1LL << lg_n1).as<double2>(); B.tw2 = tw.at(1LL << lg_n2).as<double2>();
        B.tlo = tlo.as<double2>(); B.thi = thi.as<double2>();
        const int lg_len = (KIND == BIG_A || KIND == BIG_AP) ? lg_n1 : lg_n2;
        // lines side by side per unit: A: 4 neighbouring columns; A': 4 neighbouring columns + their 4 mirrors
        // (big_ap_lgL, 1 = one column pair); B kernels: a row and its mirror
        int lgL = (KIND == BIG_A) ? 2 : ((KIND == BIG_AP) ? big_ap_lgL : 1);
        while (lgL > 1 && (lg_len - 3 + lgL > 10 || ((size_t)16 << (lgL + lg_len)) > 200 * 1024 ||
                           (KIND == BIG_AP && (1LL << lg_n2) < (4LL << lgL)))) --lgL;   // <= 1024 threads per unit
        if (KIND != BIG_A && lg_len - 3 + lgL > 10) throw RtError{"1-D field too long for the four-step path", 3};
        B.lgL = lgL;
        const long long upt = 1LL << (lg_len - 3 + lgL);
        long long G = std::max<long long>(1, big_block_threads / upt);
        B.units_per_pair = (KIND == BIG_A) ? ((1LL << lg_n2) >> lgL) : ((KIND == BIG_AP) ? ((1LL << lg_n2) >> lgL) : ((1LL << lg_n1) >> 1));
        G = std::min(G, B.units_per_pair);
        G = std::max<long long>(G, (32 + upt - 1) / upt);
        B.G = (int)G;
        B.npairs = npairs;
        const int npgrp = npairs;
        B.nloop = 1;
        const long long ubl = (B.units_per_pair + G - 1) / G;
        L.grid = ubl * npgrp;
        P.N = N; P.nvec = nvec;
        if (P.fin != FIN_NONE) {
            const size_t np = (P.fin == FIN_TOTAL) ? (size_t)L.grid : (size_t)(nvec + 1) * (size_t)ubl;
            reduce_ws_ensure(rws, ctx, np, (size_t)nvec + 1);
            P.partials = rws.partials.as<double>(); P.counters = rws.counters.as<unsigned>();
        }
        B.f = P;
        L.threads = (int)(G * upt);
        L.smem = kSmemHeader + (size_t)G * ((size_t)16 << (lgL + lg_len));
        return L;
    }

    template <int KIND, int PRO, int EPI>
    void launch_big(FieldPassParams P, int nvec) {
        const BigLaunch L = prep_big<KIND>(P, nvec);
        ctx->prof_pre(5000 + KIND * 100 + PRO * 10 + EPI);
        if (L.threads <= 256) rt_launch_pdl(k_big1d<KIND, PRO, EPI, 256>, L.grid, L.threads, L.smem, ctx->stream, L.B);
        else rt_launch_pdl(k_big1d<KIND, PRO, EPI, 1024>, L.grid, L.threads, L.smem, ctx->stream, L.B);
        ctx->prof_post();
        ctx->launches++;
    }

    // ------------------------------------------------------------------------------------------
    // geometry of one axis pass
    // ------------------------------------------------------------------------------------------
    PassGeo geo_axis(int axis, int nvec) const {
        PassGeo g;
        g.n = (int)dims[axis];
        g.log2n = ilog2(g.n);
        const int tpl = g.n / 8;
        g.inner = 1;
        for (int i = axis + 1; i < ndim; ++i) g.inner *= dims[i];
        if (ndim == 1) {
            g.geo = GEO_PAIRVEC; g.L = 1;
            g.units_per_vec = 1;
            const long long npairs = (nvec + 1) / 2;
            long long G = std::max<long long>(1, 256 / tpl);
            long long np2 = 1; while (np2 < npairs) np2 <<= 1;
            G = std::min(G, np2);
            G = std::max<long long>(G, (32 + tpl - 1) / tpl);
            g.G = (int)G;
            g.ntb = (npairs + G - 1) / G;      // blocks (with nloop = 1)
        } else if (axis == ndim - 1) {
            g.geo = GEO_CONTIG; g.L = 1;
            g.units_per_vec = (N / g.n) / 2;
            long long G = std::max<long long>(1, row_threads / tpl);
            G = std::min(G, g.units_per_vec);
            G = std::max<long long>(G, (32 + tpl - 1) / tpl);
            g.G = (int)G;
            g.ntb = (g.units_per_vec + G - 1) / G;
        } else {
            g.geo = GEO_STRIDED;
            // tile width by axis length: wide enough that a unit is 256..512 threads, at most 8 complex
            // columns (a full 128-byte line per row); 2048-point axes: 2 columns (measured best at C3),
            // 256-point axes: 8 columns (C5)
            long long Lw = max_L > 0 ? max_L : std::min<long long>(8, std::max<long long>(2, 4096 / g.n));
            long long L = std::min<long long>(Lw, g.inner / 2);
            while (L * tpl > 1024) L >>= 1;
            while (kSmemHeader + L * g.n * 16 > 220 * 1024) L >>= 1;
            if (L < 1) L = 1;
            g.L = (int)L;
            g.units_per_vec = (N / g.n) / (2 * L);
            const long long upt = L * tpl;
            long long G = std::max<long long>(1, 256 / upt);
            G = std::min(G, g.units_per_vec);
            G = std::max<long long>(G, (32 + upt - 1) / upt);
            g.G = (int)G;
            g.ntb = (g.units_per_vec + G - 1) / G;
        }
        g.threads = g.G * g.L * tpl;
        g.smem = (size_t)field_pass_smem_bytes(g.n, g.L, g.G);
        g.lgL = ilog2(g.L);
        g.lg_tpo = (g.geo == GEO_STRIDED) ? ilog2(g.inner / (2 * g.L)) : 0;
        return g;
    }

    // run-time-shape build
    template <int GEO, int PRO, int EPI, bool DBL>
    void launch_lg(const FieldPassParams& P, const PassGeo& g, long long grid) {
        if (g.threads <= 256)
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 256, 2, 0, -1, true>, grid, g.threads, g.smem, ctx->stream, P);
        else if (g.threads <= 512)
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 512, 1, 0, -1, true>, grid, g.threads, g.smem, ctx->stream, P);
        else
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 1024, 1, 0, -1, true>, grid, g.threads, g.smem, ctx->stream, P);
    }

    // compile-time-shape build; the block size follows from (LG, LGL): 2^(LG - 3 + LGL) threads per unit
    template <int GEO, int PRO, int EPI, bool DBL, int LG, int LGL, int PF>
    void launch_hot_pf(const FieldPassParams& P, const PassGeo& g, long long grid) {
        if (g.threads <= 256)
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 256, MGVI_MINB, LG, LGL, false, PF>, grid, g.threads, g.smem, ctx->stream, P);
        else if (g.threads <= 512)
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 512, 2, LG, LGL, false, PF>, grid, g.threads, g.smem, ctx->stream, P);
        else
            rt_launch_pdl(k_field_pass<GEO, PRO, EPI, DBL, 1024, 1, LG, LGL, false, PF>, grid, g.threads, g.smem, ctx->stream, P);
    }

    template <int GEO, int PRO, int EPI, bool DBL, int LG, int LGL>
    void launch_hot(const FieldPassParams& P0, const PassGeo& g0, long long grid) {
        FieldPassParams P = P0;
        PassGeo g = g0;
#ifdef MGVI_STAGE_VIN
        if (GEO == GEO_CONTIG && EPI == EPI_METRIC && stage_vin && P.vin != P.out &&
            field_pass_smem_bytes_staged(g.n, g.L, g.G) <= 72 * 1024) {
            P.stage_vin = 1;
            g.smem = (size_t)field_pass_smem_bytes_staged(g.n, g.L, g.G);
        }
#endif
        // transform scheme of this kernel (field_pass.h): the paired schemes drop one shared-memory exchange per
        // transform; first pass of a CG iteration = 1, last pass = 2, fused axis-0 pass = 3
        constexpr int PF = (LG % 3 != 2 || LG < 8) ? 0
                           : (GEO == GEO_CONTIG && PRO == PRO_CG && EPI == EPI_STORE) ? 1
                           : (GEO == GEO_CONTIG && PRO == PRO_LOAD && EPI == EPI_METRIC) ? 2
                           : (GEO == GEO_STRIDED && DBL) ? 3 : 0;
        if constexpr (PF != 0) {
            if (use_paired && !P.stage_vin) {
                P.tw2 = twp.at(g.n).as<double2>();
                launch_hot_pf<GEO, PRO, EPI, DBL, LG, LGL, PF>(P, g, grid);
                return;
            }
        }
        launch_hot_pf<GEO, PRO, EPI, DBL, LG, LGL, 0>(P, g, grid);
    }

    template <int GEO, int PRO, int EPI, bool DBL>
    void launch_pass(FieldPassParams P, const PassGeo& g, int nvec, int nsgrp, int nloop) {
        if (g.geo != GEO) throw RtError{"internal: geometry class mismatch", 1};
        P.geo = g.geo; P.n = g.n; P.log2n = g.log2n; P.L = g.L; P.G = g.G; P.lgL = g.lgL; P.lg_tpo = g.lg_tpo;
        P.N = N; P.inner = g.inner; P.units_per_vec = g.units_per_vec; P.ntb = g.ntb;
        P.nvec = nvec; P.nsgrp = nsgrp; P.nloop = nloop;
        P.tw = tw.at(g.n).as<double2>();
        long long grid;
        if (g.geo == GEO_PAIRVEC) {
            const long long npairs = (nvec + 1) / 2;
            grid = (npairs + (long long)g.G * nloop - 1) / ((long long)g.G * nloop);
        } else {
            grid = g.ntb * nsgrp;
        }
        P.prefetch_epi = prefetch_epi;
        // prefetch distance: what one wave of resident blocks covers (blocks/SM from the thread count)
        {
            const int per_sm = g.threads <= 256 ? MGVI_MINB : (g.threads <= 512 ? 2 : 1);
            P.prefetch_blocks = (grid > 4LL * ctx->sm_count * per_sm) ? prefetch_waves * ctx->sm_count * per_sm : 0;
        }
        if (P.fin != FIN_NONE) {
            const size_t np = (P.fin == FIN_TOTAL) ? (size_t)grid : (size_t)nvec * (size_t)g.ntb;
            reduce_ws_ensure(rws, ctx, np, (size_t)std::max(nvec, 1));
            P.partials = rws.partials.as<double>();
            P.counters = rws.counters.as<unsigned>();
        }
        ctx->prof_pre(PRO * 100 + EPI * 10 + (DBL ? 1 : 0));
        // The kernels of the CG iteration get fully specialised builds for the benchmark shapes
        // (2048-point axes: C3; 256-point axes: C5): compile-time length, tile width and no vector loop.
        constexpr bool hot_c = (GEO == GEO_CONTIG) && !DBL &&
                               ((PRO == PRO_CG && EPI == EPI_STORE) || (PRO == PRO_LOAD && EPI == EPI_METRIC));
        constexpr bool hot_s = (GEO == GEO_STRIDED) && (PRO == PRO_LOAD && EPI == EPI_STORE);
        bool done = false;
        // `if constexpr`: only the shape classes that have a specialised build instantiate one
        if (nloop == 1 && use_hot) {
            if constexpr (hot_c) {
                if (g.log2n == 11) { launch_hot<GEO, PRO, EPI, DBL, 11, 0>(P, g, grid); done = true; }
                else if (g.log2n == 8) { launch_hot<GEO, PRO, EPI, DBL, 8, 0>(P, g, grid); done = true; }
            }
            if constexpr (hot_s && DBL) {
                if (g.log2n == 11 && g.lgL == 2) { launch_hot<GEO, PRO, EPI, DBL, 11, 2>(P, g, grid); done = true; }
                else if (g.log2n == 11 && g.lgL == 1) { launch_hot<GEO, PRO, EPI, DBL, 11, 1>(P, g, grid); done = true; }
            }
            if constexpr (hot_s) {
                if (!done && g.log2n == 8 && g.lgL == 3) { launch_hot<GEO, PRO, EPI, DBL, 8, 3>(P, g, grid); done = true; }
                else if (!done && g.log2n == 8 && g.lgL == 2) { launch_hot<GEO, PRO, EPI, DBL, 8, 2>(P, g, grid); done = true; }
            }
        }
        if (!done) launch_lg<GEO, PRO, EPI, DBL>(P, g, grid);
        ctx->prof_post();
        ctx->launches++;
    }

    // ------------------------------------------------------------------------------------------
    // generic (any length) chain pieces
    // ------------------------------------------------------------------------------------------
    void naive_transform_all(double*& cur, double*& other, int nvec, const double* w_after, const int* active = nullptr) {
        for (int a = 0; a < ndim; ++a) {
            if (dims[a] > kBlueMinAxis) {
                BlueParams B;
                memset(&B, 0, sizeof B);
                const long long n = dims[a];
                B.in = cur; B.out = other; B.w = (a == ndim - 1) ? w_after : nullptr;
                B.chirp = blue_chirp.at(n).as<double2>(); B.hf = blue_hf.at(n).as<double2>();
                B.lg_m = blue_lg_m(n);
                B.tw = tw.at(1LL << B.lg_m).as<double2>();
                B.N = N; B.n = (int)n; B.nlines = (long long)nvec * (N / n);
                B.lines_per_vec = N / n; B.active = active;
                B.inner = 1;
                for (int i = a + 1; i < ndim; ++i) B.inner *= dims[i];
                const int threads = 1 << (B.lg_m - 3);
                const size_t smem = kSmemHeader + ((size_t)16 << B.lg_m);
                const long long grid = (B.nlines + 1) / 2;
                ctx->prof_pre(6000 + a);
                // one build per padded length m = 2^lg_m (256 ... 8192)
                switch (B.lg_m) {
                    case 8: rt_launch(k_bluestein_axis<256, 8>, grid, threads, smem, ctx->stream, B); break;
                    case 9: rt_launch(k_bluestein_axis<256, 9>, grid, threads, smem, ctx->stream, B); break;
                    case 10: rt_launch(k_bluestein_axis<256, 10>, grid, threads, smem, ctx->stream, B); break;
                    case 11: rt_launch(k_bluestein_axis<256, 11>, grid, threads, smem, ctx->stream, B); break;
                    case 12: rt_launch(k_bluestein_axis<512, 12>, grid, threads, smem, ctx->stream, B); break;
                    case 13: rt_launch(k_bluestein_axis<1024, 13>, grid, threads, smem, ctx->stream, B); break;
                    default: throw RtError{"internal: Bluestein length out of range", 1};
                }
                ctx->prof_post();
                ctx->launches++;
                std::swap(cur, other);
                continue;
            }
            NaiveParams Q;
            Q.in = cur; Q.out = other; Q.cas = cas.at(dims[a]).as<double>();
            Q.w = (a == ndim - 1) ? w_after : nullptr;
            Q.N = N; Q.n = (int)dims[a]; Q.total = (long long)nvec * N;
            Q.inner = 1;
            for (int i = a + 1; i < ndim; ++i) Q.inner *= dims[i];
            rt_launch(k_naive_axis, (Q.total + 255) / 256, 256, 0, ctx->stream, Q);
            ctx->launches++;
            std::swap(cur, other);
        }
    }

    template <int PRO, int EPI>
    void chain_generic(FieldPassParams P, int nvec, bool metric, int fin_pro, double* red_pro, int fin_epi,
                       double* red_epi) {
        tmpA.ensure(sizeof(double) * N * nvec);
        tmpB.ensure(sizeof(double) * N * nvec);
        double* cur = tmpA.as<double>();
        double* other = tmpB.as<double>();
        const long long nchunk = (N + 255) / 256;
        P.N = N; P.nvec = nvec;
        {
            ElemParams E;
            E.f = P; E.tmp = cur; E.total = (long long)nvec * N;
            E.f.fin = fin_pro; E.f.red_out = red_pro;
            const long long grid = nchunk * nvec;
            if (fin_pro != FIN_NONE) {
                reduce_ws_ensure(rws, ctx, (size_t)grid, (size_t)std::max(nvec, 1));
                E.f.partials = rws.partials.as<double>(); E.f.counters = rws.counters.as<unsigned>();
            }
            rt_launch(k_elem_pro<PRO>, grid, 256, 512, ctx->stream, E);
            ctx->launches++;
        }
        const int* act = P.mask_active ? P.cg.active : nullptr;
        naive_transform_all(cur, other, nvec, metric ? P.w : nullptr, act);
        if (metric) naive_transform_all(cur, other, nvec, nullptr, act);
        {
            ElemParams E;
            E.f = P; E.tmp = cur; E.total = (long long)nvec * N;
            E.f.fin = fin_epi; E.f.red_out = red_epi;
            const long long grid = nchunk * nvec;
            if (fin_epi != FIN_NONE) {
                reduce_ws_ensure(rws, ctx, (size_t)grid, (size_t)std::max(nvec, 1));
                E.f.partials = rws.partials.as<double>(); E.f.counters = rws.counters.as<unsigned>();
            }
            rt_launch(k_elem_epi<EPI>, grid, 256, 512, ctx->stream, E);
            ctx->launches++;
        }
    }

    // ------------------------------------------------------------------------------------------
    // chains.  `scratch` is [nvec][N].  Per-vector reductions: fin_pro over the prologue terms (first
    // pass), fin_epi over the epilogue terms (last pass); single-pass geometries put both into the
    // fin_epi slots.
    // ------------------------------------------------------------------------------------------
    template <int PRO, int EPI>
    void chain_forward(FieldPassParams P, int nvec, double* scratch, int fin_pro, double* red_pro, int fin_epi,
                       double* red_epi) {
        P.scale = c * pow(0.5, ndim);
        if (!fast) {
            chain_generic<PRO, EPI>(P, nvec, false, fin_pro, red_pro, fin_epi, red_epi);
            return;
        }
        if (big) {
            FieldPassParams Q = P;
            Q.fin = fin_pro; Q.red_out = red_pro;
            launch_big<BIG_A, PRO, EPI_STORE>(Q, nvec);
            FieldPassParams E = P;
            E.fin = fin_epi; E.red_out = red_epi; E.data = data_perm.as<double>();
            launch_big<BIG_B_END, PRO_LOAD, EPI>(E, nvec);
            return;
        }
        if (ndim == 1) {
            P.fin = (fin_epi != FIN_NONE) ? fin_epi : fin_pro;
            P.red_out = (fin_epi != FIN_NONE) ? red_epi : red_pro;
            launch_pass<GEO_PAIRVEC, PRO, EPI, false>(P, geo_axis(0, nvec), nvec, 1, 1);
            return;
        }
        // contiguous axis first
        {
            FieldPassParams Q = P;
            Q.out = scratch; Q.out_ld = N; Q.fin = fin_pro; Q.red_out = red_pro;
            launch_pass<GEO_CONTIG, PRO, EPI_STORE, false>(Q, geo_axis(ndim - 1, nvec), nvec, nvec, 1);
        }
        for (int a = ndim - 2; a >= 1; --a) {
            FieldPassParams Q = P;
            Q.in = scratch; Q.in_ld = N; Q.out = scratch; Q.out_ld = N; Q.fin = FIN_NONE;
            launch_pass<GEO_STRIDED, PRO_LOAD, EPI_STORE, false>(Q, geo_axis(a, nvec), nvec, nvec, 1);
        }
        {
            FieldPassParams Q = P;
            Q.in = scratch; Q.in_ld = N; Q.fin = fin_epi; Q.red_out = red_epi;
            launch_pass<GEO_STRIDED, PRO_LOAD, EPI, false>(Q, geo_axis(0, nvec), nvec, nvec, 1);
        }
    }

    template <int PRO, int EPI>
    void chain_adjoint(FieldPassParams P, int nvec, double* scratch, int fin_epi, double* red_epi) {
        if (!fast) {
            chain_generic<PRO, EPI>(P, nvec, false, FIN_NONE, nullptr, fin_epi, red_epi);
            return;
        }
        if (big) {
            FieldPassParams Q = P;
            Q.fin = FIN_NONE;
            launch_big<BIG_B_START, PRO, EPI_STORE>(Q, nvec);
            FieldPassParams E = P;
            E.fin = fin_epi; E.red_out = red_epi;
            launch_big<BIG_AP, PRO_LOAD, EPI>(E, nvec);
            return;
        }
        if (ndim == 1) {
            P.fin = fin_epi; P.red_out = red_epi;
            launch_pass<GEO_PAIRVEC, PRO, EPI, false>(P, geo_axis(0, nvec), nvec, 1, 1);
            return;
        }
        {
            FieldPassParams Q = P;
            Q.out = scratch; Q.out_ld = N; Q.fin = FIN_NONE;
            launch_pass<GEO_STRIDED, PRO, EPI_STORE, false>(Q, geo_axis(0, nvec), nvec, nvec, 1);
        }
        for (int a = 1; a <= ndim - 2; ++a) {
            FieldPassParams Q = P;
            Q.in = scratch; Q.in_ld = N; Q.out = scratch; Q.out_ld = N; Q.fin = FIN_NONE;
            launch_pass<GEO_STRIDED, PRO_LOAD, EPI_STORE, false>(Q, geo_axis(a, nvec), nvec, nvec, 1);
        }
        {
            FieldPassParams Q = P;
            Q.in = scratch; Q.in_ld = N; Q.fin = fin_epi; Q.red_out = red_epi;
            launch_pass<GEO_CONTIG, PRO_LOAD, EPI, false>(Q, geo_axis(ndim - 1, nvec), nvec, nvec, 1);
        }
    }

    // Tile-major layout of the 2-D metric product (field_pass.h): decides whether it applies and keeps the re-laid
    // copy of the weight `wsrc` current (one VOP_TILE launch per linearisation, on the current stream).
    bool tiled_weight(const double* wsrc, int nvec, int& lg_wt, int& lg_rows) {
        lg_wt = 0; lg_rows = 0;
        if (!(ndim == 2 && tiled_mode != 0)) return false;
        const PassGeo g0 = geo_axis(0, nvec);
        const long long wt = 2LL * g0.L;
        if (!(g0.geo == GEO_STRIDED && (tiled_mode == 1 || wt < 16) && (dims[1] / 8) % wt == 0 && dims[1] >= 8 * wt)) return false;
        lg_wt = ilog2((int)wt); lg_rows = g0.log2n;
        if (w_tiled_src != wsrc || w_tiled_gen != w_gen || w_tiled_lg_wt != lg_wt) {
            w_tiled.ensure(sizeof(double) * N);
            VecParams T;
            memset(&T, 0, sizeof T);
            T.op = VOP_TILE; T.nvec = 1; T.N = N; T.in = wsrc; T.out = w_tiled.as<double>();
            T.lg_wt = lg_wt; T.lg_rows = lg_rows; T.lg_cols = ilog2((int)dims[1]);
            vec_launch(ctx, rws, T);
            w_tiled_src = wsrc; w_tiled_gen = w_gen; w_tiled_lg_wt = lg_wt;
        }
        return true;
    }

    // M v = v + c^2 A . H( w . H(A . v) ), 2d-1 passes.  P carries PRO operands, vin, out, w.
    template <int PRO>
    void chain_metric(FieldPassParams P, int nvec, double* scratch, int fin, double* red) {
        P.scale = c * c * pow(0.5, 2 * ndim);
        if (!fast) {
            chain_generic<PRO, EPI_METRIC>(P, nvec, true, FIN_NONE, nullptr, fin, red);
            return;
        }
        if (big) {
            FieldPassParams Q = P;
            Q.fin = FIN_NONE;
            launch_big<BIG_A, PRO, EPI_STORE>(Q, nvec);
            launch_big<BIG_B_MID, PRO_LOAD, EPI_STORE>(Q, nvec);
            FieldPassParams E = P;
            E.fin = fin; E.red_out = red;
            launch_big<BIG_AP, PRO_LOAD, EPI_METRIC>(E, nvec);
            return;
        }
        if (ndim == 1) {
            P.fin = fin; P.red_out = red;
            launch_pass<GEO_PAIRVEC, PRO, EPI_METRIC, true>(P, geo_axis(0, nvec), nvec, 1, 1);
            return;
        }
        // 2-D with narrow column tiles (rows of 2L reals < one 128-byte line): keep the scratch and the
        // weight tile-major so the column pass reads and writes whole lines (field_pass.h)
        int lg_wt = 0, lg_rows = 0;
        if (tiled_weight(P.w, nvec, lg_wt, lg_rows)) {
            // the CG chains write A p over their scratch (same rows in, same rows out); with the
            // scratch tile-major the last pass would overwrite tiles other blocks still have to
            // read, so the scratch becomes a buffer of its own
            if (scratch == P.out) {
                ws_tile.ensure(sizeof(double) * N * nvec);
                scratch = ws_tile.as<double>();
            }
        }
        {
            FieldPassParams Q = P;
            Q.out = scratch; Q.out_ld = N; Q.fin = FIN_NONE;
            Q.t_lg_wt = lg_wt; Q.t_lg_rows = lg_rows;
            launch_pass<GEO_CONTIG, PRO, EPI_STORE, false>(Q, geo_axis(ndim - 1, nvec), nvec, nvec, 1);
        }
        FieldPassParams S = P;
        S.in = scratch; S.in_ld = N; S.out = scratch; S.out_ld = N; S.fin = FIN_NONE;
        if (lg_wt) { S.w = w_tiled.as<double>(); S.t_lg_wt = lg_wt; S.t_lg_rows = lg_rows; }
        for (int a = ndim - 2; a >= 1; --a) launch_pass<GEO_STRIDED, PRO_LOAD, EPI_STORE, false>(S, geo_axis(a, nvec), nvec, nvec, 1);
        launch_pass<GEO_STRIDED, PRO_LOAD, EPI_STORE, true>(S, geo_axis(0, nvec), nvec, nvec, 1);
        for (int a = 1; a <= ndim - 2; ++a) launch_pass<GEO_STRIDED, PRO_LOAD, EPI_STORE, false>(S, geo_axis(a, nvec), nvec, nvec, 1);
        {
            FieldPassParams Q = P;
            Q.in = scratch; Q.in_ld = N; Q.fin = fin; Q.red_out = red;
            Q.t_lg_wt = lg_wt; Q.t_lg_rows = lg_rows;
            launch_pass<GEO_CONTIG, PRO_LOAD, EPI_METRIC, false>(Q, geo_axis(ndim - 1, nvec), nvec, nvec, 1);
        }
    }

    // 1-D fields pack TWO residual samples into one complex transform (fft_core.h), so a sample's rounding
    // depends on its partner: with an odd number of residuals per rank the partners would depend on the
    // sharding.  Refused rather than silently rank-count dependent.
    void check_pairing(int64_t n_local) const {
        if (ctx->world > 1 && ndim == 1 && fast && (n_local & 1))
            throw RtError{"multi-GPU, 1-D field: every rank needs an even number of residual samples", 1};
    }

    FieldPassParams base_params() const {
        FieldPassParams P;
        memset(&P, 0, sizeof P);
        P.N = N;
        P.A = A.as<double>();
        P.data = data.as<double>();
        P.offset = offset; P.sigma = sigma; P.inv_sigma2 = inv_sigma2;
        P.link = link; P.likelihood = like;
        P.fin = FIN_NONE;
        return P;
    }

    // ------------------------------------------------------------------------------------------
    // Engine interface
    // ------------------------------------------------------------------------------------------
    void linearize(const double* x) override {
        // a4: w = F g'^2 and q = g' sqrt(F) at lambda(center)  (src/residual_samplers.jl:4-8)
        FieldPassParams P = base_params();
        P.x = x; P.point_mode = POINT_CENTER;
        P.w_out = w.as<double>(); P.q_out = q.as<double>();
        pt_t.ensure(sizeof(double) * N);
        chain_forward<PRO_POINT, EPI_LINEARIZE>(P, 1, pt_t.as<double>(), FIN_NONE, nullptr, FIN_NONE, nullptr);
        rt_d2d(center.p, x, sizeof(double) * N, ctx->stream);
        ++w_gen;
    }

    void get_weights(double* w_host, double* q_host) override {
        if (w_host) rt_d2h(w_host, w.p, sizeof(double) * N, ctx->stream);
        if (q_host) rt_d2h(q_host, q.p, sizeof(double) * N, ctx->stream);
        rt_sync(ctx->stream);
        if (big)
            for (double* arr : {w_host, q_host})
                if (arr) unpermute_big(arr);
    }

    void metric_apply(const double* V, int64_t k, double* MV) override {
        FieldPassParams P = base_params();
        P.in = V; P.in_ld = N; P.vin = V; P.vin_ld = N; P.out = MV; P.out_ld = N; P.w = w.as<double>();
        ws_t.ensure(sizeof(double) * N * k);
        chain_metric<PRO_AMP>(P, (int)k, ws_t.as<double>(), FIN_NONE, nullptr);
    }

    void sample_cg(const double* Z1, const double* Z2, int64_t n, const mgvib200_cg_opts& o, double* R,
                   int64_t* iters_host, double* rnorm_host) override {
        check_pairing(n);
        const int nvec = (int)n;
        if (use_groups(nvec)) { sample_cg_grouped(Z1, Z2, nvec, o, R, iters_host, rnorm_host); return; }
        const size_t vb = sizeof(double) * N * nvec;
        ws_r.ensure(vb); ws_p.ensure(vb); ws_t.ensure(vb);
        cgws.ensure(nvec);
        const long long itmax = o.maxiters > 0 ? o.maxiters : N;
        // deferred solution update: 2-D / 3-D fast path only (the fused axis-0 pass carries it)
        // Measured (profiles/README.md, round 2): C3 (2048-point axis 0) 3.50 -> 3.36 ms per iteration; C5 (256-point
        // axis 0, a pass that is already close to its HBM time) 4.30 -> 4.34 ms: on by default for 2-D fields with a long
        // axis 0 only (MGVIB200_DEFER_X=1/2 forces it on wherever it applies, =0 off)
        const bool defer = (defer_force || (defer_on && ndim == 2 && dims[0] >= 1024)) && fast && !big && ndim >= 2 && (N & 1) == 0;
        cgws.defer = defer;
        if (defer) {
            ws_p2.ensure(vb);
            rt_memset(cgws.x_pending.p, 0, sizeof(int) * nvec, ctx->stream);
            rt_memset(cgws.x_count.p, 0, sizeof(int) * nvec, ctx->stream);
        }
        CgScalars sc = cgws.scalars(o.abstol, o.reltol, itmax, CG_KRYLOV);
        rt_memset(cgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        rt_event_record(ctx->ev0, ctx->stream);
        // b = J'(sqrt(F) . n1) + eta ; r = p = b ; x = 0 ; gamma = b.b   (src/residual_samplers.jl:75-77)
        {
            FieldPassParams P = base_params();
            P.q = q.as<double>(); P.Z1 = Z1; P.Z1_ld = z1_ld; P.Z1_off = 0; P.Z1_stride = z1_stride;
            P.Z2 = Z2; P.Z2_ld = N;
            P.r = ws_r.as<double>(); P.p = ws_p.as<double>(); P.xout = R;
            P.scale = c * pow(0.5, ndim);
            P.cg = sc;
            chain_adjoint<PRO_RHS, EPI_RHS>(P, nvec, ws_t.as<double>(), FIN_RHS, nullptr);
        }
        // lock-step CG over all residual samples (Krylov.jl cg!).  Every per-iteration scalar
        // (alpha, beta, gamma, active masks, iteration counts, the itmax stop) lives in device memory, so an
        // iteration is a FIXED launch sequence: after the first one (p = r, no update of p) a chunk of
        // kCgChunk iterations is captured into a CUDA graph once and replayed.  The host only polls the
        // number of active samples, one chunk behind the launches, so the stream never drains; iterations
        // launched past convergence are no-ops (every block of an inactive sample exits at once).
        int* h = ctx->h_pinned;
        h[0] = 1;
        rt_d2h(h, cgws.n_active.p, sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        // `k` = index of the lock-step iteration.  Deferred mode: p_k lives in buffer k % 2 (the first pass reads
        // p_(k-1) from the other one), the fused axis-0 pass applies x += alpha_(k-1) p_(k-1), the update kernel
        // only touches r
        double* pbuf[2] = {ws_p.as<double>(), defer ? ws_p2.as<double>() : ws_p.as<double>()};
        auto launch_iter = [&](bool first, long long k) {
            FieldPassParams P = base_params();
            double* p_cur = pbuf[k & 1];
            double* p_prev = pbuf[(k & 1) ^ 1];
            P.r = ws_r.as<double>();
            P.p = first ? p_cur : p_prev;
            P.p_out = defer ? p_cur : nullptr;
            P.vin = p_cur; P.vin_ld = N;
            P.out = ws_t.as<double>(); P.out_ld = N;
            P.w = w.as<double>();
            P.first_iter = first ? 1 : 0;
            P.cg = sc;
            P.mask_active = 1;
            if (defer) { P.defer_x = defer_mode; P.x_prev_p = p_prev; P.xout = R; }
            chain_metric<PRO_CG>(P, nvec, ws_t.as<double>(), FIN_CG_PAP, nullptr);
            if (defer) vec_cg_update_r(ctx, rws, sc, ws_r.as<double>(), ws_t.as<double>(), N, nvec);
            else vec_cg_update(ctx, rws, sc, R, ws_r.as<double>(), ws_p.as<double>(), ws_t.as<double>(), N, nvec);
        };
        long long it = 0;
        if (h[0] > 0 && itmax > 0) {
            launch_iter(true, 0);
            it = 1;
            const bool use_graph = rt_has_graphs && graphs_on && !ctx->profile && it < itmax;
            if (use_graph) {
                // (the first, host-launched iteration above has already sized every buffer for this nvec)
                const CgGraphKey key{R, ws_r.p, ws_p.p, defer ? ws_p2.p : nullptr, ws_t.p, ws_tile.p, w_tiled.p, tmpA.p, tmpB.p, cs.p,
                                     rws.partials.p, rws.counters.p, nvec, o.abstol, o.reltol, itmax};
                if (!cg_graph_valid || !(key == cg_graph_key)) {
                    if (cg_graph_valid) { rt_graph_destroy(cg_graph); cg_graph_valid = false; }
                    const int64_t l0 = ctx->launches;
                    rt_capture_begin(ctx->stream);
                    try {
                        for (int k = 0; k < kCgChunk; ++k) launch_iter(false, 1 + k);     // replays start at odd k: same buffer parity
                        cg_graph = rt_capture_end(ctx->stream);
                    } catch (...) {
                        rt_capture_abort(ctx->stream);
                        throw;
                    }
                    cg_graph_launches = ctx->launches - l0;
                    ctx->launches = l0;
                    cg_graph_key = key;
                    cg_graph_valid = true;
                }
            }
            if (it >= itmax) {
                // a single iteration was asked for
            } else if (!use_graph) {
                // host-launched iterations (emulator build, per-kernel profiling, MGVIB200_GRAPH=0): poll after
                // 1, 2, 4, 8, 8, ... iterations
                int poll = 1;
                while (h[0] > 0 && it < itmax) {
                    for (int k = 0; k < poll && it < itmax; ++k, ++it) launch_iter(false, it);
                    rt_d2h(h, cgws.n_active.p, sizeof(int), ctx->stream);
                    rt_sync(ctx->stream);
                    if (poll < kCgChunk) poll *= 2;
                }
            } else {
                int slot = 0;
                bool pending = false;
                const long long it_cap = itmax + 2 * kCgChunk;      // the device stops every sample at itmax by itself
                while (true) {
                    rt_graph_launch(cg_graph, ctx->stream);
                    ctx->launches += cg_graph_launches;
                    it += kCgChunk;
                    rt_d2h(h + 16 * (1 + slot), cgws.n_active.p, sizeof(int), ctx->stream);
                    rt_event_record(ctx->ev_poll[slot], ctx->stream);
                    if (pending) {
                        rt_event_sync(ctx->ev_poll[slot ^ 1]);
                        if (h[16 * (1 + (slot ^ 1))] <= 0) break;
                    }
                    if (it >= it_cap) break;
                    pending = true;
                    slot ^= 1;
                }
            }
        }
        // deferred mode: vectors that stopped in the very last launched update still lack its alpha p
        // (`it` iterations were launched; the last one built p in buffer (it - 1) % 2)
        if (defer && it > 0) vec_x_flush(ctx, rws, sc, R, pbuf[(it - 1) & 1], N, nvec);
        rt_event_record(ctx->ev1, ctx->stream);
        std::vector<int> hi(nvec);
        std::vector<double> hr(nvec);
        rt_d2h(hi.data(), cgws.iters.p, sizeof(int) * nvec, ctx->stream);
        rt_d2h(hr.data(), cgws.rnorm.p, sizeof(double) * nvec, ctx->stream);
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

    bool use_groups(int nvec) const {
        if (!rt_has_graphs || !graphs_on || ctx->profile || !fast || nvec < 2 || groups_mode == 1) return false;
        if (ndim == 1) {
            // 1-D fields pair two samples per complex transform: split at a pair boundary (same partners, same bits),
            // at least one pair per group.  Measured at C2 (16 samples): 0.0481 ms per iteration in one group, 0.0487 ms
            // in two -- no gain, so only on request
            return nvec >= 4 && groups_mode == 2;
        }
        if (groups_mode == 2) return true;
        // auto: the row pass of the whole batch is fewer than ~100 waves of resident blocks (measured at C3: 4 samples
        // 0.442 -> 0.416 ms per iteration, 8 samples 0.838 -> 0.809, 32 samples 3.382 -> 3.355; profiles/README.md)
        const PassGeo g = geo_axis(ndim - 1, nvec);
        return (double)g.ntb * nvec < 100.0 * MGVI_MINB * ctx->sm_count;
    }

    // The sampler's batched CG as two independent lock-step solves (samples [0, nA) and [nA, nvec)) on two streams,
    // issued alternately by this thread: same kernels, same per-sample arithmetic and reduction order (results are
    // bit-for-bit those of the single-stream path), the blocks of one group fill the ramp-up and tail of the other's
    // kernels.  Each group has its own stream, reduction workspace, tile scratch, active counter and CUDA graph.
    void sample_cg_grouped(const double* Z1, const double* Z2, int nvec, const mgvib200_cg_opts& o, double* R,
                           int64_t* iters_host, double* rnorm_host) {
        const size_t vb = sizeof(double) * N * nvec;
        ws_r.ensure(vb); ws_p.ensure(vb); ws_t.ensure(vb);
        cgws.ensure(nvec);
        const long long itmax = o.maxiters > 0 ? o.maxiters : N;
        const bool defer = (defer_force || (defer_on && ndim == 2 && dims[0] >= 1024)) && !big && ndim >= 2 && (N & 1) == 0;
        cgws.defer = defer;
        if (defer) {
            ws_p2.ensure(vb);
            rt_memset(cgws.x_pending.p, 0, sizeof(int) * nvec, ctx->stream);
            rt_memset(cgws.x_count.p, 0, sizeof(int) * nvec, ctx->stream);
        }
        rt_memset(cgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        if (!groups_ready) {
            for (auto& G : grp) { G.st = rt_stream_create(); G.ev[0] = rt_event_create(); G.ev[1] = rt_event_create(); G.ev_done = rt_event_create(); }
            groups_ready = true;
        }
        { int a, b; tiled_weight(w.as<double>(), nvec, a, b); }      // the shared re-laid weight, before the streams fork
        rt_event_record(ctx->ev0, ctx->stream);
        const CgScalars sc_all = cgws.scalars(o.abstol, o.reltol, itmax, CG_KRYLOV);
        const int nA = (ndim == 1) ? 2 * ((nvec / 2 + 1) / 2) : (nvec + 1) / 2;      // 1-D: whole pairs per group
        for (int g = 0; g < 2; ++g) {
            CgGroup& G = grp[g];
            G.off = g ? nA : 0; G.n = g ? nvec - nA : nA;
            G.sc = sc_all;
            G.sc.gamma += G.off; G.sc.alpha += G.off; G.sc.beta += G.off; G.sc.eps += G.off; G.sc.rnorm += G.off;
            G.sc.pAp += G.off; G.sc.iters += G.off; G.sc.active += G.off; G.sc.n_active += g;
            if (defer) { G.sc.x_pending += G.off; G.sc.x_count += G.off; }
            G.h = ctx->h_pinned + 64 * (1 + g);
            G.slot = 0; G.pending = false; G.done = false; G.it = 0;
            rt_stream_wait(G.st, ctx->ev0);
        }
        rt_stream main = ctx->stream;
        auto enter = [&](CgGroup& G) { ctx->stream = G.st; std::swap(rws, G.rws); std::swap(ws_tile, G.tile); std::swap(cs, G.cs); };
        auto leave = [&](CgGroup& G) { std::swap(rws, G.rws); std::swap(ws_tile, G.tile); std::swap(cs, G.cs); ctx->stream = main; };
        auto vecs = [&](CgGroup& G, DevBuf& b) { return b.as<double>() + (size_t)G.off * N; };
        auto launch_iter = [&](CgGroup& G, bool first, long long k) {
            FieldPassParams P = base_params();
            double* pb[2] = {vecs(G, ws_p), defer ? vecs(G, ws_p2) : vecs(G, ws_p)};
            double* p_cur = pb[k & 1];
            double* p_prev = pb[(k & 1) ^ 1];
            double* Rg = R + (size_t)G.off * N;
            P.r = vecs(G, ws_r);
            P.p = first ? p_cur : p_prev;
            P.p_out = defer ? p_cur : nullptr;
            P.vin = p_cur; P.vin_ld = N;
            P.out = vecs(G, ws_t); P.out_ld = N;
            P.w = w.as<double>();
            P.first_iter = first ? 1 : 0;
            P.cg = G.sc;
            P.mask_active = 1;
            if (defer) { P.defer_x = defer_mode; P.x_prev_p = p_prev; P.xout = Rg; }
            chain_metric<PRO_CG>(P, G.n, vecs(G, ws_t), FIN_CG_PAP, nullptr);
            if (defer) vec_cg_update_r(ctx, rws, G.sc, vecs(G, ws_r), vecs(G, ws_t), N, G.n);
            else vec_cg_update(ctx, rws, G.sc, Rg, vecs(G, ws_r), vecs(G, ws_p), vecs(G, ws_t), N, G.n);
        };
        // right-hand sides, first iteration and (once per shape) graph capture, group after group
        for (int g = 0; g < 2; ++g) {
            CgGroup& G = grp[g];
            enter(G);
            try {
                FieldPassParams P = base_params();
                P.q = q.as<double>(); P.Z1 = Z1 + (size_t)G.off * z1_ld; P.Z1_ld = z1_ld; P.Z1_off = 0; P.Z1_stride = z1_stride;
                P.Z2 = Z2 + (size_t)G.off * N; P.Z2_ld = N;
                P.r = vecs(G, ws_r); P.p = vecs(G, ws_p); P.xout = R + (size_t)G.off * N;
                P.scale = c * pow(0.5, ndim);
                P.cg = G.sc;
                chain_adjoint<PRO_RHS, EPI_RHS>(P, G.n, vecs(G, ws_t), FIN_RHS, nullptr);
                G.h[0] = 1;
                rt_d2h(G.h, G.sc.n_active, sizeof(int), ctx->stream);
            } catch (...) { leave(G); throw; }
            leave(G);
        }
        for (int g = 0; g < 2; ++g) {
            CgGroup& G = grp[g];
            rt_sync(G.st);
            if (!(G.h[0] > 0 && itmax > 0)) { G.done = true; continue; }
            enter(G);
            try {
                launch_iter(G, true, 0);
                G.it = 1;
                if (G.it >= itmax) { G.done = true; }
                else {
                    const CgGraphKey key{R + (size_t)G.off * N, vecs(G, ws_r), vecs(G, ws_p), defer ? (void*)vecs(G, ws_p2) : nullptr,
                                         vecs(G, ws_t), ws_tile.p, w_tiled.p, tmpA.p, tmpB.p, cs.p, rws.partials.p,
                                         rws.counters.p, G.n, o.abstol, o.reltol, itmax};
                    if (!G.gvalid || !(key == G.key)) {
                        if (G.gvalid) { rt_graph_destroy(G.graph); G.gvalid = false; }
                        const int64_t l0 = ctx->launches;
                        rt_capture_begin(ctx->stream);
                        try {
                            for (int k = 0; k < kCgChunk; ++k) launch_iter(G, false, 1 + k);
                            G.graph = rt_capture_end(ctx->stream);
                        } catch (...) {
                            rt_capture_abort(ctx->stream);
                            throw;
                        }
                        G.glaunch = ctx->launches - l0;
                        ctx->launches = l0;
                        G.key = key;
                        G.gvalid = true;
                    }
                }
            } catch (...) { leave(G); throw; }
            leave(G);
        }
        // replay: one chunk per group in turn, each group polled one chunk behind its launches
        const long long it_cap = itmax + 2 * kCgChunk;
        while (!(grp[0].done && grp[1].done)) {
            for (int g = 0; g < 2; ++g) {
                CgGroup& G = grp[g];
                if (G.done) continue;
                rt_graph_launch(G.graph, G.st);
                ctx->launches += G.glaunch;
                G.it += kCgChunk;
                rt_d2h(G.h + 16 * (1 + G.slot), G.sc.n_active, sizeof(int), G.st);
                rt_event_record(G.ev[G.slot], G.st);
                if (G.pending) {
                    rt_event_sync(G.ev[G.slot ^ 1]);
                    if (G.h[16 * (1 + (G.slot ^ 1))] <= 0) { G.done = true; continue; }
                }
                if (G.it >= it_cap) { G.done = true; continue; }
                G.pending = true;
                G.slot ^= 1;
            }
        }
        for (int g = 0; g < 2; ++g) {
            CgGroup& G = grp[g];
            enter(G);
            try {
                if (defer && G.it > 0) {
                    double* pb[2] = {vecs(G, ws_p), vecs(G, ws_p2)};
                    vec_x_flush(ctx, rws, G.sc, R + (size_t)G.off * N, pb[(G.it - 1) & 1], N, G.n);
                }
                rt_event_record(G.ev_done, ctx->stream);
            } catch (...) { leave(G); throw; }
            leave(G);
            rt_stream_wait(ctx->stream, G.ev_done);
        }
        rt_event_record(ctx->ev1, ctx->stream);
        std::vector<int> hi(nvec);
        std::vector<double> hr(nvec);
        rt_d2h(hi.data(), cgws.iters.p, sizeof(int) * nvec, ctx->stream);
        rt_d2h(hr.data(), cgws.rnorm.p, sizeof(double) * nvec, ctx->stream);
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

    // MatrixInversion (src/residual_samplers.jl:95-123): the reference materialises M = J'FJ + I by n_theta
    // operator applications (:102-104) -- here one batched metric product on the identity -- then the
    // shared Cholesky sampler (dense_engine.cu).  Meant for small fields (the dense-vs-CG Bartlett check of
    // test/test_samplers.jl:17-39); the matrix is n_theta^2 doubles.
    void sample_dense(const double* Z, int64_t n, double* R, double* L) override {
        if (N > kDenseFieldMax)
            throw RtError{"FIELD_DHT: the MatrixInversion sampler materialises the n_theta x n_theta metric; provided for fields of up to 8192 pixels", 3};
        Mdense.ensure(sizeof(double) * N * N);
        eye.ensure(sizeof(double) * N * N);
        if (!chol) chol = dense_chol_create(ctx);
        dense_chol_identity(chol, eye.as<double>(), N);
        for (long long c0 = 0; c0 < N; c0 += 1024) {           // bounded scratch: 1024 columns per batch
            const long long nc = std::min<long long>(1024, N - c0);
            metric_apply(eye.as<double>() + c0 * N, nc, Mdense.as<double>() + c0 * N);
        }
        dense_chol_sample(chol, Mdense.as<double>(), N, Z, n, R, L);
    }

    double kl_value_grad(const double* x, const double* R, int64_t n, double* grad) override {
        const int npoints = (int)(2 * n);
        const CanonShape cs = canon_shape(ctx, n);
        // per-point scalars: [prior term 1/2 |p|^2 (npoints)] [variable part of -loglike (npoints)]
        kl_red.ensure(sizeof(double) * 2 * npoints);
        double* red = kl_red.as<double>();
        rt_memset(red, 0, sizeof(double) * 2 * npoints, ctx->stream);
        pt_t.ensure(sizeof(double) * N * npoints);
        FieldPassParams P = base_params();
        P.x = x; P.R = R; P.R_ld = N; P.point_mode = POINT_PM; P.want_grad = grad ? 1 : 0;
        // per-point gradient inputs g'(s_p) dl/dlambda(p) land in the scratch itself (the last pass writes
        // the positions it read)
        P.out = pt_t.as<double>(); P.out_ld = N;
        chain_forward<PRO_POINT, EPI_KL>(P, npoints, pt_t.as<double>(), FIN_VEC_STORE, red, FIN_VEC_STORE, red + npoints);
        // sums over the residual samples in the canonical, sharding-independent order (engine.h)
        const double ftot = canon_scalar_sum(ctx, sc_gather, red, 2, npoints);
        if (grad) {
            double* qs = qsum.as<double>();
            vec_canon_sum(ctx, rws, cs, pt_t.as<double>(), N, 2, 1.0, qs, N, gather);
            // grad = x - (1/2n) c A . H( sum_p g'(s_p) dl/dlambda(p) )
            FieldPassParams Q = base_params();
            Q.in = qs; Q.in_ld = N; Q.x = x; Q.out = grad;
            Q.scale = c * pow(0.5, ndim) / (2.0 * (double)cs.n_glob);
            chain_adjoint<PRO_LOAD, EPI_GRAD>(Q, 1, pt_t.as<double>(), FIN_NONE, nullptr);
        }
        return ftot / (2.0 * (double)cs.n_glob) + const_term;
    }

    void mean_metric_prepare(const double* x, const double* R, int64_t n) override {
        check_pairing(n);
        const CanonShape cs = canon_shape(ctx, n);
        pt_t.ensure(sizeof(double) * N * n);
        FieldPassParams P = base_params();
        P.x = x; P.R = R; P.R_ld = N; P.point_mode = POINT_PLUS;
        P.out = pt_t.as<double>(); P.out_ld = N;      // per-sample weights w_i = F g'^2 at x + r_i
        chain_forward<PRO_POINT, EPI_LINEARIZE>(P, (int)n, pt_t.as<double>(), FIN_NONE, nullptr, FIN_NONE, nullptr);
        vec_canon_sum(ctx, rws, cs, pt_t.as<double>(), N, 1, 1.0 / (double)cs.n_glob, wbar.as<double>(), N, gather);
        ++w_gen;
    }

    void mean_metric_apply(const double* u, double* out) override {
        FieldPassParams P = base_params();
        P.in = u; P.in_ld = N; P.vin = u; P.vin_ld = N; P.out = out; P.out_ld = N; P.w = wbar.as<double>();
        n_t.ensure(sizeof(double) * N);
        chain_metric<PRO_AMP>(P, 1, n_t.as<double>(), FIN_NONE, nullptr);
    }

    void ncg_init(const double* g) override {
        const size_t b = sizeof(double) * N;
        n_x.ensure(b); n_r.ensure(b); n_p.ensure(b); n_t.ensure(b);
        ncgws.ensure(1);
        // IterativeSolvers cg_iterator!(dx, A, b; initially_zero = true): reltol = sqrt(eps), abstol = 0,
        // maxiter = size(A, 2)  (src/newtoncg.jl:120)
        ncg_sc = ncgws.scalars(0.0, 1.4901161193847656e-08, N, CG_ITSOLVERS);
        rt_memset(ncgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        vec_cg_init(ctx, rws, ncg_sc, g, N, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), N, 1);
        ncg_first = true;
    }

    void ncg_set_state(const double* x, const double* r, const double* u, double res, double prev) override {
        const size_t b = sizeof(double) * N;
        n_x.ensure(b); n_r.ensure(b); n_p.ensure(b); n_t.ensure(b);
        ncgws.ensure(1);
        ncg_sc = ncgws.scalars(0.0, 1.4901161193847656e-08, N, CG_ITSOLVERS);
        rt_d2d(n_x.p, x, b, ctx->stream);
        rt_d2d(n_r.p, r, b, ctx->stream);
        rt_d2d(n_p.p, u, b, ctx->stream);
        // device form of the iterator state: gamma = residual^2, beta = residual^2 / prev^2 (what the previous
        // update's scalar tail would have left, field_pass.h: FIN_CG_RR), the vector active, tolerance 0
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
        rt_d2h(h_ncg, ncgws.active.p, sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        return h_ncg[0] != 0;
    }

    void ncg_update() override {
        FieldPassParams P = base_params();
        P.r = n_r.as<double>(); P.p = n_p.as<double>();
        P.vin = n_p.as<double>(); P.vin_ld = N;
        P.out = n_t.as<double>(); P.out_ld = N;
        P.w = wbar.as<double>();
        P.first_iter = ncg_first ? 1 : 0;
        P.cg = ncg_sc;
        P.mask_active = 1;
        chain_metric<PRO_CG>(P, 1, n_t.as<double>(), FIN_CG_PAP, nullptr);
        vec_cg_update(ctx, rws, ncg_sc, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), n_t.as<double>(), N, 1);
        ncg_first = false;
    }

    const double* ncg_dx() override { return n_x.as<double>(); }

    double ncg_residual() override {
        double h = 0.0;
        rt_d2h(&h, ncgws.rnorm.p, sizeof(double), ctx->stream);
        rt_sync(ctx->stream);
        return h;
    }

    void unpermute_big(double* arr) const {
        // device order is the permuted data-space order [k1][k2] <-> k = k1 + n1 k2
        const long long n1 = 1LL << lg_n1, n2 = 1LL << lg_n2;
        std::vector<double> t(N);
        for (long long r = 0; r < n1; ++r)
            for (long long cc = 0; cc < n2; ++cc) t[r + n1 * cc] = arr[r * n2 + cc];
        memcpy(arr, t.data(), sizeof(double) * N);
    }

    void get_mean_weights(double* w_host) override {
        rt_d2h(w_host, wbar.p, sizeof(double) * N, ctx->stream);
        rt_sync(ctx->stream);
        if (big) unpermute_big(w_host);
    }
};

}  // namespace

Engine* make_field_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d) {
    FieldEngine* e = new FieldEngine();
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
