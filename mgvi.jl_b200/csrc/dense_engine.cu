// Dense-linear engine -- BASELINE.json configs[3]: mu = J xi with a dense (m x k) Jacobian,
// Normal(mu, sigma) data; the `DenseMatrix` operator type `_get_operator_type(::MatrixInversion)`
// selects (src/residual_samplers.jl:56) and the FullJacobianFunc / FullResidualSampler pair of
// BASELINE.json's north_star.
//
//   M = J' F J + I  (F = 1/sigma^2)            FP64 tensor-core rank-k update (dmma_gemm.h)
//   MatrixInversion sampler (src/residual_samplers.jl:95-123): the reference forms inv(M) and its
//   lower Cholesky factor L_Sigma; here (restated for an upper factor):
//       Mf = P M P (index reversal) = R' R  (blocked upper Cholesky; every block operation except the 64 x 64
//                                            diagonal factorisation is a DMMA GEMM, see DenseChol::factorize)
//       L_Sigma = P R^-1 P   (lower, L_Sigma L_Sigma' = M^-1),   residuals = P R^-1 (P Z)
//   no explicit inverse of M is formed.
//   CG sampler and Newton phase work on M directly (k x k GEMM per iteration).
//   KL value / gradient use the quadratic form of the Normal-linear model:
//       -loglike(p) = 1/2 d'd/sigma^2 - p'h + 1/2 p'G p,  h = J'd/sigma^2,  G = M - I
#include <algorithm>
#include "engine.h"
#include "kernels.h"
#include "dmma_gemm.h"

namespace mgvi {

static const int kNb = 64;     // Cholesky / triangular-solve block size

template <bool KW>
SIMT_GLOBAL2(256, 1) k_dmma_gemm(SIMT_KARGS(GemmParams)) { SIMT_SMEM_DECL gemm_body<KW>(P, simt_smem); }

struct TriParams {
    double* A; long long ld;      // matrix holding the nb x nb diagonal block at (j0, j0)
    double* X; long long ldx;     // right-hand sides / panel
    int j0, nb, ncols, c0, mode, n;
    double* Dinv;                 // TRI_POTRF: where the inverse of the factored diagonal block goes (kNb x kNb, column-major)
};
enum { TRI_POTRF = 0, TRI_SYM = 3, TRI_EYE = 4 };

// TRI_POTRF: upper Cholesky of the diagonal block in place (one block of 256 threads).
// TRI_SYM  : mirror the upper triangle into the lower one.   TRI_EYE: A = identity.
//
// The diagonal kernel works on a full kNb x kNb block held in registers; a short last block (nb < kNb) is
// padded with the identity, which the factorisation and the inversion leave alone.
SIMT_DEVICE void tri_fill_body(const TriParams& P) {
    const long long total = (long long)P.n * P.n;
    const int tid = SIMT_TID, nt = SIMT_NTID;
    for (long long idx = SIMT_BID * nt + tid; idx < total; idx += SIMT_NBID * nt) {
        const long long r = idx % P.n, c = idx / P.n;
        if (P.mode == TRI_EYE) P.A[r + c * P.ld] = (r == c) ? 1.0 : 0.0;
        else if (r > c) P.A[r + c * P.ld] = P.A[c + r * P.ld];
    }
}

// diagonal block -> shared S[row * lds + col], identity-padded to kNb x kNb
SIMT_DEVICE void tri_load_diag(const TriParams& P, double* S, int lds) {
    const int nb = P.nb, tid = SIMT_TID, nt = SIMT_NTID;
    for (int idx = tid; idx < kNb * kNb; idx += nt) {
        const int r = idx % kNb, c = idx / kNb;
        S[r * lds + c] = (r < nb && c < nb) ? P.A[(P.j0 + r) + (long long)(P.j0 + c) * P.ld] : (r == c ? 1.0 : 0.0);
    }
}

// Right-looking Cholesky of one kNb x kNb block, 256 threads.  Thread (g = tid / 64, b = tid % 64)
// keeps the 16 elements S[4 i + g][b] in registers; per column j: the 64 owners of row j publish
// it, 64 threads scale it (R[j][c] = S[j][c] / sqrt(S[j][j])), everyone applies the rank-1 update
// to its registers.  Two barriers per column, no shared-memory traffic for the trailing block.
SIMT_DEVICE void potrf_body(const TriParams& P, unsigned char* smem) {
    constexpr int lds = kNb + 1;
    double* S = reinterpret_cast<double*>(smem);          // [kNb][kNb + 1]: input block, then the factor
    double* rowbuf = S + kNb * lds;                       // [2][kNb] unscaled pivot row, double-buffered
    const int tid = SIMT_TID, g = tid >> 6, b = tid & 63;
    tri_load_diag(P, S, lds);
    SIMT_SYNC();
    double reg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) reg[i] = S[(4 * i + g) * lds + b];
    SIMT_SYNC();
    for (int j = 0; j < kNb; ++j) {
        double* rb_ = rowbuf + (j & 1) * kNb;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (4 * i + g == j) rb_[b] = reg[i];
        SIMT_SYNC();
        // every thread scales by 1 / sqrt(pivot) itself (as LAPACK's dpotf2 scales by the reciprocal):
        // one barrier per column, no serial divide behind the square root
        const double rinv = SIMT_RSQRT(rb_[j]);
        const double rb = rb_[b] * rinv;
        if (g == 0) S[j * lds + b] = (b >= j) ? rb : 0.0;  // row j of R
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int a = 4 * i + g;
            if (a > j && a <= b) reg[i] -= (rb_[a] * rinv) * rb;
        }
    }
    SIMT_SYNC();
    for (int idx = tid; idx < kNb * kNb; idx += SIMT_NTID) {
        const int r = idx % kNb, c = idx / kNb;
        if (r < P.nb && c < P.nb) P.A[(P.j0 + r) + (long long)(P.j0 + c) * P.ld] = S[r * lds + c];
    }
    if (!P.Dinv) return;
    // inverse of the diagonal block, X = R^-1 (upper triangular): thread c < 64 solves R x = e_c by right-looking
    // back substitution in 64 registers.  With it the panel solve R_jj^-T A_panel and the block steps of the
    // triangular solves are plain GEMMs on the tensor pipe (the usual GPU formulation of trsm); the padded
    // identity of a short last block inverts to itself.
    double* dinv = rowbuf;                                // [kNb] reciprocal diagonal
    if (tid < kNb) dinv[tid] = 1.0 / S[tid * lds + tid];
    SIMT_SYNC();
    double t[kNb];
    if (tid < kNb) {
#pragma unroll
        for (int i = 0; i < kNb; ++i) t[i] = (i == tid) ? 1.0 : 0.0;
#pragma unroll
        for (int i = kNb - 1; i >= 0; --i) {
            const double y = t[i] * dinv[i];
            t[i] = y;
#pragma unroll
            for (int l = 0; l < i; ++l) t[l] -= S[l * lds + i] * y;
        }
    }
    SIMT_SYNC();                                          // everyone is done reading S
    if (tid < kNb) {
#pragma unroll
        for (int i = 0; i < kNb; ++i) S[i * lds + tid] = t[i];
    }
    SIMT_SYNC();
    for (int idx = tid; idx < kNb * kNb; idx += SIMT_NTID) {
        const int r = idx % kNb, c = idx / kNb;
        P.Dinv[r + c * kNb] = S[r * lds + c];
    }
}

SIMT_GLOBAL(128) k_tri_fill(SIMT_KARGS(TriParams)) { SIMT_SMEM_DECL (void)simt_smem; tri_fill_body(P); }
SIMT_GLOBAL(256) k_potrf(SIMT_KARGS(TriParams)) { SIMT_SMEM_DECL potrf_body(P, simt_smem); }

// MatrixInversion sampler (src/residual_samplers.jl:95-123) on a materialised symmetric metric M (k x k,
// device memory): shared by the dense-linear engine and by the field engines' small-model dense path.
struct DenseChol {
    mgvib200_ctx* ctx = nullptr;
    long long k = 0;
    DevBuf Mf, W, Rinv, Dinv;      // Dinv: inverses of the diagonal blocks of R, [nblk][kNb x kNb]
    ReduceWs rws;
    bool have_chol = false;

    void release() {
        DevBuf* all[] = {&Mf, &W, &Rinv, &Dinv, &rws.partials, &rws.counters, &rws.red};
        for (auto* b : all) b->release();
        have_chol = false;
    }

    void gemm(const double* A, long long a_row, long long a_k, const double* B, long long b_k, long long b_col,
              double* C, long long ldc, int Mr, int Nc, int Kr, double alpha, double beta, bool upper_only = false,
              const double* kweight = nullptr) {
        if (Mr <= 0 || Nc <= 0) return;
        GemmParams g;
        g.A = A; g.a_row = a_row; g.a_k = a_k; g.B = B; g.b_k = b_k; g.b_col = b_col; g.C = C; g.ldc = ldc;
        g.M = Mr; g.N = Nc; g.K = Kr; g.alpha = alpha; g.beta = beta; g.upper_only = upper_only ? 1 : 0;
        g.kweight = kweight;
        g.ntm = (Mr + kGemmBM - 1) / kGemmBM; g.ntn = (Nc + kGemmBN - 1) / kGemmBN;
        ctx->prof_pre(3000 + (upper_only ? 1 : 0));
        if (kweight) rt_launch(k_dmma_gemm<true>, (long long)g.ntm * g.ntn, 256, sizeof(double) * 2 * kGemmBM * kGemmLd, ctx->stream, g);
        else rt_launch(k_dmma_gemm<false>, (long long)g.ntm * g.ntn, 256, sizeof(double) * 2 * kGemmBM * kGemmLd, ctx->stream, g);
        ctx->prof_post();
        ctx->launches++;
    }

    void tri(int mode, double* A, long long ld, double* X, long long ldx, int j0, int nb, int c0, int ncols, int n,
             double* dinv = nullptr) {
        TriParams t;
        t.A = A; t.ld = ld; t.X = X; t.ldx = ldx; t.j0 = j0; t.nb = nb; t.c0 = c0; t.ncols = ncols; t.mode = mode; t.n = n;
        t.Dinv = dinv;
        ctx->prof_pre(3100 + mode);
        if (mode == TRI_SYM || mode == TRI_EYE) {
            rt_launch(k_tri_fill, std::min<long long>(((long long)n * n + 127) / 128, 4096), 128, 0, ctx->stream, t);
        } else {
            rt_launch(k_potrf, 1, 256, sizeof(double) * ((size_t)kNb * (kNb + 1) + 2 * kNb), ctx->stream, t);
        }
        ctx->prof_post();
        ctx->launches++;
    }

    VecParams vp() const {
        VecParams P;
        memset(&P, 0, sizeof P);
        P.nvec = 1; P.N = k;
        return P;
    }

    static const int kOuter = 256;    // columns per outer panel: the big trailing updates run with K = 256

    // Mf = P M P = R'R, R upper; the factor stays in Mf (upper triangle), inv(R_jj) of every 64 x 64 diagonal
    // block in Dinv.  Two-level right-looking: inside an outer panel of 256 columns each 64-block is factored
    // (one CTA, emits its inverse), its block row becomes X_j = inv(R_jj)' A[j, :] by ONE GEMM over all columns
    // to the right, and only the remaining rows of the panel are updated (K = 64, at most 192 rows); the rest of
    // the matrix gets one rank-256 update per panel -- 16 large DMMA launches at n = 4096 instead of 63 rank-64
    // ones, no thread-per-column substitution kernels on the critical path.
    void factorize(const double* M) {
        if (have_chol) return;
        Mf.ensure(8 * k * k);
        const long long nblk = (k + kNb - 1) / kNb;
        Dinv.ensure(8 * (size_t)nblk * kNb * kNb);
        // flip rows and columns: Mf[i][j] = M[k-1-i][k-1-j]; with M symmetric a flip of the flat array does it
        VecParams P = vp();
        P.op = VOP_FLIP; P.N = k * k; P.in = M; P.out = Mf.as<double>();
        vec_launch(ctx, rws, P);
        double* A = Mf.as<double>();
        for (long long J0 = 0; J0 < k; J0 += kOuter) {
            const long long Jend = std::min<long long>(k, J0 + kOuter);
            for (long long j0 = J0; j0 < Jend; j0 += kNb) {
                const int nb = (int)std::min<long long>(kNb, k - j0);
                double* di = Dinv.as<double>() + (j0 / kNb) * kNb * kNb;
                tri(TRI_POTRF, A, k, nullptr, 0, (int)j0, nb, 0, 0, (int)k, di);
                const long long c0 = j0 + nb;
                const int rest = (int)(k - c0);
                if (rest <= 0) continue;
                // X_j = inv(R_jj)' A[j, c0:]  in place (one tile row: a block reads its own columns, then writes them)
                double* Xj = A + j0 + c0 * k;
                gemm(di, kNb, 1, Xj, 1, k, Xj, k, nb, rest, nb, 1.0, 0.0);
                // rows of this panel below block j: A[c0:Jend, c0:] -= X_j[:, c0:Jend]' X_j   (upper block tiles)
                const int mrows = (int)(Jend - c0);
                if (mrows > 0) gemm(Xj, k, 1, Xj, 1, k, A + c0 + c0 * k, k, mrows, rest, nb, -1.0, 1.0, true);
            }
            // everything below the panel: A[Jend:, Jend:] -= X_J' X_J with X_J = A[J0:Jend, Jend:]
            const int rest = (int)(k - Jend);
            if (rest > 0) {
                const double* XJ = A + J0 + Jend * k;
                gemm(XJ, k, 1, XJ, 1, k, A + Jend + Jend * k, k, rest, rest, (int)(Jend - J0), -1.0, 1.0, true);
            }
        }
        have_chol = true;
    }

    // X (k x ncols, leading dimension k) <- R^-1 X, blocked back substitution with the same two levels:
    // Y_j = inv(R_jj) X_j (GEMM, in place), rows of the outer panel above it updated with K = 64, everything above
    // the panel with one K = 256 GEMM
    void back_solve(double* X, int ncols) {
        const double* A = Mf.as<double>();
        const long long npan = (k + kOuter - 1) / kOuter;
        for (long long pb = npan - 1; pb >= 0; --pb) {
            const long long J0 = pb * kOuter, Jend = std::min<long long>(k, J0 + kOuter);
            const long long nin = (Jend - J0 + kNb - 1) / kNb;
            for (long long ib = nin - 1; ib >= 0; --ib) {
                const long long j0 = J0 + ib * kNb;
                const int nb = (int)std::min<long long>(kNb, k - j0);
                const double* di = Dinv.as<double>() + (j0 / kNb) * kNb * kNb;
                // Y_j = inv(R_jj) X[j0:j0+nb, :]   (A-operand (i, l) = Dinv[i + l * 64])
                gemm(di, 1, kNb, X + j0, 1, k, X + j0, k, nb, ncols, nb, 1.0, 0.0);
                // X[J0:j0, :] -= R[J0:j0, j0:j0+nb] Y_j
                if (j0 > J0) gemm(A + J0 + j0 * k, 1, k, X + j0, 1, k, X + J0, k, (int)(j0 - J0), ncols, nb, -1.0, 1.0);
            }
            // X[0:J0, :] -= R[0:J0, J0:Jend] Y_J
            if (J0 > 0) gemm(A + J0 * k, 1, k, X + J0, 1, k, X, k, (int)J0, ncols, (int)(Jend - J0), -1.0, 1.0);
        }
    }

    void sample(const double* M, const double* Z, int64_t n, double* R, double* L) {
        factorize(M);
        const int nvec = (int)n;
        W.ensure(8 * (size_t)k * std::max<long long>(nvec, 1));
        // W = P Z ; W <- R^-1 W ; residuals = P W
        VecParams P = vp();
        P.op = VOP_FLIP; P.nvec = nvec; P.in = Z; P.out = W.as<double>();
        vec_launch(ctx, rws, P);
        back_solve(W.as<double>(), nvec);
        P.in = W.as<double>(); P.out = R;
        vec_launch(ctx, rws, P);
        if (L) {
            // L_Sigma = P R^-1 P: X = R^-1 (back substitution on the identity); reversing the flat
            // column-major array maps (i, j) -> (k-1-i, k-1-j)
            Rinv.ensure(8 * k * k);
            tri(TRI_EYE, Rinv.as<double>(), k, nullptr, 0, 0, kNb, 0, 0, (int)k);
            back_solve(Rinv.as<double>(), (int)k);
            VecParams Q = vp();
            Q.op = VOP_FLIP; Q.N = k * k; Q.in = Rinv.as<double>(); Q.out = L;
            vec_launch(ctx, rws, Q);
        }
    }
};

DenseChol* dense_chol_create(mgvib200_ctx* ctx) {
    DenseChol* c = new DenseChol();
    c->ctx = ctx;
    return c;
}
void dense_chol_destroy(DenseChol* c) {
    if (!c) return;
    c->release();
    delete c;
}
void dense_chol_identity(DenseChol* c, double* eye, long long k) {
    c->tri(TRI_EYE, eye, k, nullptr, 0, 0, kNb, 0, 0, (int)k);
}
void dense_chol_sample(DenseChol* c, const double* M, long long k, const double* Z, int64_t n, double* R, double* L) {
    c->k = k;
    c->have_chol = false;        // M belongs to the caller and may have changed (new linearisation)
    c->sample(M, Z, n, R, L);
}

namespace {

struct DenseEngine : Engine {
    long long m = 0, k = 0;
    double sigma = 1.0, inv_s2 = 1.0, dd = 0.0, const_term = 0.0;
    int layout = 0;
    long long z1_stride = 1;
    DevBuf J, data, M, h, center, P2, MP, gsum, ws_r, ws_p, ws_t, n_x, n_r, n_p, n_t;
    DevBuf fvec, sqf;            // per-datum Fisher weight 1/sigma_i^2 and its square root (null: constant sigma)
    std::vector<double> h_f;     // host copy (get_weights)
    ReduceWs rws;
    CgWs cgws, ncgws;
    CgScalars ncg_sc;
    bool ncg_first = true;
    int h_act = 0;

    bool counted = false;
    // "replicas only" (SURVEY.md 8e): mgvib200_comm_init refuses contexts that hold a dense model, setup()
    // refuses a context that already has a communicator; this covers every entry point that would sum
    // over residual samples
    void single_rank_only() const {
        if (ctx->world > 1) throw RtError{"DENSE_LINEAR: replicas only (SURVEY.md 8e); run one replica per GPU", 3};
    }

    ~DenseEngine() override {
        if (counted) ctx->n_models_single_rank--;
        DevBuf* all[] = {&J, &data, &M, &h, &center, &P2, &MP, &gsum, &ws_r, &ws_p, &ws_t, &n_x, &n_r, &n_p,
                         &n_t, &fvec, &sqf, &rws.partials, &rws.counters, &rws.red};
        for (auto* b : all) b->release();
        chol.release();
    }

    DenseChol chol;
    void gemm(const double* A, long long a_row, long long a_k, const double* B, long long b_k, long long b_col,
              double* C, long long ldc, int Mr, int Nc, int Kr, double alpha, double beta, bool upper_only = false,
              const double* kweight = nullptr) {
        chol.gemm(A, a_row, a_k, B, b_k, b_col, C, ldc, Mr, Nc, Kr, alpha, beta, upper_only, kweight);
    }
    void tri(int mode, double* A, long long ld, double* X, long long ldx, int j0, int nb, int c0, int ncols, int n) {
        chol.tri(mode, A, ld, X, ldx, j0, nb, c0, ncols, n);
    }

    VecParams vp() const {
        VecParams P;
        memset(&P, 0, sizeof P);
        P.nvec = 1; P.N = k;
        return P;
    }

    void setup(const mgvib200_model_desc& d) {
        m = d.m; k = d.k; sigma = d.sigma; layout = d.lambda_layout;
        if (d.likelihood != MGVIB200_LIKE_NORMAL) throw RtError{"DENSE_LINEAR: Normal likelihood only", 1};
        if (!d.J || !d.data || m < 1 || k < 1 || d.n_data != m) throw RtError{"DENSE_LINEAR: J (m x k) and m data values are required", 1};
        if (!(sigma > 0)) throw RtError{"DENSE_LINEAR: sigma must be positive", 1};
        if (ctx->world > 1) throw RtError{"DENSE_LINEAR: replicas only (SURVEY.md 8e); run one replica per GPU", 3};
        inv_s2 = (1.0 / sigma) * (1.0 / sigma);
        n_theta = k;
        n_lambda = (layout == MGVIB200_LAYOUT_MU_ONLY) ? m : 2 * m;
        z1_stride = (layout == MGVIB200_LAYOUT_INTERLEAVED) ? 2 : 1;
        chol.ctx = ctx;
        J.ensure(8 * m * k); data.ensure(8 * m); M.ensure(8 * k * k); h.ensure(8 * k); center.ensure(8 * k);
        rt_h2d(J.p, d.J, 8 * m * k, ctx->stream);
        rt_h2d(data.p, d.data, 8 * m, ctx->stream);
        dd = 0.0;
        const double* fw = nullptr;
        double a_scale = inv_s2;
        if (d.sigma_vec) {
            // per-datum noise level: F = diag(1 / sigma_i^2) (src/information.jl:12-18), folded into the
            // A-operand load of the DMMA rank-m update
            h_f.resize(m);
            std::vector<double> hq(m);
            double logs = 0.0;
            for (long long i = 0; i < m; ++i) {
                const double sg = d.sigma_vec[i];
                if (!(sg > 0)) throw RtError{"DENSE_LINEAR: every sigma_vec entry must be positive", 1};
                h_f[i] = (1.0 / sg) * (1.0 / sg);
                hq[i] = 1.0 / sg;
                dd += h_f[i] * d.data[i] * d.data[i];
                logs += log(sg);
            }
            fvec.ensure(8 * m); sqf.ensure(8 * m);
            rt_h2d(fvec.p, h_f.data(), 8 * m, ctx->stream);
            rt_h2d(sqf.p, hq.data(), 8 * m, ctx->stream);
            rt_sync(ctx->stream);
            fw = fvec.as<double>();
            a_scale = 1.0;
            const_term = logs + (double)m * 0.5 * log(2.0 * M_PI) + 0.5 * (double)k * log(2.0 * M_PI);
        } else {
            for (long long i = 0; i < m; ++i) dd += d.data[i] * d.data[i];
            dd *= inv_s2;
            const_term = (double)m * (log(sigma) + 0.5 * log(2.0 * M_PI)) + 0.5 * (double)k * log(2.0 * M_PI);
        }
        // M = I + J'FJ : identity, rank-m update of the upper block tiles on DMMA, mirror
        tri(TRI_EYE, M.as<double>(), k, nullptr, 0, 0, kNb, 0, 0, (int)k);
        gemm(J.as<double>(), m, 1, J.as<double>(), 1, m, M.as<double>(), k, (int)k, (int)k, (int)m, a_scale, 1.0, true, fw);
        tri(TRI_SYM, M.as<double>(), k, nullptr, 0, 0, kNb, 0, 0, (int)k);
        // h = J'F d
        gemm(J.as<double>(), m, 1, data.as<double>(), 1, m, h.as<double>(), k, (int)k, 1, (int)m, a_scale, 0.0, false, fw);
        rt_sync(ctx->stream);
    }

    void linearize(const double* x) override { rt_d2d(center.p, x, 8 * k, ctx->stream); }

    void get_weights(double* w_host, double* q_host) override {
        for (long long i = 0; i < m; ++i) {
            if (w_host) w_host[i] = h_f.empty() ? inv_s2 : h_f[i];
            if (q_host) q_host[i] = h_f.empty() ? 1.0 / sigma : sqrt(h_f[i]);
        }
    }

    void apply_M(const double* V, int nv, double* out) {
        // M symmetric: out[a, v] = sum_i M[i + a k] V[i + v k]
        gemm(M.as<double>(), k, 1, V, 1, k, out, k, (int)k, nv, (int)k, 1.0, 0.0);
    }

    void metric_apply(const double* V, int64_t nv, double* MV) override { apply_M(V, (int)nv, MV); }

    void sample_cg(const double* Z1, const double* Z2, int64_t n, const mgvib200_cg_opts& o, double* R,
                   int64_t* iters_host, double* rnorm_host) override {
        const int nvec = (int)n;
        const size_t vb = 8 * (size_t)k * nvec;
        ws_r.ensure(vb); ws_p.ensure(vb); ws_t.ensure(vb);
        cgws.ensure(nvec);
        const long long itmax = o.maxiters > 0 ? o.maxiters : k;
        CgScalars sc = cgws.scalars(o.abstol, o.reltol, itmax, CG_KRYLOV);
        rt_memset(cgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        rt_event_record(ctx->ev0, ctx->stream);
        // b = J'(sqrt(F) n1) + eta   (src/residual_samplers.jl:77): t = eta, t += (1/sigma) J' n1_mu
        rt_d2d(ws_t.p, Z2, vb, ctx->stream);
        gemm(J.as<double>(), m, 1, Z1, z1_stride, n_lambda, ws_t.as<double>(), k, (int)k, nvec, (int)m,
             sqf.p ? 1.0 : 1.0 / sigma, 1.0, false, sqf.p ? sqf.as<double>() : nullptr);
        vec_cg_init(ctx, rws, sc, ws_t.as<double>(), k, R, ws_r.as<double>(), ws_p.as<double>(), k, nvec);
        int* hp = ctx->h_pinned;
        rt_d2h(hp, cgws.n_active.p, sizeof(int), ctx->stream);
        rt_sync(ctx->stream);
        long long it = 0;
        int poll = 1;
        while (hp[0] > 0 && it < itmax) {
            for (int q = 0; q < poll && it < itmax; ++q, ++it) {
                if (it > 0) {
                    VecParams P = vp();
                    P.op = VOP_CG_PUPD; P.nvec = nvec; P.r = ws_r.as<double>(); P.p = ws_p.as<double>(); P.cg = sc;
                    vec_launch(ctx, rws, P);
                }
                apply_M(ws_p.as<double>(), nvec, ws_t.as<double>());
                {
                    VecParams P = vp();
                    P.op = VOP_CG_PAP; P.nvec = nvec; P.p = ws_p.as<double>(); P.Ap = ws_t.as<double>(); P.cg = sc;
                    P.fin = FIN_CG_PAP;
                    vec_launch(ctx, rws, P);
                }
                vec_cg_update(ctx, rws, sc, R, ws_r.as<double>(), ws_p.as<double>(), ws_t.as<double>(), k, nvec);
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

    void sample_dense(const double* Z, int64_t n, double* R, double* L) override {
        // M is fixed at model creation, so the factor is kept across calls
        chol.k = k;
        chol.sample(M.as<double>(), Z, n, R, L);
    }

    double kl_value_grad(const double* x, const double* R, int64_t n, double* grad) override {
        single_rank_only();
        const int np = (int)(2 * n);
        P2.ensure(8 * (size_t)k * np); MP.ensure(8 * (size_t)k * np); gsum.ensure(8 * k);
        vec_build_samples(ctx, rws, x, R, k, (int)n, P2.as<double>());          // points x +- r_i
        apply_M(P2.as<double>(), np, MP.as<double>());
        VecParams Q = vp();
        Q.op = VOP_QUAD; Q.nvec = np; Q.in = P2.as<double>(); Q.y = MP.as<double>(); Q.Ap = h.as<double>();
        Q.fin = FIN_TOTAL;
        vec_launch(ctx, rws, Q);
        double tot = 0.0;
        rt_d2h(&tot, rws.red.p, 8, ctx->stream);
        if (grad) {
            // grad = mean_p (M p - h)
            vec_sum_chunks(ctx, rws, MP.as<double>(), k, np, 1.0 / (double)np, gsum.as<double>(), k);
            vec_axpy(ctx, rws, grad, gsum.as<double>(), -1.0, h.as<double>(), k);
        }
        rt_sync(ctx->stream);
        return tot / (double)np + 0.5 * dd + const_term;
    }

    void mean_metric_prepare(const double*, const double*, int64_t) override { single_rank_only(); }   // M does not depend on the point
    void mean_metric_apply(const double* u, double* out) override { apply_M(u, 1, out); }

    void ncg_init(const double* g) override {
        const size_t b = 8 * k;
        n_x.ensure(b); n_r.ensure(b); n_p.ensure(b); n_t.ensure(b);
        ncgws.ensure(1);
        ncg_sc = ncgws.scalars(0.0, 1.4901161193847656e-08, k, CG_ITSOLVERS);
        rt_memset(ncgws.n_active.p, 0, sizeof(int) * 4, ctx->stream);
        vec_cg_init(ctx, rws, ncg_sc, g, k, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), k, 1);
        ncg_first = true;
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
        apply_M(n_p.as<double>(), 1, n_t.as<double>());
        VecParams P = vp();
        P.op = VOP_CG_PAP; P.p = n_p.as<double>(); P.Ap = n_t.as<double>(); P.cg = ncg_sc; P.fin = FIN_CG_PAP;
        vec_launch(ctx, rws, P);
        vec_cg_update(ctx, rws, ncg_sc, n_x.as<double>(), n_r.as<double>(), n_p.as<double>(), n_t.as<double>(), k, 1);
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

Engine* make_dense_engine(mgvib200_ctx* ctx, const mgvib200_model_desc& d) {
    DenseEngine* e = new DenseEngine();
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
