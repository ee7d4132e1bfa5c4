// Launch wrappers for the streaming vector kernels (vec_ops.h).
#include "engine.h"
#include <vector>
#include "kernels.h"

namespace mgvi {

SIMT_GLOBAL(256) k_vec(SIMT_KARGS(VecParams)) {
    SIMT_SMEM_DECL
    SIMT_PDL_ENTRY();
    vec_body(P, simt_smem);
}

static const int kVecThreads = 256;
static long long vec_per_block() {
    static long long v = 0;
    if (!v) { const char* e = getenv("MGVIB200_VEC_PER_BLOCK"); v = (e && atoll(e) >= 256) ? (atoll(e) & ~1LL) : 4096; }
    return v;
}
#define kVecPerBlock (vec_per_block())

void reduce_ws_ensure(ReduceWs& ws, mgvib200_ctx* ctx, size_t npartials, size_t ncounters) {
    ws.partials.ensure(sizeof(double) * npartials);
    ws.red.ensure(sizeof(double) * 64);
    if (ncounters > ws.ncounters) {
        ws.counters.release();
        ws.counters.ensure(sizeof(unsigned) * ncounters);
        rt_memset(ws.counters.p, 0, sizeof(unsigned) * ncounters, ctx->stream);
        ws.ncounters = ncounters;
    }
}

void vec_launch(mgvib200_ctx* ctx, ReduceWs& ws, VecParams P) {
    P.per_block = kVecPerBlock;
    P.nchunk = (P.N + kVecPerBlock - 1) / kVecPerBlock;
    const long long grid = P.nchunk * P.nvec;
    if (P.fin != FIN_NONE) {
        reduce_ws_ensure(ws, ctx, (size_t)grid, (size_t)(P.nvec > 1 ? P.nvec : 1));
        P.partials = ws.partials.as<double>();
        P.counters = ws.counters.as<unsigned>();
        if (!P.red_out) P.red_out = ws.red.as<double>();
    }
    ctx->prof_pre(1000 + P.op);
    rt_launch_pdl(k_vec, grid, kVecThreads, 512, ctx->stream, P);
    ctx->prof_post();
    ctx->launches++;
}

static VecParams vp_zero() {
    VecParams P;
    memset(&P, 0, sizeof P);
    P.nvec = 1;
    return P;
}

double vec_dot(mgvib200_ctx* ctx, ReduceWs& ws, const double* a, const double* b, long long N) {
    VecParams P = vp_zero();
    P.op = VOP_DOT; P.N = N; P.in = a; P.y = b; P.fin = FIN_TOTAL;
    vec_launch(ctx, ws, P);
    double h = 0.0;
    rt_d2h(&h, ws.red.p, sizeof(double), ctx->stream);
    rt_sync(ctx->stream);
    return h;
}

void vec_axpy(mgvib200_ctx* ctx, ReduceWs& ws, double* out, const double* x, double a, const double* y, long long N) {
    VecParams P = vp_zero();
    P.op = VOP_AXPY; P.N = N; P.in = x; P.y = y; P.a = a; P.out = out;
    vec_launch(ctx, ws, P);
}

void vec_axpy_inplace(mgvib200_ctx* ctx, ReduceWs& ws, double* x, double a, const double* y, long long N) {
    VecParams P = vp_zero();
    P.op = VOP_AXPY_INPLACE; P.N = N; P.out = x; P.y = y; P.a = a;
    vec_launch(ctx, ws, P);
}

void vec_build_samples(mgvib200_ctx* ctx, ReduceWs& ws, const double* c, const double* R, long long N, int n,
                       double* samples) {
    VecParams P = vp_zero();
    P.op = VOP_BUILD; P.N = N; P.nvec = n; P.in = c; P.y = R; P.out = samples;
    vec_launch(ctx, ws, P);
}

void vec_sum_chunks(mgvib200_ctx* ctx, ReduceWs& ws, const double* in, long long ld, int nsum, double scale,
                    double* out, long long N) {
    VecParams P = vp_zero();
    P.op = VOP_SUM_CHUNKS; P.N = N; P.in = in; P.in_ld = ld; P.nsum = nsum; P.scale = scale; P.out = out;
    vec_launch(ctx, ws, P);
}

void vec_canon_sum(mgvib200_ctx* ctx, ReduceWs& ws, const CanonShape& cs, const double* in, long long ld,
                   int terms_per_sample, double scale, double* out, long long N, DevBuf& gather) {
    const int per = (int)(cs.per_chunk * terms_per_sample);
    VecParams P = vp_zero();
    P.op = VOP_SUM_CANON; P.N = N; P.in = in; P.in_ld = ld; P.per = per;
    if (ctx->world <= 1) {
        // both levels in one kernel: same additions in the same order as the gathered form below
        P.nvec = 1; P.nsum = (int)cs.chunks; P.scale = scale; P.out = out;
        vec_launch(ctx, ws, P);
        return;
    }
    // chunk partials of this rank -> all-gather (no arithmetic in the collective) -> chunks in order
    gather.ensure(sizeof(double) * (size_t)N * (size_t)cs.chunks);
    double* G = gather.as<double>();
    double* mine = G + (size_t)ctx->rank * cs.chunks_local * N;
    P.nvec = (int)cs.chunks_local; P.nsum = 1; P.scale = 1.0; P.out = mine;
    vec_launch(ctx, ws, P);
    comm_allgather(ctx, mine, G, cs.chunks_local * N);
    VecParams Q = vp_zero();
    Q.op = VOP_SUM_CANON; Q.N = N; Q.in = G; Q.in_ld = N; Q.per = 1; Q.nvec = 1; Q.nsum = (int)cs.chunks;
    Q.scale = scale; Q.out = out;
    vec_launch(ctx, ws, Q);
}

double canon_scalar_sum(mgvib200_ctx* ctx, DevBuf& gather, const double* dev_local, int groups, long long count) {
    const long long per_rank = (long long)groups * count;
    const int W = ctx->world > 1 ? ctx->world : 1;
    std::vector<double> h((size_t)per_rank * W);
    if (W == 1) {
        rt_d2h(h.data(), dev_local, sizeof(double) * per_rank, ctx->stream);
    } else {
        gather.ensure(sizeof(double) * per_rank * W);
        double* G = gather.as<double>();
        double* mine = G + (size_t)ctx->rank * per_rank;
        if (mine != dev_local) rt_d2d(mine, dev_local, sizeof(double) * per_rank, ctx->stream);
        comm_allgather(ctx, mine, G, per_rank);
        rt_d2h(h.data(), G, sizeof(double) * per_rank * W, ctx->stream);
    }
    rt_sync(ctx->stream);
    // group after group, every group in global term order: the sequence of additions does not depend on W
    double tot = 0.0;
    for (int g = 0; g < groups; ++g)
        for (int r = 0; r < W; ++r)
            for (long long i = 0; i < count; ++i) tot += h[(size_t)r * per_rank + (size_t)g * count + i];
    return tot;
}

void vec_cg_init(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, const double* b, long long b_ld, double* x,
                 double* r, double* p, long long N, int nvec) {
    VecParams P = vp_zero();
    P.op = VOP_CG_INIT; P.N = N; P.nvec = nvec; P.in = b; P.in_ld = b_ld; P.x = x; P.r = r; P.p = p;
    P.fin = FIN_RHS; P.cg = cg;
    vec_launch(ctx, ws, P);
}

void vec_cg_update(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* x, double* r, const double* p,
                   const double* Ap, long long N, int nvec) {
    VecParams P = vp_zero();
    P.op = VOP_CG_UPDATE; P.N = N; P.nvec = nvec; P.x = x; P.r = r; P.p = p; P.Ap = Ap;
    P.fin = FIN_CG_RR; P.cg = cg;
    vec_launch(ctx, ws, P);
}

void vec_cg_update_r(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* r, const double* Ap, long long N, int nvec) {
    VecParams P = vp_zero();
    P.op = VOP_CG_UPDATE_R; P.N = N; P.nvec = nvec; P.r = r; P.Ap = Ap;
    P.fin = FIN_CG_RR; P.cg = cg;
    vec_launch(ctx, ws, P);
}

void vec_x_flush(mgvib200_ctx* ctx, ReduceWs& ws, const CgScalars& cg, double* x, const double* p, long long N, int nvec) {
    VecParams P = vp_zero();
    P.op = VOP_X_FLUSH; P.N = N; P.nvec = nvec; P.x = x; P.p = p; P.cg = cg;
    vec_launch(ctx, ws, P);
}

}  // namespace mgvi
