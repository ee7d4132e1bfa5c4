// Streaming vector kernels of the batched conjugate-gradient solve and of the Newton step
// (fused axpy pairs with the dot product of the same pass; src/residual_samplers.jl:79-81 via
// Krylov.jl cg!, src/newtoncg.jl:120 via IterativeSolvers cg_iterator!), plus sample assembly
// (src/mgvi_impl.jl:152-156).
#pragma once
#include "field_pass.h"

namespace mgvi {

enum {
    VOP_CG_UPDATE = 0,   // x += alpha p ; r -= alpha Ap ; acc = r.r          (per vector)
    VOP_CG_INIT = 1,     // r = b ; p = b ; x = 0 ; acc = b.b                 (per vector)
    VOP_SUM_CHUNKS = 2,  // out = scale * sum_c in[c]                          (nvec = 1)
    VOP_AXPY = 3,        // out = x + a * y                                    (nvec = 1)
    VOP_DOT = 4,         // acc = x . y                                        (nvec = 1)
    VOP_BUILD = 5,       // samples[:, 2i] = c + r_i ; samples[:, 2i+1] = c - r_i   (vector = i)
    VOP_AXPY_INPLACE = 6,// x += a * y                                         (nvec = 1)
    VOP_CG_PUPD = 7,     // p = r + beta p                                     (per vector; dense CG)
    VOP_CG_PAP = 8,      // acc = p . Ap                                       (per vector; dense CG)
    VOP_QUAD = 9,        // acc = sum x . (0.5 y - h)  over all vectors         (dense KL value)
    VOP_FLIP = 10,       // out[i, s] = in[N-1-i, s]                            (index reversal)
    VOP_COPY = 11,       // out = in                                            (per vector)
    VOP_TILE = 12,       // out[tile-major i] = in[natural index of i]           (2-D relayout, field_pass.h)
    VOP_CG_UPDATE_R = 14,// r -= alpha Ap ; acc = r.r   (per vector; the x half is deferred, field_pass.h: defer_x)
    VOP_X_FLUSH = 15,    // x += alpha p for vectors whose deferred update is still pending  (per vector)
    VOP_SUM_CANON = 13   // out[s] = scale * sum_{c < nsum} ( sum_{j < per} in[(s nsum + c) per + j] )
                         //   two-level sequential sum in the canonical sample order (engine.h: CanonShape)
};

struct VecParams {
    int op, nvec, fin;
    long long N, nchunk, per_block;
    double* x; double* r; const double* p; const double* Ap;   // CG vectors, ld = N
    const double* in; double* out; long long in_ld;
    const double* y;
    int nsum, per;
    int lg_wt, lg_rows, lg_cols;      // VOP_TILE geometry
    double a, scale;
    double* partials; unsigned* counters; double* red_out;
    CgScalars cg;
};

SIMT_DEVICE void vec_elem(const VecParams& P, int s, long long i, double alpha, double& acc) {
    const long long gi = (long long)s * P.N + i;
    switch (P.op) {
        case VOP_CG_UPDATE: {
            const double pv = P.p[gi], av = P.Ap[gi];
            P.x[gi] += alpha * pv;
            const double rv = P.r[gi] - alpha * av;
            P.r[gi] = rv;
            acc += rv * rv;
        } break;
        case VOP_CG_UPDATE_R: {
            const double rv = P.r[gi] - alpha * P.Ap[gi];
            P.r[gi] = rv;
            acc += rv * rv;
        } break;
        case VOP_X_FLUSH: P.x[gi] += alpha * P.p[gi]; break;
        case VOP_CG_INIT: {
            const double b = P.in[(long long)s * P.in_ld + i];
            P.r[gi] = b;
            const_cast<double*>(P.p)[gi] = b;
            P.x[gi] = 0.0;
            acc += b * b;
        } break;
        case VOP_SUM_CHUNKS: {
            double t = 0.0;
            for (int c = 0; c < P.nsum; ++c) t += P.in[(long long)c * P.in_ld + i];
            P.out[i] = P.scale * t;
        } break;
        case VOP_SUM_CANON: {
            double tot = 0.0;
            const double* src = P.in + ((long long)s * P.nsum * P.per) * P.in_ld + i;
            for (int c = 0; c < P.nsum; ++c) {
                double t = 0.0;
                for (int j = 0; j < P.per; ++j) t += src[((long long)c * P.per + j) * P.in_ld];
                tot += t;
            }
            P.out[gi] = P.scale * tot;
        } break;
        case VOP_AXPY: P.out[i] = P.in[i] + P.a * P.y[i]; break;
        case VOP_AXPY_INPLACE: P.out[i] += P.a * P.y[i]; break;
        case VOP_DOT: acc += P.in[i] * P.y[i]; break;
        case VOP_CG_PUPD: {
            const_cast<double*>(P.p)[gi] = P.r[gi] + P.cg.beta[s] * P.p[gi];
        } break;
        case VOP_CG_PAP: acc += P.p[gi] * P.Ap[gi]; break;
        case VOP_QUAD: acc += P.in[gi] * (0.5 * P.y[gi] - P.Ap[i]); break;
        case VOP_FLIP: P.out[gi] = P.in[(long long)s * P.N + (P.N - 1 - i)]; break;
        case VOP_COPY: P.out[gi] = P.in[(long long)s * P.in_ld + i]; break;
        case VOP_TILE: {
            const long long tile = i >> (P.lg_wt + P.lg_rows), r = (i >> P.lg_wt) & ((1LL << P.lg_rows) - 1);
            const long long k = (tile << P.lg_wt) + (i & ((1LL << P.lg_wt) - 1));
            P.out[gi] = P.in[(long long)s * P.N + (r << P.lg_cols) + k];
        } break;
        case VOP_BUILD: {
            const double c = P.in[i], rv = P.y[gi];
            P.out[(long long)(2 * s) * P.N + i] = c + rv;
            P.out[(long long)(2 * s + 1) * P.N + i] = c - rv;
        } break;
    }
}

// block -> (chunk, vector); a block covers `per_block` consecutive elements of one vector.
SIMT_DEVICE void vec_body(const VecParams& P, unsigned char* smem_raw) {
    double* red = reinterpret_cast<double*>(smem_raw);
    int* flag = reinterpret_cast<int*>(smem_raw + 256);
    const long long BID = SIMT_BID;
    const int s = (int)(BID % P.nvec);
    const long long chunk = BID / P.nvec;
    const bool cg_op = (P.op == VOP_CG_UPDATE || P.op == VOP_CG_PUPD || P.op == VOP_CG_PAP || P.op == VOP_CG_UPDATE_R);
    if (cg_op && !P.cg.active[s]) return;
    if (P.op == VOP_X_FLUSH && !P.cg.x_pending[s]) return;
    const double alpha = (P.op == VOP_CG_UPDATE || P.op == VOP_CG_UPDATE_R || P.op == VOP_X_FLUSH) ? P.cg.alpha[s] : 0.0;
    const long long lo = chunk * P.per_block;
    long long hi = lo + P.per_block;
    if (hi > P.N) hi = P.N;
    double acc = 0.0;
    if (P.op == VOP_CG_UPDATE && (P.N & 1) == 0) {
        // the streaming update of the CG iteration moves 48 bytes per element: 16-byte accesses,
        // all loads of a pair issued before its stores
        const long long base = (long long)s * P.N;
        const double2* __restrict__ p2 = reinterpret_cast<const double2*>(P.p + base);
        const double2* __restrict__ a2 = reinterpret_cast<const double2*>(P.Ap + base);
        double2* __restrict__ x2 = reinterpret_cast<double2*>(P.x + base);
        double2* __restrict__ r2 = reinterpret_cast<double2*>(P.r + base);
        // four independent 16-byte quadruples in flight per thread (the loop is a chain of memory round trips
        // when the vectors are L2-resident); the order of the additions into acc is that of the plain loop
        const long long i1 = hi >> 1, nt = SIMT_NTID;
        for (long long i0 = (lo >> 1) + SIMT_TID; i0 < i1; i0 += 4 * nt) {
            double2 pv[4], av[4], xv[4], rv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long i = i0 + u * nt;
                if (i < i1) { pv[u] = p2[i]; av[u] = a2[i]; xv[u] = x2[i]; rv[u] = r2[i]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long i = i0 + u * nt;
                if (i < i1) {
                    xv[u].x += alpha * pv[u].x; xv[u].y += alpha * pv[u].y;
                    rv[u].x -= alpha * av[u].x; rv[u].y -= alpha * av[u].y;
                    x2[i] = xv[u]; r2[i] = rv[u];
                    acc += rv[u].x * rv[u].x;
                    acc += rv[u].y * rv[u].y;
                }
            }
        }
    } else if (P.op == VOP_CG_UPDATE_R && (P.N & 1) == 0) {
        const long long base = (long long)s * P.N;
        const double2* __restrict__ a2 = reinterpret_cast<const double2*>(P.Ap + base);
        double2* __restrict__ r2 = reinterpret_cast<double2*>(P.r + base);
        // two streams in, one out: eight 16-byte pairs in flight per thread to keep as many bytes outstanding
        // as the four-stream update does
        const long long i1 = hi >> 1, nt = SIMT_NTID;
        for (long long i0 = (lo >> 1) + SIMT_TID; i0 < i1; i0 += 8 * nt) {
            double2 av[8], rv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long i = i0 + u * nt;
                if (i < i1) { av[u] = a2[i]; rv[u] = r2[i]; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long i = i0 + u * nt;
                if (i < i1) {
                    rv[u].x -= alpha * av[u].x; rv[u].y -= alpha * av[u].y;
                    r2[i] = rv[u];
                    acc += rv[u].x * rv[u].x;
                    acc += rv[u].y * rv[u].y;
                }
            }
        }
    } else {
        for (long long i = lo + SIMT_TID; i < hi; i += SIMT_NTID) vec_elem(P, s, i, alpha, acc);
    }
    if (P.fin == FIN_NONE) return;
    const double tot = seg_reduce(acc, SIMT_NTID, red);
    if (P.fin == FIN_TOTAL)
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, 0, BID, SIMT_NBID, tot, red,
                                   flag);
    else
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, s, chunk, P.nchunk, tot, red,
                                   flag);
}

}  // namespace mgvi
