// Fused axis-pass kernel of the correlated-field model:  load (prologue op) -> Hartley transform
// along one axis [-> pointwise weight -> Hartley transform again] -> store (epilogue op), with an
// optional scalar reduction finished by the last block ("last block done" pattern, deterministic
// order).  One kernel body serves every step of the path:
//
//   a5  metric-vector product  M v = v + c^2 A . H( w . H(A . v) )       (src/residual_samplers.jl:72)
//   a6  sampler right-hand side b = c A . H(q . n1) + eta                (src/residual_samplers.jl:77)
//   a4  linearisation weights  w = F g'^2, q = g' sqrt(F)                (src/residual_samplers.jl:4-8,
//                                                                         src/information.jl:12-18,74-78)
//   a9  KL value terms, a10 KL gradient input                            (src/mgvi_impl.jl:3-28, newtoncg.jl:100)
//
// For a d-dimensional field H is separable; the two axis-0 transforms of M v sit next to each
// other (the axis transforms commute), so they are fused into ONE pass with the weight applied
// in shared memory/registers: 2d-1 passes over HBM instead of 2d.
#pragma once
#include "fft_core.h"

namespace mgvi {

enum { GEO_CONTIG = 0, GEO_STRIDED = 1, GEO_PAIRVEC = 2 };
enum { PRO_LOAD = 0, PRO_AMP = 1, PRO_CG = 2, PRO_POINT = 3, PRO_RHS = 4 };
enum { EPI_STORE = 0, EPI_METRIC = 1, EPI_LINEARIZE = 2, EPI_RHS = 3, EPI_KL = 4, EPI_GRAD = 5 };
enum { FIN_NONE = 0, FIN_VEC_STORE = 1, FIN_TOTAL = 2, FIN_CG_PAP = 3, FIN_RHS = 4, FIN_CG_RR = 5 };
enum { LINK_IDENTITY = 0, LINK_EXP = 1 };
enum { LIKE_NORMAL = 0, LIKE_POISSON = 1 };
enum { POINT_PM = 0, POINT_PLUS = 1, POINT_CENTER = 2 };
enum { CG_KRYLOV = 0, CG_ITSOLVERS = 1 };

// Per-vector conjugate-gradient scalars, all resident in device memory so that a lock-step
// iteration over every residual sample is a fixed sequence of launches with no host round trip.
struct CgScalars {
    double* gamma;     // r.r  (Krylov.jl) / residual^2 (IterativeSolvers)
    double* alpha;
    double* beta;
    double* eps;       // stopping threshold per vector
    double* rnorm;     // current ||r||
    double* pAp;
    int* iters;
    int* active;       // 1 while the vector still iterates
    int* n_active;     // number of active vectors (single int)
    int* x_pending;    // deferred solution update (FieldPassParams::defer_x): 1 = x still lacks alpha * p of the last update
    unsigned* x_count; // ... blocks of the vector that have applied it (self-resetting)
    double atol, rtol;
    long long itmax;
    int flavor;        // CG_KRYLOV (sampler) or CG_ITSOLVERS (Newton inner CG, A.2)
};

struct FieldPassParams {
    // ---- geometry -------------------------------------------------------------------------
    int geo, n, log2n, L, G, lgL, lg_tpo;
    long long N, inner, units_per_vec, ntb;
    int nvec, nsgrp, nloop;
    const double2* tw;
    const double2* tw2;                       // paired transforms (fft_core.h: fft_line_post), lengths 4 * 8^s
    // ---- operands -------------------------------------------------------------------------
    const double* in;  long long in_ld;
    double* out;       long long out_ld;
    const double* vin; long long vin_ld;      // EPI_METRIC: the "+ v" operand
    const double* A;                          // amplitude (N)
    const double* w;                          // mid weight (N), double-transform passes
    const double* x;                          // PRO_POINT centre / EPI_GRAD prior term (N)
    const double* R;   long long R_ld;        // residuals (N x n)
    int point_mode;
    const double* q;                          // g' sqrt(F) (N), PRO_RHS
    const double* Z1;  long long Z1_ld, Z1_off, Z1_stride;
    const double* Z2;  long long Z2_ld;
    double* r; double* p; double* xout;       // CG vectors (leading dimension N)
    double* p_out;                            // PRO_CG: where p = r + beta p is written (null: in place)
    // Deferred solution update of the sampler CG (2-D / 3-D fields): x += alpha p moves 24 N bytes that need
    // nothing but alpha, so it rides in the fused axis-0 pass of the NEXT iteration -- a kernel bound by the L1
    // data pipe with two thirds of its HBM time idle -- instead of the streaming update kernel, which is bound
    // by HBM.  p is double-buffered (the next first pass writes p = r + beta p elsewhere), x_prev_p is the
    // previous iteration's p.
    int defer_x;
    const double* x_prev_p;
    const double* data;                       // observations (N)
    double* w_out; double* q_out; double* lam_out;   // EPI_LINEARIZE at the centre (out == null)
    double scale, offset, sigma, inv_sigma2;
    int link, likelihood, first_iter, want_grad;
    int mask_active;                          // skip vectors whose cg.active flag is 0 (CG iterations only)
    int prefetch_blocks;                      // > 0: hint the inputs of block (this + prefetch_blocks) into L2
    int prefetch_epi;                         // hint the epilogue / mid operands into L1 before the transform
    int t_lg_wt, t_lg_rows;                   // tile-major scratch (2-D): log2(reals per tile row) (0 = natural), log2(rows)
    int stage_vin;                            // contiguous EPI_METRIC: TMA-stage the '+ v' rows into shared memory at block start
    // ---- reduction ------------------------------------------------------------------------
    int fin;
    double* partials; unsigned* counters; double* red_out;
    CgScalars cg;
};

// ------------------------------------------------------------------------------------------
// pair helpers: (a, b) are the two real lines of a complex line.  `adj` (strided geometry) means
// b sits right after a in memory, so both move as one 16-byte access.
// ------------------------------------------------------------------------------------------
SIMT_DEVICE double2 ld_pair(const double* __restrict__ base, long long ia, long long ib, bool va, bool vb,
                            bool adj) {
    if (adj) return *reinterpret_cast<const double2*>(base + ia);
    return make_double2(va ? base[ia] : 0.0, vb ? base[ib] : 0.0);
}
SIMT_DEVICE double2 ldg_pair(const double* __restrict__ base, long long ia, long long ib, bool va, bool vb,
                             bool adj) {
    if (adj) return SIMT_LDG(reinterpret_cast<const double2*>(base + ia));
    return make_double2(va ? SIMT_LDG(base + ia) : 0.0, vb ? SIMT_LDG(base + ib) : 0.0);
}
SIMT_DEVICE void st_pair(double* __restrict__ base, long long ia, long long ib, bool va, bool vb, bool adj,
                         double2 v) {
    if (adj) { *reinterpret_cast<double2*>(base + ia) = v; return; }
    if (va) base[ia] = v.x;
    if (vb) base[ib] = v.y;
}

// ------------------------------------------------------------------------------------------
// pointwise model pieces (Appendix B of SURVEY.md; src/information.jl:12-18,74-78)
// ------------------------------------------------------------------------------------------
SIMT_DEVICE void lin_point(const FieldPassParams& P, double sig, double& w, double& q, double& lam) {
    lam = (P.link == LINK_EXP) ? exp(sig) : sig;
    const double gp = (P.link == LINK_EXP) ? lam : 1.0;
    const double F = (P.likelihood == LIKE_POISSON) ? 1.0 / lam : P.inv_sigma2;
    w = F * gp * gp;
    q = gp * sqrt(F);
}

// variable part of -loglike at one datum, and g' * dloglike/dlambda
SIMT_DEVICE void kl_point(const FieldPassParams& P, double sig, double d, double& nl, double& qv) {
    const double lam = (P.link == LINK_EXP) ? exp(sig) : sig;
    const double gp = (P.link == LINK_EXP) ? lam : 1.0;
    if (P.likelihood == LIKE_POISSON) {
        const double loglam = (P.link == LINK_EXP) ? sig : log(lam);
        nl = lam - d * loglam;
        qv = (P.link == LINK_EXP) ? (d - lam) : (d / lam - 1.0);
    } else {
        const double z = (d - lam) / P.sigma;
        nl = 0.5 * z * z;
        qv = gp * ((d - lam) * P.inv_sigma2);
    }
}

SIMT_DEVICE void point_decode(int mode, int s, int& i, double& sign) {
    if (mode == POINT_PM) { i = s >> 1; sign = (s & 1) ? -1.0 : 1.0; }
    else if (mode == POINT_PLUS) { i = s; sign = 1.0; }
    else { i = 0; sign = 0.0; }
}

template <int PRO>
SIMT_DEVICE double2 pro_pair(const FieldPassParams& P, int sa, int sb, long long pa, long long pb, bool va,
                             bool vb, bool adj, double& acc_a, double& acc_b) {
    if (PRO == PRO_LOAD) {
        return ld_pair(P.in, sa * P.in_ld + pa, sb * P.in_ld + pb, va, vb, adj);
    } else if (PRO == PRO_AMP) {
        double2 v = ld_pair(P.in, sa * P.in_ld + pa, sb * P.in_ld + pb, va, vb, adj);
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        return make_double2(a.x * v.x, a.y * v.y);
    } else if (PRO == PRO_CG) {
        const long long ia = sa * P.N + pa, ib = sb * P.N + pb;
        double2 pp = ld_pair(P.p, ia, ib, va, vb, adj);
        if (!P.first_iter) {
            double2 rr = ld_pair(P.r, ia, ib, va, vb, adj);
            const double ba = va ? P.cg.beta[sa] : 0.0, bb = vb ? P.cg.beta[sb] : 0.0;
            pp = make_double2(rr.x + ba * pp.x, rr.y + bb * pp.y);
            st_pair(P.p_out ? P.p_out : P.p, ia, ib, va, vb, adj, pp);
        }
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        return make_double2(a.x * pp.x, a.y * pp.y);
    } else if (PRO == PRO_POINT) {
        double2 xv = ldg_pair(P.x, pa, pb, va, vb, adj);
        if (P.point_mode != POINT_CENTER) {
            int i_a, i_b; double sg_a, sg_b;
            point_decode(P.point_mode, sa, i_a, sg_a);
            point_decode(P.point_mode, sb, i_b, sg_b);
            double2 rv = ld_pair(P.R, i_a * P.R_ld + pa, i_b * P.R_ld + pb, va, vb, adj);
            xv = make_double2(xv.x + sg_a * rv.x, xv.y + sg_b * rv.y);
        }
        if (va) acc_a += 0.5 * xv.x * xv.x;      // prior term 1/2 |p|^2 (src/mgvi_impl.jl:3-7)
        if (vb) acc_b += 0.5 * xv.y * xv.y;
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        return make_double2(a.x * xv.x, a.y * xv.y);
    } else {  // PRO_RHS
        double2 qv = ldg_pair(P.q, pa, pb, va, vb, adj);
        const long long za = sa * P.Z1_ld + P.Z1_off + P.Z1_stride * pa;
        const long long zb = sb * P.Z1_ld + P.Z1_off + P.Z1_stride * pb;
        double2 z = ld_pair(P.Z1, za, zb, va, vb, adj && P.Z1_stride == 1);
        return make_double2(qv.x * z.x, qv.y * z.y);
    }
}

// Per-point fields (the gradient input g' dl/dlambda of EPI_KL, the metric weight of EPI_LINEARIZE away
// from the centre) are STORED per vector; sums over residual samples are formed afterwards in the
// canonical order (vec_ops.h: VOP_SUM_CANON) so that they do not depend on the sharding.
template <int EPI>
SIMT_DEVICE void epi_pair(const FieldPassParams& P, int sa, int sb, long long pa, long long pb, bool va,
                          bool vb, bool adj, double2 val, double& acc_a, double& acc_b) {
    if (EPI == EPI_STORE) {
        st_pair(P.out, sa * P.out_ld + pa, sb * P.out_ld + pb, va, vb, adj, val);
    } else if (EPI == EPI_METRIC) {
        double2 vi = ld_pair(P.vin, sa * P.vin_ld + pa, sb * P.vin_ld + pb, va, vb, adj);
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        double2 o = make_double2(vi.x + P.scale * (a.x * val.x), vi.y + P.scale * (a.y * val.y));
        st_pair(P.out, sa * P.out_ld + pa, sb * P.out_ld + pb, va, vb, adj, o);
        if (va) acc_a += vi.x * o.x;
        if (vb) acc_b += vi.y * o.y;
    } else if (EPI == EPI_LINEARIZE) {
        double wa = 0, qa = 0, la = 0, wb = 0, qb = 0, lb = 0;
        if (va) lin_point(P, P.scale * val.x + P.offset, wa, qa, la);
        if (vb) lin_point(P, P.scale * val.y + P.offset, wb, qb, lb);
        if (P.out) {
            st_pair(P.out, sa * P.out_ld + pa, sb * P.out_ld + pb, va, vb, adj, make_double2(wa, wb));
        } else {
            st_pair(P.w_out, pa, pb, va, vb, adj, make_double2(wa, wb));
            st_pair(P.q_out, pa, pb, va, vb, adj, make_double2(qa, qb));
            if (P.lam_out) st_pair(P.lam_out, pa, pb, va, vb, adj, make_double2(la, lb));
        }
    } else if (EPI == EPI_RHS) {
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        double2 e = ld_pair(P.Z2, sa * P.Z2_ld + pa, sb * P.Z2_ld + pb, va, vb, adj);
        double2 b = make_double2(P.scale * (a.x * val.x) + e.x, P.scale * (a.y * val.y) + e.y);
        const long long ia = sa * P.N + pa, ib = sb * P.N + pb;
        st_pair(P.r, ia, ib, va, vb, adj, b);
        st_pair(P.p, ia, ib, va, vb, adj, b);
        st_pair(P.xout, ia, ib, va, vb, adj, make_double2(0.0, 0.0));
        if (va) acc_a += b.x * b.x;
        if (vb) acc_b += b.y * b.y;
    } else if (EPI == EPI_KL) {
        double2 d = ldg_pair(P.data, pa, pb, va, vb, adj);
        double na = 0, qa = 0, nb = 0, qb = 0;
        if (va) kl_point(P, P.scale * val.x + P.offset, d.x, na, qa);
        if (vb) kl_point(P, P.scale * val.y + P.offset, d.y, nb, qb);
        acc_a += na; acc_b += nb;
        if (P.want_grad) st_pair(P.out, sa * P.out_ld + pa, sb * P.out_ld + pb, va, vb, adj, make_double2(qa, qb));
    } else {  // EPI_GRAD: grad = x - scale * A . val   (scale carries c / (2n) and the 1/2 factors)
        double2 xv = ldg_pair(P.x, pa, pb, va, vb, adj);
        double2 a = ldg_pair(P.A, pa, pb, va, vb, adj);
        st_pair(P.out, pa, pb, va, vb, adj,
                make_double2(xv.x - P.scale * (a.x * val.x), xv.y - P.scale * (a.y * val.y)));
    }
}

// ------------------------------------------------------------------------------------------
// Whole-thread versions of the two hot load/store steps of the CG iteration (contiguous and
// strided geometry, both halves of the pair valid).  The per-element forms above interleave loads
// with stores to arrays the compiler must assume may alias, which serialises the memory
// pipeline (and re-reads beta for every element); here every batch issues all its loads first,
// through __restrict__ pointers, then computes, then stores.
// ------------------------------------------------------------------------------------------
// Slot 4 h + j of a thread's 8 elements sits at offset off + h * d1 + j * stepj: the standard arrangement
// (k = lt + m n/8) is d1 = 4 step, stepj = step; the paired arrangement (fft_core.h) is d1 = (jB - jA) * kstride,
// stepj = 2 step.
template <bool ADJ>
SIMT_DEVICE void pro_cg_block(double2 (&v)[8], const double* __restrict__ p, double* pout, const double* __restrict__ r,
                              const double* __restrict__ A, unsigned off_a, unsigned off_b, int d1, unsigned stepj,
                              long long vec_off, double beta, bool first_iter) {
    // p = r + beta p (skipped on the first iteration, where p = r already) ; v = A . p
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double2 pp[4], rr[4], aa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned oa = off_a + h * d1 + j * stepj, ob = off_b + h * d1 + j * stepj;
            pp[j] = ld_pair(p, vec_off + oa, vec_off + ob, true, true, ADJ);
            aa[j] = ldg_pair(A, oa, ob, true, true, ADJ);
            if (!first_iter) rr[j] = ld_pair(r, vec_off + oa, vec_off + ob, true, true, ADJ);
        }
        if (!first_iter) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned oa = off_a + h * d1 + j * stepj, ob = off_b + h * d1 + j * stepj;
                pp[j] = make_double2(rr[j].x + beta * pp[j].x, rr[j].y + beta * pp[j].y);
                st_pair(pout, vec_off + oa, vec_off + ob, true, true, ADJ, pp[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[4 * h + j] = make_double2(aa[j].x * pp[j].x, aa[j].y * pp[j].y);
    }
}

// same as epi_metric_block below, with the '+ v' rows already in shared memory (row a at stg[0..n),
// row b at stg[n..2n)); contiguous geometry only
SIMT_DEVICE void epi_metric_block_staged(const double2 (&v)[8], double* __restrict__ out, const double* stg,
                                         const double* __restrict__ A, unsigned off_a, unsigned off_b, unsigned step,
                                         long long out_off, int lt, int n, double scale, double& acc) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double2 vi[4], aa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = lt + (4 * h + j) * (int)step;
            const unsigned oa = off_a + (4 * h + j) * step, ob = off_b + (4 * h + j) * step;
            vi[j] = make_double2(stg[k], stg[n + k]);
            aa[j] = ld_pair(A, oa, ob, true, true, false);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned oa = off_a + (4 * h + j) * step, ob = off_b + (4 * h + j) * step;
            const double2 val = v[4 * h + j];
            const double2 o = make_double2(vi[j].x + scale * (aa[j].x * val.x), vi[j].y + scale * (aa[j].y * val.y));
            st_pair(out, out_off + oa, out_off + ob, true, true, false, o);
            acc += vi[j].x * o.x;
            acc += vi[j].y * o.y;
        }
    }
}

template <bool ADJ>
SIMT_DEVICE void epi_metric_block(const double2 (&v)[8], double* __restrict__ out, const double* __restrict__ vin,
                                  const double* __restrict__ A, unsigned off_a, unsigned off_b, int d1, unsigned stepj,
                                  long long out_off, long long vin_off, double scale, double& acc) {
    // out = vin + scale * A . val ; acc += vin . out
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double2 vi[4], aa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned oa = off_a + h * d1 + j * stepj, ob = off_b + h * d1 + j * stepj;
            vi[j] = ld_pair(vin, vin_off + oa, vin_off + ob, true, true, ADJ);
            aa[j] = ld_pair(A, oa, ob, true, true, ADJ);       // plain load: must not be hoisted above the transform
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned oa = off_a + h * d1 + j * stepj, ob = off_b + h * d1 + j * stepj;
            const double2 val = v[4 * h + j];
            const double2 o = make_double2(vi[j].x + scale * (aa[j].x * val.x), vi[j].y + scale * (aa[j].y * val.y));
            st_pair(out, out_off + oa, out_off + ob, true, true, ADJ, o);
            acc += vi[j].x * o.x;
            acc += vi[j].y * o.y;
        }
    }
}

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
// Sum `v` over segments of `seg` consecutive threads (seg a power of two dividing the block size,
// block size a multiple of 32).  Result valid in the first thread of each segment.  `red` is
// >= 32 doubles of shared scratch.  All threads of the block must call.
SIMT_DEVICE double seg_reduce(double v, int seg, double* red) {
    const int tid = SIMT_TID;
    if (seg <= 32) {
        for (int m = seg >> 1; m >= 1; m >>= 1) v += SIMT_SHFL_XOR_F64(v, m);
        return v;
    }
    for (int m = 16; m >= 1; m >>= 1) v += SIMT_SHFL_XOR_F64(v, m);
    SIMT_SYNC();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    SIMT_SYNC();
    double tot = 0.0;
    if ((tid & (seg - 1)) == 0) {
        const int w0 = tid >> 5, nw = seg >> 5;
        for (int i = 0; i < nw; ++i) tot += red[w0 + i];
    }
    return tot;
}

SIMT_DEVICE void cg_deactivate(const CgScalars& cg, int s) {
    cg.active[s] = 0;
    SIMT_ATOMIC_ADD_I32(cg.n_active, -1);
}

// Scalar tail of a reduction, executed by exactly one thread per vector.
SIMT_DEVICE void finalize_vec(int fin, const CgScalars& cg, double* red_out, int s, double tot) {
    if (fin == FIN_VEC_STORE) {
        red_out[s] = tot;
    } else if (fin == FIN_TOTAL) {
        red_out[0] = tot;
    } else if (fin == FIN_CG_PAP) {
        cg.pAp[s] = tot;
        cg.alpha[s] = cg.gamma[s] / tot;
    } else if (fin == FIN_RHS) {
        // Krylov.jl cg! start-up: gamma = b.b, eps = atol + rtol*||b||
        const double rn = sqrt(tot);
        cg.gamma[s] = tot;
        cg.rnorm[s] = rn;
        cg.iters[s] = 0;
        cg.beta[s] = 0.0;
        if (cg.flavor == CG_KRYLOV) {
            const double eps = cg.atol + cg.rtol * rn;
            cg.eps[s] = eps;
            const int act = (tot != 0.0) && !(rn <= eps) && (cg.itmax > 0);
            cg.active[s] = act;
            if (act) SIMT_ATOMIC_ADD_I32(cg.n_active, 1);
        } else {
            // IterativeSolvers cg_iterator! start-up: tol = max(reltol*res, abstol)
            const double tol = fmax(cg.rtol * rn, cg.atol);
            cg.eps[s] = tol;
            cg.gamma[s] = rn * rn;
            const int act = !(rn <= tol) && (cg.itmax > 0);
            cg.active[s] = act;
            if (act) SIMT_ATOMIC_ADD_I32(cg.n_active, 1);
        }
    } else if (fin == FIN_CG_RR) {
        if (cg.x_pending) cg.x_pending[s] = 1;       // deferred solution update: x now lacks alpha * p of this update
        const double rn = sqrt(tot);
        cg.rnorm[s] = rn;
        const int it = cg.iters[s] + 1;
        cg.iters[s] = it;
        if (cg.flavor == CG_KRYLOV) {
            const bool solved = (rn <= cg.eps[s]) || (rn + 1.0 <= 1.0);
            cg.beta[s] = tot / cg.gamma[s];
            cg.gamma[s] = tot;
            if (solved || it >= cg.itmax) cg_deactivate(cg, s);
        } else {
            const double g_new = rn * rn;
            cg.beta[s] = g_new / cg.gamma[s];
            cg.gamma[s] = g_new;
            if (rn <= cg.eps[s] || it >= cg.itmax) cg_deactivate(cg, s);
        }
    }
}

// Cross-block part of a reduction: thread 0 of the block publishes its partial; the last block to
// arrive sums all partials of the vector in a fixed order and runs the scalar tail.
// slot = index of this block among the `count` blocks contributing to vector s.
SIMT_DEVICE void block_publish_and_finalize(int fin, const CgScalars& cg, double* red_out, double* partials,
                                            unsigned* counters, int s, long long slot, long long count,
                                            double block_total, double* red, int* flag) {
    const int tid = SIMT_TID;
    if (tid == 0) {
        partials[(long long)s * count + slot] = block_total;
        SIMT_FENCE();
        const unsigned c = SIMT_ATOMIC_INC_U32(&counters[s]);
        *flag = (c == (unsigned)(count - 1));
    }
    SIMT_SYNC();
    if (*flag) {
        SIMT_FENCE();
        double a = 0.0;
        for (long long i = tid; i < count; i += SIMT_NTID) a += SIMT_LDCG(&partials[(long long)s * count + i]);
        const double tot = seg_reduce(a, SIMT_NTID, red);
        if (tid == 0) {
            counters[s] = 0;
            finalize_vec(fin, cg, red_out, s, tot);
        }
    }
}

// shared memory: [0,256) reduction scratch, [256,260) flag, FFT buffers from byte 512
static const int kSmemHeader = 512;

SIMT_HD long long field_pass_smem_bytes(int n, int L, int G) {
    return kSmemHeader + (long long)G * L * n * 16;
}
// with the TMA-staged '+ v' operand: one row pair (2n doubles) per unit behind the FFT buffers
SIMT_HD long long field_pass_smem_bytes_staged(int n, int L, int G) {
    return field_pass_smem_bytes(n, L, G) + (long long)G * 2 * n * 8;
}

// Deferred solution update, one block's share: the slab [bt, bt + 1) * N / ntb of vector s; all loads of a round
// first, through __restrict__ pointers.  The last block of the vector to finish clears the flag (every block --
// and, by the caller's barrier, every thread of it -- has read it by then).
SIMT_DEVICE void deferred_x_update(const FieldPassParams& P, int s, long long bt, int tid) {
    const double alpha = P.cg.alpha[s];
    const long long slab2 = (P.N / P.ntb) >> 1;                 // double2 elements per block
    const long long o2 = (((long long)s * P.N) >> 1) + bt * slab2;
    double2* __restrict__ x2 = reinterpret_cast<double2*>(P.xout) + o2;
    const double2* __restrict__ p2 = reinterpret_cast<const double2*>(P.x_prev_p) + o2;
    const int nt = SIMT_NTID;
#ifndef MGVI_DEFER_UNROLL
#define MGVI_DEFER_UNROLL 4
#endif
    for (long long i0 = tid; i0 < slab2; i0 += MGVI_DEFER_UNROLL * nt) {
        double2 xv[MGVI_DEFER_UNROLL], pv[MGVI_DEFER_UNROLL];
#pragma unroll
        for (int u = 0; u < MGVI_DEFER_UNROLL; ++u)
            if (i0 + u * nt < slab2) { xv[u] = x2[i0 + u * nt]; pv[u] = p2[i0 + u * nt]; }
#pragma unroll
        for (int u = 0; u < MGVI_DEFER_UNROLL; ++u)
            if (i0 + u * nt < slab2) {
                xv[u].x += alpha * pv[u].x; xv[u].y += alpha * pv[u].y;
                x2[i0 + u * nt] = xv[u];
            }
    }
    if (tid == 0) {
        SIMT_FENCE();
        const unsigned c = SIMT_ATOMIC_INC_U32(&P.cg.x_count[s]);
        if (c == (unsigned)(P.ntb - 1)) { P.cg.x_count[s] = 0; P.cg.x_pending[s] = 0; }
    }
}

// GEO is a compile-time geometry class:
//   GEO_CONTIG : lines along the last (contiguous) axis; a unit = two neighbouring rows of one vector
//   GEO_STRIDED: lines along an outer axis; a unit = a tile of 2L neighbouring columns
//   GEO_PAIRVEC: 1-D fields; a unit = two vectors
// All sizes are powers of two, so thread/unit decoding is shifts and masks.  Element offsets
// inside a vector are 32-bit (N < 2^31, checked on the host).
// LG > 0 / LGL >= 0 fix log2(n) / log2(L) at compile time; LOOP = false drops the loop over vectors
// (nloop == 1): with everything known the shared-memory swizzle offsets fold into immediates and
// nothing loop-invariant is left for the compiler to hoist and spill.
// PF selects the transform scheme (fft_core.h): 0 = fft_line_forward + partner read; the paired schemes
// (hot kernels of the 2048- / 256-point axes only) drop one shared-memory exchange per transform:
//   1 = operands loaded in the paired arrangement, combine on the INPUT, standard arrangement out (first pass)
//   2 = standard arrangement in, combine on the OUTPUT in registers, epilogue in the paired arrangement (last pass)
//   3 = DOUBLE: scheme 2, mid weight in the paired arrangement, scheme 1 (fused axis-0 pass)
template <int GEO, int PRO, int EPI, bool DOUBLE, int LG, int LGL, bool LOOP, int PF = 0>
SIMT_DEVICE void field_pass_body(const FieldPassParams& P, unsigned char* smem_raw) {
    const int log2n = LG ? LG : P.log2n, lgL = (LGL >= 0) ? LGL : P.lgL;
    const int n8 = 1 << (log2n - 3);
    const int tid = SIMT_TID;
    const int lg_upt = log2n - 3 + lgL;          // threads per unit = 2^lg_upt
    const int g = tid >> lg_upt;
    const int tu = tid & ((1 << lg_upt) - 1);
    const int l = tu & ((1 << lgL) - 1), lt = tu >> lgL;
    double* red = reinterpret_cast<double*>(smem_raw);
    int* flag = reinterpret_cast<int*>(smem_raw + 256);
    LineView lv;
    lv.sm = reinterpret_cast<double2*>(smem_raw + kSmemHeader) + ((long long)g << (lgL + log2n));
    lv.lgL = lgL;
    lv.l = l;

    const bool pairvec = (GEO == GEO_PAIRVEC);
    const bool adj = (GEO == GEO_STRIDED);
    long long bt;
    int sgrp;
    if (pairvec) { bt = SIMT_BID; sgrp = 0; }
    else { bt = SIMT_BID / P.nsgrp; sgrp = (int)(SIMT_BID - bt * P.nsgrp); }
    const bool use_active = (P.mask_active != 0);
    int x_pend = 0;
    if (DOUBLE && GEO == GEO_STRIDED && P.defer_x) {
        // deferred x += alpha p of the previous update (see FieldPassParams::defer_x).  The vector may have
        // converged in that update: then this is all the block does.
        x_pend = P.cg.x_pending[sgrp];              // uniform for the whole block
        const int act = P.cg.active[sgrp];
        SIMT_SYNC();        // every thread has read the flags before thread 0 may clear x_pending (deferred_x_update)
        if (!x_pend && !act) return;
        if (x_pend && (!act || P.defer_x == 1)) { deferred_x_update(P, sgrp, bt, tid); x_pend = 0; }
        if (!act) return;
        // defer_x == 2: the update runs AFTER the transforms of this block; its operands are pulled into L2 now
        if (x_pend && tid == 0) {
            const long long slab = P.N / P.ntb, o = (long long)sgrp * P.N + bt * slab;
            SIMT_PREFETCH_BULK_L2(P.xout + o, (unsigned)(slab * 8));
            SIMT_PREFETCH_BULK_L2(P.x_prev_p + o, (unsigned)(slab * 8));
        }
    } else if (use_active && !pairvec && P.nloop == 1) {
        if (!P.cg.active[sgrp]) return;    // uniform for the whole block
    }

    // Software prefetch: blocks are dispatched in index order, so the block `prefetch_blocks` ahead
    // starts about one residency later; pulling its per-sample input rows into L2 now takes the DRAM
    // latency off that block's load phase (the shared operands A and w stay in L2 by themselves).
    if (!pairvec && P.prefetch_blocks > 0 && P.nloop == 1) {
        const long long b2 = SIMT_BID + P.prefetch_blocks;
        if (b2 < SIMT_NBID) {
            const long long bt2 = b2 / P.nsgrp;
            const long long s2 = b2 - bt2 * P.nsgrp;
            const long long u2 = bt2 * P.G + g;
            if (u2 < P.units_per_vec) {
                if (GEO == GEO_CONTIG) {
                    // a row pair is one contiguous chunk of 2n doubles
                    const long long o2 = s2 * P.N + ((2 * u2) << log2n);
                    const unsigned bytes = 16u << log2n;
                    if (tu == 0) {
                        if (PRO == PRO_CG) { SIMT_PREFETCH_BULK_L2(P.p + o2, bytes); if (!P.first_iter) SIMT_PREFETCH_BULK_L2(P.r + o2, bytes); }
                        if (PRO == PRO_LOAD || PRO == PRO_AMP) SIMT_PREFETCH_BULK_L2(P.in + s2 * P.in_ld + ((2 * u2) << log2n), bytes);
                        if (EPI == EPI_METRIC) SIMT_PREFETCH_BULK_L2(P.vin + s2 * P.vin_ld + ((2 * u2) << log2n), bytes);
                    }
                } else if (PRO == PRO_LOAD) {
                    const long long o2 = u2 >> P.lg_tpo, cb2 = u2 & ((1LL << P.lg_tpo) - 1);
                    const double* q = P.in + s2 * P.in_ld + ((o2 << log2n) + lt) * P.inner + (cb2 << (lgL + 1)) + 2 * l;
                    // one hint per row segment (the L = 2^lgL lanes of a row share a line)
                    if (l == 0) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) SIMT_PREFETCH_L2(q + (long long)m * n8 * P.inner);
                    }
                }
            }
        }
    }

    double acc_a = 0.0, acc_b = 0.0;
    // element offsets of my 8 elements: off_a + m * step (and off_b + m * step); kstr = offset of one step
    // along the transform axis (step = n/8 * kstr)
    unsigned off_a = 0, off_b = 0, step = (unsigned)n8, kstr = 1;
    bool unit_ok = true;
    if (!pairvec) {
        const long long u = bt * P.G + g;
        unit_ok = u < P.units_per_vec;
        if (GEO == GEO_CONTIG) {
            off_a = (unsigned)((2 * u) << log2n) + (unsigned)lt;
            off_b = off_a + (1u << log2n);
            step = (unsigned)n8;
        } else if (P.t_lg_wt) {
            // tile-major scratch (2-D, narrow tiles): tile u is one contiguous block [row][2L reals],
            // so the 2L-real row segments of this pass are back to back instead of one per 128-byte line
            off_a = (unsigned)((u << (lgL + 1 + log2n)) + ((long long)lt << (lgL + 1)) + 2 * l);
            off_b = off_a + 1;
            kstr = 1u << (lgL + 1);
            step = (unsigned)(n8 << (lgL + 1));
        } else {
            const long long o = u >> P.lg_tpo, cb = u & ((1LL << P.lg_tpo) - 1);
            off_a = (unsigned)(((o << log2n) + lt) * P.inner + (cb << (lgL + 1)) + 2 * l);
            off_b = off_a + 1;
            kstr = (unsigned)P.inner;
            step = (unsigned)(n8 * P.inner);
        }
    } else {
        off_a = off_b = (unsigned)lt;
    }
    // paired arrangement (PF != 0): slot 4 h + j at off + h * pd1 + j * 2 step
    int pd1 = 0;
    if (PF != 0) pd1 = (paired_jB<(LG > 0 ? LG : 8)>(lt) - lt) * (int)kstr;
    // row passes next to a tile-major column pass: the scratch operand (the output of the first
    // pass, the input of the last) is addressed tile-major -- element (row r, column k) lives at
    // (k / wt) * wt * rows + r * wt + k % wt; rows 2u and 2u+1 of a unit are wt reals apart
    unsigned toff_a = off_a, toff_b = off_b, tstep = step;
    if (GEO == GEO_CONTIG && P.t_lg_wt) {
        const long long u = bt * P.G + g;
        const int lw = P.t_lg_wt;
        toff_a = (unsigned)((((long long)(lt >> lw)) << (lw + P.t_lg_rows)) + ((2 * u) << lw) + (lt & ((1 << lw) - 1)));
        toff_b = toff_a + (1u << lw);
        tstep = (unsigned)(((long long)(n8 >> lw)) << (lw + P.t_lg_rows));
    }
    int sa = 0, sb = 0;
    bool va = false, vb = false;
    const int nloop = LOOP ? P.nloop : 1;
    for (int it = 0; it < nloop; ++it) {
        if (pairvec) {
            const long long pair = (bt * nloop + it) * P.G + g;
            sa = (int)(2 * pair); sb = sa + 1;
            va = sa < P.nvec; vb = sb < P.nvec;
        } else {
            sa = sb = sgrp * nloop + it;
            va = vb = unit_ok && sa < P.nvec;
        }
        if (use_active) {
            va = va && P.cg.active[sa];
            vb = vb && P.cg.active[sb];
        }
        const bool valid = va || vb;
        double2 v[8];
        if (pairvec) {
            // the two halves of a pair are different vectors: per-half predicates
#pragma unroll
            for (int m = 0; m < 8; ++m)
                v[m] = valid ? pro_pair<PRO>(P, sa, sb, off_a + m * step, off_b + m * step, va, vb, false, acc_a, acc_b)
                             : make_double2(0.0, 0.0);
        } else if (valid) {
            if (PRO == PRO_CG) {
                if (PF == 1)
                    pro_cg_block<GEO == GEO_STRIDED>(v, P.p, P.p_out ? P.p_out : P.p, P.r, P.A, off_a, off_b, pd1, 2 * step, (long long)sa * P.N,
                                                     P.first_iter ? 0.0 : P.cg.beta[sa], P.first_iter != 0);
                else
                    pro_cg_block<GEO == GEO_STRIDED>(v, P.p, P.p_out ? P.p_out : P.p, P.r, P.A, off_a, off_b, 4 * (int)step, step, (long long)sa * P.N,
                                                     P.first_iter ? 0.0 : P.cg.beta[sa], P.first_iter != 0);
            } else if (PRO == PRO_LOAD && GEO == GEO_CONTIG && P.t_lg_wt && n8 >= 32) {
                // tile-major scratch, row pass: a unit's two rows sit wt reals apart, so lane pairs
                // fetch 16 bytes of ONE row each (even lane: row a, columns k, k+1; odd lane: row b,
                // columns k-1, k) and swap halves -- half the load instructions, whole 32-byte sectors
                const int q = lt & 1;
                const double* src = P.in + (long long)sa * P.in_ld;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const unsigned o = q ? (toff_b + m * tstep - 1) : (toff_a + m * tstep);
                    const double2 d2 = *reinterpret_cast<const double2*>(src + o);
                    const double rcv = SIMT_SHFL_XOR_F64(q ? d2.x : d2.y, 1);
                    v[m] = q ? make_double2(rcv, d2.y) : make_double2(d2.x, rcv);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    v[m] = (PRO == PRO_LOAD) ? pro_pair<PRO>(P, sa, sb, toff_a + m * tstep, toff_b + m * tstep, true, true, adj, acc_a, acc_b)
                                             : pro_pair<PRO>(P, sa, sb, off_a + m * step, off_b + m * step, true, true, adj, acc_a, acc_b);
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = make_double2(0.0, 0.0);
        }
        // TMA staging of the '+ v' rows (contiguous last pass): the row pair is one contiguous chunk of 2n
        // doubles; one thread arms an mbarrier and issues a bulk copy (UBLKCP) that lands while the
        // transform runs -- the epilogue then reads shared memory instead of stalling on a second
        // global load phase
        double* stg = nullptr;
        unsigned long long* mbar = nullptr;
#ifdef MGVI_STAGE_VIN
        const bool staged = (EPI == EPI_METRIC) && (GEO == GEO_CONTIG) && !LOOP && (P.stage_vin != 0);
#else
        // measured on B200 (profiles/README.md, v6): 0.771 ms staged vs 0.690 ms unstaged for the C3 last
        // pass -- the extra 32 KB of shared memory per block shrinks L1 under the twiddle / A reads and
        // the early copy competes with the primary loads; compiled out unless -DMGVI_STAGE_VIN
        const bool staged = false;
#endif
        if (staged) {
            stg = reinterpret_cast<double*>(smem_raw + kSmemHeader + ((long long)P.G << (lgL + log2n + 4))) + ((long long)g << (log2n + 1));
            mbar = reinterpret_cast<unsigned long long*>(smem_raw + 264) + g;
            if (tu == 0 && valid) {
                SIMT_MBAR_INIT(mbar, 1);
                SIMT_BULK_G2S(stg, P.vin + (long long)sa * P.vin_ld + (off_a - lt), 16u << log2n, mbar);
            }
        }
        // operands only needed AFTER the transform (the "+ v" term, the mid weight): start them towards
        // L1 now -- a hint costs no register and no scoreboard entry, the later load then hits
        if (P.prefetch_epi && valid && !pairvec) {
            if (EPI == EPI_METRIC) {
                const double* q = P.vin + (long long)sa * P.vin_ld;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    SIMT_PREFETCH_L1(q + off_a + m * step);
                    if (GEO != GEO_STRIDED) SIMT_PREFETCH_L1(q + off_b + m * step);
                }
            }
            if (DOUBLE) {
#pragma unroll
                for (int m = 0; m < 8; ++m) SIMT_PREFETCH_L1(P.w + off_a + m * step);
            }
        }
        if (it > 0) SIMT_SYNC();   // previous iteration's combine reads vs this iteration's writes
        if (PF == 1) {
            dht_combine_paired(v, lt);                     // combine on the input (zeros stay zeros)
            fft_line_pre<(LG > 0 ? LG : 8)>(v, lv, lt, P.tw);
        } else if (PF == 2 || PF == 3) {
            fft_line_post<(LG > 0 ? LG : 8)>(v, lv, lt, P.tw2);
            dht_combine_paired(v, lt);
        } else {
            fft_line_forward<LG>(v, lv, lt, log2n, P.tw);
            dht_combine_read<LG>(v, lv, lt, log2n);
        }
        SIMT_OPAQUE_U32(off_a); SIMT_OPAQUE_U32(off_b); SIMT_OPAQUE_U32(step);
        if (DOUBLE) {
            if (valid) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 wv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int m = 4 * h + j;
                        // plain (coherent) loads on purpose: read-only-path loads may be hoisted above the
                        // barriers to the top of the kernel, which keeps 32 registers live through the
                        // whole first transform and spills
                        if (PF == 3) {
                            const unsigned o = off_a + h * pd1 + j * 2 * step;
                            wv[j] = ld_pair(P.w, o, o + 1, true, true, adj);
                        } else {
                            wv[j] = pairvec ? ld_pair(P.w, off_a + m * step, off_b + m * step, va, vb, false)
                                            : ld_pair(P.w, off_a + m * step, off_b + m * step, true, true, adj);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        v[4 * h + j] = make_double2(v[4 * h + j].x * wv[j].x, v[4 * h + j].y * wv[j].y);
                }
            }
            if (PF == 3) {
                dht_combine_paired(v, lt);
                fft_line_pre<(LG > 0 ? LG : 8)>(v, lv, lt, P.tw);
            } else {
                SIMT_SYNC();
                fft_line_forward<LG>(v, lv, lt, log2n, P.tw);
                dht_combine_read<LG>(v, lv, lt, log2n);
            }
            SIMT_OPAQUE_U32(off_a); SIMT_OPAQUE_U32(off_b); SIMT_OPAQUE_U32(step);
        }
        if (valid) {
            if (pairvec) {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    epi_pair<EPI>(P, sa, sb, off_a + m * step, off_b + m * step, va, vb, false, v[m], acc_a, acc_b);
            } else if (EPI == EPI_METRIC && staged) {
                SIMT_MBAR_WAIT(mbar, 0);
                epi_metric_block_staged(v, P.out, stg, P.A, off_a, off_b, step, (long long)sa * P.out_ld, lt, 1 << log2n,
                                        P.scale, acc_a);
            } else if (EPI == EPI_METRIC) {
                if (PF == 2)
                    epi_metric_block<GEO == GEO_STRIDED>(v, P.out, P.vin, P.A, off_a, off_b, pd1, 2 * step, (long long)sa * P.out_ld,
                                                         (long long)sa * P.vin_ld, P.scale, acc_a);
                else
                    epi_metric_block<GEO == GEO_STRIDED>(v, P.out, P.vin, P.A, off_a, off_b, 4 * (int)step, step, (long long)sa * P.out_ld,
                                                         (long long)sa * P.vin_ld, P.scale, acc_a);
            } else if (EPI == EPI_STORE && GEO == GEO_CONTIG && P.t_lg_wt && n8 >= 32) {
                // tile-major scratch, row pass: the mirror image of the paired load above
                const int q = lt & 1;
                double* dst = P.out + (long long)sa * P.out_ld;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const double rcv = SIMT_SHFL_XOR_F64(q ? v[m].x : v[m].y, 1);
                    const unsigned o = q ? (toff_b + m * tstep - 1) : (toff_a + m * tstep);
                    *reinterpret_cast<double2*>(dst + o) = q ? make_double2(rcv, v[m].y) : make_double2(v[m].x, rcv);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    if (EPI == EPI_STORE)
                        epi_pair<EPI>(P, sa, sb, toff_a + m * tstep, toff_b + m * tstep, true, true, adj, v[m], acc_a, acc_b);
                    else
                        epi_pair<EPI>(P, sa, sb, off_a + m * step, off_b + m * step, true, true, adj, v[m], acc_a, acc_b);
            }
        }
    }
    if (DOUBLE && GEO == GEO_STRIDED && x_pend) deferred_x_update(P, sgrp, bt, tid);
    // ---- scalar reductions -----------------------------------------------------------------
    if (P.fin == FIN_NONE) return;
    if (P.fin == FIN_TOTAL) {
        const double tot = seg_reduce(acc_a + acc_b, SIMT_NTID, red);
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, 0, SIMT_BID, SIMT_NBID, tot, red,
                                   flag);
    } else if (pairvec) {
        // every vector lives entirely inside one unit (nloop == 1 for per-vector reductions)
        const double ta = seg_reduce(acc_a, 1 << lg_upt, red);
        SIMT_SYNC();
        const double tb = seg_reduce(acc_b, 1 << lg_upt, red);
        if (tu == 0) {
            if (va) finalize_vec(P.fin, P.cg, P.red_out, sa, ta);
            if (vb) finalize_vec(P.fin, P.cg, P.red_out, sb, tb);
        }
    } else {
        const double tot = seg_reduce(acc_a + acc_b, SIMT_NTID, red);
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, sgrp, bt, P.ntb, tot, red, flag);
    }
}

// ------------------------------------------------------------------------------------------
// Generic (any length) path for small or non-power-of-two axes: the same prologue / epilogue
// operators as elementwise kernels around a direct O(n^2) Hartley pass per axis.  This is what
// the reference's own 40-pixel fixture (test/test_models/model_fft_gp.jl:17) runs through.
// ------------------------------------------------------------------------------------------
struct ElemParams {
    FieldPassParams f;
    double* tmp;           // [nvec][N] staging values
    long long total;       // nvec * N
};

template <int PRO>
SIMT_DEVICE void elem_pro_body(const ElemParams& E, unsigned char* smem_raw) {
    const FieldPassParams& P = E.f;
    double* red = reinterpret_cast<double*>(smem_raw);
    int* flag = reinterpret_cast<int*>(smem_raw + 256);
    // block -> (chunk of pixels, vector)
    const long long nchunk = (P.N + SIMT_NTID - 1) / SIMT_NTID;
    const int s = (int)(SIMT_BID % P.nvec);
    const long long chunk = SIMT_BID / P.nvec;
    const long long pix = chunk * SIMT_NTID + SIMT_TID;
    const bool use_active = (P.mask_active != 0);
    if (use_active && !P.cg.active[s]) return;
    double acc_a = 0.0, acc_b = 0.0;
    if (pix < P.N) {
        double2 v = pro_pair<PRO>(P, s, s, pix, pix, true, false, false, acc_a, acc_b);
        E.tmp[(long long)s * P.N + pix] = v.x;
    }
    if (P.fin == FIN_NONE) return;
    const double tot = seg_reduce(acc_a, SIMT_NTID, red);
    if (P.fin == FIN_TOTAL)
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, 0, SIMT_BID, SIMT_NBID, tot, red,
                                   flag);
    else
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, s, chunk, nchunk, tot, red, flag);
}

template <int EPI>
SIMT_DEVICE void elem_epi_body(const ElemParams& E, unsigned char* smem_raw) {
    const FieldPassParams& P = E.f;
    double* red = reinterpret_cast<double*>(smem_raw);
    int* flag = reinterpret_cast<int*>(smem_raw + 256);
    const long long nchunk = (P.N + SIMT_NTID - 1) / SIMT_NTID;
    double acc_a = 0.0, acc_b = 0.0;
    const int s = (int)(SIMT_BID % P.nvec);
    const long long chunk = SIMT_BID / P.nvec;
    const long long pix = chunk * SIMT_NTID + SIMT_TID;
    const bool use_active = (P.mask_active != 0);
    if (use_active && !P.cg.active[s]) return;
    if (pix < P.N) {
        double2 val = make_double2(E.tmp[(long long)s * P.N + pix], 0.0);
        epi_pair<EPI>(P, s, s, pix, pix, true, false, false, val, acc_a, acc_b);
    }
    if (P.fin == FIN_NONE) return;
    const double tot = seg_reduce(acc_a, SIMT_NTID, red);
    if (P.fin == FIN_TOTAL)
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, 0, SIMT_BID, SIMT_NBID, tot, red,
                                   flag);
    else
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, s, chunk, nchunk, tot, red, flag);
}

// Direct Hartley pass along one axis of [outer][n][inner], any n: out = 2 * H(in) (the factor 2
// matches what dht_combine_read produces so both paths share scale constants), optionally
// multiplied by a weight AFTER the transform (mid weight of the metric).
struct NaiveParams {
    const double* in; double* out;
    const double* cas;     // n x n table, cas[k*n + j] = cos + sin (2 pi j k / n)
    const double* w;       // optional pointwise weight on the output (N), or null
    long long N, inner, total;   // total = nvec * N
    int n;
};

SIMT_DEVICE void naive_axis_body(const NaiveParams& P, unsigned char*) {
    const long long idx = SIMT_BID * SIMT_NTID + SIMT_TID;
    if (idx >= P.total) return;
    const long long s = idx / P.N, pix = idx - s * P.N;
    const long long in_idx = pix % P.inner;
    const long long k = (pix / P.inner) % P.n;
    const long long o = pix / (P.inner * P.n);
    const double* src = P.in + s * P.N + o * P.n * P.inner + in_idx;
    const double* row = P.cas + k * P.n;
    double a = 0.0;
    for (int j = 0; j < P.n; ++j) a += SIMT_LDG(row + j) * src[(long long)j * P.inner];
    a *= 2.0;
    if (P.w) a *= SIMT_LDG(P.w + pix);
    P.out[idx] = a;
}

// ------------------------------------------------------------------------------------------
// Any-length Hartley pass in O(m log m): Bluestein's chirp-z on top of the power-of-two FFT core.
//   Z[k] = sum_j z_j e^(-2 pi i j k / n),  j k = (j^2 + k^2 - (k - j)^2) / 2
//        = c_k * sum_j (z_j c_j) conj(c_(k-j)),        c_j = e^(-i pi j^2 / n)
// i.e. a cyclic convolution of length m >= 2n - 1 (m a power of two) of y_j = z_j c_j with the wrapped chirp
// h_j = conj(c_|j|): two length-m FFTs per complex line (the inverse one as conj(FFT(conj(.)))), the FFT of h
// precomputed (scaled by 1/m).  As everywhere a complex line packs TWO real lines z = a + i b and the Hartley
// pair comes out of Z[k], Z[n-k]; the result is 2 H(.) like the direct pass it replaces (naive_axis_body).
// One block = one pair of neighbouring lines, m/8 threads.
// ------------------------------------------------------------------------------------------
struct BlueParams {
    const double* in; double* out;
    const double* w;            // optional pointwise weight on the output (N), or null
    const double2* chirp;       // c_j, j < n
    const double2* hf;          // FFT_m(h) / m
    const double2* tw;          // per-stage twiddles of the length-m FFT (fft_core.h layout)
    long long N, inner, nlines; // nlines = nvec * N / n real lines
    long long lines_per_vec;    // N / n
    const int* active;          // CG iterations: per-vector active flags (a converged vector's stale line must not
                                // share a complex transform with a live one: its magnitude drifts away and the
                                // rounding of the partner is relative to the larger of the two), or null
    int n, lg_m;
};

// LGM > 0 fixes log2(m) at compile time (stage loops unrolled, swizzle offsets folded into immediates).
template <int LGM>
SIMT_DEVICE void bluestein_axis_body(const BlueParams& P, unsigned char* smem_raw) {
    const int lt = SIMT_TID;
    const int lg_m = LGM ? LGM : P.lg_m;
    const int m8 = 1 << (lg_m - 3);
    LineView lv;
    lv.sm = reinterpret_cast<double2*>(smem_raw + kSmemHeader);
    lv.lgL = 0;
    lv.l = 0;
    const long long la = 2 * SIMT_BID, lb = la + 1;
    bool va = true, vb = lb < P.nlines;
    if (P.active) {
        va = P.active[la / P.lines_per_vec] != 0;
        vb = vb && P.active[lb / P.lines_per_vec] != 0;
        if (!va && !vb) return;                  // uniform for the block
    }
    const long long ni = (long long)P.n * P.inner;
    const long long base_a = (la / P.inner) * ni + (la % P.inner);
    const long long base_b = vb ? (lb / P.inner) * ni + (lb % P.inner) : base_a;
    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int j = lt + t * m8;
        if (j < P.n) {
            const double a = va ? P.in[base_a + (long long)j * P.inner] : 0.0;
            const double b = vb ? P.in[base_b + (long long)j * P.inner] : 0.0;
            v[t] = cmul(make_double2(a, b), SIMT_LDG(P.chirp + j));
        } else {
            v[t] = make_double2(0.0, 0.0);
        }
    }
    fft_line_forward<LGM>(v, lv, lt, lg_m, P.tw);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const double2 x = cmul(v[t], SIMT_LDG(P.hf + lt + t * m8));
        v[t] = make_double2(x.x, -x.y);
    }
    fft_line_forward<LGM>(v, lv, lt, lg_m, P.tw);
    // shared memory now holds conj(convolution) in natural order; Z[k] = c_k * conj(.)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int k = lt + t * m8;
        if (k < P.n) {
            const int kp = k ? P.n - k : 0;
            const double2 zk = cmul(SIMT_LDG(P.chirp + k), make_double2(v[t].x, -v[t].y));
            const double2 sp = lv.raw(kp);
            const double2 zn = cmul(SIMT_LDG(P.chirp + kp), make_double2(sp.x, -sp.y));
            double oa = zk.x - zk.y + zn.x + zn.y, ob = zk.x + zk.y + zn.y - zn.x;
            const long long pa = base_a + (long long)k * P.inner, pb = base_b + (long long)k * P.inner;
            if (P.w) { oa *= SIMT_LDG(P.w + pa % P.N); if (vb) ob *= SIMT_LDG(P.w + pb % P.N); }
            if (va) P.out[pa] = oa;
            if (vb) P.out[pb] = ob;
        }
    }
}

}  // namespace mgvi
