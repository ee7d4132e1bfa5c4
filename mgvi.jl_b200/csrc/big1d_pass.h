// Long 1-D fields (n > 4096, BASELINE.json configs[1]: 2^16 pixels): the transform of one line does
// not fit one block, so it is split four-step style, n = n1 * n2, over two kernels per transform,
// with TWO residual samples packed into each complex line (a + i b) exactly as in fft_core.h.
//
//   forward  H: x (natural order j = j1 n2 + j2)  ->  y in "permuted" order: y[k1 + n1 k2] stored at [k1][k2]
//       A : for every column j2: FFT over j1 (stride n2), times W_n^(j2 k1)          -> complex scratch
//       B : for every row k1: FFT over j2 (contiguous), Hartley combine with row n1-k1
//   adjoint  H: y (permuted order)  ->  x (natural order m = m1 n2 + m2)
//       B': for every row k1: FFT over k2 (contiguous), times W_n^(k1 m2)             -> complex scratch
//       A': for every column m2: FFT over k1 (stride n2), Hartley combine with column n2-m2
//
// Everything that lives in data space (w, q, the observations, gradient accumulators) is kept in
// the permuted order, so no transposition pass exists, and the two row kernels of the metric
// product M v = v + c^2 A . H(w . H(A . v)) fuse into one:  A, [B, w., B'], A'  -- three kernels.
#pragma once
#include "field_pass.h"

namespace mgvi {

enum { BIG_A = 0, BIG_B_END = 1, BIG_B_MID = 2, BIG_B_START = 3, BIG_AP = 4 };

struct BigParams {
    FieldPassParams f;          // operands / epilogue constants (geometry fields unused)
    int lg_n1, lg_n2;           // n = n1 * n2
    int lgL, G;                 // lines side by side per unit, units per block
    long long units_per_pair;   // A: n2 / L tiles ; B*: n1 / 2 row pairs ; A': n2 / 2 column pairs
    int npairs, nloop;          // pairs of vectors; nloop = 1 (one pair per block group)
    double2* cs;                // complex scratch [npairs][n]
    const double2* tw1;         // per-stage twiddles for length n1 (fft_core.h layout)
    const double2* tw2;         // ... for length n2
    const double2* tlo;         // W_n^e, e < 2^h
    const double2* thi;         // W_n^(e 2^h), e < n / 2^h
    int lg_h;
};

SIMT_DEVICE double2 big_twiddle(const BigParams& B, long long e) {
    const double2 lo = SIMT_LDG(&B.tlo[e & ((1LL << B.lg_h) - 1)]);
    const double2 hi = SIMT_LDG(&B.thi[e >> B.lg_h]);
    return cmul(lo, hi);
}

// storage index (row k1, column k2) -> logical index k = k1 + n1 k2 of the permuted data space
SIMT_DEVICE long long big_logical(const BigParams& B, long long pix) {
    const long long r = pix >> B.lg_n2, c = pix & ((1LL << B.lg_n2) - 1);
    return r + (c << B.lg_n1);
}

template <int KIND, int PRO, int EPI>
SIMT_DEVICE void big1d_body(const BigParams& B, unsigned char* smem_raw) {
    const FieldPassParams& P = B.f;
    const long long BID = SIMT_BID;
    const int lg_len = (KIND == BIG_A || KIND == BIG_AP) ? B.lg_n1 : B.lg_n2;     // transform length of this kernel
    const int n8 = 1 << (lg_len - 3);
    const int lgL = B.lgL;
    const int tid = SIMT_TID;
    const int lg_upt = lg_len - 3 + lgL;
    const int g = tid >> lg_upt;
    const int tu = tid & ((1 << lg_upt) - 1);
    const int l = tu & ((1 << lgL) - 1), lt = tu >> lgL;
    const long long n2 = 1LL << B.lg_n2, n1 = 1LL << B.lg_n1, n = n1 * n2;
    double* red = reinterpret_cast<double*>(smem_raw);
    int* flag = reinterpret_cast<int*>(smem_raw + 256);
    LineView lv;
    lv.sm = reinterpret_cast<double2*>(smem_raw + kSmemHeader) + ((long long)g << (lgL + lg_len));
    lv.lgL = lgL;
    lv.l = l;
    const double2* tw = (KIND == BIG_A || KIND == BIG_AP) ? B.tw1 : B.tw2;

    // block -> (unit block, pair group); units of a block share the pair
    const long long ubl = (B.units_per_pair + B.G - 1) / B.G;
    const long long bt = BID % ubl;
    const long long pgrp = BID / ubl;
    const long long u = bt * B.G + g;
    const bool unit_ok = u < B.units_per_pair;

    // element positions of my line: storage index = base + k * stride (k = lt + m n8)
    long long base = 0, stride = 1;
    long long row = 0, col = 0;           // row (B kernels) / column (A kernels) of my line
    bool self_pair = false;               // the partner of my line is my own line
    if (KIND == BIG_A) {
        col = (u << lgL) + l; base = col; stride = n2;
    } else if (KIND == BIG_AP) {
        if (lgL == 1) {
            // column pair (u, n2 - u); unit 0 holds the two self-mirrored columns 0 and n2/2
            if (u == 0) { col = l ? (n2 >> 1) : 0; self_pair = true; }
            else col = l ? (n2 - u) : u;
        } else {
            // H = 2^(lgL-1) neighbouring columns u H .. u H + H-1 (lines 0..H-1) and their mirrors (lines H..2H-1,
            // line H + i <-> column n2 - (u H + i)): whole 32-byte sectors per row instead of one double per
            // sector.  Column 0 mirrors onto itself; its mirror slot holds the other self-mirrored column n2/2.
            const int H = 1 << (lgL - 1), i = l & (H - 1);
            const long long cdir = u * H + i;
            if (l < H) col = cdir;
            else col = (cdir == 0) ? (n2 >> 1) : (n2 - cdir);
            self_pair = (cdir == 0);
        }
        base = col; stride = n2;
    } else {
        if (u == 0) { row = l ? (n1 >> 1) : 0; self_pair = true; }
        else row = l ? (n1 - u) : u;
        base = row * n2; stride = 1;
    }

    double acc_a = 0.0, acc_b = 0.0;
    int sa = 0, sb = 0;
    bool va = false, vb = false;
    for (int it = 0; it < B.nloop; ++it) {
        const long long pair = pgrp * B.nloop + it;
        sa = (int)(2 * pair); sb = sa + 1;
        va = unit_ok && pair < B.npairs && sa < P.nvec;
        vb = unit_ok && pair < B.npairs && sb < P.nvec;
        if (P.mask_active) { va = va && P.cg.active[sa]; vb = vb && P.cg.active[sb]; }
        const bool valid = va || vb;
        double2* cs = B.cs + pair * n;
        double2 v[8];
        // ---------------- load ----------------
        if (KIND == BIG_A && PRO == PRO_CG && valid) {
            // CG iteration: p = r + beta p, v = A . p -- all loads of the thread's 8 elements first, then the
            // stores (the per-element form alternates loads with stores to arrays that may alias and pays one
            // memory round trip per element)
            double* __restrict__ pA = P.p + (long long)sa * P.N;
            double* __restrict__ pB = P.p + (long long)sb * P.N;
            const double* __restrict__ rA = P.r + (long long)sa * P.N;
            const double* __restrict__ rB = P.r + (long long)sb * P.N;
            double2 pp[8], rr[8];
            double aa[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long pix = base + (long long)(lt + m * n8) * stride;
                pp[m] = make_double2(va ? pA[pix] : 0.0, vb ? pB[pix] : 0.0);
                aa[m] = SIMT_LDG(P.A + pix);
                if (!P.first_iter) rr[m] = make_double2(va ? rA[pix] : 0.0, vb ? rB[pix] : 0.0);
            }
            if (!P.first_iter) {
                const double ba = va ? P.cg.beta[sa] : 0.0, bb = vb ? P.cg.beta[sb] : 0.0;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const long long pix = base + (long long)(lt + m * n8) * stride;
                    pp[m] = make_double2(rr[m].x + ba * pp[m].x, rr[m].y + bb * pp[m].y);
                    if (va) pA[pix] = pp[m].x;
                    if (vb) pB[pix] = pp[m].y;
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = make_double2(aa[m] * pp[m].x, aa[m] * pp[m].y);
        } else
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const long long pix = base + (long long)(lt + m * n8) * stride;
            if (!valid) v[m] = make_double2(0.0, 0.0);
            else if (KIND == BIG_A) v[m] = pro_pair<PRO>(P, sa, sb, pix, pix, va, vb, false, acc_a, acc_b);
            else if (KIND == BIG_B_START) {
                if (PRO == PRO_RHS) {
                    // draws are indexed by the LOGICAL lambda index
                    const long long kl = big_logical(B, pix);
                    const double qv = SIMT_LDG(&P.q[pix]);
                    const double za = va ? P.Z1[sa * P.Z1_ld + P.Z1_off + P.Z1_stride * kl] : 0.0;
                    const double zb = vb ? P.Z1[sb * P.Z1_ld + P.Z1_off + P.Z1_stride * kl] : 0.0;
                    v[m] = make_double2(qv * za, qv * zb);
                } else {
                    v[m] = pro_pair<PRO>(P, sa, sb, pix, pix, va, vb, false, acc_a, acc_b);
                }
            } else v[m] = cs[pix];
        }
        if (it > 0) SIMT_SYNC();
        fft_line_forward<0>(v, lv, lt, lg_len, tw);
        if (KIND == BIG_A) {
            // times W_n^(j2 k1), store the complex intermediate
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long k1 = lt + m * n8;
                if (valid) cs[k1 * n2 + col] = cmul(v[m], big_twiddle(B, k1 * col));
            }
            continue;
        }
        if (KIND == BIG_B_START) {
            // times W_n^(k1 m2), store the complex intermediate
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long m2 = lt + m * n8;
                if (valid) cs[row * n2 + m2] = cmul(v[m], big_twiddle(B, row * m2));
            }
            continue;
        }
        // ---------------- Hartley combine with the mirrored line ----------------
        {
            LineView pv = lv;
            pv.l = self_pair ? l : ((KIND == BIG_AP) ? (l ^ (1 << (lgL - 1))) : (l ^ 1));
            const long long len = 1LL << lg_len;
            // partner position along the line: the line with index 0 (row 0 / column 0) mirrors onto itself
            // with (len - k) mod len; every other line mirrors with len - 1 - k
            const bool zero_line = (KIND == BIG_AP) ? (col == 0) : (row == 0);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int k = lt + m * n8;
                const int kp = zero_line ? (int)((len - k) & (len - 1)) : (int)(len - 1 - k);
                const double2 zn = pv.raw(kp);
                const double2 zk = v[m];
                v[m] = make_double2(zk.x - zk.y + zn.x + zn.y, zk.x + zk.y + zn.y - zn.x);
            }
        }
        if (KIND == BIG_B_MID) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long pix = base + (lt + m * n8);
                const double wv = valid ? P.w[pix] : 0.0;
                v[m] = make_double2(v[m].x * wv, v[m].y * wv);
            }
            SIMT_SYNC();
            fft_line_forward<0>(v, lv, lt, lg_len, tw);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long m2 = lt + m * n8;
                if (valid) cs[row * n2 + m2] = cmul(v[m], big_twiddle(B, row * m2));
            }
            continue;
        }
        // ---------------- epilogue (BIG_B_END in permuted data space, BIG_AP in natural order) ----------------
        if (valid && EPI == EPI_METRIC) {
            // out = vin + scale * A . val ; acc += vin . out -- loads first, then stores (see the CG prologue above)
            const double* __restrict__ iA = P.vin + (long long)sa * P.vin_ld;
            const double* __restrict__ iB = P.vin + (long long)sb * P.vin_ld;
            double* __restrict__ oA = P.out + (long long)sa * P.out_ld;
            double* __restrict__ oB = P.out + (long long)sb * P.out_ld;
            double2 vi[8];
            double aa[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long pix = base + (long long)(lt + m * n8) * stride;
                vi[m] = make_double2(va ? iA[pix] : 0.0, vb ? iB[pix] : 0.0);
                aa[m] = SIMT_LDG(P.A + pix);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long pix = base + (long long)(lt + m * n8) * stride;
                const double2 o = make_double2(vi[m].x + P.scale * (aa[m] * v[m].x), vi[m].y + P.scale * (aa[m] * v[m].y));
                if (va) { oA[pix] = o.x; acc_a += vi[m].x * o.x; }
                if (vb) { oB[pix] = o.y; acc_b += vi[m].y * o.y; }
            }
        } else if (valid) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const long long pix = base + (long long)(lt + m * n8) * stride;
                epi_pair<EPI>(P, sa, sb, pix, pix, va, vb, false, v[m], acc_a, acc_b);
            }
        }
    }
    // reductions: prologue sums (prior term of the KL estimate) in the first kernel of a chain, epilogue
    // sums (likelihood terms, CG dots) in the last; the middle kernels carry none
    if ((KIND == BIG_B_START || KIND == BIG_B_MID) && P.fin != FIN_TOTAL) return;
    if (P.fin == FIN_NONE) return;
    if (P.fin == FIN_TOTAL) {
        const double tot = seg_reduce(acc_a + acc_b, SIMT_NTID, red);
        block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, 0, BID, SIMT_NBID, tot, red,
                                   flag);
    } else {
        // per-vector reductions (CG dots): blocks of one pair group publish two partials each
        const double ta = seg_reduce(acc_a, SIMT_NTID, red);
        SIMT_SYNC();
        const double tb = seg_reduce(acc_b, SIMT_NTID, red);
        const bool pa_ok = (2 * pgrp) < P.nvec && (!P.mask_active || P.cg.active[2 * pgrp]);
        const bool pb_ok = (2 * pgrp + 1) < P.nvec && (!P.mask_active || P.cg.active[2 * pgrp + 1]);
        if (pa_ok)
            block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, (int)(2 * pgrp), bt, ubl, ta, red,
                                       flag);
        SIMT_SYNC();
        if (pb_ok)
            block_publish_and_finalize(P.fin, P.cg, P.red_out, P.partials, P.counters, (int)(2 * pgrp + 1), bt, ubl, tb,
                                       red, flag);
    }
}

}  // namespace mgvi
