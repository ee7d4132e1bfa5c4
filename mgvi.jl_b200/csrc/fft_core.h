// Register-resident complex FFT for lines staged through shared memory (FP64).
//
// One "complex line" packs TWO real lines (a + i b).  The discrete Hartley transform of both
// comes out of one complex FFT Z = FFT(a + i b) through
//     H(a)[k] + i H(b)[k] = 1/2 [ (1+i) Z[k] + (1-i) Z[n-k] ]
// (the unnormalised separable DHT FFTW's r2r/DHT plan computes -- reference models
//  test/test_models/model_fft_gp.jl:21, docs/src/advanced_tutorial_lit.jl:208).
//
// Layout: a thread owns 8 elements of its line, k = lt + m*n/8 (m = 0..7, lt = thread index
// within the line, n/8 threads per line).  Stockham autosort stages (radix 8, with one leading
// radix-4 or radix-2 stage when log2 n is not a multiple of 3) exchange through ONE shared
// buffer per line: write scattered, barrier, read back in the k = lt + m*n/8 arrangement,
// barrier.  Shared memory holds double2 (16 B) chunks at index (k*L + l) for line l of L
// side-by-side lines (L a power of two), XOR-swizzled on the chunk index (bits 0..2 ^= bits
// 3..5) -- the same pattern as the TMA 128-byte swizzle -- which keeps every stage's scattered
// writes and linear reads free of bank conflicts.
//
// Everything is written for a small register footprint (the kernels want 3-4 resident blocks
// of 256 threads per SM): power-of-two index arithmetic is shifts and masks, twiddles are
// consumed one at a time, butterflies work in place.
#pragma once
#include "simt.h"

namespace mgvi {

SIMT_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
SIMT_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
SIMT_HD double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i
SIMT_HD double2 cmul_mi(double2 a) { return make_double2(a.y, -a.x); }

// XOR swizzle of the 16-byte chunk index (only bits 3..5 are looked at, only bits 0..2 change).
// One line per unit (lgL = 0): bits 0..2 ^= bits 3..5 (the TMA 128-byte pattern).  Side-by-side
// lines (lgL >= 1) keep the line index in bit 0, so row bits 3..4 are folded into bits 1..2 instead
// (same cost: one shift, one mask, one xor).
// Checked by simulation of every stage's access pattern (tools/bank_conflicts.py): conflict-free
// for the benchmark shapes.
SIMT_HD int swz(int chunk, int lgL) {
    return lgL == 0 ? (chunk ^ ((chunk >> 3) & 7)) : (chunk ^ ((chunk >> 2) & 6));
}

// in-place radix-2 butterfly
SIMT_HD void bfly2(double2& a, double2& b) {
    const double2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// 4-point forward DFT in place (natural order out).
SIMT_HD void dft4(double2& c0, double2& c1, double2& c2, double2& c3) {
    bfly2(c0, c2);              // c0 = d0, c2 = d1
    bfly2(c1, c3);              // c1 = d2, c3 = c1 - c3
    c3 = cmul_mi(c3);           // d3
    bfly2(c0, c1);              // c0 = Y0, c1 = Y2
    bfly2(c2, c3);              // c2 = Y1, c3 = Y3
    const double2 t = c1; c1 = c2; c2 = t;
}

// 8-point forward DFT in place (natural order out).
SIMT_HD void dft8(double2 (&a)[8]) {
    const double h = 0.70710678118654752440;
    bfly2(a[0], a[4]);
    bfly2(a[1], a[5]);
    bfly2(a[2], a[6]);
    bfly2(a[3], a[7]);
    // odd branch twiddles W8^1 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
    a[5] = make_double2((a[5].x + a[5].y) * h, (a[5].y - a[5].x) * h);
    a[6] = cmul_mi(a[6]);
    a[7] = make_double2((a[7].y - a[7].x) * h, -(a[7].x + a[7].y) * h);
    dft4(a[0], a[1], a[2], a[3]);     // X0 X2 X4 X6
    dft4(a[4], a[5], a[6], a[7]);     // X1 X3 X5 X7
    double2 t;
    // permute (X0 X2 X4 X6 X1 X3 X5 X7) -> natural order
    t = a[1]; a[1] = a[4]; a[4] = a[2]; a[2] = t;      // a1 = X1, a4 = X4, a2 = X2
    t = a[3]; a[3] = a[5]; a[5] = a[6]; a[6] = t;      // a3 = X3, a5 = X5, a6 = X6
}

// Shared-memory view of one complex line among L = 2^lgL side-by-side lines.
struct LineView {
    double2* sm;     // base of the unit's buffer (L*n chunks)
    int lgL, l;
    SIMT_DEVICE int idx(int k) const { return swz((k << lgL) | l, lgL); }
    SIMT_DEVICE double2& at(int k) const { return sm[idx(k)]; }
    // unswizzled position: the final exchange (own outputs out, mirrored partners in) touches runs of
    // consecutive chunks that straddle 8-chunk blocks, which the swizzle would turn into 2-way conflicts
    SIMT_DEVICE double2& raw(int k) const { return sm[(k << lgL) | l]; }
};

// Forward FFT of the line whose elements the calling threads hold in v (v[m] = x[lt + m*n/8]).
// On return v[m] = Z[lt + m*n/8] (each thread keeps its own outputs in registers) AND the whole
// line sits in shared memory in natural order (all threads have passed the final barrier), so
// partners can be read.  EVERY thread of the block must call this (uniform barriers); threads
// whose unit is out of range simply transform zeros inside their own shared buffer (no branches:
// a branch around each phase costs a register move per live value at every merge point).
// n = 2^log2n >= 8.  tw is the per-stage twiddle table: for every radix-8 stage with Ns > 1, in
// stage order, three runs of Ns entries  W^(j), W^(2j), W^(4j)  (j < Ns, W = exp(-2 pi i / (8 Ns))),
// so that the lanes of a warp (consecutive j) read consecutive 16-byte entries.
// LG > 0 fixes log2 n at compile time (stage loop unrolled, shifts and swizzle offsets folded
// into immediates); LG = 0 takes it from log2n_rt.
template <int LG>
SIMT_DEVICE void fft_line_forward(double2 (&v)[8], const LineView& lv, int lt, int log2n_rt,
                                  const double2* __restrict__ tw) {
    const int log2n = LG ? LG : log2n_rt;
    const int n8 = 1 << (log2n - 3);
    const int lgL = lv.lgL;
    const int lead = log2n % 3;   // 0: all radix-8; 1: leading radix-2; 2: leading radix-4
    // linear read-back positions: swz(((lt + m n8) << lgL) | l) = rd0 + m * rd_step once the step is a
    // multiple of 64 chunks (the swizzle only looks at bits 3..5)
    const int rd_step = n8 << lgL;
    const bool rd_fast = (rd_step & 63) == 0;
    const int rd0 = lv.idx(lt);
    if (lead == 2) {
        // leading radix-4 stage (Ns = 1): two butterflies j = lt + q*n8, inputs v[q + 2t]
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            dft4(v[q], v[q + 2], v[q + 4], v[q + 6]);
            const int j4 = (lt + q * n8) << 2;
#pragma unroll
            for (int t = 0; t < 4; ++t) lv.at(j4 + t) = v[q + 2 * t];
        }
    } else if (lead == 1) {
        // leading radix-2 stage (Ns = 1): four butterflies j = lt + q*n8, inputs v[q], v[q + 4]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            bfly2(v[q], v[q + 4]);
            const int j2 = (lt + q * n8) << 1;
            lv.at(j2) = v[q];
            lv.at(j2 + 1) = v[q + 4];
        }
    }
    if (lead != 0) {
        SIMT_SYNC();
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = lv.sm[rd_fast ? rd0 + m * rd_step : lv.idx(lt + m * n8)];
        SIMT_SYNC();
    }
    // radix-8 stages
#pragma unroll
    for (int lgNs = lead; lgNs < log2n; lgNs += 3) {
        {
            const int jlo = lt & ((1 << lgNs) - 1);
            if (lgNs > 0) {
                // twiddles W^(jlo*t): three coalesced table loads, the other four by products
                // stage offset: 3 * (sum of the Ns of the earlier radix-8 stages that have twiddles, i.e. Ns > 1)
                const double2* tws = tw + 3 * (((1 << lgNs) - (1 << lead)) / 7 - (lead == 0 ? 1 : 0)) + jlo;
                const double2 w1 = SIMT_LDG(tws);
#ifndef MGVI_TW_TABLE
                // one table load per stage, W^2j and W^4j by squaring: 2 L1 wavefront groups less per stage
                // for 8 more FP64 instructions on a pipe that is 40 % busy; costs <= 2 ulp on the twiddles.
                // Measured on B200 (profiles/README.md, round 2): axis0_fused 0.90 -> 0.84 ms, C3 iteration
                // 3.553 -> 3.488 ms.  -DMGVI_TW_TABLE restores the three table loads.
                const double2 w2 = cmul(w1, w1);
#else
                const double2 w2 = SIMT_LDG(tws + (1 << lgNs));
#endif
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                double2 w3 = cmul(w1, w2);
                v[3] = cmul(v[3], w3);
#ifndef MGVI_TW_TABLE
                const double2 w4 = cmul(w2, w2);
#else
                const double2 w4 = SIMT_LDG(tws + (2 << lgNs));
#endif
                v[4] = cmul(v[4], w4);
                w3 = cmul(w3, w4);                // W^7
                v[7] = cmul(v[7], w3);
                w3 = cmul(w1, w4);                // W^5
                v[5] = cmul(v[5], w3);
                w3 = cmul(w2, w4);                // W^6
                v[6] = cmul(v[6], w3);
            }
            dft8(v);
            const int base = ((lt >> lgNs) << (lgNs + 3)) | jlo;
            const int cb = (base << lgL) | lv.l;
            const int sh = lgNs + lgL;
            if (lgNs + 3 >= log2n) {
                // final stage: plain positions (see LineView::raw)
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[cb + (t << sh)] = v[t];
            } else if (sh >= 6) {
                const int w0 = swz(cb, lgL);
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[w0 + (t << sh)] = v[t];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[swz(cb + (t << sh), lgL)] = v[t];
            }
        }
        SIMT_SYNC();
        if (lgNs + 3 < log2n) {
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = lv.sm[rd_fast ? rd0 + m * rd_step : lv.idx(lt + m * n8)];
            SIMT_SYNC();
        }
    }
}

// After fft_line_forward: v[m] <- 2 * ( H(a)[k], H(b)[k] ), k = lt + m*n/8.  Own values come from
// registers, the partners Z[n-k] from shared memory (they live in thread n/8 - lt, register 7-m).
// (The factor 2 is folded into the callers' scale constants; powers of two are exact.)
template <int LG>
SIMT_DEVICE void dht_combine_read(double2 (&v)[8], const LineView& lv, int lt, int log2n_rt) {
    const int log2n = LG ? LG : log2n_rt;
    const int n8 = 1 << (log2n - 3);
    const int rd_step = n8 << lv.lgL;
    const int plt = (n8 - lt) & (n8 - 1);
    const int p0 = (plt << lv.lgL) | lv.l;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double2 zn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = 4 * h + j;
            const int pm = (lt == 0) ? ((8 - m) & 7) : (7 - m);
            zn[j] = lv.sm[p0 + pm * rd_step];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 zk = v[4 * h + j];
            v[4 * h + j] = make_double2(zk.x - zk.y + zn[j].x + zn[j].y, zk.x + zk.y + zn[j].y - zn[j].x);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Paired arrangement for lengths n = 4 * 8^s (log2 n = 2 mod 3: the 2048- and 256-point axes of the
// benchmark shapes).  A thread's 8 registers hold
//     slot t     : element jA + t * n/4,   jA = lt
//     slot 4 + t : element jB + t * n/4,   jB = n/4 - lt   (thread 0: jB = n/8)          t = 0..3
// so that an element k and its Hartley partner n - k sit in the SAME thread (slot t <-> slot 7 - t;
// thread 0: slot t <-> slot (4 - t) & 3 and slot 4 + t <-> slot 7 - t).  The combine
//     H(a)[k] + i H(b)[k] = 1/2 [ (1+i) Z[k] + (1-i) Z[n-k] ]
// is linear, so it can be applied to the OUTPUT of a transform (Z = FFT(a + i b)) or -- because
// Z[n-k] = FFT(index-reversed input)[k] -- to its INPUT (y[j] = 1/2[(1+i) z[j] + (1-i) z[n-j]], FFT(y) is
// the Hartley pair directly).  With the radix-4 stage of the Stockham factorisation placed LAST
// (8 . 8 . 8 . 4: its two butterflies jA, jB produce exactly the paired slots) or FIRST (4 . 8 . 8 . 8: its
// two butterflies consume exactly the paired slots) the combine needs NO shared-memory exchange:
//     fft_line_post : standard arrangement in  -> paired arrangement out   (then dht_combine_paired)
//     fft_line_pre  : (dht_combine_paired, then) paired arrangement in -> standard arrangement out
// each with three exchanges instead of the four (three stages + partner read) of fft_line_forward +
// dht_combine_read: 25 % less shared-memory traffic on kernels that are bound by the L1 data pipe.
// ------------------------------------------------------------------------------------------
template <int LG>
SIMT_DEVICE int paired_jB(int lt) { return lt ? ((1 << (LG - 2)) - lt) : (1 << (LG - 3)); }

// v <- 2 * combine(v) in registers (output side: Hartley pair from Z; input side: y from z)
SIMT_DEVICE void dht_combine_paired(double2 (&v)[8], int lt) {
    double2 o[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double2 pa = lt ? v[7 - t] : v[(4 - t) & 3];
        const double2 pb = lt ? v[3 - t] : v[7 - t];
        const double2 za = v[t], zb = v[4 + t];
        o[t] = make_double2(za.x - za.y + pa.x + pa.y, za.x + za.y + pa.y - pa.x);
        o[4 + t] = make_double2(zb.x - zb.y + pb.x + pb.y, zb.x + zb.y + pb.y - pb.x);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = o[m];
}

// radix-8 input twiddles W^(j t), t = 1..7, from the table entry w1 = W^j (W^2j, W^4j by squaring)
SIMT_DEVICE void r8_twiddle(double2 (&v)[8], double2 w1) {
    const double2 w2 = cmul(w1, w1);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    double2 w3 = cmul(w1, w2);
    v[3] = cmul(v[3], w3);
    const double2 w4 = cmul(w2, w2);
    v[4] = cmul(v[4], w4);
    w3 = cmul(w3, w4);                // W^7
    v[7] = cmul(v[7], w3);
    w3 = cmul(w1, w4);                // W^5
    v[5] = cmul(v[5], w3);
    w3 = cmul(w2, w4);                // W^6
    v[6] = cmul(v[6], w3);
}

// exchange-specific chunk swizzles (tools/bank_conflicts_paired.py): 0 = the usual one, 1 = bits 4..5 folded
// into bits 1..2 (first exchange of fft_line_post with side-by-side lines), 2 = none
template <int SW>
SIMT_DEVICE int xswz(int c, int lgL) {
    return SW == 2 ? c : (SW == 1 ? (c ^ ((c >> 3) & 6)) : swz(c, lgL));
}

// Standard arrangement in (v[m] = x[lt + m n/8]), paired arrangement out (slot t = Z[jA + t n/4],
// slot 4 + t = Z[jB + t n/4]).  twp: per radix-8 stage with Ns = 8, 64, ... one run of Ns entries W^j
// (W = exp(-2 pi i / (8 Ns))), then n/4 entries exp(-2 pi i j / n) for the final radix-4 stage.
// Every thread of the block must call (uniform barriers).  Leaves no pending shared-memory reads.
template <int LG>
SIMT_DEVICE void fft_line_post(double2 (&v)[8], const LineView& lv, int lt, const double2* __restrict__ twp) {
    static_assert(LG % 3 == 2 && LG >= 8, "paired transforms: n = 4 * 8^s, n >= 256");
    constexpr int n8 = 1 << (LG - 3), n4 = 1 << (LG - 2);
    const int lgL = lv.lgL;
    const int rd_step = n8 << lgL;
    const int std0 = ((lt << lgL) | lv.l);
#pragma unroll
    for (int lgNs = 0; lgNs + 2 < LG; lgNs += 3) {
        const bool first = (lgNs == 0), last = (lgNs + 5 >= LG);
        const int jlo = lt & ((1 << lgNs) - 1);
        if (lgNs > 0) r8_twiddle(v, SIMT_LDG(twp + (((1 << lgNs) - 8) / 7) + jlo));
        dft8(v);
        const int base = ((lt >> lgNs) << (lgNs + 3)) | jlo;
        const int cb = (base << lgL) | lv.l;
        const int sh = lgNs + lgL;
        if (last) {
#pragma unroll
            for (int t = 0; t < 8; ++t) lv.sm[cb + (t << sh)] = v[t];
        } else if (first) {
            if (lgL == 0) {
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[xswz<0>(cb + (t << sh), lgL)] = v[t];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[xswz<1>(cb + (t << sh), lgL)] = v[t];
            }
        } else {
            // middle exchanges: t << sh is a multiple of 64 chunks here (sh >= 3 + lgL ... checked below)
            if (sh >= 6) {
                const int w0 = xswz<0>(cb, lgL);
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[w0 + (t << sh)] = v[t];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[xswz<0>(cb + (t << sh), lgL)] = v[t];
            }
        }
        SIMT_SYNC();
        if (!last) {
            // linear read-back in the standard arrangement (rd_step is a multiple of 64 chunks for n >= 512 or
            // side-by-side lines; otherwise the swizzle is evaluated per element)
            if ((rd_step & 63) == 0) {
                const int r0 = (first && lgL != 0) ? xswz<1>(std0, lgL) : xswz<0>(std0, lgL);
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = lv.sm[r0 + m * rd_step];
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    v[m] = lv.sm[(first && lgL != 0) ? xswz<1>(std0 + m * rd_step, lgL) : xswz<0>(std0 + m * rd_step, lgL)];
            }
            SIMT_SYNC();
        }
    }
    // paired read for the final radix-4 stage (plain positions: descending runs straddle 8-chunk blocks)
    const int jB = paired_jB<LG>(lt);
    const int a0 = (lt << lgL) | lv.l, b0 = (jB << lgL) | lv.l;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        v[t] = lv.sm[a0 + ((t * n4) << lgL)];
        v[4 + t] = lv.sm[b0 + ((t * n4) << lgL)];
    }
    SIMT_SYNC();
    const double2* twf = twp + (n4 - 8) / 7;
    {
        const double2 w1 = SIMT_LDG(twf + lt);
        const double2 w2 = cmul(w1, w1);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], cmul(w1, w2));
        dft4(v[0], v[1], v[2], v[3]);
    }
    {
        const double2 w1 = SIMT_LDG(twf + jB);
        const double2 w2 = cmul(w1, w1);
        v[5] = cmul(v[5], w1);
        v[6] = cmul(v[6], w2);
        v[7] = cmul(v[7], cmul(w1, w2));
        dft4(v[4], v[5], v[6], v[7]);
    }
}

// Paired arrangement in (slot t = y[jA + t n/4], slot 4 + t = y[jB + t n/4]), standard arrangement out
// (v[m] = Z[lt + m n/8], registers only: nothing is left in shared memory).  tw: the per-stage table of
// fft_line_forward (stages Ns = 4, 32, ..., n/8).
template <int LG>
SIMT_DEVICE void fft_line_pre(double2 (&v)[8], const LineView& lv, int lt, const double2* __restrict__ tw) {
    static_assert(LG % 3 == 2 && LG >= 8, "paired transforms: n = 4 * 8^s, n >= 256");
    constexpr int n8 = 1 << (LG - 3);
    const int lgL = lv.lgL;
    const int rd_step = n8 << lgL;
    const bool rd_fast = (rd_step & 63) == 0;
    const int rd0 = lv.idx(lt);
    const int jB = paired_jB<LG>(lt);
    // leading radix-4 stage (Ns = 1): butterflies jA = lt and jB
    dft4(v[0], v[1], v[2], v[3]);
    dft4(v[4], v[5], v[6], v[7]);
    if (lgL == 0) {
        // one line per unit: this exchange gets a swizzle of its own, chunk ^= (chunk >> 3) & 3.  The descending
        // jB runs (4 (n/4 - lt) + t) put two lanes of every quarter warp on the same bank group under the usual
        // three-bit swizzle (ncu: 4 of the first pass's STS.128 at twice their ideal wavefronts); two bits folded
        // into the slot bits keep jA writes, jB writes and the linear read-back conflict-free
        // (tools/bank_conflicts_paired.py).  n/8 is a multiple of 32 chunks, so the read-back stays linear.
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int ca = (lt << 2) + t, cb = (jB << 2) + t;
            lv.sm[ca ^ ((ca >> 3) & 3)] = v[t];
            lv.sm[cb ^ ((cb >> 3) & 3)] = v[4 + t];
        }
        SIMT_SYNC();
        const int r3 = lt ^ ((lt >> 3) & 3);
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = lv.sm[r3 + m * n8];
        SIMT_SYNC();
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            lv.at((lt << 2) + t) = v[t];
            lv.at((jB << 2) + t) = v[4 + t];
        }
        SIMT_SYNC();
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = lv.sm[rd_fast ? rd0 + m * rd_step : lv.idx(lt + m * n8)];
        SIMT_SYNC();
    }
#pragma unroll
    for (int lgNs = 2; lgNs < LG; lgNs += 3) {
        const int jlo = lt & ((1 << lgNs) - 1);
        r8_twiddle(v, SIMT_LDG(tw + 3 * (((1 << lgNs) - 4) / 7) + jlo));
        dft8(v);
        if (lgNs + 3 < LG) {
            const int base = ((lt >> lgNs) << (lgNs + 3)) | jlo;
            const int cb = (base << lgL) | lv.l;
            const int sh = lgNs + lgL;
            if (sh >= 6) {
                const int w0 = swz(cb, lgL);
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[w0 + (t << sh)] = v[t];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.sm[swz(cb + (t << sh), lgL)] = v[t];
            }
            SIMT_SYNC();
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = lv.sm[rd_fast ? rd0 + m * rd_step : lv.idx(lt + m * n8)];
            SIMT_SYNC();
        }
    }
}

}  // namespace mgvi
