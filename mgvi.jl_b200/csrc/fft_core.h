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
// side-by-side lines, XOR-swizzled on the chunk index (bits 0..2 ^= bits 3..5) -- the same
// pattern as the TMA 128-byte swizzle -- which keeps every stage's scattered writes and linear
// reads free of bank conflicts.
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

SIMT_HD int swz(int chunk) { return chunk ^ ((chunk >> 3) & 7); }

// 4-point forward DFT in place (natural order out).
SIMT_HD void dft4(double2& c0, double2& c1, double2& c2, double2& c3) {
    double2 d0 = cadd(c0, c2), d1 = csub(c0, c2), d2 = cadd(c1, c3), d3 = cmul_mi(csub(c1, c3));
    c0 = cadd(d0, d2);
    c1 = cadd(d1, d3);
    c2 = csub(d0, d2);
    c3 = csub(d1, d3);
}

// 8-point forward DFT in place (natural order out).
SIMT_HD void dft8(double2 (&a)[8]) {
    const double h = 0.70710678118654752440;
    double2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    double2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    double2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
    double2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    // odd branch twiddles W8^1 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
    b5 = make_double2((b5.x + b5.y) * h, (b5.y - b5.x) * h);
    b6 = cmul_mi(b6);
    b7 = make_double2((b7.y - b7.x) * h, -(b7.x + b7.y) * h);
    dft4(b0, b1, b2, b3);
    dft4(b4, b5, b6, b7);
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = b4; a[3] = b5; a[5] = b6; a[7] = b7;
}

// Shared-memory view of one complex line among L side-by-side lines.
struct LineView {
    double2* sm;   // base of the unit's buffer (L*n chunks)
    int L, l;      // lines side by side, my line
    SIMT_DEVICE double2& at(int k) const { return sm[swz(k * L + l)]; }
};

// Forward FFT of the line whose elements the calling threads hold in v (v[m] = x[lt + m*n/8]).
// On return the line sits in shared memory in natural order (all threads have passed the final
// barrier).  EVERY thread of the block must call this (uniform barriers); threads whose unit is
// out of range pass valid = false and only take part in the barriers.
// n >= 8, power of two; tw[k] = exp(-2 pi i k / n), k < n.
SIMT_DEVICE void fft_line_forward(double2 (&v)[8], const LineView& lv, int lt, int n, int log2n,
                                  const double2* __restrict__ tw, bool valid) {
    const int n8 = n >> 3;
    int Ns = 1;
    const int lead = log2n % 3;   // 0: all radix-8; 1: leading radix-2; 2: leading radix-4
    bool first = true;
    while (Ns < n) {
        const int R = (first && lead == 1) ? 2 : ((first && lead == 2) ? 4 : 8);
        if (valid) {
            if (R == 8) {
                const int j = lt;
                if (Ns > 1) {
                    const int tstride = n / (Ns * 8);
                    const int jm = (j & (Ns - 1)) * tstride;
#pragma unroll
                    for (int t = 1; t < 8; ++t) v[t] = cmul(v[t], SIMT_LDG(&tw[jm * t]));
                }
                dft8(v);
                const int base = (j / Ns) * Ns * 8 + (j & (Ns - 1));
#pragma unroll
                for (int t = 0; t < 8; ++t) lv.at(base + t * Ns) = v[t];
            } else if (R == 4) {
                // two butterflies: j = lt + q*n8, inputs v[q + 2t] (always the leading stage: Ns = 1)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    dft4(v[q], v[q + 2], v[q + 4], v[q + 6]);
                    const int j = lt + q * n8;
#pragma unroll
                    for (int t = 0; t < 4; ++t) lv.at(j * 4 + t) = v[q + 2 * t];
                }
            } else {
                // four butterflies: j = lt + q*n8, inputs v[q], v[q + 4] (leading stage: Ns = 1)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double2 s = cadd(v[q], v[q + 4]), d = csub(v[q], v[q + 4]);
                    const int j = lt + q * n8;
                    lv.at(j * 2) = s;
                    lv.at(j * 2 + 1) = d;
                }
            }
        }
        SIMT_SYNC();
        Ns *= R;
        first = false;
        if (Ns < n) {
            if (valid) {
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = lv.at(lt + m * n8);
            }
            SIMT_SYNC();
        }
    }
}

// After fft_line_forward: v[m] <- 2 * ( H(a)[k], H(b)[k] ), k = lt + m*n/8.
// (The factor 2 is folded into the callers' scale constants; powers of two are exact.)
SIMT_DEVICE void dht_combine_read(double2 (&v)[8], const LineView& lv, int lt, int n) {
    const int n8 = n >> 3;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = lt + m * n8;
        const double2 zk = lv.at(k);
        const double2 zn = lv.at((n - k) & (n - 1));
        v[m] = make_double2(zk.x - zk.y + zn.x + zn.y, zk.x + zk.y + zn.y - zn.x);
    }
}

}  // namespace mgvi
