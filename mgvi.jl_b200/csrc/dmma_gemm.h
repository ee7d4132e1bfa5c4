// FP64 tensor-core GEMM for the dense Fisher path (BASELINE.json configs[3]; reference
// src/residual_samplers.jl:95-109 materialises J'FJ + I by n_theta operator applications):
//
//     C[M x N] = alpha * op(A) * op(B) + beta * C          (C column-major, leading dimension ldc)
//
// with A(a, i) = A[a * a_row + i * a_k] and B(i, v) = B[i * b_k + v * b_col]: the reduction index may
// be the contiguous one ("T" operand, e.g. J' in J'FJ where columns of J are contiguous) or the
// strided one ("N" operand).  The inner product runs on the FP64 tensor pipe with
// mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA) -- tcgen05 has no FP64 kind, so on sm_100a this
// legacy path IS the FP64 tensor path.  Block tile 128 x 128 x 16, 256 threads = 8 warps as 2 x 4,
// warp tile 64 x 32 (8 x 4 mma tiles, 64 accumulators per thread); the next K-slab is fetched
// into registers while the current one is multiplied.  `upper_only` skips block tiles strictly
// below the diagonal (symmetric rank-k updates: J'FJ and the Cholesky trailing update).
#pragma once
#include "field_pass.h"

namespace mgvi {

struct GemmParams {
    const double* A; long long a_row, a_k;
    const double* B; long long b_k, b_col;
    double* C; long long ldc;
    int M, N, K;
    double alpha, beta;
    int upper_only;     // compute only block tiles with tile_col >= tile_row
    int ntm, ntn;
    const double* kweight;   // optional diagonal weight along the reduction index, folded into the A-operand
                             // load: C = alpha * op(A) diag(kweight) op(B) + beta C  (the Fisher diagonal of J'FJ)
};

static const int kGemmBM = 128, kGemmBN = 128, kGemmBK = 16, kGemmLd = 20;   // row stride 20 doubles: 2-way at most

SIMT_DEVICE void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#else
    // emulator: lane holds A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
    // C[row = lane/4][col = 2*(lane%4) + {0,1}]
    const int lane = SIMT_TID & 31, row = lane >> 2, c2 = (lane & 3) * 2;
    double wa[32], wb[32];
    simt_emu::warp_allgather2(a, b, wa, wb);
    for (int k = 0; k < 4; ++k) {
        c0 += wa[row * 4 + k] * wb[c2 * 4 + k];
        c1 += wa[row * 4 + k] * wb[(c2 + 1) * 4 + k];
    }
#endif
}

// Global -> registers for one 128 x 16 operand slab.  `kcontig`: the K index is the contiguous one.
// Each thread fetches 8 doubles: kcontig: (row = t/8 + 32 r, k = 2 (t%8) + {0,1}), r = 0..3
//                                 else   : (row = 2 (t%64) + {0,1}, k = t/64 + 4 r),   r = 0..3
template <bool KW>
SIMT_DEVICE void gemm_fetch(double (&reg)[8], const double* __restrict__ X, long long s_row, long long s_k, int row0,
                            int k0, int nrow, int nk, bool kcontig, const double* __restrict__ kweight = nullptr) {
    const int t = SIMT_TID;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int row, k;
            if (kcontig) { row = (t >> 3) + 32 * r; k = 2 * (t & 7) + e; }
            else { row = 2 * (t & 63) + e; k = (t >> 6) + 4 * r; }
            const bool ok = (row0 + row < nrow) && (k0 + k < nk);
            double xv = ok ? X[(long long)(row0 + row) * s_row + (long long)(k0 + k) * s_k] : 0.0;
            if (KW) { if (ok) xv *= SIMT_LDG(kweight + k0 + k); }
            reg[2 * r + e] = xv;
        }
    }
}

SIMT_DEVICE void gemm_stash(const double (&reg)[8], double* sm, bool kcontig) {
    const int t = SIMT_TID;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int row, k;
            if (kcontig) { row = (t >> 3) + 32 * r; k = 2 * (t & 7) + e; }
            else { row = 2 * (t & 63) + e; k = (t >> 6) + 4 * r; }
            sm[row * kGemmLd + k] = reg[2 * r + e];
        }
    }
}

template <bool KW>
SIMT_DEVICE void gemm_body(const GemmParams& P, unsigned char* smem) {
    double* As = reinterpret_cast<double*>(smem);
    double* Bs = As + kGemmBM * kGemmLd;
    // block -> tile (column-major over tiles so that consecutive blocks share the B slab)
    long long bid = SIMT_BID;
    const int tm = (int)(bid % P.ntm), tn = (int)(bid / P.ntm);
    if (P.upper_only && tn < tm) return;
    const int row0 = tm * kGemmBM, col0 = tn * kGemmBN;
    const int warp = SIMT_TID >> 5, lane = SIMT_TID & 31;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const bool a_kc = (P.a_k == 1), b_kc = (P.b_k == 1);
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    double ra[8], rb[8];
    gemm_fetch<KW>(ra, P.A, P.a_row, P.a_k, row0, 0, P.M, P.K, a_kc, P.kweight);
    gemm_fetch<false>(rb, P.B, P.b_col, P.b_k, col0, 0, P.N, P.K, b_kc);
    for (int k0 = 0; k0 < P.K; k0 += kGemmBK) {
        gemm_stash(ra, As, a_kc);
        gemm_stash(rb, Bs, b_kc);
        SIMT_SYNC();
        if (k0 + kGemmBK < P.K) {
            gemm_fetch<KW>(ra, P.A, P.a_row, P.a_k, row0, k0 + kGemmBK, P.M, P.K, a_kc, P.kweight);
            gemm_fetch<false>(rb, P.B, P.b_col, P.b_k, col0, k0 + kGemmBK, P.N, P.K, b_kc);
        }
#pragma unroll
        for (int kk = 0; kk < kGemmBK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = As[(wm + 8 * i + (lane >> 2)) * kGemmLd + kk + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn + 8 * j + (lane >> 2)) * kGemmLd + kk + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        SIMT_SYNC();
    }
    // epilogue: lane holds C[row = lane/4][col = 2 (lane%4) + {0,1}] of each 8 x 8 tile
    // The old values of C are fetched eight at a time BEFORE the stores of that group: element by element the
    // compiler must keep every load behind the previous store (it cannot prove the addresses distinct), and 64
    // dependent round trips cost more than the whole K loop of the rank-64 updates of the blocked Cholesky.
    double* __restrict__ Cw = P.C;
    const double* __restrict__ Cr = P.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + wm + 8 * i + (lane >> 2);
        double old[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = col0 + wn + 8 * j + 2 * (lane & 3) + e;
                old[j][e] = (P.beta != 0.0 && r < P.M && c < P.N) ? Cr[(long long)c * P.ldc + r] : 0.0;
            }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = col0 + wn + 8 * j + 2 * (lane & 3) + e;
                if (r < P.M && c < P.N) Cw[(long long)c * P.ldc + r] = P.alpha * acc[i][j][e] + P.beta * old[j][e];
            }
    }
}

}  // namespace mgvi
