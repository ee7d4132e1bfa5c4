// Small dense symmetric systems (n <= 32) handled by ONE warp: the polynomial model's metric
// (test/test_models/model_polyfit.jl has 5 parameters) and any dense-linear model this small.
// Lane a owns component a of every vector; the matrix lives in shared memory.
#pragma once
#include "field_pass.h"

namespace mgvi {

static const int kSmallMax = 32;

SIMT_DEVICE double warp_sum(double v) {
    for (int m = 16; m >= 1; m >>= 1) v += SIMT_SHFL_XOR_F64(v, m);
    return v;
}

SIMT_DEVICE double warp_bcast(double v, int src, double* scratch) {
    // broadcast through shared scratch (one double per warp); all lanes call
    const int lane = SIMT_TID & 31;
    if (lane == src) *scratch = v;
    SIMT_SYNC();
    const double r = *scratch;
    SIMT_SYNC();
    return r;
}

// y = M v for an n x n row-major matrix in shared memory; lane a returns y[a].
// vs: shared staging of v (n doubles).  All 32 lanes call (blocks are exactly one warp).
SIMT_DEVICE double small_matvec(const double* Msh, int n, double v_lane, double* vs) {
    const int lane = SIMT_TID & 31;
    if (lane < n) vs[lane] = v_lane;
    SIMT_SYNC();
    double y = 0.0;
    if (lane < n)
        for (int b = 0; b < n; ++b) y += Msh[lane * n + b] * vs[b];
    SIMT_SYNC();
    return y;
}

// ---------------------------------------------------------------------------------------------
// Krylov.jl cg! for one right-hand side per block (one warp per block).
// ---------------------------------------------------------------------------------------------
struct SmallCgParams {
    const double* M;      // n x n (row-major; symmetric)
    const double* B;      // n x nvec right-hand sides (column per vector)
    double* X;            // n x nvec solutions
    int* iters;           // nvec
    double* rnorm;        // nvec
    int n;
    double atol, rtol;
    long long itmax;
};

SIMT_DEVICE void small_cg_body(const SmallCgParams& P, unsigned char* smem) {
    double* Msh = reinterpret_cast<double*>(smem);
    double* vs = Msh + kSmallMax * kSmallMax;
    const int lane = SIMT_TID & 31, n = P.n;
    const long long s = SIMT_BID;
    for (int i = lane; i < n * n; i += 32) Msh[i] = P.M[i];
    SIMT_SYNC();
    double x = 0.0;
    double r = lane < n ? P.B[s * n + lane] : 0.0;
    double p = r;
    double gamma = warp_sum(r * r);
    double rn = sqrt(gamma);
    long long it = 0;
    if (gamma != 0.0) {
        const double eps = P.atol + P.rtol * rn;
        bool solved = rn <= eps;
        while (!solved && it < P.itmax) {
            const double Ap = small_matvec(Msh, n, p, vs);
            const double pAp = warp_sum(p * Ap);
            const double alpha = gamma / pAp;
            x += alpha * p;
            r -= alpha * Ap;
            const double gnext = warp_sum(r * r);
            rn = sqrt(gnext);
            solved = (rn <= eps) || (rn + 1.0 <= 1.0);
            if (!solved) {
                const double beta = gnext / gamma;
                p = r + beta * p;
            }
            gamma = gnext;
            ++it;
        }
    }
    if (lane < n) P.X[s * n + lane] = x;
    if (lane == 0) {
        if (P.iters) P.iters[s] = (int)it;
        if (P.rnorm) P.rnorm[s] = rn;
    }
}

// ---------------------------------------------------------------------------------------------
// IterativeSolvers cg_iterator!, one update per launch; state in global memory:
// st[0] = residual, st[1] = prev_residual, st[2] = tol, st[3] = active (0/1), st[4] = iteration
// ---------------------------------------------------------------------------------------------
struct SmallNcgParams {
    const double* M;
    const double* g;     // right-hand side (init)
    double* x; double* r; double* u;
    double* st;
    int n, init;
    long long maxiter;
};

SIMT_DEVICE void small_ncg_body(const SmallNcgParams& P, unsigned char* smem) {
    double* Msh = reinterpret_cast<double*>(smem);
    double* vs = Msh + kSmallMax * kSmallMax;
    const int lane = SIMT_TID & 31, n = P.n;
    if (P.init) {
        const double b = lane < n ? P.g[lane] : 0.0;
        if (lane < n) { P.x[lane] = 0.0; P.r[lane] = b; P.u[lane] = 0.0; }
        const double res = sqrt(warp_sum(b * b));
        if (lane == 0) {
            const double tol = fmax(1.4901161193847656e-08 * res, 0.0);
            P.st[0] = res; P.st[1] = 1.0; P.st[2] = tol; P.st[4] = 0.0;
            P.st[3] = (!(res <= tol) && P.maxiter > 0) ? 1.0 : 0.0;
        }
        return;
    }
    for (int i = lane; i < n * n; i += 32) Msh[i] = P.M[i];
    SIMT_SYNC();
    const double res = P.st[0], prev = P.st[1];
    double x = lane < n ? P.x[lane] : 0.0, r = lane < n ? P.r[lane] : 0.0, u = lane < n ? P.u[lane] : 0.0;
    const double beta = (res * res) / (prev * prev);
    u = r + beta * u;
    const double c = small_matvec(Msh, n, u, vs);
    const double alpha = (res * res) / warp_sum(u * c);
    x += alpha * u;
    r -= alpha * c;
    const double rnew = sqrt(warp_sum(r * r));
    if (lane < n) { P.x[lane] = x; P.r[lane] = r; P.u[lane] = u; }
    if (lane == 0) {
        const double it = P.st[4] + 1.0;
        P.st[1] = res; P.st[0] = rnew; P.st[4] = it;
        P.st[3] = (!(rnew <= P.st[2]) && it < (double)P.maxiter) ? 1.0 : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// MatrixInversion sampler for a small metric (src/residual_samplers.jl:95-123): the LOWER Cholesky
// factor of M^-1 without forming the inverse: with P the index reversal,
// chol(P M P) = L', U = P L' P is upper with M = U U', and L_Sigma = U^-T.  Then R = L_Sigma Z.
// One block (one warp); O(n^3) serial work is irrelevant at n <= 32.
// ---------------------------------------------------------------------------------------------
struct SmallDenseParams {
    const double* M;     // n x n
    const double* Z;     // n x nvec
    double* R;           // n x nvec
    double* Lout;        // n x n column-major (Julia layout) or null
    int n, nvec;
};

SIMT_DEVICE void small_dense_body(const SmallDenseParams& P, unsigned char* smem) {
    double* A = reinterpret_cast<double*>(smem);                 // P M P, then its lower Cholesky L'
    double* Ls = A + kSmallMax * kSmallMax;                      // L_Sigma (row-major)
    const int lane = SIMT_TID & 31, n = P.n;
    for (int i = lane; i < n * n; i += 32) {
        const int a = i / n, b = i % n;
        A[a * n + b] = P.M[(n - 1 - a) * n + (n - 1 - b)];
    }
    SIMT_SYNC();
    if (lane == 0) {
        // in-place lower Cholesky of A (serial, row by row)
        for (int j = 0; j < n; ++j) {
            double d = A[j * n + j];
            for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
            d = sqrt(d);
            A[j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double t = A[i * n + j];
                for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
                A[i * n + j] = t / d;
            }
        }
        // U[a][b] = L'[n-1-a][n-1-b] (upper).  L_Sigma = U^-T is lower: solve U' L_Sigma = I, i.e.
        // forward substitution with the lower matrix U' (U'[a][b] = U[b][a]).
        for (int col = 0; col < n; ++col) {
            for (int a = 0; a < n; ++a) {
                double t = (a == col) ? 1.0 : 0.0;
                for (int b = 0; b < a; ++b) {
                    const double Uba = A[(n - 1 - b) * n + (n - 1 - a)];
                    t -= Uba * Ls[b * n + col];
                }
                const double Uaa = A[(n - 1 - a) * n + (n - 1 - a)];
                Ls[a * n + col] = (a >= col) ? t / Uaa : 0.0;
            }
        }
    }
    SIMT_SYNC();
    if (P.Lout)
        for (int i = lane; i < n * n; i += 32) {
            const int a = i / n, b = i % n;
            P.Lout[b * n + a] = Ls[a * n + b];
        }
    for (int v = 0; v < P.nvec; ++v) {
        if (lane < n) {
            double t = 0.0;
            for (int b = 0; b <= lane; ++b) t += Ls[lane * n + b] * P.Z[(long long)v * n + b];
            P.R[(long long)v * n + lane] = t;
        }
    }
}

}  // namespace mgvi
