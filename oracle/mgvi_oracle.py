"""NumPy float64 restatement of the MGVI.jl v0.4.5 hot path (TEST INFRASTRUCTURE, see
oracle/__init__.py for the rules and the pinning status).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Arithmetic that lives in un-vendored third-party Julia packages
(Krylov.jl, IterativeSolvers.jl, LineSearches.jl, Distributions.jl, FFTW r2r/DHT) is
restated from the published algorithm; the call site in the reference is cited instead.

All random draws are INPUTS (standard-normal arrays handed in by the caller), in the
order the reference consumes them: per residual `n1` (n_lambda) then `eta` (n_theta)
(src/residual_samplers.jl:75-76); dense sampler: one (n_theta x n) matrix (:121).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft
import scipy.linalg
import scipy.special

__all__ = [
    "SQRT_EPS", "LAYOUT_MU_ONLY", "LAYOUT_BLOCKED", "LAYOUT_INTERLEAVED",
    "LINK_IDENTITY", "LINK_EXP", "LIKE_NORMAL", "LIKE_POISSON",
    "dht", "dht_matrix",
    "fisher_normal", "fisher_mvnormal_diag", "fisher_mvnormal_full", "fisher_exponential", "fisher_poisson",
    "fisher_blockdiag", "fisher_blockdiag_dense",
    "FieldModel", "PolyModel", "DenseLinearModel", "HyperFieldModel",
    "metric_apply", "sampler_rhs", "krylov_cg", "sample_residuals_cg", "residual_pushfwd_operator",
    "sample_residuals_dense", "mv_normal_logdensity", "posterior_loglike", "mean_neg_log_pstr",
    "kl_gradient", "mean_metric_apply", "itsolvers_cg_iterator", "strong_wolfe", "NewtonCGOpts",
    "NewtonCGTrace", "newtoncg_optimize", "newtoncg_step_record", "build_samples", "mgvi_step", "mgvi_sample",
]

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))  # 1.4901161193847656e-08, LinearSolve default_tol
LOG2PI = math.log(2.0 * math.pi)

LAYOUT_MU_ONLY, LAYOUT_BLOCKED, LAYOUT_INTERLEAVED = 0, 1, 2
LINK_IDENTITY, LINK_EXP = 0, 1
LIKE_NORMAL, LIKE_POISSON = 0, 1

_FFT_WORKERS = 1


def set_fft_workers(n: int) -> None:
    """Threads scipy.fft may use (CPU-baseline timing only)."""
    global _FFT_WORKERS
    _FFT_WORKERS = int(n)


# --------------------------------------------------------------------------------------------
# Discrete Hartley transform -- FFTW `plan_r2r(..., DHT)` (test/test_models/model_fft_gp.jl:21,
# docs/src/advanced_tutorial_lit.jl:208): unnormalised, H[k] = sum_j x[j] (cos + sin)(2 pi j k / n);
# a multi-dimensional r2r plan is the separable product of 1-D DHTs.
# --------------------------------------------------------------------------------------------
def dht(x: np.ndarray, axes=None) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    if axes is None:
        axes = range(x.ndim)
    for ax in axes:
        X = scipy.fft.fft(x, axis=ax, workers=_FFT_WORKERS)
        x = X.real - X.imag
    return x


def dht_matrix(n: int) -> np.ndarray:
    """Dense cas-kernel matrix (O(n^2) definition), used to pin `dht` itself."""
    jk = np.outer(np.arange(n), np.arange(n)) % n
    ang = 2.0 * np.pi * jk / n
    return np.cos(ang) + np.sin(ang)


# --------------------------------------------------------------------------------------------
# Fisher information in distribution-parameter space -- src/information.jl
# (diagonal families only; returned as the diagonal vector, `sqrt` of it is cholesky_L,
#  src/util.jl:24)
# --------------------------------------------------------------------------------------------
def fisher_normal(mu, sigma) -> np.ndarray:
    """src/information.jl:12-18: Normal(mu, sigma) -> diag(1/sigma^2, 2/sigma^2)."""
    inv_s = 1.0 / float(sigma)
    inv_s2 = inv_s * inv_s
    return np.array([inv_s2, 2.0 * inv_s2])


def fisher_mvnormal_diag(mu, var) -> np.ndarray:
    """src/information.jl:60-65: MvNormal with PDiagMat covariance -> [1/var ...; 2/var ...]
    (blocked; the variance block carries the weight of sigma, a reference quirk that parity
    copies)."""
    inv = 1.0 / np.asarray(var, dtype=np.float64)
    return np.concatenate([inv, 2.0 * inv])


def fisher_exponential(theta) -> np.ndarray:
    """src/information.jl:67-72."""
    inv_l = 1.0 / float(theta)
    return np.array([inv_l * inv_l])


def fisher_poisson(lam) -> np.ndarray:
    """src/information.jl:74-78."""
    return np.array([1.0 / float(lam)])


def fisher_mvnormal_full(mu, cov) -> np.ndarray:
    """src/information.jl:20-58: MvNormal with a full covariance.  Flat parameters are
    [mu (d); Sigma[0:i+1, i] for i = 0..d-1] (src/shapes.jl:14-21: columns of the upper triangle),
    an off-diagonal Sigma_ab counting once for both symmetric entries.  Returns the dense
    (d + d(d+1)/2)-square block matrix blockdiag(P, C), P = inv(Sigma),
    C[(a,b),(c,d)] = 1/2 tr(P D_ab P D_cd), D_ab = dSigma/dSigma_ab (E_ab + E_ba, or E_aa) --
    the closed form of the reference's loops (P_ac P_bd + P_ad P_bc, halved once per diagonal
    parameter involved)."""
    cov = np.asarray(cov, dtype=np.float64)
    d = cov.shape[0]
    P = np.linalg.inv(cov)
    idx = [(i, j) for j in range(d) for i in range(j + 1)]
    nc = len(idx)
    C = np.zeros((nc, nc))
    for m, (a, b) in enumerate(idx):
        Dm = np.zeros((d, d)); Dm[a, b] = 1.0; Dm[b, a] = 1.0
        for n, (c, e) in enumerate(idx):
            Dn = np.zeros((d, d)); Dn[c, e] = 1.0; Dn[e, c] = 1.0
            C[m, n] = 0.5 * np.trace(P @ Dm @ P @ Dn)
    out = np.zeros((d + nc, d + nc))
    out[:d, :d] = P
    out[d:, d:] = C
    return out


def fisher_blockdiag_dense(parts) -> np.ndarray:
    """`_blockdiag` (src/custom_linear_maps.jl:84-94) when at least one part is a dense block:
    diagonal parts (1-D arrays) and dense parts (2-D arrays) on the diagonal of one matrix."""
    mats = [np.diag(np.asarray(p, dtype=np.float64)) if np.ndim(p) == 1 else np.asarray(p, dtype=np.float64)
            for p in parts]
    n = sum(m.shape[0] for m in mats)
    out = np.zeros((n, n))
    o = 0
    for m in mats:
        k = m.shape[0]
        out[o:o + k, o:o + k] = m
        o += k
    return out


def fisher_blockdiag(parts) -> np.ndarray:
    """src/information.jl:80-96 + src/custom_linear_maps.jl:84-94: Product /
    ProductDistribution / NamedTupleDist -> concatenation of the parts' diagonals, in order."""
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


# --------------------------------------------------------------------------------------------
# lambda layout helpers (src/shapes.jl:4-14,40-54 -- `flat_params` is a layout contract)
# --------------------------------------------------------------------------------------------
def _mu_index(n: int, layout: int) -> tuple[np.ndarray, int]:
    """Rows of the flat lambda vector that hold the n mean parameters, and n_lambda."""
    if layout == LAYOUT_MU_ONLY:
        return np.arange(n), n
    if layout == LAYOUT_BLOCKED:        # MvNormal{PDiagMat}: [mu(n); var(n)]  (shapes.jl:7,14,40-41)
        return np.arange(n), 2 * n
    if layout == LAYOUT_INTERLEAVED:    # Product of Normal: [mu1, s1, mu2, s2, ...] (shapes.jl:48-50)
        return 2 * np.arange(n), 2 * n
    raise ValueError("bad layout")


def _noise_index(n: int, layout: int) -> np.ndarray:
    if layout == LAYOUT_BLOCKED:
        return n + np.arange(n)
    if layout == LAYOUT_INTERLEAVED:
        return 2 * np.arange(n) + 1
    return np.zeros(0, dtype=np.int64)


@dataclass
class Linearization:
    """What `_fisher_information_and_jac` returns (src/residual_samplers.jl:4-8): the diagonal
    Fisher information at lambda(xi) and the Jacobian d lambda / d xi as a pair of closures
    (`with_jacobian(flat_params o model, xi, LinearMap, ad)`: J*v is the directional derivative
    at the fixed point xi, J'*l the pullback at the same xi)."""
    n_theta: int
    n_lambda: int
    fisher: np.ndarray            # (n_lambda,)
    jvp: callable                 # v (n_theta) -> (n_lambda)
    vjp: callable                 # l (n_lambda) -> (n_theta)
    dense: callable               # () -> (n_lambda, n_theta)


# --------------------------------------------------------------------------------------------
# Model families
# --------------------------------------------------------------------------------------------
class FieldModel:
    """Standardised correlated field: s = c * H(A . xi) + offset, lambda = g(s).

    Follows test/test_models/model_fft_gp.jl:60-72 (A = 50/(k^2.3+1), c = 1/n via inv(plan),
    exp link, Normal(mean, 0.4)) and docs/src/advanced_tutorial_lit.jl:253-256,275-277,301-307
    (A = sqrt_kernel, c = _GP_HARMONIC_DIST, exp link, Poisson).  `dims` is in C order
    (dims[0] slowest); a Julia caller passes reversed `size(field)`.
    """

    def __init__(self, dims, amplitude, scale, data, *, offset=0.0, link=LINK_EXP,
                 likelihood=LIKE_POISSON, sigma=1.0, layout=LAYOUT_MU_ONLY):
        self.dims = tuple(int(d) for d in dims)
        self.N = int(np.prod(self.dims))
        self.A = np.asarray(amplitude, dtype=np.float64).reshape(self.N)
        self.c = float(scale)
        self.offset = float(offset)
        self.link = int(link)
        self.likelihood = int(likelihood)
        self.sigma = float(sigma)
        self.data = np.asarray(data, dtype=np.float64).reshape(self.N)
        if self.likelihood == LIKE_POISSON:
            layout = LAYOUT_MU_ONLY        # Poisson has one parameter (information.jl:74-78)
        self.layout = int(layout)
        self.mu_idx, self.n_lambda = _mu_index(self.N, self.layout)
        self.noise_idx = _noise_index(self.N, self.layout)
        self.n_theta = self.N

    # -- forward pieces -------------------------------------------------------------------
    def _H(self, x):
        return dht(x.reshape(self.dims)).reshape(self.N)

    def signal(self, xi):
        return self.c * self._H(self.A * xi) + self.offset

    def rate(self, xi):
        s = self.signal(xi)
        return np.exp(s) if self.link == LINK_EXP else s

    def lam(self, xi):
        """flat_params(model(xi)) (src/shapes.jl:40-54)."""
        out = np.empty(self.n_lambda)
        out[self.mu_idx] = self.rate(xi)
        out[self.noise_idx] = self.sigma if self.layout == LAYOUT_INTERLEAVED else self.sigma ** 2
        return out

    def _gprime(self, lam_mu):
        return lam_mu if self.link == LINK_EXP else np.ones_like(lam_mu)

    def _fisher_mu(self, lam_mu):
        if self.likelihood == LIKE_POISSON:
            return 1.0 / lam_mu                          # information.jl:74-78
        inv_s = 1.0 / self.sigma
        return np.full(self.N, inv_s * inv_s)            # information.jl:12-18 / 60-65

    def metric_weight(self, xi):
        """w = F_mu . g'(s)^2, the pixel-space diagonal of J'FJ = c^2 A H w H A."""
        lam_mu = self.rate(xi)
        gp = self._gprime(lam_mu)
        return self._fisher_mu(lam_mu) * gp * gp

    def linearize(self, xi) -> Linearization:
        lam_mu = self.rate(xi)
        gp = self._gprime(lam_mu)
        F = np.empty(self.n_lambda)
        F[self.mu_idx] = self._fisher_mu(lam_mu)
        if self.noise_idx.size:
            inv_s = 1.0 / self.sigma
            F[self.noise_idx] = 2.0 * (inv_s * inv_s)
        mu_idx, n_lambda, N = self.mu_idx, self.n_lambda, self.N

        def jvp(v):
            out = np.zeros(n_lambda)
            out[mu_idx] = gp * (self.c * self._H(self.A * v))
            return out

        def vjp(l):
            return self.c * (self.A * self._H(gp * l[mu_idx]))

        def dense():
            J = np.zeros((n_lambda, N))
            eye = np.eye(N)
            for j in range(N):
                J[:, j] = jvp(eye[:, j])
            return J

        return Linearization(N, n_lambda, F, jvp, vjp, dense)

    # -- likelihood -----------------------------------------------------------------------
    def loglike(self, xi):
        """logdensityof(model(xi), data) (src/mgvi_impl.jl:10); closed forms of Distributions.jl
        `logpdf` for Normal / Poisson."""
        lam_mu = self.rate(xi)
        d = self.data
        if self.likelihood == LIKE_POISSON:
            s = self.signal(xi) if self.link == LINK_EXP else np.log(lam_mu)
            return float(np.sum(d * s - lam_mu - scipy.special.gammaln(d + 1.0)))
        z = (d - lam_mu) / self.sigma
        return float(np.sum(-0.5 * z * z) - self.N * (math.log(self.sigma) + 0.5 * LOG2PI))

    def grad_loglike(self, xi):
        """J(xi)' * d loglike / d lambda (what Zygote returns for the likelihood term)."""
        lam_mu = self.rate(xi)
        gp = self._gprime(lam_mu)
        d = self.data
        if self.likelihood == LIKE_POISSON:
            dl = d / lam_mu - 1.0
        else:
            dl = (d - lam_mu) / (self.sigma * self.sigma)
        return self.c * (self.A * self._H(gp * dl))


class PolyModel:
    """Polynomial fit with a shared, parameterised noise level -- test/test_models/model_polyfit.jl.

    mean_g(x) = sum_k coef_scale[k] * p[k] * x^k   (model_polyfit.jl:15-17: scales 10, 40, 600, 80)
    sigma     = noise_scale * p[P]^2                (model_polyfit.jl:20-21: 60 * p5^2)
    The model is a NamedTupleDist of one product of Normals per grid (model_polyfit.jl:19-24).
    `layout` picks which concrete Distributions type `product_distribution(Normal.(...))`
    yields (SURVEY.md note N1): BLOCKED = MvNormal{PDiagMat} -> [mu_g; var_g] per group,
    INTERLEAVED = Product{Normal} -> [mu1, sigma1, mu2, sigma2, ...] per group.
    """

    def __init__(self, x_groups, data, *, coef_scale=(10.0, 40.0, 600.0, 80.0), noise_scale=60.0,
                 layout=LAYOUT_BLOCKED):
        self.x_groups = [np.asarray(x, dtype=np.float64) for x in x_groups]
        self.sizes = [x.size for x in self.x_groups]
        self.x = np.concatenate(self.x_groups)
        self.N = self.x.size
        self.coef_scale = np.asarray(coef_scale, dtype=np.float64)
        self.P = self.coef_scale.size
        self.noise_scale = float(noise_scale)
        self.data = np.asarray(data, dtype=np.float64).reshape(self.N)
        if layout not in (LAYOUT_BLOCKED, LAYOUT_INTERLEAVED):
            raise ValueError("PolyModel needs a noise parameter layout")
        self.layout = int(layout)
        self.n_theta = self.P + 1
        self.n_lambda = 2 * self.N
        mu_idx, noise_idx, off, goff = [], [], 0, 0
        for n in self.sizes:
            mi, _ = _mu_index(n, self.layout)
            mu_idx.append(off + mi)
            noise_idx.append(off + _noise_index(n, self.layout))
            off += 2 * n
        self.mu_idx = np.concatenate(mu_idx)
        self.noise_idx = np.concatenate(noise_idx)
        # design matrix X[i, k] = coef_scale[k] * x_i^k
        self.X = np.stack([self.coef_scale[k] * self.x ** k for k in range(self.P)], axis=1)

    def mean(self, p):
        x = self.x
        out = np.zeros_like(x)
        for k in range(self.P):
            out = out + p[k] * self.coef_scale[k] * x ** k
        return out

    def sigma(self, p):
        return p[self.P] ** 2 * self.noise_scale

    def lam(self, p):
        out = np.empty(self.n_lambda)
        out[self.mu_idx] = self.mean(p)
        s = self.sigma(p)
        out[self.noise_idx] = s if self.layout == LAYOUT_INTERLEAVED else s * s
        return out

    def linearize(self, p) -> Linearization:
        s = self.sigma(p)
        inv = 1.0 / (s * s)
        F = np.empty(self.n_lambda)
        F[self.mu_idx] = inv
        F[self.noise_idx] = 2.0 * inv
        ds_dp = 2.0 * self.noise_scale * p[self.P]             # d sigma / d p_P
        dn_dp = ds_dp if self.layout == LAYOUT_INTERLEAVED else 2.0 * s * ds_dp   # d(sigma^2)/dp
        J = np.zeros((self.n_lambda, self.n_theta))
        J[self.mu_idx, : self.P] = self.X
        J[self.noise_idx, self.P] = dn_dp
        return Linearization(self.n_theta, self.n_lambda, F, lambda v: J @ v, lambda l: J.T @ l,
                             lambda: J.copy())

    def loglike(self, p):
        mu, s, d = self.mean(p), self.sigma(p), self.data
        z = (d - mu) / s
        return float(np.sum(-0.5 * z * z) - self.N * (math.log(abs(s)) + 0.5 * LOG2PI))

    def grad_loglike(self, p):
        mu, s, d = self.mean(p), self.sigma(p), self.data
        r = d - mu
        g = np.empty(self.n_theta)
        g[: self.P] = self.X.T @ (r / (s * s))
        g[self.P] = np.sum(r * r / s ** 3 - 1.0 / s) * (2.0 * self.noise_scale * p[self.P])
        return g


class DenseLinearModel:
    """mu = Jm @ xi, Normal(mu, sigma) with constant or per-datum sigma -- BASELINE.json configs[3]; the
    operator type `DenseMatrix` that `_get_operator_type(::MatrixInversion)` selects
    (src/residual_samplers.jl:56)."""

    def __init__(self, Jm, data, *, sigma=0.5, layout=LAYOUT_MU_ONLY):
        self.Jm = np.asarray(Jm, dtype=np.float64)
        self.m, self.k = self.Jm.shape
        self.N = self.m
        self.sigma = float(sigma) if np.ndim(sigma) == 0 else np.asarray(sigma, dtype=np.float64).reshape(self.m)
        self.data = np.asarray(data, dtype=np.float64).reshape(self.m)
        self.layout = int(layout)
        self.mu_idx, self.n_lambda = _mu_index(self.m, self.layout)
        self.noise_idx = _noise_index(self.m, self.layout)
        self.n_theta = self.k
        self.log_sigma_sum = float(np.sum(np.log(np.broadcast_to(self.sigma, (self.m,)))))

    def lam(self, xi):
        out = np.empty(self.n_lambda)
        out[self.mu_idx] = self.Jm @ xi
        if self.noise_idx.size:
            out[self.noise_idx] = self.sigma if self.layout == LAYOUT_INTERLEAVED else self.sigma ** 2
        return out

    def linearize(self, xi) -> Linearization:
        inv_s = 1.0 / self.sigma
        F = np.empty(self.n_lambda)
        F[self.mu_idx] = inv_s * inv_s
        if self.noise_idx.size:
            F[self.noise_idx] = 2.0 * (inv_s * inv_s)
        mu_idx, n_lambda = self.mu_idx, self.n_lambda

        def jvp(v):
            out = np.zeros(n_lambda)
            out[mu_idx] = self.Jm @ v
            return out

        def dense():
            J = np.zeros((n_lambda, self.k))
            J[mu_idx, :] = self.Jm
            return J

        return Linearization(self.k, n_lambda, F, jvp, lambda l: self.Jm.T @ l[mu_idx], dense)

    def loglike(self, xi):
        z = (self.data - self.Jm @ xi) / self.sigma
        return float(np.sum(-0.5 * z * z) - self.log_sigma_sum - self.m * 0.5 * LOG2PI)

    def grad_loglike(self, xi):
        return self.Jm.T @ ((self.data - self.Jm @ xi) / (self.sigma * self.sigma))


class HyperFieldModel:
    """The tutorial's model (docs/src/advanced_tutorial_lit.jl:194-307): a 1-D Gaussian process whose
    amplitude spectrum depends on two hyperparameters, exponentiated, integrated over data bins and
    observed through Poisson counts.

        theta = [p1, p2, xi (N)]                               (PARIDX :153, hyperparameters FIRST)
        ampl  = a0 exp(a1 p1),  corrlen = l0 exp(l1 p2)        (sqrt_kernel :198-203: a0 = 2 GRAIN, a1 = 0.9,
                                                                l0 = 12 / GRAIN^0.3, l1 = 1/15)
        A[k]  = ampl sqrt(2 pi corrlen) exp(-pi^2 d_k^2 corrlen^2)      (amplitude_spectrum :194-196)
        s     = c H(A . xi),  c = _GP_HARMONIC_DIST            (gp_sample :253-256)
        fine  = exp(s)                                         (poisson_gp_link :275-277)
        Lambda_b = binsize sum_{j = idx_b .. idx_b + grain - 1} fine[j]   (agg_lambdas :286-294)
        data_b ~ Poisson(Lambda_b)                             (model :301-307), Fisher 1/Lambda_b.

    `data_idxs` are 0-based start indices of the data bins."""

    def __init__(self, kdist, scale, data_idxs, grain, binsize, data, *, a0, a1, l0, l1):
        self.kd = np.asarray(kdist, dtype=np.float64)
        self.N = self.kd.size
        self.c = float(scale)
        self.idx = np.asarray(data_idxs, dtype=np.int64)
        self.grain = int(grain)
        self.binsize = float(binsize)
        self.data = np.asarray(data, dtype=np.float64).reshape(self.idx.size)
        self.a0, self.a1, self.l0, self.l1 = float(a0), float(a1), float(l0), float(l1)
        self.n_theta = self.N + 2
        self.n_lambda = self.idx.size
        self.cols = self.idx[:, None] + np.arange(self.grain)[None, :]        # fine pixels of every bin

    def amplitude(self, p):
        ampl = self.a0 * math.exp(self.a1 * p[0])
        cl = self.l0 * math.exp(self.l1 * p[1])
        A = ampl * math.sqrt(2.0 * math.pi * cl) * np.exp(-(math.pi ** 2) * self.kd ** 2 * cl ** 2)
        dA1 = self.a1 * A                                                     # dA/dp1
        dA2 = A * (self.l1 * (0.5 - 2.0 * math.pi ** 2 * self.kd ** 2 * cl ** 2))   # dA/dp2
        return A, dA1, dA2

    def fine_rate(self, p):
        A, _, _ = self.amplitude(p)
        return np.exp(self.c * dht(A * p[2:]))

    def agg(self, fine):
        return self.binsize * fine[self.cols].sum(axis=1)

    def lam(self, p):
        return self.agg(self.fine_rate(np.asarray(p, dtype=np.float64)))

    def linearize(self, p) -> Linearization:
        p = np.asarray(p, dtype=np.float64)
        A, dA1, dA2 = self.amplitude(p)
        xi = p[2:]
        fine = np.exp(self.c * dht(A * xi))
        Lam = self.agg(fine)
        F = 1.0 / Lam                                                         # information.jl:74-78
        N, nb = self.N, self.n_lambda

        def jvp(v):
            t = self.c * dht(A * v[2:] + (v[0] * dA1 + v[1] * dA2) * xi)
            return self.agg(fine * t)

        def vjp(l):
            z = np.zeros(N)
            z[self.cols] = (self.binsize * l)[:, None] * fine[self.cols]
            g = self.c * dht(z)
            out = np.empty(N + 2)
            out[0] = float(np.sum(dA1 * xi * g))
            out[1] = float(np.sum(dA2 * xi * g))
            out[2:] = A * g
            return out

        def dense():
            J = np.zeros((nb, N + 2))
            eye = np.eye(N + 2)
            for j in range(N + 2):
                J[:, j] = jvp(eye[:, j])
            return J

        return Linearization(N + 2, nb, F, jvp, vjp, dense)

    def loglike(self, p):
        Lam = self.lam(p)
        d = self.data
        return float(np.sum(d * np.log(Lam) - Lam - scipy.special.gammaln(d + 1.0)))

    def grad_loglike(self, p):
        p = np.asarray(p, dtype=np.float64)
        lin = self.linearize(p)
        Lam = 1.0 / lin.fisher
        return lin.vjp(self.data / Lam - 1.0)


# --------------------------------------------------------------------------------------------
# Metric and residual samplers -- src/residual_samplers.jl
# --------------------------------------------------------------------------------------------
def metric_apply(lin: Linearization, v: np.ndarray) -> np.ndarray:
    """(J' * F * J + I) * v  (src/residual_samplers.jl:72, src/mgvi_impl.jl:147-150); LinearMaps
    evaluates the composite right to left and adds the identity term last."""
    return lin.vjp(lin.fisher * lin.jvp(v)) + v


def sampler_rhs(lin: Linearization, n1: np.ndarray, eta: np.ndarray) -> np.ndarray:
    """J' * (cholesky_L(F) * n1) + eta  (src/residual_samplers.jl:77; util.jl:24)."""
    return lin.vjp(np.sqrt(lin.fisher) * n1) + eta


def krylov_cg(matvec, b, *, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None):
    """Krylov.jl `cg!` as LinearSolve's `KrylovJL_CG()` drives it (call site
    src/residual_samplers.jl:79-80; algorithm in SURVEY.md appendix A.1): x0 = 0, no
    preconditioner, stop on ||r|| <= atol + rtol*||b||, on ||r|| + 1 <= 1, or after itmax
    (default length(b)) iterations.  Returns (x, iterations, final ||r||)."""
    b = np.asarray(b, dtype=np.float64)
    n = b.size
    if itmax is None:
        itmax = n
    x = np.zeros(n)
    r = b.copy()
    p = r.copy()
    gamma = float(r @ r)
    rnorm = math.sqrt(gamma)
    if gamma == 0.0:
        return x, 0, rnorm
    eps = atol + rtol * rnorm
    pnorm2 = gamma
    it = 0
    solved = rnorm <= eps
    tired = it >= itmax
    while not (solved or tired):
        Ap = matvec(p)
        pAp = float(p @ Ap)
        if pAp <= np.finfo(np.float64).eps * pnorm2 and abs(pAp) <= np.finfo(np.float64).eps * pnorm2:
            break                      # zero curvature; unreachable for M >= I
        alpha = gamma / pAp
        x += alpha * p
        r -= alpha * Ap
        gamma_next = float(r @ r)
        rnorm = math.sqrt(gamma_next)
        solved = (rnorm <= eps) or (rnorm + 1.0 <= 1.0)
        if not solved:
            beta = gamma_next / gamma
            pnorm2 = gamma_next + beta * beta * pnorm2
            p = r + beta * p
        gamma = gamma_next
        it += 1
        tired = it >= itmax
    return x, it, rnorm


def sample_residuals_cg(model, center, Z1, Z2, *, atol=SQRT_EPS, rtol=SQRT_EPS, itmax=None):
    """`sample_residuals(s, n)` for the CG sampler (src/residual_samplers.jl:66-92) with the
    draws as inputs: residual i uses n1 = Z1[:, i] (n_lambda) and eta = Z2[:, i] (n_theta).
    Returns (R (n_theta x n), iterations (n), final residual norms (n))."""
    lin = model.linearize(np.asarray(center, dtype=np.float64))
    Z1 = np.asarray(Z1, dtype=np.float64).reshape(lin.n_lambda, -1)
    Z2 = np.asarray(Z2, dtype=np.float64).reshape(lin.n_theta, -1)
    n = Z2.shape[1]
    R = np.empty((lin.n_theta, n))
    iters = np.zeros(n, dtype=np.int64)
    rn = np.zeros(n)
    for i in range(n):
        b = sampler_rhs(lin, Z1[:, i], Z2[:, i])
        R[:, i], iters[i], rn[i] = krylov_cg(lambda v: metric_apply(lin, v), b, atol=atol, rtol=rtol,
                                             itmax=itmax)
    return R, iters, rn


def residual_pushfwd_operator(model, center):
    """src/residual_samplers.jl:95-109: materialise M = J'FJ + I, invert it, return the LOWER
    Cholesky factor of the inverse (PositiveFactorizations `cholesky(Positive, .)` is the
    plain Cholesky on an SPD input)."""
    lin = model.linearize(np.asarray(center, dtype=np.float64))
    J = lin.dense()
    M = J.T @ (lin.fisher[:, None] * J) + np.eye(lin.n_theta)
    Sigma = np.linalg.inv(M)
    Sigma = 0.5 * (Sigma + Sigma.T)
    return np.linalg.cholesky(Sigma)


def sample_residuals_dense(model, center, Z):
    """src/residual_samplers.jl:118-123: R = L_Sigma * Z, Z = randn(n_theta, n)."""
    L = residual_pushfwd_operator(model, center)
    return L @ np.asarray(Z, dtype=np.float64).reshape(L.shape[1], -1), L


# --------------------------------------------------------------------------------------------
# KL estimate -- src/mgvi_impl.jl:3-28
# --------------------------------------------------------------------------------------------
def mv_normal_logdensity(x):
    """src/mgvi_impl.jl:3-7."""
    return -float(x @ x) / 2.0 - float(x.size) * (LOG2PI / 2.0)


def posterior_loglike(model, p):
    """src/mgvi_impl.jl:9-14."""
    return model.loglike(p) + mv_normal_logdensity(p)


def mean_neg_log_pstr(model, R, center):
    """src/mgvi_impl.jl:19-28: mean over the 2n points center +- r_i of the negative
    non-normalised log posterior."""
    res = 0.0
    for i in range(R.shape[1]):
        r = R[:, i]
        res += -(posterior_loglike(model, center + r) + posterior_loglike(model, center - r))
    return res / (2 * R.shape[1])


def kl_gradient(model, R, center):
    """`gradient_func(f, adsel)` of mean_neg_log_pstr (src/newtoncg.jl:100), restated
    analytically: grad = (1/2n) sum_i sum_+- [ p - J(p)' dl/dlambda(p) ], p = center +- r_i."""
    g = np.zeros_like(center)
    for i in range(R.shape[1]):
        for sgn in (1.0, -1.0):
            p = center + sgn * R[:, i]
            g += p - model.grad_loglike(p)
    return g / (2 * R.shape[1])


def mean_metric_apply(model, R, xi, u, lins=None):
    """src/mgvi_impl.jl:137-138: mean over the '+' points xi + r_i (n terms, not 2n) of
    J_i' F_i J_i + I, applied to u."""
    n = R.shape[1]
    if lins is None:
        lins = [model.linearize(xi + R[:, i]) for i in range(n)]
    acc = np.zeros_like(u)
    for lin in lins:
        acc += metric_apply(lin, u)
    return acc / n


# --------------------------------------------------------------------------------------------
# Newton-CG -- src/newtoncg.jl
# --------------------------------------------------------------------------------------------
def itsolvers_cg_iterator(matvec, b, *, reltol=SQRT_EPS, abstol=0.0, maxiter=None, states=None):
    """IterativeSolvers `cg_iterator!(x, A, b; initially_zero=true)` (call site
    src/newtoncg.jl:120; algorithm SURVEY.md appendix A.2) as a Python generator: yields
    (k, x, residual) AFTER the k-th update, exactly when Julia's `enumerate(cgiterator)` runs
    the loop body; `x` is updated in place.  `states` (a list, tests only) receives the iterator
    state (x, r, u, residual, prev_residual) as it stands BEFORE every update."""
    b = np.asarray(b, dtype=np.float64)
    n = b.size
    if maxiter is None:
        maxiter = n
    x = np.zeros(n)
    u = np.zeros(n)
    r = b.copy()
    residual = float(np.linalg.norm(r))
    prev_residual = 1.0
    tol = max(reltol * residual, abstol)
    it = 0
    while True:
        if it >= maxiter or residual <= tol:
            return
        if states is not None:
            states.append((x.copy(), r.copy(), u.copy(), residual, prev_residual))
        beta = residual ** 2 / prev_residual ** 2
        u = r + beta * u
        c = matvec(u)
        alpha = residual ** 2 / float(u @ c)
        x += alpha * u
        r -= alpha * c
        prev_residual = residual
        residual = float(np.linalg.norm(r))
        it += 1
        yield it, x, residual


def _sw_interpolate(a1, a2, f1, f2, d1_, d2_):
    """LineSearches `interpolate` (cubic), SURVEY.md appendix A.3."""
    d1 = d1_ + d2_ - 3.0 * (f1 - f2) / (a1 - a2)
    d2 = math.sqrt(d1 * d1 - d1_ * d2_)
    return a2 - (a2 - a1) * ((d2_ + d2 - d1) / (d2_ - d1_ + 2.0 * d2))


def _sw_zoom(a_lo, a_hi, dphi0, phi0, phi, dphi, c1, c2):
    a_j = float("nan")
    for _ in range(10):
        f_lo, g_lo = phi(a_lo), dphi(a_lo)
        f_hi, g_hi = phi(a_hi), dphi(a_hi)
        if a_lo < a_hi:
            a_j = _sw_interpolate(a_lo, a_hi, f_lo, f_hi, g_lo, g_hi)
        else:
            a_j = _sw_interpolate(a_hi, a_lo, f_hi, f_lo, g_hi, g_lo)
        f_j = phi(a_j)
        if f_j > phi0 + c1 * a_j * dphi0 or f_j > f_lo:
            a_hi = a_j
        else:
            g_j = dphi(a_j)
            if abs(g_j) <= -c2 * dphi0:
                return a_j
            if g_j * (a_hi - a_lo) >= 0.0:
                a_hi = a_lo
            a_lo = a_j
    return a_j


def strong_wolfe(phi, dphi, alpha0, phi0, dphi0, *, c1=1e-4, c2=0.9, rho=2.0, a_max=65536.0):
    """LineSearches `StrongWolfe{Float64}()` (src/newtoncg.jl:30, call :142 with the arguments
    of `linesearch_args` :64-71).  Nocedal & Wright Alg. 3.5/3.6 as restated in SURVEY.md
    appendix A.3.  Returns (a*, phi(a*))."""
    a_prev, a, f_prev, i = 0.0, float(alpha0), phi0, 1
    while a < a_max:
        f_a = phi(a)
        if f_a > phi0 + c1 * a * dphi0 or (f_a >= f_prev and i > 1):
            a_star = _sw_zoom(a_prev, a, dphi0, phi0, phi, dphi, c1, c2)
            return a_star, phi(a_star)
        g_a = dphi(a)
        if abs(g_a) <= -c2 * dphi0:
            return a, f_a
        if g_a >= 0.0:
            a_star = _sw_zoom(a, a_prev, dphi0, phi0, phi, dphi, c1, c2)
            return a_star, phi(a_star)
        a_prev, a, f_prev, i = a, a * rho, f_a, i + 1
    return a_max, phi(a_max)


@dataclass
class NewtonCGOpts:
    """src/newtoncg.jl:14-31."""
    alpha: float = 0.1
    steps: int = 4
    i0: int = 5
    i1: int = 50


@dataclass
class NewtonCGTrace:
    """`NewtonCGResults` (src/newtoncg.jl:34-46) plus the discrete branch trace SURVEY.md
    section 7 asks to compare: CG updates actually done and accepted step per Newton step."""
    f_history: list = field(default_factory=list)
    cg_iterations: list = field(default_factory=list)
    cg_updates: list = field(default_factory=list)
    step_lengths: list = field(default_factory=list)
    f_calls: int = 0
    g_calls: int = 0


def _newton_step(model, R, x, step, f_prev, df_n, opts, f, gradf, tr, rec, force_k=0):
    """The body of the Newton loop, src/newtoncg.jl:112-146.  Returns (x_new, f_n, df_n).  `rec` (a dict
    or None) receives every intermediate; `force_k` > 0 overrides the exit rules of the inner CG
    (teacher forcing, tests only)."""
    n = R.shape[1]
    g = gradf(x)                                 # :115
    lins = [model.linearize(x + R[:, i]) for i in range(n)]     # Sigma_bar_inv(x_n), :120
    dx = np.zeros_like(x)
    updates = 0
    it = itsolvers_cg_iterator(lambda u: mean_metric_apply(model, R, x, u, lins), g,
                               states=rec["cg_states"] if rec is not None else None)
    if step == 1:
        for k, xk, _res in it:                   # :121-127
            dx, updates = xk, k
            if rec is not None:
                rec["dx_iters"].append(xk.copy()); rec["cg_resid"].append(_res); rec["f_k"].append(float("nan"))
            if (k >= force_k) if force_k else (k >= opts.i0):
                break
    else:
        f_km1 = f_prev                           # :129
        for k, xk, _res in it:                   # :132-139
            dx, updates = xk, k
            f_k = f(x - dx)
            if rec is not None:
                rec["dx_iters"].append(xk.copy()); rec["cg_resid"].append(_res); rec["f_k"].append(f_k)
                rec["exit_margin"].append(abs(f_k - f_km1) - opts.alpha * df_n)
            if (k >= force_k) if force_k else (k >= opts.i1 or abs(f_k - f_km1) < opts.alpha * df_n):
                tr.cg_iterations.append(k)
                break
            f_km1 = f_k
    tr.cg_updates.append(updates)
    d = -dx                                      # linesearch_args(..., -dx, ...) :142

    def phi(a):                                  # :66
        v = f(x + a * d)
        if rec is not None:
            rec["ls"].append((0, float(a), v))
        return v

    def dphi(a):                                 # :67
        v = float(gradf(x + a * d) @ d)
        if rec is not None:
            rec["ls"].append((1, float(a), v))
        return v

    dphi0 = float(g @ d)
    beta, f_n = strong_wolfe(phi, dphi, 1.0, f_prev, dphi0)
    tr.step_lengths.append(beta)
    x_new = x - beta * dx                        # :143
    if rec is not None:
        rec.update(grad=g, k=updates, dphi0=dphi0, beta=beta, f_new=f_n, x_new=x_new.copy(), dx=dx.copy())
    return x_new, f_n, abs(f_n - f_prev)         # :144


def _new_record(step, x, f_prev, df_n):
    return dict(step=step, x_n=x.copy(), f_prev=f_prev, df_n=df_n, dx_iters=[], f_k=[], cg_resid=[],
                exit_margin=[], ls=[], cg_states=[])


def newtoncg_optimize(model, R, x0, opts: NewtonCGOpts = NewtonCGOpts(), record: list | None = None):
    """`_optimize(f, adsel, Sigma_bar_inv, x0, ::NewtonCG, opts)` (src/newtoncg.jl:85-167) with
    f = mean_neg_log_pstr(model, R, .) and Sigma_bar_inv = mean metric (src/mgvi_impl.jl:135-139).
    Returns (x, f(x), trace).  `record` (a list) receives one dict per Newton step with every
    intermediate of that step -- the state at the top of the step (x_n, f_prev, df_n), grad f(x_n),
    the inner-CG iterates dx_k with residual norms and f(x_n - dx_k), the exit margin of the
    stagnation test (:134), the line-search evaluations in call order, beta, x_{n+1} -- which is
    what the teacher-forced parity tests compare the device path with (tests/helpers.py)."""
    tr = NewtonCGTrace()

    def f(x):
        tr.f_calls += 1
        return mean_neg_log_pstr(model, R, x)

    def gradf(x):
        tr.g_calls += 1
        return kl_gradient(model, R, x)

    x0 = np.asarray(x0, dtype=np.float64)
    f_prev = f(x0)                                   # :106
    tr.f_history.append(f_prev)
    tr.cg_iterations.append(opts.i0)                 # :108
    f_n = 0.0
    df_n = 0.0
    x = x0.copy()
    for step in range(1, opts.steps + 1):            # :112
        rec = _new_record(step, x, f_prev, df_n) if record is not None else None
        x, f_n, df_n = _newton_step(model, R, x, step, f_prev, df_n, opts, f, gradf, tr, rec)
        f_prev = f_n
        tr.f_history.append(f_n)                     # :146
        if rec is not None:
            record.append(rec)
    return x, f_n, tr


def newtoncg_step_record(model, R, x_n, step, f_prev, df_n, opts: NewtonCGOpts = NewtonCGOpts(), force_k=0):
    """ONE Newton step (src/newtoncg.jl:112-146) from an injected state, fully recorded -- the oracle
    twin of mgvib200_newton_probe (tests only: teacher-forced parity and the oracle's own sensitivity)."""
    tr = NewtonCGTrace()
    x_n = np.asarray(x_n, dtype=np.float64)
    rec = _new_record(step, x_n, f_prev, df_n)
    _newton_step(model, R, x_n, step, f_prev, df_n, opts, lambda x: mean_neg_log_pstr(model, R, x),
                 lambda x: kl_gradient(model, R, x), tr, rec, force_k=force_k)
    return rec


# --------------------------------------------------------------------------------------------
# Driver -- src/mgvi_impl.jl:129-174
# --------------------------------------------------------------------------------------------
def build_samples(R, center):
    """src/mgvi_impl.jl:152-156: columns [c + r1, c - r1, c + r2, c - r2, ...]."""
    n_theta, n = R.shape
    out = np.empty((n_theta, 2 * n))
    out[:, 0::2] = center[:, None] + R
    out[:, 1::2] = center[:, None] - R
    return out


def mgvi_sample(model, center, Z1, Z2, **cg_kw):
    """src/mgvi_impl.jl:166-174."""
    R, iters, rn = sample_residuals_cg(model, center, Z1, Z2, **cg_kw)
    return build_samples(R, np.asarray(center, dtype=np.float64)), iters, rn


def mgvi_step(model, center, *, Z1=None, Z2=None, Z=None, dense=False,
              opts: NewtonCGOpts = NewtonCGOpts(), **cg_kw):
    """src/mgvi_impl.jl:129-144.  CG sampler when `dense` is False (needs Z1, Z2), the
    MatrixInversion sampler otherwise (needs Z).  Returns a dict with samples, mnlp, the
    updated center, the residuals and the traces."""
    center = np.asarray(center, dtype=np.float64)
    if dense:
        R, _L = sample_residuals_dense(model, center, Z)
        iters = rn = None
    else:
        R, iters, rn = sample_residuals_cg(model, center, Z1, Z2, **cg_kw)
    x, fmin, tr = newtoncg_optimize(model, R, center, opts)
    return dict(samples=build_samples(R, x), mnlp=fmin, center=x, residuals=R, sampler_iters=iters,
                sampler_rnorm=rn, trace=tr)
