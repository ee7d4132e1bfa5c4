"""A SECOND restatement of the MGVI.jl hot path for 1-D correlated-field models, written
independently of oracle/mgvi_oracle.py and evaluated in arbitrary precision (mpmath).  TEST
INFRASTRUCTURE: it exists to bound the Float64 oracle's own rounding (tests/test_oracle_pinning.py)
-- the reference cannot be run in this image, so "the exact answer of the restated algorithm" is
the strongest pin available for rows a5-a12.

Differences from the oracle on purpose (so that agreement means something):
  * the Hartley transform is the O(n^2) cas-kernel sum, not an FFT;
  * the mean metric is the literal average of n per-sample operators J_i' F_i J_i + I
    (src/mgvi_impl.jl:137-138), not the collapsed form with an averaged weight;
  * the gradient of the KL estimate is accumulated per sample point (src/newtoncg.jl:100).

Follows: src/residual_samplers.jl:66-83 (sampler), Krylov.jl cg! as driven by KrylovJL_CG (SURVEY.md
A.1), src/mgvi_impl.jl:3-28 (KL estimate), src/newtoncg.jl:85-167 with IterativeSolvers cg_iterator!
(A.2) and LineSearches StrongWolfe (A.3)."""
import math

import mpmath as mp
import numpy as np


def small_config(n=40, n_res=2, seed=3):
    """40 pixels (the reference fixture's size), power-law amplitude with field std 0.5, Normal(exp(s), 0.4)
    data, interleaved [mu, sigma] layout -- well conditioned, unlike the reference's own constants."""
    rng = np.random.default_rng(seed)
    k = np.array([i if i <= n // 2 else i - n for i in range(n)], dtype=np.float64)
    A = 1.0 / np.sqrt(np.abs(k) ** 2.3 + 1.0)
    A *= 0.5 / np.sqrt(np.mean(A * A))
    c = 1.0 / math.sqrt(n)
    xi_true = rng.standard_normal(n)
    X = np.fft.fft(A * xi_true)
    lam = np.exp(c * (X.real - X.imag))
    sigma = 0.4
    data = lam + sigma * rng.standard_normal(n)
    center = 0.5 * rng.standard_normal(n)
    return dict(kind="field", dims=(n,), amplitude=A, scale=c, offset=0.0, link=1, likelihood=0, sigma=sigma,
                layout=2, data=data, center=center, n_residuals=n_res, n_theta=n, n_lambda=2 * n,
                Z1=rng.standard_normal((n_res, 2 * n)).T, Z2=rng.standard_normal((n_res, n)).T,
                V=rng.standard_normal(n))


class _Model:
    def __init__(self, cfg):
        self.n = int(cfg["n_theta"])
        n = self.n
        self.A = [mp.mpf(float(v)) for v in cfg["amplitude"]]
        self.c = mp.mpf(float(cfg["scale"]))
        self.offset = mp.mpf(float(cfg["offset"]))
        self.link = int(cfg["link"])
        self.like = int(cfg["likelihood"])
        self.sigma = mp.mpf(float(cfg["sigma"]))
        self.data = [mp.mpf(float(v)) for v in cfg["data"]]
        self.layout = int(cfg["layout"])
        two_pi = 2 * mp.pi
        self.cas = [[mp.cos(two_pi * ((j * k) % n) / n) + mp.sin(two_pi * ((j * k) % n) / n) for j in range(n)]
                    for k in range(n)]
        if self.layout == 0:
            self.mu_rows = list(range(n))
        elif self.layout == 1:
            self.mu_rows = list(range(n))
        else:
            self.mu_rows = [2 * i for i in range(n)]

    def H(self, x):
        return [mp.fsum(row[j] * x[j] for j in range(self.n)) for row in self.cas]

    def signal(self, xi):
        h = self.H([a * x for a, x in zip(self.A, xi)])
        return [self.c * v + self.offset for v in h]

    def rate(self, xi):
        s = self.signal(xi)
        return [mp.e ** v for v in s] if self.link == 1 else s

    def fisher_gp(self, xi):
        lam = self.rate(xi)
        gp = lam if self.link == 1 else [mp.mpf(1)] * self.n
        F = [1 / l for l in lam] if self.like == 1 else [1 / self.sigma ** 2] * self.n
        return lam, gp, F

    def jvp(self, gp, v):
        h = self.H([a * x for a, x in zip(self.A, v)])
        return [g * (self.c * t) for g, t in zip(gp, h)]

    def vjp(self, gp, l):
        h = self.H([g * t for g, t in zip(gp, l)])
        return [self.c * (a * t) for a, t in zip(self.A, h)]

    def metric(self, gp, F, v):
        jv = self.jvp(gp, v)
        out = self.vjp(gp, [f * t for f, t in zip(F, jv)])
        return [o + x for o, x in zip(out, v)]

    def loglike(self, xi):
        lam = self.rate(xi)
        if self.like == 1:
            s = self.signal(xi) if self.link == 1 else [mp.log(l) for l in lam]
            return mp.fsum(d * t - l - mp.loggamma(d + 1) for d, t, l in zip(self.data, s, lam))
        return mp.fsum(-((d - l) / self.sigma) ** 2 / 2 for d, l in zip(self.data, lam)) \
            - self.n * (mp.log(self.sigma) + mp.log(2 * mp.pi) / 2)

    def grad_loglike(self, xi):
        lam, gp, _ = self.fisher_gp(xi)
        if self.like == 1:
            dl = [d / l - 1 for d, l in zip(self.data, lam)]
        else:
            dl = [(d - l) / self.sigma ** 2 for d, l in zip(self.data, lam)]
        return self.vjp(gp, dl)


def _dot(a, b):
    return mp.fsum(x * y for x, y in zip(a, b))


def _axpy(a, x, y):
    return [a * xi + yi for xi, yi in zip(x, y)]


def _krylov_cg(matvec, b, atol, rtol, itmax):
    n = len(b)
    x = [mp.mpf(0)] * n
    r = list(b)
    p = list(b)
    gamma = _dot(r, r)
    rn = mp.sqrt(gamma)
    if gamma == 0:
        return x, 0
    eps = atol + rtol * rn
    it = 0
    solved = rn <= eps
    while not (solved or it >= itmax):
        Ap = matvec(p)
        alpha = gamma / _dot(p, Ap)
        x = _axpy(alpha, p, x)
        r = _axpy(-alpha, Ap, r)
        g2 = _dot(r, r)
        rn = mp.sqrt(g2)
        solved = rn <= eps
        if not solved:
            p = _axpy(g2 / gamma, p, r)
        gamma = g2
        it += 1
    return x, it


def _cg_iterator(matvec, b, reltol, maxiter):
    n = len(b)
    x = [mp.mpf(0)] * n
    u = [mp.mpf(0)] * n
    r = list(b)
    res = mp.sqrt(_dot(r, r))
    prev = mp.mpf(1)
    tol = reltol * res
    it = 0
    while True:
        if it >= maxiter or res <= tol:
            return
        beta = res ** 2 / prev ** 2
        u = _axpy(beta, u, r)
        c = matvec(u)
        alpha = res ** 2 / _dot(u, c)
        x = _axpy(alpha, u, x)
        r = _axpy(-alpha, c, r)
        prev = res
        res = mp.sqrt(_dot(r, r))
        it += 1
        yield it, x


def _interp(a1, a2, f1, f2, g1, g2):
    d1 = g1 + g2 - 3 * (f1 - f2) / (a1 - a2)
    d2 = mp.sqrt(d1 * d1 - g1 * g2)
    return a2 - (a2 - a1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))


def _zoom(a_lo, a_hi, dphi0, phi0, phi, dphi, c1, c2):
    a_j = None
    for _ in range(10):
        f_lo, g_lo, f_hi, g_hi = phi(a_lo), dphi(a_lo), phi(a_hi), dphi(a_hi)
        a_j = _interp(a_lo, a_hi, f_lo, f_hi, g_lo, g_hi) if a_lo < a_hi else _interp(a_hi, a_lo, f_hi, f_lo, g_hi, g_lo)
        f_j = phi(a_j)
        if f_j > phi0 + c1 * a_j * dphi0 or f_j > f_lo:
            a_hi = a_j
        else:
            g_j = dphi(a_j)
            if abs(g_j) <= -c2 * dphi0:
                return a_j
            if g_j * (a_hi - a_lo) >= 0:
                a_hi = a_lo
            a_lo = a_j
    return a_j


def _strong_wolfe(phi, dphi, phi0, dphi0):
    c1, c2, rho, a_max = mp.mpf("1e-4"), mp.mpf("0.9"), mp.mpf(2), mp.mpf(65536)
    a_prev, a, f_prev, i = mp.mpf(0), mp.mpf(1), phi0, 1
    while a < a_max:
        f_a = phi(a)
        if f_a > phi0 + c1 * a * dphi0 or (f_a >= f_prev and i > 1):
            a_s = _zoom(a_prev, a, dphi0, phi0, phi, dphi, c1, c2)
            return a_s, phi(a_s)
        g_a = dphi(a)
        if abs(g_a) <= -c2 * dphi0:
            return a, f_a
        if g_a >= 0:
            a_s = _zoom(a, a_prev, dphi0, phi0, phi, dphi, c1, c2)
            return a_s, phi(a_s)
        a_prev, a, f_prev, i = a, a * rho, f_a, i + 1
    return a_max, phi(a_max)


def _kl(m, R, x):
    n = len(R)
    tot = mp.mpf(0)
    for r in R:
        for sg in (1, -1):
            p = [xi + sg * ri for xi, ri in zip(x, r)]
            tot += -(m.loglike(p) - _dot(p, p) / 2 - m.n * mp.log(2 * mp.pi) / 2)
    return tot / (2 * n)


def _kl_grad(m, R, x):
    n = len(R)
    g = [mp.mpf(0)] * m.n
    for r in R:
        for sg in (1, -1):
            p = [xi + sg * ri for xi, ri in zip(x, r)]
            gl = m.grad_loglike(p)
            g = [gi + (pi - li) for gi, pi, li in zip(g, p, gl)]
    return [gi / (2 * n) for gi in g]


def _mean_metric(m, R, x):
    lins = []
    for r in R:
        _, gp, F = m.fisher_gp([xi + ri for xi, ri in zip(x, r)])
        lins.append((gp, F))

    def apply(u):
        acc = [mp.mpf(0)] * m.n
        for gp, F in lins:
            acc = [a + b for a, b in zip(acc, m.metric(gp, F, u))]
        return [a / len(lins) for a in acc]
    return apply


def _newton(m, R, x0, steps=4, i0=5, i1=50, alpha=mp.mpf("0.1"), sqrt_eps=mp.mpf(2) ** -26):
    x = list(x0)
    f_prev = _kl(m, R, x)
    df_n = mp.mpf(0)
    f_n = mp.mpf(0)
    cg_updates, step_lengths, first = [], [], None
    for step in range(1, steps + 1):
        g = _kl_grad(m, R, x)
        Mbar = _mean_metric(m, R, x)
        dx = [mp.mpf(0)] * m.n
        k_done = 0
        if step == 1:
            for k, xk in _cg_iterator(Mbar, g, sqrt_eps, m.n):
                dx, k_done = xk, k
                if k >= i0:
                    break
        else:
            f_km1 = f_prev
            for k, xk in _cg_iterator(Mbar, g, sqrt_eps, m.n):
                dx, k_done = xk, k
                f_k = _kl(m, R, [a - b for a, b in zip(x, dx)])
                if k >= i1 or abs(f_k - f_km1) < alpha * df_n:
                    break
                f_km1 = f_k
        cg_updates.append(k_done)
        if first is None:
            first = dict(grad=list(g), dx=list(dx))
        d = [-v for v in dx]
        xs = list(x)
        phi = lambda a: _kl(m, R, _axpy(a, d, xs))                      # noqa: E731
        dphi = lambda a: _dot(_kl_grad(m, R, _axpy(a, d, xs)), d)       # noqa: E731
        beta, f_n = _strong_wolfe(phi, dphi, f_prev, _dot(g, d))
        step_lengths.append(beta)
        x = _axpy(-beta, dx, x)
        df_n = abs(f_n - f_prev)
        f_prev = f_n
    return x, f_n, cg_updates, step_lengths, first


def _tof(v):
    return np.array([float(t) for t in v])


def mgvi_step(cfg, digits=50, cg_rtol=1e-12):
    """src/mgvi_impl.jl:129-144 in `digits`-digit arithmetic.  Returns Float64 roundings of the exact
    results: metric product of cfg['V'], residual samples, KL value / gradient at the centre, the updated
    centre, mnlp and the discrete trace."""
    with mp.workdps(digits):
        m = _Model(cfg)
        n, nr = m.n, int(cfg["n_residuals"])
        c = [mp.mpf(float(v)) for v in cfg["center"]]
        _, gp, Fm = m.fisher_gp(c)
        MV = m.metric(gp, Fm, [mp.mpf(float(v)) for v in cfg["V"]])
        R = []
        for i in range(nr):
            n1 = [mp.mpf(float(cfg["Z1"][r, i])) for r in m.mu_rows]
            eta = [mp.mpf(float(v)) for v in cfg["Z2"][:, i]]
            b = m.vjp(gp, [mp.sqrt(f) * z for f, z in zip(Fm, n1)])
            b = [bi + ei for bi, ei in zip(b, eta)]
            x, _ = _krylov_cg(lambda v: m.metric(gp, Fm, v), b, mp.mpf(0), mp.mpf(cg_rtol), n)
            R.append(x)
        f0 = _kl(m, R, c)
        g0 = _kl_grad(m, R, c)
        x, f_n, cg_updates, step_lengths, _ = _newton(m, R, c)
        return dict(MV=_tof(MV), R=np.stack([_tof(r) for r in R], axis=1), f0=float(f0), grad0=_tof(g0),
                    center=_tof(x), mnlp=float(f_n), cg_updates=cg_updates, step_lengths=[float(b) for b in step_lengths])


def newton_first_step(cfg, R, digits=50):
    """grad f(x_0) and dx after the first Newton step's 5 CG updates (src/newtoncg.jl:115-127), exact."""
    with mp.workdps(digits):
        m = _Model(cfg)
        c = [mp.mpf(float(v)) for v in cfg["center"]]
        Rm = [[mp.mpf(float(v)) for v in R[:, i]] for i in range(R.shape[1])]
        _, _, _, _, first = _newton(m, Rm, c, steps=1)
        return dict(grad=_tof(first["grad"]), dx=_tof(first["dx"]))


def mean_metric_condition(cfg, R):
    """2-norm condition number of the mean metric at the centre (Float64, dense)."""
    n = int(cfg["n_theta"])
    jk = np.outer(np.arange(n), np.arange(n)) % n
    Hm = np.cos(2 * np.pi * jk / n) + np.sin(2 * np.pi * jk / n)
    A, c = np.asarray(cfg["amplitude"], float), float(cfg["scale"])
    M = np.zeros((n, n))
    for i in range(R.shape[1]):
        s = c * (Hm @ (A * (cfg["center"] + R[:, i]))) + float(cfg["offset"])
        lam = np.exp(s) if int(cfg["link"]) == 1 else s
        gp = lam if int(cfg["link"]) == 1 else np.ones(n)
        F = 1.0 / lam if int(cfg["likelihood"]) == 1 else np.full(n, 1.0 / float(cfg["sigma"]) ** 2)
        J = (gp[:, None] * (c * Hm)) * A[None, :]
        M += J.T @ (F[:, None] * J) + np.eye(n)
    return float(np.linalg.cond(M / R.shape[1]))
