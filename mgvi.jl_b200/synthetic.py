"""Synthetic inputs for the configurations BASELINE.json names.

Plain NumPy on the host: seeds, grids, amplitude spectra, data and the standard-normal draws
that are handed to BOTH the CUDA library and (in tests) the oracle.  Nothing here computes
any part of the MGVI path itself; the Hartley transform below is only used to synthesise
data from a "true" latent field.

All field arrays are in C order (dims[0] slowest).
"""
from __future__ import annotations

import numpy as np

LAYOUT_MU_ONLY, LAYOUT_BLOCKED, LAYOUT_INTERLEAVED = 0, 1, 2
LINK_IDENTITY, LINK_EXP = 0, 1
LIKE_NORMAL, LIKE_POISSON = 0, 1

# test/test_models/model_polyfit.jl:26-39
POLYFIT_TRUE_PARAMS = np.array([-0.3, -1.5, 0.2, -0.5, 0.3])
POLYFIT_STARTING_POINT = np.array([0.2, 0.5, -0.1, 0.3, -0.6])
POLYFIT_COEF_SCALE = np.array([10.0, 40.0, 600.0, 80.0])   # model_polyfit.jl:15-17
POLYFIT_NOISE_SCALE = 60.0                                  # model_polyfit.jl:20-21


def _hartley(x: np.ndarray) -> np.ndarray:
    for ax in range(x.ndim):
        X = np.fft.fft(x, axis=ax)
        x = X.real - X.imag
    return x


def harmonic_distance(dims) -> np.ndarray:
    """|k| with the mirror symmetry of docs/src/advanced_tutorial_lit.jl:157-172
    (`map_idx`: i <= n/2 ? i : i - n) and test/test_models/model_fft_gp.jl:18."""
    axes = []
    for n in dims:
        i = np.arange(n, dtype=np.float64)
        axes.append(np.where(i <= n // 2, i, i - n))
    grids = np.meshgrid(*axes, indexing="ij")
    return np.sqrt(sum(g * g for g in grids))


def power_law_amplitude(dims, slope=2.3, field_std=1.0) -> tuple[np.ndarray, float]:
    """A = 1/sqrt(k^slope + 1) rescaled so that the field c*H(A.xi) has standard deviation
    `field_std` for xi ~ N(0, I) with c = 1/sqrt(N)."""
    k = harmonic_distance(dims)
    A = 1.0 / np.sqrt(k ** slope + 1.0)
    A *= field_std / np.sqrt(np.mean(A * A))
    N = int(np.prod(dims))
    return A.reshape(N), 1.0 / np.sqrt(N)


def field_config(dims, likelihood, n_residuals, *, sigma=0.4, offset=None, field_std=1.0,
                 layout=None, seed_true=128, seed_data=7, seed_center=12, seed_draws=42,
                 draws=True, center_std=1.0):
    """Correlated-field configuration (C2: dims (65536,), Poisson, n=16; C3: (2048, 2048),
    Normal sigma 0.4, n=32; C5: (256,)*3, Poisson, n=64)."""
    dims = tuple(int(d) for d in dims)
    N = int(np.prod(dims))
    A, c = power_law_amplitude(dims, field_std=field_std)
    if offset is None:
        offset = 2.0 if likelihood == LIKE_POISSON else 0.0
    xi_true = np.random.default_rng(seed_true).standard_normal(N)
    s = c * _hartley((A * xi_true).reshape(dims)).reshape(N) + offset
    lam = np.exp(s)
    rng_d = np.random.default_rng(seed_data)
    if likelihood == LIKE_POISSON:
        data = rng_d.poisson(lam).astype(np.float64)
        layout = LAYOUT_MU_ONLY
    else:
        data = lam + sigma * rng_d.standard_normal(N)
        if layout is None:
            layout = LAYOUT_MU_ONLY
    n_lambda = N if layout == LAYOUT_MU_ONLY else 2 * N
    center = center_std * np.random.default_rng(seed_center).standard_normal(N)
    cfg = dict(kind="field", dims=dims, amplitude=A, scale=c, offset=float(offset), link=LINK_EXP,
               likelihood=likelihood, sigma=float(sigma), layout=layout, data=data, center=center,
               n_residuals=int(n_residuals), n_theta=N, n_lambda=n_lambda, xi_true=xi_true)
    if draws:
        rng = np.random.default_rng(seed_draws)
        # column-major (n_lambda x n) / (n_theta x n): residual i owns a contiguous column
        cfg["Z1"] = rng.standard_normal((n_residuals, n_lambda)).T
        cfg["Z2"] = rng.standard_normal((n_residuals, N)).T
    return cfg


def polyfit_config(n_residuals=8, *, reference_grid=False, layout=LAYOUT_BLOCKED, seed_data=145,
                   seed_draws=42):
    """C1: the model of test/test_models/model_polyfit.jl:11-24 on 1000
    points, or on the reference's own 40-point grids when `reference_grid`."""
    if reference_grid:
        x1 = np.array([i / 10.0 for i in range(1, 26)])            # model_polyfit.jl:11
        x2 = np.array([i / 10.0 + 0.1 for i in range(1, 16)])      # model_polyfit.jl:12
    else:
        x1 = np.array([i / 250.0 for i in range(1, 626)])
        x2 = np.array([i / 250.0 + 0.1 for i in range(1, 376)])
    x = np.concatenate([x1, x2])
    p = POLYFIT_TRUE_PARAMS
    mu = sum(p[k] * POLYFIT_COEF_SCALE[k] * x ** k for k in range(4))
    sig = POLYFIT_NOISE_SCALE * p[4] ** 2
    data = mu + sig * np.random.default_rng(seed_data).standard_normal(x.size)
    rng = np.random.default_rng(seed_draws)
    n_lambda = 2 * x.size
    return dict(kind="poly", x_groups=[x1, x2], coef_scale=POLYFIT_COEF_SCALE.copy(),
                noise_scale=POLYFIT_NOISE_SCALE, layout=layout, data=data,
                center=POLYFIT_STARTING_POINT.copy(), true_params=p.copy(), n_residuals=int(n_residuals),
                n_theta=5, n_lambda=n_lambda,
                Z1=rng.standard_normal((n_residuals, n_lambda)).T,
                Z2=rng.standard_normal((n_residuals, 5)).T,
                Z=rng.standard_normal((n_residuals, 5)).T)


def dense_linear_config(m=16384, k=4096, n_residuals=32, *, sigma=0.5, seed_J=3, seed_draws=42,
                        seed_true=128, seed_data=7, seed_center=12):
    """C4: J_mu = randn(m, k)/sqrt(m), Normal sigma 0.5."""
    J = np.random.default_rng(seed_J).standard_normal((m, k)) / np.sqrt(m)
    xi_true = np.random.default_rng(seed_true).standard_normal(k)
    data = J @ xi_true + sigma * np.random.default_rng(seed_data).standard_normal(m)
    center = np.random.default_rng(seed_center).standard_normal(k)
    rng = np.random.default_rng(seed_draws)
    return dict(kind="dense", J=J, sigma=float(sigma), layout=LAYOUT_MU_ONLY, data=data, center=center,
                n_residuals=int(n_residuals), n_theta=k, n_lambda=m,
                Z=rng.standard_normal((n_residuals, k)).T,
                Z1=rng.standard_normal((n_residuals, m)).T,
                Z2=rng.standard_normal((n_residuals, k)).T)


def tutorial_grid(n_data, xlim=(1851.0, 1962.0), grain=3, padding=80.0):
    """`produce_bins` of docs/src/advanced_tutorial_lit.jl:84-103 in exact rational arithmetic: the
    fine GP grid (bin centres), its bin width and the 0-based start index of every data bin."""
    from fractions import Fraction as Fr
    lo, hi = Fr(xlim[0]), Fr(xlim[1])
    data_binsize = (hi - lo) / n_data
    gp_binsize = data_binsize / grain
    gp_dim = int(((hi - lo) + 2 * Fr(padding)) // gp_binsize)
    left = right = (gp_dim - n_data) // 2
    if (2 * left + n_data * grain) % 2 == 1:
        left += 1
    gp_left_xlim = lo - left * gp_binsize
    gp_right_xlim = hi + right * gp_binsize

    def rng(a, step, b):             # collect(a:step:b)
        out, v = [], a
        while v <= b:
            out.append(v)
            v += step
        return out
    xs = rng(gp_left_xlim + gp_binsize / 2, gp_binsize, lo) + rng(lo + gp_binsize / 2, gp_binsize, hi) + \
        rng(hi + gp_binsize / 2, gp_binsize, gp_right_xlim)
    n_left = len(rng(gp_left_xlim + gp_binsize / 2, gp_binsize, lo))
    idxs = np.arange(n_left, n_left + n_data * grain, grain)      # Julia: left+1 : grain : left + n_data grain (1-based)
    assert n_left == left
    return np.array([float(x) for x in xs]), float(gp_binsize), idxs


def hyper_field_config(counts, n_residuals=3, *, grain=3, padding=80.0, xlim=(1851.0, 1962.0), seed_start=1,
                       seed_draws=42, start_scale=1.0):
    """The coal-mining model of docs/src/advanced_tutorial_lit.jl (hyperparameter field + bin
    aggregation + Poisson counts).  `counts`: disasters per data bin (tests/golden/coal_mining_counts.npz)."""
    counts = np.asarray(counts, dtype=np.float64)
    xs, binsize, idxs = tutorial_grid(counts.size, xlim, grain, padding)
    N = xs.size
    hd = 1.0 / N / binsize                                        # _GP_HARMONIC_DIST :106
    i = np.arange(N, dtype=np.float64)
    kd = np.abs(np.where(i <= N // 2, i, i - N)) * hd             # dist_array :157-172
    n_theta = N + 2
    center = start_scale * np.random.default_rng(seed_start).standard_normal(n_theta)
    rng = np.random.default_rng(seed_draws)
    return dict(kind="hyper", kdist=kd, scale=hd, data_idxs=idxs, grain=int(grain), binsize=binsize, data=counts,
                a0=2.0 * grain, a1=0.9, l0=12.0 / grain ** 0.3, l1=1.0 / 15.0, center=center,
                n_residuals=int(n_residuals), n_theta=n_theta, n_lambda=counts.size, gp_xs=xs,
                Z1=rng.standard_normal((n_residuals, counts.size)).T, Z2=rng.standard_normal((n_residuals, n_theta)).T,
                Z=rng.standard_normal((n_residuals, n_theta)).T)
