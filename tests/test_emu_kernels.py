"""CPU checks of the REAL kernel and host sources: the .cu files compiled with g++ against the
fiber SIMT emulator (tests/simt_emu/, test infrastructure) and driven through the same C ABI,
compared with the oracle.  Sizes are tiny (the emulator runs one thread at a time); the GPU
tests (test_gpu_parity.py) repeat this on the product library at real sizes."""
import numpy as np
import pytest

import oracle as O
from golden_io import load_golden
from helpers import RTOL, TIGHT, assert_field_ops, check_field_ops, oracle_model, rel


def _field(pkg, dims, like, n, layout=None, std=0.5):
    syn = pkg.synthetic
    return syn.field_config(dims, getattr(syn, like), n, field_std=std, layout=layout)


# every leading-radix case of the FFT (log2 n mod 3 = 0, 1, 2), single pass 1-D geometry
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512])
def test_metric_1d_all_radix_leads(pkg, emu_ctx, n):
    cfg = _field(pkg, (n,), "LIKE_POISSON", 3)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    lin = om.linearize(cfg["center"])
    V = np.random.default_rng(5).standard_normal((n, 3))       # odd count: one half-empty pair
    MV = s.metric_apply(V)
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(3)], axis=1)
    assert rel(MV, MVo) < RTOL
    w, q = s.weights()
    lam = om.rate(cfg["center"])
    assert rel(w, lam) < RTOL and rel(q, np.sqrt(lam)) < RTOL     # Poisson + exp: w = lambda, q = sqrt(lambda)
    m.close()


@pytest.mark.parametrize("dims,like,layout", [
    ((64,), "LIKE_POISSON", None),
    ((16, 32), "LIKE_NORMAL", 2),       # interleaved [mu, sigma] layout
    ((8, 64), "LIKE_POISSON", None),
    ((32, 8), "LIKE_NORMAL", 1),        # blocked [mu; var] layout
    ((8, 16, 8), "LIKE_POISSON", None),
    ((8, 256), "LIKE_POISSON", None),   # 256-point rows: compile-time-length kernels (contiguous)
    ((256, 8), "LIKE_NORMAL", 0),       # 256-point columns: compile-time-length kernels (strided, fused)
    ((256, 16), "LIKE_POISSON", None),  # ... with the 8-column tile the 256^3 workload uses
    ((40,), "LIKE_NORMAL", 2),          # the reference fixture's length: generic (non power of two) path
    ((6, 10), "LIKE_POISSON", None),
])
def test_field_ops_vs_oracle(pkg, emu_ctx, dims, like, layout):
    cfg = _field(pkg, dims, like, 3 if len(dims) == 1 else 2, layout)
    assert_field_ops(check_field_ops(pkg, emu_ctx, cfg, check_step=True), cfg)


def test_metric_3d_wide_tiles(pkg, emu_ctx):
    """3-D field with a 256-point middle axis: the single strided pass with the 8-column tile (the
    256^3 workload's shape class) plus the fused axis-0 pass; metric product only (an emulated CG
    solve at this size takes minutes)."""
    from helpers import oracle_model, rel
    cfg = _field(pkg, (2, 256, 16), "LIKE_POISSON", 2, None)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    V = np.random.default_rng(11).standard_normal((m.n_theta, 2))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12
    m.close()


def test_identity_link_and_normal(pkg, emu_ctx):
    syn = pkg.synthetic
    cfg = syn.field_config((32,), syn.LIKE_NORMAL, 2, field_std=0.5)
    cfg["link"] = syn.LINK_IDENTITY
    # linear-Gaussian: Newton converges in step 1, later steps divide rounding noise by rounding
    # noise -> judged against the oracle's own reproducibility (helpers.assert_step_parity)
    assert_field_ops(check_field_ops(pkg, emu_ctx, cfg, check_step=True), cfg)


@pytest.mark.parametrize("name", ["fft_gp_40", "field_1d_64_poisson", "field_2d_16x32_normal",
                                  "field_3d_8x8x16_poisson"])
def test_golden_vectors(pkg, emu_ctx, name):
    cfg, gold = load_golden(name)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    assert rel(s.metric_apply(gold["V"]), gold["MV"]) < RTOL
    R = pkg.sample_residuals(s, cfg["n_residuals"], draws=(cfg["Z1"], cfg["Z2"]))
    assert rel(R, gold["R"]) < RTOL
    f, g = pkg.kl_value_grad(m, cfg["center"], gold["R"])
    assert abs(f - gold["f"]) / abs(gold["f"]) < RTOL and rel(g, gold["grad"]) < RTOL
    res, c = pkg.mgvi_step(m, cfg["data"], cfg["n_residuals"], cfg["center"],
                           pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT)), emu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    # residuals (antithetic half-differences) do not pass through the optimiser: tight against the fixture
    assert rel(res.samples[:, 0::2] - res.samples[:, 1::2],
               gold["step_samples"][:, 0::2] - gold["step_samples"][:, 1::2]) < RTOL
    assert abs(res.info["optimizer_output"]["f_history"][0] - gold["tr_f_history"][0]) <= RTOL * abs(gold["tr_f_history"][0])
    # Newton phase: strict end-to-end where the oracle's trajectory is reproduced, otherwise every Newton
    # step teacher-forced from the oracle's x_n (helpers.assert_step_parity)
    from helpers import TIGHT_O, assert_step_parity
    om = oracle_model(cfg)
    ref = O.mgvi_step(om, cfg["center"], Z1=cfg["Z1"], Z2=cfg["Z2"], **TIGHT_O)
    assert_step_parity(pkg, m, om, cfg, res, c, ref, label="golden " + name)
    m.close()


def test_cg_maxiters_and_default_tolerance(pkg, emu_ctx):
    cfg = _field(pkg, (64,), "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    # itmax reached: identical truncated iterates
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(abstol=0.0, reltol=1e-30, maxiters=7), emu_ctx)
    R, info = pkg.sample_residuals(s, 2, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], atol=0.0, rtol=1e-30, itmax=7)
    assert info["iters"].tolist() == [7, 7] == ito.tolist()
    assert rel(R, Ro) < RTOL
    # reference default stopping rule (sqrt(eps) abs + rel): same iteration counts
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    R, info = pkg.sample_residuals(s, 2, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, rno = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"])
    assert info["iters"].tolist() == ito.tolist()
    assert rel(R, Ro) < 1e-6
    np.testing.assert_allclose(info["final_rnorm"], rno, rtol=1e-3)
    m.close()


def test_single_residual_and_rng_draws(pkg, emu_ctx):
    cfg = _field(pkg, (16, 8), "LIKE_POISSON", 1)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    r1 = pkg.sample_residuals(s)                 # n omitted -> one vector (src/residual_samplers.jl:66-83)
    assert r1.shape == (m.n_theta,)
    emu_ctx.rng = np.random.default_rng(9)
    a = pkg.sample_residuals(s, 2)
    emu_ctx.rng = np.random.default_rng(9)
    b = pkg.sample_residuals(s, 2)
    np.testing.assert_array_equal(a, b)
    m.close()


@pytest.mark.parametrize("layout", [1, 2])
def test_polyfit_reference_model(pkg, emu_ctx, layout):
    """BASELINE.json configs[0] on the reference's own 40-point grids (model_polyfit.jl:11-12)."""
    from helpers import check_poly_ops
    cfg = pkg.synthetic.polyfit_config(4, reference_grid=True, layout=layout)
    out = check_poly_ops(pkg, emu_ctx, cfg)
    for k in ("metric", "fisher", "dense", "kl", "grad", "mean_metric"):
        assert out[k] < 1e-9, (k, out)      # the 5x5 metric has condition number ~1e7
    assert out["iters_default"][0] == out["iters_default"][1] == [5] * 4
    assert out["residuals"] < 1e-8, out
    assert out["cg_step"] < 1e-8 and out["dense_step"] < 1e-9, out


@pytest.mark.parametrize("m,k,n", [(96, 40, 3), (128, 100, 2)])
def test_dense_linear_model(pkg, emu_ctx, m, k, n):
    """FullJacobianFunc + FullResidualSampler shape (dense J, MatrixInversion) at sizes the
    emulator handles; tile edges are exercised (96 x 40 does not fill a 128 x 128 x 16 tile; k = 100
    is one full 64-column Cholesky block, its panel and trailing update, and a short last block)."""
    from helpers import check_dense_ops
    cfg = pkg.synthetic.dense_linear_config(m, k, n)
    out = check_dense_ops(pkg, emu_ctx, cfg)
    for k in ("metric", "residuals", "dense", "L", "kl", "grad", "mean_metric", "dense_step_residuals"):
        assert out[k] < 1e-10, (k, out)
    assert out["L_lower"] == 0.0
    assert out["iters"][0] == out["iters"][1]


@pytest.mark.parametrize("N,like,layout", [(64, "LIKE_POISSON", None), (128, "LIKE_NORMAL", 2), (1024, "LIKE_POISSON", None)])
def test_long_1d_four_step_path(pkg, emu_ctx, monkeypatch, N, like, layout):
    """The four-step split used for 1-D fields longer than 4096 pixels (BASELINE.json configs[1],
    2^16 pixels), forced on at small sizes: n = n1 * n2 with both square and non-square splits."""
    monkeypatch.setenv("MGVIB200_BIG_MIN", "64")
    cfg = _field(pkg, (N,), like, 3, layout)
    assert_field_ops(check_field_ops(pkg, emu_ctx, cfg, check_step=True), cfg)


def test_zero_draws_and_single_iteration(pkg, emu_ctx):
    """gamma = 0 leaves x = 0 without iterating (Krylov.jl: 'x = 0 is a zero-residual solution');
    maxiters = 1 stops after exactly one update."""
    cfg = _field(pkg, (16, 16), "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    R, info = pkg.sample_residuals(s, 2, draws=(np.zeros_like(cfg["Z1"]), np.zeros_like(cfg["Z2"])), return_info=True)
    assert np.all(R == 0.0) and info["iters"].tolist() == [0, 0]
    # one zero right-hand side next to a live one: masks are per vector
    Z1, Z2 = cfg["Z1"].copy(), cfg["Z2"].copy()
    Z1[:, 0] = 0.0
    Z2[:, 0] = 0.0
    om = oracle_model(cfg)
    R, info = pkg.sample_residuals(s, 2, draws=(Z1, Z2), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], Z1, Z2)
    assert info["iters"][0] == 0 and np.all(R[:, 0] == 0.0)
    assert abs(int(info["iters"][1]) - int(ito[1])) <= 2 and rel(R[:, 1], Ro[:, 1]) < 1e-6
    s1 = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(maxiters=1), emu_ctx)
    R1, info1 = pkg.sample_residuals(s1, 2, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro1, ito1, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], itmax=1)
    assert info1["iters"].tolist() == [1, 1] == ito1.tolist() and rel(R1, Ro1) < RTOL
    m.close()


def test_poisson_identity_link(pkg, emu_ctx):
    """Poisson rate with the identity link (F = 1/lambda, g' = 1): lambda = c H(A xi) + offset > 0."""
    syn = pkg.synthetic
    cfg = syn.field_config((32,), syn.LIKE_POISSON, 2, field_std=0.3, offset=6.0)
    cfg["link"] = syn.LINK_IDENTITY
    rate = oracle_model(cfg).rate(cfg["xi_true"])
    assert rate.min() > 0
    cfg["data"] = np.random.default_rng(7).poisson(rate).astype(float)
    cfg["center"] = 0.2 * cfg["center"]
    assert oracle_model(cfg).rate(cfg["center"]).min() > 0
    out = check_field_ops(pkg, emu_ctx, cfg, check_step=False)
    assert_field_ops(out)


def test_mvnormal_pushfwd_function(pkg, emu_ctx):
    """`mgvi_mvnormal_pushfwd_function` (src/mgvi_impl.jl:191-198): z -> L_Sigma z + center."""
    cfg = pkg.synthetic.polyfit_config(2, reference_grid=True)
    m = pkg.model_from_config(cfg, emu_ctx)
    c = cfg["true_params"]
    f = pkg.mgvi_mvnormal_pushfwd_function(m, cfg["data"], pkg.MGVIConfig(), c, emu_ctx)
    L = O.residual_pushfwd_operator(oracle_model(cfg), c)
    z = np.random.default_rng(3).standard_normal(5)
    assert rel(f(z), L @ z + c) < 1e-9
    Z = np.random.default_rng(4).standard_normal((5, 3))
    assert rel(f(Z), L @ Z + c[:, None]) < 1e-9
    m.close()


@pytest.mark.parametrize("dims", [(1,), (2,), (3,), (1, 8), (8, 1), (1, 1, 8), (1, 5), (2, 2)])
def test_degenerate_shapes(pkg, emu_ctx, dims):
    """Singleton axes, one-pixel and two-pixel fields: metric product and CG residuals vs the oracle."""
    from helpers import TIGHT_O
    cfg = _field(pkg, dims, "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    V = np.random.default_rng(1).standard_normal((m.n_theta, 2))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12
    R = pkg.sample_residuals(s, 2, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert rel(R, Ro) < RTOL
    m.close()


def test_empty_batches(pkg, emu_ctx):
    """Zero residual samples / zero probe vectors: n_theta x 0 results, as the reference's
    `sample_residuals(s, 0)` gives; the averaged quantities refuse an empty sample set."""
    cfg = _field(pkg, (16, 8), "LIKE_NORMAL", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    assert pkg.sample_residuals(s, 0).shape == (m.n_theta, 0)
    assert s.metric_apply(np.zeros((m.n_theta, 0))).shape == (m.n_theta, 0)
    assert pkg.build_samples(m, np.zeros((m.n_theta, 0)), cfg["center"]).shape == (m.n_theta, 0)
    with pytest.raises(pkg.MGVIB200Error):
        pkg.kl_value_grad(m, cfg["center"], np.zeros((m.n_theta, 0)))
    m.close()


@pytest.mark.parametrize("dims,L", [((32, 64), 2), ((32, 128), 4), ((256, 32), 2), ((16, 256), 2)])
def test_tile_major_scratch_forced(pkg, emu_ctx, monkeypatch, dims, L):
    """Tile-major scratch/weight layout of the 2-D CG chain (what 2048-point columns use), forced on
    at small sizes: every field operation incl. the full step against the oracle."""
    monkeypatch.setenv("MGVIB200_TILED", "1")
    monkeypatch.setenv("MGVIB200_STRIDED_L", str(L))
    cfg = _field(pkg, dims, "LIKE_POISSON", 3)
    assert_field_ops(check_field_ops(pkg, emu_ctx, cfg, check_step=(dims == (32, 64))), cfg)


def test_tile_major_scratch_2048_columns(pkg, emu_ctx):
    """The benchmark's column length with its default tile (2 complex columns -> tile-major scratch,
    compile-time-shape kernels): metric product vs the oracle."""
    cfg = _field(pkg, (2048, 32), "LIKE_NORMAL", 2, 0)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    V = np.random.default_rng(12).standard_normal((m.n_theta, 3))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(3)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12
    m.close()


def _random_shapes(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        d = int(rng.integers(1, 4))
        dims = tuple(int(rng.choice([1, 3, 5, 8, 8, 12, 16, 16, 32, 32, 64, 128])) for _ in range(d))
        if 2 <= int(np.prod(dims)) <= 16384:
            out.append(dims)
    return out


@pytest.mark.parametrize("dims,like", [((40,), "LIKE_NORMAL"), ((100,), "LIKE_POISSON"), ((33,), "LIKE_POISSON"),
                                       ((129,), "LIKE_NORMAL"), ((257,), "LIKE_POISSON"), ((1000,), "LIKE_POISSON"),
                                       ((12, 40), "LIKE_POISSON"), ((36, 8), "LIKE_NORMAL"), ((5, 50), "LIKE_POISSON"),
                                       ((2, 35, 34), "LIKE_NORMAL"), ((64, 100), "LIKE_POISSON")])
def test_any_length_axes_bluestein(pkg, emu_ctx, dims, like):
    """Axes that are not a power of two and longer than 32 points take the O(m log m) Bluestein pass
    (field_pass.h: bluestein_axis_body) -- odd and even lengths, lengths next to a power of two, an odd number
    of lines (the unpaired last line), mixed with short direct-pass axes and with power-of-two axes."""
    from helpers import TIGHT_O
    cfg = _field(pkg, dims, like, 3)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    V = np.random.default_rng(4).standard_normal((m.n_theta, 3))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(3)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12, dims
    if m.n_theta > 1100:        # the emulator runs one fiber per thread: CG solves only on the small shapes
        m.close()
        return
    R = pkg.sample_residuals(s, 3, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert rel(R, Ro) < RTOL, dims
    f, g = pkg.kl_value_grad(m, cfg["center"], Ro)
    assert abs(f - O.mean_neg_log_pstr(om, Ro, cfg["center"])) < 1e-12 * abs(f)
    assert rel(g, O.kl_gradient(om, Ro, cfg["center"])) < RTOL
    m.close()


@pytest.mark.parametrize("dims", [(16, 32), (8, 16, 8), (256, 16)])
def test_deferred_solution_update_is_bitwise_the_plain_one(pkg, emu_ctx, monkeypatch, dims):
    """Sampler CG of 2-D / 3-D fields: x += alpha p rides in the fused axis-0 pass of the next iteration
    (field_pass.h: defer_x, p double-buffered, per-vector pending flags, final flush).  Same arithmetic in
    the same order: results must equal those of the streaming update kernel bit for bit -- with residuals
    that converge at very different iterations, an iteration cap that stops mid-flight, and a single iteration."""
    cfg = _field(pkg, dims, "LIKE_POISSON", 5)
    scale = np.logspace(-7, 2, 5)
    Z1, Z2 = cfg["Z1"] * scale, cfg["Z2"] * scale
    om = oracle_model(cfg)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MGVIB200_DEFER_X", mode)
        m = pkg.model_from_config(cfg, emu_ctx)
        res = []
        for opts in (dict(abstol=1e-9, reltol=1e-12), dict(abstol=0.0, reltol=0.0, maxiters=1),
                     dict(abstol=0.0, reltol=0.0, maxiters=6), dict(abstol=1e-3, reltol=0.0, maxiters=9)):
            s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**opts), emu_ctx)
            R, info = pkg.sample_residuals(s, 5, draws=(Z1, Z2), return_info=True)
            res.append((R.copy(), info["iters"].copy()))
        out[mode] = res
        m.close()
    for (Ra, ia), (Rb, ib) in zip(out["0"], out["1"]):
        assert np.array_equal(ia, ib) and np.array_equal(Ra, Rb)
    assert len(set(out["1"][0][1].tolist())) >= 3            # the residuals really stop at different iterations
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], Z1, Z2, atol=1e-9, rtol=1e-12)
    for i in range(5):
        assert np.linalg.norm(out["1"][0][0][:, i] - Ro[:, i]) <= 1e-8 * np.linalg.norm(Ro[:, i]) + 2e-9


@pytest.mark.parametrize("dims", [(8192, 8), (8, 8192)])
def test_longest_power_of_two_axes(pkg, emu_ctx, dims):
    """8192-point axes (2-D / 3-D limit): 1024 threads per line, 128 KB of shared memory per block."""
    cfg = _field(pkg, dims, "LIKE_POISSON", 1)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(abstol=0.0, reltol=0.0, maxiters=3), emu_ctx)
    V = np.random.default_rng(1).standard_normal((m.n_theta, 1))
    lin = om.linearize(cfg["center"])
    assert rel(s.metric_apply(V)[:, 0], O.metric_apply(lin, V[:, 0])) < 1e-12
    R = pkg.sample_residuals(s, 1, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], atol=0.0, rtol=0.0, itmax=3)
    assert rel(R, Ro) < RTOL
    m.close()


@pytest.mark.parametrize("dims", [(5000,), (4097, 8), (8, 8200), (16384, 16)])
def test_unsupported_axis_lengths_fail_loudly(pkg, emu_ctx, dims):
    """Beyond the documented limits (include/mgvi_b200.h: any length <= 4096, powers of two <= 8192 in 2-D / 3-D,
    1-D powers of two <= 2^24) model creation raises -- no silent slow path, no CPU fallback."""
    cfg = _field(pkg, dims, "LIKE_POISSON", 1)
    with pytest.raises(pkg.MGVIB200Error, match="axis length"):
        pkg.model_from_config(cfg, emu_ctx)


def test_bluestein_converged_partner_is_masked(pkg, emu_ctx):
    """Two residual samples of a 1-D any-length field share one complex transform.  Once one of them has
    converged its line must leave the transform (a stale line is transformed over and over, grows by ~1e6 per
    iteration and swamps the rounding of its live partner -- found on the GPU as NaN residuals)."""
    cfg = _field(pkg, (100,), "LIKE_POISSON", 2)
    Z1, Z2 = cfg["Z1"].copy(), cfg["Z2"].copy()
    Z1[:, 1] *= 1e-9
    Z2[:, 1] *= 1e-9
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(abstol=1e-10, reltol=1e-12), emu_ctx)
    R, info = pkg.sample_residuals(s, 2, draws=(Z1, Z2), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], Z1, Z2, atol=1e-10, rtol=1e-12)
    assert info["iters"][1] + 10 < info["iters"][0], info["iters"]
    assert rel(R[:, 0], Ro[:, 0]) < RTOL and np.linalg.norm(R[:, 1] - Ro[:, 1]) < 1e-9 * np.linalg.norm(Ro[:, 0])
    m.close()


@pytest.mark.parametrize("dims", _random_shapes(2026, 24))
def test_metric_random_shapes(pkg, emu_ctx, dims):
    """Seeded random shapes mixing power-of-two (fast kernels) and other axis lengths (generic
    path), 1-D to 3-D: metric product against the oracle, symmetry, and M >= I."""
    like = "LIKE_POISSON" if sum(dims) % 2 else "LIKE_NORMAL"
    cfg = _field(pkg, dims, like, 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), emu_ctx)
    V = np.random.default_rng(3).standard_normal((m.n_theta, 2))
    MV = s.metric_apply(V)
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
    assert rel(MV, MVo) < 1e-12, dims
    a, b = float(V[:, 0] @ MV[:, 1]), float(MV[:, 0] @ V[:, 1])
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), 1.0)
    assert float(V[:, 0] @ MV[:, 0]) >= float(V[:, 0] @ V[:, 0]) * (1 - 1e-14)
    m.close()


# ---- API semantics (host mirror of the reference interface) --------------------------------------
def test_two_samplers_two_centres_own_their_linearisation(pkg, emu_ctx):
    """`ResidualSampler` owns its linearisation in the reference (src/residual_samplers.jl:45-63).  The
    device handle holds one linearisation at a time; a sampler must re-target it to its own centre
    after another sampler, mgvi_step or mgvi_sample used the model."""
    from helpers import TIGHT_O
    cfg = _field(pkg, (16, 16), "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    c1, c2 = cfg["center"], cfg["center"] + 0.3
    s1 = pkg.ResidualSampler(m, c1, pkg.B200CG(**TIGHT), emu_ctx)
    s2 = pkg.ResidualSampler(m, c2, pkg.B200CG(**TIGHT), emu_ctx)          # re-linearises the handle at c2
    R1 = pkg.sample_residuals(s1, 2, draws=(cfg["Z1"], cfg["Z2"]))
    R2 = pkg.sample_residuals(s2, 2, draws=(cfg["Z1"], cfg["Z2"]))
    Ro1, _, _ = O.sample_residuals_cg(om, c1, cfg["Z1"], cfg["Z2"], **TIGHT_O)
    Ro2, _, _ = O.sample_residuals_cg(om, c2, cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert rel(R1, Ro1) < RTOL and rel(R2, Ro2) < RTOL
    assert rel(Ro1, Ro2) > 1e-3                                             # the two centres really differ
    V = np.random.default_rng(5).standard_normal((m.n_theta, 1))
    assert rel(s1.metric_apply(V)[:, 0], O.metric_apply(om.linearize(c1), V[:, 0])) < RTOL
    w2, _ = s2.weights()
    assert rel(w2, om.metric_weight(c2)) < RTOL
    # mgvi_step / mgvi_sample re-linearise the handle as well; s1 must still answer for c1
    pkg.mgvi_sample(m, cfg["data"], 2, c2, pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT)), emu_ctx,
                    draws=(cfg["Z1"], cfg["Z2"]))
    w1, _ = s1.weights()
    assert rel(w1, om.metric_weight(c1)) < RTOL
    m.close()


def test_step_rejects_foreign_data_and_optimizer_opts(pkg, emu_ctx):
    cfg = _field(pkg, (16,), "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, emu_ctx)
    conf = pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT))
    pkg.mgvi_sample(m, None, 2, cfg["center"], conf, emu_ctx, draws=(cfg["Z1"], cfg["Z2"]))       # None: model's own data
    with pytest.raises(pkg.MGVIB200Error):
        pkg.mgvi_step(m, cfg["data"] + 1.0, 2, cfg["center"], conf, emu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    with pytest.raises(pkg.MGVIB200Error):
        pkg.mgvi_sample(m, cfg["data"][:-1], 2, cfg["center"], conf, emu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    with pytest.raises(pkg.MGVIB200Error):
        pkg.mgvi_step(m, cfg["data"], 2, cfg["center"],
                      pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT), optimizer_opts=dict(maxiter=3)), emu_ctx,
                      draws=(cfg["Z1"], cfg["Z2"]))
    with pytest.raises(pkg.MGVIB200Error):
        pkg.mgvi_step(m, cfg["data"], 2, cfg["center"], pkg.MGVIConfig(optimizer="gradient descent"), emu_ctx)
    m.close()


def test_rng_draw_order_matches_the_reference_stream(pkg, emu_ctx):
    """Draws come off the context's stream per residual, n1 (n_lambda) THEN eta (n_theta), residual
    after residual (src/residual_samplers.jl:75-76 inside the loop of :88)."""
    cfg = _field(pkg, (8, 8), "LIKE_NORMAL", 3, 1)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    emu_ctx.rng = np.random.default_rng(77)
    R = pkg.sample_residuals(s, 3)
    rng = np.random.default_rng(77)
    Z1 = np.empty((m.n_lambda, 3)); Z2 = np.empty((m.n_theta, 3))
    for i in range(3):
        Z1[:, i] = rng.standard_normal(m.n_lambda)
        Z2[:, i] = rng.standard_normal(m.n_theta)
    np.testing.assert_array_equal(R, pkg.sample_residuals(s, 3, draws=(Z1, Z2)))
    m.close()


def test_mgvi_sample_matches_oracle(pkg, emu_ctx):
    """`mgvi_sample` (src/mgvi_impl.jl:166-174): samples around the GIVEN centre, no optimisation;
    CG and MatrixInversion samplers."""
    from helpers import TIGHT_O
    cfg = _field(pkg, (16, 8), "LIKE_POISSON", 3)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    S = pkg.mgvi_sample(m, cfg["data"], 3, cfg["center"], pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT)), emu_ctx,
                        draws=(cfg["Z1"], cfg["Z2"]))
    So, _, _ = O.mgvi_sample(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert S.shape == (m.n_theta, 6) and rel(S, So) < RTOL
    assert np.abs(S[:, 0::2] + S[:, 1::2] - 2 * cfg["center"][:, None]).max() < 1e-12
    m.close()
    pcfg = pkg.synthetic.polyfit_config(3, reference_grid=True)
    pm = pkg.model_from_config(pcfg, emu_ctx)
    Sd = pkg.mgvi_sample(pm, pcfg["data"], 3, pcfg["center"], pkg.MGVIConfig(linsolver=pkg.MatrixInversion()), emu_ctx,
                         draws=pcfg["Z"])
    Rd, _ = O.sample_residuals_dense(oracle_model(pcfg), pcfg["center"], pcfg["Z"])
    assert rel(Sd, O.build_samples(Rd, pcfg["center"])) < 1e-9
    pm.close()


def test_lbfgs_optimizer_seam(pkg, emu_ctx):
    """`MGVI._optimize` for a non-NewtonCG optimizer (ext/MGVIOptimExt.jl:12-23,
    ext/MGVIOptimizationBaseExt.jl:19-36; test/test_mgvi_impl.jl:44-60): L-BFGS on the host over the
    device objective.  Checked against the same SciPy driver over the ORACLE's objective."""
    import scipy.optimize
    from helpers import TIGHT_O
    cfg = _field(pkg, (16, 16), "LIKE_POISSON", 3)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    opts = dict(maxiter=12, ftol=0.0, gtol=0.0, maxcor=8)
    conf = pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT), optimizer=pkg.LBFGS(), optimizer_opts=opts)
    res, center = pkg.mgvi_step(m, cfg["data"], 3, cfg["center"], conf, emu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    # objective and gradient of the seam itself, at a point away from the centre
    obj = pkg.KLObjective(m, Ro)
    xp = cfg["center"] + 0.05 * np.random.default_rng(2).standard_normal(m.n_theta)
    f, g = obj(xp)
    assert abs(f - O.mean_neg_log_pstr(om, Ro, xp)) <= RTOL * abs(f) and rel(g, O.kl_gradient(om, Ro, xp)) < RTOL
    assert obj.value(xp) == f
    # the same driver over the oracle's objective: same iterates while rounding has not been amplified
    ref = scipy.optimize.minimize(lambda x: (O.mean_neg_log_pstr(om, Ro, x), O.kl_gradient(om, Ro, x)), cfg["center"],
                                  jac=True, method="L-BFGS-B", options=opts)
    assert rel(center, ref.x) < 1e-8 and abs(res.mnlp - ref.fun) <= 1e-10 * abs(ref.fun)
    assert res.mnlp < O.mean_neg_log_pstr(om, Ro, cfg["center"])
    assert abs(O.mean_neg_log_pstr(om, Ro, center) - res.mnlp) <= RTOL * abs(res.mnlp)      # mnlp IS f(center)
    assert rel(res.samples, O.build_samples(Ro, center)) < RTOL
    assert res.info["optimizer_output"].nit >= 1
    m.close()


def _small_hyper_cfg(pkg, n_data=10, grain=2, n=2, start_scale=0.3):
    """A small instance of the tutorial's model (docs/src/advanced_tutorial_lit.jl): `produce_bins` on a
    10-bin histogram with padding, two hyperparameters in front of the latent field."""
    counts = np.random.default_rng(3).poisson(4.0, n_data)
    return pkg.synthetic.hyper_field_config(counts, n, grain=grain, padding=4.0, xlim=(0.0, 10.0),
                                            start_scale=start_scale)


@pytest.mark.parametrize("grain", [1, 2, 3])
def test_hyper_field_model_vs_oracle(pkg, emu_ctx, grain):
    """amplitude-spectrum hyperparameters (two extra Jacobian columns) + bin aggregation
    (`agg_lambdas`) + Poisson counts -- every row of the path against the oracle, then a full mgvi_step."""
    # (grain 3 triples the kernel amplitude a0 = 2 grain: a start closer to the prior mean keeps exp() finite in
    # the first line search -- beyond that the reference itself returns NaN)
    cfg = _small_hyper_cfg(pkg, grain=grain, start_scale=0.3 if grain < 3 else 0.1)
    out = check_field_ops(pkg, emu_ctx, cfg, check_step=True)
    assert_field_ops(out, cfg)
    # Fisher weights in lambda space: 1 / Lambda_b (src/information.jl:74-78)
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), emu_ctx)
    w, q = s.weights()
    lam = om.lam(cfg["center"])
    assert rel(w, 1.0 / lam) < RTOL and rel(q, np.sqrt(1.0 / lam)) < RTOL
    m.close()


def test_hyper_field_jacobian_consistency(pkg, emu_ctx):
    """test/test_jacobians.jl:28-32,45-49 on the tutorial model: J v against central differences of the
    forward model, and <J v, l> = <v, J' l>; the device metric against the dense J'FJ + I."""
    cfg = _small_hyper_cfg(pkg)
    om = oracle_model(cfg)
    c = cfg["center"]
    lin = om.linearize(c)
    rng = np.random.default_rng(0)
    v, l = rng.standard_normal(om.n_theta), rng.standard_normal(om.n_lambda)
    h = 1e-6
    fd = (om.lam(c + h * v) - om.lam(c - h * v)) / (2 * h)
    assert rel(lin.jvp(v), fd) < 1e-5                          # the reference's own tolerance
    assert abs(lin.jvp(v) @ l - v @ lin.vjp(l)) < 1e-10 * abs(lin.jvp(v) @ l)
    J = lin.dense()
    Md = J.T @ (lin.fisher[:, None] * J) + np.eye(om.n_theta)
    m = pkg.model_from_config(cfg, emu_ctx)
    s = pkg.ResidualSampler(m, c, pkg.B200CG(), emu_ctx)
    assert rel(s.metric_apply(np.eye(om.n_theta)), Md) < RTOL
    m.close()


@pytest.mark.parametrize("kind", ["field1d", "field2d", "field_generic", "hyper"])
def test_matrix_inversion_sampler_on_field_models(pkg, emu_ctx, kind):
    """src/residual_samplers.jl:95-123 on the field families (the dense-vs-CG comparison of
    test/test_samplers.jl:17-39 needs it): L_Sigma and R = L_Sigma Z against the oracle's dense sampler,
    and L_Sigma L_Sigma' = inv(M) against the device's own metric."""
    syn = pkg.synthetic
    cfg = {"field1d": lambda: syn.field_config((32,), syn.LIKE_POISSON, 3, field_std=0.5),
           "field2d": lambda: syn.field_config((8, 16), syn.LIKE_NORMAL, 3, field_std=0.5, layout=1),
           "field_generic": lambda: syn.field_config((6, 10), syn.LIKE_POISSON, 3, field_std=0.5),
           "hyper": lambda: _small_hyper_cfg(pkg, n=3)}[kind]()
    m = pkg.model_from_config(cfg, emu_ctx)
    om = oracle_model(cfg)
    c = cfg["center"]
    Z = np.random.default_rng(42).standard_normal((m.n_theta, 3))
    s = pkg.ResidualSampler(m, c, pkg.MatrixInversion(), emu_ctx)
    R = pkg.sample_residuals(s, 3, draws=Z)
    Ro, Lo = O.sample_residuals_dense(om, c, Z)
    assert rel(R, Ro) < RTOL
    Zf = np.asfortranarray(Z)
    R2 = np.empty((m.n_theta, 3), order="F")
    L = np.empty((m.n_theta, m.n_theta), order="F")
    emu_ctx.check(emu_ctx.lib.sample_residuals_dense(m.handle, pkg.capi.ptr(Zf), 3, pkg.capi.ptr(R2), pkg.capi.ptr(L)))
    assert rel(L, Lo) < RTOL and np.array_equal(R2, R)
    M = s.metric_apply(np.eye(m.n_theta))
    assert rel(L @ L.T @ M, np.eye(m.n_theta)) < 1e-9
    m.close()


def test_dense_linear_per_datum_sigma(pkg, emu_ctx):
    """Diagonal Fisher with one sigma per datum folded into the A-operand load of the DMMA J'FJ GEMM
    (BASELINE north_star d)."""
    from helpers import check_dense_ops
    cfg = pkg.synthetic.dense_linear_config(70, 24, 3)
    cfg["sigma"] = 0.3 + np.random.default_rng(9).random(70)
    cfg["data"] = cfg["J"] @ np.random.default_rng(128).standard_normal(24) + cfg["sigma"] * np.random.default_rng(7).standard_normal(70)
    out = check_dense_ops(pkg, emu_ctx, cfg)
    for k in ("metric", "residuals", "dense", "L", "kl", "grad", "mean_metric", "dense_step_residuals"):
        assert out[k] < RTOL, (k, out)
