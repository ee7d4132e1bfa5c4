"""Parity of the CUDA path (product library, through the C ABI, cuda:0) against the oracle.

Tolerance: relative 1e-10 in Float64 on M.v, CG solutions and the updated mean (BASELINE.json
north_star), CG stopping rule tightened to reltol 1e-12 / abstol 0 on both sides;
bit-exact sample ordering and antithetic pairing.  At BASELINE.json's full sizes the oracle is
used where it finishes in seconds (one metric product) and size-independent properties
elsewhere (symmetry of M, H.H = N I, residual of the CG solution)."""
import numpy as np
import pytest

import oracle as O
from golden_io import load_golden
from helpers import RTOL, TIGHT, TIGHT_O, assert_field_ops, check_field_ops, oracle_model, rel

pytestmark = pytest.mark.gpu


def _field(pkg, dims, like, n, layout=None, std=0.5, **kw):
    syn = pkg.synthetic
    return syn.field_config(dims, getattr(syn, like), n, field_std=std, layout=layout, **kw)


def test_native_library_is_loaded(pkg, gpu_ctx):
    assert gpu_ctx.lib.path.endswith("libmgvib200.so")
    assert b"sm_100a" in gpu_ctx.lib.version()
    maps = open("/proc/self/maps").read()
    assert "libmgvib200.so" in maps


@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_metric_1d_every_length(pkg, gpu_ctx, n):
    cfg = _field(pkg, (n,), "LIKE_POISSON", 5)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), gpu_ctx)
    lin = om.linearize(cfg["center"])
    V = np.random.default_rng(5).standard_normal((n, 5))
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(5)], axis=1)
    assert rel(s.metric_apply(V), MVo) < RTOL
    m.close()


@pytest.mark.parametrize("dims,like,layout,n", [
    ((4096,), "LIKE_POISSON", None, 4),
    ((256, 256), "LIKE_NORMAL", 1, 4),
    ((128, 512), "LIKE_POISSON", None, 3),
    ((8, 2048), "LIKE_NORMAL", 2, 2),
    ((2048, 8), "LIKE_POISSON", None, 2),
    ((8, 4096), "LIKE_POISSON", None, 2),   # 512-thread row blocks
    ((4096, 16), "LIKE_NORMAL", 0, 2),      # 1024-thread column tiles
    ((256, 64, 8), "LIKE_POISSON", None, 2),    # 3-D with a 256-point axis 0: compile-time-length fused kernel
    ((16, 256, 16), "LIKE_NORMAL", 0, 2),       # ... with the 8-column tile of the 256^3 workload
    ((32, 64, 16), "LIKE_POISSON", None, 3),
    ((16, 8, 128), "LIKE_NORMAL", 0, 2),
    ((40,), "LIKE_NORMAL", 2, 4),          # reference fixture length, generic path
    ((30, 50), "LIKE_POISSON", None, 3),   # non power of two 2-D, generic path
])
def test_field_ops_vs_oracle(pkg, gpu_ctx, dims, like, layout, n):
    cfg = _field(pkg, dims, like, n, layout)
    assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


def test_identity_link(pkg, gpu_ctx):
    syn = pkg.synthetic
    cfg = syn.field_config((64, 64), syn.LIKE_NORMAL, 3, field_std=0.5)
    cfg["link"] = syn.LINK_IDENTITY
    assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


@pytest.mark.parametrize("name", ["fft_gp_40", "field_1d_64_poisson", "field_2d_16x32_normal",
                                  "field_3d_8x8x16_poisson"])
def test_golden_vectors(pkg, gpu_ctx, name):
    cfg, gold = load_golden(name)
    m = pkg.model_from_config(cfg, gpu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), gpu_ctx)
    assert rel(s.metric_apply(gold["V"]), gold["MV"]) < RTOL
    R = pkg.sample_residuals(s, cfg["n_residuals"], draws=(cfg["Z1"], cfg["Z2"]))
    assert rel(R, gold["R"]) < RTOL
    f, g = pkg.kl_value_grad(m, cfg["center"], gold["R"])
    assert abs(f - gold["f"]) / abs(gold["f"]) < RTOL and rel(g, gold["grad"]) < RTOL
    res, c = pkg.mgvi_step(m, cfg["data"], cfg["n_residuals"], cfg["center"],
                           pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT)), gpu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
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


def test_default_tolerance_iteration_counts(pkg, gpu_ctx):
    cfg = _field(pkg, (128, 128), "LIKE_POISSON", 4)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), gpu_ctx)
    R, info = pkg.sample_residuals(s, 4, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, rno = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"])
    # ||r|| is not monotone in CG: the first dip below the threshold moves by a few iterations
    assert np.all(np.abs(info["iters"] - ito) <= 3 + 0.05 * ito), (info["iters"], ito)
    assert rel(R, Ro) < 1e-6          # both stop at ||r|| ~ 1.5e-8 (1 + ||b||)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(abstol=0.0, reltol=0.0, maxiters=9), gpu_ctx)
    R, info = pkg.sample_residuals(s, 4, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], atol=0.0, rtol=0.0, itmax=9)
    assert info["iters"].tolist() == [9] * 4 == ito.tolist()
    assert rel(R, Ro) < RTOL
    m.close()


def test_many_residuals_lockstep_masking(pkg, gpu_ctx):
    """33 residuals (odd count) with very different right-hand-side scales converge at different
    iterations; every one must match its own independent oracle solve."""
    cfg = _field(pkg, (64, 32), "LIKE_POISSON", 33)
    scale = np.logspace(-6, 3, 33)
    cfg["Z1"] = cfg["Z1"] * scale
    cfg["Z2"] = cfg["Z2"] * scale
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(abstol=1e-9, reltol=1e-12), gpu_ctx)
    R, info = pkg.sample_residuals(s, 33, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], atol=1e-9, rtol=1e-12)
    assert len(set(info["iters"].tolist())) > 3
    assert np.all(np.abs(info["iters"] - ito) <= 3 + 0.05 * ito)
    for i in range(33):
        assert np.linalg.norm(R[:, i] - Ro[:, i]) <= 1e-8 * np.linalg.norm(Ro[:, i]) + 2e-9, i
    m.close()


@pytest.mark.parametrize("dims", [(2048, 32), (1024, 256), (64, 32, 16), (2048, 2048)])
def test_deferred_solution_update_bitwise(pkg, gpu_ctx, monkeypatch, dims):
    """x += alpha p inside the fused axis-0 pass of the next iteration (field_pass.h: defer_x; default for 2-D
    fields with a long axis 0, forced here for the others) against the streaming update kernel: bit for bit,
    with residuals that converge at different iterations and an iteration cap that stops mid-flight (CUDA-graph
    replay of 8-iteration chunks included)."""
    cfg = _field(pkg, dims, "LIKE_NORMAL", 5)
    scale = np.logspace(-6, 2, 5)
    Z1, Z2 = cfg["Z1"] * scale, cfg["Z2"] * scale
    out = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("MGVIB200_DEFER_X", mode)
        m = pkg.model_from_config(cfg, gpu_ctx)
        res = []
        for opts in (dict(abstol=1e-9, reltol=1e-12), dict(abstol=0.0, reltol=0.0, maxiters=1),
                     dict(abstol=0.0, reltol=0.0, maxiters=13), dict(abstol=1e-3, reltol=0.0, maxiters=40)):
            s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**opts), gpu_ctx)
            R, info = pkg.sample_residuals(s, 5, draws=(Z1, Z2), return_info=True)
            res.append((R.copy(), info["iters"].copy()))
        out[mode] = res
        m.close()
    for mode in ("1", "2"):
        for (Ra, ia), (Rb, ib) in zip(out["0"], out[mode]):
            assert np.array_equal(ia, ib) and np.array_equal(Ra, Rb), mode
    assert len(set(out["1"][0][1].tolist())) >= 3


@pytest.mark.parametrize("dims,n", [((2048, 64), 4), ((256, 256), 5), ((32, 64, 16), 3), ((1024, 1024), 2), ((64, 64), 7),
                                    ((65536,), 8), ((4096,), 6), ((8192,), 5), ((256,), 4)])
def test_two_sample_groups_bitwise(pkg, gpu_ctx, monkeypatch, dims, n):
    """Few samples per GPU: the batched CG runs as two independent lock-step solves on two streams
    (field_engine.cu: sample_cg_grouped).  Same kernels, same per-sample arithmetic: residuals, iteration counts
    and final residual norms must equal the single-stream path bit for bit -- odd sample counts, samples that
    converge at different iterations, iteration caps inside and at the end of a replayed chunk."""
    cfg = _field(pkg, dims, "LIKE_POISSON", n)
    scale = np.logspace(-5, 2, n)
    Z1, Z2 = cfg["Z1"] * scale, cfg["Z2"] * scale
    out = {}
    for mode in ("1", "2"):
        monkeypatch.setenv("MGVIB200_GROUPS", mode)
        m = pkg.model_from_config(cfg, gpu_ctx)
        res = []
        for opts in (dict(abstol=1e-9, reltol=1e-12), dict(abstol=0.0, reltol=0.0, maxiters=1),
                     dict(abstol=0.0, reltol=0.0, maxiters=9), dict(abstol=0.0, reltol=0.0, maxiters=17),
                     dict(abstol=1e-3, reltol=0.0, maxiters=40)):
            s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**opts), gpu_ctx)
            R, info = pkg.sample_residuals(s, n, draws=(Z1, Z2), return_info=True)
            res.append((R.copy(), info["iters"].copy(), info["final_rnorm"].copy()))
        out[mode] = res
        m.close()
    for (Ra, ia, ra), (Rb, ib, rb) in zip(out["1"], out["2"]):
        assert np.array_equal(ia, ib) and np.array_equal(Ra, Rb)
        if ra is not None:
            assert np.array_equal(ra, rb)


def test_consecutive_steps_reuse_cached_graphs(pkg, gpu_ctx):
    """Three MGVI iterations on one model handle (new centre and new draws each time): the cached CUDA graphs of
    the two sample groups, the re-laid weight and the deferred-update buffers are reused across re-linearisations.
    The residual samples of every step (they do not pass through the optimiser) must match the oracle's CG solves at
    that step's centre at 1e-10, and the KL estimate must decrease."""
    cfg = _field(pkg, (1024, 64), "LIKE_NORMAL", 4)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    c = cfg["center"].copy()
    conf = pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT))
    mnlp = []
    for it in range(3):
        rng = np.random.default_rng(100 + it)
        Z1 = np.asfortranarray(rng.standard_normal(cfg["Z1"].shape))
        Z2 = np.asfortranarray(rng.standard_normal(cfg["Z2"].shape))
        Ro, _, _ = O.sample_residuals_cg(om, c, Z1, Z2, **TIGHT_O)
        res, c_new = pkg.mgvi_step(m, cfg["data"], 4, c, conf, gpu_ctx, draws=(Z1, Z2))
        Rdev = 0.5 * (res.samples[:, 0::2] - res.samples[:, 1::2])
        assert rel(Rdev, Ro) < RTOL, (it, rel(Rdev, Ro))
        f_at = O.mean_neg_log_pstr(om, Rdev, c_new)
        assert abs(f_at - res.mnlp) <= RTOL * abs(f_at)
        tr = res.info["optimizer_output"]
        assert tr["f_history"][-1] <= tr["f_history"][0]
        mnlp.append(res.mnlp)
        c = c_new
    m.close()


# ---- BASELINE.json full sizes -------------------------------------------------------------------
def _full_metric_checks(pkg, gpu_ctx, cfg, n_solve):
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    c = cfg["center"]
    s = pkg.ResidualSampler(m, c, pkg.B200CG(abstol=0.0, reltol=1e-10), gpu_ctx)
    lin = om.linearize(c)
    rng = np.random.default_rng(5)
    V = rng.standard_normal((m.n_theta, 2))
    MV = s.metric_apply(V)
    # direct parity with the oracle on one product (the oracle needs ~1 s at this size)
    assert rel(MV[:, 0], O.metric_apply(lin, V[:, 0])) < RTOL
    # symmetry <u, M v> = <M u, v> and M >= I
    a, b = float(V[:, 0] @ MV[:, 1]), float(MV[:, 0] @ V[:, 1])
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b), np.linalg.norm(V[:, 0]) * np.linalg.norm(MV[:, 1]))
    assert float(V[:, 0] @ MV[:, 0]) >= float(V[:, 0] @ V[:, 0])
    # CG solutions satisfy M x = b (b rebuilt by the oracle from the same draws)
    Z1 = np.asfortranarray(cfg["Z1"][:, :n_solve])
    Z2 = np.asfortranarray(cfg["Z2"][:, :n_solve])
    R, info = pkg.sample_residuals(s, n_solve, draws=(Z1, Z2), return_info=True)
    for i in range(n_solve):
        bvec = O.sampler_rhs(lin, Z1[:, i], Z2[:, i])
        res = O.metric_apply(lin, R[:, i]) - bvec
        assert np.linalg.norm(res) <= 2e-10 * np.linalg.norm(bvec), (i, info)
    m.close()
    return info


def test_full_size_c3_2048x2048(pkg, gpu_ctx):
    cfg = _field(pkg, (2048, 2048), "LIKE_NORMAL", 2)
    _full_metric_checks(pkg, gpu_ctx, cfg, 2)


@pytest.mark.parametrize("dims,like", [((4096, 4096), "LIKE_POISSON"), ((2048, 512), "LIKE_NORMAL"),
                                       ((512, 2048), "LIKE_POISSON"), ((1024, 4096), "LIKE_NORMAL"),
                                       ((8192, 512), "LIKE_POISSON"), ((256, 8192), "LIKE_NORMAL"),
                                       ((16, 8192, 16), "LIKE_POISSON")])
def test_large_2d_shapes(pkg, gpu_ctx, dims, like):
    """Largest supported axis lengths (8192 points in 2-D / 3-D: one complex line fills 128 KB of shared memory)
    and non-square fields: 1024-thread column tiles, tile-major scratch with 2- and 4-column tiles next to
    natural-layout 8-column tiles."""
    cfg = _field(pkg, dims, like, 1)
    _full_metric_checks(pkg, gpu_ctx, cfg, 1)


@pytest.mark.parametrize("dims,like", [((1000,), "LIKE_POISSON"), ((4095,), "LIKE_NORMAL"), ((1000, 1000), "LIKE_POISSON"),
                                       ((1536, 640), "LIKE_NORMAL"), ((100, 90, 80), "LIKE_POISSON"), ((3000, 7), "LIKE_NORMAL")])
def test_any_length_axes(pkg, gpu_ctx, dims, like):
    """Axis lengths that are not a power of two: the O(m log m) Bluestein pass
    (field_pass.h: bluestein_axis_body) up to the 4096-point limit, 1-D to 3-D, next to short direct-pass axes."""
    cfg = _field(pkg, dims, like, 1)
    _full_metric_checks(pkg, gpu_ctx, cfg, 1)


def test_any_length_full_ops_1d_1000(pkg, gpu_ctx):
    """Every row of the path on a 1000-pixel line and a 96 x 80 field (Bluestein axes), end to end incl. the step."""
    for dims, like in (((1000,), "LIKE_POISSON"), ((96, 80), "LIKE_NORMAL")):
        cfg = _field(pkg, dims, like, 4)
        assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


def test_full_size_c5_256cubed(pkg, gpu_ctx):
    cfg = _field(pkg, (256, 256, 256), "LIKE_POISSON", 1)
    _full_metric_checks(pkg, gpu_ctx, cfg, 1)


def test_full_size_step_properties_1024(pkg, gpu_ctx):
    """A complete mgvi_step at 1024 x 1024 with 4 residuals: antithetic pairing is exact and the
    KL estimate decreases; the KL value of the result is reproduced by the oracle."""
    cfg = _field(pkg, (1024, 1024), "LIKE_NORMAL", 4)
    m = pkg.model_from_config(cfg, gpu_ctx)
    res, c = pkg.mgvi_step(m, cfg["data"], 4, cfg["center"], pkg.MGVIConfig(), gpu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    S = res.samples
    assert np.abs(S[:, 0::2] + S[:, 1::2] - 2 * c[:, None]).max() < 1e-12
    tr = res.info["optimizer_output"]
    assert tr["f_history"][-1] < tr["f_history"][0]
    om = oracle_model(cfg)
    R = 0.5 * (S[:, 0::2] - S[:, 1::2])
    assert abs(O.mean_neg_log_pstr(om, R, c) - res.mnlp) <= 1e-10 * abs(res.mnlp)
    m.close()


@pytest.mark.parametrize("layout,reference_grid", [(1, True), (2, True), (1, False)])
def test_polyfit_model(pkg, gpu_ctx, layout, reference_grid):
    """BASELINE.json configs[0]: the polynomial fit of test/test_models/model_polyfit.jl on the
    reference's 40-point grids and on the 1000-point C1 grids, 8 residual samples."""
    from helpers import check_poly_ops
    cfg = pkg.synthetic.polyfit_config(8, reference_grid=reference_grid, layout=layout)
    out = check_poly_ops(pkg, gpu_ctx, cfg)
    for k in ("metric", "fisher", "dense", "kl", "grad", "mean_metric"):
        assert out[k] < 1e-9, (k, out)
    assert out["iters_default"][0] == out["iters_default"][1] == [5] * 8
    assert out["residuals"] < 1e-8, out
    assert out["cg_step"] < 1e-8 and out["dense_step"] < 1e-9, out


@pytest.mark.parametrize("m,k,n", [(96, 40, 3), (700, 300, 5), (2048, 1024, 8)])
def test_dense_linear_model(pkg, gpu_ctx, m, k, n):
    """BASELINE.json configs[3] family: dense J, FP64 tensor-core J'FJ, blocked Cholesky, both
    samplers, against the oracle (LAPACK)."""
    from helpers import check_dense_ops
    cfg = pkg.synthetic.dense_linear_config(m, k, n)
    out = check_dense_ops(pkg, gpu_ctx, cfg, with_step=(k <= 300))
    for key in ("metric", "residuals", "dense", "L", "kl", "grad", "mean_metric"):
        assert out[key] < 1e-10, (key, out)
    assert out["L_lower"] == 0.0
    assert all(abs(a - b) <= 2 for a, b in zip(*out["iters"]))


def test_dense_c4_full_size_properties(pkg, gpu_ctx):
    """C4 itself (J 16384 x 4096): L_Sigma is lower triangular with L L' M = I on probe vectors, and
    the residuals are L_Sigma Z."""
    cfg = pkg.synthetic.dense_linear_config(16384, 4096, 4)
    m = pkg.model_from_config(cfg, gpu_ctx)
    sd = pkg.ResidualSampler(m, cfg["center"], pkg.MatrixInversion(), gpu_ctx)
    Z = np.asfortranarray(cfg["Z"])
    R = np.empty((4096, 4), order="F")
    L = np.empty((4096, 4096), order="F")
    gpu_ctx.check(gpu_ctx.lib.sample_residuals_dense(m.handle, pkg.capi.ptr(Z), 4, pkg.capi.ptr(R), pkg.capi.ptr(L)))
    assert np.abs(np.triu(L, 1)).max() == 0.0
    assert rel(R, L @ Z) < 1e-12
    V = np.random.default_rng(1).standard_normal((4096, 2))
    MV = sd.metric_apply(V)                                  # M V
    assert rel(L @ (L.T @ MV), V) < 1e-10                    # (L L') M = I
    J = cfg["J"]
    assert rel(MV, V + J.T @ (J @ V) / cfg["sigma"] ** 2) < 1e-11
    m.close()


def test_c2_1d_65536_poisson(pkg, gpu_ctx):
    """BASELINE.json configs[1]: 1-D field of 2^16 pixels, Poisson counts, 16 residual samples
    (four-step transform path), every row of the path against the oracle."""
    cfg = _field(pkg, (65536,), "LIKE_POISSON", 16)
    assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


@pytest.mark.parametrize("N", [8192, 32768])
def test_long_1d_other_lengths(pkg, gpu_ctx, N):
    cfg = _field(pkg, (N,), "LIKE_NORMAL", 3, layout=2)
    assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


def test_long_1d_2pow20(pkg, gpu_ctx):
    """2^20 pixels (n1 = n2 = 1024 in the four-step split): one metric product against the oracle,
    symmetry, and M x = b for a CG solution (a full oracle CG solve at this size takes minutes)."""
    cfg = _field(pkg, (1 << 20,), "LIKE_NORMAL", 2, layout=2)
    _full_metric_checks(pkg, gpu_ctx, cfg, 2)


@pytest.mark.parametrize("dims", [(1,), (2,), (3,), (1, 8), (8, 1), (1, 1, 8), (1, 5), (2, 2)])
def test_degenerate_shapes(pkg, gpu_ctx, dims):
    """Singleton axes, one-pixel and two-pixel fields."""
    from helpers import TIGHT_O
    cfg = _field(pkg, dims, "LIKE_POISSON", 2)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), gpu_ctx)
    V = np.random.default_rng(1).standard_normal((m.n_theta, 2))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12
    R = pkg.sample_residuals(s, 2, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert rel(R, Ro) < RTOL
    m.close()


def test_empty_batches(pkg, gpu_ctx):
    """Zero residual samples / probe vectors give n_theta x 0 results without a launch."""
    cfg = _field(pkg, (16, 8), "LIKE_NORMAL", 2)
    m = pkg.model_from_config(cfg, gpu_ctx)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(), gpu_ctx)
    assert pkg.sample_residuals(s, 0).shape == (m.n_theta, 0)
    assert s.metric_apply(np.zeros((m.n_theta, 0))).shape == (m.n_theta, 0)
    with pytest.raises(pkg.MGVIB200Error):
        pkg.kl_value_grad(m, cfg["center"], np.zeros((m.n_theta, 0)))
    m.close()


def _random_shapes(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        d = int(rng.integers(1, 4))
        dims = tuple(int(rng.choice([1, 3, 5, 8, 8, 12, 16, 16, 32, 32, 64, 128, 256, 512])) for _ in range(d))
        if 2 <= int(np.prod(dims)) <= 65536:
            out.append(dims)
    return out


def _random_shapes_any_length(seed, count):
    """like _random_shapes, with axis lengths on both sides of the Bluestein threshold (32) and of powers of two"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        d = int(rng.integers(1, 4))
        dims = tuple(int(rng.choice([2, 7, 16, 31, 33, 40, 63, 64, 65, 100, 127, 129, 250, 256, 1000])) for _ in range(d))
        if 2 <= int(np.prod(dims)) <= 70000:
            out.append(dims)
    return out


@pytest.mark.parametrize("dims", _random_shapes(7, 16) + _random_shapes_any_length(11, 14))
def test_random_shapes(pkg, gpu_ctx, dims):
    """Seeded random shapes (power-of-two and other axis lengths, 1-D to 3-D): metric product and
    CG residual samples against the oracle."""
    from helpers import TIGHT_O
    like = "LIKE_POISSON" if sum(dims) % 2 else "LIKE_NORMAL"
    cfg = _field(pkg, dims, like, 3)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    s = pkg.ResidualSampler(m, cfg["center"], pkg.B200CG(**TIGHT), gpu_ctx)
    V = np.random.default_rng(3).standard_normal((m.n_theta, 2))
    lin = om.linearize(cfg["center"])
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(2)], axis=1)
    assert rel(s.metric_apply(V), MVo) < 1e-12, dims
    R, info = pkg.sample_residuals(s, 3, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    # tiny fields: the solve runs into the reference's iteration cap (itmax = n_theta) before reltol 1e-12; the last
    # iterates of such a truncated CG amplify rounding (measured on the emulator at 16 pixels: 3e-11 between two
    # correct implementations), so only there the bar is 1e-8
    capped = bool(np.any(ito >= om.n_theta))
    assert rel(R, Ro) < (1e-8 if capped else RTOL), (dims, info["iters"], ito)
    m.close()


@pytest.mark.parametrize("m,k,n", [(65, 64, 1), (130, 65, 2), (200, 127, 3), (257, 128, 2), (400, 129, 4), (333, 191, 2),
                                   (512, 192, 1), (300, 193, 3)])
def test_dense_block_edges(pkg, gpu_ctx, m, k, n):
    """Dense path around the Cholesky block size (64) and the GEMM tile (128): k = 64, 65, 127,
    128, 129, 191, 192, 193 columns; odd row counts."""
    from helpers import check_dense_ops
    cfg = pkg.synthetic.dense_linear_config(m, k, n)
    out = check_dense_ops(pkg, gpu_ctx, cfg, with_step=False)
    for key in ("metric", "residuals", "dense", "L", "kl", "grad", "mean_metric"):
        assert out[key] < 1e-10, (key, out)
    assert out["L_lower"] == 0.0


# ------------------------------------------------------------------------------------------------
# the tutorial's model (amplitude-spectrum hyperparameters + bin aggregation) and the
# coal-mining run of docs/src/advanced_tutorial_lit.jl:489-556
# ------------------------------------------------------------------------------------------------
def _coal_cfg(pkg, n=3, start_scale=1.0):
    from golden_io import GOLDEN
    import os
    counts = np.load(os.path.join(GOLDEN, "coal_mining_counts.npz"))["counts"]
    return pkg.synthetic.hyper_field_config(counts, n, start_scale=start_scale)


def test_hyper_field_every_row_at_tutorial_size(pkg, gpu_ctx):
    """The tutorial's grid (110 yearly bins, GP_GRAIN_FACTOR 3, padding 80: ~800 fine pixels + 2
    hyperparameters), 3 residuals: metric product, CG residuals, KL value / gradient, mean metric (n
    per-sample metrics, canonical sum), sample assembly and one full mgvi_step against the oracle."""
    cfg = _coal_cfg(pkg, start_scale=0.3)
    assert_field_ops(check_field_ops(pkg, gpu_ctx, cfg, check_step=True), cfg)


def test_coal_mining_tutorial_run(pkg, gpu_ctx):
    """docs/src/advanced_tutorial_lit.jl:489-556: first iteration with KrylovJL_CG(itmax = 10) and
    NewtonCG(steps = 1), then iterations with KrylovJL_CG(atol = 1e-2) and the default NewtonCG, 3
    residuals each, every iteration fed the previous centre.  The chain is compared ITERATION BY
    ITERATION: the device's mgvi_step starts from the oracle's centre of that iteration with the same
    draws (teacher forcing: end-to-end chains of truncated Newton steps amplify rounding), and a
    free-running device chain must reproduce the tutorial's convergence."""
    from helpers import assert_step_parity
    cfg = _coal_cfg(pkg, start_scale=1.0)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    n, nt, nl = 3, cfg["n_theta"], cfg["n_lambda"]
    rng = np.random.default_rng(2024)

    def draws():
        return rng.standard_normal((nl, n)), rng.standard_normal((nt, n))

    def avg_nlp(samples):                                # compute_avg_likelihood (:533-539)
        return float(np.mean([-O.posterior_loglike(om, samples[:, j]) for j in range(samples.shape[1])]))

    c_or = cfg["center"].copy()
    c_dev = cfg["center"].copy()
    series_or, series_dev = [], []
    stages = [(dict(maxiters=10), dict(steps=1), dict(itmax=10))] + \
             [(dict(abstol=1e-2), dict(), dict(atol=1e-2))] * 8
    for it, (cg_kw, ncg_kw, ocg_kw) in enumerate(stages):
        Z1, Z2 = draws()
        conf = pkg.MGVIConfig(linsolver=pkg.B200CG(**cg_kw), optimizer=pkg.NewtonCG(**ncg_kw))
        ref = O.mgvi_step(om, c_or, Z1=Z1, Z2=Z2, opts=O.NewtonCGOpts(**ncg_kw), **ocg_kw)
        # The tutorial stops the sampler CG early (itmax = 10 / atol = 1e-2).  The k-th iterate of a truncated CG
        # is an ill-conditioned function of its inputs (measured on this very start: one-ulp input perturbations
        # move the ORACLE's 10th iterate by 1e-5, its 4th by 3e-15), so the residual tolerance is the oracle's
        # own measured reproducibility of exactly this solve, x 100 -- and never below 1e-10.
        sens = 0.0
        for sd in range(2):
            pr = np.random.default_rng(500 + sd)
            Rp, itp, _ = O.sample_residuals_cg(om, c_or * (1 + np.finfo(float).eps * pr.standard_normal(nt)), Z1,
                                               Z2 * (1 + np.finfo(float).eps * pr.standard_normal(Z2.shape)), **ocg_kw)
            sens = max(sens, rel(Rp, ref["residuals"]))
        rtol_R = max(RTOL, 100.0 * sens)
        assert rtol_R < 1e-2, (it, sens)
        # teacher-forced: the device's step from the oracle's centre
        res, cn = pkg.mgvi_step(m, cfg["data"], n, c_or, conf, gpu_ctx, draws=(Z1, Z2))
        it_dev, it_ref = res.info["linsolver_output"]["iters"], ref["sampler_iters"]
        if it_dev.tolist() != it_ref.tolist():
            # ||r|| sits on the atol = 1e-2 threshold within the amplified rounding (and is not monotone in CG): a
            # sample may stop a few iterations apart (seen: 12 vs 10 at a measured oracle reproducibility of 7e-5).
            # Compare against the oracle stopped at the DEVICE's counts -- same algorithm, same draws.
            assert np.abs(it_dev - it_ref).max() <= 3 and sens > 1e-12, (it, it_dev, it_ref, sens)
            Rr = ref["residuals"].copy()
            for i in np.nonzero(it_dev != it_ref)[0]:
                Ri, iti, _ = O.sample_residuals_cg(om, c_or, Z1[:, [i]], Z2[:, [i]], atol=0.0, rtol=0.0, itmax=int(it_dev[i]))
                assert int(iti[0]) == int(it_dev[i])
                Rr[:, i] = Ri[:, 0]
            xo, fo, tro = O.newtoncg_optimize(om, Rr, c_or, O.NewtonCGOpts(**ncg_kw))
            ref = dict(samples=O.build_samples(Rr, xo), mnlp=fo, center=xo, residuals=Rr, sampler_iters=it_dev, trace=tro)
        step_cfg = dict(cfg, center=c_or)
        assert_step_parity(pkg, m, om, step_cfg, res, cn, ref, label="coal mining it %d" % it, opts_kw=ncg_kw,
                           residual_tol=rtol_R)
        # the solve itself, where it is well-conditioned: the first 4 CG iterations at 1e-10
        s4 = pkg.ResidualSampler(m, c_or, pkg.B200CG(maxiters=4), gpu_ctx)
        R4 = pkg.sample_residuals(s4, n, draws=(Z1, Z2))
        R4o, _, _ = O.sample_residuals_cg(om, c_or, Z1, Z2, itmax=4)
        assert rel(R4, R4o) < RTOL, (it, rel(R4, R4o))
        series_or.append(avg_nlp(ref["samples"]))
        # free-running device chain
        res_d, c_dev = pkg.mgvi_step(m, cfg["data"], n, c_dev, conf, gpu_ctx, draws=(Z1, Z2))
        series_dev.append(avg_nlp(res_d.samples))
        c_or = ref["center"]
    # the tutorial's observation: the average -log posterior of the samples falls steeply and flattens
    assert series_dev[-1] < series_dev[0] and series_or[-1] < series_or[0]
    assert abs(series_dev[-1] - series_or[-1]) < 0.05 * abs(series_or[-1]), (series_dev, series_or)
    # posterior mean rate is positive and of the order of the observed counts
    lam = om.lam(c_dev)
    assert np.all(lam > 0) and 0.2 < lam.sum() / cfg["data"].sum() < 5.0
    m.close()


@pytest.mark.parametrize("kind", ["field1d", "field2d", "hyper"])
def test_matrix_inversion_on_field_models_and_bartlett(pkg, gpu_ctx, kind):
    """src/residual_samplers.jl:95-123 on the field families, and the reference's own sampler test
    (test/test_samplers.jl:17-39) on them: L_Sigma against the oracle (LAPACK); the empirical covariance
    of many CG residual samples against Sigma = L_Sigma L_Sigma' within the reference's factor-of-10
    element tolerance and far tighter in the trace."""
    syn = pkg.synthetic
    cfg = {"field1d": lambda: syn.field_config((256,), syn.LIKE_POISSON, 4, field_std=0.5),
           "field2d": lambda: syn.field_config((16, 32), syn.LIKE_NORMAL, 4, field_std=0.5, layout=1),
           "hyper": lambda: _coal_cfg(pkg, 4, start_scale=0.3)}[kind]()
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    c, nt, nl = cfg["center"], cfg["n_theta"], cfg["n_lambda"]
    Z = np.asfortranarray(np.random.default_rng(42).standard_normal((nt, 4)))
    sd = pkg.ResidualSampler(m, c, pkg.MatrixInversion(), gpu_ctx)
    R = np.empty((nt, 4), order="F")
    L = np.empty((nt, nt), order="F")
    gpu_ctx.check(gpu_ctx.lib.sample_residuals_dense(m.handle, pkg.capi.ptr(Z), 4, pkg.capi.ptr(R), pkg.capi.ptr(L)))
    Ro, Lo = O.sample_residuals_dense(om, c, Z)
    assert rel(R, Ro) < RTOL and rel(L, Lo) < RTOL
    assert np.abs(np.triu(L, 1)).max() == 0.0
    # Bartlett-style comparison: covariance of n_s CG samples vs Sigma
    ns = 4000
    rng = np.random.default_rng(7)
    sc = pkg.ResidualSampler(m, c, pkg.B200CG(), gpu_ctx)
    Rcg = pkg.sample_residuals(sc, ns, draws=(rng.standard_normal((nl, ns)), rng.standard_normal((nt, ns))))
    Sig = L @ L.T
    emp = Rcg @ Rcg.T / ns
    d_emp, d_sig = np.diag(emp), np.diag(Sig)
    assert np.all(d_emp < 10 * d_sig) and np.all(d_sig < 10 * d_emp)           # test_samplers.jl:31,37-38
    assert abs(np.trace(emp) - np.trace(Sig)) < 0.05 * np.trace(Sig)
    # whitened samples have unit covariance: mean squared norm of L^-1 r = n_theta
    import scipy.linalg
    wn = scipy.linalg.solve_triangular(L, Rcg, lower=True)
    assert abs(np.mean(np.sum(wn * wn, axis=0)) / nt - 1.0) < 0.02
    m.close()


def test_dense_linear_per_datum_sigma(pkg, gpu_ctx):
    """Per-datum Fisher weights folded into the A-operand load of the DMMA J'FJ GEMM, at a size that
    crosses GEMM tile and Cholesky block edges."""
    from helpers import check_dense_ops
    mm, k, n = 1500, 333, 5
    cfg = pkg.synthetic.dense_linear_config(mm, k, n)
    cfg["sigma"] = 0.3 + np.random.default_rng(9).random(mm)
    cfg["data"] = cfg["J"] @ np.random.default_rng(128).standard_normal(k) + cfg["sigma"] * np.random.default_rng(7).standard_normal(mm)
    out = check_dense_ops(pkg, gpu_ctx, cfg)
    for key in ("metric", "residuals", "dense", "L", "kl", "grad", "mean_metric", "dense_step_residuals"):
        assert out[key] < RTOL, (key, out)


def test_mgvi_sample_and_pushfwd_on_gpu(pkg, gpu_ctx):
    """`mgvi_sample` (src/mgvi_impl.jl:166-174: samples around a fixed centre, no
    optimisation) and `mgvi_mvnormal_pushfwd_function` (:191-198) on the product library."""
    syn = pkg.synthetic
    cfg = syn.field_config((32, 64), syn.LIKE_POISSON, 4, field_std=0.5)
    m = pkg.model_from_config(cfg, gpu_ctx)
    om = oracle_model(cfg)
    c = cfg["center"]
    conf = pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT))
    S = pkg.mgvi_sample(m, cfg["data"], 4, c, conf, gpu_ctx, draws=(cfg["Z1"], cfg["Z2"]))
    Ro, _, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert S.shape == (m.n_theta, 8) and rel(S, O.build_samples(Ro, c)) < RTOL
    assert np.abs(S[:, 0::2] + S[:, 1::2] - 2 * c[:, None]).max() < 1e-12          # antithetic pairs around the GIVEN centre
    # ... and with the MatrixInversion sampler on a field model
    Zd = np.random.default_rng(11).standard_normal((m.n_theta, 4))
    Sd = pkg.mgvi_sample(m, cfg["data"], 4, c, pkg.MGVIConfig(linsolver=pkg.MatrixInversion()), gpu_ctx, draws=Zd)
    Rd, _ = O.sample_residuals_dense(om, c, Zd)
    assert rel(Sd, O.build_samples(Rd, c)) < RTOL
    # push-forward: x -> centre + L_Sigma x on a small field (dense operator)
    scfg = syn.field_config((64,), syn.LIKE_POISSON, 2, field_std=0.5)
    sm = pkg.model_from_config(scfg, gpu_ctx)
    som = oracle_model(scfg)
    f = pkg.mgvi_mvnormal_pushfwd_function(sm, scfg["data"], pkg.MGVIConfig(), scfg["center"], gpu_ctx)
    X = np.random.default_rng(3).standard_normal((64, 5))
    _, Lo = O.sample_residuals_dense(som, scfg["center"], X)
    Y = np.stack([f(X[:, j]) for j in range(5)], axis=1)
    assert rel(Y, scfg["center"][:, None] + Lo @ X) < RTOL
    m.close(); sm.close()
