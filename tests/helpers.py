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
    if cfg["kind"] == "hyper":
        return O.HyperFieldModel(cfg["kdist"], cfg["scale"], cfg["data_idxs"], cfg["grain"], cfg["binsize"], cfg["data"],
                                 a0=cfg["a0"], a1=cfg["a1"], l0=cfg["l0"], l1=cfg["l1"])
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
        out["step_regime"] = assert_step_parity(pkg, m, om, cfg, res, cn, ref,
                                                label="%s %s" % (cfg["kind"], "x".join(map(str, cfg.get("dims", ())))))
    m.close()
    return out


# ------------------------------------------------------------------------------------------------
# Newton phase (rows a12 / a14): teacher-forced, per-step parity.
#
# MGVI's NewtonCG (src/newtoncg.jl:85-167) is a TRUNCATED iteration with data-dependent exits (the
# KL-stagnation test :134, the strong-Wolfe bracket/zoom branches :142).  On ill-conditioned
# fixtures a 1e-15 perturbation of the inputs moves the ORACLE's own end-to-end result by up to
# O(1) (measured: 2e-2 at 64 x 64, 0.89 at 256 x 256), so comparing only the final centre says
# nothing there.  Instead every Newton step is compared on its own: the oracle's state at the top
# of step n (x_n, f_prev, df_n) is injected into the device path (mgvib200_newton_probe) and that
# step's gradient, mean metric weight, every inner-CG iterate dx_k with its residual norm, every
# f(x_n - dx_k), every line-search evaluation, the accepted step and x_{n+1} are compared.  No code
# path returns without numeric assertions.
# ------------------------------------------------------------------------------------------------
PARITY_LOG = []      # (label, regime, worst relative error) per checked configuration; printed by conftest


def oracle_wbar(om, x, R):
    """mean_i w_i at the '+' points x + r_i (src/mgvi_impl.jl:137-138; field models)."""
    return np.mean([om.metric_weight(x + R[:, i]) for i in range(R.shape[1])], axis=0)


def _scal_err(a, b, scale):
    return abs(a - b) / max(abs(scale), 1e-300)


def _step_errors(got, rec, om, R, check_wbar, k):
    """Per-quantity relative errors of one recorded Newton step (`got`: the device probe or a perturbed
    oracle run, as a dict with the probe's keys) against the oracle record `rec`."""
    e = {}
    e["grad"] = rel(got["grad"], rec["grad"])
    if check_wbar and got.get("wbar") is not None:
        e["wbar"] = rel(got["wbar"], oracle_wbar(om, rec["x_n"], R))
    gnorm = np.linalg.norm(rec["grad"])
    if k:
        dxo = np.stack(rec["dx_iters"], axis=1)
        e["dx_k"] = max(rel(got["dx_iters"][:, j], dxo[:, j]) for j in range(k))
        # residual norms fall by many orders of magnitude: compare on the scale of ||g||
        e["cg_resid"] = max(_scal_err(got["cg_resid"][j], rec["cg_resid"][j], gnorm) for j in range(k))
        if rec["step"] > 1:
            e["f_k"] = max(_scal_err(got["f_k"][j], rec["f_k"][j], rec["f_k"][j]) for j in range(k))
    dnorm = np.linalg.norm(rec["dx"])
    e["dphi0"] = _scal_err(got["dphi0"], rec["dphi0"], gnorm * dnorm)
    # line search: the common prefix of the two evaluation sequences (same kinds) -- abscissae and values
    npre = 0
    for (k1, _, _), (k2, _, _) in zip(got["ls"], rec["ls"]):
        if k1 != k2:
            break
        npre += 1
    pairs = list(zip(got["ls"][:npre], rec["ls"][:npre]))
    e["ls_a"] = max([_scal_err(a1, a2, a2) for (_, a1, _), (_, a2, _) in pairs] or [0.0])
    e["ls_phi"] = max([_scal_err(v1, v2, v2) for (kd, _, v1), (_, _, v2) in pairs if kd == 0] or [0.0])
    # phi'(a) = grad f(x + a d) . d  -- compared on the scale ||grad|| ||d|| (it passes through zero)
    e["ls_dphi"] = max([_scal_err(v1, v2, gnorm * dnorm) for (kd, _, v1), (_, _, v2) in pairs if kd == 1] or [0.0])
    same_ls = npre == len(got["ls"]) == len(rec["ls"])
    if same_ls:
        e["beta"] = _scal_err(got["beta"], rec["beta"], rec["beta"])
        e["f_new"] = _scal_err(got["f_new"], rec["f_new"], rec["f_new"])
        e["x_new"] = rel(got["x_new"], rec["x_new"])
    return e, same_ls


class _Reassociated:
    """The oracle model with the metric product evaluated the way ANOTHER valid implementation of the same
    reference formula v + J'(F . (J v)) would round it: J v and J' l through the materialised Jacobian in a
    permuted row order, and every lambda-space component carrying the rounding of a length-n_lambda reduction
    done in a different order (relative sqrt(n_lambda) eps, random sign) -- the device reduces over the data
    points block-wise where NumPy sums pairwise.  A backward-stable product is only reproducible to
    ~ sqrt(N) eps ||M|| ||u||, which for an ill-scaled metric (polynomial fit: eigenvalues 8 ... 4e7, measured
    normwise difference device vs oracle 3.6e-15) is far more than a one-ulp change of the INPUTS produces:
    the 5th update of a 5 x 5 CG then moves by O(1) (measured: both sides' 5th iterates are 5 % / 35 % off the
    exact solve while iterates 1-4 agree to 1e-17 ... 8e-10)."""

    def __init__(self, om, seed):
        self._om, self._seed, self._calls = om, seed, 0

    def __getattr__(self, name):
        return getattr(self._om, name)

    def linearize(self, xi):
        lin = self._om.linearize(xi)
        J = lin.dense()
        perm = np.random.default_rng(self._seed).permutation(J.shape[0])
        inv = np.argsort(perm)
        Jp = np.ascontiguousarray(J[perm])
        amp = np.sqrt(J.shape[0]) * np.finfo(float).eps

        def jvp(v):
            self._calls += 1
            noise = np.random.default_rng((self._seed, self._calls)).standard_normal(J.shape[0])
            return (Jp @ v)[inv] * (1.0 + amp * noise)

        return O.mgvi_oracle.Linearization(lin.n_theta, lin.n_lambda, lin.fisher, jvp, lambda l: Jp.T @ l[perm], lin.dense)


def _oracle_step_sensitivity(om, R, rec, opts, nseeds=None):
    """The ORACLE's own reproducibility of one Newton step: the same step from x_n and R perturbed at
    the rounding level (x (1 + eps randn), one ulp), per quantity.  The k-th iterate of a TRUNCATED CG amplifies
    rounding like cond^(k/2) (measured against 50-digit arithmetic in tests/test_oracle_pinning.py: the
    reference's 40-pixel fixture loses 6 digits in 5 updates at a condition number of 6e3, and iterates
    beyond k ~ 25 of a 32 x 32 field move by > 1e-9 under a 1e-16 perturbation); no implementation can
    agree with the reference better than the reference agrees with itself, so this is what the tolerance
    is widened to -- per quantity, measured, and reported."""
    sens, ls_stable = {}, True
    k = rec["k"]
    # the amplification is heavy-tailed (a 5 x 5 polynomial metric of condition 1e6+ moves by 2e-4 or by 3e-1
    # depending on the direction of a one-ulp perturbation): small models take the maximum over more directions
    if nseeds is None:
        nseeds = 12 if rec["x_n"].size <= 64 else (4 if rec["x_n"].size <= 16384 else 2)
    ulp = np.finfo(float).eps
    for seed in range(nseeds):
        rng = np.random.default_rng(1000 + seed)
        xp = rec["x_n"] * (1.0 + ulp * rng.standard_normal(rec["x_n"].shape))
        Rp = R * (1.0 + ulp * rng.standard_normal(R.shape))
        # small models additionally re-associate the metric product (see _Reassociated)
        omp = _Reassociated(om, seed) if (rec["x_n"].size <= 64 and seed % 2 == 1) else om
        pert = O.newtoncg_step_record(omp, Rp, xp, rec["step"], rec["f_prev"], rec["df_n"], opts, force_k=k)
        got = dict(pert)
        got["dx_iters"] = np.stack(pert["dx_iters"], axis=1) if pert["dx_iters"] else np.zeros((xp.size, 0))
        got["wbar"] = None
        e, same = _step_errors(got, rec, om, Rp, False, min(k, len(pert["dx_iters"])))
        ls_stable = ls_stable and same
        for key, v in e.items():
            sens[key] = max(sens.get(key, 0.0), float(v))
    return sens, ls_stable


def _strict_from_oracle_state(pkg, model, om, R, rec, check_wbar, tol):
    """Tier A -- every quantity of one Newton step recomputed by the DEVICE from the ORACLE's own state,
    so that nothing has been amplified yet: asserted at `tol` (1e-10) with no widening.
      * every inner-CG update from the oracle's iterator state (x, r, u, residual, prev) -> dx_j, residual
        (mgvib200_ncg_probe; FIELD_DHT models)
      * f(x_n - dx_j) at the oracle's dx_j                                     (src/newtoncg.jl:133)
      * every line-search evaluation phi(a) = f(x_n + a d), phi'(a) = grad f(x_n + a d) . d at the oracle's
        abscissae and direction                                              (src/newtoncg.jl:64-71,142)
    Returns the worst relative error."""
    worst = 0.0
    k = rec["k"]
    gnorm = np.linalg.norm(rec["grad"])
    if k and model.kind in (pkg.api.MODEL_FIELD_DHT, pkg.api.MODEL_FIELD_HYPER):
        Xo, ro = pkg.ncg_probe(model, R, rec["x_n"], rec["cg_states"][:k])
        for j in range(k):
            e = rel(Xo[:, j], rec["dx_iters"][j])
            er = _scal_err(ro[j], rec["cg_resid"][j], gnorm)
            assert e < tol and er < tol, ("CG update from the oracle state", rec["step"], j + 1, e, er)
            worst = max(worst, e, er)
    obj = pkg.KLObjective(model, R)
    if rec["step"] > 1:
        for j in range(k):
            fk = obj.value(rec["x_n"] - rec["dx_iters"][j])
            e = _scal_err(fk, rec["f_k"][j], rec["f_k"][j])
            assert e < tol, ("f_k at the oracle's dx_k", rec["step"], j + 1, e)
            worst = max(worst, e)
    d = -rec["dx"]
    dnorm = np.linalg.norm(d)
    for kind, a, val in rec["ls"]:
        xp = rec["x_n"] + a * d
        if kind == 0:
            e = _scal_err(obj.value(xp), val, val)
        else:
            _, g = obj(xp)
            # phi' passes through zero: on the scale ||grad f(x_p)|| ||d||
            e = _scal_err(float(g @ d), val, max(np.linalg.norm(g) * dnorm, abs(val)))
        assert e < tol, ("line-search evaluation at the oracle's abscissa", rec["step"], kind, a, e)
        worst = max(worst, e)
    return worst


def assert_newton_teacher_forced(pkg, model, om, R, x0, opts_kw=None, label="", tol=RTOL, check_wbar=True):
    """Runs the oracle's NewtonCG with recording, then checks EVERY Newton step on the device in two tiers.

    Tier A (strict, 1e-10, never widened; `_strict_from_oracle_state`): from the oracle's x_n -- grad f(x_n)
    and w_bar; from the oracle's CG iterator state -- every single CG update (dx_k, residual); at the
    oracle's dx_k / line-search abscissae -- every f_k and every (phi, phi') pair.

    Tier B (the step free-running on the device from the oracle's x_n, mgvib200_newton_probe with the CG
    update count forced to the oracle's): dx_k, residuals, f_k, phi'(0), the line-search sequence, beta,
    f_{n+1}, x_{n+1} at 1e-10 -- and where a quantity misses that, against 100 x the oracle's own
    measured reproducibility of it (`_oracle_step_sensitivity`: the k-th iterate of a truncated CG
    amplifies rounding like cond^(k/2)).  The device's own exit rule must fire at the oracle's k wherever
    the oracle's margin exceeds what the two f_k sequences differ by.

    Returns the per-step tier-B error table; logs the regime."""
    opts_kw = opts_kw or {}
    oopts = O.NewtonCGOpts(**opts_kw)
    recs = []
    O.newtoncg_optimize(om, R, x0, oopts, record=recs)
    table, widened, worst_a = [], [], 0.0
    for rec in recs:
        k = rec["k"]
        pr = pkg.newton_probe(model, R, rec["x_n"], rec["step"], rec["f_prev"], rec["df_n"], force_k=k,
                              max_rec=max(k, 1), optimizer=pkg.NewtonCG(**opts_kw), want_wbar=check_wbar)
        assert pr["k_done"] == k, (label, rec["step"], pr["k_done"], k)
        e, same_ls = _step_errors(pr, rec, om, R, check_wbar, k)
        # ---- tier A: strict --------------------------------------------------------------------
        assert e["grad"] < tol, (label, "grad", rec["step"], e["grad"])
        if "wbar" in e:
            assert e["wbar"] < tol, (label, "wbar", rec["step"], e["wbar"])
        worst_a = max(worst_a, e["grad"], e.get("wbar", 0.0),
                      _strict_from_oracle_state(pkg, model, om, R, rec, check_wbar, tol))
        # ---- the step's own exit rule -------------------------------------------------------------
        if k and rec["step"] > 1:
            # |f_k - f_{k-1}| < alpha df_n compares differences of f: comparable only where the oracle's margin
            # exceeds what the two f_k sequences differ by (and the rounding of f itself)
            df = max(abs(pr["f_k"][j] - rec["f_k"][j]) for j in range(k))
            noise = 4.0 * df + 64 * np.finfo(float).eps * abs(rec["f_prev"])
            if all(abs(mg) > noise for mg in rec["exit_margin"]):
                assert pr["k_own_exit"] == k, (label, rec["step"], pr["k_own_exit"], k)
        elif k:
            i0 = opts_kw.get("i0", 5)
            assert pr["k_own_exit"] == (i0 if k >= i0 else 0), (label, pr["k_own_exit"], k)
        # ---- tier B: free-running within the step ------------------------------------------------------
        bad = [key for key, v in e.items() if not v < tol]
        if bad or not same_ls:
            sens, ls_stable = _oracle_step_sensitivity(om, R, rec, oopts)
            assert same_ls or not ls_stable, (label, "step", rec["step"], "line-search branch sequence differs while "
                                              "the oracle's own is stable", pr["ls"], rec["ls"])
            for key in bad:
                lim = max(tol, 100.0 * sens.get(key, 0.0))
                assert e[key] < lim, (label, "step", rec["step"], key, e[key], "oracle self-sensitivity", sens.get(key), e)
                widened.append("%s@%d %.0e (oracle self %.0e)" % (key, rec["step"], e[key], sens.get(key, 0.0)))
        table.append(dict(step=rec["step"], k=k, **e))
    worst_b = max(max(v for kk, v in t.items() if kk not in ("step", "k")) for t in table) if table else 0.0
    PARITY_LOG.append((label, "teacher-forced: per-update strict %.1e; in-step free-running %.1e%s" % (
        worst_a, worst_b, (" (beyond 1e-10 only within the oracle's own reproducibility: %s)" % "; ".join(widened[:4] + (
            ["..."] if len(widened) > 4 else []))) if widened else ""), worst_a))
    return table


def assert_step_parity(pkg, model, om, cfg, res, center, ref, *, label="", opts_kw=None, residual_tol=RTOL):
    """End-to-end mgvi_step against the oracle's.  The residual samples (which do not pass through the
    optimiser) and the KL value at the device's own result are ALWAYS asserted at 1e-10.  The
    free-running updated mean is asserted at 1e-10 with full trace equality where it agrees
    ("strict"); where it does not, the Newton phase is replayed teacher-forced (every step's
    quantities at 1e-10 from the oracle's x_n) -- there is no path without numeric assertions."""
    tr, otr = res.info["optimizer_output"], ref["trace"]
    Rdev = 0.5 * (res.samples[:, 0::2] - res.samples[:, 1::2])
    Rref = ref["residuals"]
    # `residual_tol` > 1e-10 only for deliberately truncated sampler solves (the tutorial's itmax = 10), where the
    # caller has measured the oracle's own reproducibility of that solve
    assert rel(Rdev, Rref) < residual_tol, (label, "residuals", rel(Rdev, Rref))
    assert np.abs(res.samples[:, 0::2] + res.samples[:, 1::2] - 2 * center[:, None]).max() < 1e-12
    # the returned mnlp IS the KL estimate at the returned centre (oracle evaluation at the device's result)
    f_at = O.mean_neg_log_pstr(om, Rdev, center)
    assert abs(f_at - res.mnlp) <= RTOL * abs(f_at), (label, "mnlp", f_at, res.mnlp)
    assert tr["f_history"][-1] <= tr["f_history"][0] + 1e-9 * abs(tr["f_history"][0]), (label, tr["f_history"])
    # first Newton-phase quantities are free of accumulated drift: f(x_0) at 1e-10
    assert abs(tr["f_history"][0] - otr.f_history[0]) <= residual_tol * abs(otr.f_history[0]), (label, "f0")
    ec, es = rel(center, ref["center"]), rel(res.samples, ref["samples"])
    trace_equal = (tr["cg_updates"] == otr.cg_updates and tr["f_calls"] == otr.f_calls and
                   tr["g_calls"] == otr.g_calls and tr["cg_iterations"] == otr.cg_iterations)
    if ec < RTOL and es < RTOL and trace_equal:
        PARITY_LOG.append((label, "strict", max(ec, es)))
        return "strict"
    assert_newton_teacher_forced(pkg, model, om, Rref, cfg["center"], opts_kw, label=label)
    return "teacher-forced"


def assert_field_ops(out, cfg=None):
    for k in ("metric", "residuals", "kl", "grad", "mean_metric"):
        assert out[k] < RTOL, (k, out)
    assert out["build_exact"], out
    a, b = out["iters"]
    # at reltol 1e-12 the tail of the solve sits on the rounding floor: counts may differ by a few
    assert all(abs(x - y) <= 3 + 0.10 * y for x, y in zip(a, b)), out["iters"]
    if "step_regime" in out:
        assert out["step_regime"] in ("strict", "teacher-forced"), out


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
        key = "dense_step" if dense else "cg_step"
        out[key + "_regime"] = assert_step_parity(pkg, m, om, cfg, res, cn, ref, label="poly %s" % key)
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
        out["dense_step_regime"] = assert_step_parity(pkg, m, om, cfg, res, cn, ref, label="dense %dx%d" % cfg["J"].shape)
        out["dense_step_residuals"] = rel(res.samples[:, 0::2] - res.samples[:, 1::2],
                                          ref["samples"][:, 0::2] - ref["samples"][:, 1::2])
    m.close()
    return out
