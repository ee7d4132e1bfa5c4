"""Independent cross-checks of the oracle's restatement of the UN-VENDORED third-party arithmetic
(Krylov.jl `cg!`, IterativeSolvers `cg_iterator!`, LineSearches
`StrongWolfe`, FFTW DHT) with what this image offers, since the reference itself cannot be run
here (no Julia):

 (i)   every iterate of `krylov_cg` and `itsolvers_cg_iterator` against SciPy's conjugate-gradient
       solver (an independent implementation of the same Hestenes-Stiefel recurrence) for
       k = 1 .. 50 (1e-12 for the first ten iterates, 2e-11 after, where two Float64 codes have
       drifted apart by rounding);
 (ii)  one complete `mgvi_step` on a 40-pixel field redone in 50-digit arithmetic (mpmath) by a
       second, independently written restatement (tests/mp_reference.py): bounds the oracle's own
       rounding and shows how far Float64 itself can resolve the updated mean;
 (iii) fixtures written by REAL MGVI.jl v0.4.5 (oracle/make_ref_fixtures.jl, draws as inputs) are
       consumed when present under tests/golden/julia_ref/ -- absent in this image, so that test
       reports "skipped: parity unpinned by the reference" instead of passing silently.
"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import oracle as O
from helpers import TIGHT_O, oracle_model, rel

HERE = os.path.dirname(os.path.abspath(__file__))


def _spd_operator(pkg, dims=(16, 16), like="LIKE_POISSON", std=0.5):
    syn = pkg.synthetic
    cfg = syn.field_config(dims, getattr(syn, like), 2, field_std=std)
    om = oracle_model(cfg)
    lin = om.linearize(cfg["center"])
    b = O.sampler_rhs(lin, cfg["Z1"][:, 0], cfg["Z2"][:, 0])
    return (lambda v: O.metric_apply(lin, v)), b, om.n_theta


def _scipy_cg_iterates(matvec, b, n, kmax):
    its = []
    A = spla.LinearOperator((n, n), matvec=matvec, dtype=np.float64)
    try:
        spla.cg(A, b, x0=np.zeros(n), rtol=0.0, atol=0.0, maxiter=kmax, callback=lambda xk: its.append(xk.copy()))
    except TypeError:      # older SciPy spelling
        spla.cg(A, b, x0=np.zeros(n), tol=0.0, atol=0.0, maxiter=kmax, callback=lambda xk: its.append(xk.copy()))
    return its


@pytest.mark.parametrize("dims,like,kmax", [((16, 16), "LIKE_POISSON", 50), ((32, 32), "LIKE_NORMAL", 50),
                                            ((64,), "LIKE_NORMAL", 25), ((8, 8, 4), "LIKE_POISSON", 40)])
def test_krylov_cg_iterates_match_scipy(pkg, dims, like, kmax):
    """Krylov.jl `cg!` restatement (call site src/residual_samplers.jl:79-80): iterate k for every
    k = 1..kmax (50 where the solve has not hit the rounding floor by then) equals SciPy's k-th CG
    iterate at 1e-12."""
    matvec, b, n = _spd_operator(pkg, dims, like)
    ref = _scipy_cg_iterates(matvec, b, n, kmax)
    assert len(ref) == kmax
    for k in range(1, kmax + 1):
        x, it, _ = O.krylov_cg(matvec, b, atol=0.0, rtol=0.0, itmax=k)
        assert it == k
        # two Float64 implementations of the same recurrence drift apart by rounding as k grows
        assert rel(x, ref[k - 1]) < (1e-12 if k <= 10 else 2e-11), (k, rel(x, ref[k - 1]))


@pytest.mark.parametrize("dims,like,kmax", [((16, 16), "LIKE_POISSON", 50), ((32, 32), "LIKE_NORMAL", 50),
                                            ((64,), "LIKE_NORMAL", 25)])
def test_itsolvers_iterator_iterates_match_scipy(pkg, dims, like, kmax):
    """IterativeSolvers `cg_iterator!` restatement (call site src/newtoncg.jl:120): the iterate after
    every update equals SciPy's, and the reported residual is ||b - A x_k||."""
    matvec, b, n = _spd_operator(pkg, dims, like)
    ref = _scipy_cg_iterates(matvec, b, n, kmax)
    got = []
    for k, x, res in O.itsolvers_cg_iterator(matvec, b, reltol=0.0, abstol=0.0, maxiter=kmax):
        got.append((x.copy(), res))
    assert len(got) == kmax
    bn = np.linalg.norm(b)
    # CG iterates are NOT stable under rounding once Ritz values have converged (Lanczos loss of
    # orthogonality): the same code on b (1 + 1e-16 randn) drifts away exponentially in k.  norm(r)^2
    # instead of dot(r, r) (what IterativeSolvers does) is such a rounding-level difference from SciPy.
    # So: 1e-12 while nothing has been amplified, afterwards 100 x the restatement's own measured drift --
    # an ALGORITHMIC difference (beta, alpha, update order) would show up at O(1e-2) from k = 2 on.
    perts = [[x.copy() for _, x, _ in O.itsolvers_cg_iterator(
        matvec, b * (1.0 + 1e-16 * np.random.default_rng(sd).standard_normal(n)), reltol=0.0, abstol=0.0, maxiter=kmax)]
        for sd in range(3)]
    compared = 0
    for k in range(kmax):
        self_drift = max(rel(p[k], got[k][0]) for p in perts)
        true_res = np.linalg.norm(b - matvec(got[k][0]))
        assert abs(got[k][1] - true_res) <= 1e-10 * bn, (k + 1, got[k][1], true_res)
        if self_drift > 1e-10:
            break          # from here on no two Float64 runs of the SAME code agree to the parity tolerance
        assert rel(got[k][0], ref[k]) < max(1e-12, 1000.0 * self_drift), (k + 1, rel(got[k][0], ref[k]), self_drift)
        compared += 1
    assert compared >= 8, compared


def test_truncated_cg_iterates_amplify_rounding(pkg):
    """Why the Newton phase (inner CG truncated at 5..50 updates, src/newtoncg.jl:121-139) cannot be
    compared end to end at 1e-10: the k-th CG iterate of the SAME Float64 code moves by orders of
    magnitude more than 1e-10 when b is perturbed by 1e-16 (Lanczos loss of orthogonality), although
    the converged solution does not.  Measured on the mean-metric-like operator of a 32 x 32 field."""
    matvec, b, n = _spd_operator(pkg, (32, 32), "LIKE_NORMAL")
    base = [x.copy() for _, x, _ in O.itsolvers_cg_iterator(matvec, b, reltol=0.0, abstol=0.0, maxiter=50)]
    bp = b * (1.0 + 1e-16 * np.random.default_rng(0).standard_normal(n))
    pert = [x.copy() for _, x, _ in O.itsolvers_cg_iterator(matvec, bp, reltol=0.0, abstol=0.0, maxiter=50)]
    drift = [rel(p, q) for p, q in zip(pert, base)]
    assert drift[4] < 1e-13 and max(drift[25:]) > 1e-9, drift
    xs, _, _ = O.krylov_cg(matvec, b, atol=0.0, rtol=1e-13)
    xp, _, _ = O.krylov_cg(matvec, bp, atol=0.0, rtol=1e-13)
    assert rel(xp, xs) < 1e-11             # the CONVERGED solve is well conditioned


def test_krylov_stopping_rule_matches_scipy_semantics(pkg):
    """Stop on ||r|| <= atol + rtol ||b|| (Krylov.jl) -- SciPy stops on ||r|| <= max(rtol ||b||, atol): with
    atol = 0 both count the same number of iterations and return the same solution."""
    matvec, b, n = _spd_operator(pkg, (16, 16), "LIKE_NORMAL")
    A = spla.LinearOperator((n, n), matvec=matvec, dtype=np.float64)
    cnt = [0]
    try:
        xs, _ = spla.cg(A, b, rtol=1e-9, atol=0.0, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
    except TypeError:
        xs, _ = spla.cg(A, b, tol=1e-9, atol=0.0, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
    x, it, rn = O.krylov_cg(matvec, b, atol=0.0, rtol=1e-9)
    assert it == cnt[0] and rel(x, xs) < 1e-12
    assert rn <= 1e-9 * np.linalg.norm(b)


def test_strong_wolfe_against_scipy_line_search_conditions():
    """LineSearches `StrongWolfe` restatement: on smooth 1-D problems the returned step satisfies the
    strong Wolfe conditions with c1 = 1e-4, c2 = 0.9 (what SciPy's independent `line_search` is built
    to deliver as well), a = 1 is accepted at once when it qualifies (1 phi + 1 phi' evaluation), and
    the bracketing doubles the step when the slope is still steep."""
    import scipy.optimize
    c1, c2 = 1e-4, 0.9
    cases = [
        (lambda a: (a - 0.3) ** 2, lambda a: 2 * (a - 0.3)),               # minimiser inside (0, 1): zoom
        (lambda a: (a - 1.0) ** 2 + 2.0, lambda a: 2 * (a - 1.0)),        # a = 1 exact
        (lambda a: (a - 7.0) ** 2, lambda a: 2 * (a - 7.0)),              # far away: doubling
        (lambda a: np.cosh(a - 2.5), lambda a: np.sinh(a - 2.5)),
        (lambda a: -np.log1p(a) + 0.2 * a * a, lambda a: -1 / (1 + a) + 0.4 * a),
    ]
    for phi, dphi in cases:
        calls = []
        a, fa = O.strong_wolfe(lambda t: calls.append(("f", t)) or phi(t), lambda t: calls.append(("g", t)) or dphi(t),
                               1.0, phi(0.0), dphi(0.0))
        assert fa == phi(a)
        assert phi(a) <= phi(0.0) + c1 * a * dphi(0.0)
        assert abs(dphi(a)) <= -c2 * dphi(0.0)
        # SciPy's strong-Wolfe search (independent implementation) also accepts a step with these constants
        a_sp = scipy.optimize.line_search(lambda x: phi(x[0]), lambda x: np.array([dphi(x[0])]), np.array([0.0]),
                                          np.array([1.0]), c1=c1, c2=c2)[0]
        assert a_sp is not None and abs(dphi(a_sp)) <= -c2 * dphi(0.0)
    # first-trial acceptance costs exactly one value and one slope
    calls = []
    O.strong_wolfe(lambda t: calls.append("f") or (t - 1.0) ** 2, lambda t: calls.append("g") or 2 * (t - 1.0), 1.0, 1.0, -2.0)
    assert calls == ["f", "g"]


# ---- (ii) 50-digit arithmetic -----------------------------------------------------------------
def test_mgvi_step_in_50_digit_arithmetic():
    """One mgvi_step on a 40-pixel field (the reference fixture's size, a well-conditioned amplitude)
    recomputed with 50 significant digits: the Float64 oracle agrees with it to rounding on every
    well-conditioned row (metric product, CG residual samples, KL value and gradient) and on the
    updated mean, which bounds the oracle's own rounding well inside the 1e-10 parity tolerance."""
    import mp_reference as MP
    cfg = MP.small_config()
    om = oracle_model(cfg)
    ref = MP.mgvi_step(cfg, digits=50)
    lin = om.linearize(cfg["center"])
    assert rel(O.metric_apply(lin, cfg["V"]), ref["MV"]) < 1e-13
    R, iters, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    assert rel(R, ref["R"]) < 1e-11
    f = O.mean_neg_log_pstr(om, ref["R"], cfg["center"])
    assert abs(f - ref["f0"]) <= 1e-13 * abs(ref["f0"])
    assert rel(O.kl_gradient(om, ref["R"], cfg["center"]), ref["grad0"]) < 1e-13
    step = O.mgvi_step(om, cfg["center"], Z1=cfg["Z1"], Z2=cfg["Z2"], **TIGHT_O)
    assert step["trace"].cg_updates == ref["cg_updates"], (step["trace"].cg_updates, ref["cg_updates"])
    np.testing.assert_allclose(step["trace"].step_lengths, ref["step_lengths"], rtol=1e-8)
    assert rel(step["center"], ref["center"]) < 1e-10, rel(step["center"], ref["center"])
    assert abs(step["mnlp"] - ref["mnlp"]) <= 1e-12 * abs(ref["mnlp"])


def test_reference_fixture_truncated_cg_loses_digits_in_float64():
    """The reference's own 40-pixel model (test/test_models/model_fft_gp.jl: A = 50/(k^2.3+1), exp link)
    against 50-digit arithmetic: the gradient of the Float64 oracle is exact to rounding, but the
    FIVE-update truncated CG solve of the first Newton step (src/newtoncg.jl:121-127) already carries a
    relative error of ~3e-10 -- the mean metric's condition number is only ~6e3, yet the k-th CG
    iterate amplifies rounding like cond^(k/2).  No Float64 implementation (the Julia reference
    included) pins that iterate to 1e-10, which is why the Newton phase is compared step by step
    (teacher-forced) and, where a quantity misses 1e-10, against the oracle's own measured
    reproducibility of that quantity (helpers.assert_newton_teacher_forced)."""
    import mp_reference as MP
    from golden_io import load_golden
    cfg, _ = load_golden("fft_gp_40")
    om = oracle_model(cfg)
    R, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], **TIGHT_O)
    recs = []
    O.newtoncg_optimize(om, R, cfg["center"], O.NewtonCGOpts(steps=1), record=recs)
    exact = MP.newton_first_step(cfg, R, digits=50)
    e_grad = rel(recs[0]["grad"], exact["grad"])
    e_dx = rel(recs[0]["dx_iters"][-1], exact["dx"])
    cond = MP.mean_metric_condition(cfg, R)
    assert e_grad < 1e-13                                  # the gradient itself is exact to rounding
    assert 1e3 < cond < 1e5                                # a moderately conditioned mean metric ...
    assert 1e-12 < e_dx < 1e-16 * cond ** 2.5 * 100        # ... whose 5th CG iterate has lost ~6 digits


# ---- (iii) fixtures from real MGVI.jl ------------------------------------------------------------
def _julia_cases():
    return sorted(glob.glob(os.path.join(HERE, "golden", "julia_ref", "*", "manifest.json")))


def test_reference_julia_fixtures_when_present():
    """Fixtures written by oracle/make_ref_fixtures.jl run against real MGVI.jl v0.4.5 (same draws fed
    through a replaying RNG).  None are committed -- Julia exists neither in this image nor on the GPU
    box -- so here the test SKIPS, and the skip reason is the parity status."""
    cases = _julia_cases()
    if not cases:
        pytest.skip("no tests/golden/julia_ref/ fixtures: parity of rows a5-a12 is UNPINNED by the reference "
                    "(run oracle/make_ref_fixtures.jl with Julia + MGVI.jl v0.4.5 to pin it)")
    from julia_ref_io import load_case
    for manifest in cases:
        cfg, out = load_case(os.path.dirname(manifest))
        om = oracle_model(cfg)
        lin = om.linearize(cfg["center"])
        assert rel(O.metric_apply(lin, out["V"]), out["MV"]) < 1e-10
        R, _, _ = O.sample_residuals_cg(om, cfg["center"], cfg["Z1"], cfg["Z2"], atol=out["cg_atol"], rtol=out["cg_rtol"])
        assert rel(R, out["R"]) < 1e-8
        assert abs(O.mean_neg_log_pstr(om, out["R"], cfg["center"]) - out["f0"]) <= 1e-10 * abs(out["f0"])
        assert rel(O.kl_gradient(om, out["R"], cfg["center"]), out["grad0"]) < 1e-10


# ------------------------------------------------------------------------------------------------
# Closed forms the oracle restates from un-vendored packages, against independent implementations
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["field_poisson", "field_normal", "poly", "hyper"])
def test_loglike_matches_scipy_stats(pkg, kind):
    """Distributions.jl `logpdf` of Normal / Poisson / diagonal MvNormal (src/mgvi_impl.jl:10) as restated in
    the oracle, against scipy.stats on the same rates -- the normalisation constants included (they enter the
    KL value the Newton exit test `|f_k - f_(k-1)| < alpha df_n` looks at, src/newtoncg.jl:134)."""
    import scipy.stats as st
    syn = pkg.synthetic
    if kind == "field_poisson":
        cfg = syn.field_config((12, 10), syn.LIKE_POISSON, 2, field_std=0.5)
    elif kind == "field_normal":
        cfg = syn.field_config((40,), syn.LIKE_NORMAL, 2, field_std=0.5, layout=syn.LAYOUT_INTERLEAVED)
    elif kind == "poly":
        cfg = syn.polyfit_config(2, reference_grid=True)
    else:
        counts = np.random.default_rng(11).poisson(3.0, size=12)
        cfg = syn.hyper_field_config(counts, 2, grain=3, padding=4.0, xlim=(0.0, 10.0), start_scale=0.3)
    om = oracle_model(cfg)
    x = cfg["center"] + 0.1 * np.random.default_rng(3).standard_normal(cfg["center"].shape)
    got = om.loglike(x)
    if kind == "field_poisson":
        want = st.poisson.logpmf(cfg["data"].astype(np.int64), om.rate(x)).sum()
    elif kind == "field_normal":
        want = st.norm.logpdf(cfg["data"], loc=om.rate(x), scale=cfg["sigma"]).sum()
    elif kind == "poly":
        want = st.norm.logpdf(cfg["data"], loc=om.mean(x), scale=om.sigma(x)).sum()
    else:
        want = st.poisson.logpmf(np.asarray(cfg["data"]).astype(np.int64), om.lam(x)).sum()
    assert abs(got - want) <= 1e-12 * abs(want), (kind, got, want)


@pytest.mark.parametrize("dims,like", [((12,), "LIKE_POISSON"), ((6, 5), "LIKE_NORMAL"), ((4, 3, 5), "LIKE_POISSON")])
def test_kl_gradient_and_metric_against_finite_differences(pkg, dims, like):
    """What Zygote / ForwardDiff compute for the reference (src/newtoncg.jl:100, src/residual_samplers.jl:6) is
    the exact derivative; the oracle restates it analytically.  Checked here against central finite differences
    of the oracle's own KL value (gradient) and against J'FJ + I assembled from a finite-difference Jacobian of
    flat_params o model (metric) -- an implementation that shares no derivative code with the restatement."""
    syn = pkg.synthetic
    cfg = syn.field_config(dims, getattr(syn, like), 3, field_std=0.5)
    om = oracle_model(cfg)
    c = cfg["center"]
    R = 0.3 * np.random.default_rng(8).standard_normal((om.n_theta, 3))
    g = O.kl_gradient(om, R, c)
    h = 1e-5
    fd = np.empty_like(g)
    for i in range(om.n_theta):
        e = np.zeros(om.n_theta)
        e[i] = h
        fd[i] = (O.mean_neg_log_pstr(om, R, c + e) - O.mean_neg_log_pstr(om, R, c - e)) / (2 * h)
    assert rel(g, fd) < 1e-7
    # metric: J by central differences of the rate (the mu-part of lambda; the sigma-part is constant), F from the oracle
    lin = om.linearize(c)
    N = om.n_theta
    J = np.empty((N, N))
    for i in range(N):
        e = np.zeros(N)
        e[i] = h
        J[:, i] = (om.rate(c + e) - om.rate(c - e)) / (2 * h)
    lam = om.rate(c)
    F = 1.0 / lam if like == "LIKE_POISSON" else np.full(N, 1.0 / cfg["sigma"] ** 2)
    M = J.T @ (F[:, None] * J) + np.eye(N)
    V = np.random.default_rng(9).standard_normal((N, 2))
    MV = np.stack([O.metric_apply(lin, V[:, k]) for k in range(2)], axis=1)
    assert rel(MV, M @ V) < 1e-7
