"""Host-side mirror of the MGVI.jl v0.4.5 interface for the hot path (same names, argument
meaning and error behaviour), over the C ABI of libmgvib200.so.

reference                                              here
---------                                              ----
MGVIContext (src/mgvi_context.jl:9-15)                 MGVIContext(rng, device)
MGVIConfig (src/mgvi_impl.jl:51-55)                    MGVIConfig(linsolver, optimizer, optimizer_opts)
MGVIResult (src/mgvi_impl.jl:71-78)                    MGVIResult(samples, mnlp, info)
mgvi_step / mgvi_sample (src/mgvi_impl.jl:129-174)     mgvi_step / mgvi_sample
ResidualSampler, sample_residuals                      ResidualSampler, sample_residuals
  (src/residual_samplers.jl:45-123)
MatrixInversion (src/residual_samplers.jl:18)          MatrixInversion
LinearSolve.KrylovJL_CG() (src/mgvi_impl.jl:52)        B200CG(abstol, reltol, maxiters)
NewtonCG (src/newtoncg.jl:14-31)                       NewtonCG(alpha, steps, i0, i1)
fisher_information (src/information.jl)                fisher_information (host, diagonal families)
forward_model::Function                                declarative model structs (CorrelatedFieldModel,
                                                       PolynomialModel, DenseLinearModel): arbitrary
                                                       closures are rejected, there is no CPU fallback.

Older upstream spellings named in BASELINE.json map onto the same objects:
ImplicitResidualSampler, FullResidualSampler, FwdRevADJacobianFunc, FullJacobianFunc,
mgvi_kl_optimize_step.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import capi
from .capi import MGVIB200Error, as_f64, ptr

SQRT_EPS = 1.4901161193847656e-08

MODEL_FIELD_DHT, MODEL_POLY, MODEL_DENSE_LINEAR, MODEL_FIELD_HYPER = 0, 1, 2, 3
LINK_IDENTITY, LINK_EXP = 0, 1
LIKE_NORMAL, LIKE_POISSON = 0, 1
LAYOUT_MU_ONLY, LAYOUT_BLOCKED, LAYOUT_INTERLEAVED = 0, 1, 2


# ----------------------------------------------------------------------------------------------
# context / configuration carriers
# ----------------------------------------------------------------------------------------------
class MGVIContext:
    """`MGVIContext(gen, ad)` (src/mgvi_context.jl:9-15).  `gen` carries the RNG and the compute
    unit in the reference; here: a NumPy Generator (draw source, src/residual_samplers.jl:75-76)
    and the CUDA device.  `ad` has no role: derivatives of the declarative models are analytic."""

    def __init__(self, rng: np.random.Generator | int | None = None, device: int = 0, ad: Any = None,
                 library: capi.Library | None = None):
        self.lib = library or capi.default_library()
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        self.device = int(device)
        self.ad = ad
        h = C.c_void_p()
        rc = self.lib.init(self.device, C.byref(h))
        if rc != 0:
            raise MGVIB200Error("mgvib200_init failed (%d): %s" % (rc, self.lib.last_error(None).decode()))
        self.handle = h
        self.rank, self.world_size = 0, 1

    def check(self, rc: int):
        if rc != 0:
            raise MGVIB200Error(self.lib.last_error(self.handle).decode())

    def init_comm(self, rank: int, world_size: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self.check(self.lib.comm_init(self.handle, rank, world_size, buf))
        self.rank, self.world_size = rank, world_size

    def init_comm_external(self, rank: int, world_size: int, allgather):
        """Host-supplied collective (mgvib200_comm_init_external): `allgather(send_ptr, recv_ptr, count)`
        gathers `count` doubles from every rank into recv (rank-major); pointers are integers
        (memory of the context's device)."""
        def _cb(_user, send, recv, count):
            try:
                allgather(send, recv, int(count))
                return 0
            except Exception:       # noqa: BLE001 -- must not propagate through the C frame
                return 1
        self._allgather_cb = capi.ALLGATHER_FN(_cb)      # keep the trampoline alive
        self.check(self.lib.comm_init_external(self.handle, rank, world_size,
                                               C.cast(self._allgather_cb, C.c_void_p), None))
        self.rank, self.world_size = rank, world_size

    def unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.lib.comm_unique_id(buf)
        if rc != 0:
            raise MGVIB200Error("ncclGetUniqueId failed: %s" % self.lib.last_error(None).decode())
        return bytes(buf)

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.launch_count(self.handle, 1 if reset else 0))

    def last_timing(self):
        a, b, k = C.c_double(), C.c_double(), C.c_int64()
        self.lib.last_timing(self.handle, C.byref(a), C.byref(b), C.byref(k))
        return dict(sampler_ms=a.value, newton_ms=b.value, lockstep_iters=k.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class B200CG:
    """The CG linear solver (stands where `LinearSolve.KrylovJL_CG()` stands, src/mgvi_impl.jl:52).
    Defaults are LinearSolve's: abstol = reltol = sqrt(eps(Float64)), maxiters = n_theta."""
    abstol: float = SQRT_EPS
    reltol: float = SQRT_EPS
    maxiters: int = 0


KrylovJL_CG = B200CG


@dataclass
class MatrixInversion:
    """`MGVI.MatrixInversion` (src/residual_samplers.jl:18): instantiates the metric explicitly."""


B200Dense = MatrixInversion


@dataclass
class NewtonCG:
    """`MGVI.NewtonCG` (src/newtoncg.jl:14-31); the line searcher is StrongWolfe{Float64}()."""
    alpha: float = 0.1
    steps: int = 4
    i0: int = 5
    i1: int = 50

    # reference field names
    @property
    def α(self):  # noqa: PLC2401
        return self.alpha


B200NewtonCG = NewtonCG


@dataclass
class LBFGS:
    """Stands where `Optim.LBFGS()` / `OptimizationLBFGSB.LBFGSB()` stand in `MGVIConfig.optimizer`
    (ext/MGVIOptimExt.jl:12-23, ext/MGVIOptimizationBaseExt.jl:19-36; exercised by
    test/test_mgvi_impl.jl:44-60): a host-side quasi-Newton driver (SciPy L-BFGS-B) over the DEVICE
    objective -- every f / grad f evaluation is `mgvib200_objective_eval` with the residual samples
    resident in HBM.  `MGVIConfig.optimizer_opts` is forwarded as the solver options
    (`Optim.Options(; optimization_opts...)`): maxiter, ftol, gtol, maxcor, ..."""
    method: str = "L-BFGS-B"

    def minimize(self, fun_and_grad, x0, opts):
        import scipy.optimize
        res = scipy.optimize.minimize(fun_and_grad, x0, jac=True, method=self.method, options=dict(opts or {}))
        return np.asarray(res.x, dtype=np.float64), float(res.fun), res


LBFGSB = LBFGS


class KLObjective:
    """`OnceDifferentiable(f, x0)` of the reference's Optim extension: f = _mean_neg_log_pstr over the
    bound residual samples (src/mgvi_impl.jl:19-28), gradient by the analytic adjoint; evaluated on
    the device."""

    def __init__(self, forward_model: "_Model", R):
        m = forward_model
        self.model, self.ctx = m, m.context
        R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
        self.ctx.check(self.ctx.lib.objective_bind(m.handle, ptr(R), R.shape[1]))
        self.f_calls = 0

    def __call__(self, x):
        m = self.model
        x = as_f64(x, (m.n_theta,))
        f = C.c_double()
        g = np.empty(m.n_theta)
        self.ctx.check(self.ctx.lib.objective_eval(m.handle, ptr(x), C.byref(f), ptr(g)))
        self.f_calls += 1
        return float(f.value), g

    def value(self, x):
        m = self.model
        x = as_f64(x, (m.n_theta,))
        f = C.c_double()
        self.ctx.check(self.ctx.lib.objective_eval(m.handle, ptr(x), C.byref(f), None))
        return float(f.value)


@dataclass
class MGVIConfig:
    """`MGVIConfig(linsolver, optimizer, optimizer_opts)` (src/mgvi_impl.jl:51-55)."""
    linsolver: Any = field(default_factory=B200CG)
    optimizer: Any = field(default_factory=NewtonCG)
    optimizer_opts: dict = field(default_factory=dict)


@dataclass
class MGVIResult:
    """`MGVIResult(samples, mnlp, info)` (src/mgvi_impl.jl:71-78).  info.linsolver_output is
    `nothing` in the reference; here it carries the per-sample CG iteration counts."""
    samples: np.ndarray
    mnlp: float
    info: dict


# ----------------------------------------------------------------------------------------------
# declarative forward models
# ----------------------------------------------------------------------------------------------
class _Model:
    kind = -1

    def __init__(self, context: MGVIContext):
        self.context = context
        self.handle = None
        self._keep = []
        self.data = None
        # which ResidualSampler's linearisation currently sits in the device handle (the handle holds ONE
        # linearisation: w, q and the centre); None after mgvi_step / mgvi_sample re-linearised it
        self._lin_owner = None

    def _check_data(self, data):
        """`data` is part of the reference signature (src/mgvi_impl.jl:129-132); the declarative model
        owns its observations, so a different array here is an error, not silently ignored."""
        if data is None:
            return
        d = np.asarray(data, dtype=np.float64).ravel()
        if d.shape != self.data.shape or not np.array_equal(d, self.data):
            raise MGVIB200Error("`data` differs from the observations the model was built with; create a new "
                                "model for new data")

    def _create(self, desc: capi.ModelDesc):
        h = C.c_void_p()
        self.context.check(self.context.lib.model_create(self.context.handle, C.byref(desc), C.byref(h)))
        self.handle = h
        self.n_theta = int(self.context.lib.model_n_theta(h))
        self.n_lambda = int(self.context.lib.model_n_lambda(h))

    def close(self):
        if self.handle and self.context.handle:
            self.context.lib.model_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, xi):
        raise MGVIB200Error("declarative models are evaluated on the device; use mgvi_step / ResidualSampler")


class CorrelatedFieldModel(_Model):
    """s = scale * H(amplitude . xi) + offset, lambda = link(s); Normal(lambda, sigma) or
    Poisson(lambda) data (test/test_models/model_fft_gp.jl:60-72,
    docs/src/advanced_tutorial_lit.jl:253-307).  `dims` in C order."""
    kind = MODEL_FIELD_DHT

    def __init__(self, dims, amplitude, scale, data, context: MGVIContext, *, offset=0.0, link=LINK_EXP,
                 likelihood=LIKE_POISSON, sigma=1.0, layout=LAYOUT_MU_ONLY):
        super().__init__(context)
        dims = tuple(int(d) for d in dims)
        if not 1 <= len(dims) <= 3:
            raise MGVIB200Error("CorrelatedFieldModel: 1 to 3 dimensions")
        N = int(np.prod(dims))
        A = as_f64(amplitude, (N,))
        d = as_f64(data, (N,))
        desc = capi.ModelDesc()
        desc.kind = self.kind
        desc.ndim = len(dims)
        for i, v in enumerate(dims):
            desc.dims[i] = v
        desc.amplitude = ptr(A)
        desc.scale, desc.offset, desc.link = float(scale), float(offset), int(link)
        desc.likelihood, desc.sigma, desc.lambda_layout = int(likelihood), float(sigma), int(layout)
        desc.data, desc.n_data = ptr(d), N
        self._create(desc)
        self.dims = dims
        self.data = d.ravel().copy()

    @classmethod
    def from_config(cls, cfg: dict, context: MGVIContext):
        return cls(cfg["dims"], cfg["amplitude"], cfg["scale"], cfg["data"], context, offset=cfg["offset"],
                   link=cfg["link"], likelihood=cfg["likelihood"], sigma=cfg["sigma"], layout=cfg["layout"])


class PolynomialModel(_Model):
    """test/test_models/model_polyfit.jl:15-24: mean = sum_k coef_scale[k] p[k] x^k per grid,
    sigma = noise_scale * p[end]^2, NamedTupleDist of one product of Normals per grid."""
    kind = MODEL_POLY

    def __init__(self, x_groups, data, context: MGVIContext, *, coef_scale=(10.0, 40.0, 600.0, 80.0),
                 noise_scale=60.0, layout=LAYOUT_BLOCKED):
        super().__init__(context)
        xs = [as_f64(x).ravel() for x in x_groups]
        x = np.concatenate(xs)
        sizes = np.array([g.size for g in xs], dtype=np.int64)
        d = as_f64(data, (x.size,))
        cs = as_f64(coef_scale)
        desc = capi.ModelDesc()
        desc.kind = self.kind
        desc.likelihood, desc.lambda_layout = LIKE_NORMAL, int(layout)
        desc.data, desc.n_data = ptr(d), x.size
        desc.n_groups = len(xs)
        desc.group_sizes = sizes.ctypes.data_as(capi.c_int64_p)
        desc.x = ptr(x)
        desc.n_coef = cs.size
        desc.coef_scale = ptr(cs)
        desc.noise_scale = float(noise_scale)
        self._create(desc)
        self.data = d.ravel().copy()

    @classmethod
    def from_config(cls, cfg: dict, context: MGVIContext):
        return cls(cfg["x_groups"], cfg["data"], context, coef_scale=cfg["coef_scale"],
                   noise_scale=cfg["noise_scale"], layout=cfg["layout"])


class DenseLinearModel(_Model):
    """mu = J xi with a dense (m x k) Jacobian, Normal(mu, sigma): the `DenseMatrix` operator type
    `_get_operator_type(::MatrixInversion)` selects (src/residual_samplers.jl:56)."""
    kind = MODEL_DENSE_LINEAR

    def __init__(self, J, data, context: MGVIContext, *, sigma=0.5, layout=LAYOUT_MU_ONLY):
        """`sigma`: one noise level, or one per datum (m values) -- then F = diag(1/sigma_i^2) is folded
        into the operand load of the J'FJ tensor-core GEMM."""
        super().__init__(context)
        J = as_f64(J)
        m, k = J.shape
        d = as_f64(data, (m,))
        desc = capi.ModelDesc()
        desc.kind = self.kind
        sv = None
        if np.ndim(sigma) > 0:
            sv = as_f64(sigma, (m,))
            desc.sigma_vec = ptr(sv)
            sigma = float(sv[0])
        desc.likelihood, desc.sigma, desc.lambda_layout = LIKE_NORMAL, float(sigma), int(layout)
        desc.data, desc.n_data = ptr(d), m
        desc.J, desc.m, desc.k = ptr(J), m, k
        self._create(desc)
        self.data = d.ravel().copy()

    @classmethod
    def from_config(cls, cfg: dict, context: MGVIContext):
        return cls(cfg["J"], cfg["data"], context, sigma=cfg["sigma"], layout=cfg["layout"])


class HyperFieldModel(_Model):
    """The reference tutorial's model (docs/src/advanced_tutorial_lit.jl:194-307): a 1-D Gaussian process
    whose amplitude spectrum depends on two hyperparameters (`sqrt_kernel` :198-203), exponentiated
    (`poisson_gp_link` :275-277), integrated over data bins of `grain` fine pixels (`agg_lambdas`
    :286-294) and observed through Poisson counts (`model` :301-307).

        theta = [p1, p2, xi(N)];  A[k] = a0 e^{a1 p1} sqrt(2 pi cl) exp(-pi^2 kdist[k]^2 cl^2),  cl = l0 e^{l1 p2}
        s = scale * H(A . xi);  Lambda_b = binsize * sum_{t < grain} exp(s)[bin_start + grain b + t]

    `data_idxs`: 0-based first fine pixel of every data bin; must be regular (start + grain * b)."""
    kind = MODEL_FIELD_HYPER

    def __init__(self, kdist, scale, data_idxs, grain, binsize, data, context: MGVIContext, *, a0, a1, l0, l1):
        super().__init__(context)
        kd = as_f64(kdist)
        N = kd.size
        d = as_f64(data)
        idx = np.asarray(data_idxs, dtype=np.int64).ravel()
        if idx.size != d.size or idx.size < 1:
            raise MGVIB200Error("HyperFieldModel: one start index per data bin")
        if not np.array_equal(idx, idx[0] + int(grain) * np.arange(idx.size)):
            raise MGVIB200Error("HyperFieldModel: data bins must be regular (start + grain * b), as produce_bins makes them")
        desc = capi.ModelDesc()
        desc.kind = self.kind
        desc.ndim = 1
        desc.dims[0] = N
        desc.scale, desc.link, desc.likelihood = float(scale), LINK_EXP, LIKE_POISSON
        desc.lambda_layout = LAYOUT_MU_ONLY
        desc.data, desc.n_data = ptr(d), d.size
        desc.kdist = ptr(kd)
        for i, v in enumerate((a0, a1, l0, l1)):
            desc.hyper[i] = float(v)
        desc.bin_start, desc.grain, desc.binsize = int(idx[0]), int(grain), float(binsize)
        self._create(desc)
        self.dims = (N,)
        self.data = d.ravel().copy()

    @classmethod
    def from_config(cls, cfg: dict, context: MGVIContext):
        return cls(cfg["kdist"], cfg["scale"], cfg["data_idxs"], cfg["grain"], cfg["binsize"], cfg["data"], context,
                   a0=cfg["a0"], a1=cfg["a1"], l0=cfg["l0"], l1=cfg["l1"])


def model_from_config(cfg: dict, context: MGVIContext) -> _Model:
    return {"field": CorrelatedFieldModel, "poly": PolynomialModel, "dense": DenseLinearModel,
            "hyper": HyperFieldModel}[cfg["kind"]].from_config(cfg, context)


# ----------------------------------------------------------------------------------------------
# Fisher information of the distribution families (src/information.jl) -- host-side, diagonal
# ----------------------------------------------------------------------------------------------
@dataclass
class Normal:
    mu: float
    sigma: float


@dataclass
class Poisson:
    lam: float


@dataclass
class Exponential:
    theta: float


@dataclass
class MvNormalDiag:
    mu: Any
    var: Any


@dataclass
class MvNormalFull:
    """`MvNormal(mu, Sigma)` with a full covariance (src/information.jl:20-58): the one family whose
    Fisher block is dense; flat parameters [mu; columns of the upper triangle of Sigma]."""
    mu: Any
    cov: Any


@dataclass
class Product:
    dists: list


NamedTupleDist = dict


def _fisher_mvnormal_full(cov) -> np.ndarray:
    # src/information.jl:20-58 in closed form: mean block inv(Sigma); covariance block
    # C[(a,b),(c,d)] = P_ac P_bd + P_ad P_bc, halved once for every diagonal parameter (a == b, c == d)
    cov = np.asarray(cov, dtype=np.float64)
    d = cov.shape[0]
    P = np.linalg.inv(cov)
    idx = [(i, j) for j in range(d) for i in range(j + 1)]
    nc = len(idx)
    out = np.zeros((d + nc, d + nc))
    out[:d, :d] = P
    for m, (a, b) in enumerate(idx):
        for n, (c, e) in enumerate(idx):
            v = P[a, c] * P[b, e] + P[a, e] * P[b, c]
            if a == b:
                v *= 0.5
            if c == e:
                v *= 0.5
            out[d + m, d + n] = v
    return out


def fisher_information(dist) -> np.ndarray:
    """`MGVI.fisher_information(distribution)` in the order of `flat_params`
    (src/information.jl:12-96; src/shapes.jl:40-54): a 1-D array (the diagonal) for the diagonal
    families and their products -- what the device path consumes --, a dense 2-D array as soon as a
    full-covariance `MvNormalFull` block is involved (host-side only)."""
    if isinstance(dist, MvNormalFull):
        return _fisher_mvnormal_full(dist.cov)
    if isinstance(dist, (Product, dict)):
        ds = (list(np.asarray(dist.dists, dtype=object).ravel(order="F")) if isinstance(dist, Product)
              else list(dist.values()))
        parts = [fisher_information(d) for d in ds]
        if all(p.ndim == 1 for p in parts):
            return np.concatenate(parts)
        mats = [np.diag(p) if p.ndim == 1 else p for p in parts]
        n = sum(m.shape[0] for m in mats)
        out = np.zeros((n, n))
        o = 0
        for m in mats:
            out[o:o + m.shape[0], o:o + m.shape[0]] = m
            o += m.shape[0]
        return out
    if isinstance(dist, Normal):
        inv = 1.0 / dist.sigma
        return np.array([inv * inv, 2.0 * (inv * inv)])
    if isinstance(dist, Poisson):
        return np.array([1.0 / dist.lam])
    if isinstance(dist, Exponential):
        inv = 1.0 / dist.theta
        return np.array([inv * inv])
    if isinstance(dist, MvNormalDiag):
        inv = 1.0 / np.asarray(dist.var, dtype=np.float64)
        return np.concatenate([inv, 2.0 * inv])
    raise MGVIB200Error("fisher_information: unsupported distribution %r" % (dist,))


# ----------------------------------------------------------------------------------------------
# residual samplers
# ----------------------------------------------------------------------------------------------
def _cg_opts(solver) -> capi.CGOpts:
    return capi.CGOpts(float(solver.abstol), float(solver.reltol), int(solver.maxiters))


def _draws(context: MGVIContext, rows: int, n: int) -> np.ndarray:
    # dense sampler: one (n_theta x n) matrix, column-major (src/residual_samplers.jl:114,121)
    return context.rng.standard_normal((n, rows)).T


def _draw_pairs(context: MGVIContext, n_lambda: int, n_theta: int, n: int):
    """CG sampler draws in the order the reference consumes the stream (src/residual_samplers.jl:75-76,
    once per residual, :88): for residual i first n1_i (n_lambda) then eta_i (n_theta)."""
    Z1 = np.empty((n_lambda, n), order="F")
    Z2 = np.empty((n_theta, n), order="F")
    for i in range(n):
        Z1[:, i] = context.rng.standard_normal(n_lambda)
        Z2[:, i] = context.rng.standard_normal(n_theta)
    return Z1, Z2


class ResidualSampler:
    """`ResidualSampler(f_model, center_point, linear_solver, context)`
    (src/residual_samplers.jl:45-63): linearises the model at the centre on construction."""

    def __init__(self, f_model, center_point, linear_solver, context: MGVIContext):
        if not isinstance(f_model, _Model):
            raise MGVIB200Error("forward_model must be a declarative mgvi_b200 model (CorrelatedFieldModel, "
                                "PolynomialModel, DenseLinearModel); arbitrary functions have no device path")
        self.f_model = f_model
        self.center_point = as_f64(center_point, (f_model.n_theta,))
        self.linear_solver = linear_solver
        self.context = context
        self._linearize()

    def _linearize(self):
        self.context.check(self.context.lib.linearize(self.f_model.handle, ptr(self.center_point)))
        self.f_model._lin_owner = self

    def _ensure(self):
        """The reference's sampler OWNS its linearisation (src/residual_samplers.jl:45-63); here the
        device handle holds one linearisation at a time, so a sampler re-linearises at its own centre
        whenever another sampler, mgvi_step or mgvi_sample has used the model since."""
        if self.f_model._lin_owner is not self:
            self._linearize()

    def metric_apply(self, V):
        """(J' F J + I) V for the columns of V (src/residual_samplers.jl:72)."""
        self._ensure()
        V = as_f64(V).reshape(self.f_model.n_theta, -1, order="F")
        V = np.asfortranarray(V)
        out = np.empty_like(V, order="F")
        self.context.check(self.context.lib.metric_apply(self.f_model.handle, ptr(V), V.shape[1], ptr(out)))
        return out

    def weights(self):
        self._ensure()
        N = self.f_model.n_lambda if self.f_model.kind != MODEL_FIELD_DHT else self.f_model.n_theta
        w, q = np.empty(N), np.empty(N)
        self.context.check(self.context.lib.get_weights(self.f_model.handle, ptr(w), ptr(q)))
        return w, q


def ImplicitResidualSampler(f_model, center_point, context, **kw):
    return ResidualSampler(f_model, center_point, B200CG(**kw), context)


def FullResidualSampler(f_model, center_point, context):
    return ResidualSampler(f_model, center_point, MatrixInversion(), context)


class FwdRevADJacobianFunc:
    """Old-API Jacobian strategy (JVP + VJP operator): selects the implicit CG path."""
    linsolver = B200CG


class FullJacobianFunc:
    """Old-API Jacobian strategy (materialised Jacobian): selects the dense path."""
    linsolver = MatrixInversion


FwdDerJacobianFunc = FwdRevADJacobianFunc


def sample_residuals(s: ResidualSampler, n: int | None = None, draws=None, return_info: bool = False):
    """`MGVI.sample_residuals(s[, n])` (src/residual_samplers.jl:66-123).  `draws` = (Z1, Z2) for
    the CG sampler or Z for MatrixInversion; generated from context.rng when omitted."""
    single = n is None
    n = 1 if single else int(n)
    m, ctx = s.f_model, s.context
    s._ensure()
    R = np.empty((m.n_theta, n), order="F")
    info = {}
    if isinstance(s.linear_solver, MatrixInversion):
        Z = as_f64(draws if draws is not None else _draws(ctx, m.n_theta, n)).reshape(m.n_theta, n, order="F")
        Z = np.asfortranarray(Z)
        ctx.check(ctx.lib.sample_residuals_dense(m.handle, ptr(Z), n, ptr(R), None))
    else:
        if draws is None:
            Z1, Z2 = _draw_pairs(ctx, m.n_lambda, m.n_theta, n)
        else:
            Z1, Z2 = draws
        Z1 = np.asfortranarray(as_f64(Z1).reshape(m.n_lambda, n, order="F"))
        Z2 = np.asfortranarray(as_f64(Z2).reshape(m.n_theta, n, order="F"))
        iters = np.zeros(n, dtype=np.int64)
        rn = np.zeros(n)
        opts = _cg_opts(s.linear_solver)
        ctx.check(ctx.lib.sample_residuals_cg(m.handle, ptr(Z1), ptr(Z2), n, C.byref(opts), ptr(R),
                                              iters.ctypes.data_as(capi.c_int64_p), ptr(rn)))
        info = dict(iters=iters, final_rnorm=rn)
    out = R[:, 0].copy() if single else R
    return (out, info) if return_info else out


# ----------------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------------
def _ncg_opts(opt: NewtonCG) -> capi.NewtonCGOpts:
    return capi.NewtonCGOpts(float(opt.alpha), int(opt.steps), int(opt.i0), int(opt.i1))


def _step_draws(model, n, dense, context, draws):
    if dense:
        Z = draws if draws is not None else _draws(context, model.n_theta, n)
        if isinstance(Z, tuple):
            Z = Z[0]
        Z1 = np.asfortranarray(as_f64(Z).reshape(model.n_theta, n, order="F"))
        return Z1, None
    if draws is None:
        Z1, Z2 = _draw_pairs(context, model.n_lambda, model.n_theta, n)
    else:
        Z1, Z2 = draws
    Z1 = np.asfortranarray(as_f64(Z1).reshape(model.n_lambda, n, order="F"))
    Z2 = np.asfortranarray(as_f64(Z2).reshape(model.n_theta, n, order="F"))
    return Z1, Z2


def mgvi_step(forward_model, data, n_residuals: int, center_init, config: MGVIConfig, context: MGVIContext,
              draws=None):
    """`mgvi_step(forward_model, data, n_residuals, center_init, config, context)`
    (src/mgvi_impl.jl:129-144) -> (MGVIResult, updated_center).  `data` must be the observations the
    declarative model was built with (or None); `config.optimizer_opts` is accepted by NewtonCG in the
    reference but never read there (src/newtoncg.jl:85-167), so anything but an empty set is rejected."""
    if not isinstance(forward_model, _Model):
        raise MGVIB200Error("forward_model must be a declarative mgvi_b200 model; there is no CPU fallback")
    forward_model._check_data(data)
    if not isinstance(config.optimizer, NewtonCG):
        return _mgvi_step_host_optimizer(forward_model, n_residuals, center_init, config, context, draws)
    if config.optimizer_opts:
        raise MGVIB200Error("NewtonCG takes no optimizer_opts (src/newtoncg.jl:85-88 ignores them); got %r"
                            % (dict(config.optimizer_opts),))
    m, n = forward_model, int(n_residuals)
    dense = isinstance(config.linsolver, MatrixInversion)
    center = as_f64(center_init, (m.n_theta,))
    Z1, Z2 = _step_draws(m, n, dense, context, draws)
    samples = np.empty((m.n_theta, 2 * n), order="F")
    center_out = np.empty(m.n_theta)
    mnlp = C.c_double()
    iters = np.zeros(n, dtype=np.int64)
    trace = capi.NewtonCGTrace()
    cg = _cg_opts(config.linsolver) if not dense else capi.CGOpts(SQRT_EPS, SQRT_EPS, 0)
    opts = _ncg_opts(config.optimizer)
    context.check(context.lib.step(m.handle, ptr(center), ptr(Z1), ptr(Z2) if Z2 is not None else None, n,
                                   1 if dense else 0, C.byref(cg), C.byref(opts), ptr(samples), ptr(center_out),
                                   C.byref(mnlp), iters.ctypes.data_as(capi.c_int64_p), C.byref(trace)))
    m._lin_owner = None          # the step linearised the handle at center_init
    info = dict(linsolver_output=dict(iters=iters), optimizer_output=trace.as_dict())
    return MGVIResult(samples, float(mnlp.value), info), center_out


def _mgvi_step_host_optimizer(m, n_residuals, center_init, config, context, draws):
    """mgvi_step with `MGVI._optimize` dispatched to a host-side optimizer (the ext/*.jl methods): the
    sampler and every objective / gradient evaluation run on the device, the optimizer's own vector
    algebra on the host, exactly the split the reference's Optim extension has."""
    opt = config.optimizer
    if not hasattr(opt, "minimize"):
        raise MGVIB200Error("optimizer must be MGVI.NewtonCG or provide minimize(fun_and_grad, x0, opts) "
                            "(e.g. LBFGS()); got %r" % (opt,))
    sampler = ResidualSampler(m, center_init, config.linsolver, context)
    R, info = sample_residuals(sampler, int(n_residuals), draws=draws, return_info=True)
    obj = KLObjective(m, R)
    x, f_x, optres = opt.minimize(obj, sampler.center_point.copy(), config.optimizer_opts)
    f_x = obj.value(x)                       # f(x_res) itself (ext/MGVIOptimizationBaseExt.jl:33)
    samples = build_samples(m, R, x)
    out = dict(linsolver_output=info, optimizer_output=optres)
    return MGVIResult(samples, f_x, out), x


mgvi_kl_optimize_step = mgvi_step


def mgvi_sample(forward_model, data, n_residuals: int, center, config: MGVIConfig, context: MGVIContext,
                draws=None):
    """`mgvi_sample` (src/mgvi_impl.jl:166-174): samples at the given centre, no optimisation."""
    if not isinstance(forward_model, _Model):
        raise MGVIB200Error("forward_model must be a declarative mgvi_b200 model; there is no CPU fallback")
    forward_model._check_data(data)
    m, n = forward_model, int(n_residuals)
    dense = isinstance(config.linsolver, MatrixInversion)
    c = as_f64(center, (m.n_theta,))
    Z1, Z2 = _step_draws(m, n, dense, context, draws)
    samples = np.empty((m.n_theta, 2 * n), order="F")
    iters = np.zeros(n, dtype=np.int64)
    cg = _cg_opts(config.linsolver) if not dense else capi.CGOpts(SQRT_EPS, SQRT_EPS, 0)
    context.check(context.lib.sample(m.handle, ptr(c), ptr(Z1), ptr(Z2) if Z2 is not None else None, n,
                                     1 if dense else 0, C.byref(cg), ptr(samples),
                                     iters.ctypes.data_as(capi.c_int64_p)))
    m._lin_owner = None
    return samples


def mgvi_mvnormal_pushfwd_function(forward_model, data, config: MGVIConfig, center_point, context: MGVIContext):
    """`mgvi_mvnormal_pushfwd_function` (src/mgvi_impl.jl:191-198): a function that pushes standard
    normal vectors forward to the MGVI posterior approximation, `z -> L_Sigma z + center`
    (`MulAdd(residual_pushfwd_operator(sampler), center)`).  As in the reference this instantiates
    the metric densely (MatrixInversion), so it is for low-dimensional models; the product
    `L_Sigma z` runs on the device."""
    forward_model._check_data(data)
    sampler = ResidualSampler(forward_model, center_point, MatrixInversion(), context)
    center = sampler.center_point.copy()

    def pushfwd(z):
        z = as_f64(z)
        single = z.ndim == 1
        Z = np.asfortranarray(z.reshape(forward_model.n_theta, -1, order="F"))
        sampler._ensure()        # re-linearises at `center` if the model has been used elsewhere since
        R = np.empty_like(Z, order="F")
        context.check(context.lib.sample_residuals_dense(forward_model.handle, ptr(Z), Z.shape[1], ptr(R), None))
        out = center[:, None] + R
        return out[:, 0] if single else out

    return pushfwd


def kl_value_grad(forward_model: _Model, x, R, want_grad: bool = True):
    """`_mean_neg_log_pstr` and its gradient (src/mgvi_impl.jl:19-28, src/newtoncg.jl:100)."""
    m, ctx = forward_model, forward_model.context
    x = as_f64(x, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    f = C.c_double()
    g = np.empty(m.n_theta) if want_grad else None
    ctx.check(ctx.lib.kl_value_grad(m.handle, ptr(x), ptr(R), R.shape[1], C.byref(f), ptr(g) if want_grad else None))
    return float(f.value), g


def mean_metric_apply(forward_model: _Model, x, R, u):
    """`mean(Sigma_inv.(x .+ r_i))` applied to u (src/mgvi_impl.jl:137-138)."""
    m, ctx = forward_model, forward_model.context
    x = as_f64(x, (m.n_theta,))
    u = as_f64(u, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    out = np.empty(m.n_theta)
    ctx.check(ctx.lib.mean_metric_apply(m.handle, ptr(x), ptr(R), R.shape[1], ptr(u), ptr(out)))
    return out


def newtoncg_optimize(forward_model: _Model, R, x0, optimizer: NewtonCG = NewtonCG()):
    """`MGVI._optimize(f, adsel, Sigma_bar_inv, x0, ::NewtonCG, opts)` (src/newtoncg.jl:85-167)
    -> (x, f(x), trace dict)."""
    m, ctx = forward_model, forward_model.context
    x0 = as_f64(x0, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    x = np.empty(m.n_theta)
    f = C.c_double()
    trace = capi.NewtonCGTrace()
    opts = _ncg_opts(optimizer)
    ctx.check(ctx.lib.newtoncg(m.handle, ptr(x0), ptr(R), R.shape[1], C.byref(opts), ptr(x), C.byref(f),
                               C.byref(trace)))
    return x, float(f.value), trace.as_dict()


def newton_probe(forward_model: _Model, R, x_n, step: int, f_prev: float, df_n: float, *, force_k: int = 0,
                 max_rec: int = 64, optimizer: NewtonCG = NewtonCG(), want_wbar: bool = True):
    """One Newton step of `MGVI._optimize(::NewtonCG)` (src/newtoncg.jl:112-146) from an injected state,
    with every intermediate returned (mgvib200_newton_probe): parity instrumentation for the
    teacher-forced tests -- the same device code path as `newtoncg_optimize`."""
    m, ctx = forward_model, forward_model.context
    x_n = as_f64(x_n, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    pr = capi.NewtonProbe()
    pr.step, pr.f_prev, pr.df_n, pr.force_k, pr.max_rec = int(step), float(f_prev), float(df_n), int(force_k), int(max_rec)
    grad = np.empty(m.n_theta)
    wbar = np.empty(m.n_theta) if (want_wbar and m.kind == MODEL_FIELD_DHT) else None
    dx = np.full((m.n_theta, max_rec), np.nan, order="F")
    f_k = np.full(max_rec, np.nan)
    resid = np.full(max_rec, np.nan)
    x_new = np.empty(m.n_theta)
    pr.grad, pr.dx_iters, pr.f_k, pr.cg_resid, pr.x_new = ptr(grad), ptr(dx), ptr(f_k), ptr(resid), ptr(x_new)
    pr.wbar = ptr(wbar) if wbar is not None else None
    opts = _ncg_opts(optimizer)
    ctx.check(ctx.lib.newton_probe(m.handle, ptr(x_n), ptr(R), R.shape[1], C.byref(opts), C.byref(pr)))
    k = int(pr.k_done)
    nls = int(pr.n_ls)
    return dict(grad=grad, wbar=wbar, dx_iters=dx[:, :min(k, max_rec)], f_k=f_k[:min(k, max_rec)],
                cg_resid=resid[:min(k, max_rec)], x_new=x_new, k_done=k, k_own_exit=int(pr.k_own_exit),
                dphi0=float(pr.dphi0), beta=float(pr.beta), f_new=float(pr.f_new),
                ls=[(int(pr.ls_kind[i]), float(pr.ls_a[i]), float(pr.ls_val[i])) for i in range(nls)])


def ncg_probe(forward_model: _Model, R, x_n, states):
    """Single updates of the Newton inner CG from injected iterator states (mgvib200_ncg_probe).
    `states`: list of (x, r, u, residual, prev_residual) before an update -> (X_out (n_theta x k), residuals)."""
    m, ctx = forward_model, forward_model.context
    x_n = as_f64(x_n, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    k = len(states)
    X = np.asfortranarray(np.stack([s[0] for s in states], axis=1)) if k else np.zeros((m.n_theta, 0), order="F")
    Rr = np.asfortranarray(np.stack([s[1] for s in states], axis=1)) if k else X
    U = np.asfortranarray(np.stack([s[2] for s in states], axis=1)) if k else X
    res = np.array([s[3] for s in states], dtype=np.float64)
    prev = np.array([s[4] for s in states], dtype=np.float64)
    Xo = np.empty((m.n_theta, k), order="F")
    ro = np.empty(k)
    ctx.check(ctx.lib.ncg_probe(m.handle, ptr(x_n), ptr(R), R.shape[1], k, ptr(X), ptr(Rr), ptr(U), ptr(res), ptr(prev),
                                ptr(Xo), ptr(ro)))
    return Xo, ro


def build_samples(forward_model: _Model, R, center):
    """`_build_samples` (src/mgvi_impl.jl:152-156)."""
    m, ctx = forward_model, forward_model.context
    c = as_f64(center, (m.n_theta,))
    R = np.asfortranarray(as_f64(R).reshape(m.n_theta, -1, order="F"))
    out = np.empty((m.n_theta, 2 * R.shape[1]), order="F")
    ctx.check(ctx.lib.build_samples(m.handle, ptr(c), ptr(R), R.shape[1], ptr(out)))
    return out
