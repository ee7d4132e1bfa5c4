# MGVIB200.jl -- Julia shim that puts libmgvib200.so behind the MGVI.jl v0.4.5 API.
#
# STATUS: written against include/mgvi_b200.h and the reference sources, NOT executed in the build
# environment (no Julia binary in the image or on the GPU box; SURVEY.md section 0.3).  The Python
# mirror (mgvi.jl_b200/api.py) binds exactly the same symbols with ctypes and is what the tests run.
#
# It hooks the reference at its own dispatch seams, without editing MGVI.jl (SURVEY.md 8b):
#   * forward model type     -> declarative callable structs (<: Function)        src/residual_samplers.jl:59
#   * solver type            -> B200CG / MGVI.MatrixInversion select the sampler  src/residual_samplers.jl:56-63,111-123
#   * optimizer type         -> MGVI._optimize(..., ::B200NewtonCG, ...)          src/newtoncg.jl:85-88, ext/MGVIOptimExt.jl:12-15
#   * draw source            -> randn(context.gen, ...) in the reference order    src/residual_samplers.jl:75-76,121
module MGVIB200

using MGVI
using Random
import HeterogeneousComputing: GenContext
import LinearAlgebra

export CorrelatedFieldModel, PolynomialModel, DenseLinearModel, HyperFieldModel
export B200CG, B200Dense, B200NewtonCG
export ImplicitResidualSampler, FullResidualSampler, FwdRevADJacobianFunc, FwdDerJacobianFunc, FullJacobianFunc
export mgvi_kl_optimize_step

const libmgvib200 = get(ENV, "MGVIB200_LIB", "libmgvib200.so")

# ---- C structs (include/mgvi_b200.h) ------------------------------------------------------------
struct ModelDesc
    kind::Int32
    ndim::Int32
    dims::NTuple{3,Int64}
    amplitude::Ptr{Float64}
    scale::Float64
    offset::Float64
    link::Int32
    likelihood::Int32
    sigma::Float64
    lambda_layout::Int32
    data::Ptr{Float64}
    n_data::Int64
    n_groups::Int32
    group_sizes::Ptr{Int64}
    x::Ptr{Float64}
    n_coef::Int32
    coef_scale::Ptr{Float64}
    noise_scale::Float64
    J::Ptr{Float64}
    m::Int64
    k::Int64
    sigma_vec::Ptr{Float64}
    kdist::Ptr{Float64}
    hyper::NTuple{4,Float64}
    bin_start::Int64
    grain::Int32
    binsize::Float64
end
const _NOHYPER = (C_NULL, (0.0, 0.0, 0.0, 0.0), Int64(0), Int32(0), 0.0)      # kdist, hyper, bin_start, grain, binsize

struct CGOpts
    abstol::Float64
    reltol::Float64
    maxiters::Int64
end

struct NewtonCGOpts
    alpha::Float64
    steps::Int64
    i0::Int64
    i1::Int64
end

const TRACE_MAX = 64
mutable struct NewtonCGTrace
    steps::Int64
    f_history::NTuple{TRACE_MAX + 1,Float64}
    n_cg_iterations::Int64
    cg_iterations::NTuple{TRACE_MAX + 1,Int64}
    cg_updates::NTuple{TRACE_MAX,Int64}
    step_lengths::NTuple{TRACE_MAX,Float64}
    f_calls::Int64
    g_calls::Int64
    initial_f::Float64
    minimum::Float64
    NewtonCGTrace() = new()
end

const MODEL_FIELD_DHT, MODEL_POLY, MODEL_DENSE_LINEAR, MODEL_FIELD_HYPER = Int32(0), Int32(1), Int32(2), Int32(3)
const LINK_IDENTITY, LINK_EXP = Int32(0), Int32(1)
const LIKE_NORMAL, LIKE_POISSON = Int32(0), Int32(1)
const LAYOUT_MU_ONLY, LAYOUT_BLOCKED, LAYOUT_INTERLEAVED = Int32(0), Int32(1), Int32(2)

# ---- context: one library context per Julia process/device ---------------------------------------
mutable struct B200Context
    handle::Ptr{Cvoid}
    function B200Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:mgvib200_init, libmgvib200), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error("mgvib200_init failed: ", unsafe_string(ccall((:mgvib200_last_error, libmgvib200), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> ccall((:mgvib200_destroy, libmgvib200), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const _default_ctx = Ref{Union{Nothing,B200Context}}(nothing)
b200_context() = (_default_ctx[] === nothing && (_default_ctx[] = B200Context(0)); _default_ctx[])

function _check(ctx::B200Context, rc::Integer)
    rc == 0 || throw(ErrorException(unsafe_string(ccall((:mgvib200_last_error, libmgvib200), Cstring, (Ptr{Cvoid},), ctx.handle))))
    nothing
end

# ---- declarative forward models (callable, <: Function as src/residual_samplers.jl:59 requires) ----
abstract type B200Model <: Function end

mutable struct ModelHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
end

function _create(ctx::B200Context, desc::ModelDesc, keep...)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        _check(ctx, ccall((:mgvib200_model_create, libmgvib200), Cint, (Ptr{Cvoid}, Ref{ModelDesc}, Ref{Ptr{Cvoid}}), ctx.handle, desc, h))
    end
    mh = ModelHandle(h[], ctx)
    finalizer(m -> ccall((:mgvib200_model_destroy, libmgvib200), Cvoid, (Ptr{Cvoid},), m.ptr), mh)
    return mh
end

"""
    CorrelatedFieldModel(amplitude::Array, scale, data; offset=0, link=:exp, likelihood=:poisson, sigma=1, layout=:mu_only)

`s = scale * DHT(amplitude .* ξ) .+ offset`, `λ = link.(s)`; `Normal.(λ, sigma)` or `Poisson.(λ)` data
(test/test_models/model_fft_gp.jl:60-72, docs/src/advanced_tutorial_lit.jl:253-307).
"""
struct CorrelatedFieldModel{N} <: B200Model
    handle::ModelHandle
    dims::NTuple{N,Int}
end

function CorrelatedFieldModel(amplitude::Array{Float64,N}, scale::Real, data::Array; offset::Real=0.0, link::Symbol=:exp,
                              likelihood::Symbol=:poisson, sigma::Real=1.0, layout::Symbol=:mu_only,
                              ctx::B200Context=b200_context()) where N
    1 <= N <= 3 || throw(ArgumentError("1 to 3 dimensions"))
    d = Float64.(vec(data))
    cdims = reverse(size(amplitude))                      # the C ABI takes dims slowest-first
    dims3 = ntuple(i -> i <= N ? Int64(cdims[i]) : Int64(0), 3)
    lay = Dict(:mu_only => LAYOUT_MU_ONLY, :blocked => LAYOUT_BLOCKED, :interleaved => LAYOUT_INTERLEAVED)[layout]
    desc = ModelDesc(MODEL_FIELD_DHT, Int32(N), dims3, pointer(amplitude), scale, offset,
                     link === :exp ? LINK_EXP : LINK_IDENTITY, likelihood === :poisson ? LIKE_POISSON : LIKE_NORMAL,
                     sigma, lay, pointer(d), length(d), Int32(0), C_NULL, C_NULL, Int32(0), C_NULL, 0.0, C_NULL, 0, 0,
                     C_NULL, _NOHYPER...)
    CorrelatedFieldModel{N}(_create(ctx, desc, amplitude, d), size(amplitude))
end

"""
    PolynomialModel(x_groups, data; coef_scale=[10, 40, 600, 80], noise_scale=60, layout=:blocked)

test/test_models/model_polyfit.jl:15-24.
"""
struct PolynomialModel <: B200Model
    handle::ModelHandle
end

function PolynomialModel(x_groups::Vector{Vector{Float64}}, data::Vector{Float64}; coef_scale=[10.0, 40.0, 600.0, 80.0],
                         noise_scale::Real=60.0, layout::Symbol=:blocked, ctx::B200Context=b200_context())
    x = reduce(vcat, x_groups)
    sizes = Int64.(length.(x_groups))
    cs = Float64.(coef_scale)
    lay = layout === :blocked ? LAYOUT_BLOCKED : LAYOUT_INTERLEAVED
    desc = ModelDesc(MODEL_POLY, Int32(0), (0, 0, 0), C_NULL, 0.0, 0.0, Int32(0), LIKE_NORMAL, 0.0, lay, pointer(data),
                     length(data), Int32(length(sizes)), pointer(sizes), pointer(x), Int32(length(cs)), pointer(cs),
                     noise_scale, C_NULL, 0, 0, C_NULL, _NOHYPER...)
    PolynomialModel(_create(ctx, desc, x, sizes, cs, data))
end

"""
    DenseLinearModel(J::Matrix, data; sigma=0.5)

`μ = J ξ`, `Normal.(μ, sigma)`: the `DenseMatrix` operator type of `MatrixInversion` (src/residual_samplers.jl:56).
"""
struct DenseLinearModel <: B200Model
    handle::ModelHandle
end

function DenseLinearModel(J::Matrix{Float64}, data::Vector{Float64}; sigma::Union{Real,Vector{Float64}}=0.5,
                          ctx::B200Context=b200_context())
    # a vector `sigma` (one noise level per datum) is folded into the operand load of the J'FJ tensor-core GEMM
    sv = sigma isa Vector ? sigma : Float64[]
    desc = ModelDesc(MODEL_DENSE_LINEAR, Int32(0), (0, 0, 0), C_NULL, 0.0, 0.0, Int32(0), LIKE_NORMAL,
                     sigma isa Vector ? sigma[1] : Float64(sigma), LAYOUT_MU_ONLY,
                     pointer(data), length(data), Int32(0), C_NULL, C_NULL, Int32(0), C_NULL, 0.0, pointer(J),
                     size(J, 1), size(J, 2), isempty(sv) ? C_NULL : pointer(sv), _NOHYPER...)
    DenseLinearModel(_create(ctx, desc, J, data, sv))
end

"""
    HyperFieldModel(kdist, scale, data_idxs, grain, binsize, data; a0, a1, l0, l1)

The tutorial's model (docs/src/advanced_tutorial_lit.jl:194-307): `θ = [p₁, p₂, ξ]`, `sqrt_kernel` amplitude
`a0·exp(a1·p₁)·√(2π·cl)·exp(-π²k²cl²)` with `cl = l0·exp(l1·p₂)`, `gp_sample`, `exp` link, `agg_lambdas` over
`grain` fine pixels per data bin, Poisson counts.  `data_idxs` are `_DATA_IDXS` (1-based, regular).
"""
struct HyperFieldModel <: B200Model
    handle::ModelHandle
end

function HyperFieldModel(kdist::Vector{Float64}, scale::Real, data_idxs::AbstractVector{<:Integer}, grain::Integer,
                         binsize::Real, data::AbstractVector{<:Real}; a0::Real, a1::Real, l0::Real, l1::Real,
                         ctx::B200Context=b200_context())
    d = Float64.(data)
    all(diff(data_idxs) .== grain) || error("HyperFieldModel: data bins must be regular (start + grain * b)")
    N = length(kdist)
    desc = ModelDesc(MODEL_FIELD_HYPER, Int32(1), (Int64(N), 0, 0), C_NULL, Float64(scale), 0.0, LINK_EXP, LIKE_POISSON, 1.0,
                     LAYOUT_MU_ONLY, pointer(d), length(d), Int32(0), C_NULL, C_NULL, Int32(0), C_NULL, 0.0, C_NULL, 0, 0,
                     C_NULL, pointer(kdist), (Float64(a0), Float64(a1), Float64(l0), Float64(l1)),
                     Int64(first(data_idxs) - 1), Int32(grain), Float64(binsize))
    HyperFieldModel(_create(ctx, desc, kdist, d))
end

(m::B200Model)(ξ) = error("declarative B200 models are evaluated on the device; there is no CPU fallback")

n_theta(m::B200Model) = Int(ccall((:mgvib200_model_n_theta, libmgvib200), Int64, (Ptr{Cvoid},), m.handle.ptr))
n_lambda(m::B200Model) = Int(ccall((:mgvib200_model_n_lambda, libmgvib200), Int64, (Ptr{Cvoid},), m.handle.ptr))

# ---- solver / optimizer types ------------------------------------------------------------------
"CG sampler on the device; stands where `LinearSolve.KrylovJL_CG()` stands (src/mgvi_impl.jl:52)."
Base.@kwdef struct B200CG
    abstol::Float64 = sqrt(eps(Float64))
    reltol::Float64 = sqrt(eps(Float64))
    maxiters::Int = 0
end
const B200Dense = MGVI.MatrixInversion

"`MGVI.NewtonCG` on the device (src/newtoncg.jl:14-31)."
Base.@kwdef struct B200NewtonCG
    α::Float64 = 0.1
    steps::Int = 4
    i₀::Int = 5
    i₁::Int = 50
end

_cgopts(s::B200CG) = CGOpts(s.abstol, s.reltol, s.maxiters)
_cgopts(::Any) = CGOpts(sqrt(eps(Float64)), sqrt(eps(Float64)), 0)
_ncgopts(o::B200NewtonCG) = NewtonCGOpts(o.α, o.steps, o.i₀, o.i₁)
_ncgopts(o::MGVI.NewtonCG) = NewtonCGOpts(o.α, o.steps, o.i₀, o.i₁)

# ---- residual sampler: same constructor signature as src/residual_samplers.jl:59-63 ----------------
struct B200ResidualSampler{M<:B200Model,SLV,CTX<:MGVIContext}
    f_model::M
    center_point::Vector{Float64}
    linear_solver::SLV
    context::CTX
end

function MGVI.ResidualSampler(f_model::B200Model, center_point::Vector{<:Real}, linear_solver, context::MGVIContext)
    c = Float64.(center_point)
    _check(f_model.handle.ctx, ccall((:mgvib200_linearize, libmgvib200), Cint, (Ptr{Cvoid}, Ptr{Float64}), f_model.handle.ptr, c))
    B200ResidualSampler(f_model, c, linear_solver, context)
end

ImplicitResidualSampler(f_model, center, context; kw...) = MGVI.ResidualSampler(f_model, center, B200CG(; kw...), context)
FullResidualSampler(f_model, center, context) = MGVI.ResidualSampler(f_model, center, MGVI.MatrixInversion(), context)
struct FwdRevADJacobianFunc end     # old-API strategy names: JVP+VJP operator -> CG path
const FwdDerJacobianFunc = FwdRevADJacobianFunc
struct FullJacobianFunc end         # materialised Jacobian -> dense path

# draws in the order the reference consumes them (src/residual_samplers.jl:75-76): n1 then eta, per residual
function _draws(context::MGVIContext, nλ::Int, nθ::Int, n::Int)
    Z1 = Matrix{Float64}(undef, nλ, n); Z2 = Matrix{Float64}(undef, nθ, n)
    for i in 1:n
        Z1[:, i] .= randn(context.gen, nλ)
        Z2[:, i] .= randn(context.gen, nθ)
    end
    Z1, Z2
end

function MGVI.sample_residuals(s::B200ResidualSampler, n::Integer)
    m = s.f_model; nθ = n_theta(m); nλ = n_lambda(m)
    R = Matrix{Float64}(undef, nθ, n)
    if s.linear_solver isa MGVI.MatrixInversion
        Z = randn(s.context.gen, nθ, n)                                   # src/residual_samplers.jl:121
        _check(m.handle.ctx, ccall((:mgvib200_sample_residuals_dense, libmgvib200), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}), m.handle.ptr, Z, n, R, C_NULL))
    else
        Z1, Z2 = _draws(s.context, nλ, nθ, n)
        opts = Ref(_cgopts(s.linear_solver))
        iters = Vector{Int64}(undef, n); rn = Vector{Float64}(undef, n)
        _check(m.handle.ctx, ccall((:mgvib200_sample_residuals_cg, libmgvib200), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ref{CGOpts}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
               m.handle.ptr, Z1, Z2, n, opts, R, iters, rn))
    end
    R
end
MGVI.sample_residuals(s::B200ResidualSampler) = vec(MGVI.sample_residuals(s, 1))

# ---- mgvi_step / mgvi_sample for declarative models (src/mgvi_impl.jl:129-174) ----------------------
function MGVI.mgvi_step(forward_model::B200Model, data, n_residuals::Integer, center_init::AbstractVector{<:Real},
                        config::MGVIConfig, context::MGVIContext)
    m = forward_model; nθ = n_theta(m); nλ = n_lambda(m); n = Int(n_residuals)
    dense = config.linsolver isa MGVI.MatrixInversion
    config.optimizer isa Union{MGVI.NewtonCG,B200NewtonCG} || error("only NewtonCG runs on the device path")
    c = Float64.(center_init)
    Z1, Z2 = dense ? (randn(context.gen, nθ, n), Matrix{Float64}(undef, 0, 0)) : _draws(context, nλ, nθ, n)
    samples = Matrix{Float64}(undef, nθ, 2n); cout = Vector{Float64}(undef, nθ)
    mnlp = Ref{Float64}(0.0); iters = Vector{Int64}(undef, n); trace = NewtonCGTrace()
    cg = Ref(_cgopts(config.linsolver)); opts = Ref(_ncgopts(config.optimizer))
    GC.@preserve c Z1 Z2 samples cout iters trace begin
        _check(m.handle.ctx, ccall((:mgvib200_step, libmgvib200), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ref{CGOpts}, Ref{NewtonCGOpts},
                Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Int64}, Ref{NewtonCGTrace}),
               m.handle.ptr, c, Z1, dense ? C_NULL : pointer(Z2), n, dense ? 1 : 0, cg, opts, samples, cout, mnlp, iters, trace))
    end
    info = (linsolver_output = (iters = iters,), optimizer_output = trace)
    return MGVI.MGVIResult(samples, mnlp[], info), oftype(center_init, cout)
end

const mgvi_kl_optimize_step = MGVI.mgvi_step

function MGVI.mgvi_sample(forward_model::B200Model, data, n_residuals::Integer, center::AbstractVector{<:Real},
                          config::MGVIConfig, context::MGVIContext)
    m = forward_model; nθ = n_theta(m); nλ = n_lambda(m); n = Int(n_residuals)
    dense = config.linsolver isa MGVI.MatrixInversion
    c = Float64.(center)
    Z1, Z2 = dense ? (randn(context.gen, nθ, n), Matrix{Float64}(undef, 0, 0)) : _draws(context, nλ, nθ, n)
    samples = Matrix{Float64}(undef, nθ, 2n); iters = Vector{Int64}(undef, n)
    cg = Ref(_cgopts(config.linsolver))
    _check(m.handle.ctx, ccall((:mgvib200_sample, libmgvib200), Cint,
           (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ref{CGOpts}, Ptr{Float64}, Ptr{Int64}),
           m.handle.ptr, c, Z1, dense ? C_NULL : pointer(Z2), n, dense ? 1 : 0, cg, samples, iters))
    samples
end

# ---- the pieces, for users who drive the optimiser themselves (ext/MGVIOptimExt.jl idiom) -----------
"`(f, ∇f)` of `_mean_neg_log_pstr` (src/mgvi_impl.jl:19-28) at `x` for residuals `R`."
function kl_value_grad(m::B200Model, x::Vector{Float64}, R::Matrix{Float64})
    f = Ref{Float64}(0.0); g = similar(x)
    _check(m.handle.ctx, ccall((:mgvib200_kl_value_grad, libmgvib200), Cint,
           (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Float64}, Ptr{Float64}), m.handle.ptr, x, R, size(R, 2), f, g))
    f[], g
end

"`(J'FJ + I) * V` at the linearisation point (src/residual_samplers.jl:72)."
function metric_apply(s::B200ResidualSampler, V::Matrix{Float64})
    out = similar(V)
    _check(s.f_model.handle.ctx, ccall((:mgvib200_metric_apply, libmgvib200), Cint,
           (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), s.f_model.handle.ptr, V, size(V, 2), out))
    out
end

end # module
