# TEST INFRASTRUCTURE -- UNEXECUTED IN THIS REPOSITORY'S IMAGE (no Julia here).
#
# Turns "parity unpinned" into "pinned": runs the REAL bat/MGVI.jl v0.4.5 on the inputs of the golden
# cases of tests/golden/ -- with the normal draws given as INPUT, so that nothing depends on Julia's RNG
# streams -- and writes what the reference computes (metric products, residual samples, KL value and
# gradient, one full mgvi_step) as raw little-endian Float64 files that tests/test_ref_fixtures.py
# compares the oracle (CPU tier) and the CUDA path (GPU tier) against.
#
#   1.  python tests/golden/export_ref_inputs.py            # writes tests/golden/ref_inputs/<case>/
#   2.  JULIA_NUM_THREADS=1 julia --project=<env with MGVI 0.4.5, LinearSolve, Zygote, FFTW, Distributions, ValueShapes> \
#           oracle/make_ref_fixtures.jl /path/to/MGVI.jl    # writes tests/golden/ref/<case>/
#   3.  python -m pytest tests/test_ref_fixtures.py -q      # skipped while tests/golden/ref/ is absent
#
# JULIA_NUM_THREADS=1 matters: sample_residuals(s, n) fills its columns under Threads.@threads
# (src/residual_samplers.jl:85-92) and every column consumes draws from the shared RNG -- n_lambda normals,
# then n_theta normals (:75-76).  With one thread column i consumes the i-th (Z1[:, i], Z2[:, i]) pair.
#
# File format (both directions): <dir>/meta.txt with `name rows cols` per line, <dir>/<name>.f64 raw
# column-major Float64.

using Random, LinearAlgebra, Statistics
using MGVI, Distributions, ValueShapes
import LinearSolve, Zygote, FFTW
import AbstractFFTs

const REPO = normpath(joinpath(@__DIR__, ".."))
const REF_SRC = length(ARGS) >= 1 ? ARGS[1] : error("usage: make_ref_fixtures.jl /path/to/MGVI.jl")

# ---------------------------------------------------------------------------------------------
# raw array I/O
# ---------------------------------------------------------------------------------------------
function read_case(dir)
    d = Dict{String,Array{Float64}}()
    for ln in eachline(joinpath(dir, "meta.txt"))
        isempty(strip(ln)) && continue
        name, r, c = split(ln)
        r = parse(Int, r); c = parse(Int, c)
        a = Array{Float64}(undef, r, c)
        read!(joinpath(dir, name * ".f64"), a)
        d[name] = c == 1 ? vec(a) : a
    end
    return d
end

function write_case(dir, d::Dict)
    mkpath(dir)
    open(joinpath(dir, "meta.txt"), "w") do io
        for (name, a) in d
            m = a isa Number ? fill(Float64(a), 1, 1) : reshape(Float64.(a), size(a, 1), :)
            println(io, name, " ", size(m, 1), " ", size(m, 2))
            write(joinpath(dir, name * ".f64"), m)
        end
    end
end

# ---------------------------------------------------------------------------------------------
# an RNG that replays a given sequence of standard normals (randn is all the path draws)
# ---------------------------------------------------------------------------------------------
mutable struct ReplayRNG <: Random.AbstractRNG
    buf::Vector{Float64}
    pos::Int
end
ReplayRNG(buf) = ReplayRNG(collect(Float64, buf), 0)
function _next!(r::ReplayRNG)
    r.pos += 1
    r.pos <= length(r.buf) || error("ReplayRNG exhausted after $(length(r.buf)) draws")
    return r.buf[r.pos]
end
Random.randn(r::ReplayRNG) = _next!(r)
Random.randn(r::ReplayRNG, ::Type{Float64}) = _next!(r)
function Random.randn!(r::ReplayRNG, A::AbstractArray{Float64})
    for i in eachindex(A)
        A[i] = _next!(r)
    end
    return A
end

# draws in the order the reference consumes them: per residual sample, n_lambda then n_theta
function draw_stream(Z1, Z2)
    n = size(Z1, 2)
    return vcat((vcat(Z1[:, i], Z2[:, i]) for i in 1:n)...)
end

replay_context(Z1, Z2) = MGVIContext(MGVI.GenContext(ReplayRNG(draw_stream(Z1, Z2))), MGVI.ADSelector(Zygote))

# ---------------------------------------------------------------------------------------------
# models: the reference's own test models, with the case's data instead of the MersenneTwister draws
# ---------------------------------------------------------------------------------------------
include(joinpath(REF_SRC, "test", "test_models", "model_fft_gp.jl"))
include(joinpath(REF_SRC, "test", "test_models", "model_polyfit.jl"))

function run_case(name, model, inp; tight = true)
    center = inp["center"]; data = inp["data"]; Z1 = inp["Z1"]; Z2 = inp["Z2"]; V = inp["V"]
    Z1 = reshape(Z1, :, size(Z2, 2)); n = size(Z2, 2)
    out = Dict{String,Any}()
    # CG tolerances: the parity runs use reltol = 1e-12, abstol = 0 on both sides; if this LinearSolve version
    # refuses the keywords the defaults (sqrt(eps)) are used and recorded
    solver, used_tight = try
        tight ? (LinearSolve.KrylovJL_CG(atol = 0.0, rtol = 1e-12), 1.0) : (LinearSolve.KrylovJL_CG(), 0.0)
    catch
        (LinearSolve.KrylovJL_CG(), 0.0)
    end
    ctx = replay_context(Z1, Z2)
    s = ResidualSampler(model, center, solver, ctx)
    # a5: metric products through the sampler's own operators (src/residual_samplers.jl:72)
    M = s.jac_dλ_dθ' * s.λ_information * s.jac_dλ_dθ + I
    out["MV"] = hcat((M * V[:, i] for i in 1:size(V, 2))...)
    # a6/a7: residual samples with the replayed draws
    R = MGVI.sample_residuals(s, n)
    out["R"] = R
    out["tight"] = used_tight
    # a9/a10: KL value and gradient at the centre for these residuals (src/mgvi_impl.jl:3-28)
    f = MGVI._mean_neg_log_pstr(model, data, R, center)
    out["f"] = f
    out["grad"] = Zygote.gradient(p -> MGVI._mean_neg_log_pstr(model, data, R, p), center)[1]
    # a12/a14: one full step, same draws
    cfg = MGVIConfig(linsolver = solver, optimizer = MGVI.NewtonCG())
    res, c1 = mgvi_step(model, data, n, center, cfg, replay_context(Z1, Z2))
    out["step_center"] = c1
    out["step_samples"] = res.samples
    out["step_mnlp"] = res.mnlp
    # a8: MatrixInversion sampler (draws: n_theta normals per residual = Z2)
    ctxd = MGVIContext(MGVI.GenContext(ReplayRNG(vec(Z2))), MGVI.ADSelector(Zygote))
    sd = ResidualSampler(model, center, MGVI.MatrixInversion(), ctxd)
    out["L_sigma"] = Matrix(MGVI.residual_pushfwd_operator(sd))
    out["R_dense"] = MGVI.sample_residuals(sd, n)
    write_case(joinpath(REPO, "tests", "golden", "ref", name), out)
    println(name, ": mnlp ", res.mnlp, " tight ", used_tight)
end

for (name, model) in (("fft_gp_40", ModelFFTGP.model), ("polyfit_40", ModelPolyfit.model))
    dir = joinpath(REPO, "tests", "golden", "ref_inputs", name)
    isdir(dir) || (println("skip ", name, " (run tests/golden/export_ref_inputs.py first)"); continue)
    run_case(name, model, read_case(dir))
end
