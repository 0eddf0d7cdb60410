# ext/TuringCUDANUTSExt.jl -- Turing.jl package extension that routes
#     sample(model, NUTS(δ), MCMCThreads(), N, nchains)
# to the B200 engine (libtnuts.so, C ABI in include/tnuts.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The file is
# the binding a Turing maintainer would add (pattern: ext/TuringDynamicHMCExt.jl:21-28 for the
# sampler type, src/mcmc/abstractmcmc.jl:133-166 for the ensemble `sample` override,
# src/mcmc/emcee.jl:124-140 for bundling many walkers into one chain object).  The Python
# host in turing.jl_b200/ binds the very same symbols and is what the tests exercise.
#
# Project.toml additions (cf. Project.toml:40-46):
#   [weakdeps]    TuringCUDANUTS_jll = "<uuid of the jll that ships libtnuts.so>"
#   [extensions]  TuringCUDANUTSExt = "TuringCUDANUTS_jll"
module TuringCUDANUTSExt

using Turing
using Turing: AbstractMCMC, Random, LogDensityProblems, DynamicPPL
using Turing.Inference: Hamiltonian, NUTS
using Turing.Inference: AdvancedHMC   # `import AdvancedHMC` in src/mcmc/Inference.jl:31
import MCMCChains

const libtnuts = get(ENV, "TNUTS_LIBRARY", "libtnuts.so")

# ---- mirror of include/tnuts.h -------------------------------------------------------------
struct TnutsModel
    family::Int32; D::Int32; N::Int64; G::Int32
    X::Ptr{Float64}; y::Ptr{Float64}; sigma::Ptr{Float64}; group::Ptr{Int32}
    hyper::NTuple{8,Float64}
end
struct TnutsConfig
    algorithm::Int32; delta::Float64; max_depth::Int32; delta_max::Float64; init_eps::Float64
    n_adapts::Int64; n_leapfrog::Int32; lambda::Float64; seed::UInt64
    n_chains::Int32; chain_offset::Int32; device::Int32; row_shards::Int32; row_rank::Int32
    strategy::Int32; metric::Int32; reserved::NTuple{4,Int32}
end
const TNUTS_NSTAT = 9
const STAT_NAMES = (:n_steps, :acceptance_rate, :log_density, :hamiltonian_energy,
    :hamiltonian_energy_error, :max_hamiltonian_energy_error, :tree_depth, :numerical_error,
    :step_size)

function check(h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:tnuts_last_error, libtnuts), Cstring, (Ptr{Cvoid},), h))
    # TNUTS_E_INIT carries the text pinned by test/mcmc/hmc.jl:292-295
    rc == -3 ? error(msg) : throw(ErrorException("tnuts error $rc: $msg"))
end

# ---- sampler type ---------------------------------------------------------------------------
"""
    CUDANUTS(spl::NUTS; family, data, device=0)

NUTS on the B200 engine.  `family`/`data` describe the registered model family the engine
evaluates instead of the opaque `DynamicPPL.Model` closure; `Turing.cuda(spl, model)` below
recognises the supported models and fills them in.
"""
struct CUDANUTS{AD} <: Hamiltonian
    n_adapts::Int; δ::Float64; max_depth::Int; Δ_max::Float64; ϵ::Float64; adtype::AD
    family::Int32; data::NamedTuple; device::Int32
    options::NamedTuple   # engine knobs passed to tnuts_set_option, e.g. (logistic_tc = 1,)
    metric::Int32         # TNUTS_METRIC_*: the sampler's metricT (src/mcmc/hmc.jl:404-406)
end
metric_id(::Type{<:AdvancedHMC.DiagEuclideanMetric}) = Int32(0)
metric_id(::Type{<:AdvancedHMC.UnitEuclideanMetric}) = Int32(1)
metric_id(::Type{<:AdvancedHMC.DenseEuclideanMetric}) = Int32(2)   # thread-per-chain families only
CUDANUTS(s::NUTS; family, data, device=0, options=NamedTuple()) =
    CUDANUTS(s.n_adapts, s.δ, s.max_depth, s.Δ_max, s.ϵ, s.adtype, Int32(family), data, Int32(device),
             options, metric_id(Turing.Inference.getmetricT(s)))

# ---- engine handle --------------------------------------------------------------------------
mutable struct Engine
    h::Ptr{Cvoid}; C::Int; D::Int; P::Int
end
function Engine(spl::CUDANUTS, nchains::Int, nadapts::Int, seed::UInt64)
    cfg = Ref(TnutsConfig(0, spl.δ, spl.max_depth, spl.Δ_max, spl.ϵ, nadapts, 0, 0.0, seed,
                          nchains, 0, spl.device, 1, 0, 0, spl.metric,
                          ntuple(_ -> Int32(0), 4)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:tnuts_create, libtnuts), Cint, (Ref{TnutsConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    d = spl.data
    X = get(d, :X, Float64[]); y = get(d, :y, Float64[]); σ = get(d, :sigma, Float64[])
    g = get(d, :group, Int32[])
    hyper = ntuple(i -> i <= length(d.hyper) ? Float64(d.hyper[i]) : 0.0, 8)
    GC.@preserve X y σ g begin
        m = Ref(TnutsModel(spl.family, d.D, d.N, get(d, :G, 0), pointer(X), pointer(y), pointer(σ),
                           pointer(g), hyper))
        check(h[], ccall((:tnuts_set_model, libtnuts), Cint, (Ptr{Cvoid}, Ref{TnutsModel}), h[], m))
    end
    for (k, v) in pairs(spl.options)
        check(h[], ccall((:tnuts_set_option, libtnuts), Cint, (Ptr{Cvoid}, Cstring, Cdouble),
                         h[], String(k), Float64(v)))
    end
    e = Engine(h[], nchains, d.D, ccall((:tnuts_num_params, libtnuts), Cint, (Ptr{Cvoid},), h[]))
    finalizer(x -> ccall((:tnuts_destroy, libtnuts), Cint, (Ptr{Cvoid},), x.h), e)
    return e
end

# ---- LogDensityProblems interface on the engine (cross-check against DynamicPPL's LDF) ------
struct EngineLogDensity
    e::Engine
end
LogDensityProblems.dimension(l::EngineLogDensity) = l.e.D
LogDensityProblems.capabilities(::Type{EngineLogDensity}) = LogDensityProblems.LogDensityOrder{1}()
function LogDensityProblems.logdensity_and_gradient(l::EngineLogDensity, θ::AbstractVector{Float64})
    lp = Ref(0.0); g = zeros(l.e.D)
    check(l.e.h, ccall((:tnuts_logdensity_and_gradient, libtnuts), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Float64}, Ptr{Float64}), l.e.h, θ, 1, lp, g))
    return lp[], g
end
LogDensityProblems.logdensity(l::EngineLogDensity, θ) = first(LogDensityProblems.logdensity_and_gradient(l, θ))

"verify, before sampling, that the engine's family reproduces the model's own log density"
function crosscheck(rng, model, spl::CUDANUTS, e::Engine; n=4,
                    rtol=get(spl.options, :logistic_tc, 0) != 0 ? 3e-7 : 1e-10)
    ldf = DynamicPPL.LogDensityFunction(model, DynamicPPL.getlogjoint_internal, DynamicPPL.LinkAll();
                                        adtype=spl.adtype)
    for _ in 1:n
        θ = 4 .* rand(rng, e.D) .- 2
        a, ga = LogDensityProblems.logdensity_and_gradient(ldf, θ)
        b, gb = LogDensityProblems.logdensity_and_gradient(EngineLogDensity(e), θ)
        (isapprox(a, b; rtol) && isapprox(ga, gb; rtol=max(1e-8, 300rtol))) ||
            error("the model is not the registered family the engine was given")
    end
end

# ---- the batched ensemble override (src/mcmc/abstractmcmc.jl:133-166, src/mcmc/hmc.jl:85-154) --
function AbstractMCMC.sample(rng::Random.AbstractRNG, model::DynamicPPL.Model, spl::CUDANUTS,
        ::AbstractMCMC.MCMCThreads, N::Integer, n_chains::Integer;
        chain_type=MCMCChains.Chains, check_model::Bool=true, verbose::Bool=true,
        initial_params=nothing, nadapts=spl.n_adapts, discard_adapt=true, discard_initial=-1,
        thinning=1, kwargs...)
    check_model && Turing._check_model(model, spl)
    _nadapts = nadapts == -1 ? min(1000, N ÷ 2) : nadapts                      # hmc.jl:106-110
    _discard = discard_initial == -1 ? (discard_adapt ? _nadapts : 0) : discard_initial  # :113-117
    e = Engine(spl, n_chains, _nadapts, rand(rng, UInt64))
    crosscheck(rng, model, spl, e)
    θ0 = initial_params === nothing ? C_NULL : permutedims(reduce(hcat, initial_params))  # [chains][D], linked
    t = @elapsed begin
        check(e.h, ccall((:tnuts_init, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, θ0))
        kept = Ref{Int64}(0)
        ntr, first_keep = _discard == 0 ? ((N - 1) * thinning, thinning) : (_discard + (N - 1) * thinning, _discard)
        check(e.h, ccall((:tnuts_run, libtnuts), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ref{Int64}),
                         e.h, ntr, first_keep, thinning, kept))
        ncol = ccall((:tnuts_num_columns, libtnuts), Cint, (Ptr{Cvoid},), e.h)
        # device layout [keep][ncol][chains] in C order == Julia Array (chains, ncol, keep)
        raw = Array{Float64}(undef, n_chains, ncol, kept[])
        check(e.h, ccall((:tnuts_fetch_draws, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, raw))
    end
    vals = permutedims(raw, (3, 2, 1))                          # (iterations × names × chains)
    pnames = Symbol.(DynamicPPL.varname_leaves_flat(model))     # model order, x[1], x[2] …
    inames = (:logprior, :loglikelihood, :logjoint, STAT_NAMES...)
    chn = MCMCChains.Chains(vals, vcat(collect(pnames), collect(inames)), (internals=collect(inames),);
                            info=(sampling_time=fill(t, n_chains),))
    Turing.Inference.post_sample_hook(chn, spl; verbose)        # divergence warning, hmc.jl:476-486
    return chn
end

Turing.Inference.getstepsize(::CUDANUTS, e::Engine) = begin
    ϵ = zeros(e.C)
    check(e.h, ccall((:tnuts_get_stepsize, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, ϵ)); ϵ
end

end # module
