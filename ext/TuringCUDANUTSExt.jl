# ext/TuringCUDANUTSExt.jl -- Turing.jl package extension that routes
#     sample(model, NUTS(δ), MCMCThreads(), N, nchains)
# to the B200 engine (libtnuts.so, C ABI in include/tnuts.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The file is the
# binding a Turing maintainer would add.  Patterns followed:
#   sampler type + converter ........ ext/TuringDynamicHMCExt.jl:21-28
#   AbstractMCMC.step (init/iterate)  src/mcmc/hmc.jl:156-208, 210-252; ext/TuringDynamicHMCExt.jl:46-96
#   ensemble `sample` override ...... src/mcmc/abstractmcmc.jl:133-166
#   nadapts / discard / resume ...... src/mcmc/hmc.jl:85-154
#   chain assembly .................. ext/TuringMCMCChainsExt.jl, src/mcmc/emcee.jl:124-140
# The Python host in turing.jl_b200/ binds the very same symbols with the same semantics and is
# what the tests exercise; tests/c_abi/drive.c drives them from plain C.
#
# Project.toml additions (cf. Project.toml:40-46):
#   [weakdeps]    TuringCUDANUTS_jll = "<uuid of the jll that ships libtnuts.so>"
#   [extensions]  TuringCUDANUTSExt = "TuringCUDANUTS_jll"
module TuringCUDANUTSExt

using Turing
using Turing: AbstractMCMC, Random, LogDensityProblems, DynamicPPL
using Turing.Inference: AdaptiveHamiltonian, NUTS, find_initial_params_ldf
using Turing.Inference: AdvancedHMC   # `import AdvancedHMC` in src/mcmc/Inference.jl:31
import MCMCChains

const libtnuts = get(ENV, "TNUTS_LIBRARY", "libtnuts.so")
"`TuringCUDANUTSExt.ENABLED[] = false` sends recognised models back to the CPU path"
const ENABLED = Ref(true)

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
const TNUTS_GDEMO, TNUTS_LINREG, TNUTS_EIGHT_SCHOOLS, TNUTS_LOGISTIC, TNUTS_HIER_LINREG, TNUTS_STD_NORMAL,
      TNUTS_BETA_BERNOULLI, TNUTS_DIRICHLET_CATEGORICAL = Int32.(0:7)

function check(h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:tnuts_last_error, libtnuts), Cstring, (Ptr{Cvoid},), h))
    # TNUTS_E_INIT carries the text pinned by test/mcmc/hmc.jl:292-295
    rc == -3 ? error(msg) : throw(ErrorException("tnuts error $rc: $msg"))
end

# ---- registered model families --------------------------------------------------------------
# Turing hands its samplers an opaque closure; the engine needs (family id, data arrays, prior
# hyper-parameters).  A family is registered per `@model` function: `extract(model)` returns the
# NamedTuple the engine's tnuts_model descriptor is filled from,
#     (D, N, [G,] [X,] [y,] [sigma,] [group,] hyper)
# with theta in linked space and variables in model statement order (DESIGN.md section 1).
const FAMILIES = IdDict{Any,Tuple{Int32,Function}}()
"""
    register_family!(f, family, extract)

Route `sample(f(args...), NUTS(δ), MCMCThreads(), N, n)` to the engine.  `f` is the function an
`@model` definition creates, `family` one of `TNUTS_*`, `extract(model)::NamedTuple` the data.
Before the first transition the engine's log density is cross-checked against the model's own
`LogDensityFunction` (`crosscheck`), so a wrong registration fails loudly instead of sampling
another posterior.
"""
register_family!(f, family::Integer, extract::Function) = (FAMILIES[f] = (Int32(family), extract); nothing)
recognise(model::DynamicPPL.Model) = get(FAMILIES, model.f, nothing)

# ---- sampler type ---------------------------------------------------------------------------
"""
    CUDANUTS(spl::NUTS; family, data, device=0, options=(;))

NUTS on the B200 engine: the same knobs as `NUTS` (`n_adapts, δ, max_depth, Δ_max, ϵ`, `metricT`)
plus what the engine needs.  Users do not construct it: the `sample` method below converts a plain
`NUTS(δ)` when the model's family is registered (`register_family!`), exactly like
`Turing.externalsampler(::DynamicHMC.NUTS)` wraps its sampler (ext/TuringDynamicHMCExt.jl:28).
"""
struct CUDANUTS{AD} <: AdaptiveHamiltonian
    n_adapts::Int; δ::Float64; max_depth::Int; Δ_max::Float64; ϵ::Float64; adtype::AD
    family::Int32; data::NamedTuple; device::Int32
    options::NamedTuple   # engine knobs passed to tnuts_set_option, e.g. (logistic_tc = 1,)
    metric::Int32         # TNUTS_METRIC_*: the sampler's metricT (src/mcmc/hmc.jl:404-406)
end
metric_id(::Type{<:AdvancedHMC.DiagEuclideanMetric}) = Int32(0)
metric_id(::Type{<:AdvancedHMC.UnitEuclideanMetric}) = Int32(1)
metric_id(::Type{<:AdvancedHMC.DenseEuclideanMetric}) = Int32(2)   # thread-per-chain families; D <= 256 on the pump strategy
CUDANUTS(s::NUTS; family, data, device=0, options=NamedTuple()) =
    CUDANUTS(s.n_adapts, s.δ, s.max_depth, s.Δ_max, s.ϵ, s.adtype, Int32(family), data, Int32(device),
             options, metric_id(Turing.Inference.getmetricT(s)))
CUDANUTS(s::NUTS, model::DynamicPPL.Model; kwargs...) = begin
    fam = recognise(model)
    fam === nothing && throw(ArgumentError("no engine family is registered for this model (register_family!)"))
    CUDANUTS(s; family=fam[1], data=fam[2](model), kwargs...)
end

# The unchanged user call.  More specific than src/mcmc/abstractmcmc.jl:133 (spl::AbstractSampler,
# ensemble::AbstractMCMCEnsemble): a recognised model goes to the engine, everything else -- and
# everything when ENABLED[] is false -- takes Turing's own method.
function AbstractMCMC.sample(rng::Random.AbstractRNG, model::DynamicPPL.Model, spl::NUTS,
        ensemble::AbstractMCMC.MCMCThreads, N::Integer, n_chains::Integer; kwargs...)
    if ENABLED[] && recognise(model) !== nothing
        return AbstractMCMC.sample(rng, model, CUDANUTS(spl, model), ensemble, N, n_chains; kwargs...)
    end
    return invoke(AbstractMCMC.sample,
                  Tuple{Random.AbstractRNG,DynamicPPL.Model,AbstractMCMC.AbstractSampler,
                        AbstractMCMC.AbstractMCMCEnsemble,Integer,Integer},
                  rng, model, spl, ensemble, N, n_chains; kwargs...)
end

# ---- engine handle --------------------------------------------------------------------------
mutable struct Engine
    h::Ptr{Cvoid}; C::Int; D::Int; P::Int; seed::UInt64
end
function Engine(spl::CUDANUTS, nchains::Int, nadapts::Int, seed::UInt64)
    # AHMCAdaptor (src/mcmc/hmc.jl:454): `iszero(alg.n_adapts)` means NoAdaptation whatever nadapts is
    nad = iszero(spl.n_adapts) ? 0 : nadapts
    cfg = Ref(TnutsConfig(0, spl.δ, spl.max_depth, spl.Δ_max, spl.ϵ, nad, 0, 0.0, seed,
                          nchains, 0, spl.device, 1, 0, 0, spl.metric,
                          ntuple(_ -> Int32(0), 4)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:tnuts_create, libtnuts), Cint, (Ref{TnutsConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    d = spl.data
    X = get(d, :X, Float64[]); y = get(d, :y, Float64[]); σ = get(d, :sigma, Float64[])
    g = get(d, :group, Int32[])
    # the ABI wants X row-major [N][D]: a Julia Matrix (N x D, column-major) is transposed once here
    Xc = X isa AbstractMatrix ? collect(transpose(X)) : collect(Float64, X)
    hyper = ntuple(i -> i <= length(d.hyper) ? Float64(d.hyper[i]) : 0.0, 8)
    GC.@preserve Xc y σ g begin
        m = Ref(TnutsModel(spl.family, d.D, d.N, get(d, :G, 0), pointer(Xc), pointer(y), pointer(σ),
                           pointer(g), hyper))
        check(h[], ccall((:tnuts_set_model, libtnuts), Cint, (Ptr{Cvoid}, Ref{TnutsModel}), h[], m))
    end
    for (k, v) in pairs(spl.options)
        check(h[], ccall((:tnuts_set_option, libtnuts), Cint, (Ptr{Cvoid}, Cstring, Cdouble),
                         h[], String(k), Float64(v)))
    end
    e = Engine(h[], nchains, d.D, ccall((:tnuts_num_params, libtnuts), Cint, (Ptr{Cvoid},), h[]), seed)
    finalizer(x -> ccall((:tnuts_destroy, libtnuts), Cint, (Ptr{Cvoid},), x.h), e)
    return e
end
set_option!(e::Engine, k, v) =
    check(e.h, ccall((:tnuts_set_option, libtnuts), Cint, (Ptr{Cvoid}, Cstring, Cdouble), e.h, String(k), Float64(v)))
"current linked-space positions, D x chains"
function positions(e::Engine)
    θ = Matrix{Float64}(undef, e.D, e.C)       # C layout [chains][D] == Julia (D, chains)
    check(e.h, ccall((:tnuts_get_theta, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, θ)); θ
end
"constrained parameters (P x n) and user-space logprior / loglikelihood / logjoint (3 x n) of θ (D x n)"
function constrain(e::Engine, θ::Matrix{Float64})
    n = size(θ, 2); p = Matrix{Float64}(undef, e.P, n); lp = Matrix{Float64}(undef, 3, n)
    check(e.h, ccall((:tnuts_constrain, libtnuts), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}), e.h, θ, n, p, lp))
    return p, lp
end
function state_blob(e::Engine)
    nb = Ref{Csize_t}(0)
    check(e.h, ccall((:tnuts_get_state, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Csize_t}), e.h, C_NULL, nb))
    blob = Vector{UInt8}(undef, nb[])
    check(e.h, ccall((:tnuts_get_state, libtnuts), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ref{Csize_t}), e.h, blob, nb))
    return blob
end
restore!(e::Engine, blob::Vector{UInt8}) =
    check(e.h, ccall((:tnuts_set_state, libtnuts), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), e.h, blob, length(blob)))

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

make_ldf(model, spl::CUDANUTS; fix_transforms=false) =
    DynamicPPL.LogDensityFunction(model, DynamicPPL.getlogjoint_internal, DynamicPPL.LinkAll();
                                  adtype=spl.adtype, fix_transforms=fix_transforms)   # src/mcmc/hmc.jl:170-176

"verify, before sampling, that the engine's family reproduces the model's own log density"
function crosscheck(rng, ldf, spl::CUDANUTS, e::Engine; n=4,
                    rtol=get(spl.options, :logistic_tc, 0) != 0 ? 3e-7 : 1e-10)
    for _ in 1:n
        θ = 4 .* rand(rng, e.D) .- 2
        a, ga = LogDensityProblems.logdensity_and_gradient(ldf, θ)
        b, gb = LogDensityProblems.logdensity_and_gradient(EngineLogDensity(e), θ)
        (isapprox(a, b; rtol) && isapprox(ga, gb; rtol=max(1e-8, 300rtol))) ||
            error("the model is not the registered family the engine was given")
    end
end

stats_tuple(v::AbstractVector{Float64}) =
    (n_steps=Int(v[1]), is_accept=true, acceptance_rate=v[2], log_density=v[3], hamiltonian_energy=v[4],
     hamiltonian_energy_error=v[5], max_hamiltonian_energy_error=v[6], tree_depth=Int(v[7]),
     numerical_error=v[8] != 0, step_size=v[9], nom_step_size=v[9])

# ---- AbstractMCMC step interface on a one-chain engine (src/mcmc/hmc.jl:156-208, 210-252) -----
struct CUDANUTSState{L}
    i::Int
    engine::Engine
    ldf::L
end

function AbstractMCMC.step(rng::Random.AbstractRNG, model::DynamicPPL.Model, spl::CUDANUTS;
        initial_params::DynamicPPL.AbstractInitStrategy, nadapts=0, discard_sample=false,
        verbose::Bool=true, fix_transforms::Bool=false, kwargs...)
    ldf = make_ldf(model, spl; fix_transforms)
    e = Engine(spl, 1, nadapts, rand(rng, UInt64))
    crosscheck(rng, ldf, spl, e)
    # initial values: the reference's own retry loop (<= 1000 tries, warning at 10, URL in the error)
    θ = find_initial_params_ldf(rng, ldf, initial_params)          # linked space
    check(e.h, ccall((:tnuts_init, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, θ))  # + find_good_stepsize
    verbose && iszero(spl.ϵ) && @info "Found initial step size" ϵ = Turing.Inference.getstepsize(spl, e)[1]
    transition = discard_sample ? nothing : DynamicPPL.ParamsWithStats(θ, ldf, NamedTuple())
    return transition, CUDANUTSState(0, e, ldf)
end

function AbstractMCMC.step(rng::Random.AbstractRNG, model::DynamicPPL.Model, spl::CUDANUTS,
        state::CUDANUTSState; nadapts=0, discard_sample=false, kwargs...)
    e = state.engine
    kept = Ref{Int64}(0)
    # one transition + AHMC.adapt! (the engine counts i itself); kept so that its stats can be read
    check(e.h, ccall((:tnuts_run, libtnuts), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ref{Int64}), e.h, 1, 1, 1, kept))
    transition = if discard_sample
        nothing
    else
        st = Vector{Float64}(undef, TNUTS_NSTAT)
        check(e.h, ccall((:tnuts_fetch, libtnuts), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, C_NULL, C_NULL, st, C_NULL))
        # a16: ParamsWithStats re-runs the model at theta for the constrained values and log-probs
        DynamicPPL.ParamsWithStats(vec(positions(e)), state.ldf, stats_tuple(st))
    end
    return transition, CUDANUTSState(state.i + 1, e, state.ldf)
end

Turing.Inference.getstepsize(::CUDANUTS, e::Engine) = begin
    ϵ = zeros(e.C)
    check(e.h, ccall((:tnuts_get_stepsize, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, ϵ)); ϵ
end
Turing.Inference.getstepsize(spl::CUDANUTS, s::CUDANUTSState) = Turing.Inference.getstepsize(spl, s.engine)[1]

# ---- the batched ensemble method (src/mcmc/abstractmcmc.jl:133-166, src/mcmc/hmc.jl:85-154) -----
"final state of an ensemble run: what `save_state=true` stores and `initial_state` accepts"
struct CUDANUTSEnsembleState
    blob::Vector{UInt8}
    seed::UInt64
    n_chains::Int
end

function AbstractMCMC.sample(rng::Random.AbstractRNG, model::DynamicPPL.Model, spl::CUDANUTS,
        ::AbstractMCMC.MCMCThreads, N::Integer, n_chains::Integer;
        chain_type=Turing.Inference.DEFAULT_CHAIN_TYPE, check_model::Bool=true, verbose::Bool=true,
        initial_params=fill(Turing.Inference.init_strategy(spl), n_chains), initial_state=nothing,
        nadapts=spl.n_adapts, discard_adapt=true, discard_initial=-1, thinning=1, save_state=false,
        progress=false, kwargs...)
    check_model && Turing._check_model(model, spl)
    if !(initial_params isa AbstractVector) || length(initial_params) != n_chains
        throw(ArgumentError("`initial_params` must be an AbstractVector of length `n_chains`; one element per chain"))
    end
    resume = initial_state !== nothing
    if resume                                                      # hmc.jl:135-152
        _nadapts, _discard = 0, 0
    else
        _nadapts = nadapts == -1 ? min(1000, N ÷ 2) : nadapts                                   # :106-110
        _discard = discard_initial == -1 ? (discard_adapt ? _nadapts : 0) : discard_initial     # :113-117
    end
    ldf = make_ldf(model, spl)
    seed = resume ? initial_state.seed : rand(rng, UInt64)
    e = Engine(spl, n_chains, _nadapts, seed)
    crosscheck(rng, ldf, spl, e)
    generic = chain_type !== MCMCChains.Chains        # any other container goes through bundle_samples
    generic && set_option!(e, :keep_theta, 1)
    first_row = nothing
    t = @elapsed begin
        if resume
            initial_state.n_chains == n_chains || throw(ArgumentError("`initial_state` holds $(initial_state.n_chains) chains"))
            restore!(e, initial_state.blob)
            ntr, first_keep = 1 + (N - 1) * thinning, 1            # mcmcsample keeps the first step, then thins
        else
            inits = map(Turing._convert_initial_params, initial_params)
            θ0 = if all(s -> s isa DynamicPPL.InitFromUniform, inits)
                C_NULL                                             # the engine draws U(-2,2), <= 1000 tries per chain
            else
                # user-space values -> linked space through the model's own transforms; D x chains in
                # Julia memory IS the ABI's [chains][D]
                reduce(hcat, [find_initial_params_ldf(rng, ldf, s) for s in inits])
            end
            check(e.h, ccall((:tnuts_init, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, θ0))
            if _discard == 0
                # the first sample is the initial point, without sampler stats (test/mcmc/hmc.jl:173-197)
                θi = positions(e)
                first_row = (θi, constrain(e, θi)...)
                ntr, first_keep = (N - 1) * thinning, thinning
            else
                ntr, first_keep = _discard + (N - 1) * thinning, _discard
            end
        end
        kept = Ref{Int64}(0)
        ntr > 0 && check(e.h, ccall((:tnuts_run, libtnuts), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ref{Int64}),
                                    e.h, ntr, first_keep, thinning, kept))
        ncol = ccall((:tnuts_num_columns, libtnuts), Cint, (Ptr{Cvoid},), e.h)
        # device layout [keep][ncol][chains] in C order == Julia Array (chains, ncol, keep)
        raw = Array{Float64}(undef, n_chains, ncol, kept[])
        kept[] > 0 && check(e.h, ccall((:tnuts_fetch_draws, libtnuts), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, raw))
    end
    final_state = save_state ? CUDANUTSEnsembleState(state_blob(e), seed, n_chains) : nothing
    chn = generic ? bundle_generic(e, raw, first_row, model, spl, ldf, chain_type, t, final_state; kwargs...) :
                    bundle_chains(e, raw, first_row, model, t, final_state)
    Turing.Inference.post_sample_hook(chn, spl; verbose)        # divergence warning, hmc.jl:476-486
    return chn
end

# MCMCChains.Chains straight from the engine's (iterations x names x chains) array
function bundle_chains(e::Engine, raw, first_row, model, t, final_state)
    vals = permutedims(raw, (3, 2, 1))
    if first_row !== nothing
        _, p, lp = first_row
        row = Array{Union{Missing,Float64}}(missing, 1, size(raw, 2), e.C)
        row[1, 1:e.P, :] = p
        row[1, e.P+1:e.P+3, :] = lp
        vals = vcat(row, vals)
    end
    pnames = Symbol.(DynamicPPL.varname_leaves_flat(model))     # model order, x[1], x[2] ...
    inames = (:logprior, :loglikelihood, :logjoint, STAT_NAMES...)
    info = final_state === nothing ? (sampling_time=fill(t, e.C),) :
           (sampling_time=fill(t, e.C), samplerstate=final_state)
    return MCMCChains.Chains(vals, vcat(collect(pnames), collect(inames)), (internals=collect(inames),); info)
end

# any other chain_type (VNChain is the default, src/mcmc/Inference.jl:65): per chain, the vector of
# DynamicPPL.ParamsWithStats the CPU path would have produced, bundled by the container's own method
function bundle_generic(e::Engine, raw, first_row, model, spl, ldf, chain_type, t, final_state; kwargs...)
    K = size(raw, 3)
    θk = Array{Float64}(undef, e.C, e.D, K)                      # [keep][D][chains] in C order
    K > 0 && check(e.h, ccall((:tnuts_fetch, libtnuts), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, C_NULL, C_NULL, C_NULL, θk))
    s0 = e.P + 3
    chains = map(1:e.C) do c
        samples = Any[]
        first_row === nothing || push!(samples, DynamicPPL.ParamsWithStats(first_row[1][:, c], ldf, NamedTuple()))
        for k in 1:K
            push!(samples, DynamicPPL.ParamsWithStats(θk[c, :, k], ldf, stats_tuple(raw[c, s0+1:s0+TNUTS_NSTAT, k])))
        end
        AbstractMCMC.bundle_samples(identity.(samples), model, spl, final_state, chain_type;
                                    stats=(stop=t, start=0.0, duration=t), kwargs...)
    end
    return AbstractMCMC.chainscat(chains...)
end

Turing.Inference.loadstate(s::CUDANUTSEnsembleState) = s

end # module
