# device.jl -- device-backed particle filters for kleinschmidt/Particles.jl over libparticles_b200.so
#
# ONE file added to the package, `include`d at the end of src/Particles.jl (after labeled.jl):
#
#     include("device.jl")
#
# and a guard in the package's two constructor one-liners (a method cannot be overwritten from a second file while the
# module precompiles, so the switch sits where the constructors are defined):
#
#     # src/fearnhead.jl:29-30
#     FearnheadParticles(n::Int, prior::Union{Tuple,<:Distribution}, stateprior::T) where T<:StatePrior =
#         USE_DEVICE[] ? DeviceParticles(PF_FEARNHEAD, n, prior, stateprior) :
#         FearnheadParticles([InfiniteParticle(prior, stateprior)], n)
#     # src/chenliu.jl:25-26
#     ChenLiuParticles(n::Int, prior::Union{Tuple,<:Distribution}, stateprior::T; rejuv::Float64=50.) where T<:StatePrior =
#         USE_DEVICE[] ? DeviceParticles(PF_CHENLIU, n, prior, stateprior; rejuv=rejuv) :
#         ChenLiuParticles([InfiniteParticle(prior, stateprior) for _ in 1:n], n, rejuv)
#
# The file adds no module of its own: it lives inside `module Particles` and uses the names that file already imports
# (ConjugatePriors' NormalInverseChisq / NormalInverseWishart, StatsBase.fit!, the package's StatePrior types).  Every
# computation is forwarded to the C ABI declared in include/particles_b200.h with `ccall`; there is no CUDA.jl and no
# CPU fallback (without a CUDA device the device constructors throw).
#
# Drop-in switch: after `Particles.use_device!()` the constructors above return device-backed filters
# (`DeviceParticles`; they can also be built directly), and the package's generic functions --
# fit!, filter!, particles, weight, components, ncomponents, nobs, assignments, ncomponents_dist,
# assignment_similarity, randindex, posterior_predictive, state_entropy, mean(f, ps) -- have methods
# for them below.  bench/1d.jl, bench/2d.jl, test/clustering.jl and test/labeled.jl then run
# unchanged.  `Particles.use_device!(false)` restores the CPU constructors.
#
# Julia is not installed in the build image, so this file has not been executed there; the C ABI it
# binds is exercised call for call by tests/host/abi_smoke.c and by the Python mirror
# (particles_b200/__init__.py), and tests/test_julia_shim.py checks every `ccall` below against the
# header (symbol, argument count) and the struct layouts against the C sizes.

const PF_LIB = get(ENV, "PARTICLES_B200_LIB", joinpath(@__DIR__, "..", "libparticles_b200.so"))

const PF_FEARNHEAD, PF_CHENLIU = Int32(0), Int32(1)
const PF_NIX2, PF_NIW = Int32(0), Int32(1)
const PF_F64, PF_F32 = Int32(0), Int32(1)
const PF_ORDER_INDEX, PF_ORDER_SORTED = Int32(0), Int32(1)
const PF_SP_CRP, PF_SP_STICKY, PF_SP_CHANGEPOINT, PF_SP_NSTATE = Int32(0), Int32(1), Int32(2), Int32(3)
const PF_MEAN_NCOMPONENTS, PF_MEAN_STATE_ENTROPY = Int32(0), Int32(1)

# pf_config (include/particles_b200.h), 144 bytes
struct PfConfig
    filter_kind::Int32; prior_kind::Int32; d::Int32; precision::Int32
    n_particles::Int64
    k_max::Int32; order::Int32
    history::Int64
    device::Int32; stateprior_kind::Int32; rank::Int32; nranks::Int32
    alpha::Float64; rejuv_threshold::Float64; kappa0::Float64; nu0::Float64; sigma2_0::Float64
    mu0::Ptr{Float64}; lambda0::Ptr{Float64}
    seed::UInt64
    sp_param::Float64
    sp_counts::Ptr{Float64}
    sp_n::Int32; reserved2::Int32
end
# pf_step_stats, 160 bytes
struct PfStepStats
    step::Int64; n_in::Int64; n_putative::Int64; n_kept::Int64; n_resamp_req::Int64
    n_resamp_got::Int64; n_out::Int64; overflow::Int64
    log_total_w::Float64; log_w_resamp::Float64; threshold::Float64; cv::Float64
    resampled::Int32; select_rounds::Int32
    n_candidates::Int64
    select_phase_us::NTuple{4,Float64}
    n_plateau::Int64
    select_status::Int32; reserved1::Int32
end

# status -> the exception the reference throws at the corresponding site
function pf_check(rc::Int32, h::Ptr{Cvoid}=C_NULL)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:pf_last_error, PF_LIB), Cstring, (Ptr{Cvoid},), h))
    rc == 1 && throw(ArgumentError(msg))        # src/particle.jl:137-139, src/statepriors.jl:141,169,175,212
    error("particles_b200 error $rc: $msg")
end

# ---------------------------------------------------------------- the device-backed filter
mutable struct DeviceParticles{P,S<:StatePrior} <: ParticleFilter
    handle::Ptr{Cvoid}
    kind::Int32                       # PF_FEARNHEAD | PF_CHENLIU
    N::Int
    prior::P                          # ConjugatePriors prior, as given
    stateprior::S
    d::Int
    k_max::Int
    rejuvination_threshold::Float64   # (the reference's spelling, src/chenliu.jl:10)
    overflow_seen::Int64
end

const USE_DEVICE = Ref(false)
"""
    use_device!(on=true)

Route `FearnheadParticles(n, prior, stateprior)` / `ChenLiuParticles(n, prior, stateprior)` to the B200 library.
"""
use_device!(on::Bool=true) = (USE_DEVICE[] = on)

# prior -> (kind, d, mu0, Lambda0 (d x d column-major), kappa0, nu0, sigma2_0); fields as in ConjugatePriors 0.3.2
pf_prior(p::NormalInverseChisq) = (PF_NIX2, 1, Float64[p.μ], zeros(1, 1), Float64(p.κ), Float64(p.ν), Float64(p.σ2))
function pf_prior(p::NormalInverseWishart)
    Λ = Matrix{Float64}(Matrix(p.Lamchol))        # Cholesky -> the matrix it factorises (src/component.jl:19-20,32)
    (PF_NIW, Int(p.dim), Vector{Float64}(p.mu), Λ, Float64(p.kappa), Float64(p.nu), 0.0)
end
pf_prior(p::Tuple) = pf_prior(NormalInverseChisq(p...))      # Component(params::NTuple), src/component.jl:8

# state prior -> (kind, alpha, param, counts)
pf_stateprior(s::ChineseRestaurantProcess) = (PF_SP_CRP, s.α, 0.0, Float64[])
pf_stateprior(s::StickyCRP) = (PF_SP_STICKY, s.α, s.ρ, Float64[])
pf_stateprior(s::ChangePoint) = (PF_SP_CHANGEPOINT, 1.0, exp(s.logp), Float64[])
pf_stateprior(s::NStatePrior) = (PF_SP_NSTATE, 1.0, 0.0, Vector{Float64}(s.counts))

"""
    DeviceParticles(kind, n, prior, stateprior; rejuv=50., k_max=32, order=PF_ORDER_INDEX, history=-1,
                    device=0, seed=0, precision=PF_F64)

`k_max`: cluster slots per particle (the reference's component vector is unbounded, src/particle.jl:141-146; a
dropped new-cluster candidate is reported by a warning).  `history`: generations kept for `assignments`
(-1 grows on demand like the reference's ancestor chain).  `order`: PF_ORDER_INDEX is the sort-free traversal the
reference's TODO proposes (src/fearnhead.jl:140-145, the fast path); PF_ORDER_SORTED restates `sort!` literally.
"""
function DeviceParticles(kind::Int32, n::Int, prior, stateprior::StatePrior; rejuv::Float64=50., k_max::Int=32,
                         order::Int32=PF_ORDER_INDEX, history::Int=-1, device::Int=0, seed::Integer=0,
                         precision::Int32=PF_F64)
    pk, d, mu0, lam0, κ0, ν0, σ20 = pf_prior(prior)
    sk, α, spp, counts = pf_stateprior(stateprior)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve mu0 lam0 counts begin
        cfg = PfConfig(kind, pk, d, precision, n, k_max, order, history, device, sk, 0, 1,
                       α, rejuv, κ0, ν0, σ20, pointer(mu0), pointer(lam0), UInt64(seed), spp,
                       isempty(counts) ? Ptr{Float64}(C_NULL) : pointer(counts), length(counts), 0)
        pf_check(ccall((:pf_create, PF_LIB), Int32, (Ref{PfConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    end
    f = DeviceParticles(h[], kind, n, prior, stateprior, d, k_max, rejuv, Int64(0))
    finalizer(x -> ccall((:pf_destroy, PF_LIB), Int32, (Ptr{Cvoid},), x.handle), f)
    return f
end

# direct constructors with the library's extra knobs (k_max, history, order, precision, device, seed)
DeviceFearnheadParticles(n::Int, prior, stateprior::StatePrior; kw...) = DeviceParticles(PF_FEARNHEAD, n, prior, stateprior; kw...)
DeviceChenLiuParticles(n::Int, prior, stateprior::StatePrior; rejuv::Float64=50., kw...) =
    DeviceParticles(PF_CHENLIU, n, prior, stateprior; rejuv=rejuv, kw...)

Base.length(ps::DeviceParticles) = Int(ccall((:pf_n_particles, PF_LIB), Int64, (Ptr{Cvoid},), ps.handle))
nsteps(ps::DeviceParticles) = Int(ccall((:pf_n_steps, PF_LIB), Int64, (Ptr{Cvoid},), ps.handle))

function last_step(ps::DeviceParticles)
    s = Ref{PfStepStats}()
    pf_check(ccall((:pf_get_last_step, PF_LIB), Int32, (Ptr{Cvoid}, Ref{PfStepStats}), ps.handle, s), ps.handle)
    s[]
end

function warn_overflow(ps::DeviceParticles)
    ov = last_step(ps).overflow
    if ov > ps.overflow_seen
        @warn "$(ov - ps.overflow_seen) new-cluster candidates were dropped: a particle already had k_max = $(ps.k_max) components (the reference's component vector is unbounded); construct the filter with a larger k_max"
        ps.overflow_seen = ov
    end
end

# uniforms one observation consumes when the caller supplies them: Fearnhead 1 (the rand() of src/fearnhead.jl:99);
# Chen-Liu N categorical draws (src/chenliu.jl:35) + N for a rejuvenation (src/chenliu.jl:61).  Without `uniforms`
# the library draws them from its counter-based device stream (pf_config.seed).
nuniforms(ps::DeviceParticles) = ps.kind == PF_FEARNHEAD ? 1 : 2 * ps.N

observation(y::Labeled) = y.obs
observation(y) = y
label0(y::Labeled) = y.label === missing ? Int32(-1) : Int32(y.label - 1)      # src/labeled.jl:1-8
label0(y) = Int32(-1)

# fit!(ps, y)   src/fearnhead.jl:130-184, src/chenliu.jl:54-69; fit!(ps, Labeled(y, label)) src/labeled.jl:10-11
function fit!(ps::DeviceParticles, y; uniforms::Union{Nothing,Vector{Float64}}=nothing)
    yv = Float64[observation(y)...]
    length(yv) == ps.d || throw(DimensionMismatch("observation has $(length(yv)) entries, the prior has dimension $(ps.d)"))
    lab = Int32[label0(y)]
    up = uniforms === nothing ? Ptr{Float64}(C_NULL) : pointer(uniforms)
    nu = uniforms === nothing ? 0 : length(uniforms)
    GC.@preserve yv uniforms lab begin
        if lab[1] >= 0
            pf_check(ccall((:pf_filter_labeled, PF_LIB), Int32,
                           (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                           ps.handle, yv, lab, 1, up, nu, C_NULL), ps.handle)
        else
            pf_check(ccall((:pf_step, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                           ps.handle, yv, up, nu), ps.handle)
        end
    end
    warn_overflow(ps)
    return ps
end

# filter!(ps, ys, progress; cb)   src/filters.jl:6-13.  Without a callback the whole stream is ONE library call (the
# loop runs inside the .so, observations pipelined on the device); with one, the callback sees the filter after every
# observation as in the reference.
function Base.filter!(ps::DeviceParticles, ys::AbstractVector, progress=true; cb=nothing,
                      uniforms::Union{Nothing,Matrix{Float64}}=nothing)
    if cb !== nothing
        for (t, y) in enumerate(ys)
            fit!(ps, y; uniforms = uniforms === nothing ? nothing : uniforms[:, t])
            cb(ps, y)
        end
        return ps
    end
    T = length(ys)
    Y = Matrix{Float64}(undef, ps.d, T)                 # d x T column-major
    for (t, y) in enumerate(ys); Y[:, t] .= observation(y); end
    labs = Int32[label0(y) for y in ys]
    le = Vector{Float64}(undef, T)
    up = uniforms === nothing ? Ptr{Float64}(C_NULL) : pointer(uniforms)
    ups = uniforms === nothing ? 0 : size(uniforms, 1)
    GC.@preserve Y uniforms labs le begin
        if any(>=(0), labs)
            pf_check(ccall((:pf_filter_labeled, PF_LIB), Int32,
                           (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                           ps.handle, Y, labs, T, up, ups, le), ps.handle)
        else
            pf_check(ccall((:pf_filter, PF_LIB), Int32,
                           (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                           ps.handle, Y, T, up, ups, le), ps.handle)
        end
    end
    warn_overflow(ps)
    return ps
end

# ---------------------------------------------------------------- read-out
function logweights(ps::DeviceParticles)
    out = Vector{Float64}(undef, length(ps))
    pf_check(ccall((:pf_get_logweights, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ps.handle, out), ps.handle)
    out
end

function Distributions.ncomponents(ps::DeviceParticles)
    out = Vector{Int32}(undef, length(ps))
    pf_check(ccall((:pf_get_ncomponents, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ps.handle, out), ps.handle)
    Int.(out)
end

# One particle of the device-resident population (src/particle.jl:38-45), materialised lazily: weight and K are read
# with the population, the components when asked for.
struct DeviceParticle{P,S} <: AbstractParticle
    filter::DeviceParticles{P,S}
    index::Int                 # 1-based
    weight::Float64
    K::Int
end
weight(p::DeviceParticle) = p.weight
Distributions.ncomponents(p::DeviceParticle, includeprior::Bool=false) = p.K + includeprior

# A cluster as the store keeps it: count, running mean and the lower Cholesky factor of the posterior scale Lambda_n;
# `Component(p)` rebuilds the reference's Component{P,S} (prior + sufficient statistics, src/component.jl:1-4)
struct DeviceComponent{P}
    prior::P
    n::Int
    mean::Vector{Float64}
    chol::Matrix{Float64}
    logdet::Float64
end
nobs(c::DeviceComponent) = c.n
Base.isempty(c::DeviceComponent) = c.n == 0

function device_components(p::DeviceParticle)
    ps = p.filter; d, km = ps.d, ps.k_max
    K = Ref{Int32}(0)
    counts = Vector{Float64}(undef, km); means = Matrix{Float64}(undef, d, km)
    chol = Array{Float64,3}(undef, d, d, km); logdet = Vector{Float64}(undef, km)
    pf_check(ccall((:pf_get_particle, PF_LIB), Int32,
                   (Ptr{Cvoid}, Int64, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   ps.handle, p.index - 1, K, counts, means, chol, logdet), ps.handle)
    [DeviceComponent(ps.prior, Int(counts[j]), means[:, j], chol[:, :, j], logdet[j]) for j in 1:K[]]
end
Distributions.components(p::DeviceParticle) = device_components(p)
nobs(p::DeviceParticle) = sum(nobs, device_components(p))

# particles(ps)   src/filters.jl:15
function particles(ps::DeviceParticles)
    lw, K = logweights(ps), ncomponents(ps)
    [DeviceParticle(ps, i, exp(lw[i]), K[i]) for i in eachindex(lw)]
end

# assignments(ps)   src/filters.jl:26-34 : T x N matrix of 1-based labels (the ancestor chain is walked on the device)
function assignments(ps::DeviceParticles)
    out = Matrix{Int32}(undef, nsteps(ps), length(ps))
    pf_check(ccall((:pf_get_assignments, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ps.handle, out), ps.handle)
    Int.(out) .+ 1
end

# mean(f, ps)   src/filters.jl:17-18: ncomponents and state_entropy are reduced on the device, any other f is a host
# function of the particle views, as in the reference
function Statistics.mean(f::Function, ps::DeviceParticles)
    what = f === ncomponents ? PF_MEAN_NCOMPONENTS : f === state_entropy ? PF_MEAN_STATE_ENTROPY : Int32(-1)
    if what < 0
        p = particles(ps)
        return mean(f.(p), Weights(weight.(p)))
    end
    out = Ref{Float64}(0.0)
    pf_check(ccall((:pf_weighted_mean, PF_LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}), ps.handle, what, out), ps.handle)
    out[]
end
state_entropy(ps::DeviceParticles) = mean(state_entropy, ps)               # src/filters.jl:20

# ncomponents_dist(ps)   src/filters.jl:36-37
function ncomponents_dist(ps::DeviceParticles)
    p = Vector{Float64}(undef, ps.k_max + 1)
    pf_check(ccall((:pf_ncomponents_dist, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ps.handle, p), ps.handle)
    Categorical(p[2:findlast(>(0), p)])                                     # categories 1..max K
end

# assignment_similarity(ps)   src/filters.jl:39-42 (co-clustering counts on the device)
function assignment_similarity(ps::DeviceParticles)
    T = nsteps(ps)
    out = Matrix{Float64}(undef, T, T)
    pf_check(ccall((:pf_assignment_similarity, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ps.handle, out), ps.handle)
    out
end

# randindex(ps, truth, type)   src/filters.jl:53-80
function randindex(ps::DeviceParticles, truth::T, type=:adjusted) where {T<:AbstractVector{<:Integer}}
    length(truth) == nsteps(ps) ||
        throw(DimensionMismatch("Ground truth and particles have different number of obs!"))
    t = Int32.(truth)
    out = Ref{Float64}(0.0)
    pf_check(ccall((:pf_randindex, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int64, Int32, Ref{Float64}),
                   ps.handle, t, length(t), Int32(type == :adjusted), out), ps.handle)
    out[]
end

# pdf.(posterior_predictive(ps), xs)   src/filters.jl:22-24 (the mixture is evaluated on the device; xs: d x Q)
function posterior_predictive_pdf(ps::DeviceParticles, xs::AbstractMatrix{<:Real})
    X = Matrix{Float64}(xs)
    size(X, 1) == ps.d || throw(DimensionMismatch("query points must be $(ps.d) x Q"))
    out = Vector{Float64}(undef, size(X, 2))
    pf_check(ccall((:pf_posterior_predictive, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                   ps.handle, X, size(X, 2), out), ps.handle)
    out
end
posterior_predictive_pdf(ps::DeviceParticles, xs::AbstractVector{<:Real}) = posterior_predictive_pdf(ps, reshape(Float64.(xs), 1, :))

# checkpoint / restore of the device-resident state
save(ps::DeviceParticles, path::AbstractString) =
    pf_check(ccall((:pf_save, PF_LIB), Int32, (Ptr{Cvoid}, Cstring), ps.handle, path), ps.handle)
function restore!(ps::DeviceParticles, path::AbstractString; device::Int=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    pf_check(ccall((:pf_load, PF_LIB), Int32, (Cstring, Int32, Ref{Ptr{Cvoid}}), path, device, h))
    ccall((:pf_destroy, PF_LIB), Int32, (Ptr{Cvoid},), ps.handle)
    ps.handle = h[]
    ps
end

# ---------------------------------------------------------------- one process driving several GPUs
# The Fearnhead / Chen-Liu population block-sharded over `devices` (pf_mg_create_local); every exchange between the
# shards happens inside the library's kernels over NVLink.  filter! is ONE library call.
mutable struct ShardedDeviceParticles <: ParticleFilter
    handles::Vector{Ptr{Cvoid}}
    N::Int
    d::Int
end
function ShardedDeviceParticles(kind::Int32, n::Int, prior, stateprior::ChineseRestaurantProcess, devices::Vector{Int};
                                rejuv::Float64=50., k_max::Int=32, history::Int=0, seed::Integer=0)
    pk, d, mu0, lam0, κ0, ν0, σ20 = pf_prior(prior)
    hs = fill(Ptr{Cvoid}(C_NULL), length(devices))
    devs = Int32.(devices)
    GC.@preserve mu0 lam0 begin
        cfg = PfConfig(kind, pk, d, PF_F64, n, k_max, PF_ORDER_INDEX, history, 0, PF_SP_CRP, 0, length(devices),
                       stateprior.α, rejuv, κ0, ν0, σ20, pointer(mu0), pointer(lam0), UInt64(seed), 0.0, Ptr{Float64}(C_NULL), 0, 0)
        pf_check(ccall((:pf_mg_create_local, PF_LIB), Int32, (Ref{PfConfig}, Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}),
                       cfg, devs, length(devs), hs))
    end
    f = ShardedDeviceParticles(hs, n, d)
    finalizer(x -> foreach(h -> ccall((:pf_destroy, PF_LIB), Int32, (Ptr{Cvoid},), h), x.handles), f)
    f
end
function Base.filter!(ps::ShardedDeviceParticles, ys::AbstractVector, progress=true)
    T = length(ys)
    Y = Matrix{Float64}(undef, ps.d, T)
    for (t, y) in enumerate(ys); Y[:, t] .= y; end
    le = Vector{Float64}(undef, T)
    GC.@preserve Y le pf_check(ccall((:pf_mg_filter, PF_LIB), Int32,
        (Ptr{Ptr{Cvoid}}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
        ps.handles, length(ps.handles), Y, T, C_NULL, 0, le), ps.handles[1])
    ps
end
# assignments(ps): every shard walks its particles' ancestor chains through the peers' genealogy arrays
function assignments(ps::ShardedDeviceParticles)
    blocks = map(ps.handles) do h
        n = Int(ccall((:pf_n_particles, PF_LIB), Int64, (Ptr{Cvoid},), h))
        T = Int(ccall((:pf_n_steps, PF_LIB), Int64, (Ptr{Cvoid},), h))
        out = Matrix{Int32}(undef, T, n)
        pf_check(ccall((:pf_mg_get_assignments, PF_LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), h, out), h)
        out
    end
    Int.(hcat(blocks...)) .+ 1
end
