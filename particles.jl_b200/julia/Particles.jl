# Particles.jl shim over libparticles_b200.so
#
# Keeps the reference's names and call shapes for the accelerated path (src/Particles.jl:23-56
# of kleinschmidt/Particles.jl) and forwards every computation to the C ABI declared in
# include/particles_b200.h via `ccall`.  No CUDA.jl, no CPU fallback: without a CUDA device
# every constructor throws.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed there.
# It is kept declarative on purpose: one struct mirroring pf_config, one `ccall` per entry
# point, index shifts (+1) where the ABI is 0-based.  The Python mirror with the same surface
# (particles.jl_b200/particles_b200/__init__.py) is what the test-suite drives.
module Particles

using LinearAlgebra

export ParticleFilter, FearnheadParticles, ChenLiuParticles, InfiniteParticle, Component,
       NormalInverseChisq, NormalInverseWishart, ChineseRestaurantProcess,
       fit!, particles, weight, nobs, ncomponents, components, assignments,
       ncomponents_dist, assignment_similarity

const LIB = get(ENV, "PARTICLES_B200_LIB", joinpath(@__DIR__, "..", "libparticles_b200.so"))

# ---------------------------------------------------------------- priors (ConjugatePriors.jl names)
struct NormalInverseChisq
    μ::Float64; σ2::Float64; κ::Float64; ν::Float64
end
struct NormalInverseWishart
    mu::Vector{Float64}; kappa::Float64; Lam::Matrix{Float64}; nu::Float64
end
struct ChineseRestaurantProcess          # src/statepriors.jl:47-52
    α::Float64
    N::Vector{Float64}
end
ChineseRestaurantProcess(α::Float64) = ChineseRestaurantProcess(α, Float64[])

# ---------------------------------------------------------------- pf_config / pf_step_stats
struct PfConfig
    filter_kind::Int32; prior_kind::Int32; d::Int32; precision::Int32
    n_particles::Int64
    k_max::Int32; order::Int32
    history::Int64
    device::Int32; reserved0::Int32; rank::Int32; nranks::Int32
    alpha::Float64; rejuv_threshold::Float64; kappa0::Float64; nu0::Float64; sigma2_0::Float64
    mu0::Ptr{Float64}; lambda0::Ptr{Float64}
    seed::UInt64
end
struct PfStepStats
    step::Int64; n_in::Int64; n_putative::Int64; n_kept::Int64; n_resamp_req::Int64
    n_resamp_got::Int64; n_out::Int64; overflow::Int64
    log_total_w::Float64; log_w_resamp::Float64; threshold::Float64; cv::Float64
    resampled::Int32; select_rounds::Int32; n_candidates::Int64
    select_phase_us::NTuple{4,Float64}
end

const PF_FEARNHEAD, PF_CHENLIU = Int32(0), Int32(1)
const PF_NIX2, PF_NIW = Int32(0), Int32(1)
const PF_ORDER_INDEX, PF_ORDER_SORTED = Int32(0), Int32(1)

function check(rc::Int32, h::Ptr{Cvoid}=C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:pf_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    rc == 1 && throw(ArgumentError(msg))           # the reference's ArgumentError sites
    error("particles_b200 error $rc: $msg")
end

# ---------------------------------------------------------------- filters
abstract type ParticleFilter end               # src/filters.jl:4

mutable struct DeviceFilter{P} <: ParticleFilter
    handle::Ptr{Cvoid}
    kind::Int32                 # PF_FEARNHEAD | PF_CHENLIU
    N::Int
    prior::P
    stateprior::ChineseRestaurantProcess
    d::Int
    k_max::Int
    rejuvination_threshold::Float64
end

function create(kind::Int32, n::Int, prior, stateprior::ChineseRestaurantProcess;
                rejuv::Float64=50., k_max::Int=16, order::Int32=PF_ORDER_SORTED, history::Int=0,
                device::Int=0, seed::Integer=0)
    if prior isa NormalInverseChisq
        pk, d, mu0, lam0 = PF_NIX2, 1, [prior.μ], zeros(1, 1)
        κ0, ν0, σ20 = prior.κ, prior.ν, prior.σ2
    else
        pk, d, mu0, lam0 = PF_NIW, length(prior.mu), copy(prior.mu), Matrix{Float64}(prior.Lam)
        κ0, ν0, σ20 = prior.kappa, prior.nu, 0.0
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve mu0 lam0 begin
        cfg = PfConfig(kind, pk, d, 0, n, k_max, order, history, device, 0, 0, 1,
                       stateprior.α, rejuv, κ0, ν0, σ20, pointer(mu0), pointer(lam0), UInt64(seed))
        check(ccall((:pf_create, LIB), Int32, (Ref{PfConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    end
    f = DeviceFilter(h[], kind, n, prior, stateprior, d, k_max, rejuv)
    finalizer(x -> ccall((:pf_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), f)
    return f
end

# FearnheadParticles(n, prior, stateprior)   src/fearnhead.jl:29-30
# The default traversal order is the reference-literal one (PF_ORDER_SORTED); pass
# order=PF_ORDER_INDEX for the sort-free variant of src/fearnhead.jl:140-145.
FearnheadParticles(n::Int, prior, stateprior::ChineseRestaurantProcess; kw...) =
    create(PF_FEARNHEAD, n, prior isa Tuple ? NormalInverseChisq(prior...) : prior, stateprior; kw...)
# ChenLiuParticles(n, prior, stateprior; rejuv=50.)   src/chenliu.jl:25-26
ChenLiuParticles(n::Int, prior, stateprior::ChineseRestaurantProcess; rejuv::Float64=50., kw...) =
    create(PF_CHENLIU, n, prior isa Tuple ? NormalInverseChisq(prior...) : prior, stateprior; rejuv=rejuv, kw...)

# fit!(ps, y)   src/fearnhead.jl:130-184, src/chenliu.jl:54-69
# rand() sites of the reference become host-supplied uniforms (Fearnhead: 1, Chen-Liu: 2N).
function fit!(ps::DeviceFilter, y; uniforms::Union{Nothing,Vector{Float64}}=nothing)
    yv = Float64[y...]
    length(yv) == ps.d || throw(DimensionMismatch("observation has $(length(yv)) entries, prior has dimension $(ps.d)"))
    u = uniforms === nothing ? rand(nuniforms(ps)) : uniforms
    GC.@preserve yv u check(ccall((:pf_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                                  ps.handle, yv, u, length(u)), ps.handle)
    return ps
end
# uniforms one observation consumes: Fearnhead 1 (src/fearnhead.jl:99); Chen-Liu N for the
# categorical draws (src/chenliu.jl:35) + N for a rejuvenation (src/chenliu.jl:61)
nuniforms(ps::DeviceFilter) = ps.kind == PF_FEARNHEAD ? 1 : 2 * ps.N

# filter!(ps, ys, progress; cb)   src/filters.jl:6-13
# Without a callback the whole stream is ONE library call (the loop runs inside the .so);
# with one, the callback sees the filter after every observation as in the reference.
function Base.filter!(ps::DeviceFilter, ys::AbstractVector, progress=true; cb=nothing,
                      uniforms::Union{Nothing,Matrix{Float64}}=nothing)
    if cb !== nothing
        for (t, y) in enumerate(ys)
            fit!(ps, y; uniforms = uniforms === nothing ? nothing : uniforms[:, t])
            cb(ps, y)
        end
        return ps
    end
    T = length(ys)
    Y = Matrix{Float64}(undef, ps.d, T)                 # d x T column-major, src/filters.jl:8 order
    for (t, y) in enumerate(ys); Y[:, t] .= y; end
    U = uniforms === nothing ? rand(nuniforms(ps), T) : uniforms
    le = Vector{Float64}(undef, T)
    GC.@preserve Y U le check(ccall((:pf_filter, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
        ps.handle, Y, T, U, size(U, 1), le), ps.handle)
    return ps
end

# ---------------------------------------------------------------- read-out
Base.length(ps::DeviceFilter) = Int(ccall((:pf_n_particles, LIB), Int64, (Ptr{Cvoid},), ps.handle))

function logweights(ps::DeviceFilter)
    out = Vector{Float64}(undef, length(ps))
    check(ccall((:pf_get_logweights, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), ps.handle, out), ps.handle)
    out
end

function ncomponents(ps::DeviceFilter)
    out = Vector{Int32}(undef, length(ps))
    check(ccall((:pf_get_ncomponents, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ps.handle, out), ps.handle)
    Int.(out)
end

# Component / InfiniteParticle views (src/component.jl:1-4, src/particle.jl:38-45), materialised lazily
struct Component{P}
    prior::P
    n::Int                     # nobs (src/component.jl:88)
    mean::Vector{Float64}      # NormalStats.m / MvNormalStats.m
    chol::Matrix{Float64}      # lower Cholesky factor of the posterior scale Λ_n
    logdet::Float64
end
nobs(c::Component) = c.n
Base.isempty(c::Component) = c.n == 0

struct InfiniteParticle{P}
    filter::DeviceFilter{P}
    index::Int                 # 1-based
    weight::Float64
    K::Int
end
weight(p::InfiniteParticle) = p.weight
ncomponents(p::InfiniteParticle) = p.K

function components(p::InfiniteParticle)
    ps = p.filter; d, km = ps.d, ps.k_max
    K = Ref{Int32}(0)
    counts = Vector{Float64}(undef, km); means = Matrix{Float64}(undef, d, km)
    chol = Array{Float64,3}(undef, d, d, km); logdet = Vector{Float64}(undef, km)
    check(ccall((:pf_get_particle, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ps.handle, p.index - 1, K, counts, means, chol, logdet), ps.handle)
    [Component(ps.prior, Int(counts[j]), means[:, j], chol[:, :, j], logdet[j]) for j in 1:K[]]
end
nobs(p::InfiniteParticle) = sum(nobs, components(p))

# particles(ps)   src/filters.jl:15
function particles(ps::DeviceFilter)
    lw, K = logweights(ps), ncomponents(ps)
    [InfiniteParticle(ps, i, exp(lw[i]), K[i]) for i in eachindex(lw)]
end

# assignments(ps)   src/filters.jl:26-34 : T x N matrix of 1-based labels
function assignments(ps::DeviceFilter)
    T = Int(ccall((:pf_n_steps, LIB), Int64, (Ptr{Cvoid},), ps.handle))
    out = Matrix{Int32}(undef, T, length(ps))
    check(ccall((:pf_get_assignments, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ps.handle, out), ps.handle)
    Int.(out) .+ 1
end

# ncomponents_dist(ps)   src/filters.jl:36-37 (weighted distribution of the component count)
function ncomponents_dist(ps::DeviceFilter)
    K, w = ncomponents(ps), exp.(logweights(ps))
    p = zeros(maximum(K) + 1)
    for (k, wi) in zip(K, w); p[k + 1] += wi; end
    p ./ sum(p)
end

# assignment_similarity(ps)   src/filters.jl:39-42
function assignment_similarity(ps::DeviceFilter)
    a = assignments(ps)
    T = size(a, 1)
    [count(a[s, :] .== a[t, :]) / size(a, 2) for s in 1:T, t in 1:T]
end

function last_step(ps::DeviceFilter)
    s = Ref{PfStepStats}()
    check(ccall((:pf_get_last_step, LIB), Int32, (Ptr{Cvoid}, Ref{PfStepStats}), ps.handle, s), ps.handle)
    s[]
end

end # module
