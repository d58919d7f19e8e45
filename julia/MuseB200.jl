# MuseB200.jl — Julia glue that makes libmuse_b200.so a drop-in backend of MuseInference.jl for the
# registered model families.  UNEXECUTED: there is no Julia toolchain in the build image or on the GPU
# box; this file is the binding a maintainer would add (INTEGRATION.md), written against
# include/muse_b200.h.  The Python ctypes host (museinference.jl_b200/) binds the same symbols and is
# what the tests and the benchmark run.
#
# It replaces exactly the three mapped blocks of the reference,
#     muse!   src/muse.jl:169-176      get_J!  src/muse.jl:508-525      get_H!  src/muse.jl:417-442 (+ src/util.jl:9-26)
# and keeps every other line of those functions (outer θ solver, prior terms, covariance assembly).
module MuseB200

using MuseInference
using MuseInference: AbstractMuseProblem, MuseResult, standardizeθ, finalize_result!
using Random, Statistics, LinearAlgebra

const libmuse = get(ENV, "LIBMUSE_B200", "libmuse_b200.so")

const FAMILY = Dict(:funnel => Cint(1), :hiergauss => Cint(2), :corrgauss => Cint(3), :twolayer => Cint(4))
const START_ZEROS, START_PREV, START_TRUTH, START_USER = Cint(0), Cint(1), Cint(2), Cint(3)

# mirrors `struct muse_cfg` (include/muse_b200.h)
Base.@kwdef struct MuseCfg
    abi_version::Cint = 1
    family::Cint
    d::Cint
    ntheta::Cint
    nsims::Cint
    device::Cint = 0
    sim_offset::Int64 = 0
    nsims_h::Cint = 0
    kernel::Cint = 0
    h_sim_offset::Int64 = 0
    lbfgs_m::Cint = 0
    max_iters::Cint = 0
    group::Cint = 0
    cluster::Cint = 0
    P::Ptr{Cdouble} = C_NULL
    L::Ptr{Cdouble} = C_NULL
    stream::Ptr{Cvoid} = C_NULL
end

struct B200Error <: Exception
    code::Cint
    msg::String
end

function check(h, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:muse_b200_last_error, libmuse), Cstring, (Ptr{Cvoid},), h))
    throw(B200Error(rc, msg))
end

"""
    B200MuseProblem(x, family; logPriorθ = θ -> 0)

The B200 counterpart of `SimpleMuseProblem` (src/simple.jl:4-12).  A GPU backend cannot introspect Julia
closures, so the model is *named*: `:funnel` (src/simple.jl:58-76), `:hiergauss`, `:corrgauss` or `:twolayer` (the toy hierarchy of
src/turing.jl:63-79 with `x = vcat(sim.x, sim.y)`, parameter σ).
Any other `AbstractMuseProblem` (Turing, Soss, arbitrary closures) is not supported: calling `muse` on it
through this backend throws; there is no CPU fallback.
"""
mutable struct B200MuseProblem <: AbstractMuseProblem
    x::Vector{Float64}
    family::Symbol
    logPriorθ
    prior::Union{Nothing,Tuple{Vector{Float64},Vector{Float64}}}   # (mean, sigma) of an independent Normal prior, if that is what logPriorθ is:
                                                                   # lets muse! run the whole solve inside the library (muse_b200_muse_solve)
    P::Union{Nothing,Matrix{Float64}}        # :corrgauss — Σ₀⁻¹ and chol(Σ₀).L, passed ROW-major to the library (= the transposes of
    L::Union{Nothing,Matrix{Float64}}        # Julia's column-major arrays; P is symmetric, L is transposed once below)
    transform::Vector{Symbol}                # per component :identity | :log — transform_θ / inv_transform_θ (src/interface.jl:14-28)
    handle::Ptr{Cvoid}
    nsims::Int
    seed::UInt64
end
function B200MuseProblem(x, family::Symbol; logPriorθ = θ -> 0.0, prior = nothing, Σ₀ = nothing, transform = nothing)
    haskey(FAMILY, family) || error("model family $family is not registered with the B200 backend")
    P = L = nothing
    if family === :corrgauss
        Σ₀ === nothing && error(":corrgauss needs the fixed covariance Σ₀ (d × d, SPD)")
        P = Matrix{Float64}(inv(Symmetric(Σ₀)))
        L = permutedims(Matrix{Float64}(cholesky(Symmetric(Σ₀)).L))    # row-major L for the library = column-major Lᵀ here
    end
    nθ = family === :hiergauss ? 2 : 1
    tr = transform === nothing ? fill(:identity, nθ) : collect(Symbol, transform)
    length(tr) == nθ && all(t -> t in (:identity, :log), tr) || error("transform: one of :identity | :log per θ component")
    if prior !== nothing
        logPriorθ = θ -> -sum(abs2, (collect(θ) .- prior[1]) ./ prior[2]) / 2
        prior = (collect(Float64, prior[1]), collect(Float64, prior[2]))
    end
    B200MuseProblem(collect(Float64, x), family, logPriorθ, prior, P, L, tr, C_NULL, 0, 0)
end
MuseInference.logPriorθ(p::B200MuseProblem, θ) = p.logPriorθ(θ)
ntheta(p::B200MuseProblem) = p.family === :hiergauss ? 2 : 1
# the kernels are parameterised in the unconstrained space: θ′ = transform_θ(θ) is what they take, and ∇θ′ logLike what they return
MuseInference.transform_θ(p::B200MuseProblem, θ) = [t === :log ? log(v) : v for (t, v) in zip(p.transform, θ)]
MuseInference.inv_transform_θ(p::B200MuseProblem, θ′) = [t === :log ? exp(v) : v for (t, v) in zip(p.transform, θ′)]
dθdθ′(p::B200MuseProblem, θ′) = [t === :log ? exp(v) : 1.0 for (t, v) in zip(p.transform, θ′)]
has_transform(p::B200MuseProblem) = any(!=(:identity), p.transform)

function backend!(p::B200MuseProblem, nsims::Integer, seed::UInt64)
    if p.handle == C_NULL || p.nsims != nsims
        p.handle != C_NULL && ccall((:muse_b200_destroy, libmuse), Cint, (Ptr{Cvoid},), p.handle)
        Pp = p.P === nothing ? Ptr{Cdouble}(C_NULL) : pointer(p.P)
        Lp = p.L === nothing ? Ptr{Cdouble}(C_NULL) : pointer(p.L)
        cfg = Ref(MuseCfg(family = FAMILY[p.family], d = length(p.x), ntheta = ntheta(p), nsims = nsims, P = Pp, L = Lp))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = GC.@preserve p ccall((:muse_b200_create, libmuse), Cint, (Ref{MuseCfg}, Ref{Ptr{Cvoid}}), cfg, h)   # P, L are copied to the device
        rc == 0 || throw(B200Error(rc, unsafe_string(ccall((:muse_b200_last_error, libmuse), Cstring, (Ptr{Cvoid},), C_NULL))))
        p.handle, p.nsims, p.seed = h[], nsims, typemax(UInt64)
        GC.@preserve p check(p.handle, ccall((:muse_b200_set_data, libmuse), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.handle, p.x))
    end
    if p.seed != seed   # split_rng (src/util.jl:85-92): the same child streams at every call for a given master rng
        check(p.handle, ccall((:muse_b200_seed_draws, libmuse), Cint, (Ptr{Cvoid}, UInt64), p.handle, seed))
        p.seed = seed
    end
    p.handle
end

# body of src/muse.jl:169-176 / 508-514 for the whole batch
function map_score(p::B200MuseProblem, h, θsim, θeval, atol; include_data::Bool, warm_start::Cint, first_sim = 0, count = p.nsims)
    units = count + include_data
    g = Matrix{Float64}(undef, ntheta(p), units)          # column-major nθ × units == row-major units × nθ
    iters = Vector{Cint}(undef, units); fg = similar(iters); status = similar(iters)
    gnorm = Vector{Float64}(undef, units)
    ts, te = collect(Float64, θsim), collect(Float64, θeval)
    GC.@preserve ts te g iters fg gnorm status check(h, ccall((:muse_b200_map_score, libmuse), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cint}),
        h, ts, te, atol, include_data, warm_start, first_sim, count, g, iters, fg, gnorm, status))
    any(==(4), status) && error("MAP solution failed with a non-finite objective (src/interface.jl:170)")
    (; g, iters, fg, gnorm, status)
end

seed_of(rng::AbstractRNG) = rand(copy(rng), UInt64)       # a pure function of the master rng, which is not advanced

function MuseInference.muse!(result::MuseResult, prob::B200MuseProblem, θ₀ = nothing;
        rng = nothing, maxsteps = 50, θ_rtol = 1e-1, ∇z_logLike_atol = 1e-2, nsims = 100, α = 0.7,
        regularize = identity, get_covariance = false, kwargs...)
    result.rng = rng = something(rng, result.rng, copy(Random.default_rng()))           # src/muse.jl:134
    θ = θunreg = standardizeθ(prob, something(result.θ, θ₀))                            # :135
    h = backend!(prob, nsims, seed_of(rng))
    history = result.history
    # Common configuration (fresh result, constant α, identity regularize and transform, flat or Normal prior, isotropic family):
    # the loop AND the covariance stage run inside the library — one kernel launch per solve (muse_b200_muse_solve, DESIGN.md §3.6)
    if isempty(history) && regularize === identity && α isa Real && !has_transform(prob) && prob.family !== :corrgauss && prob.family !== :twolayer &&
       maxsteps <= 64 && isempty(kwargs) && (prob.prior !== nothing || prob.logPriorθ(θ) == 0)
        return fused_solve!(result, prob, h, θ; maxsteps, θ_rtol, ∇z_logLike_atol, nsims, α, get_covariance)
    end
    first_pass = true
    for i = (length(history) + 1):maxsteps                                              # :159
        if i > 2                                                                        # :163-166
            Δθ = history[end].θ .- history[end-1].θ
            sqrt(-(Δθ' * history[end].H⁻¹_post * Δθ)) < θ_rtol && break
        end
        out = map_score(prob, h, θ, θ, ∇z_logLike_atol; include_data = true,            # :169-176 → one ccall
                        warm_start = first_pass ? START_ZEROS : START_PREV)
        first_pass = false
        g_like_dat, g_like_sims = out.g[:, 1], [out.g[:, k] for k = 2:size(out.g, 2)]
        g_like = g_like_dat .- mean(g_like_sims)                                        # :183
        g_prior = MuseInference.AD.gradient(MuseInference.AD.ForwardDiffBackend(), θ -> prob.logPriorθ(θ), θ)[1]
        g_post = g_like .+ g_prior
        H⁻¹_like = Diagonal(-1 ./ var(g_like_sims))                                     # :188-189
        H_prior = MuseInference.AD.hessian(MuseInference.AD.ForwardDiffBackend(), θ -> prob.logPriorθ(θ), θ)[1]
        H⁻¹_post = inv(inv(H⁻¹_like) + H_prior)                                         # :207-208
        push!(history, (; θ, θunreg, g_like_sims, g_like_dat, g_like, g_prior, g_post, H⁻¹_post, H_prior, H⁻¹_like))
        θunreg = θ .- α .* (H⁻¹_post * g_post)                                          # :224
        θ = regularize(θunreg)
        result.θ = θunreg; result.gs = g_like_sims                                      # :230-231
    end
    if get_covariance                                                                   # :244-247
        MuseInference.get_J!(result, prob; rng, nsims, ∇z_logLike_atol)
        MuseInference.get_H!(result, prob; rng, nsims = max(1, nsims ÷ 10), ∇z_logLike_atol)
    end
    result
end

# mirrors `muse_iterate_out` / `muse_cov_out` (include/muse_b200.h): pointers into Julia arrays that outlive the call
struct IterateOut
    n_iter::Cint
    theta_final::Ptr{Cdouble}; theta_hist::Ptr{Cdouble}; g_dat_hist::Ptr{Cdouble}; g_sims_hist::Ptr{Cdouble}
    g_like_hist::Ptr{Cdouble}; g_prior_hist::Ptr{Cdouble}; h_inv_like_hist::Ptr{Cdouble}; h_prior_hist::Ptr{Cdouble}
    h_inv_post_hist::Ptr{Cdouble}; seconds_hist::Ptr{Cdouble}
    iters_hist::Ptr{Cint}; fg_hist::Ptr{Cint}; gnorm_hist::Ptr{Cdouble}; status_hist::Ptr{Cint}
end
struct CovOut
    J::Ptr{Cdouble}; step::Ptr{Cdouble}; Hs::Ptr{Cdouble}; H::Ptr{Cdouble}; Sigma_inv::Ptr{Cdouble}; Sigma::Ptr{Cdouble}
end

function fused_solve!(result, prob, h, θ₀; maxsteps, θ_rtol, ∇z_logLike_atol, nsims, α, get_covariance)
    nθ, K, units, nH = ntheta(prob), maxsteps, nsims + 1, max(1, nsims ÷ 10)
    f(dims...) = zeros(Float64, dims...)
    θfin, θh, gd, gs = f(nθ), f(nθ, K), f(nθ, K), f(nθ, nsims, K)          # column-major (c, k, row) == C row-major [row][k][c]
    gl, gp, hil, hp, hip, secs = f(nθ, K), f(nθ, K), f(nθ, K), f(nθ, K), f(nθ, K), f(K)
    its, fg, st = zeros(Cint, units, K), zeros(Cint, units, K), zeros(Cint, units, K)
    gn = f(units, K)
    J, step, Hs, H, Σi, Σ = f(nθ, nθ), f(nθ), f(nθ, nθ, nH), f(nθ, nθ), f(nθ, nθ), f(nθ, nθ)
    t0 = collect(Float64, θ₀)
    pm, ps = prob.prior === nothing ? (Ptr{Cdouble}(C_NULL), Ptr{Cdouble}(C_NULL)) : (pointer(prob.prior[1]), pointer(prob.prior[2]))
    GC.@preserve prob t0 θfin θh gd gs gl gp hil hp hip secs its fg gn st J step Hs H Σi Σ begin
        out = Ref(IterateOut(0, pointer(θfin), pointer(θh), pointer(gd), pointer(gs), pointer(gl), pointer(gp), pointer(hil), pointer(hp),
                             pointer(hip), pointer(secs), pointer(its), pointer(fg), pointer(gn), pointer(st)))
        cov = Ref(CovOut(pointer(J), pointer(step), pointer(Hs), pointer(H), pointer(Σi), pointer(Σ)))
        check(h, ccall((:muse_b200_muse_solve, libmuse), Cint,
            (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Cdouble, Cdouble, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cint},
             Ref{IterateOut}, Ref{CovOut}),
            h, t0, nsims, C_NULL, maxsteps, θ_rtol, ∇z_logLike_atol, α, START_ZEROS, pm, ps, get_covariance, nH, C_NULL, out, cov))
        n = out[].n_iter
        for i = 1:n                                                                     # the rows muse! pushes at :211-221
            g_like_sims = [gs[:, k, i] for k = 1:nsims]
            push!(result.history, (; θ = θh[:, i], θunreg = i == 1 ? t0 : θh[:, i], g_like_sims, g_like_dat = gd[:, i], g_like = gl[:, i],
                                   g_prior = gp[:, i], g_post = gl[:, i] .+ gp[:, i], H⁻¹_post = Diagonal(hip[:, i]),
                                   H_prior = Diagonal(hp[:, i]), H⁻¹_like = Diagonal(hil[:, i]), t = secs[i]))
        end
        result.θ = θfin; result.gs = [gs[:, k, n] for k = 1:nsims]                      # :230-231
        if get_covariance                                                               # :244-247 ran on the same stream
            result.J, result.H = permutedims(J), permutedims(H)
            result.Hs = [permutedims(Hs[:, :, k]) for k = 1:nH]
            finalize_result!(result, prob)
        end
    end
    result
end

function MuseInference.get_J!(result::MuseResult, prob::B200MuseProblem, θ₀ = nothing;
        rng = nothing, nsims = 100, ∇z_logLike_atol = 1e-2, kwargs...)
    rng = something(rng, result.rng, copy(Random.default_rng()))
    θ₀ = standardizeθ(prob, something(θ₀, result.θ))                                    # :498
    existing = length(result.gs)
    if nsims > existing                                                                 # :499-506
        h = backend!(prob, nsims, seed_of(rng))
        out = map_score(prob, h, θ₀, θ₀, ∇z_logLike_atol; include_data = false, warm_start = START_TRUTH,
                        first_sim = existing, count = nsims - existing)                 # :508-514
        append!(result.gs, [out.g[:, k] for k = 1:size(out.g, 2)])
    end
    result.J = θ₀ isa Number || length(θ₀) == 1 ? var(result.gs) : cov(result.gs)       # :529
    finalize_result!(result, prob)
end

function MuseInference.get_H!(result::MuseResult, prob::B200MuseProblem, θ₀ = nothing;
        rng = nothing, nsims = 10, step = nothing, ∇z_logLike_atol = 1e-2, z₀ = nothing, implicit_diff = false,
        implicit_diff_cg_kwargs = (maxiter = 100,), kwargs...)
    rng = something(rng, result.rng, copy(Random.default_rng()))
    θ₀ = standardizeθ(prob, something(θ₀, result.θ))                                    # :315
    remaining = nsims - length(result.Hs)
    remaining > 0 || return result
    h = backend!(prob, max(prob.nsims, remaining), seed_of(rng))
    nθ = ntheta(prob)
    if z₀ !== nothing                                                                   # :309, 419 — start of the fiducial solve
        zz = collect(Float64, z₀)
        GC.@preserve zz check(h, ccall((:muse_b200_set_z0, libmuse), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h, zz))
    end
    if implicit_diff                                                                    # :335-405
        Hs = Array{Float64}(undef, nθ, nθ, remaining); its = Matrix{Cint}(undef, nθ, remaining)
        t0 = collect(Float64, θ₀)
        GC.@preserve t0 Hs its check(h, ccall((:muse_b200_implicit_h, libmuse), Cint,
            (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}),
            h, t0, remaining, z₀ === nothing ? START_ZEROS : START_USER, get(implicit_diff_cg_kwargs, :maxiter, 100), Hs, its, C_NULL))
        append!(result.Hs, [permutedims(Hs[:, :, k]) for k = 1:remaining])
        append!(get!(() -> [], result.metadata, :implicit_diff_cg_hists), [its[:, k] for k = 1:remaining])
        result.H = mean(result.Hs)
        return finalize_result!(result, prob)
    end
    z₀ === nothing || check(h, ccall((:muse_b200_fd_start, libmuse), Cint, (Ptr{Cvoid}, Cint), h, START_USER))
    # :411-413.  With neither `step` nor scores the reference leaves step = nothing and FiniteDifferences estimates one per sim and
    # component (src/util.jl:13); museinference.jl_b200/muse.py (_fd_jacobian_adaptive) does that with muse_b200_fd_scores calls — here
    # the caller is asked for a step instead.
    (step === nothing && isempty(result.gs)) && error("get_H!: pass `step` or run muse!/get_J! first")
    step = something(step, 0.1 ./ std(result.gs))
    Hs = Array{Float64}(undef, nθ, nθ, remaining)     # column-major (n, i, k) == C row-major [k][i][n]
    t0, st = collect(Float64, θ₀), collect(Float64, step)
    GC.@preserve t0 st Hs check(h, ccall((:muse_b200_fd_jacobian, libmuse), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cint}),
        h, t0, st, remaining, ∇z_logLike_atol, Hs, C_NULL))                             # :417-442 + src/util.jl:9-26
    z₀ === nothing || check(h, ccall((:muse_b200_fd_start, libmuse), Cint, (Ptr{Cvoid}, Cint), h, START_ZEROS))
    append!(result.Hs, [permutedims(Hs[:, :, k]) for k = 1:remaining])
    result.H = mean(result.Hs)                                                          # :446
    finalize_result!(result, prob)
end

# Bounded θ (transform_θ / inv_transform_θ, src/interface.jl:14-28).  The kernels take the unconstrained θ′ and return
# ∇θ′ logLike, so a problem with positive components passes θ′ = transform_θ(prob, θ) to `map_score`, divides the scores by
# ∂θ/∂θ′ where the reference asks for UnTransformedθ() (src/muse.jl:172, 432, 513), and — because pjacobian perturbs the
# UNtransformed θ₀ (src/util.jl:15) — gets the raw ± scores of get_H!'s virtual sims at the mapped points from this call,
# forming sum(fs .* [-1/2, 0, 1/2]) / step itself (museinference.jl_b200/muse.py does exactly this; DESIGN.md §8.1).
function fd_scores(p::B200MuseProblem, h, θeval′, θsims′::Matrix{Float64}, nsims_H, atol)
    nθ = ntheta(p)
    size(θsims′) == (nθ, 2nθ) || error("θsims′ must be nθ × 2nθ (column 2n-1 / 2n = the − / + point of Jacobian column n)")
    g = Array{Float64}(undef, nθ, 2nθ, nsims_H)       # column-major (i, point, k) == C row-major [k][point][i]
    te = collect(Float64, θeval′)
    GC.@preserve te θsims′ g check(h, ccall((:muse_b200_fd_scores, libmuse), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cint}),
        h, te, θsims′, nsims_H, atol, g, C_NULL))
    g
end

end # module
