# CFMMRouterB200.jl -- thin `ccall` shim that keeps CFMMRouter.jl's
# Router / route! / CFMM / Objective API and sends the dual-decomposition inner
# loop (find_arb! sweep + Ψ/acc folds) to libcfmm_b200.so (include/cfmm_b200.h).
#
# STATUS: Julia is not installed in the build image, so this file has never been
# executed; it is written against the C ABI that the Python host
# (cfmmrouter.jl_b200/router.py, same call sequence) exercises in tests/.  It is
# kept deliberately small: everything that is not a ccall is the reference's own
# logic, re-used from the CFMMRouter package (objectives, pool structs, L-BFGS-B).
#
# Usage (drop-in for `using CFMMRouter` on the route! path):
#     using CFMMRouterB200           # re-exports CFMMRouter's names
#     r = B200Router(LinearNonnegative(c), pools, n)    # instead of Router(...)
#     route!(r); netflows(r); r.Δs; r.Λs; r.v           # unchanged
module CFMMRouterB200

using CFMMRouter
using CFMMRouter: CFMM, ProductTwoCoin, GeometricMeanTwoCoin, UniV3, Objective,
                  f, grad!, lower_limit, upper_limit
using LBFGSB
import CFMMRouter: route!, find_arb!, netflows!, netflows, update_reserves!

export B200Router, sync_reserves!

const LIB = get(ENV, "CFMM_B200_LIB", joinpath(@__DIR__, "..", "libcfmm_b200.so"))

struct B200Error <: Exception
    code::Cint
    msg::String
end

function chk(ctx::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:cfmm_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    rc == -1 ? throw(ArgumentError(msg)) : throw(B200Error(rc, msg))   # ArgumentError as the ctors do (cfmms.jl:77-78)
end

# Same fields as Router (src/router.jl:4-10) + the device context and the cached
# folds of the last sweep.
mutable struct B200Router{O,T}
    objective::O
    cfmms::Vector{CFMM{T}}
    Δs::Vector{AbstractVector{T}}
    Λs::Vector{AbstractVector{T}}
    v::Vector{T}
    ctx::Ptr{Cvoid}
    order::Vector{Int}      # library insertion order -> index into cfmms
    ψ::Vector{T}            # Σ A_i(Λ_i − Δ_i) of the last sweep
    acc::Base.RefValue{T}   # Σ ν[A_i]ᵀ(Λ_i − Δ_i) of the last sweep
end

# Router(objective, cfmms, n_tokens), src/router.jl:18-35: pack Vector{CFMM} -> SoA, upload.
function B200Router(objective::O, cfmms::Vector{C}, n_tokens; device::Integer=0) where {T,O<:Objective,C<:CFMM{T}}
    T === Float64 || throw(ArgumentError("libcfmm_b200 is fp64-only"))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cfmm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64), out, device, n_tokens)
    rc == 0 || throw(B200Error(rc, unsafe_string(ccall((:cfmm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    ctx = out[]
    order = Int[]
    prod = findall(c -> c isa ProductTwoCoin, cfmms)
    geo = findall(c -> c isa GeometricMeanTwoCoin, cfmms)
    uni = findall(c -> c isa UniV3, cfmms)
    length(prod) + length(geo) + length(uni) == length(cfmms) ||
        throw(MethodError(find_arb!, (cfmms,)))    # what the reference would hit
    if !isempty(prod)
        R = Float64[c.R[j] for c in cfmms[prod] for j in 1:2]
        γ = Float64[c.γ for c in cfmms[prod]]
        Ai = Int64[c.Ai[j] for c in cfmms[prod] for j in 1:2]
        GC.@preserve R γ Ai chk(ctx, ccall((:cfmm_add_product, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), ctx, length(prod), R, γ, Ai))
        append!(order, prod)
    end
    if !isempty(geo)
        R = Float64[c.R[j] for c in cfmms[geo] for j in 1:2]
        w = Float64[c.w[j] for c in cfmms[geo] for j in 1:2]
        γ = Float64[c.γ for c in cfmms[geo]]
        Ai = Int64[c.Ai[j] for c in cfmms[geo] for j in 1:2]
        GC.@preserve R w γ Ai chk(ctx, ccall((:cfmm_add_geomean, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}), ctx, length(geo), R, γ, Ai, w))
        append!(order, geo)
    end
    if !isempty(uni)
        cp = Float64[c.current_price for c in cfmms[uni]]
        γ = Float64[c.γ for c in cfmms[uni]]
        Ai = Int64[c.Ai[j] for c in cfmms[uni] for j in 1:2]
        off = Int64[0; cumsum(Int64[length(c.lower_ticks) for c in cfmms[uni]])]
        lt = reduce(vcat, (Float64.(c.lower_ticks) for c in cfmms[uni]))
        lq = reduce(vcat, (Float64.(c.liquidity) for c in cfmms[uni]))
        GC.@preserve cp γ Ai off lt lq chk(ctx, ccall((:cfmm_add_univ3, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            ctx, length(uni), cp, γ, Ai, off, lt, lq))
        append!(order, uni)
    end
    chk(ctx, ccall((:cfmm_finalize, LIB), Cint, (Ptr{Cvoid},), ctx))
    Δs = AbstractVector{T}[zeros(T, 2) for _ in cfmms]     # zerotrade, router.jl:23-26
    Λs = AbstractVector{T}[zeros(T, 2) for _ in cfmms]
    r = B200Router{O,T}(objective, convert(Vector{CFMM{T}}, cfmms), Δs, Λs, zeros(T, n_tokens),
                        ctx, order, zeros(T, n_tokens), Ref(zero(T)))
    finalizer(x -> ccall((:cfmm_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), r)
    return r
end

# One dual-gradient sweep on the GPU: find_arb!(r, v) (router.jl:38-42) + both
# folds (router.jl:79-83, 98-100).  materialize=true also refreshes r.Δs / r.Λs.
function sweep!(r::B200Router{O,T}, v::Vector{T}; materialize::Bool=false) where {O,T}
    GC.@preserve v chk(r.ctx, ccall((:cfmm_sweep, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Cint), r.ctx, v, r.ψ, r.acc, materialize ? 1 : 0))
    if materialize
        m = length(r.order)
        D = Vector{Float64}(undef, 2m); L = Vector{Float64}(undef, 2m)
        chk(r.ctx, ccall((:cfmm_get_trades, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), r.ctx, D, L))
        for (k, i) in enumerate(r.order)          # library order -> r.cfmms order
            r.Δs[i][1] = D[2k-1]; r.Δs[i][2] = D[2k]
            r.Λs[i][1] = L[2k-1]; r.Λs[i][2] = L[2k]
        end
    end
    return nothing
end

find_arb!(r::B200Router, v) = sweep!(r, collect(Float64, v); materialize=true)

# route!, src/router.jl:58-108, with the three find_arb!(r, v) call sites and the
# two fold loops replaced by sweep!.  Everything else is the reference's code path.
function route!(r::B200Router; v=nothing, verbose=false, m=5, factr=1e1, pgtol=1e-5, maxfun=15_000, maxiter=15_000)
    optimizer = L_BFGS_B(length(r.v), 17)
    if isnothing(v)
        r.v .= ones(length(r.v)) / length(r.v)
    else
        r.v .= v
    end
    bounds = zeros(3, length(r.v))
    bounds[1, :] .= 2
    bounds[2, :] .= lower_limit(r.objective)
    bounds[3, :] .= upper_limit(r.objective)

    function fn(x)                       # router.jl:73-86
        if !all(x .== r.v)
            sweep!(r, x)
            r.v .= x
        end
        return f(r.objective, x) + r.acc[]
    end
    function g!(G, x)                    # router.jl:89-102
        G .= 0
        if !all(x .== r.v)
            sweep!(r, x)
            r.v .= x
        end
        grad!(G, r.objective, x)
        G .+= r.ψ
    end

    sweep!(r, r.v)                       # router.jl:104
    _, x = optimizer(fn, g!, r.v, bounds, m=m, factr=factr, pgtol=pgtol,
                     iprint=verbose ? 1 : -1, maxfun=maxfun, maxiter=maxiter)
    r.v .= x
    sweep!(r, r.v; materialize=true)     # router.jl:107
    return nothing
end

# netflows!, src/router.jl:111-119: host-side pool-order sum of the stored trades
function netflows!(ψ, r::B200Router)
    fill!(ψ, 0)
    for (Δ, Λ, c) in zip(r.Δs, r.Λs, r.cfmms)
        ψ[c.Ai] += Λ - Δ
    end
    return nothing
end
netflows(r::B200Router) = (ψ = zero(r.v); netflows!(ψ, r); ψ)

# A working update_reserves!(r) (the reference's, src/router.jl:127-132, calls a per-CFMM
# method that is defined nowhere): R <- R + γΔ − Λ (test/cfmms.jl:10) applied on the device
# from the materialised trades, and mirrored on the host objects with the same expression.
function update_reserves!(r::B200Router)
    for (Δ, Λ, c) in zip(r.Δs, r.Λs, r.cfmms)
        (c isa ProductTwoCoin || c isa GeometricMeanTwoCoin) || continue
        c.R .= c.R .+ c.γ .* Δ .- Λ
    end
    if any(c -> c isa UniV3, r.cfmms)
        sync_reserves!(r)                       # mixed set: push the two-coin reserves
    else
        chk(r.ctx, ccall((:cfmm_apply_trades, LIB), Cint, (Ptr{Cvoid},), r.ctx))
    end
    return nothing
end

# The reference reads cfmm.R live on every sweep; push mutated reserves explicitly.
function sync_reserves!(r::B200Router)
    for (ptype, T) in ((0, ProductTwoCoin), (1, GeometricMeanTwoCoin))
        ids = [i for i in r.order if r.cfmms[i] isa T]
        isempty(ids) && continue
        R = Float64[r.cfmms[i].R[j] for i in ids for j in 1:2]
        chk(r.ctx, ccall((:cfmm_update_reserves, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), r.ctx, ptype, 0, length(ids), R))
    end
end

end # module
