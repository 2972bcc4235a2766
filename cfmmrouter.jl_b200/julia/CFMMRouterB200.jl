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
#     route!(r; optimizer=:device)                       # outer iteration on the GPU too (cfmm_solve)
#     r = B200Router(obj, pools, n; devices=0:7)         # pools sharded over 8 GPUs of this process
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
    ctx::Ptr{Cvoid}         # context of the first device (all of them when there is one)
    order::Vector{Int}      # library insertion order -> index into cfmms (first device's shard first)
    ψ::Vector{T}            # Σ A_i(Λ_i − Δ_i) of the last sweep: a view of PINNED memory (cfmm_host_alloc)
    acc::Base.RefValue{T}   # Σ ν[A_i]ᵀ(Λ_i − Δ_i) of the last sweep
    ctxs::Vector{Ptr{Cvoid}}        # one context per device (multi-GPU: pools sharded in list order)
    shard::Vector{UnitRange{Int}}   # positions of `order` each context holds
    pin::Ptr{Float64}               # pinned staging: ν [n] | ψ [n] | acc [1], per context (n_ctx blocks)
end

# One device context holding the pools cfmms[ids] (positions in the caller's list); returns
# (ctx, order) with order = library insertion order -> position in the caller's list.
function make_context(cfmms::Vector{C}, ids::AbstractVector{Int}, n_tokens, device::Integer) where {T,C<:CFMM{T}}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cfmm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64), out, device, n_tokens)
    rc == 0 || throw(B200Error(rc, unsafe_string(ccall((:cfmm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    ctx = out[]
    order = Int[]
    prod = [i for i in ids if cfmms[i] isa ProductTwoCoin]
    geo = [i for i in ids if cfmms[i] isa GeometricMeanTwoCoin]
    uni = [i for i in ids if cfmms[i] isa UniV3]
    length(prod) + length(geo) + length(uni) == length(ids) ||
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
    return ctx, order
end

# Router(objective, cfmms, n_tokens), src/router.jl:18-35: pack Vector{CFMM} -> SoA, upload.
# devices: one GPU (default) or several of THIS process; the pool list is split into contiguous
# shards, one context per device, and the contexts are joined through the NVLink peer exchange
# (cfmm_comm_export / cfmm_comm_attach, same-process path: raw pointers + peer access), so
# every context's sweep returns the global [ψ; acc].
function B200Router(objective::O, cfmms::Vector{C}, n_tokens; device::Integer=0,
                    devices::AbstractVector{<:Integer}=[device]) where {T,O<:Objective,C<:CFMM{T}}
    T === Float64 || throw(ArgumentError("libcfmm_b200 is fp64-only"))
    W = length(devices)
    m = length(cfmms)
    ctxs = Ptr{Cvoid}[]; order = Int[]; shard = UnitRange{Int}[]
    for (k, dev) in enumerate(devices)
        ids = (div(m * (k - 1), W) + 1):div(m * k, W)
        ctx, ord = make_context(cfmms, collect(ids), n_tokens, dev)
        push!(ctxs, ctx); push!(shard, (length(order) + 1):(length(order) + length(ord))); append!(order, ord)
    end
    if W > 1
        handles = zeros(UInt8, 128 * W)                       # CFMM_COMM_HANDLE_BYTES per context
        for k in 1:W
            GC.@preserve handles chk(ctxs[k], ccall((:cfmm_comm_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}),
                ctxs[k], pointer(handles, 128 * (k - 1) + 1)))
        end
        for k in 1:W
            GC.@preserve handles chk(ctxs[k], ccall((:cfmm_comm_attach, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctxs[k], W, k - 1, handles))
        end
    end
    # pinned staging per context: the library replays {H2D ν, sweep, D2H [ψ; acc]} as one CUDA
    # graph on buffers it has seen twice (cfmm_b200.h, "sweep_graphs"); pageable Vectors cannot
    pin = convert(Ptr{Float64}, ccall((:cfmm_host_alloc, LIB), Ptr{Cvoid}, (Csize_t,), 8 * (2n_tokens + 1) * W))
    pin == C_NULL && throw(OutOfMemoryError())
    ψ = unsafe_wrap(Array, pin + 8n_tokens, n_tokens)         # context 1's ψ block
    Δs = AbstractVector{T}[zeros(T, 2) for _ in cfmms]     # zerotrade, router.jl:23-26
    Λs = AbstractVector{T}[zeros(T, 2) for _ in cfmms]
    r = B200Router{O,T}(objective, convert(Vector{CFMM{T}}, cfmms), Δs, Λs, zeros(T, n_tokens),
                        ctxs[1], order, ψ, Ref(zero(T)), ctxs, shard, pin)
    finalizer(r) do x
        foreach(c -> ccall((:cfmm_destroy, LIB), Cvoid, (Ptr{Cvoid},), c), x.ctxs)
        ccall((:cfmm_host_free, LIB), Cvoid, (Ptr{Cvoid},), x.pin)
    end
    return r
end

# One dual-gradient sweep on the GPU: find_arb!(r, v) (router.jl:38-42) + both
# folds (router.jl:79-83, 98-100).  materialize=true also refreshes r.Δs / r.Λs.
function sweep!(r::B200Router{O,T}, v::AbstractVector{T}; materialize::Bool=false) where {O,T}
    n = length(r.v); W = length(r.ctxs); blk = 2n + 1
    # ν into every context's pinned block, then one blocking cfmm_sweep per context.  With
    # several contexts the calls MUST run concurrently (the fused exchange kernels wait for
    # each other): one task per context on Julia's thread pool (start julia with -t >= W).
    for k in 1:W
        unsafe_copyto!(r.pin + 8blk * (k - 1), pointer(v), n)
    end
    rcs = zeros(Cint, W)
    GC.@preserve v begin
        @sync for k in 1:W
            base = r.pin + 8blk * (k - 1)
            Threads.@spawn rcs[k] = ccall((:cfmm_sweep, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                r.ctxs[k], base, base + 8n, base + 16n, materialize ? 1 : 0)
        end
    end
    foreach(k -> chk(r.ctxs[k], rcs[k]), 1:W)
    r.acc[] = unsafe_load(r.pin, 2n + 1)            # every context holds the bitwise-identical sum
    if materialize
        for k in 1:W
            mk = length(r.shard[k])
            D = Vector{Float64}(undef, 2mk); L = Vector{Float64}(undef, 2mk)
            chk(r.ctxs[k], ccall((:cfmm_get_trades, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), r.ctxs[k], D, L))
            for (j, i) in enumerate(r.order[r.shard[k]])   # library order -> r.cfmms order
                r.Δs[i][1] = D[2j-1]; r.Δs[i][2] = D[2j]
                r.Λs[i][1] = L[2j-1]; r.Λs[i][2] = L[2j]
            end
        end
    end
    return nothing
end

find_arb!(r::B200Router, v) = sweep!(r, collect(Float64, v); materialize=true)

# route!, src/router.jl:58-108, with the three find_arb!(r, v) call sites and the
# two fold loops replaced by sweep!.  Everything else is the reference's code path.
# cfmm_solve_opts / cfmm_solve_info of include/cfmm_b200.h
struct SolveOpts; max_iter::Cint; max_fun::Cint; pgtol::Cdouble; factr::Cdouble; end
mutable struct SolveInfo; iterations::Cint; fun_evals::Cint; status::Cint; f::Cdouble; pg_norm::Cdouble; solve_ms::Cdouble; end

# f(ν) = linᵀν on the box for both objectives of the reference (src/objectives.jl:62-79, 106-129)
linear_term(o::CFMMRouter.LinearNonnegative) = zero(o.c)
linear_term(o::CFMMRouter.BasketLiquidation) = (l = copy(o.Δin); l[o.i] = 0; l)

function route!(r::B200Router; v=nothing, verbose=false, m=5, factr=1e1, pgtol=1e-5, maxfun=15_000, maxiter=15_000,
                optimizer::Symbol=:host)
    if optimizer === :device       # the whole outer iteration on the GPU: only scalars cross PCIe
        length(r.ctxs) == 1 || throw(ArgumentError("optimizer=:device drives one GPU"))
        lin = linear_term(r.objective); lo = lower_limit(r.objective); up = upper_limit(r.objective)
        v0 = isnothing(v) ? C_NULL : pointer(v)
        info = SolveInfo(0, 0, 0, 0.0, 0.0, 0.0); opts = SolveOpts(maxiter, maxfun, pgtol, factr)
        GC.@preserve lin lo up v chk(r.ctx, ccall((:cfmm_solve, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{SolveOpts}, Ptr{Float64}, Ref{SolveInfo}),
            r.ctx, lin, lo, up, v0, opts, r.v, info))
        mk = length(r.order); D = Vector{Float64}(undef, 2mk); L = Vector{Float64}(undef, 2mk)
        chk(r.ctx, ccall((:cfmm_get_trades, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), r.ctx, D, L))
        for (j, i) in enumerate(r.order)
            r.Δs[i] .= (D[2j-1], D[2j]); r.Λs[i] .= (L[2j-1], L[2j])
        end
        return nothing
    end
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
        foreach(c -> chk(c, ccall((:cfmm_apply_trades, LIB), Cint, (Ptr{Cvoid},), c)), r.ctxs)
    end
    return nothing
end

# The reference reads cfmm.R live on every sweep; push mutated reserves explicitly.
function sync_reserves!(r::B200Router)
    for (k, ctx) in enumerate(r.ctxs), (ptype, T) in ((0, ProductTwoCoin), (1, GeometricMeanTwoCoin))
        ids = [i for i in r.order[r.shard[k]] if r.cfmms[i] isa T]
        isempty(ids) && continue
        R = Float64[r.cfmms[i].R[j] for i in ids for j in 1:2]
        chk(ctx, ccall((:cfmm_update_reserves, LIB), Cint,
            (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}), ctx, ptype, 0, length(ids), R))
    end
end

end # module
