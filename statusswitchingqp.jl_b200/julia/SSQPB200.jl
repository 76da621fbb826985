# SSQPB200.jl — Julia host glue for libssqp_b200.so (NOT executed in the build environment: no julia binary there).
#
# Keeps the API surface of StatusSwitchingQP.jl (QP, Settings, Status, solveQP) and adds batch methods that reach
# the CUDA kernels through `ccall` only (no CUDA.jl, no kernel codegen, no CPU fallback):
#
#     solveQP(Qs::AbstractVector{QP{Float64}}; settings, settingsLP)  ->  Vector{Tuple{Vector{Float64},Vector{Status},Int}}
#     solveQP_batch(V, A, G, q, b, g, d, u; ...)                      ->  (X, S, status)
#
# Each call below mirrors one prototype of include/ssqp_b200.h; the Python ctypes binding
# (statusswitchingqp.jl_b200/capi.py) makes exactly the same calls and is what the test-suite exercises.
module SSQPB200

using StatusSwitchingQP: QP, LP, Settings, Status, IN, DN, UP, OE, EO
import StatusSwitchingQP: solveQP, SimplexLP
import StatusSwitchingQP
import LinearAlgebra

const LIB = get(ENV, "SSQP_B200_LIB", joinpath(@__DIR__, "..", "libssqp_b200.so"))

# mirrors ssqp_settings (include/ssqp_b200.h) == Settings{Float64} (src/types.jl:390-408)
struct CSettings
    max_iter::Int32
    tol::Float64
    tolG::Float64
    rule::Int32
    pivot::Int32
end
CSettings(s::Settings{Float64}) = CSettings(Int32(s.maxIter), s.tol, s.tolG,
    Int32(s.rule == :Dantzig ? 0 : s.rule == :stpEdgeLP ? 1 : s.rule == :maxImprovement ? 2 : 0), Int32(0))   # the rule symbols initQP / SimplexLP test for (src/SSQP.jl:477-481)

mutable struct Context
    h::Ptr{Cvoid}
    N::Int; M::Int; J::Int
end

function check(ctx, rc, what)
    rc == 0 && return
    msg = unsafe_string(ccall((:ssqp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx === nothing ? C_NULL : ctx.h))
    error("$what failed ($rc): $msg  (libssqp_b200 has no CPU fallback)")
end

"Context(devices = [0]) — one per host thread; sharding over the listed CUDA devices is by QP index."
function Context(devices::Vector{<:Integer}=[0])
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ids = Int32.(devices)
    rc = ccall((:ssqp_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Int32}, Int32), h, ids, length(ids))
    check(nothing, rc, "ssqp_create")
    ctx = Context(h[], 0, 0, 0)
    finalizer(c -> (c.h != C_NULL && ccall((:ssqp_destroy, LIB), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
    return ctx
end

"Upload V (N×N), A (M×N), G (J×N) — Julia matrices are already column-major, no copy besides H2D."
function set_shared!(ctx::Context, V::Matrix{Float64}, A::Matrix{Float64}, G::Matrix{Float64})
    N = size(V, 1); M = size(A, 1); J = size(G, 1)
    rc = ccall((:ssqp_set_shared, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.h, N, M, J, V, M > 0 ? pointer(A) : C_NULL, J > 0 ? pointer(G) : C_NULL)
    check(ctx, rc, "ssqp_set_shared")
    ctx.N, ctx.M, ctx.J = N, M, J
    return ctx
end

"""
    solveQP_batch(ctx, q, b, g, d, u; settings, settingsLP, S0=nothing, x0=nothing) -> (X, S, status)

q,d,u: N×nb; b: M×nb; g: J×nb (one column per QP).  X: N×nb, S: (N+J)×nb Matrix{Status}, status: Vector{Int}
with the meaning of solveQP's third return value (src/SSQP.jl:224-377).
"""
function solveQP_batch(ctx::Context, q::Matrix{Float64}, b::Matrix{Float64}, g::Matrix{Float64},
        d::Matrix{Float64}, u::Matrix{Float64};
        settings=Settings{Float64}(), settingsLP=settings,
        S0::Union{Nothing,Matrix{Status}}=nothing, x0::Union{Nothing,Matrix{Float64}}=nothing)
    N, M, J = ctx.N, ctx.M, ctx.J
    nb = size(q, 2)
    X = Matrix{Float64}(undef, N, nb)
    S = Matrix{Status}(undef, N + J, nb)          # @enum Status has an Int32 base: bit-compatible with int32_t*
    status = Vector{Int64}(undef, nb)
    st = Ref(CSettings(settings)); stlp = Ref(CSettings(settingsLP))
    rc = ccall((:ssqp_solve_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Status}, Ptr{Float64}, Ref{CSettings}, Ref{CSettings}, Ptr{Float64}, Ptr{Status}, Ptr{Int64}),
        ctx.h, nb, C_NULL, q, M > 0 ? pointer(b) : C_NULL, J > 0 ? pointer(g) : C_NULL, d, u,
        S0 === nothing ? C_NULL : pointer(S0), x0 === nothing ? C_NULL : pointer(x0), st, stlp, X, S, status)
    check(ctx, rc, "ssqp_solve_batch")
    return X, S, status
end

"""
    SimplexLP(Ps::AbstractVector{LP{Float64}}; settings, ctx) -> Vector of (x, S, status)

Drop-in for `[SimplexLP(P; settings) for P in Ps]` (src/Simplex.jl:831) when the LPs share A and G.
"""
function SimplexLP(Ps::AbstractVector{LP{Float64}}; settings=Settings{Float64}(), ctx::Context=default_context())
    isempty(Ps) && return Tuple{Vector{Float64},Vector{Status},Int}[]
    P = first(Ps)
    all(Q -> Q.A == P.A && Q.G == P.G, Ps) || error("a device batch must share A and G; split the list")
    N, M, J = P.N, P.M, P.J
    # The device wants [A 0; G I] of full row rank.  The reference's redundancy purge (src/Simplex.jl:889-902) stays on this
    # side of the C ABI: rank-deficient inputs are purged per LP with the reference's own getRowsGJr, grouped by the rows
    # kept, and each group goes to the device as its own batch.  (Same logic as solver.py: SimplexLP_batch.)
    if M > 0
        slack(Q) = begin
            iv = findall((Q.u .== Inf) .& (Q.d .== -Inf)); id = findall((Q.d .== -Inf) .& (Q.u .< Inf))
            A0 = [Q.A zeros(M, J) -Q.A[:, iv]; Q.G Matrix{Float64}(LinearAlgebra.I, J, J) -Q.G[:, iv]]
            A0[:, id] .= -A0[:, id]
            A0
        end
        if LinearAlgebra.rank(slack(P)) < M + J || any(Q -> (Q.d .== -Inf) != (P.d .== -Inf) || (Q.u .== Inf) != (P.u .== Inf), Ps)
            res = Vector{Tuple{Vector{Float64},Vector{Status},Int}}(undef, length(Ps))
            groups = Dict{Vector{Int},Vector{Int}}()
            for (t, Q) in enumerate(Ps)
                A0 = slack(Q); m0 = LinearAlgebra.rank(A0)
                if m0 == M + J
                    push!(get!(groups, collect(1:M), Int[]), t); continue
                end
                ra, la = StatusSwitchingQP.getRowsGJr([A0 [Q.b; Q.g]], settings.tol)
                if length(ra) != la
                    res[t] = (zeros(N), fill(DN, N), 0)
                elseif m0 != la || any(r -> r > M && !(r in ra), 1:M+J)
                    res[t] = (zeros(N), fill(DN, N), -1)
                else
                    push!(get!(groups, filter(r -> r <= M, ra), Int[]), t)
                end
            end
            for (rows, ts) in groups
                sub = [LP(Ps[t].c, Ps[t].A[rows, :], Ps[t].b[rows]; d=Ps[t].d, u=Ps[t].u, G=Ps[t].G, g=Ps[t].g) for t in ts]
                length(rows) == M || (res[ts] .= SimplexLP(sub; settings=settings, ctx=ctx); continue)
                res[ts] .= _simplexlp_device(sub, settings, ctx)
            end
            return res
        end
    end
    return _simplexlp_device(Ps, settings, ctx)
end

function _simplexlp_device(Ps, settings, ctx)
    P = first(Ps)
    N, M, J = P.N, P.M, P.J
    rc = ccall((:ssqp_set_shared, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.h, N, M, J, C_NULL, M > 0 ? pointer(P.A) : C_NULL, J > 0 ? pointer(P.G) : C_NULL)
    check(ctx, rc, "ssqp_set_shared"); ctx.N, ctx.M, ctx.J = N, M, J
    nb = length(Ps)
    cat(f) = reduce(hcat, (f(Q) for Q in Ps))
    c, b, g, d, u = cat(Q -> Q.c), cat(Q -> Q.b), cat(Q -> Q.g), cat(Q -> Q.d), cat(Q -> Q.u)
    X = Matrix{Float64}(undef, N, nb); S = Matrix{Status}(undef, N + J, nb); status = Vector{Int64}(undef, nb)
    st = Ref(CSettings(settings))
    rc = ccall((:ssqp_solve_lp_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{CSettings},
         Ptr{Float64}, Ptr{Status}, Ptr{Int64}),
        ctx.h, nb, c, M > 0 ? pointer(b) : C_NULL, J > 0 ? pointer(g) : C_NULL, d, u, st, X, S, status)
    check(ctx, rc, "ssqp_solve_lp_batch")
    return [(X[:, t], S[:, t], Int(status[t])) for t in 1:nb]
end

const _ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = (_ctx[] === nothing && (_ctx[] = Context([0])); _ctx[])

"""
    solveQP(Qs::AbstractVector{QP{Float64}}; settings, settingsLP, ctx) -> Vector of (z, S, status)

Drop-in for `[solveQP(Q; settings, settingsLP) for Q in Qs]` when the QPs share V, A and G (e.g. the frontier
sweeps built with `QP(P, q, L)` / `QP(P, mu, q)`, src/types.jl:303-339).  QPs with `mc <= 0` get the reference's early
return (src/SSQP.jl:226-228) without touching the device.
"""
function solveQP(Qs::AbstractVector{QP{Float64}}; settings=Settings{Float64}(), settingsLP=settings,
        ctx::Context=default_context())
    isempty(Qs) && return Tuple{Vector{Float64},Vector{Status},Int}[]
    P = first(Qs)
    all(Q -> Q.V === P.V && Q.A == P.A && Q.G == P.G, Qs) ||
        error("a device batch must share V, A and G; split the list")
    set_shared!(ctx, P.V, P.A, P.G)
    good = findall(Q -> Q.mc > 0, Qs)
    res = Vector{Tuple{Vector{Float64},Vector{Status},Int}}(undef, length(Qs))
    for (i, Q) in enumerate(Qs)
        Q.mc <= 0 && (res[i] = (zeros(Q.N), fill(DN, Q.N), -1))
    end
    if !isempty(good)
        cat(f) = reduce(hcat, (f(Qs[i]) for i in good))
        X, S, status = solveQP_batch(ctx, cat(Q -> Q.q), cat(Q -> Q.b), cat(Q -> Q.g), cat(Q -> Q.d), cat(Q -> Q.u);
            settings=settings, settingsLP=settingsLP)
        for (t, i) in enumerate(good)
            res[i] = (X[:, t], S[:, t], Int(status[t]))
        end
    end
    return res
end

"""
    solveQP_sweep(Qs::AbstractVector{QP{Float64}}; chain_len=32, settings, settingsLP, ctx) -> Vector of (z, S, status)

The warm-started loop `z,S,st = solveQP(Qs[1]); for Q in Qs[2:end]; z,S,st = solveQP(Q,S,z); end` (src/SSQP.jl:237) for
chains of `chain_len` consecutive QPs that share V, A, G, b, g, d, u (e.g. `[QP(P, E, L) for L in Ls]`, src/types.jl:303-319),
the chains in parallel on the device (`ssqp_solve_sweep`).  `length(Qs)` must be a multiple of `chain_len`.
"""
function solveQP_sweep(Qs::AbstractVector{QP{Float64}}; chain_len::Integer=32, settings=Settings{Float64}(), settingsLP=settings,
        ctx::Context=default_context())
    isempty(Qs) && return Tuple{Vector{Float64},Vector{Status},Int}[]
    P = first(Qs)
    all(Q -> Q.V === P.V && Q.A == P.A && Q.G == P.G && Q.mc > 0, Qs) || error("a sweep must share V, A and G (and be valid QPs)")
    set_shared!(ctx, P.V, P.A, P.G)
    N, M, J = ctx.N, ctx.M, ctx.J
    nb = length(Qs)
    cat(f) = reduce(hcat, (f(Q) for Q in Qs))
    q, b, g, d, u = cat(Q -> Q.q), cat(Q -> Q.b), cat(Q -> Q.g), cat(Q -> Q.d), cat(Q -> Q.u)
    X = Matrix{Float64}(undef, N, nb); S = Matrix{Status}(undef, N + J, nb); status = Vector{Int64}(undef, nb)
    st = Ref(CSettings(settings)); stlp = Ref(CSettings(settingsLP))
    rc = ccall((:ssqp_solve_sweep, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ref{CSettings}, Ref{CSettings}, Ptr{Float64}, Ptr{Status}, Ptr{Int64}),
        ctx.h, nb, chain_len, C_NULL, q, M > 0 ? pointer(b) : C_NULL, J > 0 ? pointer(g) : C_NULL, d, u, st, stlp, X, S, status)
    check(ctx, rc, "ssqp_solve_sweep")
    return [(X[:, t], S[:, t], Int(status[t])) for t in 1:nb]
end

end # module
