# SSQPB200.jl — Julia host glue for libssqp_b200.so (NOT executed in the build environment: no julia binary there).
#
# Keeps the API surface of StatusSwitchingQP.jl (QP, Settings, Status, solveQP) and adds batch methods that reach
# the CUDA kernels through `ccall` only (no CUDA.jl, no kernel codegen, no CPU fallback):
#
#     solveQP(Qs::AbstractVector{QP{Float64}}; settings, settingsLP)  ->  Vector{Tuple{Vector{Float64},Vector{Status},Int}}
#     solveQP_batch(V, A, G, q, b, g, d, u; ...)                      ->  (X, S, status)
#     optimize_batch!(opts::Vector{Optimizer{Float64}})               ->  the MOI wrapper's optimize! for many models at once
#     use_device!(true)                                               ->  MOI.optimize!(::Optimizer) itself goes to the device
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
    all(Q -> (Q.V === P.V || Q.V == P.V) && Q.A == P.A && Q.G == P.G, Qs) ||      # equal matrices built independently are fine
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
    all(Q -> (Q.V === P.V || Q.V == P.V) && Q.A == P.A && Q.G == P.G && Q.mc > 0, Qs) || error("a sweep must share V, A and G (and be valid QPs)")
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

"""
    SimplexLP(cs::AbstractVector{Vector{Float64}}, A, b, d, u; settings, min, ctx) -> Vector of (x, S, status)

Batch twin of the reference's array form `SimplexLP(c, A, b, d, u; settings, min)` (src/Simplex.jl:1036-1196): LPs
`min c'x s.t. Ax = b, d <= x <= u` that share A, b, d, u and differ in the cost vector.  It is the struct form without the
slack block (J = 0), so it goes through the same device entry.
"""
function SimplexLP(cs::AbstractVector{Vector{Float64}}, A, b, d, u; settings=Settings{Float64}(), min=true, ctx::Context=default_context())
    sgn = min ? 1.0 : -1.0
    return SimplexLP([LP(sgn .* c, Matrix{Float64}(A), Vector{Float64}(b); d=Vector{Float64}(d), u=Vector{Float64}(u)) for c in cs];
                     settings=settings, ctx=ctx)
end

# ---- MathOptInterface hook (src/MOIwrapper.jl:131-171) ------------------------------------------------------------------
# `MOI.optimize!(opt::Optimizer)` of the reference solves ONE model with the scalar solveQP / SimplexLP (:165,:167).  Two ways
# to the device from JuMP / MOI:
#   optimize_batch!(opts)   — many optimizers (e.g. `backend(model)` of many JuMP models over one covariance matrix) in one
#                             device batch per group sharing (V, A, G); fills opt.Results / opt.solTime exactly as optimize! does,
#                             so TerminationStatus (:213-228, bug-compatible mapping included), ObjectiveValue, VariablePrimal
#                             keep working unchanged;
#   use_device!(true)       — overrides MOI.optimize!(::Optimizer{Float64}) so that every single optimize! is a batch of one.
# The `mc == -20` presolve branches (:133-158) never reach the solver in the reference and stay on the host here.
import MathOptInterface as MOI
const SSQ = StatusSwitchingQP

function _presolved!(opt::SSQ.Optimizer{Float64})
    P = opt.Problem
    P.mc == -20 || return false
    N = P.N
    if P.M > 0
        opt.Results = (P.A \ P.b, fill(DN, N), 1)
    elseif P isa QP
        x = P.V \ P.q
        dt = LinearAlgebra.det(P.V)
        st = ((opt.Sense == MOI.MIN_SENSE && dt > 0) || (opt.Sense == MOI.MAX_SENSE && dt < 0)) ? 1 : 3
        opt.Results = (x, fill(DN, N), st)
    else
        opt.Results = (zeros(N), fill(DN, N), LinearAlgebra.norm(P.c, Inf) == 0 ? 1 : 3)
    end
    return true
end

"""
    optimize_batch!(opts::AbstractVector{<:StatusSwitchingQP.Optimizer{Float64}}; ctx)

`MOI.optimize!` for a list of optimizers through the device path: models are grouped by kind (QP / LP), shape and shared
(V, A, G) and Settings; each group is one `ssqp_solve_batch` / `ssqp_solve_lp_batch` call.
"""
function optimize_batch!(opts::AbstractVector{<:SSQ.Optimizer{Float64}}; ctx::Context=default_context())
    t0 = time()
    groups = Vector{Vector{Int}}()
    for (i, o) in enumerate(opts)
        _presolved!(o) && continue
        P = o.Problem
        k = findfirst(groups) do g
            K = opts[g[1]].Problem; S0 = opts[g[1]].Settings
            typeof(K) == typeof(P) && (K.N, K.M, K.J) == (P.N, P.M, P.J) && K.A == P.A && K.G == P.G &&
                (!(P isa QP) || K.V === P.V || K.V == P.V) &&
                (S0.maxIter, S0.tol, S0.tolG, S0.rule) == (o.Settings.maxIter, o.Settings.tol, o.Settings.tolG, o.Settings.rule)
        end
        k === nothing ? push!(groups, [i]) : push!(groups[k], i)
    end
    for g in groups
        st = opts[g[1]].Settings
        res = opts[g[1]].Problem isa QP ? solveQP([opts[i].Problem for i in g]; settings=st, ctx=ctx) :
                                          SimplexLP([opts[i].Problem for i in g]; settings=st, ctx=ctx)
        for (t, i) in enumerate(g)
            opts[i].Results = res[t]
        end
    end
    dt = time() - t0
    for o in opts
        o.solTime = dt
    end
    return opts
end

"use_device!(on): route every `MOI.optimize!(::Optimizer{Float64})` through libssqp_b200 (a batch of one); `false` restores the CPU path."
function use_device!(on::Bool=true)
    if on
        @eval MOI.optimize!(opt::SSQ.Optimizer{Float64}) = (optimize_batch!([opt]); nothing)
    else
        @eval function MOI.optimize!(opt::SSQ.Optimizer{Float64})       # the reference's body (src/MOIwrapper.jl:131-171)
            _presolved!(opt) && return nothing
            t0 = time()
            opt.Results = opt.Problem isa QP ? SSQ.solveQP(opt.Problem; settings=opt.Settings) : SSQ.SimplexLP(opt.Problem; settings=opt.Settings)
            opt.solTime = time() - t0
            nothing
        end
    end
    return on
end

end # module
