# bench/reference_threads.jl — the TRUE reference arm: StatusSwitchingQP.jl's own `solveQP`, Julia-threaded over the batch.
#
# NOT RUN IN THE BUILD ENVIRONMENT (no julia binary, no network; SURVEY.md §8d, BASELINE.md §3).  `bench.py --impl reference`
# times the C++ restatement in reference form (oracle/, kind = "port") instead.  Anyone with Julia >= 1.6 and the package
# installed can produce the genuine number on the same inputs:
#
#     python scripts/dump_config4.py 512 /tmp/c4.bin          # the benchmark's own seeded inputs (numpy PCG64), see below
#     julia -t auto bench/reference_threads.jl /tmp/c4.bin
#
# File layout written by scripts/dump_config4.py (little endian): Int64 N, M, J, nb; then Float64 column-major
# V (N×N), A (M×N), G (J×N), q (N×nb), b (M×nb), g (J×nb), d (N×nb), u (N×nb).
# Output: one JSON line in bench.py's format (metric, value = QPs/s over all threads, cores, sample).
using StatusSwitchingQP, LinearAlgebra, Printf
using Base.Threads

function load(path)
    open(path, "r") do io
        N, M, J, nb = (read(io, Int64) for _ in 1:4)
        rd(r, c) = (A = Matrix{Float64}(undef, r, c); read!(io, A); A)
        V = rd(N, N); A = rd(M, N); G = rd(J, N)
        q = rd(N, nb); b = rd(M, nb); g = rd(J, nb); d = rd(N, nb); u = rd(N, nb)
        return (; N, M, J, nb, V, A, G, q, b, g, d, u)
    end
end

function main(path)
    W = load(path)
    BLAS.set_num_threads(1)                         # one QP per Julia thread, no nested BLAS threading
    mk(i) = QP(W.V; q=W.q[:, i], A=W.A, b=W.b[:, i], G=W.G, g=W.g[:, i], d=W.d[:, i], u=W.u[:, i])
    Qs = [mk(i) for i in 1:W.nb]                    # construction (eigmin(V) per QP, src/types.jl:246) is outside the timed region
    status = zeros(Int, W.nb)
    solveQP(Qs[1])                                  # compile
    t = @elapsed begin
        @threads for i in 1:W.nb
            _, _, st = solveQP(Qs[i])               # src/SSQP.jl:224
            status[i] = st
        end
    end
    @printf("{\"impl\": \"reference\", \"metric\": \"QPs solved/sec (FP64, N=500 batch)\", \"value\": %.6f, \"unit\": \"QPs/s\", ", W.nb / t)
    @printf("\"cpu_baseline\": {\"kind\": \"reference\", \"cores\": %d, \"sample\": \"%d QPs of config 4 from %s\"}, ", nthreads(), W.nb, path)
    @printf("\"solved_ok\": %d, \"trips_per_qp\": %.2f}\n", count(>(0), status), sum(status[status .> 0]) / max(1, count(>(0), status)))
end

main(ARGS[1])
