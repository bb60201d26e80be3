# TEST INFRASTRUCTURE — runs the UNMODIFIED reference (Jchemo.jl's src/utility.jl + src/plskern.jl) under Julia.
#
#   julia oracle/julia_ref.jl <reference root> time   n p q nlv reps warmup     -> one JSON line with timings
#   julia oracle/julia_ref.jl <reference root> golden <out dir>                  -> fixtures for tests/golden/
#
# Neither Julia nor the reference tree exists on the build image or the GPU boxes of this project (SURVEY.md §0), so
# this script has NEVER been executed: it is the hook that turns "parity unpinned" green the first time a machine
# has both.  bench.py (`cpu_baseline.kind = "reference"`) and tests/test_oracle.py::test_julia_reference_fixtures
# look for `julia` on PATH and for the tree at $JCHEMO_REFERENCE (default /root/reference) and call it when found.
# The reference's two files need only LinearAlgebra, Statistics and the DataFrame TYPE (ensure_mat(::DataFrame),
# utility.jl:548); when DataFrames.jl is not installed a stand-in type keeps the method definitions loadable.
using LinearAlgebra, Statistics, Printf
try
    @eval using DataFrames
catch
    @eval struct DataFrame end
end
try
    @eval using StatsBase
catch
end

const REF = ARGS[1]
include(joinpath(REF, "src", "utility.jl"))      # mweight, colmean, colstd, center!, cscale!, ensure_mat, vcol, list ...
include(joinpath(REF, "src", "plskern.jl"))      # struct Plsr, plskern, plskern!, transform, coef, predict

# the counter-based generator of SURVEY.md §8(d) / oracle/synth.py (bit-identical)
@inline function mix64(z::UInt64)
    z ⊻= z >> 30; z *= 0xBF58476D1CE4E5B9
    z ⊻= z >> 27; z *= 0x94D049BB133111EB
    z ⊻= z >> 31
    z
end
u01(seed, k) = Float64(mix64(UInt64(seed) * 0xD1342543DE82EF95 + (UInt64(k) + 0x1) * 0x9E3779B97F4A7C15) >> 11) * 2.0^-53
function synth_matrix(seed, n, p)
    X = Matrix{Float64}(undef, n, p)
    Threads.@threads for j in 1:p
        for i in 1:n
            X[i, j] = u01(seed, (i - 1) + (j - 1) * n)
        end
    end
    X
end
synth_weights(n) = [0.5 + u01(3, i - 1) for i in 1:n]

function run_time(n, p, q, nlv, reps, warmup)
    X = synth_matrix(1, n, p); Y = synth_matrix(2, n, q)
    ts = Float64[]
    for i in 1:(warmup + reps)
        Xc = copy(X); Yc = copy(Y)
        t = @elapsed plskern!(Xc, Yc; nlv = nlv)
        i > warmup && push!(ts, t)
    end
    @printf("{\"seconds_mean\": %.6f, \"seconds_min\": %.6f, \"julia\": \"%s\", \"blas_threads\": %d, \"threads\": %d}\n",
            mean(ts), minimum(ts), string(VERSION), BLAS.get_num_threads(), Threads.nthreads())
end

# fixtures: raw little-endian Float64 column-major files + a manifest; tests/test_oracle.py reads them
function dump(dir, name, A)
    open(joinpath(dir, name * ".f64"), "w") do io
        write(io, Float64.(vec(collect(A))))
    end
    size(A)
end
function run_golden(dir)
    mkpath(dir)
    cases = Dict("c1" => (150, 200, 2, 5, 50, true, false), "c1_wscal" => (150, 200, 2, 5, 50, false, true),
                 "c2_cut" => (20000, 500, 10, 25, 1000, true, false), "c3_cut" => (20000, 1000, 1, 30, 1000, false, true),
                 "edge_odd" => (37, 5, 3, 10, 7, false, false))
    open(joinpath(dir, "manifest.txt"), "w") do man
        for (name, (n, p, q, nlv, m, uniform, scal)) in cases
            X = synth_matrix(1, n, p); Y = synth_matrix(2, n, q); Xnew = synth_matrix(4, m, p)
            w = uniform ? ones(n) : synth_weights(n)
            fm = plskern(X, Y, w; nlv = nlv, scal = scal)
            for (f, A) in (("T", fm.T), ("P", fm.P), ("R", fm.R), ("W", fm.W), ("C", fm.C), ("TT", fm.TT),
                           ("xmeans", fm.xmeans), ("xscales", fm.xscales), ("ymeans", fm.ymeans),
                           ("yscales", fm.yscales), ("weights", fm.weights), ("B", coef(fm).B),
                           ("int", coef(fm).int), ("pred", predict(fm, Xnew).pred), ("Tnew", transform(fm, Xnew)))
                sz = dump(dir, name * "_" * f, A)
                println(man, name, " ", f, " ", join(sz, "x"))
            end
        end
    end
end

if ARGS[2] == "time"
    run_time(parse.(Int, ARGS[3:8])...)
elseif ARGS[2] == "golden"
    run_golden(ARGS[3])
end
