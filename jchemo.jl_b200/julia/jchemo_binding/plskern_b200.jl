# src/plskern_b200.jl — the binding a Jchemo.jl maintainer adds (INTEGRATION.md).
#
# Include it in src/Jchemo.jl right AFTER `include("plskern.jl")` (src/Jchemo.jl:19-165 holds the include list).
# It keeps Jchemo's own `struct Plsr` (src/plskern.jl:1-14) — every typed consumer keeps working
# (`fm::Plsr` in src/plsravg_unif.jl:2, src/plsravg_aic.jl:2,8, src/plsravg_shenk.jl:2; `Union{Pcr, Plsr}` in
# src/vip.jl:62, src/xfit.jl:33, src/wshenk.jl:47, src/occsd.jl:129, src/occod.jl:43) — and adds MORE SPECIFIC
# methods that route the Float64 path to libjchemo_b200.so:
#
#   plskern!(X::Matrix{Float64}, Y::Matrix{Float64}, ...)   more specific than plskern!(X::Matrix, Y::Matrix, ...) (:112)
#   plskern(X, Y, ...)                                      replaces :106-110 (same signature: this definition wins)
#   transform(object::Plsr, X; nlv)                         replaces :187-195 (same signature)
#   coef(object::Plsr; nlv)                                 more specific than coef(::Union{Plsr, Pcr}) (:207)
#   predict(object::Plsr, X; nlv)                           more specific than predict(::Union{Plsr, Pcr}, X) (:226)
#
# `Pcr` objects are untouched: `coef(::Pcr)` / `predict(::Pcr, X)` still dispatch to the Union methods of
# src/plskern.jl:207,226 on the CPU (Julia picks the most specific applicable method, so there is no ambiguity).
# Non-Float64 matrices passed to `plskern!` keep hitting the reference's own method (:112) by dispatch.
#
# The file is an OPT-IN build of the package, not a second backend behind a switch: src/Jchemo.jl includes it only
# when the user asked for the B200 path,
#       haskey(ENV, "JCHEMO_B200_LIB") && include("plskern_b200.jl")
# and from then on there is no CPU fallback: if the library or a B200 is missing, loading the package fails in
# `__init_b200__`, and a failing call raises.
#
# Not executed in the build image (no Julia there); every `ccall` signature in this file is checked against
# include/jchemo_b200.h by tests/test_abi.py.
const LIBB200 = ENV["JCHEMO_B200_LIB"]

function __init_b200__()            # call from Jchemo.__init__()
    rc = ccall((:jcb200_init, LIBB200), Cint, (Cint,), parse(Cint, get(ENV, "JCHEMO_B200_DEVICE", "0")))
    b200_check(rc, "jcb200_init")
    ids = get(ENV, "JCHEMO_B200_DEVICES", "")                  # e.g. "0,1,2,3,4,5,6,7": shard fits over several GPUs
    if !isempty(ids)
        v = parse.(Cint, split(ids, ","))
        ccall((:jcb200_shutdown, LIBB200), Cvoid, ())
        b200_check(ccall((:jcb200_init_multi, LIBB200), Cint, (Cint, Ptr{Cint}), length(v), v), "jcb200_init_multi")
    end
end

function b200_check(rc::Cint, what)
    rc == 0 && return
    msg = unsafe_string(ccall((:jcb200_last_error, LIBB200), Cstring, ()))
    rc == -5 && throw(ArgumentError("matrix contains Infs or NaNs ($what)"))     # what svd throws at plskern.jl:154
    error("$what: $msg (status $rc)")
end

# large outputs from the library's page-locked pool: the scores come back at PCIe speed; still a Matrix{Float64}
function b200_array(dims::Int...)
    bytes = 8 * prod(dims)
    if bytes >= (1 << 22)
        ptr = ccall((:jcb200_host_alloc, LIBB200), Ptr{Cvoid}, (Int64,), bytes)
        if ptr != C_NULL
            A = unsafe_wrap(Array, Ptr{Float64}(ptr), dims; own = false)
            finalizer(_ -> ccall((:jcb200_host_free, LIBB200), Cint, (Ptr{Cvoid},), ptr), A)
            return A
        end
    end
    Array{Float64}(undef, dims...)
end

function b200_fit(X::Matrix{Float64}, Y::Matrix{Float64}, weights, nlv, scal, writeback)
    n, p = size(X); q = nco(Y)
    nro(Y) == n || throw(DimensionMismatch("X has $n rows, Y has $(nro(Y))"))
    a = max(0, min(n, p, nlv))                                                   # plskern.jl:116
    w = Vector{Float64}(vec(weights))                                            # utility.jl:716
    length(w) == n || throw(DimensionMismatch("weights has length $(length(w)), X has $n rows"))
    T = b200_array(n, a); P = Matrix{Float64}(undef, p, a); R = similar(P); W = similar(P)
    C = Matrix{Float64}(undef, q, a); TT = Vector{Float64}(undef, a)
    xmeans = Vector{Float64}(undef, p); xscales = similar(xmeans)
    ymeans = Vector{Float64}(undef, q); yscales = similar(ymeans)
    wout = b200_array(n); nlv_out = Ref{Int32}(0)
    rc = ccall((:jcb200_plskern_fit, LIBB200), Cint,
        (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int32, Int32, Int32,
         Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
        X, max(n, 1), Y, max(n, 1), w, n, p, q, nlv, scal, writeback,
        T, max(n, 1), P, R, W, C, TT, xmeans, xscales, ymeans, yscales, wout, nlv_out)
    b200_check(rc, "plskern")
    Plsr(T, P, R, W, C, TT, xmeans, xscales, ymeans, yscales, wout, nothing)     # Jchemo's own struct, :176-177
end

# ---- fit: src/plskern.jl:106-178
function plskern!(X::Matrix{Float64}, Y::Matrix{Float64}, weights = ones(nro(X)); nlv, scal = false)
    b200_fit(X, Y, weights, nlv, scal, #=writeback_xy: X, Y leave centred / scaled, :125-129=# true)
end

function plskern(X, Y, weights = ones(nro(X)); nlv, scal = false)
    # no host copy: the device copy replaces `copy(ensure_mat(X))` (:108); views / DataFrames are densified
    Xd = ensure_mat(X); Yd = ensure_mat(Y)
    Xd = Xd isa Matrix{Float64} ? Xd : Matrix{Float64}(Xd)
    Yd = Yd isa Matrix{Float64} ? Yd : Matrix{Float64}(Yd)
    b200_fit(Xd, Yd, weights, nlv, scal, false)
end

# ---- transform: src/plskern.jl:187-195
function transform(object::Plsr, X; nlv = nothing)
    X = ensure_mat(X)
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = min(nlv, a)
    nlv = max(nlv, 0)
    (nlv == 0 || nro(X) == 0) && return Matrix{Float64}(undef, nro(X), nlv)
    X = X isa Matrix{Float64} ? X : Matrix{Float64}(X)
    m, p = size(X)
    p == nro(object.R) || throw(DimensionMismatch("X has $p columns, the model has $(nro(object.R))"))
    T = b200_array(m, nlv)
    rc = ccall((:jcb200_transform, LIBB200), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int64),
        X, m, m, p, object.xmeans, object.xscales, object.R, nlv, T, m)
    b200_check(rc, "transform")
    T
end

# ---- coef: src/plskern.jl:207-217 (Plsr only; Pcr keeps the Union method)
function coef(object::Plsr; nlv = nothing)
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = min(nlv, a)
    nlv = max(nlv, 0)
    p = nro(object.R); q = nro(object.C)
    B = Matrix{Float64}(undef, p, q); int = Matrix{Float64}(undef, 1, q)
    rc = ccall((:jcb200_coef, LIBB200), Cint,
        (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int32,
         Ptr{Float64}, Ptr{Float64}),
        a == 0 ? C_NULL : object.R, a == 0 ? C_NULL : object.C, object.xmeans, object.xscales,
        object.ymeans, object.yscales, p, q, nlv, B, int)
    b200_check(rc, "coef")
    (B = B, int = int)
end

# ---- predict: src/plskern.jl:226-238 (Plsr only) — one pass over X for the whole nlv range
function predict(object::Plsr, X; nlv = nothing)
    X = ensure_mat(X)
    X = X isa Matrix{Float64} ? X : Matrix{Float64}(X)
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = (max(0, minimum(nlv)):min(a, maximum(nlv)))   # :229
    le_nlv = length(nlv)
    m, p = size(X); q = nro(object.C)
    p == length(object.xmeans) || throw(DimensionMismatch("X has $p columns, the model has $(length(object.xmeans))"))
    pred = list(le_nlv, Matrix{Float64})                                           # :231
    for i = 1:le_nlv
        pred[i] = b200_array(m, q)
    end
    if le_nlv > 0 && m > 0
        ptrs = [pointer(z) for z in pred]
        GC.@preserve pred begin
            rc = ccall((:jcb200_predict_sweep, LIBB200), Cint,
                (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Float64}}),
                X, m, m, p, q, a == 0 ? C_NULL : object.R, a == 0 ? C_NULL : object.C, a, object.xmeans,
                object.xscales, object.ymeans, object.yscales, first(nlv), last(nlv), ptrs)
        end
        b200_check(rc, "predict")
    end
    le_nlv == 1 ? pred = pred[1] : nothing                                         # :236
    (pred = pred,)
end
