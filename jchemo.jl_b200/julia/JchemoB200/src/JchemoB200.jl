"""
    JchemoB200

Drop-in for the kernel-PLS path of Jchemo.jl (`plskern`, `plskern!`, `transform`, `coef`, `predict` on
`Plsr`; reference: `src/plskern.jl:1-238`).  Same function names, argument meaning, returned fields and
field order; the arithmetic runs on an NVIDIA B200 in `libjchemo_b200.so` (hand-written sm_100a CUDA)
through the C ABI declared in `include/jchemo_b200.h`.  There is no CPU fallback: without the library
or a B200 every call raises.

The library path is taken from `ENV["JCHEMO_B200_LIB"]`, else `libjchemo_b200.so` next to this package.

NOTE: this file could not be executed in the build container (no Julia toolchain there, nor on the GPU boxes);
its `ccall` signatures are checked statically against the header (tests/test_abi.py) and the same ABI is
exercised end to end by the Python ctypes mirror in `jchemo.jl_b200/plskern.py` and `tests/`.
"""
module JchemoB200

using LinearAlgebra
using Libdl

export Plsr, plskern, plskern!, transform, coef, predict, gridscorelv, gridcvlv, locwlv, xfit, xfit!, xresid, xresid!
export pin!, unpin!, resident, resident_add, resident_drop, last_fit_info

const LIB = get(ENV, "JCHEMO_B200_LIB",
                normpath(joinpath(@__DIR__, "..", "..", "..", "libjchemo_b200.so")))

# ---------------------------------------------------------------- struct (src/plskern.jl:1-14)
struct Plsr
    T::Matrix{Float64}
    P::Matrix{Float64}
    R::Matrix{Float64}
    W::Matrix{Float64}
    C::Matrix{Float64}
    TT::Vector{Float64}
    xmeans::Vector{Float64}
    xscales::Vector{Float64}
    ymeans::Vector{Float64}
    yscales::Vector{Float64}
    weights::Vector{Float64}
    niter::Union{Array{Float64}, Nothing}
end

# `V` is the north-star spelling of the X-loadings `P`
Base.getproperty(o::Plsr, s::Symbol) = s === :V ? getfield(o, :P) : getfield(o, s)
Base.propertynames(::Plsr) = (fieldnames(Plsr)..., :V)

# ---------------------------------------------------------------- helpers (src/utility.jl)
ensure_mat(X::AbstractMatrix) = X                                   # utility.jl:544
ensure_mat(X::AbstractVector) = Matrix(reshape(X, :, 1))            # :545
ensure_mat(X::Number) = reshape([X], 1, 1)                          # :546
ensure_mat(X::LinearAlgebra.Adjoint) = Matrix(X)                    # :547
ensure_mat(X) = Matrix(X)                                           # DataFrame etc. (:548)
nro(X) = size(X, 1)
nco(X) = size(X, 2)

dense64(X) = X isa Matrix{Float64} ? X : Matrix{Float64}(X)

function check(rc::Cint, what::AbstractString)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:jcb200_last_error, LIB), Cstring, ()))
    # JCB200_ENONFINITE: the reference fails inside LinearAlgebra.svd with this very exception (plskern.jl:154)
    rc == -5 && throw(ArgumentError("matrix contains Infs or NaNs ($what: $msg)"))
    error("JchemoB200.$what failed (status $rc): $msg")
end

# ---------------------------------------------------------------- host memory the GPU can reach at PCIe speed
# Large OUTPUTS (the scores T, the normalised weights) are taken from the library's page-locked pool
# (jcb200_host_alloc): the device-to-host copy then runs at link speed instead of through the pageable staging
# path.  They are ordinary `Matrix{Float64}` / `Vector{Float64}` for every consumer; the block goes back to the
# pool when the array is garbage collected.  Small arrays and a failed pool allocation fall back to `undef` arrays.
const PINNED_MIN_BYTES = 1 << 22
function pooled(dims::Int...)
    bytes = 8 * prod(dims)
    if bytes >= PINNED_MIN_BYTES
        ptr = ccall((:jcb200_host_alloc, LIB), Ptr{Cvoid}, (Int64,), bytes)
        if ptr != C_NULL
            A = unsafe_wrap(Array, Ptr{Float64}(ptr), dims; own = false)
            finalizer(_ -> ccall((:jcb200_host_free, LIB), Cint, (Ptr{Cvoid},), ptr), A)
            return A
        end
    end
    Array{Float64}(undef, dims...)
end

"""
    pin!(X) ; unpin!(X)

Page-lock a caller's `Array{Float64}` in place (`jcb200_host_register`): later fits / predictions on it copy at
full PCIe speed (C2: 81 ms instead of 98 ms end to end).  Worth it for an X that is used more than once;
`unpin!` before the array is freed.
"""
pin!(X::Array{Float64}) = (check(ccall((:jcb200_host_register, LIB), Cint, (Ptr{Cvoid}, Int64), X, sizeof(X)), "pin!"); X)
unpin!(X::Array{Float64}) = (check(ccall((:jcb200_host_unregister, LIB), Cint, (Ptr{Cvoid},), X), "unpin!"); X)

"""
    resident_add(X) ; resident_drop(X) ; resident(f, X, Y, ...)

Device-resident data handle (`jcb200_resident_add`): upload a matrix once; every later call that is handed this
very array as `X` (or `Y`) skips the host-to-device transfer — `plskern`, `summary`, `transform`, `predict`,
`gridscorelv`, `gridcvlv`, `xfit`.  Do not modify the array on the host while it is resident.

    resident(X, Y) do
        fm = plskern(X, Y; nlv = 25)
        summary(fm, X)
        gridscorelv(X, Y, X, Y; score = :rmsep, nlv = 0:25)
    end
"""
resident_add(X::Matrix{Float64}) =
    (check(ccall((:jcb200_resident_add, LIB), Cint, (Ptr{Float64}, Int64, Int64, Int64), X, max(nro(X), 1), nro(X), nco(X)),
           "resident_add"); X)
resident_drop(X::Matrix{Float64}) = (ccall((:jcb200_resident_drop, LIB), Cint, (Ptr{Float64},), X); nothing)
function resident(f::Function, mats::Matrix{Float64}...)
    done = Matrix{Float64}[]
    try
        for A in mats
            resident_add(A)
            push!(done, A)
        end
        return f()
    finally
        foreach(resident_drop, done)
    end
end

"""
    last_fit_info()

`(nlv_effective = k,)` for the calling thread's last fit: LVs that carry information (`TT[a] > 0` and `C[:, a] != 0`).  Where the
reference divides 0/0 (constant y, more LVs than the data carry; plskern.jl:152,166) this library returns inert
LVs — finite model, predictions equal to those of the last informative LV.
"""
function last_fit_info()
    k = Ref{Int32}(0)
    check(ccall((:jcb200_last_fit_info, LIB), Cint, (Ref{Int32},), k), "last_fit_info")
    (nlv_effective = Int(k[]),)
end

"""
    init_multi(devices)

Bind the library to several GPUs of one box (e.g. `init_multi(0:7)`); `plskern` / `plskern!` then shard
the rows over them inside the library (`jcb200_init_multi`).  Call once, before any fit.
"""
function init_multi(devices)
    ids = Cint.(collect(devices))
    check(ccall((:jcb200_init_multi, LIB), Cint, (Cint, Ptr{Cint}), length(ids), ids), "init_multi")
end

# ---------------------------------------------------------------- fit (src/plskern.jl:106-178)
function _fit(X::Matrix{Float64}, Y::Matrix{Float64}, weights, nlv::Integer, scal::Bool, writeback::Bool)
    n, p = size(X)
    q = nco(Y)
    nro(Y) == n || throw(DimensionMismatch("X has $n rows, Y has $(nro(Y))"))
    a = max(0, min(n, p, nlv))                                      # :116
    w = weights === nothing ? nothing : Vector{Float64}(vec(weights))
    w === nothing || length(w) == n || throw(DimensionMismatch("weights has length $(length(w))"))
    T = pooled(n, a); P = Matrix{Float64}(undef, p, a)
    R = Matrix{Float64}(undef, p, a); W = Matrix{Float64}(undef, p, a)
    C = Matrix{Float64}(undef, q, a); TT = Vector{Float64}(undef, a)
    xmeans = Vector{Float64}(undef, p); xscales = Vector{Float64}(undef, p)
    ymeans = Vector{Float64}(undef, q); yscales = Vector{Float64}(undef, q)
    wout = pooled(n)
    nlv_out = Ref{Int32}(0)
    rc = ccall((:jcb200_plskern_fit, LIB), Cint,
               (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int32, Int32,
                Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ref{Int32}),
               X, max(n, 1), Y, max(n, 1), w === nothing ? C_NULL : w, n, p, q, nlv, scal, writeback,
               T, max(n, 1), P, R, W, C, TT, xmeans, xscales, ymeans, yscales, wout, nlv_out)
    check(rc, "plskern")
    Plsr(T, P, R, W, C, TT, xmeans, xscales, ymeans, yscales, wout, nothing)
end

"""
    plskern(X, Y, weights = ones(nro(X)); nlv, scal = false)

Same as `Jchemo.plskern` (src/plskern.jl:106-110): the inputs are left untouched.  No host copy of X
is made (the reference's `copy` is replaced by the host-to-device transfer).
"""
function plskern(X, Y, weights = nothing; nlv, scal = false)
    _fit(dense64(ensure_mat(X)), dense64(ensure_mat(Y)), weights, nlv, scal, false)
end

"""
    plskern!(X::Matrix, Y::Matrix, weights = ones(nro(X)); nlv, scal = false)

Same as `Jchemo.plskern!` (src/plskern.jl:112-178): X and Y leave centred (and scaled) in place.
"""
function plskern!(X::Matrix{Float64}, Y::Matrix{Float64}, weights = nothing; nlv, scal = false)
    _fit(X, Y, weights, nlv, scal, true)
end
# The reference's signature is `plskern!(X::Matrix, Y::Matrix, ...)`; with a non-Float64 element type its in-place
# centring (`center!`, utility.jl:76-81) throws an InexactError (integers) or computes in another precision.  The
# device path is Float64 only: say so instead of converting behind the caller's back (the side effect could not
# reach the caller's array).
plskern!(X::Matrix, Y::Matrix, weights = nothing; nlv, scal = false) =
    throw(ArgumentError("plskern! needs Matrix{Float64} arguments (got $(eltype(X)), $(eltype(Y))); use plskern"))

# ---------------------------------------------------------------- transform (src/plskern.jl:187-195)
function transform(object::Plsr, X; nlv = nothing)
    X = dense64(ensure_mat(X))
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = min(nlv, a)
    nlv = max(nlv, 0)
    m, p = size(X)
    p == nro(object.R) || throw(DimensionMismatch("X has $p columns, the model has $(nro(object.R))"))
    T = pooled(m, nlv)
    (nlv == 0 || m == 0) && return T
    rc = ccall((:jcb200_transform, LIB), Cint,
               (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
                Ptr{Float64}, Int64),
               X, m, m, p, object.xmeans, object.xscales, object.R, nlv, T, m)
    check(rc, "transform")
    T
end

# ---------------------------------------------------------------- coef (src/plskern.jl:207-217)
function coef(object::Plsr; nlv = nothing)
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = min(nlv, a)
    nlv = max(nlv, 0)
    p = nro(object.R); q = nro(object.C)
    B = Matrix{Float64}(undef, p, q)
    int = Matrix{Float64}(undef, 1, q)
    rc = ccall((:jcb200_coef, LIB), Cint,
               (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64,
                Int64, Int32, Ptr{Float64}, Ptr{Float64}),
               a == 0 ? C_NULL : object.R, a == 0 ? C_NULL : object.C, object.xmeans, object.xscales,
               object.ymeans, object.yscales, p, q, nlv, B, int)
    check(rc, "coef")
    (B = B, int = int)
end

# ---------------------------------------------------------------- predict (src/plskern.jl:226-238)
function predict(object::Plsr, X; nlv = nothing)
    X = dense64(ensure_mat(X))
    a = nco(object.T)
    isnothing(nlv) ? nlv = a : nlv = (max(0, minimum(nlv)):min(a, maximum(nlv)))   # :229
    le_nlv = length(nlv)
    m, p = size(X)
    q = nro(object.C)
    p == length(object.xmeans) || throw(DimensionMismatch("X has $p columns"))
    pred = [pooled(m, q) for _ in 1:le_nlv]
    if le_nlv > 0 && m > 0
        ptrs = [pointer(z) for z in pred]
        GC.@preserve pred begin
            rc = ccall((:jcb200_predict_sweep, LIB), Cint,
                       (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Int32,
                        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32,
                        Ptr{Ptr{Float64}}),
                       X, m, m, p, q, a == 0 ? C_NULL : object.R, a == 0 ? C_NULL : object.C, a,
                       object.xmeans, object.xscales, object.ymeans, object.yscales,
                       first(nlv), last(nlv), ptrs)
        end
        check(rc, "predict")
    end
    le_nlv == 1 ? (pred = pred[1],) : (pred = pred,)                                # :236-237
end

# ---------------------------------------------------------------- summary (src/plskern.jl:246-260)
"""
    summary(object::Plsr, X)

Explained X-variance per LV, as `Jchemo`'s `Base.summary(::Plsr, X)`: returns
`(explvarx = (nlv, var, pvar, cumpvar),)` (NamedTuple of columns; wrap in `DataFrame` as needed).
"""
function Base.summary(object::Plsr, X)       # the reference takes Union{Matrix, DataFrame} (plskern.jl:246)
    X = dense64(ensure_mat(X))
    n, a = size(object.T)
    p = nco(X)
    xvar = Vector{Float64}(undef, a); pvar = similar(xvar); cumpvar = similar(xvar)
    rc = ccall((:jcb200_summary, LIB), Cint,
               (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               X, n, n, p, object.xmeans, object.xscales, object.weights, a == 0 ? C_NULL : object.P,
               a == 0 ? C_NULL : object.TT, a, xvar, pvar, cumpvar)
    check(rc, "summary")
    (explvarx = (nlv = collect(1:a), var = xvar, pvar = pvar, cumpvar = cumpvar),)
end

# ---------------------------------------------------------------- gridscorelv (src/gridscore.jl:167-221)
"""
    gridscorelv(Xtrain, Ytrain, X, Y; score, nlv, kwargs...)

`Jchemo.gridscorelv` for `fun = plskern` (branch `pars === nothing`): one fit with `maximum(nlv)` LVs,
then the residual sums of the validation set for every nlv in ONE pass on the GPU
(`jcb200_gridscore`); `score` is one of `:msep, :rmsep, :ssr, :bias, :sep, :r2, :rpd`
(src/scores.jl).  Returns a NamedTuple of columns `nlv, y1, ..., yq` (wrap in `DataFrame` as needed).
"""
function gridscorelv(Xtrain, Ytrain, X, Y; score::Symbol, nlv, kwargs...)
    Xtrain = dense64(ensure_mat(Xtrain)); Ytrain = dense64(ensure_mat(Ytrain))
    X = dense64(ensure_mat(X)); Y = dense64(ensure_mat(Y))
    p = nco(Xtrain)
    lo = max(0, minimum(nlv)); hi = min(p, maximum(nlv))                 # :170-173
    fm = plskern(Xtrain, Ytrain; nlv = hi, kwargs...)                    # :179
    a = nco(fm.T); hi = min(hi, a); nk = hi - lo + 1
    m = nro(X); q = nco(Y)
    ssr = Matrix{Float64}(undef, nk, q); sres = similar(ssr)
    ysum = Vector{Float64}(undef, q); ysumsq = similar(ysum)
    rc = ccall((:jcb200_gridscore, LIB), Cint,
               (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64},
                Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               X, m, Y, m, m, p, q, a == 0 ? C_NULL : fm.R, a == 0 ? C_NULL : fm.C, a, fm.xmeans,
               fm.xscales, fm.ymeans, fm.yscales, lo, hi, ssr, sres, ysum, ysumsq)
    check(rc, "gridscorelv")
    ms = ssr ./ m; bi = -sres ./ m
    vary = (ysumsq ./ m .- (ysum ./ m) .^ 2)'
    res = score === :msep ? ms : score === :rmsep ? sqrt.(ms) : score === :ssr ? ssr :
          score === :bias ? bi : score === :sep ? sqrt.(ms .- bi .^ 2) :
          score === :r2 ? 1 .- ms ./ vary : score === :rpd ? sqrt.(vary) ./ sqrt.(ms) :
          error("score must be one of :msep, :rmsep, :ssr, :bias, :sep, :r2, :rpd")
    cols = (; nlv = collect(lo:hi), (Symbol("y", j) => res[:, j] for j in 1:q)...)
    cols
end

# ---------------------------------------------------------------- gridcvlv (src/gridcv.jl:187-228)
"""
    gridcvlv(X, Y; segm, score, nlv, scal = false)

`Jchemo.gridcvlv` for `fun = plskern` (branch `pars === nothing`).  `segm` as built by `segmkf` / `segmts`
(a vector of repetitions, each a vector of 1-based row-index vectors, disjoint within a repetition).
Every repetition is one `jcb200_gridcv` call: Gram down-dating on the GPU instead of K fits on K
row-copies.  Returns `(res = ..., res_rep = ...)` as NamedTuples of columns.
"""
function gridcvlv(X, Y; segm, score::Symbol, nlv, scal = false)
    X = dense64(ensure_mat(X)); Y = dense64(ensure_mat(Y))
    n, p = size(X); q = nco(Y)
    lo = max(0, minimum(nlv)); hi = min(p, maximum(nlv)); nk = hi - lo + 1      # :193
    repl = Int[]; sg = Int[]; ks = Int[]; ys = [Float64[] for _ in 1:q]
    for (i, listsegm) in enumerate(segm)
        K = length(listsegm)
        allidx = reduce(vcat, listsegm)
        rest = setdiff(1:n, allidx)
        perm = Int64.(vcat(allidx, rest)) .- 1
        seg_start = Int64.(vcat(0, cumsum(length.(listsegm))))
        ssr = Array{Float64}(undef, nk, q, K); sres = similar(ssr)
        ysum = Matrix{Float64}(undef, q, K); ysumsq = similar(ysum)
        rc = ccall((:jcb200_gridcv, LIB), Cint,
                   (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64},
                    Int32, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   X, n, Y, n, n, p, q, perm, seg_start, K, lo, hi, scal, i > 1, ssr, sres, ysum, ysumsq)
        check(rc, "gridcvlv")
        for j in 1:K
            m = length(listsegm[j])
            ms = ssr[:, :, j] ./ m; bi = -sres[:, :, j] ./ m
            vary = (ysumsq[:, j] ./ m .- (ysum[:, j] ./ m) .^ 2)'
            tab = score === :msep ? ms : score === :rmsep ? sqrt.(ms) : score === :ssr ? ssr[:, :, j] :
                  score === :bias ? bi : score === :sep ? sqrt.(ms .- bi .^ 2) :
                  score === :r2 ? 1 .- ms ./ vary : score === :rpd ? sqrt.(vary) ./ sqrt.(ms) :
                  error("score must be one of :msep, :rmsep, :ssr, :bias, :sep, :r2, :rpd")
            for t in 1:nk
                push!(repl, i); push!(sg, j); push!(ks, lo + t - 1)
                for c in 1:q
                    push!(ys[c], tab[t, c])
                end
            end
        end
    end
    res_rep = (; repl = repl, segm = sg, nlv = ks, (Symbol("y", c) => ys[c] for c in 1:q)...)
    uk = sort(unique(ks))
    res = (; nlv = uk, (Symbol("y", c) => [sum(ys[c][ks .== k]) / count(ks .== k) for k in uk] for c in 1:q)...)
    (res = res, res_rep = res_rep)
end

# ---------------------------------------------------------------- xfit / xresid (src/xfit.jl:33-99)
function _xfit!(object::Plsr, X::Matrix{Float64}, out::Matrix{Float64}, nlv, resid::Bool)
    a = size(object.T, 2)
    nlv = isnothing(nlv) ? a : max(min(nlv, a), 0)
    m, p = size(X)
    p == size(object.R, 1) || throw(DimensionMismatch("X has $p columns, the model has $(size(object.R, 1))"))
    m == 0 && return out
    rc = ccall((:jcb200_xfit, LIB), Cint,
               (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Int32, Int32, Ptr{Float64}, Int64),
               X, m, m, p, object.xmeans, object.xscales, nlv == 0 ? C_NULL : object.R,
               nlv == 0 ? C_NULL : object.P, nlv, resid, out, m)
    check(rc, "xfit")
    out
end
xfit(object::Plsr, X; nlv = nothing) = (Z = dense64(ensure_mat(X)); _xfit!(object, Z, similar(Z), nlv, false))
xfit!(object::Plsr, X::Matrix{Float64}; nlv = nothing) = _xfit!(object, X, X, nlv, false)
xresid(object::Plsr, X; nlv = nothing) = (Z = dense64(ensure_mat(X)); _xfit!(object, Z, similar(Z), nlv, true))
xresid!(object::Plsr, X::Matrix{Float64}; nlv = nothing) = _xfit!(object, X, X, nlv, true)

# ---------------------------------------------------------------- locwlv (src/locwlv.jl:9-48)
"""
    locwlv(Xtrain, Ytrain, X; listnn, listw = nothing, nlv, scal = false)

`Jchemo.locwlv` for `fun = plskern`: every row of `X` is predicted by a weighted kernel-PLS model fitted on
its neighbours; all `m` tiny fits run in one kernel launch (`jcb200_locw_plskern`).
"""
function locwlv(Xtrain, Ytrain, X; listnn, listw = nothing, nlv, scal = false, verbose = false)
    Xtrain = dense64(ensure_mat(Xtrain)); Ytrain = dense64(ensure_mat(Ytrain)); X = dense64(ensure_mat(X))
    ntr, p = size(Xtrain); q = nco(Ytrain); m = nro(X)
    lo = max(0, minimum(nlv)); hi = min(p, maximum(nlv)); nk = hi - lo + 1        # :14
    for i in 1:m
        k = length(listnn[i])
        if min(k, p) < hi && !(q == 1 && length(unique(Ytrain[listnn[i], :])) == 1)
            throw(DimensionMismatch("neighbourhood $i has $k rows, fewer than nlv = $hi"))   # as :37
        end
    end
    idx = Int64.(reduce(vcat, [collect(s) for s in listnn])) .- 1
    off = Int64.(vcat(0, cumsum([length(s) for s in listnn])))
    w = isnothing(listw) ? C_NULL : Float64.(reduce(vcat, [collect(v) for v in listw]))
    zpred = Array{Float64}(undef, m, q, nk)
    rc = ccall((:jcb200_locw_plskern, LIB), Cint,
               (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64,
                Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}),
               Xtrain, ntr, Ytrain, ntr, ntr, p, q, X, m, m, idx, off, w, lo, hi, scal, zpred)
    check(rc, "locwlv")
    pred = [zpred[:, :, a] for a in 1:nk]
    (pred = nk == 1 ? pred[1] : pred,)
end

end # module
