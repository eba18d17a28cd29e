# VoiceConversionB200.jl -- drop-in Julia shim for the spectral-conversion hot path of
# r9y9/VoiceConversion.jl over libvcb200.so (C ABI: include/vcb200.h).
#
# Same exported names, argument meaning and exceptions as the reference for this path
# (reference src/VoiceConversion.jl:12-38); every method body is argument checks + one `ccall`.
# The reference is Julia-0.5 syntax (`immutable`, `type`, `Array(Float64, ...)`) which no current
# Julia parses; this file is Julia 1.x.  Julia is not installed in the build image, so this shim
# is exercised through its line-for-line Python twin (voiceconversion.jl_b200/__init__.py) --
# both bind exactly the same symbols with the same argument order.
#
#   using VoiceConversionB200            # instead of `using VoiceConversion`
#   mapper = GMMMap(gmm["weights"], gmm["means"], gmm["covars"])
#   mapper = TrajectoryGMMMap(mapper, 100)
#   converted = vc(mapper, src)          # bin/vc.jl:82 works unchanged
module VoiceConversionB200

using Libdl
using SparseArrays
import Base: length, size

export FrameByFrameConverter, TrajectoryConverter, GMMMapParam, GMMMap, TrajectoryGMMMap,
       TrajectoryGVGMMMap, VarianceScaling, fvpostf, fvpostf!, diffgmm, vc_static,
       fvconvert, vc, ncomponents, dim, push_delta, align

const libvcb200 = get(ENV, "LIBVCB200", joinpath(@__DIR__, "..", "libvcb200.so"))

# ---- status codes -> the reference's exception types -------------------------------------------
const VCB_OK, VCB_EDIM, VCB_ENOTPD, VCB_ESINGULAR, VCB_EARG, VCB_ENOMEM, VCB_ECUDA, VCB_EUNSUPPORTED = 0:7

function last_error()
    buf = Vector{UInt8}(undef, 1024)
    ccall((:vcb_last_error, libvcb200), Int32, (Ptr{UInt8}, Csize_t), buf, length(buf))
    unsafe_string(pointer(buf))
end

function check(rc::Int32)
    rc == VCB_OK && return nothing
    msg = last_error()
    rc == VCB_EDIM && throw(DimensionMismatch(msg))                 # src/gmmmap.jl:102
    rc == VCB_ENOTPD && throw(LinearAlgebra.PosDefException(0))     # MvNormal, src/gmm.jl:17
    rc == VCB_ESINGULAR && throw(LinearAlgebra.SingularException(0)) # `^-1`, src/gmmmap.jl:35
    rc == VCB_EARG && throw(ArgumentError(msg))
    rc == VCB_ENOMEM && throw(OutOfMemoryError())
    error("libvcb200: $msg (status $rc)")
end
import LinearAlgebra

# one process per GPU: call once with the local rank
set_device(dev::Integer) = check(ccall((:vcb_set_device, libvcb200), Int32, (Int32,), dev))

# ONE Julia process driving the whole box: after init(0) the batch methods below -- vc(mapper, fms),
# fvconvert(g, X), DTWs.fit!(d, templates, toff, sequences, soff), align(srcs, tgts) -- shard their batch
# over all visible GPUs inside the library (one host thread and copy pipeline per device, model replicated)
function init(ndev::Integer=0)
    check(ccall((:vcb_init, libvcb200), Int32, (Int32,), ndev))
    n = Ref{Int32}(0)
    check(ccall((:vcb_num_devices, libvcb200), Int32, (Ref{Int32},), n))
    Int(n[])
end

# Page-locked arrays: host<->device copies of ordinary (pageable) Julia arrays go through the driver's staging
# buffers at a fraction of PCIe speed and serialise the library's copy/compute pipelines.  `pinned_matrix`
# returns a Matrix{Float64} in page-locked memory (freed by its finalizer); `pin!` / `unpin!` page-lock a
# long-lived array in place.
function pinned_matrix(rows::Integer, cols::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vcb_host_alloc, libvcb200), Int32, (Ref{Ptr{Cvoid}}, Csize_t), p, rows * cols * sizeof(Float64)))
    A = unsafe_wrap(Array, Ptr{Float64}(p[]), (Int(rows), Int(cols)); own=false)
    finalizer(_ -> ccall((:vcb_host_free, libvcb200), Int32, (Ptr{Cvoid},), p[]), A)
    A
end
pin!(A::Array{Float64}) = (check(ccall((:vcb_host_register, libvcb200), Int32, (Ptr{Cvoid}, Csize_t), A, sizeof(A))); A)
unpin!(A::Array{Float64}) = (check(ccall((:vcb_host_unregister, libvcb200), Int32, (Ptr{Cvoid},), A)); A)

# ---- type hierarchy (src/common.jl:2-4) -----------------------------------------------------------
abstract type AbstractConverter end
abstract type FrameByFrameConverter <: AbstractConverter end
abstract type TrajectoryConverter <: AbstractConverter end

# ---- GMMMapParam / GMMMap (src/gmmmap.jl:10-96) -----------------------------------------------------
struct GMMMapParam
    weights::Vector{Float64}
    μˣ::Matrix{Float64}
    μʸ::Matrix{Float64}
    Σˣˣ::Array{Float64,3}
    Σˣʸ::Array{Float64,3}
    Σʸˣ::Array{Float64,3}
    Σʸʸ::Array{Float64,3}
    ΣʸˣΣˣˣ⁻¹::Array{Float64,3}
end

mutable struct GMMMap <: FrameByFrameConverter
    handle::Ptr{Cvoid}
    D::Int
    M::Int
    params::GMMMapParam

    function GMMMap(weights::Vector{Float64}, μ::Matrix{Float64}, Σ::Array{Float64,3}; swap::Bool=false)
        twoD, M = size(μ)
        (size(Σ) == (twoD, twoD, M) && length(weights) == M) ||
            throw(DimensionMismatch("Inconsistent dimentions."))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vcb_gmmmap_create, libvcb200), Int32,
                    (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                    weights, μ, Σ, twoD, M, swap, h))
        D = twoD >> 1
        get(which, dims...) = (a = Array{Float64}(undef, dims...);
            check(ccall((:vcb_gmmmap_get_param, libvcb200), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), h[], which, a)); a)
        p = GMMMapParam(get(7, M), get(0, D, M), get(1, D, M), get(3, D, D, M), get(4, D, D, M),
                        get(5, D, D, M), get(6, D, D, M), get(2, D, D, M))
        g = new(h[], D, M, p)
        finalizer(x -> ccall((:vcb_gmmmap_destroy, libvcb200), Int32, (Ptr{Cvoid},), x.handle), g)
        g
    end
end

# model files: `load(path)` of JLD / JLD2 returns a Dict with the keys of bin/train_gmm.jl:106-113
# ("weights", "means", "covars", "diff", "n_components"); bin/vc.jl:49-54 passes three of them on
GMMMap(gmm::AbstractDict; swap::Bool=false) =
    GMMMap(Vector{Float64}(gmm["weights"]), Matrix{Float64}(gmm["means"]), Array{Float64,3}(gmm["covars"]); swap=swap)

length(g::GMMMap) = 1                      # src/gmmmap.jl:93
dim(g::GMMMap) = g.D                       # :94
ncomponents(g::GMMMap) = g.M               # :95
size(g::GMMMap) = (dim(g), length(g))      # :96

# fvconvert(g, x)  (src/gmmmap.jl:101-118); the matrix method converts T frames in one call
function fvconvert(g::GMMMap, x::Vector{Float64})
    y = Vector{Float64}(undef, dim(g))
    check(ccall((:vcb_gmmmap_convert, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int64, Ptr{Float64}, Int64),
                g.handle, x, length(x), 1, length(x), y, dim(g)))
    y
end

function fvconvert(g::GMMMap, X::Matrix{Float64})
    Y = Matrix{Float64}(undef, dim(g), size(X, 2))
    check(ccall((:vcb_gmmmap_convert, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int64, Ptr{Float64}, Int64),
                g.handle, X, size(X, 1), size(X, 2), size(X, 1), Y, dim(g)))
    Y
end

# vc(c::FrameByFrameConverter, fm)  (src/common.jl:7-26)
function vc(c::GMMMap, fm::AbstractMatrix{Float64})
    fmd = fm isa Matrix{Float64} ? fm : Matrix{Float64}(fm)
    converted = similar(fmd)
    check(ccall((:vcb_gmmmap_vc, libvcb200), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Float64}),
                c.handle, fmd, size(fmd, 1), size(fmd, 2), converted))
    converted
end

# ---- TrajectoryGMMMap (src/trajectory_gmmmap.jl:3-110) ----------------------------------------------
mutable struct TrajectoryGMMMap <: TrajectoryConverter
    gmmmap::GMMMap
    handle::Ptr{Cvoid}
    T::Int                     # number of column blocks of the (never materialised) W, :34
    Eʸ::Vector{Float64}        # kept for the GV variant, :90-91
    Dʸ::Array{Float64,3}       # :24-28

    function TrajectoryGMMMap(g::GMMMap, T::Int)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vcb_traj_create, libvcb200), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), g.handle, h))
        Dy = Array{Float64}(undef, dim(g), dim(g), ncomponents(g))
        check(ccall((:vcb_traj_get_Dy, libvcb200), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], Dy))
        t = new(g, h[], T, zeros(0), Dy)
        finalizer(x -> ccall((:vcb_traj_destroy, libvcb200), Int32, (Ptr{Cvoid},), x.handle), t)
        t
    end
end

# constructW(D, T) (src/trajectory_gmmmap.jl:39-61): the sparse (2DT x DT) window matrix.  The library never
# builds it (R = W' D^-1 W is assembled on the fly); it exists because the reference exposes and tests it
# (test/trajectory_gmmmap.jl:1-34, 53).
function constructW(D::Int, T::Int)
    I_, J_, V_ = Int[], Int[], Float64[]
    for t in 1:T
        r0 = 2D * (t - 1)
        for k in 1:D
            push!(I_, r0 + k); push!(J_, (t - 1) * D + k); push!(V_, 1.0)                  # static: I at (t, t)
            if t >= 2
                push!(I_, r0 + D + k); push!(J_, (t - 2) * D + k); push!(V_, -0.5)         # delta: -1/2 I at (t, t-1)
            end
            if t < T
                push!(I_, r0 + D + k); push!(J_, t * D + k); push!(V_, 0.5)                # delta: +1/2 I at (t, t+1)
            end
        end
    end
    sparse(I_, J_, V_, 2D * T, D * T)
end

length(t::TrajectoryGMMMap) = t.T                  # :34
dim(t::TrajectoryGMMMap) = dim(t.gmmmap)           # :35
ncomponents(t::TrajectoryGMMMap) = ncomponents(t.gmmmap)
size(t::TrajectoryGMMMap) = (dim(t), length(t))

function fvconvert(tgmm::TrajectoryGMMMap, X::Matrix{Float64})
    rows, T = size(X)
    Y = Matrix{Float64}(undef, rows >> 1, T)
    Ey = Matrix{Float64}(undef, rows, T)
    off = Int64[0, T]
    check(ccall((:vcb_traj_convert_batch, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}),
                tgmm.handle, X, rows, rows, off, 1, 0, Y, rows >> 1, C_NULL, Ey))
    tgmm.T = T                 # W is rebuilt for the new length and stays (:70-72)
    tgmm.Eʸ = vec(Ey)
    Y
end

# vc(c::TrajectoryConverter, fm)  (src/common.jl:31-63): chunks of length(c) frames, read once
function vc(c::TrajectoryGMMMap, fm::AbstractMatrix{Float64})
    vc(c, [fm isa Matrix{Float64} ? fm : Matrix{Float64}(fm)])[1]
end

# batch extension: every utterance is converted with the chunk limit length(c) in ONE library call
function vc(c::TrajectoryGMMMap, fms::Vector{Matrix{Float64}})
    limit = length(c)
    rows = size(fms[1], 1)
    all(m -> size(m, 1) == rows, fms) || throw(DimensionMismatch("Inconsistent dimentions."))
    off = Int64[0; cumsum(size.(fms, 2))]
    fm = length(fms) == 1 ? fms[1] : hcat(fms...)
    Dout = ((rows - 1) >> 1) + 1
    out = Matrix{Float64}(undef, Dout, size(fm, 2))
    check(ccall((:vcb_traj_vc_batch, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Int64}, Int64, Int32, Ptr{Float64}),
                c.handle, fm, rows, off, length(fms), limit, out))
    Tlast = size(fms[end], 2)
    if Tlast > 0
        r = Tlast % limit
        c.T = r == 0 ? min(limit, Tlast) : r       # length of the last chunk solved (quirk Q3)
    end
    [out[:, off[i]+1:off[i+1]] for i in 1:length(fms)]
end

# bin/vc.jl:76-82 in one call: fms hold the power row and the STATIC features only, (1+Ds, T_s); the
# delta rows ([static; delta] per utterance, push_delta's boundary rule) are appended on the device
function vc_static(c::TrajectoryGMMMap, fms::Vector{Matrix{Float64}})
    limit = length(c)
    rows = size(fms[1], 1)
    all(m -> size(m, 1) == rows, fms) || throw(DimensionMismatch("Inconsistent dimentions."))
    off = Int64[0; cumsum(size.(fms, 2))]
    fm = length(fms) == 1 ? fms[1] : hcat(fms...)
    out = Matrix{Float64}(undef, rows, size(fm, 2))
    check(ccall((:vcb_traj_vc_static_batch, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Int64}, Int64, Int32, Ptr{Float64}),
                c.handle, fm, rows, off, length(fms), limit, out))
    [out[:, off[i]+1:off[i+1]] for i in 1:length(fms)]
end

# ---- TrajectoryGVGMMMap (src/trajectory_gmmmap.jl:112-189) ----------------------------------------
mutable struct TrajectoryGVGMMMap <: TrajectoryConverter
    tgmm::TrajectoryGMMMap
    μᵛ::Vector{Float64}
    Σᵛᵛ::Matrix{Float64}
    handle::Ptr{Cvoid}

    function TrajectoryGVGMMMap(tgmm::TrajectoryGMMMap, μᵛ::Vector{Float64}, Σᵛᵛ::Matrix{Float64})
        h = Ref{Ptr{Cvoid}}(C_NULL)      # the library asserts μᵛ >= 0 (:124) and inverts Σᵛᵛ (:125)
        check(ccall((:vcb_trajgv_create, libvcb200), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                    tgmm.handle, μᵛ, Σᵛᵛ, h))
        t = new(tgmm, μᵛ, Σᵛᵛ, h[])
        finalizer(x -> ccall((:vcb_trajgv_destroy, libvcb200), Int32, (Ptr{Cvoid},), x.handle), t)
        t
    end
end

length(t::TrajectoryGVGMMMap) = length(t.tgmm)      # :129
dim(t::TrajectoryGVGMMMap) = dim(t.tgmm)            # :130
ncomponents(t::TrajectoryGVGMMMap) = ncomponents(t.tgmm)

function fvconvert(tgv::TrajectoryGVGMMMap, X::Matrix{Float64}; epochs::Int=100, α::Float64=1.0e-5, verbose::Bool=false)
    rows, T = size(X)
    Y = Matrix{Float64}(undef, rows >> 1, T)
    check(ccall((:vcb_trajgv_convert_batch, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Int64}, Int64, Int32, Int32, Float64, Ptr{Float64}, Int64),
                tgv.handle, X, rows, rows, Int64[0, T], 1, 0, epochs, α, Y, rows >> 1))
    tgv.tgmm.T = T
    Y
end

function vc(c::TrajectoryGVGMMMap, fms::Vector{Matrix{Float64}})
    limit = length(c)
    rows = size(fms[1], 1)
    off = Int64[0; cumsum(size.(fms, 2))]
    fm = length(fms) == 1 ? fms[1] : hcat(fms...)
    out = Matrix{Float64}(undef, ((rows - 1) >> 1) + 1, size(fm, 2))
    check(ccall((:vcb_trajgv_vc_batch, libvcb200), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Int64}, Int64, Int32, Int32, Float64, Ptr{Float64}),
                c.handle, fm, rows, off, length(fms), limit, 100, 1.0e-5, out))
    Tlast = size(fms[end], 2)
    if Tlast > 0                                   # fvconvert(tgv, X) rebuilds tgmm.W for the last chunk (:153-156)
        r = Tlast % limit
        c.tgmm.T = r == 0 ? min(limit, Tlast) : r
    end
    [out[:, off[i]+1:off[i+1]] for i in 1:length(fms)]
end
vc(c::TrajectoryGVGMMMap, fm::AbstractMatrix{Float64}) = vc(c, [Matrix{Float64}(fm)])[1]

# ---- GV post filter (src/gv.jl:6-21) and differential model (src/diffgmm.jl:9-25) ----------------
struct VarianceScaling
    σ²::Vector{Float64}
end

function fvpostf!(vs::VarianceScaling, src::Matrix{Float64})
    D, T = size(src)
    length(vs.σ²) == D || throw(DimensionMismatch("VarianceScaling has $(length(vs.σ²)) entries, src has $D rows"))
    check(ccall((:vcb_variance_scaling_batch, libvcb200), Int32,
                (Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Float64}, Int64),
                vs.σ², D, src, D, Int64[0, T], 1, src, D))
    src
end
fvpostf(vs::VarianceScaling, src::AbstractMatrix) = fvpostf!(vs, Matrix{Float64}(copy(src)))

# joint (μ, Σ) of the differential model; GMMMap(weights, diffgmm(μ, Σ)...) is the reference's
# GMMMapParam returned by diffgmm(params)
function diffgmm(μ::Matrix{Float64}, Σ::Array{Float64,3})
    μd, Σd = similar(μ), similar(Σ)
    check(ccall((:vcb_diffgmm, libvcb200), Int32, (Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}),
                μ, Σ, size(μ, 1), size(μ, 2), μd, Σd))
    μd, Σd
end

# ---- DTWs (src/dtw.jl) ------------------------------------------------------------------------------
module DTWs
import ..libvcb200, ..check
export DTW, fit!, update!, set_template!, backward

mutable struct DTW
    fstep::Int
    bstep::Int
    template::Matrix{Float64}
    costtable::Matrix{Float64}
    backpointer::Matrix{Int}
end
DTW(; fstep=0, bstep=1) = DTW(fstep, bstep, zeros(1, 1), zeros(1, 1), zeros(Int, 1, 1))

# fit!(d, template, sequence) (src/dtw.jl:93-128).  The fused kernel keeps the cost table on the chip, so
# d.costtable / d.backpointer are NOT filled (the reference leaves its S x (T+1) tables behind, :122-123);
# tables=true rebuilds them exactly through update! (one library call per frame) for callers that read them.
function fit!(d::DTW, template::Matrix{Float64}, sequence::Matrix{Float64}; tables::Bool=false)
    size(template, 1) == size(sequence, 1) || throw(DimensionMismatch("Inconsistent dimentions."))
    if tables
        set_template!(d, template)
        for t in 1:size(sequence, 2)
            update!(d, sequence[:, t])
        end
        return backward(d)
    end
    d.template = template
    path = Vector{Int}(undef, size(sequence, 2))
    check(ccall((:vcb_dtw_fit_batch, libvcb200), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Int32, Int32, Ptr{Int64}, Ptr{Float64}),
                template, Int64[0, size(template, 2)], sequence, Int64[0, size(sequence, 2)], 1,
                size(template, 1), d.fstep, d.bstep, path, C_NULL))
    path
end
fit!(d::DTW, sequence::Matrix{Float64}) = fit!(d, d.template, sequence)

# batch extension: pairs back to back with frame offsets
function fit!(d::DTW, templates::Matrix{Float64}, toff::Vector{Int64}, sequences::Matrix{Float64}, soff::Vector{Int64})
    paths = Vector{Int}(undef, size(sequences, 2))
    check(ccall((:vcb_dtw_fit_batch, libvcb200), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Int32, Int32, Ptr{Int64}, Ptr{Float64}),
                templates, toff, sequences, soff, length(toff) - 1, size(templates, 1), d.fstep, d.bstep, paths, C_NULL))
    paths
end

function set_template!(d::DTW, template::Matrix{Float64})
    d.template = template
    S = size(template, 2)
    d.costtable = reshape(collect(1.0:S), S, 1)
    d.backpointer = reshape(collect(1:S), S, 1)
end

function update!(d::DTW, v::AbstractVector)
    D, S = size(d.template)
    newcost = Vector{Float64}(undef, S); newbp = Vector{Int}(undef, S)
    check(ccall((:vcb_dtw_update, libvcb200), Int32,
                (Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Int64}),
                d.template, D, S, d.costtable[:, end], Vector{Float64}(v), d.fstep, d.bstep, newcost, newbp))
    d.costtable = [d.costtable newcost]
    d.backpointer = [d.backpointer newbp]
end

function backward(d::DTW)        # src/dtw.jl:133-145 (index chasing on the host tables)
    T = size(d.costtable, 2) - 1
    minpath = zeros(Int, T)
    minpath[end] = argmin(d.costtable[:, T+1])
    for i in reverse(2:T)
        minpath[i-1] = d.backpointer[minpath[i], i+1]
    end
    minpath
end
# batch of align(src, tgt) in ONE library call (the loop of bin/align.jl:84-113; sharded over the GPUs after init(0))
function align(srcs::Vector{Matrix{Float64}}, tgts::Vector{Matrix{Float64}})
    length(srcs) == length(tgts) || throw(ArgumentError("srcs and tgts must have the same length"))
    D = size(srcs[1], 1)
    all(m -> size(m, 1) == D, srcs) && all(m -> size(m, 1) == D, tgts) ||
        throw(DimensionMismatch("order of feature vector must be equal"))
    soff = Int64[0; cumsum(size.(srcs, 2))]
    toff = Int64[0; cumsum(size.(tgts, 2))]
    src, tgt = hcat(srcs...), hcat(tgts...)
    newtgt = similar(src)
    check(ccall((:vcb_align_batch, libvcb200), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ptr{Int64}),
                src, soff, tgt, toff, length(srcs), D, newtgt, C_NULL))
    [(srcs[i], newtgt[:, soff[i]+1:soff[i+1]]) for i in 1:length(srcs)]
end

end # module DTWs
using .DTWs

# ---- callers either side (src/datasets.jl:6-13, src/align.jl:8-35) -----------------------------------
function push_delta(src::Matrix{Float64})
    D, T = size(src)
    out = Matrix{Float64}(undef, 2D, T)
    check(ccall((:vcb_push_delta_batch, libvcb200), Int32, (Ptr{Float64}, Int32, Ptr{Int64}, Int64, Ptr{Float64}),
                src, D, Int64[0, T], 1, out))
    out
end

function align(src::Matrix{Float64}, tgt::Matrix{Float64})
    size(src, 1) == size(tgt, 1) || throw(DimensionMismatch("order of feature vector must be equal"))
    newtgt = similar(src)
    check(ccall((:vcb_align_batch, libvcb200), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ptr{Int64}),
                src, Int64[0, size(src, 2)], tgt, Int64[0, size(tgt, 2)], 1, size(src, 1), newtgt, C_NULL))
    src, newtgt
end

# batch of align(src, tgt) in ONE library call (the loop of bin/align.jl:84-113; sharded over the GPUs after init(0))
function align(srcs::Vector{Matrix{Float64}}, tgts::Vector{Matrix{Float64}})
    length(srcs) == length(tgts) || throw(ArgumentError("srcs and tgts must have the same length"))
    D = size(srcs[1], 1)
    all(m -> size(m, 1) == D, srcs) && all(m -> size(m, 1) == D, tgts) ||
        throw(DimensionMismatch("order of feature vector must be equal"))
    soff = Int64[0; cumsum(size.(srcs, 2))]
    toff = Int64[0; cumsum(size.(tgts, 2))]
    src, tgt = hcat(srcs...), hcat(tgts...)
    newtgt = similar(src)
    check(ccall((:vcb_align_batch, libvcb200), Int32,
                (Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ptr{Int64}),
                src, soff, tgt, toff, length(srcs), D, newtgt, C_NULL))
    [(srcs[i], newtgt[:, soff[i]+1:soff[i+1]]) for i in 1:length(srcs)]
end

end # module
