# DArrayB200.jl -- reference-side binding for libdab200.so  (UNVERIFIED: there is no Julia in the build image; this file is the
# binding a DistributedArrays.jl maintainer would add, written against include/dab200.h; the executable stand-in used by the
# tests is the Python ctypes layer in distributedarrays.jl_b200/_lib.py, which binds the very same symbols).
#
# Idea: DistributedArrays.jl is generic in the chunk type `A` of `DArray{T,N,A}` (reference src/darray.jl:25).  `B200Array{T,N}`
# is a chunk type whose data lives in one GPU's HBM; it overloads exactly the Base generics the hot path calls on
# `localpart(d)`, so that the reference's own `map!` / broadcast / `mapreduce` / `mapreduce(...; dims)` code runs unchanged and
# lands in the CUDA kernels:
#
#   reference call site                                   Base generic overloaded here          C entry point
#   src/broadcast.jl:80  copyto!(localpart(dest), lbc)    Base.copyto!(::B200Array, ::Broadcasted)   dab_affine / dab_unary / dab_binary* / dab_broadcast_expr
#   src/broadcast.jl:96  copy(lbc)                        Base.copy(::Broadcasted{B200Style})        same + dab_alloc
#   src/mapreduce.jl:8   map!(f, localpart(dest), src)    Base.map!(f, ::B200Array, ::B200Array)     dab_affine / ...
#   src/mapreduce.jl:23,31 mapreduce(f, op, localpart)    Base.mapreduce / Base.reduce               dab_reduce_host
#   src/mapreduce.jl:64,77 mapreduce(...; dims)           Base.mapreducedim!                         dab_reducedim
#   src/mapreduce.jl:100-127 all/any/count/extrema        Base.all / any / count / extrema           dab_reduce_host
#   src/darray.jl:815    localpart(d)[idxs...]            Base.getindex(::B200Array, ranges...)      dab_copy_box
#   src/darray.jl:824,831 fill! / rand!                   Base.fill! / Random.rand!                  dab_fill / dab_rand_u01
#   src/mapreduce.jl:34  reduce(op, results)              (DArray method below)                      dab_mapreduce_all
#   src/linalg.jl:95-97,141 localpart(A)*xj, localpart(A)'*xj   Base.:*                                 dab_gemv
#   src/linalg.jl:1-17   transpose!(lp, rp)               LinearAlgebra.transpose! / adjoint!        dab_transpose_box
#   src/sort.jl:8,22,61  sort(localpart(d)), sort!(lp)    Base.sort / Base.sort!                     dab_sort
#   src/sort.jl:8,22,61  sort(localpart(d); by = f)       sort_by (keys = f.(a) by broadcast)        dab_sort_by_key
module DArrayB200

using Distributed, DistributedArrays, LinearAlgebra
using LinearAlgebra: Adjoint, Transpose
import Base.Broadcast: Broadcasted, BroadcastStyle, AbstractArrayStyle

const libdab = get(ENV, "LIBDAB200", "libdab200.so")

# ---- status handling: mirror the reference's exception types -----------------------------------------------------------------
const DAB_OK = Int32(0)
function check(st::Int32, ctx::Ptr{Cvoid} = C_NULL)
    st == DAB_OK && return nothing
    msg = unsafe_string(ccall((:dab_last_error, libdab), Cstring, (Ptr{Cvoid},), ctx))
    st == 2 || st == 3 ? throw(ArgumentError(msg)) :
    st == 4 ? throw(DimensionMismatch(msg)) : throw(ErrorException("libdab200 [$st]: $msg"))
end

# ---- one context per worker process (one GPU per worker) ---------------------------------------------------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        ndev = Ref{Int32}(0)
        check(ccall((:dab_device_count, libdab), Int32, (Ref{Int32},), ndev))
        dev = Int32((myid() - 2 + ndev[]) % ndev[])           # worker pid 2 -> GPU 0, ...
        check(ccall((:dab_init, libdab), Int32, (Int32, Ref{Ptr{Cvoid}}), dev, CTX))
    end
    CTX[]
end

dab_dtype(::Type{Float32}) = Int32(0); dab_dtype(::Type{Float64}) = Int32(1)
dab_dtype(::Type{Int32}) = Int32(2);   dab_dtype(::Type{Int64}) = Int32(3); dab_dtype(::Type{Bool}) = Int32(4)

# ---- the chunk type ------------------------------------------------------------------------------------------------------------
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:dab_alloc, libdab), Int32, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx(), prod(dims) * sizeof(T), p), ctx())
        a = new{T,N}(p[], dims)
        finalizer(x -> ccall((:dab_free, libdab), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), x.ptr), a)   # cf. src/darray.jl:47-49
        a
    end
end
B200Array{T}(u::UndefInitializer, dims::Int...) where {T} = B200Array{T,length(dims)}(u, dims)
Base.size(a::B200Array) = a.dims
Base.similar(a::B200Array{T}, ::Type{S}, dims::Dims) where {T,S} = B200Array{S,length(dims)}(undef, dims)
Base.getindex(::B200Array, ::Int...) = error("scalar indexing of a B200Array is disabled (cf. DistributedArrays.allowscalar(false))")

# host <-> device  (distribute / Array(d): src/darray.jl:544-555, 574-582)
function B200Array(a::Array{T,N}) where {T,N}
    d = B200Array{T,N}(undef, size(a))
    check(ccall((:dab_h2d, libdab), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), d.ptr, a, sizeof(a)), ctx())
    check(ccall((:dab_sync, libdab), Int32, (Ptr{Cvoid},), ctx()), ctx()); d
end
function Base.Array(d::B200Array{T,N}) where {T,N}
    a = Array{T,N}(undef, size(d))
    check(ccall((:dab_d2h, libdab), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx(), a, d.ptr, sizeof(a)), ctx())
    check(ccall((:dab_sync, libdab), Int32, (Ptr{Cvoid},), ctx()), ctx()); a
end
Base.convert(::Type{B200Array{T,N}}, a::Array{T,N}) where {T,N} = B200Array(a)       # empty_localpart, src/darray.jl:62

function Base.fill!(a::B200Array{T}, x) where {T}                                       # src/darray.jl:822-827
    check(ccall((:dab_fill, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t, Ref{T}), ctx(), dab_dtype(T), a.ptr, length(a), T(x)), ctx()); a
end

# ---- elementwise: map! and in-place broadcast ------------------------------------------------------------------------------------
# A closure cannot cross the C ABI.  Known shapes are pattern-matched onto the hand-written kernels; any other Broadcasted tree is
# lowered to a C expression string and JIT-compiled by dab_broadcast_expr (NVRTC) -- `lower(bc)` below is the analogue of the
# Python tracer in distributedarrays.jl_b200/_broadcast.py.
struct Affine{T}; a::T; b::T; end                      # x -> a*x + b, two roundings (Julia never contracts to FMA)
(f::Affine)(x) = f.a * x + f.b

function Base.map!(f::Affine{T}, dest::B200Array{T}, src::B200Array{T}) where {T}       # src/mapreduce.jl:8
    length(dest) == length(src) || throw(DimensionMismatch("map!: lengths differ"))
    check(ccall((:dab_affine, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ref{T}, Ref{T}, Csize_t),
                ctx(), dab_dtype(T), dest.ptr, src.ptr, f.a, f.b, length(dest)), ctx()); dest
end

struct B200Style{N} <: AbstractArrayStyle{N} end
B200Style(::Val{N}) where {N} = B200Style{N}()
Base.Broadcast.BroadcastStyle(::Type{<:B200Array{T,N}}) where {T,N} = B200Style{N}()
Base.similar(bc::Broadcasted{B200Style{N}}, ::Type{T}) where {N,T} = B200Array{T,N}(undef, map(length, axes(bc)))

# y .= a .* x .+ b  ==  Broadcasted(+, (Broadcasted(*, (a, x)), b))                      # src/broadcast.jl:80
function Base.copyto!(dest::B200Array{T}, bc::Broadcasted{<:B200Style}) where {T}
    ab = match_affine(bc, T)
    if ab !== nothing
        a, x, b = ab
        return map!(Affine{T}(a, b), dest, x)
    end
    return broadcast_expr!(dest, bc)                    # general tree -> "jl_sub(a0, jl_mul(a1, jl_sin(a2)))" -> NVRTC
end
match_affine(bc, ::Type{T}) where {T} =
    (bc.f === (+) && length(bc.args) == 2 && bc.args[1] isa Broadcasted && bc.args[1].f === (*) &&
     bc.args[1].args[1] isa T && bc.args[1].args[2] isa B200Array{T} && bc.args[2] isa T) ?
        (bc.args[1].args[1], bc.args[1].args[2], bc.args[2]) : nothing
# General trees: walk the Broadcasted and emit the C expression dab_broadcast_expr compiles with NVRTC (the same lowering the
# Python tracer in distributedarrays.jl_b200/_broadcast.py performs; helper names are the jl_* functions of the JIT prelude).
const FN2 = Dict{Any,String}((+) => "jl_add", (-) => "jl_sub", (*) => "jl_mul", (/) => "jl_div", rem => "jl_rem", mod => "jl_mod",
                             div => "jl_idiv", max => "jl_max", min => "jl_min", (^) => "jl_pow", (<) => "jl_lt", (<=) => "jl_le",
                             (>) => "jl_gt", (>=) => "jl_ge", (==) => "jl_eq", (!=) => "jl_ne", (&) => "jl_and", (|) => "jl_or", xor => "jl_xor")
const FN1 = Dict{Any,String}((-) => "jl_neg", abs => "jl_abs", abs2 => "jl_abs2", sqrt => "jl_sqrt", inv => "jl_inv", floor => "jl_floor",
                             ceil => "jl_ceil", sign => "jl_sign", sin => "jl_sin", cos => "jl_cos", tan => "jl_tan", exp => "jl_exp",
                             log => "jl_log", tanh => "jl_tanh", isnan => "jl_isnan", identity => "")
ctype(::Type{Float32}) = "float"; ctype(::Type{Float64}) = "double"; ctype(::Type{Int32}) = "int"; ctype(::Type{Int64}) = "long long"; ctype(::Type{Bool}) = "bool"
literal(x::Float32) = "__int_as_float((int)0x$(string(reinterpret(UInt32, x), base = 16)))"
literal(x::Float64) = "__longlong_as_double((long long)0x$(string(reinterpret(UInt64, x), base = 16))ULL)"
literal(x::Integer) = "(($(ctype(typeof(x))))$(x))"
literal(x::Bool) = x ? "true" : "false"

# returns (expression string, element type); `args` collects the array / Ref-scalar leaves in order -> a0, a1, ...
function lower(bc::Broadcasted, args::Vector{Any})
    parts = [lower(a, args) for a in bc.args]
    Ts = map(last, parts)
    T = Base.promote_op(bc.f, Ts...)                                  # Julia's own result type: promotion stays exactly Julia's
    conv = [Ti === Tp ? s : "(($(ctype(Tp)))($s))" for ((s, Ti), Tp) in zip(parts, promote_types(bc.f, Ts, T))]
    name = length(conv) == 1 ? get(FN1, bc.f, nothing) : get(FN2, bc.f, nothing)
    name === nothing && error("DArrayB200: $(bc.f) is not served by the broadcast lowering (no host fallback)")
    (isempty(name) ? conv[1] : "$name($(join(conv, ", ")))", T)
end
lower(a::B200Array{T}, args) where {T} = (push!(args, a); ("a$(length(args) - 1)", T))
lower(x::Number, args) = (literal(x), typeof(x))
lower(r::Base.RefValue, args) = lower(r[], args)
# comparison / arithmetic operands are promoted to a common type first (Base.promote); bitwise and comparison results keep Bool
promote_types(f, Ts, T) = (f in (<, <=, >, >=, ==, !=)) ? fill(promote_type(Ts...), length(Ts)) : fill(T, length(Ts))

function broadcast_expr!(dest::B200Array{T,N}, bc::Broadcasted) where {T,N}
    args = Any[]
    expr, Tr = lower(bc, args)
    Tr === T || (expr = "(($(ctype(T)))($expr))")
    shape = Csize_t[size(dest)..., ones(Int, 4 - N)...]
    dense(sz) = Csize_t[cumprod([1; collect(sz)[1:end-1]])..., zeros(Int, 4 - length(sz))...]
    ostr = dense(size(dest)); ostr[N+1:end] .= 0
    strides = Csize_t[]
    for a in args                                                        # 0 = extruded dim (src/broadcast.jl:112-113)
        st = dense(size(a)); for d in 1:4; (d > ndims(a) || size(a, d) == 1) && (st[d] = 0); end; append!(strides, st)
    end
    check(ccall((:dab_broadcast_expr, libdab), Int32,
                (Ptr{Cvoid}, Cstring, Int32, Ptr{Cvoid}, Ptr{Csize_t}, Ptr{Csize_t}, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Csize_t}, Ptr{UInt64}),
                ctx(), expr, dab_dtype(T), dest.ptr, shape, ostr, length(args), Int32[dab_dtype(eltype(a)) for a in args],
                Ptr{Cvoid}[a.ptr for a in args], strides, zeros(UInt64, length(args))), ctx())
    dest
end

# ---- reductions --------------------------------------------------------------------------------------------------------------------
const OPS = Dict{Any,Int32}(Base.add_sum => 0, (+) => 0, Base.mul_prod => 1, (*) => 1, max => 2, min => 3)
const MAPS = Dict{Any,Int32}(identity => 0, abs => 1, abs2 => 2, (-) => 3)
result_type(::Type{T}, op) where {T} = (T <: AbstractFloat || op >= 2) ? T : Int64          # add_sum / mul_prod widen Int32

function Base.mapreduce(f, op, a::B200Array{T}; dims = :, init = nothing) where {T}         # src/mapreduce.jl:23,31,64
    haskey(OPS, op) && haskey(MAPS, f) || error("DArrayB200: mapreduce($f, $op) is not served by a kernel (no host fallback)")
    dims === Colon() || return mapreducedim(f, op, a, dims, init)
    out = zeros(UInt64, 2)
    check(ccall((:dab_reduce_host, libdab), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                ctx(), dab_dtype(T), OPS[op], MAPS[f], C_NULL, a.ptr, length(a), out), ctx())
    reinterpret(result_type(T, OPS[op]), out)[1]
end
Base.reduce(op, a::B200Array; kw...) = mapreduce(identity, op, a; kw...)

function mapreducedim(f, op, a::B200Array{T,N}, dims, init) where {T,N}
    region = Tuple(dims)
    all(d -> d >= 1, region) || throw(ArgumentError("region dimension(s) must be ≥ 1, got $dims"))
    rdims = ntuple(i -> i in region ? 1 : size(a, i), N)
    R = B200Array{result_type(T, OPS[op]),N}(undef, rdims)
    init === nothing || fill!(R, init)
    # one (inner, reduce, outer) pass per maximal run of reduced dims; single leading / single trailing run shown
    k = findfirst(i -> !(i in region), 1:N)
    inner, red, outer = k === nothing ? (1, length(a), 1) :
                        (first(region) == 1 ? (1, prod(size(a)[1:k-1]), prod(size(a)[k:end])) :
                                              (prod(size(a)[1:first(region)-1]), prod(size(a)[collect(region)]), prod(size(a)[last(region)+1:end])))
    check(ccall((:dab_reducedim, libdab), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Cvoid}, Csize_t, Csize_t, Csize_t, Ptr{Cvoid}, Int32),
                ctx(), dab_dtype(T), OPS[op], MAPS[f], a.ptr, inner, red, outer, R.ptr, init === nothing ? 0 : 1), ctx())
    R
end
Base.mapreducedim!(f, op, R::B200Array, A::B200Array) = (copyto!(R, mapreduce(f, op, A; dims = findall(size(R) .!= size(A)))); R)  # src/mapreduce.jl:77

# ---- the combine seam: sum(d::DArray{T,N,<:B200Array}) in ONE call per worker ----------------------------------------------------------
# replaces  results = asyncmap(procs(d)) do p; remotecall_fetch(...) end;  reduce(op, results)   (src/mapreduce.jl:29-35):
# every worker launches the fused kernel (chunk reduce + peer-memory all-gather + ordered fold); the caller fetches one scalar.
function Base._mapreduce(f, op, ::IndexCartesian, d::DArray{T,N,<:B200Array}) where {T,N}
    haskey(OPS, op) && haskey(MAPS, f) || error("DArrayB200: mapreduce($f, $op) is not served by a kernel")
    results = asyncmap(procs(d)) do p
        remotecall_fetch(p) do
            a = localpart(d); out = zeros(UInt64, 2)
            check(ccall((:dab_mapreduce_all, libdab), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                        ctx(), dab_dtype(T), OPS[op], MAPS[f], C_NULL, a.ptr, length(a), out), ctx())
            reinterpret(result_type(T, OPS[op]), out)[1]
        end
    end
    first(results)        # every worker already holds the left-folded result
end

# communicator / mailbox bring-up: rank 0 creates the NCCL id, the ids and IPC handles travel over Distributed (control plane only)
function init_comm(pids = workers())
    id = remotecall_fetch(pids[1]) do
        buf = zeros(UInt8, 128); check(ccall((:dab_comm_unique_id, libdab), Int32, (Ptr{UInt8},), buf)); buf
    end
    @sync for (r, p) in enumerate(pids)
        @async remotecall_wait(p) do
            check(ccall((:dab_comm_init_rank, libdab), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx(), id, r - 1, length(pids)), ctx())
        end
    end
    handles = [remotecall_fetch(p) do
                   h = zeros(UInt8, 64); check(ccall((:dab_mailbox_create, libdab), Int32, (Ptr{Cvoid}, Ptr{UInt8}), ctx(), h), ctx()); h
               end for p in pids]
    allh = reduce(vcat, handles)
    @sync for (r, p) in enumerate(pids)
        @async remotecall_wait(p) do
            check(ccall((:dab_mailbox_attach, libdab), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx(), allh, r - 1, length(pids)), ctx())
        end
    end
end

# ---- Level-2 and sort on the chunk type (widening rows): the generic code of src/linalg.jl and src/sort.jl then runs unchanged ------
# localpart(A)*xj  /  localpart(A)'*xj  inside mul!(y::DVector, A::DMatrix, x, α, β)  (src/linalg.jl:95-97, 141)
function gemv(trans::Bool, A::B200Array{T,2}, x::B200Array{T,1}) where {T}
    m, n = size(A)
    r = B200Array{T,1}(undef, (trans ? n : m,))
    check(ccall((:dab_gemv, libdab), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Csize_t, Csize_t, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx(), dab_dtype(T), trans ? 1 : 0, A.ptr, m, n, x.ptr, r.ptr), ctx())
    r
end
Base.:*(A::B200Array{T,2}, x::B200Array{T,1}) where {T} = gemv(false, A, x)
Base.:*(A::Adjoint{T,<:B200Array{T,2}}, x::B200Array{T,1}) where {T<:Real} = gemv(true, parent(A), x)
Base.:*(A::Transpose{T,<:B200Array{T,2}}, x::B200Array{T,1}) where {T} = gemv(true, parent(A), x)

# transpose!(lp, rp) / adjoint!(lp, rp) of copy(::Transpose{T,<:DArray{T,2}})  (src/linalg.jl:1-17), real T
function LinearAlgebra.transpose!(dst::B200Array{T,2}, src::B200Array{T,2}) where {T}
    rows, cols = size(src)
    check(ccall((:dab_transpose_box, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t, Csize_t, Csize_t),
                ctx(), sizeof(T), dst.ptr, cols, src.ptr, rows, rows, cols), ctx())
    dst
end
LinearAlgebra.adjoint!(dst::B200Array{T,2}, src::B200Array{T,2}) where {T<:Real} = transpose!(dst, src)

# sort(localpart(d)) / sort!(lp_sorting)  (src/sort.jl:8, 22, 61); keys only, isless order
function Base.sort!(a::B200Array{T,1}; kw...) where {T}
    tmp = B200Array{T,1}(undef, size(a))
    check(ccall((:dab_sort, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                ctx(), dab_dtype(T), a.ptr, a.ptr, tmp.ptr, length(a)), ctx())
    a
end
function sort_keys(a::B200Array{T,1}) where {T}
    out = B200Array{T,1}(undef, size(a)); tmp = B200Array{T,1}(undef, size(a))
    check(ccall((:dab_sort, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                ctx(), dab_dtype(T), a.ptr, out.ptr, tmp.ptr, length(a)), ctx())
    out
end

# sort(localpart(d); by = f) / sort!(lp_sorting; by = f)  (src/sort.jl:8, 22, 61 with the :by keyword of :111): keys = f.(a) through
# the broadcast lowering (one fused kernel), then the values are ordered stably by the keys (packed key|position words sorted by K11).
function sort_by(a::B200Array{T,1}, keys::B200Array{K,1}) where {T,K}
    n = length(a); need = Ref{Csize_t}(0)
    check(ccall((:dab_sort_by_key_scratch_bytes, libdab), Int32, (Int32, Csize_t, Ref{Csize_t}), dab_dtype(K), n, need), ctx())
    out = B200Array{T,1}(undef, size(a)); scratch = B200Array{UInt8,1}(undef, (Int(need[]),))
    check(ccall((:dab_sort_by_key, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Csize_t),
                ctx(), dab_dtype(K), keys.ptr, sizeof(T), a.ptr, out.ptr, scratch.ptr, need[], n), ctx())
    out
end
# keyword arguments do not take part in dispatch: ONE method serves both spellings
Base.sort(a::B200Array{T,1}; by = identity, kw...) where {T} = by === identity ? sort_keys(a) : sort_by(a, by.(a))

# localpart(A) * Bjk, transpose(localpart(A)) * Bjk inside _matmatmul!  (src/linalg.jl:218-226): K12, tcgen05 3xTF32 for Float32
function gemm(transA::Bool, A::B200Array{T,2}, B::B200Array{T,2}) where {T}
    m, k = transA ? reverse(size(A)) : size(A)
    size(B, 1) == k || throw(DimensionMismatch("matrix A has dimensions ($m, $k), matrix B has dimensions $(size(B))"))
    R = B200Array{T,2}(undef, (m, size(B, 2)))
    check(ccall((:dab_gemm, libdab), Int32,
                (Ptr{Cvoid}, Int32, Int32, Csize_t, Csize_t, Csize_t, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Csize_t),
                ctx(), dab_dtype(T), transA, m, size(B, 2), k, A.ptr, size(A, 1), B.ptr, size(B, 1), R.ptr, m), ctx())
    R
end
Base.:*(A::B200Array{T,2}, B::B200Array{T,2}) where {T} = gemm(false, A, B)
Base.:*(A::Adjoint{T,<:B200Array{T,2}}, B::B200Array{T,2}) where {T<:Real} = gemm(true, parent(A), B)
Base.:*(A::Transpose{T,<:B200Array{T,2}}, B::B200Array{T,2}) where {T} = gemm(true, parent(A), B)

# add!(localpart(y), R[i,j], alpha) for all j after the beta scaling (src/linalg.jl:62-76, 101-117, 232-252): one fused launch over the
# stack of tile results that the producers PUT into this worker's exchange arena; dab_peer_barrier orders the puts (device side)
function accumulate_stack!(y::B200Array{T}, beta, alpha, stack::Ptr{Cvoid}, count::Integer) where {T}
    check(ccall((:dab_accumulate_stack, libdab), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Csize_t, Ref{T}, Ref{T}, Ptr{Cvoid}, Csize_t, Int32),
                ctx(), dab_dtype(T), y.ptr, length(y), T(beta), T(alpha), stack, length(y), count), ctx())
    y
end
peer_barrier() = check(ccall((:dab_peer_barrier, libdab), Int32, (Ptr{Cvoid},), ctx()), ctx())

# localpart(d)[idxs...] with StepRange / Vector{Int} indices (src/darray.jl:661, 798-820): strided / table-driven gather
function Base.getindex(a::B200Array{T,N}, I::Vararg{Union{AbstractRange{Int},Vector{Int}},N}) where {T,N}
    out = B200Array{T,N}(undef, map(length, I))
    sstr = cumprod((1, size(a)[1:end-1]...)); dstr = cumprod((1, size(out)[1:end-1]...))
    tabs = [B200Array(Int64.((collect(I[k]) .- first(I[k])) .* sstr[k])) for k in 1:N]       # source offsets as tables (affine ones could pass strides)
    src = a.ptr + (sum((first(I[k]) - 1) * sstr[k] for k in 1:N)) * sizeof(T)
    check(ccall((:dab_gather_box, libdab), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{Clonglong}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Clonglong}, Ptr{Ptr{Cvoid}}, Ptr{Csize_t}),
                ctx(), sizeof(T), N, out.ptr, Clonglong[dstr...], C_NULL, src, zeros(Clonglong, N), Ptr{Cvoid}[t.ptr for t in tabs],
                Csize_t[length.(I)...]), ctx())
    out
end

# user code is then unchanged:
#   d = DArray(I -> B200Array(rand(Float32, map(length, I))), (8 * 2^30,))
#   d .= 1.5f0 .* d .+ 0.25f0 ;  map!(Affine(2f0, 1f0), d, d) ;  sum(d) ;  maximum(d) ;  sum(d2, dims = 1) ;  A * B ;  A' * x ;  sort(v)
end # module
