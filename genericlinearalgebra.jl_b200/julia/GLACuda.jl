# GLACuda.jl -- the Julia host side of libgla_cuda.so (include/gla_cuda.h).
#
# NOT EXECUTED in this repository: the build image has no Julia runtime (SURVEY.md finding 5).  It is
# the binding a maintainer of GenericLinearAlgebra.jl adds; the Python/ctypes mirror in ../glacuda.py
# makes exactly the same C calls and is what tests/ and bench.py drive.
#
# Mechanism = the reference's own plugin API, multiple dispatch: the reference defines its hot-path
# functions on `StridedMatrix{T} where T` (src/qr.jl:113-118, src/qr.jl:86, src/cholesky.jl:37).  This
# module adds MORE SPECIFIC methods for `Matrix{T}`, T in {Float32, Float64, ComplexF64}, that ccall the
# library.  BigFloat, Quaternion, views and every other element type keep dispatching to the generic
# reference code.  There is no CPU fallback inside these methods: if the library fails, they throw.
#
# FFI conventions follow the reference's only FFI precedent, `module LAPACK2` (src/lapack.jl:1-9):
# ccall on (symbol, library), Ptr{T} for arrays, integer status -> exception.
module GLACuda

using LinearAlgebra
import GenericLinearAlgebra
import GenericLinearAlgebra: QR2, HouseholderBlock, qrBlocked!, qrUnblocked!, cholRecursive!, cholBlocked!, cholUnblocked!,
    rankUpdate!

const libgla = get(ENV, "GLA_CUDA_LIB", "libgla_cuda.so")
const GLAFloat = Union{Float32,Float64,ComplexF64}

prefix(::Type{Float32}) = "s"
prefix(::Type{Float64}) = "d"
prefix(::Type{ComplexF64}) = "z"

struct GLACudaError <: Exception
    code::Cint
    msg::String
end

last_error() = unsafe_string(ccall((:gla_last_error_string, libgla), Cstring, ()))

# return-code convention of include/gla_cuda.h: 0 ok, -k illegal argument k, GLA_ERR_NOT_POSDEF (900) from potrf
# with the failing minor out of band in gla_last_info(), >= 1000 CUDA/NCCL runtime failure
const GLA_ERR_NOT_POSDEF = Cint(900)
last_info() = ccall((:gla_last_info, libgla), Int64, ())
function chk(rc::Cint, what::AbstractString)
    rc == 0 && return nothing
    rc >= 1000 && throw(GLACudaError(rc, "$what: $(last_error())"))
    rc < 0 && throw(DimensionMismatch("$what: argument $(-rc) is illegal"))
    if rc == GLA_ERR_NOT_POSDEF
        # the reference fails with DomainError from sqrt of a negative real (src/cholesky.jl:40)
        k = last_info()
        throw(DomainError(k, "$what: leading minor $k is not positive definite"))
    end
    throw(GLACudaError(rc, what))
end

for T in (Float32, Float64, ComplexF64)
    p = prefix(T)
    geqr = Symbol("gla_", p, "geqr_blocked")
    potrf = Symbol("gla_", p, "potrf_recursive_L")
    potrf_unb = Symbol("gla_", p, "potrf_unblocked_L")
    potrf_blk = Symbol("gla_", p, "potrf_blocked_L")
    larft = Symbol("gla_", p, "larft")
    ormqr = Symbol("gla_", p, "ormqr_blocked")
    batched = Symbol("gla_", p, "geqr_batched")
    rapply = Symbol("gla_", p, "reflector_apply_right")
    herk = T === ComplexF64 ? :gla_zherk_lower : Symbol("gla_", p, "syrk_lower")
    RT = real(T)

    @eval begin
        # ---- qrBlocked!(A, blocksize, τ, work)            replaces src/qr.jl:113-146
        # `work` is accepted for signature compatibility and ignored; `blocksize` is a hint.
        function qrBlocked!(
            A::Matrix{$T},
            blocksize::Integer = 12,
            τ::Vector{$T} = fill(zero($T), min(size(A)...)),
            work = nothing,
        )
            m, n = size(A)
            length(τ) >= min(m, n) || throw(DimensionMismatch("τ is too short"))
            rc = GC.@preserve A τ ccall(
                ($(QuoteNode(geqr)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Int64),
                A, m, n, max(1, stride(A, 2)), τ, blocksize)
            chk(rc, "qrBlocked!")
            return QR2{$T,typeof(A),typeof(τ)}(A, τ)
        end

        # ---- qrUnblocked!(A, τ)                            replaces src/qr.jl:86-111
        # same factorisation (V, τ, R do not depend on the panel width beyond rounding)
        qrUnblocked!(A::Matrix{$T}, τ::Vector{$T} = fill(zero($T), min(size(A)...))) =
            qrBlocked!(A, 0, τ)

        # ---- getindex(::QR2, Tuple{:QBlocked})             replaces src/qr.jl:64-83
        # (with the conj the reference omits at :72 for complex element types)
        function Base.getindex(F::QR2{$T,Matrix{$T},Vector{$T}}, ::Type{Tuple{:QBlocked}})
            m, n = size(F)
            k = min(m, n)
            Tm = zeros($T, k, k)
            rc = GC.@preserve F Tm ccall(
                ($(QuoteNode(larft)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}, Int64),
                F.factors, m, n, max(1, stride(F.factors, 2)), F.τ, Tm, max(1, k))
            chk(rc, "getindex(QR2, :QBlocked)")
            return HouseholderBlock{$T,typeof(F.factors),Matrix{$T}}(F.factors, UpperTriangular(Tm))
        end

        # ---- lmul!(H, A, M) / lmul!(H', A, M)              replaces src/householder.jl:82-115,119-157
        # The library rebuilds the per-panel T factors from (V, τ); τ_j = T[j,j] (src/qr.jl:77).
        function LinearAlgebra.lmul!(H::HouseholderBlock{$T,Matrix{$T}}, A::Matrix{$T}, M = nothing)
            _ormqr!(H, A, false)
        end
        function LinearAlgebra.lmul!(
            adjH::Adjoint{$T,<:HouseholderBlock{$T,Matrix{$T}}}, A::Matrix{$T}, M = nothing)
            _ormqr!(parent(adjH), A, true)
        end
        function _ormqr!(H::HouseholderBlock{$T,Matrix{$T}}, A::Matrix{$T}, adjoint::Bool)
            V = H.V
            size(V, 1) == size(A, 1) || throw(DimensionMismatch(""))   # src/householder.jl:87,129
            τ = diag(H.T)
            rc = GC.@preserve V A τ ccall(
                ($(QuoteNode(ormqr)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Cint),
                V, size(V, 1), size(V, 2), max(1, stride(V, 2)), τ,
                A, size(A, 1), size(A, 2), max(1, stride(A, 2)), adjoint ? 1 : 0)
            chk(rc, "lmul!(HouseholderBlock)")
            return A
        end

        # ---- right reflectorApply!(A, x, τ)                replaces src/qr.jl:19-42
        function LinearAlgebra.reflectorApply!(A::Matrix{$T}, x::Vector{$T}, τ::Number)
            m, n = size(A)
            if length(x) != n     # message of src/qr.jl:22-26
                throw(DimensionMismatch(
                    "reflector must have same length as second dimension of matrix, but got $(length(x)) and $n"))
            end
            τr = Ref{$T}(convert($T, τ))
            rc = GC.@preserve A x ccall(
                ($(QuoteNode(rapply)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Int64, Ref{$T}),
                A, m, n, max(1, stride(A, 2)), x, length(x), τr)
            chk(rc, "reflectorApply!")
            return A
        end

        # ---- cholRecursive!(A, Val{:L}, cutoff)            replaces src/cholesky.jl:37-55
        function cholRecursive!(A::Matrix{$T}, ::Type{Val{:L}}, cutoff = 1)
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(
                ($(QuoteNode(potrf)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64), A, n, max(1, stride(A, 2)), cutoff)
            chk(rc, "cholRecursive!")
            return LowerTriangular(A)
        end

        # ---- cholUnblocked!(A, Val{:L}) / cholBlocked!(A, Val{:L}, blocksize)   replace src/cholesky.jl:3-15, :17-35
        # (the lower factor is unique: the same device routine; both return A like the reference)
        function cholUnblocked!(A::Matrix{$T}, ::Type{Val{:L}})
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(
                ($(QuoteNode(potrf_unb)), libgla), Cint,
                (Ptr{$T}, Int64, Int64), A, n, max(1, stride(A, 2)))
            chk(rc, "cholUnblocked!")
            return A
        end
        function cholBlocked!(A::Matrix{$T}, ::Type{Val{:L}}, blocksize::Integer)
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(
                ($(QuoteNode(potrf_blk)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64), A, n, max(1, stride(A, 2)), blocksize)
            chk(rc, "cholBlocked!")
            return A
        end

        # ---- rankUpdate!(Hermitian(C,:L), A, α::Real)      replaces src/juliaBLAS.jl:89-112
        function rankUpdate!(C::Hermitian{$T,Matrix{$T}}, A::Matrix{$T}, α::Real)
            C.uplo == 'L' || return invoke(rankUpdate!, Tuple{Hermitian,StridedVecOrMat,Real}, C, A, α)
            n = size(C, 1)
            size(A, 1) == n || throw(DimensionMismatch("first dimension of A does not match C"))
            rc = GC.@preserve C A ccall(
                ($(QuoteNode(herk)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Int64, $RT),
                C.data, n, max(1, stride(C.data, 2)), A, size(A, 2), max(1, stride(A, 2)), convert($RT, α))
            chk(rc, "rankUpdate!")
            return C
        end

        # ---- batched small QR (new workload, BASELINE configs[2]); A is m x n x batch, in place
        function qrBatched!(A::Array{$T,3}, τ::Matrix{$T} = zeros($T, min(size(A, 1), size(A, 2)), size(A, 3)))
            m, n, batch = size(A)
            rc = GC.@preserve A τ ccall(
                ($(QuoteNode(batched)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}), A, m, n, batch, τ)
            chk(rc, "qrBatched!")
            return A, τ
        end
    end
end

# module-local qr (the reference's sign convention, NOT LinearAlgebra.qr's LAPACK convention)
qr(A::Matrix{T}) where {T<:GLAFloat} = qrBlocked!(copy(A))

# R factor of a tall-skinny Float64 matrix by a TSQR tree (BASELINE configs[3]); rows signs: DESIGN.md
function tsqrR(A::Matrix{Float64})
    m, n = size(A)
    R = zeros(n, n)
    rc = GC.@preserve A R ccall((:gla_dtsqr, libgla), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64), A, m, n, max(1, stride(A, 2)), R, max(1, n))
    chk(rc, "tsqrR")
    return UpperTriangular(R)
end

end # module
