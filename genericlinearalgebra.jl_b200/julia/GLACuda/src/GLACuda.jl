# GLACuda.jl -- LAPACK2-style wrapper module over libgla_cuda.so (include/gla_cuda.h).
#
# NOT EXECUTED in this repository: the build image has no Julia runtime (SURVEY.md finding 5).  The Python/ctypes
# mirror (../../glacuda.py) makes exactly the same C calls and is what tests/ and bench.py drive.
#
# Layout follows the reference's only FFI precedent, `module LAPACK2` (src/lapack.jl:1-9): one thin function per C
# entry point, `ccall` on (symbol, library), Ptr{T} for arrays, integer status -> exception.  This module does NOT depend
# on GenericLinearAlgebra; the package extension ext/GLACudaGenericLinearAlgebraExt.jl (a weak dependency, loaded by
# `using GenericLinearAlgebra, GLACuda`) adds the more specific methods that make the library a drop-in for
# qrBlocked!, qrUnblocked!, cholRecursive!, cholBlocked!, cholUnblocked!, rankUpdate!, lmul! and the T build.
module GLACuda

using LinearAlgebra

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
version() = ccall((:gla_version, libgla), Cint, ())
device_count() = ccall((:gla_device_count, libgla), Cint, ())
set_device(dev::Integer) = chk(ccall((:gla_set_device, libgla), Cint, (Cint,), dev), "set_device")

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

# workspace query (the reference's LAPACK wrappers ask for their workspace first, src/lapack.jl:152-170)
const OP_GEQR_BLOCKED, OP_POTRF_L, OP_GEQR_BATCHED, OP_TSQR = Cint(1), Cint(2), Cint(3), Cint(4)
const OP_LDLT, OP_BIDIAGONALIZE, OP_HESSENBERG, OP_SYMTRI = Cint(5), Cint(6), Cint(7), Cint(8)
function workspace_query(op::Integer, ::Type{T}, m::Integer, n::Integer) where {T<:GLAFloat}
    b = ccall((:gla_workspace_query, libgla), Int64, (Cint, Cint, Int64, Int64), op, sizeof(T), m, n)
    b < 0 && throw(ArgumentError("workspace_query: argument $(-b) is illegal"))
    return b
end

for T in (Float32, Float64, ComplexF64)
    p = prefix(T)
    RT = real(T)
    herk = T === ComplexF64 ? :gla_zherk_lower : Symbol("gla_", p, "syrk_lower")
    @eval begin
        # gla_?geqr_blocked: A <- factors, tau filled                      (qrBlocked!, src/qr.jl:113-146)
        function geqr_blocked!(A::Matrix{$T}, tau::Vector{$T}, blocksize::Integer = 0)
            m, n = size(A)
            length(tau) >= min(m, n) || throw(DimensionMismatch("tau is too short"))
            rc = GC.@preserve A tau ccall(($(QuoteNode(Symbol("gla_", p, "geqr_blocked"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Int64), A, m, n, max(1, stride(A, 2)), tau, blocksize)
            chk(rc, "geqr_blocked!")
            return A, tau
        end
        # gla_?larft: T of all k reflectors                                 (getindex(::QR2, :QBlocked), src/qr.jl:64-83)
        function larft!(F::Matrix{$T}, tau::Vector{$T}, Tm::Matrix{$T})
            m, n = size(F)
            rc = GC.@preserve F tau Tm ccall(($(QuoteNode(Symbol("gla_", p, "larft"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}, Int64),
                F, m, n, max(1, stride(F, 2)), tau, Tm, max(1, stride(Tm, 2)))
            chk(rc, "larft!")
            return Tm
        end
        # gla_?ormqr_blocked: A <- Q A or Q^H A                              (lmul!, src/householder.jl:82-157)
        function ormqr_blocked!(V::Matrix{$T}, tau::Vector{$T}, A::Matrix{$T}, adjoint::Bool)
            size(V, 1) == size(A, 1) || throw(DimensionMismatch(""))   # src/householder.jl:87,129
            rc = GC.@preserve V A tau ccall(($(QuoteNode(Symbol("gla_", p, "ormqr_blocked"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}, Int64, Int64, Int64, Cint),
                V, size(V, 1), size(V, 2), max(1, stride(V, 2)), tau,
                A, size(A, 1), size(A, 2), max(1, stride(A, 2)), adjoint ? 1 : 0)
            chk(rc, "ormqr_blocked!")
            return A
        end
        # gla_?orgqr_thin: Q = H_1 .. H_k [I_k; 0]                            (HouseholderBlock * Matrix(I, m, k), src/householder.jl:116-117)
        function orgqr_thin!(Q::Matrix{$T}, F::Matrix{$T}, tau::Vector{$T})
            m, n = size(F)
            size(Q) == (m, min(m, n)) || throw(DimensionMismatch("Q must be m x min(m, n)"))
            rc = GC.@preserve F tau Q ccall(($(QuoteNode(Symbol("gla_", p, "orgqr_thin"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}, Int64),
                F, m, n, max(1, stride(F, 2)), tau, Q, max(1, stride(Q, 2)))
            chk(rc, "orgqr_thin!")
            return Q
        end
        # gla_?reflector_apply_right                                          (reflectorApply!(A, x, tau), src/qr.jl:19-42)
        function reflector_apply_right!(A::Matrix{$T}, x::Vector{$T}, tau::Number)
            m, n = size(A)
            if length(x) != n     # message of src/qr.jl:22-26
                throw(DimensionMismatch(
                    "reflector must have same length as second dimension of matrix, but got $(length(x)) and $n"))
            end
            tr = Ref{$T}(convert($T, tau))
            rc = GC.@preserve A x ccall(($(QuoteNode(Symbol("gla_", p, "reflector_apply_right"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Int64, Ref{$T}), A, m, n, max(1, stride(A, 2)), x, length(x), tr)
            chk(rc, "reflector_apply_right!")
            return A
        end
        # gla_?potrf_{recursive,blocked,unblocked}_L                         (src/cholesky.jl:37-55, 17-35, 3-15)
        function potrf_recursive_L!(A::Matrix{$T}, cutoff::Integer = 1)
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(($(QuoteNode(Symbol("gla_", p, "potrf_recursive_L"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64), A, n, max(1, stride(A, 2)), cutoff)
            chk(rc, "potrf_recursive_L!")
            return A
        end
        function potrf_blocked_L!(A::Matrix{$T}, blocksize::Integer)
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(($(QuoteNode(Symbol("gla_", p, "potrf_blocked_L"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64), A, n, max(1, stride(A, 2)), blocksize)
            chk(rc, "potrf_blocked_L!")
            return A
        end
        function potrf_unblocked_L!(A::Matrix{$T})
            n = LinearAlgebra.checksquare(A)
            rc = GC.@preserve A ccall(($(QuoteNode(Symbol("gla_", p, "potrf_unblocked_L"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64), A, n, max(1, stride(A, 2)))
            chk(rc, "potrf_unblocked_L!")
            return A
        end
        # gla_{s,d}syrk_lower / gla_zherk_lower: lower(C) += alpha A A^H      (rankUpdate!, src/juliaBLAS.jl:89-112)
        function herk_lower!(C::Matrix{$T}, A::Matrix{$T}, alpha::Real)
            n = LinearAlgebra.checksquare(C)
            size(A, 1) == n || throw(DimensionMismatch("first dimension of A does not match C"))
            rc = GC.@preserve C A ccall(($(QuoteNode(herk)), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Ptr{$T}, Int64, Int64, $RT),
                C, n, max(1, stride(C, 2)), A, size(A, 2), max(1, stride(A, 2)), convert($RT, alpha))
            chk(rc, "herk_lower!")
            return C
        end
        # gla_?geqr_batched: A is m x n x batch, factorised in place           (BASELINE configs[2])
        function geqr_batched!(A::Array{$T,3}, tau::Matrix{$T} = zeros($T, min(size(A, 1), size(A, 2)), size(A, 3)))
            m, n, batch = size(A)
            rc = GC.@preserve A tau ccall(($(QuoteNode(Symbol("gla_", p, "geqr_batched"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}), A, m, n, batch, tau)
            chk(rc, "geqr_batched!")
            return A, tau
        end
    end
end

# gla_?bidiagonalize / gla_?hessenberg / gla_?symtri: the two-sided reductions, in place
#   (bidiagonalize!, src/svd.jl:328-381; _hessenberg!, src/eigenGeneral.jl:18-31; symtriLower!/symtriUpper!,
#    src/eigenSelfAdjoint.jl:450-564)
for (T, p) in ((Float32, "s"), (Float64, "d"), (ComplexF64, "z"))
    @eval begin
        function bidiagonalize_inplace!(A::Matrix{$T})
            m, n = size(A)
            nl, nr = m >= n ? (n, max(n - 1, 0)) : (max(m - 1, 0), m)
            taul, taur = zeros($T, nl), zeros($T, nr)
            rc = GC.@preserve A taul taur ccall(($(QuoteNode(Symbol("gla_", p, "bidiagonalize"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Int64, Ptr{$T}, Ptr{$T}), A, m, n, max(1, stride(A, 2)), taul, taur)
            chk(rc, "bidiagonalize_inplace!")
            return A, taul, taur
        end
        function hessenberg_inplace!(A::Matrix{$T})
            n = LinearAlgebra.checksquare(A)
            tau = zeros($T, max(n - 1, 0))
            rc = GC.@preserve A tau ccall(($(QuoteNode(Symbol("gla_", p, "hessenberg"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Ptr{$T}), A, n, max(1, stride(A, 2)), tau)
            chk(rc, "hessenberg_inplace!")
            return A, tau
        end
        function symtri_inplace!(A::Matrix{$T}, uplo::Char, tau::Vector{$T} = zeros($T, max(size(A, 1) - 1, 0)))
            n = LinearAlgebra.checksquare(A)
            length(tau) >= n - 1 || throw(DimensionMismatch("tau is too short"))
            rc = GC.@preserve A tau ccall(($(QuoteNode(Symbol("gla_", p, "symtri"))), libgla), Cint,
                (Ptr{$T}, Int64, Int64, Cint, Ptr{$T}), A, n, max(1, stride(A, 2)), Cint(uplo), tau)
            chk(rc, "symtri_inplace!")
            return A, tau
        end
    end
end

# gla_{s,d,z}ldlt: LDL^H without pivoting, in place                    (ldlt!, src/ldlt.jl:155-162)
const GLA_ERR_SINGULAR = Cint(901)
for (T, sym) in ((Float32, :gla_sldlt), (Float64, :gla_dldlt), (ComplexF64, :gla_zldlt))
    @eval function ldlt_inplace!(A::Matrix{$T}, uplo::Char, blocksize::Integer)
        n = LinearAlgebra.checksquare(A)
        rc = GC.@preserve A ccall(($(QuoteNode(sym)), libgla), Cint,
            (Ptr{$T}, Int64, Int64, Cint, Int64), A, n, max(1, stride(A, 2)), Cint(uplo), blocksize)
        rc == GLA_ERR_SINGULAR && throw(SingularException(Int(last_info())))
        chk(rc, "ldlt_inplace!")
        return A
    end
end

# R factor of a tall-skinny Float64 matrix by a TSQR tree (BASELINE configs[3]); row signs: DESIGN.md
function tsqrR(A::Matrix{Float64})
    m, n = size(A)
    R = zeros(n, n)
    rc = GC.@preserve A R ccall((:gla_dtsqr, libgla), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64), A, m, n, max(1, stride(A, 2)), R, max(1, n))
    chk(rc, "tsqrR")
    return UpperTriangular(R)
end

end # module
