# GLACudaGenericLinearAlgebraExt.jl -- package extension: loaded automatically when both GenericLinearAlgebra and
# GLACuda are in the session (Project.toml [weakdeps] / [extensions]).  NOT EXECUTED here (no Julia runtime in the image).
#
# Mechanism = the reference's own plugin API, multiple dispatch: the reference defines its hot-path functions on
# `StridedMatrix{T} where T` (src/qr.jl:113-118, src/qr.jl:86, src/cholesky.jl:3,17,37).  This extension adds MORE
# SPECIFIC methods for `Matrix{T}`, T in {Float32, Float64, ComplexF64}, that call the library through GLACuda's
# wrappers.  BigFloat, Quaternion, views and every other element type keep dispatching to the generic reference code.
# There is no CPU fallback inside these methods: if the library fails, they throw.
module GLACudaGenericLinearAlgebraExt

using LinearAlgebra
import GLACuda
import GenericLinearAlgebra
import GenericLinearAlgebra: QR2, HouseholderBlock, qrBlocked!, qrUnblocked!, cholRecursive!, cholBlocked!,
    cholUnblocked!, rankUpdate!

for T in (Float32, Float64, ComplexF64)
    @eval begin
        # ---- qrBlocked!(A, blocksize, tau, work)            replaces src/qr.jl:113-146
        # `work` is accepted for signature compatibility and ignored; `blocksize` is a hint.
        function qrBlocked!(A::Matrix{$T}, blocksize::Integer = 12,
                            tau::Vector{$T} = fill(zero($T), min(size(A)...)), work = nothing)
            GLACuda.geqr_blocked!(A, tau, blocksize)
            return QR2{$T,typeof(A),typeof(tau)}(A, tau)
        end

        # ---- qrUnblocked!(A, tau)                            replaces src/qr.jl:86-111
        # same factorisation (V, tau, R do not depend on the panel width beyond rounding)
        qrUnblocked!(A::Matrix{$T}, tau::Vector{$T} = fill(zero($T), min(size(A)...))) = qrBlocked!(A, 0, tau)

        # ---- getindex(::QR2, Tuple{:QBlocked})               replaces src/qr.jl:64-83
        # (with the conj the reference omits at :72 for complex element types)
        function Base.getindex(F::QR2{$T,Matrix{$T},Vector{$T}}, ::Type{Tuple{:QBlocked}})
            k = min(size(F)...)
            Tm = GLACuda.larft!(F.factors, F.τ, zeros($T, k, k))
            return HouseholderBlock{$T,typeof(F.factors),Matrix{$T}}(F.factors, UpperTriangular(Tm))
        end

        # ---- lmul!(H, A, M) / lmul!(H', A, M)                replaces src/householder.jl:82-115,119-157
        # The library rebuilds the per-panel T factors from (V, tau) with tau_j = T[j,j] (src/qr.jl:77).  That is only
        # the same operator when H.T IS the compact-WY factor of (V, tau) -- true for every block produced by
        # getindex(:QBlocked); a user-built HouseholderBlock with another T goes to the generic reference method.
        function _is_wy_block(H::HouseholderBlock{$T,Matrix{$T}})
            k = min(size(H.V)...)
            size(H.T) == (k, k) || return false
            Tm = GLACuda.larft!(H.V, diag(H.T), zeros($T, k, k))
            return isapprox(Tm, Matrix(H.T); rtol = 64 * eps(real($T)))
        end
        function LinearAlgebra.lmul!(H::HouseholderBlock{$T,Matrix{$T}}, A::Matrix{$T}, M = nothing)
            _is_wy_block(H) || return invoke(LinearAlgebra.lmul!,
                Tuple{HouseholderBlock{$T},StridedMatrix{$T},StridedMatrix{$T}}, H, A,
                similar(A, (min(size(H.V)...), size(A, 2))))
            GLACuda.ormqr_blocked!(H.V, diag(H.T), A, false)
        end
        function LinearAlgebra.lmul!(adjH::Adjoint{$T,<:HouseholderBlock{$T,Matrix{$T}}}, A::Matrix{$T}, M = nothing)
            H = parent(adjH)
            _is_wy_block(H) || return invoke(LinearAlgebra.lmul!,
                Tuple{Adjoint{$T,<:HouseholderBlock{$T}},StridedMatrix{$T},StridedMatrix{$T}}, adjH, A,
                similar(A, (min(size(H.V)...), size(A, 2))))
            GLACuda.ormqr_blocked!(H.V, diag(H.T), A, true)
        end

        # ---- right reflectorApply!(A, x, tau)                replaces src/qr.jl:19-42
        LinearAlgebra.reflectorApply!(A::Matrix{$T}, x::Vector{$T}, tau::Number) =
            GLACuda.reflector_apply_right!(A, x, tau)

        # ---- cholRecursive! / cholBlocked! / cholUnblocked!  replace src/cholesky.jl:37-55, 17-35, 3-15
        cholRecursive!(A::Matrix{$T}, ::Type{Val{:L}}, cutoff = 1) =
            LowerTriangular(GLACuda.potrf_recursive_L!(A, cutoff))
        cholBlocked!(A::Matrix{$T}, ::Type{Val{:L}}, blocksize::Integer) = GLACuda.potrf_blocked_L!(A, blocksize)
        cholUnblocked!(A::Matrix{$T}, ::Type{Val{:L}}) = GLACuda.potrf_unblocked_L!(A)

        # ---- rankUpdate!(Hermitian(C,:L), A, alpha::Real)    replaces src/juliaBLAS.jl:89-112
        function rankUpdate!(C::Hermitian{$T,Matrix{$T}}, A::Matrix{$T}, alpha::Real)
            C.uplo == 'L' || return invoke(rankUpdate!, Tuple{Hermitian,StridedVecOrMat,Real}, C, A, alpha)
            GLACuda.herk_lower!(C.data, A, alpha)
            return C
        end
    end
end

# ---- ldlt!(A::Hermitian, blocksize)                      replaces src/ldlt.jl:155-162
for T in (Float32, Float64, ComplexF64)
    @eval function LinearAlgebra.ldlt!(A::Hermitian{$T,Matrix{$T}}, blocksize::Int = max(1, 128 ÷ sizeof($T)))
        GLACuda.ldlt_inplace!(A.data, A.uplo, blocksize)
        return LDLt(A)
    end
end

# ---- two-sided reductions (SURVEY 8 f3): the step after QR in svd / eigen of the reference
for T in (Float32, Float64, ComplexF64)
    @eval begin
        # bidiagonalize!(A)                                   replaces src/svd.jl:328-381
        function GenericLinearAlgebra.bidiagonalize!(A::Matrix{$T})
            m, n = size(A)
            _, taul, taur = GLACuda.bidiagonalize_inplace!(A)
            bd = m >= n ? Bidiagonal(real(diag(A)), real(diag(A, 1)), :U) : Bidiagonal(real(diag(A)), real(diag(A, -1)), :L)
            return GenericLinearAlgebra.BidiagonalFactorization{eltype(bd),typeof(bd.dv),typeof(A),typeof(taul)}(
                bd, A, taul, taur)
        end
        # _hessenberg!(A)                                     replaces src/eigenGeneral.jl:18-31
        function GenericLinearAlgebra._hessenberg!(A::Matrix{$T})
            _, tau = GLACuda.hessenberg_inplace!(A)
            return Hessenberg(A, tau)
        end
        # symtriLower!(AS, tau, u) / symtriUpper!(AS, tau, u) replace src/eigenSelfAdjoint.jl:450-564 (`u` is scratch: ignored)
        function GenericLinearAlgebra.symtriLower!(AS::Matrix{$T}, tau::Vector{$T} = zeros($T, size(AS, 1) - 1), u = nothing)
            GLACuda.symtri_inplace!(AS, 'L', tau)
            return GenericLinearAlgebra.SymmetricTridiagonalFactorization(
                GenericLinearAlgebra.EigenQ('L', AS, tau), SymTridiagonal(real(diag(AS)), real(diag(AS, -1))))
        end
        function GenericLinearAlgebra.symtriUpper!(AS::Matrix{$T}, tau::Vector{$T} = zeros($T, size(AS, 1) - 1), u = nothing)
            GLACuda.symtri_inplace!(AS, 'U', tau)
            return GenericLinearAlgebra.SymmetricTridiagonalFactorization(
                GenericLinearAlgebra.EigenQ('U', AS, tau), SymTridiagonal(real(diag(AS)), real(diag(AS, 1))))
        end
    end
end

end # module
