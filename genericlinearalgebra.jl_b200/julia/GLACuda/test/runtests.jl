# Mirrors the reference's own tests for the hot path (test/qr.jl:7-35, test/cholesky.jl:8-27, test/juliaBLAS.jl:10-17)
# with the GPU methods active.  NOT EXECUTED in this repository (no Julia runtime in the image); the same assertions run
# through the Python mirror in tests/test_qr_blocked_gpu.py, tests/test_cholesky_gpu.py.
using Test, LinearAlgebra, Random
using GenericLinearAlgebra, GLACuda
import GenericLinearAlgebra: qrBlocked!, cholRecursive!, cholBlocked!, cholUnblocked!, rankUpdate!

Random.seed!(123)

@testset "extension is active" begin
    @test Base.get_extension(GLACuda, :GLACudaGenericLinearAlgebraExt) !== nothing
    @test which(qrBlocked!, Tuple{Matrix{Float64}}).module !== GenericLinearAlgebra
    @test which(qrBlocked!, Tuple{Matrix{BigFloat}}).module === GenericLinearAlgebra   # BigFloat stays on the reference path
end

@testset "qrBlocked! grid of test/qr.jl, m=$m n=$n bz=$bz" for (m, n) in ((10, 5), (10, 10), (5, 10), (100, 50), (100, 100), (50, 100)),
    bz in (1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 33)
    A = randn(m, n)
    Aqr = qrBlocked!(copy(A), bz)
    AqrQ = Aqr[Tuple{:QBlocked}]
    if m >= n
        @test (AqrQ' * A)[1:min(m, n), :] ≈ Aqr[Tuple{:R}]
    else
        @test (AqrQ' * A) ≈ triu(Aqr.factors)
    end
    @test AqrQ' * (AqrQ * A) ≈ A
end

@testset "error paths of test/qr.jl" begin
    @test_throws DimensionMismatch LinearAlgebra.reflectorApply!(zeros(5, 5), zeros(4), 1.0)
    @test_throws ArgumentError qrBlocked!(randn(5, 10))[Tuple{:R}]
end

@testset "Cholesky of test/cholesky.jl, T=$T" for T in (Float32, Float64, ComplexF32, ComplexF64)
    T === ComplexF32 && continue   # no GPU method: stays on the reference path
    n = 50
    A = rand(T, n, n)
    AcA = A'A
    Lref = cholesky(Hermitian(AcA)).L
    @test LowerTriangular(cholUnblocked!(copy(AcA), Val{:L})) ≈ Lref
    @test LowerTriangular(cholBlocked!(copy(AcA), Val{:L}, 5)) ≈ Lref
    @test cholRecursive!(copy(AcA), Val{:L}, 1) ≈ Lref
    @test cholRecursive!(copy(AcA), Val{:L}, 4) ≈ Lref
    B = Matrix{T}(I, 100, 100); B[71, 71] = -1
    @test_throws DomainError cholRecursive!(B, Val{:L})
end

@testset "rankUpdate! of test/juliaBLAS.jl" begin
    C = Matrix(Hermitian(randn(5, 5) + 5I)); B = randn(5, 2)
    @test tril(rankUpdate!(Hermitian(copy(C), :L), B, 0.5).data) ≈ tril(C + 0.5 * B * B')
end

@testset "two-sided reductions (test/svd.jl:100-112, test/eigengeneral.jl:239-249, test/eigenselfadjoint.jl:61-68)" begin
    A = randn(8, 8)
    @test svdvals(Matrix(GenericLinearAlgebra.bidiagonalize!(copy(A)).bidiagonal)) ≈ svdvals(A)
    BF = GenericLinearAlgebra.bidiagonalize!(copy(A))
    @test (BF.rightQ' * Matrix(I, size(A)...)) * BF.rightQ ≈ I
    n = 10
    A = randn(n, n)
    LHF = GenericLinearAlgebra._hessenberg!(copy(A))
    @test tril(LHF.Q' * A * LHF.Q, -2) ≈ zeros(n, n) atol = 1e-14
    for uplo in (:L, :U)
        S = Hermitian(A + A', uplo)
        Tf = GenericLinearAlgebra.symtri!(copy(S))
        Q = Array(Tf.reflectors)
        @test Q'Q ≈ I
        @test Q' * S * Q ≈ Tf.diagonals
    end
    @test GenericLinearAlgebra.bidiagonalize!(big.(A)) isa GenericLinearAlgebra.BidiagonalFactorization   # BigFloat: reference path
end
