"""K6 parity: cholRecursive!(A, Val{:L}) (reference src/cholesky.jl:37-55) and the generic Hermitian
rank-k update (src/juliaBLAS.jl:89-112) through the C ABI vs the oracle; grid of test/cholesky.jl:8-27
and test/juliaBLAS.jl:10-17."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float32: 3e-5, np.float64: 1e-12, np.complex128: 1e-12}


def _spd(rng, n, dtype, shift=0.0):
    A = rng.random((n, n))
    if dtype == np.complex128:
        A = A + 1j * rng.random((n, n))
    A = A.astype(dtype)
    S = A.conj().T @ A + shift * np.eye(n, dtype=dtype)
    return np.asfortranarray(S.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("cutoff", [1, 4])
def test_reference_grid_n50(gla, oracle, dtype, cutoff):
    """test/cholesky.jl: n = 50, AcA = A'A, A = rand(n,n); compared with LAPACK potrf and the oracle."""
    rng = np.random.default_rng(123)
    S = _spd(rng, 50, dtype, shift=1.0 if dtype == np.float32 else 0.0)
    ref = oracle.chol_recursive(S, cutoff)
    got = gla.cholRecursive_(S.copy(order="F"), "L", cutoff)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= TOL[dtype] * scale * 50
    L = np.linalg.cholesky(S.astype(np.complex128 if dtype == np.complex128 else np.float64))
    assert np.max(np.abs(np.tril(got) - L)) <= TOL[dtype] * scale * 50
    # the strict upper triangle is left untouched (src/cholesky.jl:54 returns LowerTriangular(A))
    assert np.array_equal(np.triu(got, 1), np.triu(S, 1))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 200, 257, 640])
def test_sizes_vs_oracle(gla, oracle, dtype, n):
    rng = np.random.default_rng(n)
    S = _spd(rng, n, dtype, shift=float(n))
    ref = oracle.chol_recursive(S, 1, mt=True)
    got = gla.cholRecursive_(S.copy(order="F"))
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= TOL[dtype] * scale * max(n, 10)
    assert np.array_equal(np.triu(got, 1), np.triu(S, 1))


def test_known_answer(gla):
    A = np.asfortranarray(np.array([[4.0, 2.0], [2.0, 5.0]]))
    got = gla.cholRecursive_(A)
    assert np.array_equal(got, np.array([[2.0, 2.0], [1.0, 2.0]]))


def test_not_positive_definite_raises_domain_error(gla):
    A = np.asfortranarray(np.eye(100))
    A[70, 70] = -1.0
    with pytest.raises(gla.DomainError) as ei:
        gla.cholRecursive_(A)
    assert "71" in str(ei.value)
    # a failing minor beyond 1000 must not collide with the >= 1000 CUDA/NCCL code range of the ABI: the index travels out
    # of band (gla_last_info); and like the reference (src/cholesky.jl:40) the call leaves A partially factorised
    B = np.asfortranarray(np.eye(2304) * 4.0)
    B[1500, 1500] = -1.0
    with pytest.raises(gla.DomainError) as ei:
        gla.cholRecursive_(B)
    assert ei.value.args[1] == 1501 and "1501" in str(ei.value)
    assert B[0, 0] == 2.0 and np.array_equal(np.triu(B, 1), np.zeros_like(B))
    with pytest.raises(gla.DimensionMismatch):
        gla.cholRecursive_(np.zeros((3, 4), order="F"))


def test_config2_4096(gla, oracle):
    """BASELINE config 2: A = X'X + n I, n = 4096, Float64: residual + parity with the oracle."""
    n = 4096
    rng = np.random.default_rng(123)
    X = rng.standard_normal((n, n))
    S = np.asfortranarray(X.T @ X + n * np.eye(n))
    got = gla.cholRecursive_(S.copy(order="F"))
    L = np.tril(got)
    assert np.linalg.norm(L @ L.T - S) / np.linalg.norm(S) <= 10 * n * 2.2e-16
    ref = oracle.chol_recursive(S, 1, mt=True)
    assert np.max(np.abs(L - np.tril(ref))) <= 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("n,k", [(5, 2), (5, 1), (130, 70), (300, 33)])
def test_rank_update_lower(gla, oracle, dtype, n, k):
    """test/juliaBLAS.jl:10-17: C + alpha*B*B' on the lower triangle, generic method == oracle."""
    rng = np.random.default_rng(n + k)
    Cm = _spd(rng, n, dtype)
    B = rng.standard_normal((n, k))
    if dtype == np.complex128:
        B = B + 1j * rng.standard_normal((n, k))
    B = np.asfortranarray(B.astype(dtype))
    for alpha in (0.5, -1.0):
        ref = oracle.rank_update_lower(Cm, B, alpha)
        got = gla.rankUpdate_(Cm.copy(order="F"), B, alpha)
        assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= TOL[dtype] * 10 * np.max(np.abs(ref))
        assert np.array_equal(np.triu(got, 1), np.triu(Cm, 1))
        full = Cm + alpha * (B @ B.conj().T)
        assert np.max(np.abs(np.tril(got) - np.tril(full))) <= TOL[dtype] * 10 * np.max(np.abs(full))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_unblocked_and_blocked_entry_points(gla, oracle, dtype):
    """cholUnblocked!(A, Val{:L}) (src/cholesky.jl:3-15) and cholBlocked!(A, Val{:L}, blocksize) (:17-35) against their
    own restatements in the oracle; the strict upper triangle stays untouched, as in the reference."""
    rng = np.random.default_rng(17)
    for n, bs in ((50, 7), (130, 16), (257, 64)):
        S = _spd(rng, n, dtype, shift=float(n))
        for got, ref in ((gla.cholUnblocked_(S.copy(order="F")), oracle.chol_unblocked(S)),
                         (gla.cholBlocked_(S.copy(order="F"), "L", bs), oracle.chol_blocked(S, bs))):
            scale = np.max(np.abs(ref))
            assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= TOL[dtype] * scale * max(n, 10)
            assert np.array_equal(np.triu(got, 1), np.triu(S, 1))
    with pytest.raises(gla.ArgumentError):
        gla.cholBlocked_(np.asfortranarray(np.eye(4)), "L", 0)


def test_panel_path_matches_recursion_bitwise_shape(gla, oracle):
    """Float64 above the look-ahead threshold of the right-looking driver (n > 384) and above its 256-row outer block
    (n > 6144 is too slow for the oracle here, so 1000 and 2500): parity with the oracle at ragged sizes."""
    rng = np.random.default_rng(5)
    for n in (449, 1000, 2500):
        X = rng.standard_normal((n, n))
        S = np.asfortranarray(X.T @ X + n * np.eye(n))
        got = gla.cholRecursive_(S.copy(order="F"))
        ref = oracle.chol_recursive(S, 1, mt=True)
        assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= 1e-11 * np.max(np.abs(ref))
        assert np.array_equal(np.triu(got, 1), np.triu(S, 1))


@pytest.mark.parametrize("dtype", [np.float32, np.complex128])
def test_panel_path_other_types_above_lookahead(gla, oracle, dtype):
    """Float32 (trailing updates on the tcgen05 kernel with the triangle mask and bounded CTA lifetime) and ComplexF64 (panel
    kernel at 254 registers, complex DMMA updates) through the look-ahead schedule of the right-looking driver."""
    rng = np.random.default_rng(11)
    for n in (449, 1500):
        S = _spd(rng, n, dtype, shift=float(n) if dtype == np.float32 else 1.0)
        got = gla.cholRecursive_(S.copy(order="F"))
        ref = oracle.chol_recursive(S, 1, mt=True)
        assert np.max(np.abs(np.tril(got) - np.tril(ref))) <= TOL[dtype] * n * np.max(np.abs(ref))
        assert np.array_equal(np.triu(got, 1), np.triu(S, 1))
        L = np.tril(got).astype(np.complex128 if dtype == np.complex128 else np.float64)
        Sw = S.astype(L.dtype)
        assert np.linalg.norm(L @ L.conj().T - Sw) / np.linalg.norm(Sw) <= 10 * n * np.finfo(np.float32 if dtype == np.float32 else np.float64).eps


def _hermitian_dd(rng, n, dtype, indefinite):
    """Diagonally dominant symmetric / Hermitian test matrix (all leading minors well conditioned, so LDL^T without pivoting is
    stable); with `indefinite` the diagonal alternates in sign."""
    X = rng.standard_normal((n, n))
    if dtype == np.complex128:
        X = X + 1j * rng.standard_normal((n, n))
    H = (X + X.conj().T) / 2
    d = np.abs(H).sum(axis=1) + 1.0
    if indefinite:
        d = d * np.where(np.arange(n) % 3 == 1, -1.0, 1.0)
    H[np.arange(n), np.arange(n)] = d
    return np.asfortranarray(H.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [1, 2, 5, 50, 64, 65, 130, 500, 1000])
def test_ldlt_vs_oracle(gla, oracle, dtype, uplo, n):
    """ldlt!(Hermitian(A, uplo)) (src/ldlt.jl:80-162; grid of test/ldlt.jl: n in 5, 50, 500, both triangles) against the
    oracle's restatement, the factorisation identity L D L' = A (U' D U = A), and the untouched other triangle."""
    rng = np.random.default_rng(n + (7 if uplo == "U" else 0))
    for indefinite in (False, True):
        H = _hermitian_dd(rng, n, dtype, indefinite)
        ref = oracle.ldlt(H, uplo)
        got = gla.ldlt_(H.copy(order="F"), uplo)
        tol = (3e-5 if dtype == np.float32 else 1e-12) * max(n, 10)
        tri = np.tril if uplo == "L" else np.triu
        assert np.max(np.abs(tri(got) - tri(ref))) <= tol * np.max(np.abs(tri(ref)))
        other = (np.triu(got, 1), np.triu(H, 1)) if uplo == "L" else (np.tril(got, -1), np.tril(H, -1))
        assert np.array_equal(*other)
        d = np.real(np.diag(got)).astype(np.float64)
        if indefinite and n >= 5:
            assert d.min() < 0 < d.max()
        wide = np.complex128 if dtype == np.complex128 else np.float64
        F = tri(got, -1 if uplo == "L" else 1).astype(wide) + np.eye(n)
        R = F @ np.diag(d) @ F.conj().T if uplo == "L" else F.conj().T @ np.diag(d) @ F
        Hs = np.tril(H) + np.tril(H, -1).conj().T if uplo == "L" else np.triu(H) + np.triu(H, 1).conj().T
        assert np.max(np.abs(R - Hs)) <= tol * np.max(np.abs(Hs))


def test_ldlt_zero_pivot_and_errors(gla):
    A = np.asfortranarray(np.eye(70) * 3.0)
    A[10, 10] = 0.0
    with pytest.raises(gla.ZeroPivotError) as ei:
        gla.ldlt_(A)
    assert ei.value.args[1] == 11
    with pytest.raises(gla.DimensionMismatch):
        gla.ldlt_(np.zeros((3, 4), order="F"))
    with pytest.raises(TypeError):
        gla.ldlt_(np.asfortranarray(np.eye(3, dtype=np.float16)))
    # the doc example of the reference (src/ldlt.jl docstring): [1 1; 1 -1] -> L = [1 0; 1 1], D = (1, -2)
    got = gla.ldlt_(np.asfortranarray(np.array([[1.0, 1.0], [1.0, -1.0]])), "U")
    assert np.array_equal(got, np.array([[1.0, 1.0], [1.0, -2.0]]))
