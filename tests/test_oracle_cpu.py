"""CPU suite, part 1: the oracle (C++ restatement of the reference) against the known-answer vectors
(tests/golden/kat.json), against LAPACK after sign normalisation, and against the reference's own test
properties (test/qr.jl, test/cholesky.jl, test/juliaBLAS.jl).  No GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def _c(v):
    return np.array([complex(a, b) for a, b in v])


def test_kat_qr_2x2(oracle):
    k = KAT["qr_2x2"]
    for fn in (lambda A: oracle.qr_unblocked(A), lambda A: oracle.qr_blocked(A, 12), lambda A: oracle.qr_blocked(A, 1)):
        f, t = fn(np.array(k["A"]))
        np.testing.assert_allclose(f, k["factors"], rtol=1e-15)
        np.testing.assert_allclose(t, k["tau"], rtol=1e-15)


def test_kat_reflector(oracle):
    k = KAT["reflector_zero"]
    x, tau = oracle.reflector(np.array(k["x"]))
    assert tau == 0.0 and np.array_equal(x, k["x_out"])
    k = KAT["reflector_neg"]
    x, tau = oracle.reflector(np.array(k["x"]))
    np.testing.assert_allclose(x, k["x_out"], rtol=1e-15)
    assert abs(tau - k["tau"]) < 1e-15
    k = KAT["reflector_complex"]
    x, tau = oracle.reflector(_c(k["x"]))
    np.testing.assert_allclose(x, _c(k["x_out"]), rtol=1e-15)
    assert abs(tau - complex(*k["tau"])) < 1e-15
    # length-1 nonzero vector is still reflected: tau = 2, x1 <- -x1
    x, tau = oracle.reflector(np.array([0.4]))
    assert tau == 2.0 and x[0] == -0.4


def test_kat_cholesky_larft_apply_rank(oracle):
    k = KAT["chol_2x2"]
    assert np.array_equal(oracle.chol_recursive(np.array(k["A"])), np.array(k["inplace"]))
    assert np.array_equal(oracle.chol_unblocked(np.array(k["A"])), np.array(k["inplace"]))
    k = KAT["larft_2"]
    T = oracle.build_T(np.array(k["F"]), np.array(k["tau"]))
    np.testing.assert_allclose(T, k["T"], rtol=1e-15)
    k = KAT["apply_right"]
    out = oracle.reflector_apply_right(np.array(k["A"]), np.array(k["x"]), k["tau"])
    np.testing.assert_allclose(out, k["out"], rtol=1e-15)
    k = KAT["rank_update"]
    out = oracle.rank_update_lower(np.array(k["C"]), np.array(k["A"]), k["alpha"])
    assert np.array_equal(out, np.array(k["out"]))


@pytest.mark.parametrize("m,n", [(10, 5), (10, 10), (5, 10), (100, 50), (100, 100), (50, 100)])
@pytest.mark.parametrize("bz", [1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 33])
def test_reference_qr_properties(oracle, m, n, bz):
    """test/qr.jl:7-25 on the oracle: Q'A = R, Q'(QA) = A, and |R| == LAPACK's |R|."""
    rng = np.random.default_rng(m * 100 + n + bz)
    A = rng.standard_normal((m, n))
    f, tau = oracle.qr_blocked(A, bz)
    T = oracle.build_T(f, tau)
    QtA = oracle.block_apply(f, T, A, adjoint=True)
    k = min(m, n)
    np.testing.assert_allclose(QtA[:k], np.triu(f)[:k], atol=1e-12 * np.linalg.norm(A))
    if m > k:
        assert np.max(np.abs(QtA[k:])) < 1e-12 * np.linalg.norm(A)
    back = oracle.block_apply(f, T, QtA, adjoint=False)
    np.testing.assert_allclose(back, A, atol=1e-12 * np.linalg.norm(A))
    Rl = np.linalg.qr(A, mode="r")
    np.testing.assert_allclose(np.abs(np.triu(f)[:k]), np.abs(Rl), atol=1e-12 * np.linalg.norm(A))
    # Julia's convention: max tau = 2 exactly when m <= n (last reflector has length 1)
    assert (tau[-1] == 2.0) == (m <= n)


def test_blocked_equals_unblocked_real_and_conj_bug_for_complex(oracle):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((60, 40))
    fu, tu = oracle.qr_unblocked(A)
    for bs in (1, 4, 12, 33):
        fb, tb = oracle.qr_blocked(A, bs)
        assert np.max(np.abs(fb - fu)) < 1e-13 * np.max(np.abs(fu))
    Z = A + 1j * rng.standard_normal((60, 40))
    fu, tu = oracle.qr_unblocked(Z)
    fb, tb = oracle.qr_blocked(Z, 12)                      # conj-corrected T build
    assert np.max(np.abs(fb - fu)) < 1e-13 * np.max(np.abs(fu))
    fl, tl = oracle.qr_blocked(Z, 12, literal=True)        # literal src/qr.jl:72 (no conj): O(1) wrong
    assert np.max(np.abs(fl - fu)) > 1e-3
    assert np.max(np.abs(np.diag(fu).imag)) == 0.0          # diag(R) exactly real


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_reference_cholesky_grid(oracle, dtype):
    """test/cholesky.jl:8-27: every variant ~= LAPACK potrf, n = 50."""
    rng = np.random.default_rng(123)
    A = rng.random((50, 50))
    if dtype == np.complex128:
        A = A + 1j * rng.random((50, 50))
    S = (A.conj().T @ A + (1.0 if dtype == np.float32 else 0.0) * np.eye(50)).astype(dtype)
    L = np.linalg.cholesky(S.astype(np.complex128 if dtype == np.complex128 else np.float64))
    tol = 2e-3 if dtype == np.float32 else 1e-10
    for got in (oracle.chol_unblocked(S), oracle.chol_blocked(S, 5), oracle.chol_blocked(S, 10),
                oracle.chol_recursive(S, 1), oracle.chol_recursive(S, 4), oracle.chol_recursive(S, 1, mt=True)):
        assert np.max(np.abs(np.tril(got) - L)) < tol * np.max(np.abs(L))
        assert np.array_equal(np.triu(got, 1), np.triu(S, 1))
    with pytest.raises(oracle.DomainError):
        oracle.chol_recursive(-np.eye(3))


def test_reference_rank_update_identities(oracle):
    """test/juliaBLAS.jl:10-17."""
    rng = np.random.default_rng(1)
    for cplx in (False, True):
        A = rng.standard_normal((5, 5)) + (1j * rng.standard_normal((5, 5)) if cplx else 0)
        A = A + A.conj().T
        B = rng.standard_normal((5, 2)) + (1j * rng.standard_normal((5, 2)) if cplx else 0)
        got = oracle.rank_update_lower(A, B, 0.5)
        full = A + 0.5 * B @ B.conj().T
        np.testing.assert_allclose(np.tril(got), np.tril(full), atol=1e-14)
        assert np.array_equal(np.triu(got, 1), np.triu(A, 1))
        np.testing.assert_array_equal(oracle.rank_update_lower(A, B, 0.5, mt=True), got)


def test_error_paths(oracle):
    with pytest.raises(ValueError):
        oracle.reflector_apply_right(np.zeros((5, 5)), np.zeros(4), 1.0)       # test/qr.jl:29-33
    with pytest.raises(ValueError):
        oracle.block_apply(np.zeros((5, 2)), np.zeros((2, 2)), np.zeros((4, 3)))


def test_ldlt_restatement(oracle):
    """ldlt!(Hermitian(A, uplo)) (src/ldlt.jl:80-162): the factorisation identity for both triangles and several block
    sizes, the docstring example of the reference, and the untouched other triangle."""
    rng = np.random.default_rng(9)
    for dtype in (np.float64, np.complex128):
        for n in (1, 2, 5, 50, 130):
            X = rng.standard_normal((n, n))
            if dtype == np.complex128:
                X = X + 1j * rng.standard_normal((n, n))
            P = (X @ X.conj().T + n * np.eye(n)).astype(dtype)      # Hermitian positive definite: no pivoting needed
            for uplo in ("L", "U"):
                for bs in (1, 7, 16, 200):
                    F = oracle.ldlt(P, uplo, bs)
                    d = np.diag(F)
                    if uplo == "L":
                        L = np.tril(F, -1) + np.eye(n)
                        R = L @ np.diag(d) @ L.conj().T
                        assert np.array_equal(np.triu(F, 1), np.triu(P, 1))
                    else:
                        U = np.triu(F, 1) + np.eye(n)
                        R = U.conj().T @ np.diag(d) @ U
                        assert np.array_equal(np.tril(F, -1), np.tril(P, -1))
                    assert np.max(np.abs(R - P)) <= 1e-12 * n * np.max(np.abs(P))
    got = oracle.ldlt(np.array([[1.0, 1.0], [1.0, -1.0]]), "U")
    assert np.array_equal(got, np.array([[1.0, 1.0], [1.0, -2.0]]))
    got = oracle.ldlt(np.array([[1.0, 1.0], [1.0, 1.0]]), "L")
    assert np.array_equal(got, np.array([[1.0, 1.0], [1.0, 0.0]]))


# ---------------------------------------------------------------------------------------------- two-sided reductions (f3)
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
def test_twosided_restatements(oracle, dtype):
    """bidiagonalize! / _hessenberg! / symtri! restatements (src/svd.jl:328-381, src/eigenGeneral.jl:18-31,
    src/eigenSelfAdjoint.jl:450-564): the stored reflectors reproduce the condensed form from the input (what
    test/svd.jl:100-112, test/eigengeneral.jl:239-249 and test/eigenselfadjoint.jl:61-68 rely on), spectra are invariant,
    the other triangle of symtri! is untouched."""
    from twosided_helpers import bidiag_residual, hessenberg_residual, symtri_residual
    rng = np.random.default_rng(42)
    eps = np.finfo(np.float32 if dtype == np.float32 else np.float64).eps

    def rand(shape):
        A = rng.standard_normal(shape)
        if dtype == np.complex128:
            A = A + 1j * rng.standard_normal(shape)
        return np.asfortranarray(A.astype(dtype))

    for shape in [(1, 1), (5, 1), (1, 5), (10, 10), (50, 30), (30, 50), (129, 65)]:
        A = rand(shape)
        F, tl, tr, dv, ev, uplo = oracle.bidiagonalize(A.copy(order="F"))
        assert uplo == ("U" if shape[0] >= shape[1] else "L")
        assert bidiag_residual(A, F, tl, tr) <= 30 * max(shape) * eps * max(1.0, np.abs(A).max())
        k = min(shape)
        B = np.diag(dv.astype(np.float64)) + (np.diag(ev.astype(np.float64), 1 if uplo == "U" else -1) if k > 1 else 0)
        s_ref = np.linalg.svd(A.astype(np.complex128), compute_uv=False)
        assert np.max(np.abs(np.sort(s_ref) - np.sort(np.linalg.svd(B, compute_uv=False)))) <= 200 * k * eps * max(1.0, s_ref.max())
    for n in (1, 2, 3, 10, 65):
        A = rand((n, n))
        F, tau = oracle.hessenberg(A.copy(order="F"))
        assert hessenberg_residual(A, F, tau) <= 30 * n * eps * max(1.0, np.abs(A).max()) * max(1.0, np.sqrt(n))
        S = np.asfortranarray(A + A.conj().T)
        for uplo in "LU":
            poison = np.full_like(S, 7)
            Sin = np.asfortranarray(np.tril(S) + np.triu(poison, 1) if uplo == "L" else np.triu(S) + np.tril(poison, -1))
            F, tau, dv, ev = oracle.symtri(Sin.copy(order="F"), uplo)
            assert symtri_residual(S, F, tau, uplo) <= 30 * n * eps * max(1.0, np.abs(S).max()) * max(1.0, np.sqrt(n))
            if uplo == "L":
                assert np.array_equal(np.triu(F, 1), np.triu(Sin, 1))
            else:
                assert np.array_equal(np.tril(F, -1), np.tril(Sin, -1))
            if dtype != np.complex128 and n >= 2:
                assert tau[n - 2] == 0   # real element types stop one step earlier (src/eigenSelfAdjoint.jl:462)
