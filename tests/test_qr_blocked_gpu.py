"""K1-K3 parity: blocked Householder QR through the C ABI vs the oracle restatement of
qrBlocked!/qrUnblocked! (reference src/qr.jl:86-146), plus the reference's own test properties
(test/qr.jl:7-25) and the north_star tolerances."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS = {np.float64: 2.2e-16, np.float32: 1.2e-7, np.complex128: 2.2e-16}
RTOL_R = {np.float64: 1e-10, np.float32: 1e-4, np.complex128: 1e-10}   # north_star elementwise tolerance


def _randn(rng, m, n, dtype):
    A = rng.standard_normal((m, n))
    if dtype == np.complex128:
        A = A + 1j * rng.standard_normal((m, n))
    return np.asfortranarray(A.astype(dtype))


def _check_against_oracle(F, tau, ref_f, ref_t, dtype):
    m, n = ref_f.shape
    k = min(m, n)
    tol = RTOL_R[dtype]
    scale = np.max(np.abs(ref_f))
    # R (upper trapezoid incl. diagonal) elementwise: |d| <= tol*|ref| + tol*max|ref|
    Rg, Rr = np.triu(F)[:k], np.triu(ref_f)[:k]
    assert np.all(np.abs(Rg - Rr) <= tol * np.abs(Rr) + tol * scale)
    # reflectors and tau
    Vg, Vr = np.tril(F, -1), np.tril(ref_f, -1)
    assert np.max(np.abs(Vg - Vr)) <= 1e3 * tol * max(1.0, np.max(np.abs(Vr)))
    assert np.max(np.abs(tau - ref_t)) <= 1e2 * tol


@pytest.mark.parametrize("m,n", [(10, 5), (10, 10), (5, 10), (100, 50), (100, 100), (50, 100)])
@pytest.mark.parametrize("bz", [1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 33])
def test_reference_grid_float64(gla, oracle, m, n, bz):
    """test/qr.jl:7-25 grid: the oracle runs qrBlocked!(A, bz); the GPU result must match it for every bz
    (results are blocksize independent up to rounding) and satisfy Q'A = R, Q'(QA) = A."""
    rng = np.random.default_rng(1000 * m + 10 * n + bz)
    A = _randn(rng, m, n, np.float64)
    ref_f, ref_t = oracle.qr_blocked(A, bz)
    qr = gla.qrBlocked_(A.copy(order="F"), bz)
    _check_against_oracle(qr.factors, qr.tau, ref_f, ref_t, np.float64)
    Q = qr.QBlocked
    QtA = Q.adjoint_mul(A)
    k = min(m, n)
    if m >= n:
        np.testing.assert_allclose(QtA[:k], qr.R, rtol=0, atol=1.5e-8 * np.linalg.norm(A))
    else:
        np.testing.assert_allclose(QtA, np.triu(qr.factors), rtol=0, atol=1.5e-8 * np.linalg.norm(A))
    back = Q @ QtA
    np.testing.assert_allclose(back, A, rtol=0, atol=1.5e-8 * np.linalg.norm(A))


def test_reference_error_paths(gla):
    """test/qr.jl:28-35."""
    with pytest.raises(gla.DimensionMismatch):
        gla.reflectorApply_(np.zeros((5, 5), order="F"), np.zeros(4), 1.0)
    qr = gla.qrBlocked_(np.asfortranarray(np.random.default_rng(0).standard_normal((5, 10))))
    with pytest.raises(gla.ArgumentError):
        qr.R


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
@pytest.mark.parametrize("m,n", [(300, 200), (257, 257), (64, 64), (65, 130), (1000, 70), (129, 1), (1, 129),
                                 (512, 512), (700, 333)])
def test_types_and_ragged_shapes(gla, oracle, dtype, m, n):
    rng = np.random.default_rng(m * 7 + n)
    A = _randn(rng, m, n, dtype)
    # complex oracle = unblocked reference path (the blocked one drops a conj at src/qr.jl:72)
    ref_f, ref_t = (oracle.qr_unblocked(A) if dtype == np.complex128 else oracle.qr_blocked(A, 12))
    qr = gla.qrBlocked_(A.copy(order="F"))
    _check_against_oracle(qr.factors, qr.tau, ref_f, ref_t, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
def test_backward_error_and_orthogonality(gla, dtype):
    """north_star: ||A - QR||/||A|| <= 10 n eps, ||Q^H Q - I|| <= 10 n eps."""
    m, n = 640, 384
    rng = np.random.default_rng(5)
    A = _randn(rng, m, n, dtype)
    qr = gla.qrBlocked_(A.copy(order="F"))
    Q = qr.QBlocked
    R = np.zeros((m, n), dtype=dtype, order="F")
    R[:n] = qr.R
    QR = Q @ R
    eps = EPS[dtype]
    assert np.linalg.norm(A - QR) / np.linalg.norm(A) <= 10 * n * eps
    I = np.asfortranarray(np.eye(m, dtype=dtype))
    Qm = Q @ I
    assert np.linalg.norm(Qm.conj().T @ Qm - np.eye(m), 2) <= 10 * n * eps


def test_known_answers(gla):
    """SURVEY.md section 8c KATs (hand derived from Julia's reflector! semantics)."""
    A = np.asfortranarray(np.array([[3.0, 1.0], [4.0, 2.0]]))
    qr = gla.qrBlocked_(A)
    np.testing.assert_allclose(qr.factors, [[-5.0, -2.2], [0.5, -0.4]], rtol=1e-14)
    np.testing.assert_allclose(qr.tau, [1.6, 2.0], rtol=1e-15)
    Z = np.zeros((3, 2), order="F")
    qr = gla.qrBlocked_(Z)
    assert np.array_equal(qr.tau, [0.0, 0.0]) and np.array_equal(qr.factors, np.zeros((3, 2)))
    C = np.asfortranarray(np.array([[3j], [4.0 + 0j]]))
    qr = gla.qrBlocked_(C)
    np.testing.assert_allclose(qr.factors[:, 0], [-5.0, 4.0 / (5.0 + 3.0j)], rtol=1e-15)
    np.testing.assert_allclose(qr.tau, [1.0 + 0.6j], rtol=1e-15)


def test_larft_matches_oracle(gla, oracle):
    rng = np.random.default_rng(11)
    for dtype in (np.float64, np.complex128):
        A = _randn(rng, 90, 40, dtype)
        f, t = oracle.qr_unblocked(A)
        Tref = oracle.build_T(f, t)
        Tg = gla.QR2(f, t).QBlocked.T
        np.testing.assert_allclose(Tg, Tref, rtol=0, atol=1e-12 * np.max(np.abs(Tref)))


def test_right_reflector_apply(gla, oracle):
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.float32, np.complex128):
        A = _randn(rng, 37, 9, dtype)
        x = _randn(rng, 9, 1, dtype)[:, 0]
        tau = dtype(1.3) if dtype != np.complex128 else np.complex128(1.3 - 0.2j)
        ref = oracle.reflector_apply_right(A, x, tau)
        got = gla.reflectorApply_(A.copy(order="F"), x, tau)
        np.testing.assert_allclose(got, ref, rtol=0, atol=(1e-5 if dtype == np.float32 else 1e-13) * np.max(np.abs(ref)))


def test_config1_1024(gla, oracle):
    """BASELINE config 1: qrBlocked! on a 1024x1024 Float64 random matrix vs the oracle (blocksize 12)."""
    rng = np.random.default_rng(123)
    A = _randn(rng, 1024, 1024, np.float64)
    ref_f, ref_t = oracle.qr_blocked(A, 12)
    qr = gla.qrBlocked_(A.copy(order="F"))
    _check_against_oracle(qr.factors, qr.tau, ref_f, ref_t, np.float64)
    assert qr.tau[-1] == 2.0


def test_device_twins_and_workspace_query(gla, oracle):
    """`_dev` twins of the T build, the block application and the right reflector application against the
    oracle, and gla_workspace_query against the pool's high-water mark."""
    import torch
    rng = np.random.default_rng(21)
    st = torch.cuda.current_stream().cuda_stream
    for dtype, tdt in ((np.float64, torch.float64), (np.float32, torch.float32), (np.complex128, torch.complex128)):
        m, n, nA = 90, 40, 23
        A = _randn(rng, m, n, dtype)
        f, t = oracle.qr_unblocked(A)
        Tref = oracle.build_T(f, t)
        dF = torch.from_numpy(np.ascontiguousarray(f.T)).cuda()          # column-major m x n
        dtau = torch.from_numpy(t.copy()).cuda()
        dT = torch.zeros((n, n), dtype=tdt, device="cuda")
        gla.larft_dev(dF.data_ptr(), m, n, m, dtau.data_ptr(), dT.data_ptr(), n, st, dtype)
        tol = 1e-4 if dtype == np.float32 else 1e-12
        np.testing.assert_allclose(dT.cpu().numpy().T, Tref, rtol=0, atol=tol * np.max(np.abs(Tref)))
        B = _randn(rng, m, nA, dtype)
        for adjoint in (False, True):
            ref = oracle.block_apply(f, Tref, B, adjoint=adjoint)
            dB = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
            gla.ormqr_blocked_dev(dF.data_ptr(), m, n, m, dtau.data_ptr(), dB.data_ptr(), m, nA, m, adjoint, st, dtype)
            np.testing.assert_allclose(dB.cpu().numpy().T, ref, rtol=0, atol=10 * tol * np.max(np.abs(ref)))
        x = _randn(rng, nA, 1, dtype)[:, 0]
        tau = dtype(1.3) if dtype != np.complex128 else np.complex128(1.3 - 0.2j)
        ref = oracle.reflector_apply_right(B, x, tau)
        dB = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
        dx = torch.from_numpy(x.copy()).cuda()
        gla.reflector_apply_right_dev(dB.data_ptr(), m, nA, m, dx.data_ptr(), nA, tau, st, dtype)
        np.testing.assert_allclose(dB.cpu().numpy().T, ref, rtol=0, atol=10 * tol * np.max(np.abs(ref)))
        with pytest.raises(gla.DimensionMismatch):
            gla.reflector_apply_right_dev(dB.data_ptr(), m, nA, m, dx.data_ptr(), nA - 1, tau, st, dtype)
    assert gla.workspace_query(gla.OP_GEQR_BLOCKED, np.float64, 4096, 4096) > 4096 * 384 * 8
    assert gla.workspace_query(gla.OP_POTRF_L, np.float64, 4096, 4096) >= 4096 * 4096 * 8
    assert gla.workspace_query(gla.OP_GEQR_BATCHED, np.float64, 32, 32) == 0
    assert gla.workspace_query(gla.OP_LDLT, np.float64, 4096, 4096) >= 2 * 4096 * 4096 * 8
    assert gla.workspace_query(gla.OP_HESSENBERG, np.float64, 2048, 2048) < 4096                      # vectors live in shared memory
    assert gla.workspace_query(gla.OP_BIDIAGONALIZE, np.float64, 100, 300) >= 300 * 100 * 8          # the A^H copy of the wide case
    assert gla.workspace_query(gla.OP_BIDIAGONALIZE, np.float64, 40000, 8) >= 148 * 40000 * 8        # per-CTA slabs: the column exceeds the budget
    assert gla.workspace_query(gla.OP_SYMTRI, np.complex128, 500, 500) >= 500 * 500 * 16             # the flipped copy of uplo = 'U'
    with pytest.raises(gla.ArgumentError):
        gla.workspace_query(99, np.float64, 4, 4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
def test_wide_block_apply_and_thin_q(gla, oracle, dtype):
    """lmul!(H, A, M) / lmul!(H', A, M) (src/householder.jl:82-157) across several 384-reflector outer blocks against the
    oracle's block_apply with the full k x k T (src/qr.jl:64-83), and the thin Q: ||Q^H Q - I|| <= 10 n eps, Q R = A."""
    rng = np.random.default_rng(77)
    m, n, nA = 1300, 900, 37          # k = 900: three outer blocks (384 + 384 + 132), last panel ragged
    A = _randn(rng, m, n, dtype)
    qr = gla.qrBlocked_(A.copy(order="F"))
    f, t = qr.factors, qr.tau
    Tfull = oracle.build_T(f, t)
    B = _randn(rng, m, nA, dtype)
    tol = 2e-4 if dtype == np.float32 else 1e-11
    H = qr.QBlocked
    for adjoint in (False, True):
        ref = oracle.block_apply(f, Tfull, B, adjoint=adjoint)
        got = H.adjoint_lmul_(B.copy(order="F")) if adjoint else H.lmul_(B.copy(order="F"))
        assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref))
    Q = qr.thinQ()
    eps = EPS[dtype]
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(n), 2) <= 10 * n * eps
    assert np.linalg.norm(Q @ np.triu(f[:n]) - A) / np.linalg.norm(A) <= 10 * n * eps
