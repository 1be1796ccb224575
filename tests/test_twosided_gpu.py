"""f3 parity: the two-sided Householder reductions bidiagonalize! (reference src/svd.jl:328-381), _hessenberg!
(src/eigenGeneral.jl:18-31) and symtriLower!/symtriUpper! (src/eigenSelfAdjoint.jl:450-564) through the C ABI
against the oracle's literal restatements, plus the invariants the reference's own tests check
(test/svd.jl:100-112, test/eigengeneral.jl:239-249, test/eigenselfadjoint.jl:61-68)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64, np.complex128]
TOL = {np.float32: 1e-4, np.float64: 1e-10, np.complex128: 1e-10}   # north_star: R elementwise 1e-10 / 1e-4


def _rand(rng, shape, dtype):
    A = rng.standard_normal(shape)
    if dtype == np.complex128:
        A = A + 1j * rng.standard_normal(shape)
    return np.asfortranarray(A.astype(dtype))


def _close(got, ref, dtype, what, n=1, exact=None):
    """Elementwise against the oracle at the north_star tolerance.  Float32 beyond 64 steps: unlike R of a QR, the late
    reflectors of a two-sided reduction are computed from a trailing matrix that already carries the rounding of n
    two-sided updates in BOTH implementations (different summation orders), so there the bar is "as accurate as the
    reference's own Float32 arithmetic": the distance to the Float64 oracle result (`exact`) must not exceed four times
    the Float32 oracle's distance to it."""
    scale = max(1.0, float(np.max(np.abs(ref)))) if ref.size else 1.0
    err = float(np.max(np.abs(got - ref))) if ref.size else 0.0
    tol = TOL[dtype] * scale
    if err <= tol:
        return
    if dtype == np.float32 and n > 64 and exact is not None:
        e_ref = float(np.max(np.abs(ref.astype(np.float64) - exact)))
        e_got = float(np.max(np.abs(got.astype(np.float64) - exact)))
        if e_got <= 4 * e_ref + tol:
            return
        # reflector! is discontinuous where real(x[1]) changes sign (nu = copysign(norm, real(x[1]))): a pivot that is
        # zero to Float32 rounding after hundreds of steps may come out with the other sign.  That negates one row /
        # column of the remaining problem and nothing else (same v, same tau, |d|, |e| unchanged), so the magnitudes must
        # still agree to the same bar.
        a_ref = float(np.max(np.abs(np.abs(ref.astype(np.float64)) - np.abs(exact))))
        a_got = float(np.max(np.abs(np.abs(got.astype(np.float64)) - np.abs(exact))))
        assert a_got <= 4 * a_ref + tol, f"{what}: GPU {e_got:.3e} (|.|: {a_got:.3e}) vs oracle {e_ref:.3e} from the Float64 result"
        return
    assert err <= tol, f"{what}: {err:.3e} > {tol:.3e}"


from twosided_helpers import bidiag_residual as _bidiag_residual, hessenberg_residual as _hessenberg_residual, \
    symtri_residual as _symtri_residual, wide as _wide   # noqa: E402


BIDIAG_SHAPES = [(1, 1), (5, 1), (1, 5), (2, 2), (3, 2), (2, 3), (10, 10), (50, 30), (30, 50), (257, 129), (129, 257),
                 (600, 600), (1100, 37), (37, 1100)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", BIDIAG_SHAPES)
def test_bidiagonalize_vs_oracle(gla, oracle, dtype, shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    A = _rand(rng, shape, dtype)
    F, tl, tr, dv, ev, uplo = oracle.bidiagonalize(A.copy(order="F"))
    G = gla.bidiagonalize_(A.copy(order="F"))
    assert G.uplo == uplo
    k = max(shape)   # steps and the length of the dots both feed the Float32 rounding
    # backward error with the reflectors as stored: Ql^H A Qr == B (what test/svd.jl relies on)
    eps = np.finfo(np.float32 if dtype == np.float32 else np.float64).eps
    scale = max(1.0, float(np.max(np.abs(A))))
    assert _bidiag_residual(A, G.reflectors, G.taul, G.taur) <= 30 * k * eps * scale
    if dtype == np.float32 and k > 256:
        # hundreds of Float32 two-sided steps: the last reflectors (2 x 2, 3 x 3 trailing blocks) are O(1)-sensitive to the
        # accumulated rounding in BOTH implementations (the Float32 oracle itself is 0.14 away from the Float64 result at
        # 600 x 600), so the elementwise comparison stops at the bidiagonal's magnitudes
        # (measured: |d| of the Float32 ORACLE is 0.14 off the Float64 result at entry 259, the GPU's 0.065; singular values of
        # both agree with svdvals(A) to 4e-6), so the bar for the magnitudes is "no worse than the reference's own arithmetic"
        X = oracle.bidiagonalize(A.astype(np.float64, order="F"))
        for got, ref, exact in ((G.dv, dv, X[3]), (G.ev, ev, X[4])):
            e_ref = float(np.max(np.abs(np.abs(ref.astype(np.float64)) - np.abs(exact))))
            e_got = float(np.max(np.abs(np.abs(got.astype(np.float64)) - np.abs(exact))))
            assert e_got <= 4 * e_ref + 1e-4 * scale
    else:
        X = oracle.bidiagonalize(A.astype(np.float64, order="F")) if dtype == np.float32 else [None] * 5
        _close(G.reflectors, F, dtype, "factors", k, X[0])
        _close(G.taul, tl, dtype, "taul", k, X[1])
        _close(G.taur, tr, dtype, "taur", k, X[2])
        _close(G.dv, dv, dtype, "dv", k, X[3])
        _close(G.ev, ev, dtype, "ev", k, X[4])
    # singular values are invariant (test/svd.jl:110-112 checks svdvals of the bidiagonal against svdvals(A))
    wide = np.complex128 if dtype == np.complex128 else np.float64
    s_ref = np.linalg.svd(A.astype(wide), compute_uv=False)
    s_got = np.linalg.svd(G.bidiagonal.astype(np.float64), compute_uv=False)
    assert np.max(np.abs(np.sort(s_ref) - np.sort(s_got))) <= 50 * TOL[dtype] * 1e-2 * max(1.0, s_ref.max())


def test_bidiagonalize_reference_test_matrix(gla):
    """The 8 x 8 matrix of test/svd.jl:100-112: svdvals(bidiagonal) == svdvals(A)."""
    A = np.array([
        [0.3, 0.0, 0.0, 0.0, 0.0, 0.2, 0.3, 0.0],
        [0.0, 0.0, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0],
        [0.0, -0.2, 0.0, 0.0, 0.0, 0.0, 0.0, -0.2],
        [0.3, 0.0, 0.0, 0.0, 0.0, 0.2, 0.4, 0.0],
        [0.0, 0.4, -0.2, 0.0, 0.0, 0.0, 0.0, 0.3],
        [0.2, 0.0, 0.0, 0.0, 0.0, 0.0, 0.2, 0.0],
        [0.0, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0, 0.0],
        [0.0, 0.3, -0.2, 0.0, 0.0, 0.0, 0.0, 0.3]], order="F")
    G = gla.bidiagonalize_(A.copy(order="F"))
    s_ref = np.linalg.svd(A, compute_uv=False)
    s_got = np.linalg.svd(G.bidiagonal, compute_uv=False)
    assert np.allclose(np.sort(s_ref), np.sort(s_got), rtol=0, atol=1e-14)


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_matrix(gla, oracle, dtype):
    """svd(zeros(2, 2)) of test/svd.jl "Issue 119": reflector! of a zero vector is tau = 0 and leaves it alone."""
    for shape in [(2, 2), (4, 3), (3, 4)]:
        Z = np.zeros(shape, dtype=dtype, order="F")
        G = gla.bidiagonalize_(Z.copy(order="F"))
        assert not G.reflectors.any() and not G.taul.any() and not G.taur.any()
    Z = np.zeros((5, 5), dtype=dtype, order="F")
    H, tau = gla.hessenberg_(Z.copy(order="F"))
    assert not H.any() and not tau.any()
    for uplo in "LU":
        S = gla.symtri_(Z.copy(order="F"), uplo)
        assert not S.factors.any() and not S.tau.any()


@pytest.mark.parametrize("dtype", DTYPES)
def test_rank_deficient_issue_121(gla, oracle, dtype):
    """test/svd.jl "Issue 121": [0 0; 1 -1] and [1 0 0; 0 0 0; 0 1 -1] (zero pivots inside the reduction)."""
    for M in ([[0, 0], [1, -1]], [[1, 0, 0], [0, 0, 0], [0, 1, -1]]):
        A = np.asfortranarray(np.array(M, dtype=dtype))
        F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
        G = gla.bidiagonalize_(A.copy(order="F"))
        _close(G.reflectors, F, dtype, "factors")
        _close(G.taul, tl, dtype, "taul")
        _close(G.taur, tr, dtype, "taur")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 3, 10, 64, 257, 700])
def test_hessenberg_vs_oracle(gla, oracle, dtype, n):
    rng = np.random.default_rng(1000 + n)
    A = _rand(rng, (n, n), dtype)
    F, tau = oracle.hessenberg(A.copy(order="F"))
    G, gtau = gla.hessenberg_(A.copy(order="F"))
    eps = np.finfo(np.float32 if dtype == np.float32 else np.float64).eps
    scale = max(1.0, float(np.max(np.abs(A))))
    assert _hessenberg_residual(A, G, gtau) <= 30 * max(n, 1) * eps * scale * max(1.0, np.sqrt(n))   # Q^H A Q == H
    if not (dtype == np.float32 and n > 256):   # see test_bidiagonalize_vs_oracle
        X = oracle.hessenberg(A.astype(np.float64, order="F")) if dtype == np.float32 else [None] * 2
        _close(G, F, dtype, "factors", n, X[0])
        _close(gtau, tau, dtype, "tau", n, X[1])
    if n == 10:   # test/eigengeneral.jl:239-249: the Hessenberg matrix is unitarily similar to A
        wide = np.complex128
        e_ref = np.linalg.eigvals(A.astype(wide))
        e_got = list(np.linalg.eigvals(np.triu(G, -1).astype(wide)))
        worst = 0.0
        for z in e_ref:
            k = int(np.argmin([abs(z - y) for y in e_got]))
            worst = max(worst, abs(z - e_got.pop(k)))
        assert worst <= (2e-3 if dtype == np.float32 else 1e-10)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("n", [1, 2, 3, 10, 64, 257, 700])
def test_symtri_vs_oracle(gla, oracle, dtype, uplo, n):
    rng = np.random.default_rng(2000 + n)
    A = _rand(rng, (n, n), dtype)
    S = A + A.conj().T
    # only the `uplo` triangle may be read or written: poison the other one
    poison = np.full_like(S, 7)
    Sin = np.asfortranarray(np.tril(S) + np.triu(poison, 1) if uplo == "L" else np.triu(S) + np.tril(poison, -1))
    if dtype == np.complex128 and n > 1:
        Sin[1, 1] += 0.25j   # the imaginary part of the diagonal is ignored (src/eigenSelfAdjoint.jl:458-460)
    F, tau, dv, ev = oracle.symtri(Sin.copy(order="F"), uplo)
    G = gla.symtri_(Sin.copy(order="F"), uplo)
    eps = np.finfo(np.float32 if dtype == np.float32 else np.float64).eps
    scale = max(1.0, float(np.max(np.abs(S))))
    assert _symtri_residual(S, G.factors, G.tau, uplo) <= 30 * max(n, 1) * eps * scale * max(1.0, np.sqrt(n))   # Q^H A Q == T
    if not (dtype == np.float32 and n > 256):   # see test_bidiagonalize_vs_oracle
        X = oracle.symtri(Sin.astype(np.float64, order="F"), uplo) if dtype == np.float32 else [None] * 4
        _close(G.factors, F, dtype, "factors", n, X[0])
        _close(G.tau, tau, dtype, "tau", n, X[1])
        _close(G.dv, dv, dtype, "dv", n, X[2])
        _close(G.ev, ev, dtype, "ev", n, X[3])
    if uplo == "L":
        assert np.array_equal(np.triu(G.factors, 1), np.triu(Sin, 1))
    else:
        assert np.array_equal(np.tril(G.factors, -1), np.tril(Sin, -1))
    # Q' A Q == T (test/eigenselfadjoint.jl:67): the spectrum is invariant
    wide = np.complex128 if dtype == np.complex128 else np.float64
    e_ref = np.linalg.eigvalsh(S.astype(wide))
    e_got = np.linalg.eigvalsh(G.diagonals.astype(np.float64))
    assert np.max(np.abs(e_ref - e_got)) <= 100 * TOL[dtype] * 1e-2 * max(1.0, np.abs(e_ref).max())


@pytest.mark.parametrize("scale", [1e-170, 1e170])
def test_extreme_scaling(gla, oracle, scale):
    """Julia's reflector! takes norm(x), which rescales; sum-of-squares alone would under/overflow here."""
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.standard_normal((40, 24)) * scale)
    F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
    G = gla.bidiagonalize_(A.copy(order="F"))
    assert np.all(np.isfinite(G.reflectors))
    assert np.max(np.abs(G.dv - dv)) <= 1e-10 * np.max(np.abs(dv))
    assert np.max(np.abs(G.taul - tl)) <= 1e-10 and np.max(np.abs(G.taur - tr)) <= 1e-10


def test_large_vs_oracle(gla, oracle):
    """n = 1536: the trailing matrix is far larger than one wave of row blocks / columns (every CTA loops)."""
    rng = np.random.default_rng(77)
    n = 1536
    A = _rand(rng, (n, n), np.float64)
    F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
    G = gla.bidiagonalize_(A.copy(order="F"))
    _close(G.reflectors, F, np.float64, "bidiag factors")
    _close(G.dv, dv, np.float64, "dv")
    S = np.asfortranarray(A + A.T)
    F, tau, dv, ev = oracle.symtri(S.copy(order="F"), "L")
    T = gla.symtri_(S.copy(order="F"), "L")
    _close(T.factors, F, np.float64, "symtri factors")
    Hr, tau = oracle.hessenberg(A.copy(order="F"))
    Hg, gtau = gla.hessenberg_(A.copy(order="F"))
    _close(Hg, Hr, np.float64, "hessenberg factors")


def test_global_slab_path(gla, oracle, monkeypatch):
    """Vectors that do not fit shared memory live in per-CTA slabs of global memory: same results."""
    monkeypatch.setenv("GLA_TS_SMEM_BUDGET", "0")
    rng = np.random.default_rng(9)
    for dtype in DTYPES:
        A = _rand(rng, (130, 70), dtype)
        F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
        G = gla.bidiagonalize_(A.copy(order="F"))
        _close(G.reflectors, F, dtype, "factors", 70)
        B = _rand(rng, (90, 90), dtype)
        S = np.asfortranarray(B + B.conj().T)
        F, tau, dv, ev = oracle.symtri(S.copy(order="F"), "U")
        T = gla.symtri_(S.copy(order="F"), "U")
        _close(T.factors, F, dtype, "symtri", 90)
        Hr, tau = oracle.hessenberg(B.copy(order="F"))
        Hg, gtau = gla.hessenberg_(B.copy(order="F"))
        _close(Hg, Hr, dtype, "hessenberg", 90)
    monkeypatch.delenv("GLA_TS_SMEM_BUDGET")
    # and a column longer than the real budget (30000 x 8 B > 200 KB)
    A = _rand(rng, (30000, 6), np.float64)
    F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
    G = gla.bidiagonalize_(A.copy(order="F"))
    _close(G.reflectors, F, np.float64, "tall factors")


def test_error_paths(gla):
    with pytest.raises(gla.DimensionMismatch):
        gla.hessenberg_(np.zeros((3, 4), order="F"))
    with pytest.raises(gla.DimensionMismatch):
        gla.symtri_(np.zeros((3, 4), order="F"))
    with pytest.raises(gla.ArgumentError):
        gla.symtri_(np.zeros((3, 3), order="F"), "X")
    with pytest.raises(TypeError):
        gla.bidiagonalize_(np.zeros((3, 3), dtype=np.float16, order="F"))


def test_device_twins(gla, oracle):
    """`_dev` entry points on the caller's stream with a padded leading dimension."""
    import torch
    rng = np.random.default_rng(3)
    m, n, lda = 200, 120, 208
    A = _rand(rng, (m, n), np.float64)
    F, tl, tr, dv, ev, _ = oracle.bidiagonalize(A.copy(order="F"))
    buf = torch.zeros((n, lda), dtype=torch.float64, device="cuda")
    buf[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    dl = torch.zeros(n, dtype=torch.float64, device="cuda")
    dr = torch.zeros(n, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gla.bidiagonalize_dev(buf.data_ptr(), m, n, lda, dl.data_ptr(), dr.data_ptr(), st)
    torch.cuda.synchronize()
    got = buf[:, :m].cpu().numpy().T
    _close(got, F, np.float64, "factors")
    _close(dl.cpu().numpy(), tl, np.float64, "taul")
    _close(dr.cpu().numpy()[: n - 1], tr, np.float64, "taur")
    assert float(buf[:, m:].abs().max()) == 0.0   # the padding rows are not touched
    B = _rand(rng, (150, 150), np.float64)
    S = np.asfortranarray(B + B.T)
    for uplo in "LU":
        F, tau, dv, ev = oracle.symtri(S.copy(order="F"), uplo)
        d = torch.from_numpy(np.ascontiguousarray(S.T)).cuda()
        dt = torch.zeros(150, dtype=torch.float64, device="cuda")
        gla.symtri_dev(d.data_ptr(), 150, 150, uplo, dt.data_ptr(), st)
        torch.cuda.synchronize()
        _close(d.cpu().numpy().T, F, np.float64, "symtri " + uplo)
    Hr, tau = oracle.hessenberg(B.copy(order="F"))
    d = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
    dt = torch.zeros(150, dtype=torch.float64, device="cuda")
    gla.hessenberg_dev(d.data_ptr(), 150, 150, dt.data_ptr(), st)
    torch.cuda.synchronize()
    _close(d.cpu().numpy().T, Hr, np.float64, "hessenberg")
    _close(dt.cpu().numpy()[:149], tau, np.float64, "tau")
