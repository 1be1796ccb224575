"""Oracle parity THROUGH the look-ahead / outer-block schedule of the blocked QR (it only engages for m, n > 768) for all
three element types, and a leading-panel check of the n = 16384 results (reference src/qr.jl:113-146: the first k columns
of the factors and the first k taus depend on the first k columns of A only, so qrBlocked! of the leading 16384 x 128
panel is exactly what qrBlocked! of the full matrix leaves there).  Tolerances are north_star's: R elementwise 1e-10
(Float64, ComplexF64) / 1e-4 (Float32) relative, reflectors and tau likewise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-10, np.float32: 1e-4, np.complex128: 1e-10}


def _randn(rng, m, n, dtype):
    A = rng.standard_normal((m, n))
    if dtype == np.complex128:
        A = A + 1j * rng.standard_normal((m, n))
    return np.asfortranarray(A.astype(dtype))


def _check(F, tau, ref_f, ref_t, dtype):
    k = min(ref_f.shape)
    tol = RTOL[dtype]
    scale = np.max(np.abs(ref_f))
    Rg, Rr = np.triu(F)[:k], np.triu(ref_f)[:k]
    assert np.all(np.abs(Rg - Rr) <= tol * np.abs(Rr) + tol * scale)
    Vg, Vr = np.tril(F, -1), np.tril(ref_f, -1)
    assert np.all(np.abs(Vg - Vr) <= tol * np.abs(Vr) + tol * max(1.0, np.max(np.abs(Vr))))
    assert np.max(np.abs(tau - ref_t)) <= 2 * tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128])
@pytest.mark.parametrize("m,n", [(1536, 1536), (2304, 2304), (3000, 1100), (1100, 3000)])
def test_lookahead_schedule_vs_oracle(gla, oracle, dtype, m, n):
    rng = np.random.default_rng(m + 3 * n)
    A = _randn(rng, m, n, dtype)
    # complex oracle = the unblocked reference path (the blocked T build drops a conj at src/qr.jl:72)
    ref_f, ref_t = oracle.qr_unblocked(A) if dtype == np.complex128 else oracle.qr_blocked(A, 12)
    qr = gla.qrBlocked_(A.copy(order="F"))
    _check(qr.factors, qr.tau, ref_f, ref_t, dtype)
    # the reference's own test property (test/qr.jl:20-25) on what came out
    R = np.triu(qr.factors)[:min(m, n)]
    G = A.conj().T @ A
    assert np.max(np.abs(R.conj().T @ R - G)) <= 50 * max(m, n) * np.finfo(A.real.dtype).eps * np.max(np.abs(G))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_n16384_leading_panel_vs_oracle(gla, oracle, dtype):
    import torch
    n, k = 16384, 128
    tdt = {np.float64: torch.float64, np.complex128: torch.complex128, np.float32: torch.float32}[dtype]
    g = torch.Generator(device="cuda").manual_seed(123)
    dA = torch.randn((n, n), generator=g, device="cuda", dtype=tdt)       # dA[j, i] = A[i, j] (column-major storage)
    panel = np.asfortranarray(dA[:k].cpu().numpy().T.copy())               # leading n x k panel of A
    dtau = torch.zeros(n, device="cuda", dtype=tdt)
    gla.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, torch.cuda.current_stream().cuda_stream, dtype)
    torch.cuda.synchronize()
    F = np.asfortranarray(dA[:k].cpu().numpy().T.copy())
    tau = dtau[:k].cpu().numpy()
    ref_f, ref_t = oracle.qr_unblocked(panel) if dtype == np.complex128 else oracle.qr_blocked(panel, 12)
    _check(F, tau, ref_f, ref_t, dtype)
    # and the whole result through the sign-blind Gram identity on a random probe: R^H R x = A^H A x
    x = torch.randn(n, device="cuda", dtype=tdt)
    src = torch.randn((n, n), generator=torch.Generator(device="cuda").manual_seed(123), device="cuda", dtype=tdt)
    R = torch.triu(dA.t())
    y1 = R.conj().t() @ (R @ x)
    y2 = src.conj() @ (src.t() @ x)
    assert ((y1 - y2).abs().max() / y2.abs().max()).item() <= (5e-3 if dtype == np.float32 else 1e-10)
