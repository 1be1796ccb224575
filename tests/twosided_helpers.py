"""numpy (Float64 / ComplexF64) checks shared by the CPU tests of the oracle and the GPU parity tests of the two-sided
reductions: the reflectors are read from the in-place factors with the reference's conventions and applied to the input."""
import numpy as np


def wide(dtype):
    return np.complex128 if dtype == np.complex128 else np.float64


def bidiag_residual(A, F, tl, tr):
    """max |Ql^H A Qr - B| with the reflectors taken from the factors (reference conventions: left application
    (I - conj(tau) v v^H), right application (I - tau u u^H), u stored as reflector!(conj(row)))."""
    W = A.astype(wide(A.dtype)).copy()
    F = F.astype(W.dtype)
    m, n = A.shape
    if m >= n:
        for i in range(n):
            v = np.concatenate(([1.0], F[i + 1:, i]))
            W[i:, i:] -= np.conj(tl[i]) * np.outer(v, v.conj() @ W[i:, i:])
            if i < n - 1:
                u = np.concatenate(([1.0], F[i, i + 2:]))
                W[i:, i + 1:] -= tr[i] * np.outer(W[i:, i + 1:] @ u, u.conj())
        B = np.diag(np.real(np.diagonal(F))[:n]) + (np.diag(np.real(np.diagonal(F, 1)), 1)[:n, :n] if n > 1 else 0)
        return float(np.max(np.abs(W[:n] - B))) if m == n else float(max(np.max(np.abs(W[:n] - B)), np.max(np.abs(W[n:]))))
    for i in range(m):
        u = np.concatenate(([1.0], F[i, i + 1:]))
        W[i:, i:] -= tr[i] * np.outer(W[i:, i:] @ u, u.conj())
        if i < m - 1:
            v = np.concatenate(([1.0], F[i + 2:, i]))
            W[i + 1:, i:] -= np.conj(tl[i]) * np.outer(v, v.conj() @ W[i + 1:, i:])
    B = np.zeros((m, n))
    B[np.arange(m), np.arange(m)] = np.real(np.diagonal(F))[:m]
    if m > 1:
        B[np.arange(1, m), np.arange(m - 1)] = np.real(np.diagonal(F, -1))[:m - 1]
    return float(np.max(np.abs(W - B)))


def hessenberg_residual(A, F, tau):
    W = A.astype(wide(A.dtype)).copy()
    F = F.astype(W.dtype)
    n = A.shape[0]
    for i in range(n - 1):
        v = np.concatenate(([1.0], F[i + 2:, i]))
        W[i + 1:, :] -= np.conj(tau[i]) * np.outer(v, v.conj() @ W[i + 1:, :])
        W[:, i + 1:] -= tau[i] * np.outer(W[:, i + 1:] @ v, v.conj())
    return float(np.max(np.abs(W - np.triu(F, -1))))


def symtri_residual(S, F, tau, uplo):
    """max |Q^H S Q - T| (test/eigenselfadjoint.jl:67), reflectors read from the `uplo` triangle of the factors."""
    n = S.shape[0]
    W = S.astype(wide(S.dtype)).copy()
    F = F.astype(W.dtype)
    if uplo == "U":          # the upper variant is the lower one under the index reversal
        W = W[::-1, ::-1].copy()
        F = F[::-1, ::-1].copy()
    for k in range(min(len(tau), n - 1)):
        v = np.concatenate(([1.0], F[k + 2:, k]))
        W[k + 1:, :] -= np.conj(tau[k]) * np.outer(v, v.conj() @ W[k + 1:, :])
        W[:, k + 1:] -= tau[k] * np.outer(W[:, k + 1:] @ v, v.conj())
    dv = np.real(np.diagonal(F))
    ev = np.real(np.diagonal(F, -1))
    T = np.diag(dv) + (np.diag(ev, 1) + np.diag(ev, -1) if n > 1 else 0)
    return float(np.max(np.abs(W - T)))


