"""ctypes front end of libgla_oracle.so (C++ restatement of the reference, see gla_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- never imported by the product path.
All matrices are numpy arrays in Fortran (column-major) order, as Julia stores them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_PFX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex128): "z"}
_I64 = C.c_int64


def build(force: bool = False) -> str:
    """Compile the oracle (g++ only; no reference sources are involved)."""
    so = os.path.join(_HERE, "libgla_oracle.so")
    src = os.path.join(_HERE, "gla_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgla_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libgla_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
    return _LIB


def max_threads() -> int:
    """OpenMP team size the oracle's parallel loops will use."""
    return int(lib().oracle_max_threads())


def set_threads(n: int) -> None:
    lib().oracle_set_threads(C.c_int(int(n)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    if a.dtype not in _PFX:
        raise TypeError(f"oracle supports float32/float64/complex128, got {a.dtype}")
    return np.array(a, order="F", copy=True)


def _fn(name, a):
    return getattr(lib(), f"oracle_{_PFX[a.dtype]}{name}")


def reflector(x):
    """LinearAlgebra.reflector!(x) -> (x_out, tau).  Call site src/qr.jl:96."""
    x = _f(x)
    tau = np.zeros(1, dtype=x.dtype)
    _fn("reflector", x)(_p(x), _I64(x.size), _p(tau))
    return x, tau[0]


def reflector_apply_left(x, tau, A):
    """left LinearAlgebra.reflectorApply!(x, tau, A).  Call site src/qr.jl:102."""
    A = _f(A)
    x = _f(x, A.dtype)
    if x.size != A.shape[0]:
        raise ValueError("DimensionMismatch")
    t = np.array([tau], dtype=A.dtype)
    _fn("reflector_apply_left", A)(_p(x), _p(t), _p(A), _I64(A.shape[0]), _I64(A.shape[1]),
                                   _I64(A.shape[0]))
    return A


def reflector_apply_right(A, x, tau):
    """right reflectorApply!(A, x, tau), src/qr.jl:19-42; ValueError == DimensionMismatch."""
    A = _f(A)
    x = _f(x, A.dtype)
    t = np.array([tau], dtype=A.dtype)
    rc = _fn("reflector_apply_right", A)(_p(A), _I64(A.shape[0]), _I64(A.shape[1]),
                                         _I64(A.shape[0]), _p(x), _I64(x.size), _p(t))
    if rc != 0:
        raise ValueError("DimensionMismatch: reflector must have same length as second dimension of matrix")
    return A


def qr_unblocked(A):
    """qrUnblocked!(A) -> (factors, tau).  src/qr.jl:86-111."""
    A = _f(A)
    m, n = A.shape
    tau = np.zeros(min(m, n), dtype=A.dtype)
    _fn("qr_unblocked", A)(_p(A), _I64(m), _I64(n), _I64(m), _p(tau))
    return A, tau


def qr_blocked(A, blocksize=12, literal=False):
    """qrBlocked!(A, blocksize) -> (factors, tau).  src/qr.jl:113-146.

    literal=True reproduces the reference's missing conj at src/qr.jl:72 (wrong for complex
    inputs with more than one panel; identical for real types)."""
    A = _f(A)
    m, n = A.shape
    tau = np.zeros(min(m, n), dtype=A.dtype)
    _fn("qr_blocked", A)(_p(A), _I64(m), _I64(n), _I64(m), _p(tau), _I64(blocksize),
                         C.c_int(1 if literal else 0))
    return A, tau


def build_T(F, tau, literal=False):
    """getindex(::QR2, Tuple{:QBlocked}) -> T (k x k upper).  src/qr.jl:64-83."""
    F = _f(F)
    m, n = F.shape
    k = min(m, n)
    tau = _f(tau, F.dtype)
    T = np.zeros((k, k), dtype=F.dtype, order="F")
    _fn("build_T", F)(_p(F), _I64(m), _I64(n), _I64(m), _p(tau), _p(T), _I64(k),
                      C.c_int(1 if literal else 0))
    return T


def block_apply(V, T, A, adjoint=False):
    """lmul!(H, A, M) / lmul!(H', A, M) for H = HouseholderBlock(V, T).  src/householder.jl:82-157."""
    V = _f(V)
    T = _f(T, V.dtype)
    A = _f(A, V.dtype)
    if A.ndim == 1:
        A = A.reshape(-1, 1, order="F")
    rc = _fn("block_apply", V)(_p(V), _I64(V.shape[0]), _I64(V.shape[1]), _I64(V.shape[0]), _p(T),
                               _I64(T.shape[0]), _p(A), _I64(A.shape[0]), _I64(A.shape[1]),
                               _I64(A.shape[0]), C.c_int(1 if adjoint else 0))
    if rc != 0:
        raise ValueError("DimensionMismatch")
    return A


def qr_batched(A, blocksize=12):
    """Independent qrBlocked! on A[b] (batch, m, n) -- returns (factors (batch,m,n), tau (batch,k)).

    Storage: each matrix column-major, matrices contiguous (what the C ABI takes)."""
    A = np.asarray(A)
    batch, m, n = A.shape
    k = min(m, n)
    buf = np.array(np.transpose(A, (0, 2, 1)), order="C", copy=True)  # (batch, n, m) C-order == col-major mats
    tau = np.zeros((batch, k), dtype=A.dtype)
    getattr(lib(), f"oracle_{_PFX[buf.dtype]}qr_batched")(_p(buf), _I64(m), _I64(n), _I64(batch),
                                                          _p(tau), _I64(blocksize))
    return np.transpose(buf, (0, 2, 1)), tau


def qr_batched_raw(buf, m, n, batch, tau, blocksize=12):
    """In-place variant on a raw contiguous buffer (used for CPU-baseline timing)."""
    getattr(lib(), f"oracle_{_PFX[buf.dtype]}qr_batched")(_p(buf), _I64(m), _I64(n), _I64(batch),
                                                          _p(tau), _I64(blocksize))


def rank_update_lower(Cm, A, alpha=-1.0, mt=False):
    """rankUpdate!(Hermitian(C,:L), A, alpha) generic method.  src/juliaBLAS.jl:89-112."""
    Cm = _f(Cm)
    A = _f(A, Cm.dtype)
    if A.ndim == 1:
        A = A.reshape(-1, 1, order="F")
    fn = _fn("rank_update_lower", Cm)
    rt = C.c_float if Cm.dtype == np.float32 else C.c_double
    fn.argtypes = [C.c_void_p, _I64, _I64, C.c_void_p, _I64, _I64, rt, C.c_int]
    fn(_p(Cm), Cm.shape[0], Cm.shape[0], _p(A), A.shape[1], A.shape[0], alpha, 1 if mt else 0)
    return Cm


def rdiv_lower_adjoint(X, L):
    """rdiv!(X, LowerTriangular(L)')  (call site src/cholesky.jl:48)."""
    X = _f(X)
    L = _f(L, X.dtype)
    _fn("rdiv_lower_adjoint", X)(_p(X), _I64(X.shape[0]), _I64(X.shape[1]), _I64(X.shape[0]), _p(L),
                                 _I64(L.shape[0]))
    return X


class DomainError(ArithmeticError):
    """sqrt of a non-positive pivot (what Julia's sqrt throws inside cholRecursive!)."""


def _chol(name, A, *args):
    A = _f(A)
    n = A.shape[0]
    if A.shape[1] != n:
        raise ValueError("DimensionMismatch: matrix is not square")
    info = _fn(name, A)(_p(A), _I64(n), _I64(n), *args)
    if info != 0:
        raise DomainError(f"leading minor {info} is not positive definite")
    return A


def chol_unblocked(A):
    """cholUnblocked!(A, Val{:L}).  src/cholesky.jl:3-15.  Returns the full in-place array."""
    return _chol("chol_unblocked", A)


def chol_blocked(A, blocksize):
    """cholBlocked!(A, Val{:L}, blocksize).  src/cholesky.jl:17-35."""
    return _chol("chol_blocked", A, _I64(blocksize))


def chol_recursive(A, cutoff=1, mt=False):
    """cholRecursive!(A, Val{:L}, cutoff).  src/cholesky.jl:37-55 (strict upper untouched)."""
    return _chol("chol_recursive", A, _I64(cutoff), C.c_int(1 if mt else 0))


def ldlt(A, uplo="L", blocksize=None):
    """ldlt!(Hermitian(A, uplo), blocksize) (src/ldlt.jl:155-162 -> _ldlt_lower_blocked! / _ldlt_upper_blocked!):
    in place, D on the diagonal, the unit factor in the strict `uplo` triangle; the other triangle is untouched.
    Default blocksize = max(1, 128 / sizeof(T)) as in the reference."""
    A = _f(A)
    n = A.shape[0]
    if A.shape[1] != n:
        raise ValueError("DimensionMismatch: matrix is not square")
    bs = max(1, 128 // A.dtype.itemsize) if blocksize is None else int(blocksize)
    _fn("ldlt", A)(_p(A), _I64(n), _I64(n), _I64(bs), C.c_int(1 if uplo in ("U", ":U") else 0))
    return A


def bidiagonalize(A):
    """bidiagonalize!(A) (src/svd.jl:328-381).  Returns (factors, taul, taur, dv, ev, uplo): the reflectors in place,
    the bidiagonal as real(diag(A)) and real(diag(A, +-1)), uplo 'U' for m >= n and 'L' for m < n."""
    A = _f(A)
    m, n = A.shape
    k = min(m, n)
    nl, nr = (n, max(n - 1, 0)) if m >= n else (max(m - 1, 0), m)
    taul = np.zeros(max(nl, 1), dtype=A.dtype)
    taur = np.zeros(max(nr, 1), dtype=A.dtype)
    _fn("bidiagonalize", A)(_p(A), _I64(m), _I64(n), _I64(max(m, 1)), _p(taul), _p(taur))
    off = 1 if m >= n else -1
    ev = np.real(np.diagonal(A, off)).copy()[:(nr if m >= n else nl)]
    return A, taul[:nl], taur[:nr], np.real(np.diagonal(A)[:k]).copy(), ev, ("U" if m >= n else "L")


def hessenberg(A):
    """_hessenberg!(A) (src/eigenGeneral.jl:18-31): reflectors below the first subdiagonal, tau of length n-1."""
    A = _f(A)
    n = A.shape[0]
    if A.shape[1] != n:
        raise ValueError("DimensionMismatch: matrix is not square")
    tau = np.zeros(max(n - 1, 1), dtype=A.dtype)
    _fn("hessenberg", A)(_p(A), _I64(n), _I64(max(n, 1)), _p(tau))
    return A, tau[:max(n - 1, 0)]


def symtri(A, uplo="L"):
    """symtri!(Hermitian(A, uplo)) (src/eigenSelfAdjoint.jl:446-564): returns (factors, tau, dv, ev); tau has n-1
    entries (the last one stays 0 for real element types, which stop one step earlier)."""
    A = _f(A)
    n = A.shape[0]
    if A.shape[1] != n:
        raise ValueError("DimensionMismatch: matrix is not square")
    up = uplo in ("U", ":U")
    tau = np.zeros(max(n - 1, 1), dtype=A.dtype)
    _fn("symtri", A)(_p(A), _I64(n), _I64(max(n, 1)), _p(tau), C.c_int(1 if up else 0))
    return A, tau[:max(n - 1, 0)], np.real(np.diagonal(A)).copy(), np.real(np.diagonal(A, 1 if up else -1)).copy()
