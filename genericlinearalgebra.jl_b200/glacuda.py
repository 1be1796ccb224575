"""Host-side mirror of the reference's operator interface over the C ABI of libgla_cuda.so.

This is what a Julia host does with `ccall` (see julia/GLACuda/ and INTEGRATION.md), written in
Python/ctypes because no Julia runtime exists in this image.  Names and argument meaning follow
the reference (GenericLinearAlgebra.jl v0.4.0); a trailing `_` stands for Julia's `!`:

    qrBlocked_(A, blocksize=12, tau=None) -> QR2          src/qr.jl:113-146
    qrUnblocked_(A, tau=None)            -> QR2          src/qr.jl:86-111
    QR2.R / QR2.QBlocked                                  src/qr.jl:55-83
    HouseholderBlock.lmul_(A) / .adjoint_lmul_(A)         src/householder.jl:82-157
    reflectorApply_(A, x, tau)            (right apply)   src/qr.jl:19-42
    cholRecursive_(A, "L", cutoff=1)                      src/cholesky.jl:37-55
    rankUpdate_(C_lower, A, alpha)                        src/juliaBLAS.jl:89-112

Errors map as in SURVEY.md section 8b: DimensionMismatch / ArgumentError / DomainError.
Arrays are numpy, column-major (order="F"), float32 / float64 / complex128, modified in place.
`*_dev` helpers take raw device pointers (e.g. torch `tensor.data_ptr()`) and a CUDA stream handle.

There is NO CPU fallback: if the shared library is missing, import of the entry points raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLA_CUDA_LIB") or os.path.join(_HERE, "lib", "libgla_cuda.so")   # same override as julia/GLACuda/src/GLACuda.jl
_LIB = None

_I64 = C.c_int64
_PFX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex128): "z"}


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch."""


class ArgumentError(ValueError):
    """Julia's ArgumentError."""


GLA_ERR_NOT_POSDEF = 900
GLA_ERR_SINGULAR = 901


class DomainError(ArithmeticError):
    """Julia's DomainError (sqrt of a negative pivot inside cholRecursive!)."""


class ZeroPivotError(ArithmeticError):
    """ldlt!: a zero pivot (the reference divides by it and carries Inf/NaN on)."""


class GLACudaError(RuntimeError):
    """CUDA / NCCL runtime failure reported by the library (return code >= 1000)."""


def lib():
    """Load libgla_cuda.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GLACudaError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.gla_last_error_string.restype = C.c_char_p
        _LIB.gla_last_device_ms.restype = C.c_double
        _LIB.gla_last_info.restype = C.c_int64
    return _LIB


def _check(rc, what, neg=DimensionMismatch):
    if rc == 0:
        return
    if rc >= 1000:
        raise GLACudaError(f"{what}: {lib().gla_last_error_string().decode()}")
    if rc < 0:
        raise neg(f"{what}: argument {-rc} is illegal")
    if rc == GLA_ERR_NOT_POSDEF:
        k = int(lib().gla_last_info())
        raise DomainError(f"{what}: leading minor {k} is not positive definite (sqrt of a non-positive pivot)", k)
    if rc == GLA_ERR_SINGULAR:
        k = int(lib().gla_last_info())
        raise ZeroPivotError(f"{what}: pivot {k} is zero", k)
    raise GLACudaError(f"{what}: unexpected return code {rc}")


def _fn(name, dtype):
    try:
        p = _PFX[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"only Float32/Float64/ComplexF64 have a GPU method; {dtype} stays on the reference path")
    return getattr(lib(), f"gla_{p}{name}")


def _colmajor(A, what):
    if not isinstance(A, np.ndarray) or A.dtype not in _PFX:
        raise TypeError(f"{what}: need a numpy float32/float64/complex128 array")
    if A.ndim != 2:
        raise DimensionMismatch(f"{what}: need a matrix")
    if A.size and A.shape[0] > 1 and A.strides[0] != A.itemsize:
        raise ArgumentError(f"{what}: need unit row stride (column-major, order='F')")
    if A.size and A.shape[1] > 1 and (A.strides[1] % A.itemsize or A.strides[1] < A.itemsize * A.shape[0]):
        raise ArgumentError(f"{what}: bad column stride")
    ld = A.strides[1] // A.itemsize if A.shape[1] > 1 else max(A.shape[0], 1)
    return max(ld, 1)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def device_count() -> int:
    return lib().gla_device_count()


def set_device(dev: int) -> None:
    _check(lib().gla_set_device(C.c_int(dev)), "gla_set_device")


# ------------------------------------------------------------------------------------ QR objects
class HouseholderBlock:
    """HouseholderBlock{T}(V, T): src/householder.jl:8-11.  V = factors (unit-lower part used)."""

    def __init__(self, V, T, tau):
        self.V, self.T, self.tau = V, T, tau

    def _apply(self, A, adjoint):
        ldv = _colmajor(self.V, "HouseholderBlock.V")
        one_d = A.ndim == 1
        A2 = A.reshape(-1, 1, order="F") if one_d else A
        lda = _colmajor(A2, "lmul!")
        if A2.dtype != self.V.dtype:
            raise TypeError("lmul!: element types differ")
        if A2.shape[0] != self.V.shape[0]:
            raise DimensionMismatch("")  # src/householder.jl:87,129
        rc = _fn("ormqr_blocked", A2.dtype)(_ptr(self.V), _I64(self.V.shape[0]), _I64(self.V.shape[1]), _I64(ldv),
                                            _ptr(self.tau), _ptr(A2), _I64(A2.shape[0]), _I64(A2.shape[1]),
                                            _I64(lda), C.c_int(1 if adjoint else 0))
        _check(rc, "ormqr_blocked")
        return A

    def lmul_(self, A):
        """lmul!(H, A, M): A <- (I - V T V^H) A = Q A.  src/householder.jl:82-115."""
        return self._apply(A, False)

    def adjoint_lmul_(self, A):
        """lmul!(H', A, M): A <- (I - V T^H V^H) A = Q^H A.  src/householder.jl:119-157."""
        return self._apply(A, True)

    def __matmul__(self, A):  # H * A, src/householder.jl:116-117
        return self.lmul_(np.array(A, order="F", copy=True))

    def adjoint_mul(self, A):  # H' * A, src/householder.jl:158-159
        return self.adjoint_lmul_(np.array(A, order="F", copy=True))


class QR2:
    """QR2{T}(factors, tau): src/qr.jl:6-16."""

    def __init__(self, factors, tau):
        self.factors, self.tau = factors, tau

    def thinQ(self):
        """Matrix of the first k = min(m, n) columns of Q = H_1 .. H_k (gla_?orgqr_thin): what
        `F[Tuple{:QBlocked}] * Matrix(I, m, k)` gives in the reference (src/householder.jl:116-117)."""
        lda = _colmajor(self.factors, "thinQ")
        m, n = self.factors.shape
        k = min(m, n)
        Q = np.zeros((m, k), dtype=self.factors.dtype, order="F")
        rc = _fn("orgqr_thin", Q.dtype)(_ptr(self.factors), _I64(m), _I64(n), _I64(lda), _ptr(self.tau), _ptr(Q),
                                        _I64(max(m, 1)))
        _check(rc, "orgqr_thin")
        return Q

    @property
    def shape(self):
        return self.factors.shape

    @property
    def R(self):
        """A[Tuple{:R}]: src/qr.jl:55-62 (ArgumentError for wide input)."""
        m, n = self.factors.shape
        if m < n:
            raise ArgumentError("R matrix is trapezoid and cannot be extracted with indexing")
        return np.triu(self.factors[:n, :n])

    @property
    def QBlocked(self):
        """A[Tuple{:QBlocked}]: compact-WY block of all min(m,n) reflectors.  src/qr.jl:64-83."""
        F = self.factors
        ldf = _colmajor(F, "QBlocked")
        m, n = F.shape
        k = min(m, n)
        T = np.zeros((k, k), dtype=F.dtype, order="F")
        if k:
            rc = _fn("larft", F.dtype)(_ptr(F), _I64(m), _I64(n), _I64(ldf), _ptr(self.tau), _ptr(T), _I64(k))
            _check(rc, "larft")
        return HouseholderBlock(F, T, self.tau)


def qrBlocked_(A, blocksize: int = 12, tau=None) -> QR2:
    """GenericLinearAlgebra.qrBlocked!(A, blocksize, tau): in place, returns QR2(A, tau).

    `blocksize` is a hint (the GPU chooses its own panel width; R, V and tau do not depend on it
    beyond rounding)."""
    lda = _colmajor(A, "qrBlocked!")
    m, n = A.shape
    k = min(m, n)
    if tau is None:
        tau = np.zeros(k, dtype=A.dtype)
    if tau.dtype != A.dtype or tau.size < k:
        raise DimensionMismatch("qrBlocked!: tau too short or of another element type")
    rc = _fn("geqr_blocked", A.dtype)(_ptr(A), _I64(m), _I64(n), _I64(lda), _ptr(tau), _I64(blocksize))
    _check(rc, "geqr_blocked")
    return QR2(A, tau)


def qrUnblocked_(A, tau=None) -> QR2:
    """GenericLinearAlgebra.qrUnblocked!(A, tau): the same factorisation (src/qr.jl:86-111)."""
    return qrBlocked_(A, 0, tau)


def qr(A) -> QR2:
    """Module-local qr(A) = qrBlocked!(copy(A)) (the reference's sign convention, not LAPACK's)."""
    return qrBlocked_(np.array(A, order="F", copy=True))


def reflectorApply_(A, x, tau):
    """right reflectorApply!(A, x, tau): A <- A (I - tau v v^H).  src/qr.jl:19-42."""
    lda = _colmajor(A, "reflectorApply!")
    m, n = A.shape
    x = np.ascontiguousarray(x, dtype=A.dtype)
    if x.size != n:
        raise DimensionMismatch(f"reflector must have same length as second dimension of matrix, but got {x.size} and {n}")
    t = np.array([tau], dtype=A.dtype)
    rc = _fn("reflector_apply_right", A.dtype)(_ptr(A), _I64(m), _I64(n), _I64(lda), _ptr(x), _I64(x.size), _ptr(t))
    _check(rc, "reflector_apply_right")
    return A


def qr_batched_(A, tau=None):
    """`batch` independent qrBlocked! problems.  A: C-contiguous (batch, n, m) buffer holding each
    m x n matrix column-major (i.e. A[b].T is the matrix), factorised in place.
    Returns (A, tau) with tau of shape (batch, min(m,n))."""
    if A.ndim != 3 or not A.flags.c_contiguous or A.dtype not in _PFX:
        raise ArgumentError("qr_batched!: need a C-contiguous (batch, n, m) array")
    batch, n, m = A.shape
    k = min(m, n)
    if tau is None:
        tau = np.zeros((batch, k), dtype=A.dtype)
    rc = _fn("geqr_batched", A.dtype)(_ptr(A), _I64(m), _I64(n), _I64(batch), _ptr(tau))
    _check(rc, "geqr_batched", ArgumentError)
    return A, tau


def qr_batched_ptr(ptr: int, m: int, n: int, batch: int, tau_ptr: int, dtype=np.float64) -> None:
    """Host-POINTER variant (pinned buffers owned by the caller, e.g. torch pinned tensors)."""
    rc = _fn("geqr_batched", dtype)(C.c_void_p(ptr), _I64(m), _I64(n), _I64(batch), C.c_void_p(tau_ptr))
    _check(rc, "geqr_batched", ArgumentError)


def qr_batched_dev(dA: int, m: int, n: int, batch: int, dtau: int, stream: int = 0, dtype=np.float64) -> None:
    """Device-pointer twin, asynchronous on `stream`."""
    rc = _fn("geqr_batched_dev", dtype)(C.c_void_p(dA), _I64(m), _I64(n), _I64(batch), C.c_void_p(dtau),
                                        C.c_void_p(stream))
    _check(rc, "geqr_batched_dev", ArgumentError)


def qr_blocked_dev(dA: int, m: int, n: int, lda: int, dtau: int, blocksize: int = 0, stream: int = 0,
                   dtype=np.float64) -> None:
    rc = _fn("geqr_blocked_dev", dtype)(C.c_void_p(dA), _I64(m), _I64(n), _I64(lda), C.c_void_p(dtau),
                                        _I64(blocksize), C.c_void_p(stream))
    _check(rc, "geqr_blocked_dev")


def qr_blocked_ptr(ptr: int, m: int, n: int, lda: int, tau_ptr: int, blocksize: int = 0, dtype=np.float64) -> None:
    rc = _fn("geqr_blocked", dtype)(C.c_void_p(ptr), _I64(m), _I64(n), _I64(lda), C.c_void_p(tau_ptr), _I64(blocksize))
    _check(rc, "geqr_blocked")


# ------------------------------------------------------------------------------------ TSQR
def tsqr_R(A):
    """R factor (n x n upper) of a tall m x n Float64 matrix by a TSQR tree on one GPU."""
    lda = _colmajor(A, "tsqr")
    m, n = A.shape
    if A.dtype != np.float64:
        raise TypeError("tsqr: Float64 only")
    R = np.zeros((n, n), dtype=np.float64, order="F")
    rc = lib().gla_dtsqr(_ptr(A), _I64(m), _I64(n), _I64(lda), _ptr(R), _I64(n))
    _check(rc, "tsqr", ArgumentError)
    return R


def tsqr_local_dev(dA: int, m: int, n: int, lda: int, dR: int, ldr: int, stream: int = 0) -> None:
    rc = lib().gla_dtsqr_local_dev(C.c_void_p(dA), _I64(m), _I64(n), _I64(lda), C.c_void_p(dR), _I64(ldr),
                                   C.c_void_p(stream))
    _check(rc, "tsqr_local_dev", ArgumentError)


def tsqr_combine_dev(dRs: int, count: int, n: int, dR: int, ldr: int, stream: int = 0) -> None:
    rc = lib().gla_dtsqr_combine_dev(C.c_void_p(dRs), _I64(count), _I64(n), C.c_void_p(dR), _I64(ldr),
                                     C.c_void_p(stream))
    _check(rc, "tsqr_combine_dev", ArgumentError)


class TsqrComm:
    """NCCL communicator owned by the library for the TSQR R-factor exchange (one process per GPU).
    `broadcast_bytes(bytes_or_None) -> bytes` is the host framework's broadcast from rank 0 (e.g. over
    torch.distributed); it carries the 128-byte ncclUniqueId."""

    def __init__(self, rank: int, nranks: int, broadcast_bytes):
        uid = C.create_string_buffer(128)
        if rank == 0:
            _check(lib().gla_nccl_unique_id(uid), "nccl_unique_id", ArgumentError)
        raw = broadcast_bytes(bytes(uid.raw) if rank == 0 else None)
        uid = C.create_string_buffer(bytes(raw), 128)
        self.comm = C.c_void_p()
        self.nranks = nranks
        _check(lib().gla_nccl_comm_init(C.byref(self.comm), C.c_int(nranks), uid, C.c_int(rank)), "nccl_comm_init",
               ArgumentError)

    def allreduce_R(self, dRloc: int, n: int, dstack: int, dR: int, ldr: int, stream: int = 0) -> None:
        rc = lib().gla_dtsqr_allreduce_dev(self.comm, C.c_int(self.nranks), C.c_void_p(dRloc), _I64(n), C.c_void_p(dstack),
                                           C.c_void_p(dR), _I64(ldr), C.c_void_p(stream))
        _check(rc, "tsqr_allreduce_dev", ArgumentError)

    def destroy(self) -> None:
        if self.comm:
            lib().gla_nccl_comm_destroy(self.comm)
            self.comm = C.c_void_p()


# ------------------------------------------------------------------------------------ Cholesky
def cholRecursive_(A, uplo="L", cutoff: int = 1):
    """GenericLinearAlgebra.cholRecursive!(A, Val{:L}, cutoff): lower Cholesky in place; the strict
    upper triangle of A is left untouched.  Returns A (LowerTriangular(A) in the reference)."""
    if uplo not in ("L", ":L"):
        raise ArgumentError("only Val{:L} has a method (src/cholesky.jl:37)")
    lda = _colmajor(A, "cholRecursive!")
    if A.shape[0] != A.shape[1]:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")  # checksquare
    rc = _fn("potrf_recursive_L", A.dtype)(_ptr(A), _I64(A.shape[0]), _I64(lda), _I64(cutoff))
    _check(rc, "potrf_recursive_L")
    return A


def cholUnblocked_(A, uplo="L"):
    """GenericLinearAlgebra.cholUnblocked!(A, Val{:L}) (src/cholesky.jl:3-15): the same unique lower factor, in place."""
    if uplo not in ("L", ":L"):
        raise ArgumentError("only Val{:L} has a method (src/cholesky.jl:3)")
    lda = _colmajor(A, "cholUnblocked!")
    if A.shape[0] != A.shape[1]:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    _check(_fn("potrf_unblocked_L", A.dtype)(_ptr(A), _I64(A.shape[0]), _I64(lda)), "potrf_unblocked_L")
    return A


def cholBlocked_(A, uplo="L", blocksize: int = 64):
    """GenericLinearAlgebra.cholBlocked!(A, Val{:L}, blocksize) (src/cholesky.jl:17-35); blocksize is a hint."""
    if uplo not in ("L", ":L"):
        raise ArgumentError("only Val{:L} has a method (src/cholesky.jl:17)")
    lda = _colmajor(A, "cholBlocked!")
    if A.shape[0] != A.shape[1]:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    _check(_fn("potrf_blocked_L", A.dtype)(_ptr(A), _I64(A.shape[0]), _I64(lda), _I64(blocksize)), "potrf_blocked_L",
           ArgumentError)
    return A


def ldlt_(A, uplo="L", blocksize=None):
    """LinearAlgebra.ldlt!(Hermitian(A, uplo), blocksize) of the reference (src/ldlt.jl:155-162), without pivoting:
    in place, D on the diagonal, the unit factor in the strict `uplo` triangle.  Float32 / Float64 / ComplexF64 (Hermitian)."""
    lda = _colmajor(A, "ldlt!")
    if A.shape[0] != A.shape[1]:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    u = uplo[-1]
    if u not in ("L", "U"):
        raise ArgumentError("uplo must be :L or :U")
    bs = max(1, 128 // A.dtype.itemsize) if blocksize is None else int(blocksize)
    rc = _fn("ldlt", A.dtype)(_ptr(A), _I64(A.shape[0]), _I64(lda), C.c_int(ord(u)), _I64(bs))
    _check(rc, "ldlt", ArgumentError)
    return A


OP_GEQR_BLOCKED, OP_POTRF_L, OP_GEQR_BATCHED, OP_TSQR, OP_LDLT, OP_BIDIAGONALIZE, OP_HESSENBERG, OP_SYMTRI = 1, 2, 3, 4, 5, 6, 7, 8


def workspace_query(op: int, dtype, m: int, n: int) -> int:
    """Device bytes the `_dev` call of `op` takes from the library's stream-ordered pool (gla_workspace_query)."""
    f = lib().gla_workspace_query
    f.restype = C.c_int64
    rc = int(f(C.c_int(op), C.c_int(np.dtype(dtype).itemsize), _I64(m), _I64(n)))
    if rc < 0:
        raise ArgumentError(f"workspace_query: argument {-rc} is illegal")
    return rc


def larft_dev(dF: int, m: int, n: int, ldf: int, dtau: int, dT: int, ldt: int, stream: int = 0, dtype=np.float64) -> None:
    rc = _fn("larft_dev", dtype)(C.c_void_p(dF), _I64(m), _I64(n), _I64(ldf), C.c_void_p(dtau), C.c_void_p(dT), _I64(ldt),
                                 C.c_void_p(stream))
    _check(rc, "larft_dev")


def ormqr_blocked_dev(dF: int, mF: int, nF: int, ldf: int, dtau: int, dA: int, mA: int, nA: int, lda: int, adjoint: bool,
                      stream: int = 0, dtype=np.float64) -> None:
    rc = _fn("ormqr_blocked_dev", dtype)(C.c_void_p(dF), _I64(mF), _I64(nF), _I64(ldf), C.c_void_p(dtau), C.c_void_p(dA),
                                         _I64(mA), _I64(nA), _I64(lda), C.c_int(1 if adjoint else 0), C.c_void_p(stream))
    _check(rc, "ormqr_blocked_dev")


def reflector_apply_right_dev(dA: int, m: int, n: int, lda: int, dx: int, lenx: int, tau, stream: int = 0,
                              dtype=np.float64) -> None:
    t = np.array([tau], dtype=dtype)
    rc = _fn("reflector_apply_right_dev", dtype)(C.c_void_p(dA), _I64(m), _I64(n), _I64(lda), C.c_void_p(dx), _I64(lenx),
                                                 _ptr(t), C.c_void_p(stream))
    _check(rc, "reflector_apply_right_dev")


def chol_recursive_dev(dA: int, n: int, lda: int, dinfo: int, cutoff: int = 1, stream: int = 0,
                       dtype=np.float64) -> None:
    rc = _fn("potrf_recursive_L_dev", dtype)(C.c_void_p(dA), _I64(n), _I64(lda), _I64(cutoff), C.c_void_p(dinfo),
                                             C.c_void_p(stream))
    _check(rc, "potrf_recursive_L_dev")


def rankUpdate_(Cm, A, alpha=-1.0):
    """rankUpdate!(Hermitian(C, :L), A, alpha): lower triangle of C += alpha*A*A^H.  src/juliaBLAS.jl:89-112."""
    ldc = _colmajor(Cm, "rankUpdate!")
    if Cm.shape[0] != Cm.shape[1]:
        raise DimensionMismatch("rankUpdate!: C is not square")
    A2 = A.reshape(-1, 1, order="F") if A.ndim == 1 else A
    lda = _colmajor(A2, "rankUpdate!")
    if A2.shape[0] != Cm.shape[0] or A2.dtype != Cm.dtype:
        raise DimensionMismatch("rankUpdate!: A and C do not match")
    name = "herk_lower" if Cm.dtype == np.complex128 else "syrk_lower"
    fn = _fn(name, Cm.dtype)
    rt = C.c_float if Cm.dtype == np.float32 else C.c_double
    fn.argtypes = [C.c_void_p, _I64, _I64, C.c_void_p, _I64, _I64, rt]
    rc = fn(Cm.ctypes.data, Cm.shape[0], ldc, A2.ctypes.data, A2.shape[1], lda, float(alpha))
    _check(rc, name)
    return Cm


# ---------------------------------------------------------------------------------------------- two-sided reductions
class BidiagonalFactorization:
    """Mirror of the reference's BidiagonalFactorization (src/svd.jl:321-326): `bidiagonal` = (dv, ev, uplo),
    `reflectors` = the in-place array, `taul`, `taur`."""

    def __init__(self, dv, ev, uplo, reflectors, taul, taur):
        self.dv, self.ev, self.uplo = dv, ev, uplo
        self.reflectors, self.taul, self.taur = reflectors, taul, taur

    @property
    def bidiagonal(self):
        k = self.dv.size
        B = np.diag(self.dv)
        if k > 1:
            B = B + np.diag(self.ev, 1 if self.uplo == "U" else -1)
        return B


def bidiagonalize_(A) -> BidiagonalFactorization:
    """bidiagonalize!(A) (src/svd.jl:328-381) on the GPU: upper bidiagonal for m >= n, lower for m < n."""
    lda = _colmajor(A, "bidiagonalize!")
    m, n = A.shape
    nl, nr = (n, max(n - 1, 0)) if m >= n else (max(m - 1, 0), m)
    taul = np.zeros(max(nl, 1), dtype=A.dtype)
    taur = np.zeros(max(nr, 1), dtype=A.dtype)
    rc = _fn("bidiagonalize", A.dtype)(_ptr(A), _I64(m), _I64(n), _I64(lda), _ptr(taul), _ptr(taur))
    _check(rc, "bidiagonalize")
    k = min(m, n)
    off = 1 if m >= n else -1
    dv = np.real(np.diagonal(A)[:k]).copy()
    ev = np.real(np.diagonal(A, off)).copy()[:(nr if m >= n else nl)]
    return BidiagonalFactorization(dv, ev, "U" if m >= n else "L", A, taul[:nl], taur[:nr])


def hessenberg_(A):
    """_hessenberg!(A) (src/eigenGeneral.jl:18-31) on the GPU: returns (factors, tau); H = triu(factors, -1)."""
    lda = _colmajor(A, "hessenberg!")
    n = A.shape[0]
    if A.shape[1] != n:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")   # checksquare, :19
    tau = np.zeros(max(n - 1, 1), dtype=A.dtype)
    rc = _fn("hessenberg", A.dtype)(_ptr(A), _I64(n), _I64(lda), _ptr(tau))
    _check(rc, "hessenberg")
    return A, tau[:max(n - 1, 0)]


class SymmetricTridiagonalFactorization:
    """Mirror of the reference's SymmetricTridiagonalFactorization (src/eigenSelfAdjoint.jl): `reflectors` =
    (uplo, factors, tau) as in EigenQ, `diagonals` = (dv, ev)."""

    def __init__(self, uplo, factors, tau, dv, ev):
        self.uplo, self.factors, self.tau, self.dv, self.ev = uplo, factors, tau, dv, ev

    @property
    def diagonals(self):
        T = np.diag(self.dv)
        if self.dv.size > 1:
            T = T + np.diag(self.ev, 1) + np.diag(self.ev, -1)
        return T


def symtri_(A, uplo="L") -> SymmetricTridiagonalFactorization:
    """symtri!(Hermitian(A, uplo)) (src/eigenSelfAdjoint.jl:446-564) on the GPU; only the `uplo` triangle is touched."""
    lda = _colmajor(A, "symtri!")
    n = A.shape[0]
    if A.shape[1] != n:
        raise DimensionMismatch(f"matrix is not square: dimensions are {A.shape}")
    u = uplo[-1]
    if u not in ("L", "U"):
        raise ArgumentError("uplo must be :L or :U")
    tau = np.zeros(max(n - 1, 1), dtype=A.dtype)
    rc = _fn("symtri", A.dtype)(_ptr(A), _I64(n), _I64(lda), C.c_int(ord(u)), _ptr(tau))
    _check(rc, "symtri")
    dv = np.real(np.diagonal(A)).copy()
    ev = np.real(np.diagonal(A, 1 if u == "U" else -1)).copy()
    return SymmetricTridiagonalFactorization(u, A, tau[:max(n - 1, 0)], dv, ev)


def bidiagonalize_dev(dA: int, m: int, n: int, lda: int, dtaul: int, dtaur: int, stream: int = 0, dtype=np.float64) -> None:
    rc = _fn("bidiagonalize_dev", dtype)(C.c_void_p(dA), _I64(m), _I64(n), _I64(lda), C.c_void_p(dtaul), C.c_void_p(dtaur),
                                         C.c_void_p(stream))
    _check(rc, "bidiagonalize_dev")


def hessenberg_dev(dA: int, n: int, lda: int, dtau: int, stream: int = 0, dtype=np.float64) -> None:
    rc = _fn("hessenberg_dev", dtype)(C.c_void_p(dA), _I64(n), _I64(lda), C.c_void_p(dtau), C.c_void_p(stream))
    _check(rc, "hessenberg_dev")


def symtri_dev(dA: int, n: int, lda: int, uplo: str, dtau: int, stream: int = 0, dtype=np.float64) -> None:
    rc = _fn("symtri_dev", dtype)(C.c_void_p(dA), _I64(n), _I64(lda), C.c_int(ord(uplo[-1])), C.c_void_p(dtau),
                                  C.c_void_p(stream))
    _check(rc, "symtri_dev")
