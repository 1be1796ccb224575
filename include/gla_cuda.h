/* gla_cuda.h -- C ABI of libgla_cuda.so: the B200 (sm_100a) Householder-QR / Cholesky-update
 * hot path of GenericLinearAlgebra.jl.
 *
 * Every entry point replaces one reference interface (citations are file:line into the
 * reference tree, GenericLinearAlgebra.jl v0.4.0).  A Julia host reaches them with `ccall`
 * (see INTEGRATION.md and genericlinearalgebra.jl_b200/julia/GLACuda/); the Python
 * ctypes mirror used by tests/bench is genericlinearalgebra.jl_b200/glacuda.py.
 *
 * Conventions
 *  - matrices are column-major with unit row stride, leading dimension `lda` in ELEMENTS,
 *    complex is interleaved (re,im) == Julia ComplexF64 == C `double _Complex`;
 *  - results are written in place exactly as the reference leaves them:
 *      QR:   R in the upper triangle incl. diagonal, Householder vectors below it (implicit
 *            unit leading 1), tau[0..min(m,n));  sign/tau convention of Julia's
 *            LinearAlgebra.reflector!: R[k,k] = -copysign(||x||, Re x1), a length-1 column
 *            is still reflected (tau = 2);
 *      Chol: lower triangle incl. diagonal; the strict upper triangle is NOT touched;
 *  - prefix s/d/z = Float32 / Float64 / ComplexF64;
 *  - plain functions take HOST pointers and are synchronous (H2D, compute, D2H inside);
 *    `_dev` twins take DEVICE pointers plus a cudaStream_t (passed as void*) and are
 *    asynchronous on that stream;
 *  - return value: 0 ok; -k = k-th argument illegal (shim throws DimensionMismatch /
 *    ArgumentError); GLA_ERR_NOT_POSDEF (900) from potrf = a leading minor is not positive
 *    definite, its 1-based index is returned OUT OF BAND by gla_last_info() (shim throws
 *    DomainError, like sqrt of a negative real at src/cholesky.jl:40; the index can exceed any
 *    fixed code range, so it is never folded into the return value); >= 1000 = CUDA/NCCL
 *    runtime failure, text via gla_last_error_string().
 *  - magnitude range: the generic small-matrix QR kernel (batched shapes other than real
 *    32x32, ComplexF64, TSQR tree nodes) takes the scaled column norm like Julia's norm(x)
 *    (columns around 1e-160 / 1e160, 1e-20 / 1e20 in Float32, are rescaled by an exact power
 *    of two).  The 32x32 register kernel of the batched QR, the panel kernel of the blocked QR
 *    and the TSQR streaming kernel square the entries unscaled: every column (and column tail)
 *    must satisfy 1e-140 < ||x|| < 1e140 (1e-15 .. 1e15 in Float32) or be exactly zero; outside
 *    that range scale the matrix by a power of two first (R scales with it, V and tau do not).
 *  - there is NO CPU fallback anywhere in this library.
 */
#ifndef GLA_CUDA_H
#define GLA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GLA_API __attribute__((visibility("default")))
#else
#define GLA_API
#endif

#define GLA_ERR_NOT_POSDEF 900
#define GLA_ERR_SINGULAR 901   /* ldlt: zero pivot, 1-based index in gla_last_info() */
#define GLA_ERR_CUDA 1000
#define GLA_ERR_NCCL 2000

/* ---- library ------------------------------------------------------------------- */
GLA_API int gla_version(void);                       /* 100*major + minor */
GLA_API int gla_device_count(void);                  /* number of visible CUDA devices, <0 on error */
GLA_API const char* gla_last_error_string(void);     /* thread-local text of the last >=1000 error */
GLA_API int gla_set_device(int device);              /* device used by host-pointer entry points */
/* thread-local detail of the last GLA_ERR_NOT_POSDEF: 1-based index of the leading minor that failed */
GLA_API int64_t gla_last_info(void);
/* average device time (ms) of the compute part of the last host-pointer call on this thread */
GLA_API double gla_last_device_ms(void);

/* ---- blocked Householder QR ------------------------------------------------------
 * replaces GenericLinearAlgebra.qrBlocked!(A, blocksize, tau, work)   src/qr.jl:113-146
 * (panel = qrUnblocked! src/qr.jl:86-111 with stdlib reflector!/reflectorApply! call sites
 * src/qr.jl:96,102; T build src/qr.jl:64-83; trailing update src/householder.jl:119-157).
 * `blocksize_hint` <= 0 lets the library choose; results do not depend on it beyond rounding. */
GLA_API int gla_sgeqr_blocked(float* A, int64_t m, int64_t n, int64_t lda, float* tau, int64_t blocksize_hint);
GLA_API int gla_dgeqr_blocked(double* A, int64_t m, int64_t n, int64_t lda, double* tau, int64_t blocksize_hint);
GLA_API int gla_zgeqr_blocked(void* A, int64_t m, int64_t n, int64_t lda, void* tau, int64_t blocksize_hint);
GLA_API int gla_sgeqr_blocked_dev(float* dA, int64_t m, int64_t n, int64_t lda, float* dtau, int64_t blocksize_hint, void* stream);
GLA_API int gla_dgeqr_blocked_dev(double* dA, int64_t m, int64_t n, int64_t lda, double* dtau, int64_t blocksize_hint, void* stream);
GLA_API int gla_zgeqr_blocked_dev(void* dA, int64_t m, int64_t n, int64_t lda, void* dtau, int64_t blocksize_hint, void* stream);

/* ---- compact-WY T factor ---------------------------------------------------------
 * replaces getindex(::QR2, Tuple{:QBlocked})   src/qr.jl:64-83  (with the conj the reference
 * omits at :72 for complex types).  F = factors (m x n, V below the diagonal), tau[k],
 * k = min(m,n); T is k x k upper triangular, ldt >= k, strict lower part set to zero. */
GLA_API int gla_slarft(const float* F, int64_t m, int64_t n, int64_t ldf, const float* tau, float* T, int64_t ldt);
GLA_API int gla_dlarft(const double* F, int64_t m, int64_t n, int64_t ldf, const double* tau, double* T, int64_t ldt);
GLA_API int gla_zlarft(const void* F, int64_t m, int64_t n, int64_t ldf, const void* tau, void* T, int64_t ldt);
GLA_API int gla_slarft_dev(const float* dF, int64_t m, int64_t n, int64_t ldf, const float* dtau, float* dT, int64_t ldt, void* stream);
GLA_API int gla_dlarft_dev(const double* dF, int64_t m, int64_t n, int64_t ldf, const double* dtau, double* dT, int64_t ldt, void* stream);
GLA_API int gla_zlarft_dev(const void* dF, int64_t m, int64_t n, int64_t ldf, const void* dtau, void* dT, int64_t ldt, void* stream);

/* ---- block reflector application ---------------------------------------------------
 * replaces lmul!(H, A, M) (adjoint = 0, A <- Q A, src/householder.jl:82-115) and
 * lmul!(H', A, M) (adjoint = 1, A <- Q^H A, src/householder.jl:119-157) where
 * H = HouseholderBlock(F, T) built from all k = min(mF,nF) reflectors of F.
 * returns -1..: DimensionMismatch when mF != mA (src/householder.jl:87,129). */
GLA_API int gla_sormqr_blocked(const float* F, int64_t mF, int64_t nF, int64_t ldf, const float* tau, float* A, int64_t mA, int64_t nA, int64_t lda, int adjoint);
GLA_API int gla_dormqr_blocked(const double* F, int64_t mF, int64_t nF, int64_t ldf, const double* tau, double* A, int64_t mA, int64_t nA, int64_t lda, int adjoint);
GLA_API int gla_zormqr_blocked(const void* F, int64_t mF, int64_t nF, int64_t ldf, const void* tau, void* A, int64_t mA, int64_t nA, int64_t lda, int adjoint);
GLA_API int gla_sormqr_blocked_dev(const float* dF, int64_t mF, int64_t nF, int64_t ldf, const float* dtau, float* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream);
GLA_API int gla_dormqr_blocked_dev(const double* dF, int64_t mF, int64_t nF, int64_t ldf, const double* dtau, double* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream);
GLA_API int gla_zormqr_blocked_dev(const void* dF, int64_t mF, int64_t nF, int64_t ldf, const void* dtau, void* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream);

/* ---- thin Q --------------------------------------------------------------------------
 * Q (m x k, k = min(m,n), ldq >= m) = H_1 .. H_k [I_k; 0]: the product `HouseholderBlock * Matrix(I, m, k)` of the
 * reference (src/householder.jl:116-117 with T from src/qr.jl:64-83), formed with 384-reflector-wide block applies. */
GLA_API int gla_sorgqr_thin(const float* F, int64_t m, int64_t n, int64_t ldf, const float* tau, float* Q, int64_t ldq);
GLA_API int gla_dorgqr_thin(const double* F, int64_t m, int64_t n, int64_t ldf, const double* tau, double* Q, int64_t ldq);
GLA_API int gla_zorgqr_thin(const void* F, int64_t m, int64_t n, int64_t ldf, const void* tau, void* Q, int64_t ldq);
GLA_API int gla_sorgqr_thin_dev(const float* dF, int64_t m, int64_t n, int64_t ldf, const float* dtau, float* dQ, int64_t ldq, void* stream);
GLA_API int gla_dorgqr_thin_dev(const double* dF, int64_t m, int64_t n, int64_t ldf, const double* dtau, double* dQ, int64_t ldq, void* stream);
GLA_API int gla_zorgqr_thin_dev(const void* dF, int64_t m, int64_t n, int64_t ldf, const void* dtau, void* dQ, int64_t ldq, void* stream);

/* ---- right reflector application ---------------------------------------------------
 * replaces reflectorApply!(A, x, tau)   src/qr.jl:19-42   A <- A (I - tau v v^H), v = [1; x[2:]]
 * returns -6 when lenx != n (DimensionMismatch at src/qr.jl:21-27). tau passed by (host) pointer. */
GLA_API int gla_sreflector_apply_right(float* A, int64_t m, int64_t n, int64_t lda, const float* x, int64_t lenx, const float* tau);
GLA_API int gla_dreflector_apply_right(double* A, int64_t m, int64_t n, int64_t lda, const double* x, int64_t lenx, const double* tau);
GLA_API int gla_zreflector_apply_right(void* A, int64_t m, int64_t n, int64_t lda, const void* x, int64_t lenx, const void* tau);
GLA_API int gla_sreflector_apply_right_dev(float* dA, int64_t m, int64_t n, int64_t lda, const float* dx, int64_t lenx, const float* tau, void* stream);
GLA_API int gla_dreflector_apply_right_dev(double* dA, int64_t m, int64_t n, int64_t lda, const double* dx, int64_t lenx, const double* tau, void* stream);
GLA_API int gla_zreflector_apply_right_dev(void* dA, int64_t m, int64_t n, int64_t lda, const void* dx, int64_t lenx, const void* tau, void* stream);

/* ---- batched small QR --------------------------------------------------------------
 * `batch` independent qrBlocked! problems (src/qr.jl:113-146 per matrix); matrices are
 * contiguous, column-major, stride m*n elements; tau has stride min(m,n).
 * 32x32 real runs register-resident, two matrices per warp (the device stack must be
 * 16-byte aligned for this path; an unaligned stack silently takes the next one); shapes
 * with n <= 32, m <= 64 (any element type, e.g. ComplexF64 32x32) run one matrix per WARP,
 * eight per CTA, in shared memory; other shapes with m*n*sizeof(T) <= 96 KiB one per CTA. */
GLA_API int gla_sgeqr_batched(float* A, int64_t m, int64_t n, int64_t batch, float* tau);
GLA_API int gla_dgeqr_batched(double* A, int64_t m, int64_t n, int64_t batch, double* tau);
GLA_API int gla_zgeqr_batched(void* A, int64_t m, int64_t n, int64_t batch, void* tau);
GLA_API int gla_sgeqr_batched_dev(float* dA, int64_t m, int64_t n, int64_t batch, float* dtau, void* stream);
GLA_API int gla_dgeqr_batched_dev(double* dA, int64_t m, int64_t n, int64_t batch, double* dtau, void* stream);
GLA_API int gla_zgeqr_batched_dev(void* dA, int64_t m, int64_t n, int64_t batch, void* dtau, void* stream);

/* ---- tall-skinny QR (R factor only) -------------------------------------------------
 * the R that qrBlocked! (src/qr.jl:113-146) would leave in the upper triangle of a tall
 * m x n matrix (n <= 64), computed by a TSQR tree.  R (n x n, ldr >= n, upper, strict lower
 * zeroed) is normalised to the reference's sign convention row by row only in the sense
 * documented in DESIGN.md ("TSQR sign"): diag(R) <= 0 is NOT guaranteed to match the
 * sequential Householder signs; parity tests compare after row-phase normalisation.
 *   gla_dtsqr_local_dev : one row block -> its n x n R factor (no communication)
 *   gla_dtsqr_combine_dev: `count` stacked R factors (each n x n, ld n) -> one R
 *   gla_dtsqr           : host pointers, single GPU. */
GLA_API int gla_dtsqr_local_dev(const double* dA, int64_t m, int64_t n, int64_t lda, double* dR, int64_t ldr, void* stream);
GLA_API int gla_dtsqr_combine_dev(const double* dRstack, int64_t count, int64_t n, double* dR, int64_t ldr, void* stream);
GLA_API int gla_dtsqr(const double* A, int64_t m, int64_t n, int64_t lda, double* R, int64_t ldr);
/* multi-GPU exchange step (one process per GPU): NCCL is resolved with dlopen("libnccl.so.2").
 *   gla_nccl_unique_id   : rank 0 fills 128 bytes (ncclUniqueId) that the host framework broadcasts
 *   gla_nccl_comm_init   : every rank, after cudaSetDevice; *comm is an ncclComm_t
 *   gla_dtsqr_allreduce_dev: dRloc (n x n, ld n, this rank's local R) -> ncclAllGather into dstack
 *                          (nranks x n x n) -> every rank reduces the stack to the same dR (n x n, ldr). */
GLA_API int gla_nccl_unique_id(void* id128);
GLA_API int gla_nccl_comm_init(void** comm, int nranks, const void* id128, int rank);
GLA_API int gla_nccl_comm_destroy(void* comm);
GLA_API int gla_dtsqr_allreduce_dev(void* comm, int nranks, const double* dRloc, int64_t n, double* dstack, double* dR, int64_t ldr, void* stream);

/* ---- recursive Cholesky, lower ------------------------------------------------------
 * replaces cholRecursive!(A, Val{:L}, cutoff)   src/cholesky.jl:37-55
 * (trsm = rdiv!(A21, LowerTriangular(A11)') at :48; rank-k update = rankUpdate! at :51 ->
 * src/juliaBLAS.jl:89-112).  `cutoff` is accepted as a hint. */
GLA_API int gla_spotrf_recursive_L(float* A, int64_t n, int64_t lda, int64_t cutoff);
GLA_API int gla_dpotrf_recursive_L(double* A, int64_t n, int64_t lda, int64_t cutoff);
GLA_API int gla_zpotrf_recursive_L(void* A, int64_t n, int64_t lda, int64_t cutoff);
GLA_API int gla_spotrf_recursive_L_dev(float* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream);
GLA_API int gla_dpotrf_recursive_L_dev(double* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream);
GLA_API int gla_zpotrf_recursive_L_dev(void* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream);
/* replace cholUnblocked!(A, Val{:L})   src/cholesky.jl:3-15   and   cholBlocked!(A, Val{:L}, blocksize)
 * src/cholesky.jl:17-35.  The lower Cholesky factor is unique, so on the GPU they are the same computation as
 * cholRecursive! (the reference's three variants differ in loop order only); `blocksize` >= 1 is a hint. */
GLA_API int gla_spotrf_unblocked_L(float* A, int64_t n, int64_t lda);
GLA_API int gla_dpotrf_unblocked_L(double* A, int64_t n, int64_t lda);
GLA_API int gla_zpotrf_unblocked_L(void* A, int64_t n, int64_t lda);
GLA_API int gla_spotrf_blocked_L(float* A, int64_t n, int64_t lda, int64_t blocksize);
GLA_API int gla_dpotrf_blocked_L(double* A, int64_t n, int64_t lda, int64_t blocksize);
GLA_API int gla_zpotrf_blocked_L(void* A, int64_t n, int64_t lda, int64_t blocksize);

/* ---- LDL^H without pivoting -----------------------------------------------------------
 * replaces ldlt!(A::Hermitian, blocksize)   src/ldlt.jl:155-162  ->  _ldlt_lower_blocked! (:80-103) for uplo = 'L',
 * _ldlt_upper_blocked! (:122-146) for uplo = 'U'.  In place: D on the diagonal, the unit factor in the strict `uplo`
 * triangle, the other triangle untouched; `blocksize` >= 1 is a hint (the reference's default is 128 / sizeof(T)).
 * Float32 / Float64 / ComplexF64 (Hermitian: the imaginary part of the diagonal is ignored, D is real; Quaternion and Rational
 * stay on the reference path).  A zero pivot returns
 * GLA_ERR_SINGULAR with its index in gla_last_info() (the reference divides by it). */
GLA_API int gla_sldlt(float* A, int64_t n, int64_t lda, int uplo, int64_t blocksize);
GLA_API int gla_dldlt(double* A, int64_t n, int64_t lda, int uplo, int64_t blocksize);
GLA_API int gla_zldlt(void* A, int64_t n, int64_t lda, int uplo, int64_t blocksize);
GLA_API int gla_sldlt_dev(float* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream);
GLA_API int gla_dldlt_dev(double* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream);
GLA_API int gla_zldlt_dev(void* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream);

/* ---- two-sided Householder reductions (the step after QR in the reference's SVD / eigen pipelines) ----------
 * gla_?bidiagonalize   replaces bidiagonalize!(A)                      src/svd.jl:328-381
 *     in place: m >= n -> upper bidiagonal (d = real(diag(A)), e = real(diag(A, 1))), left reflectors below the
 *     diagonal (taul: n entries), right reflectors right of the superdiagonal (taur: n-1 entries, row i holds
 *     reflector!(conj(row))); m < n -> lower bidiagonal, taur: m entries, taul: m-1 entries.
 * gla_?hessenberg      replaces _hessenberg!(A) / hessenberg!(A)       src/eigenGeneral.jl:18-31
 *     in place: upper Hessenberg part + reflectors below the first subdiagonal, tau: n-1 entries.
 * gla_?symtri          replaces symtri!(Hermitian(A, uplo)) = symtriLower! / symtriUpper!
 *                                                                      src/eigenSelfAdjoint.jl:446-564
 *     in place: only the `uplo` ('L' / 'U') triangle is read and written; diagonal and first off-diagonal hold
 *     the real tridiagonal, reflectors beyond; tau: n-1 entries (real element types stop one step earlier: tau[n-2] = 0).
 * Return values: 0, -k (k-th argument illegal), >= 1000 CUDA error.  Each reduction is ONE persistent kernel launched
 * cooperatively (all SMs of the device); the `_dev` twins are asynchronous on `stream`. */
GLA_API int gla_sbidiagonalize(float* A, int64_t m, int64_t n, int64_t lda, float* taul, float* taur);
GLA_API int gla_sbidiagonalize_dev(float* dA, int64_t m, int64_t n, int64_t lda, float* dtaul, float* dtaur, void* stream);
GLA_API int gla_shessenberg(float* A, int64_t n, int64_t lda, float* tau);
GLA_API int gla_shessenberg_dev(float* dA, int64_t n, int64_t lda, float* dtau, void* stream);
GLA_API int gla_ssymtri(float* A, int64_t n, int64_t lda, int uplo, float* tau);
GLA_API int gla_ssymtri_dev(float* dA, int64_t n, int64_t lda, int uplo, float* dtau, void* stream);
GLA_API int gla_dbidiagonalize(double* A, int64_t m, int64_t n, int64_t lda, double* taul, double* taur);
GLA_API int gla_dbidiagonalize_dev(double* dA, int64_t m, int64_t n, int64_t lda, double* dtaul, double* dtaur, void* stream);
GLA_API int gla_dhessenberg(double* A, int64_t n, int64_t lda, double* tau);
GLA_API int gla_dhessenberg_dev(double* dA, int64_t n, int64_t lda, double* dtau, void* stream);
GLA_API int gla_dsymtri(double* A, int64_t n, int64_t lda, int uplo, double* tau);
GLA_API int gla_dsymtri_dev(double* dA, int64_t n, int64_t lda, int uplo, double* dtau, void* stream);
GLA_API int gla_zbidiagonalize(void* A, int64_t m, int64_t n, int64_t lda, void* taul, void* taur);
GLA_API int gla_zbidiagonalize_dev(void* dA, int64_t m, int64_t n, int64_t lda, void* dtaul, void* dtaur, void* stream);
GLA_API int gla_zhessenberg(void* A, int64_t n, int64_t lda, void* tau);
GLA_API int gla_zhessenberg_dev(void* dA, int64_t n, int64_t lda, void* dtau, void* stream);
GLA_API int gla_zsymtri(void* A, int64_t n, int64_t lda, int uplo, void* tau);
GLA_API int gla_zsymtri_dev(void* dA, int64_t n, int64_t lda, int uplo, void* dtau, void* stream);

/* ---- workspace query ------------------------------------------------------------------
 * the reference's FFI precedent asks LAPACK for its workspace before the call (src/lapack.jl:152-170,
 * :514-553); here the library owns its temporaries (stream-ordered pool), and this reports how many
 * device bytes the `_dev` call of `op` on an m x n problem (n x n for potrf, batch ignored) will take
 * from the pool, so that a host can size its own allocations around it.  <0: illegal argument. */
#define GLA_OP_GEQR_BLOCKED 1
#define GLA_OP_POTRF_L 2
#define GLA_OP_GEQR_BATCHED 3
#define GLA_OP_TSQR 4
#define GLA_OP_LDLT 5            /* n x n (m ignored) */
#define GLA_OP_BIDIAGONALIZE 6
#define GLA_OP_HESSENBERG 7      /* n x n */
#define GLA_OP_SYMTRI 8          /* n x n, upper bound over uplo */
GLA_API int64_t gla_workspace_query(int op, int elem_bytes, int64_t m, int64_t n);

/* ---- Hermitian rank-k update, lower ---------------------------------------------------
 * replaces rankUpdate!(Hermitian(C,:L), A, alpha)   src/juliaBLAS.jl:89-112
 * C (n x n, lower triangle only) += alpha * A * A^H,  A is n x k.  alpha is real. */
GLA_API int gla_ssyrk_lower(float* C, int64_t n, int64_t ldc, const float* A, int64_t k, int64_t lda, float alpha);
GLA_API int gla_dsyrk_lower(double* C, int64_t n, int64_t ldc, const double* A, int64_t k, int64_t lda, double alpha);
GLA_API int gla_zherk_lower(void* C, int64_t n, int64_t ldc, const void* A, int64_t k, int64_t lda, double alpha);

#ifdef __cplusplus
}
#endif
#endif /* GLA_CUDA_H */
