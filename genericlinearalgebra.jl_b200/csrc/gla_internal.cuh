// gla_internal.cuh -- device-level entry points behind the C ABI (all take device pointers).
#pragma once
#include "common.cuh"

namespace gla {

const char* last_error();
extern thread_local double g_last_ms;
extern thread_local i64 g_last_info;   // 1-based failing minor of the last potrf that returned GLA_ERR_NOT_POSDEF

// K4: batched small QR (batched_qr.cu)
template <class T>
int geqr_batched_dev(T* dA, i64 m, i64 n, i64 batch, T* dtau, cudaStream_t st);

// K1/K2/K3: blocked QR of one large matrix (qr_blocked.cu)
// Host destination of a factorisation driven through the host-pointer entry point: the columns of an outer block are final
// as soon as its panel chain is done, so they travel device -> host on `copy` WHILE the far updates of the later blocks
// run (PCIe is otherwise idle during the factorisation).  copied_cols = columns already on their way when the call returns.
// Upload side (optional): the matrix ARRIVES in column chunks on another stream while the first outer block is already
// being factorised; chunk c covers the columns [up_col[c], up_col[c + 1]) and is on the device when up_ev[c] has fired.
// The driver makes each of its streams wait for exactly the chunks a step reads, and cuts the first far update along the
// chunk boundaries.
template <class T>
struct QrHostSink {
  T* hA = nullptr;
  i64 ldh = 0;
  cudaStream_t copy = nullptr;
  i64 copied_cols = 0;
  int up_chunks = 0;
  const i64* up_col = nullptr;
  const cudaEvent_t* up_ev = nullptr;
};
// Host destination of a Cholesky factorisation driven through the host-pointer entry point: the rows of U = L^H of an outer
// block are final when its panel kernels are done, so the matching COLUMNS of L (rows r .. n) are mirrored into the user
// layout and sent home on `copy` while the chain goes on.  copied_cols = columns already on their way when the call returns.
template <class T>
struct CholHostSink {
  T* hA = nullptr;
  i64 ldh = 0;
  cudaStream_t copy = nullptr;
  i64 copied_cols = 0;
};
template <class T>
int geqr_blocked_dev(T* dA, i64 m, i64 n, i64 lda, T* dtau, i64 blocksize_hint, cudaStream_t st, QrHostSink<T>* sink = nullptr);
i64 geqr_blocked_workspace_bytes(i64 m, i64 n, i64 elem_bytes);
// compact-WY T of all k=min(m,n) reflectors; dT is k x k (ldt)
template <class T>
int larft_dev(const T* dF, i64 m, i64 n, i64 ldf, const T* dtau, T* dT, i64 ldt, cudaStream_t st);
// A <- Q A (adjoint=0) or Q^H A (adjoint=1)
template <class T>
int ormqr_blocked_dev(const T* dF, i64 mF, i64 nF, i64 ldf, const T* dtau, T* dA, i64 mA, i64 nA,
                      i64 lda, int adjoint, cudaStream_t st);
// thin Q = H_1 .. H_k [I_k; 0], m x k
template <class T>
int orgqr_thin_dev(const T* dF, i64 m, i64 n, i64 ldf, const T* dtau, T* dQ, i64 ldq, cudaStream_t st);
template <class T>
int reflector_apply_right_dev(T* dA, i64 m, i64 n, i64 lda, const T* dx, T tau, cudaStream_t st);

// K5: TSQR (tsqr.cu)
int tsqr_local_dev(const double* dA, i64 m, i64 n, i64 lda, double* dR, i64 ldr, cudaStream_t st);
int tsqr_combine_dev(const double* dRs, i64 count, i64 n, double* dR, i64 ldr, cudaStream_t st);
// NCCL (libnccl.so.2, dlopen'ed): communicator helpers and the R-factor exchange + reduction
int nccl_unique_id(void* id128);
int nccl_comm_init(void** comm, int nranks, const void* id128, int rank);
int nccl_comm_destroy(void* comm);
int tsqr_allreduce_dev(void* comm, int nranks, const double* dRloc, i64 n, double* dstack, double* dR, i64 ldr,
                       cudaStream_t st);

// K6: Cholesky (chol.cu)
template <class T>
int potrf_recursive_L_dev(T* dA, i64 n, i64 lda, i64 cutoff, int* dinfo, cudaStream_t st, CholHostSink<T>* sink = nullptr);
// ldlt!(Hermitian(A, uplo)) without pivoting (src/ldlt.jl:80-162); real element types
template <class T>
int ldlt_dev(T* dA, i64 n, i64 lda, int upper, int* dinfo, cudaStream_t st);
template <class T>
int herk_lower_dev(T* dC, i64 n, i64 ldc, const T* dA, i64 k, i64 lda, typename Sc<T>::real alpha,
                   cudaStream_t st);

// ---- two-sided reductions (twosided.cu), asynchronous on `st`
// bidiagonalize!(A) (src/svd.jl:328-381): taul / taur get n / n-1 entries for m >= n, m-1 / m for m < n
template <class T>
int bidiagonalize_dev(T* dA, i64 m, i64 n, i64 lda, T* dtaul, T* dtaur, cudaStream_t st);
// _hessenberg!(A) (src/eigenGeneral.jl:18-31): tau has n-1 entries
template <class T>
int hessenberg_dev(T* dA, i64 n, i64 lda, T* dtau, cudaStream_t st);
// symtriLower! / symtriUpper! (src/eigenSelfAdjoint.jl:450-564): tau has n-1 entries
template <class T>
int symtri_dev(T* dA, i64 n, i64 lda, int upper, T* dtau, cudaStream_t st);

}  // namespace gla
