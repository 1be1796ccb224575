// capi.cu -- the C ABI declared in include/gla_cuda.h.  Host-pointer entry points stage through
// device memory (H2D, kernels, D2H) and are synchronous; `_dev` twins are asynchronous on the
// caller's stream.  No C++ exception crosses this boundary and nothing here computes on the CPU.
#include "../../include/gla_cuda.h"
#include "gla_internal.cuh"
#include <vector>

#include <new>
#include <stdlib.h>

using namespace gla;

namespace {

struct DevBuf {   // device buffer from the library's stream-ordered pool, tied to the stream of the call
  void* p = nullptr;
  cudaStream_t s = nullptr;
  ~DevBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  int alloc(size_t bytes, cudaStream_t st) {
    s = st;
    return pool_malloc(&p, bytes, st);
  }
  template <class T>
  T* as() {
    return static_cast<T*>(p);
  }
};

struct Stream {
  cudaStream_t s = nullptr;
  ~Stream() {
    if (s) cudaStreamDestroy(s);
  }
  int create() { return check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), __FILE__, __LINE__); }
};

struct Event {
  cudaEvent_t e = nullptr;
  ~Event() {
    if (e) cudaEventDestroy(e);
  }
  int create() { return check_cuda(cudaEventCreate(&e), __FILE__, __LINE__); }
};

// ---- 2-D staged copies of a column-major host matrix
template <class T>
int h2d_matrix(T* d, i64 ldd, const T* h, i64 ldh, i64 m, i64 n, cudaStream_t st) {
  if (m == 0 || n == 0) return 0;
  GLA_CUDA(cudaMemcpy2DAsync(d, ldd * sizeof(T), h, ldh * sizeof(T), m * sizeof(T), n, cudaMemcpyHostToDevice, st));
  return 0;
}
template <class T>
int d2h_matrix(T* h, i64 ldh, const T* d, i64 ldd, i64 m, i64 n, cudaStream_t st) {
  if (m == 0 || n == 0) return 0;
  GLA_CUDA(cudaMemcpy2DAsync(h, ldh * sizeof(T), d, ldd * sizeof(T), m * sizeof(T), n, cudaMemcpyDeviceToHost, st));
  return 0;
}

// ------------------------------------------------------------------ batched QR, host pointers
// Chunked multi-stream pipeline: H2D of later chunks and D2H of earlier ones overlap the kernel of chunk i (PCIe is
// full duplex), so the end-to-end time tends to max(H2D, D2H) rather than their sum.  Measured on the B200 box
// (tools/time_e2e.py; raw pinned copies run at 53 / 52 GB/s per direction): 64 MiB x 3 streams 4.6 M matrices/s,
// 128 MiB x 6 streams 5.4 M matrices/s (90 GB/s both directions together) -> the default.
template <class T>
int geqr_batched_host(T* A, i64 m, i64 n, i64 batch, T* tau) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (batch < 0) return -4;
  if (m == 0 || n == 0 || batch == 0) return 0;
  if (!A) return -1;
  if (!tau) return -5;
  const i64 k = m < n ? m : n;
  const i64 mat_bytes = m * n * (i64)sizeof(T);
  if (mat_bytes > 96 * 1024) return -2;
  static const i64 chunk_mb = [] { const char* e = getenv("GLA_BATCH_CHUNK_MB"); return e ? (i64)atoi(e) : 128ll; }();
  static const int want_ns = [] { const char* e = getenv("GLA_BATCH_STREAMS"); return e ? atoi(e) : 6; }();
  i64 chunk = ((chunk_mb > 0 ? chunk_mb : 64) << 20) / mat_bytes;
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  constexpr int NS = 8;
  Stream st[NS];
  // staging buffers from the stream-ordered pool (kept warm between calls, see sm_count()): no cudaMalloc / cudaFree
  // device synchronisation per call
  struct PoolBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~PoolBuf() {
      if (p) cudaFreeAsync(p, s);
    }
  } dA[NS], dtau[NS];
  const int ns_cap = want_ns < 1 ? 1 : (want_ns > NS ? NS : want_ns);
  int ns = (int)((batch + chunk - 1) / chunk < ns_cap ? (batch + chunk - 1) / chunk : ns_cap);
  for (int s = 0; s < ns; ++s) {
    GLA_TRY(st[s].create());
    dA[s].s = dtau[s].s = st[s].s;
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&dA[s].p), chunk * mat_bytes, st[s].s));
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&dtau[s].p), chunk * k * sizeof(T), st[s].s));
  }
  i64 done = 0;
  int it = 0;
  while (done < batch) {
    const int s = it % ns;
    const i64 nb = batch - done < chunk ? batch - done : chunk;
    T* hA = A + done * m * n;
    T* ht = tau + done * k;
    GLA_CUDA(cudaMemcpyAsync(dA[s].p, hA, nb * mat_bytes, cudaMemcpyHostToDevice, st[s].s));
    GLA_TRY(geqr_batched_dev<T>(static_cast<T*>(dA[s].p), m, n, nb, static_cast<T*>(dtau[s].p), st[s].s));
    GLA_CUDA(cudaMemcpyAsync(hA, dA[s].p, nb * mat_bytes, cudaMemcpyDeviceToHost, st[s].s));
    GLA_CUDA(cudaMemcpyAsync(ht, dtau[s].p, nb * k * sizeof(T), cudaMemcpyDeviceToHost, st[s].s));
    done += nb;
    ++it;
  }
  for (int s = 0; s < ns; ++s) GLA_CUDA(cudaStreamSynchronize(st[s].s));
  return 0;
}

// ------------------------------------------------------------------ device copy of a host matrix
// ld is padded to an even number of elements (16-byte column alignment for TMA / vector access)
// Lower-trapezoid transfers for the Hermitian entry points (only the lower triangle is read and written by the reference,
// src/cholesky.jl:37-55, src/ldlt.jl:80-103): block columns of TRI_BC columns, rows from the block's first row down --
// 1/2 + TRI_BC/(2n) of the square instead of all of it, in both directions.
constexpr i64 TRI_BC = 256;
template <class T>
int h2d_lower(T* d, i64 ldd, const T* h, i64 ldh, i64 n, cudaStream_t st) {
  for (i64 j = 0; j < n; j += TRI_BC) {
    const i64 nc = n - j < TRI_BC ? n - j : TRI_BC;
    GLA_TRY(h2d_matrix<T>(d + j + j * ldd, ldd, h + j + j * ldh, ldh, n - j, nc, st));
  }
  return 0;
}
template <class T>
int d2h_lower(T* h, i64 ldh, const T* d, i64 ldd, i64 n, cudaStream_t st) {
  for (i64 j = 0; j < n; j += TRI_BC) {
    const i64 nc = n - j < TRI_BC ? n - j : TRI_BC;
    GLA_TRY(d2h_matrix<T>(h + j + j * ldh, ldh, d + j + j * ldd, ldd, n - j, nc, st));
  }
  return 0;
}

template <class T>
int h2d_upper(T* d, i64 ldd, const T* h, i64 ldh, i64 n, cudaStream_t st) {
  for (i64 j = 0; j < n; j += TRI_BC) {
    const i64 nc = n - j < TRI_BC ? n - j : TRI_BC;
    GLA_TRY(h2d_matrix<T>(d + j * ldd, ldd, h + j * ldh, ldh, j + nc, nc, st));
  }
  return 0;
}
template <class T>
int d2h_upper(T* h, i64 ldh, const T* d, i64 ldd, i64 n, cudaStream_t st) {
  for (i64 j = 0; j < n; j += TRI_BC) {
    const i64 nc = n - j < TRI_BC ? n - j : TRI_BC;
    GLA_TRY(d2h_matrix<T>(h + j * ldh, ldh, d + j * ldd, ldd, j + nc, nc, st));
  }
  return 0;
}

template <class T>
struct DevMatrix {
  DevBuf buf;
  i64 ld = 0;
  T* p() { return buf.as<T>(); }
  int upload(const T* h, i64 ldh, i64 m, i64 n, cudaStream_t st) {
    ld = round_up(m > 0 ? m : 1, 16 / sizeof(T) > 2 ? 16 / sizeof(T) : 2);
    GLA_TRY(buf.alloc((size_t)ld * (n > 0 ? n : 1) * sizeof(T), st));
    return h2d_matrix<T>(p(), ld, h, ldh, m, n, st);
  }
  int download(T* h, i64 ldh, i64 m, i64 n, cudaStream_t st) { return d2h_matrix<T>(h, ldh, p(), ld, m, n, st); }
};

template <class T>
int geqr_blocked_host(T* A, i64 m, i64 n, i64 lda, T* tau, i64 hint) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (m == 0 || n == 0) return 0;
  if (!A) return -1;
  if (!tau) return -5;
  const i64 k = m < n ? m : n;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dtau;
  // Large problems: the matrix arrives in column chunks on a stream of its own (first two chunks = the first two outer
  // blocks, then 1024 columns each) while the factorisation of the first outer block is already running; the driver waits
  // per chunk (QrHostSink::up_*).  GLA_QR_NO_UPLOAD_OVERLAP=1: one upload ahead of everything.
  static const bool no_up = getenv("GLA_QR_NO_UPLOAD_OVERLAP") != nullptr;
  const bool stream_up = !no_up && n > 4 * 384 && m > 2 * 384;
  Stream up;
  std::vector<i64> up_col;
  std::vector<Event> up_events;
  std::vector<cudaEvent_t> up_ev;
  if (!stream_up) {
    GLA_TRY(dA.upload(A, lda, m, n, st.s));
  } else {
    dA.ld = round_up(m, 16 / sizeof(T) > 2 ? 16 / sizeof(T) : 2);
    GLA_TRY(dA.buf.alloc((size_t)dA.ld * n * sizeof(T), st.s));
    GLA_TRY(up.create());
    Event ready;
    GLA_TRY(ready.create());
    GLA_CUDA(cudaEventRecord(ready.e, st.s));            // the (stream-ordered) allocation
    GLA_CUDA(cudaStreamWaitEvent(up.s, ready.e, 0));
    up_col.push_back(0);
    for (i64 c = 0; c < n;) {
      const i64 w = up_col.size() <= 2 ? 384 : 1024;
      c = c + w < n ? c + w : n;
      up_col.push_back(c);
    }
    const int nchunks = (int)up_col.size() - 1;
    up_events.resize(nchunks);
    up_ev.resize(nchunks);
    for (int c = 0; c < nchunks; ++c) {
      GLA_TRY(up_events[c].create());
      GLA_TRY(h2d_matrix<T>(dA.p() + up_col[c] * dA.ld, dA.ld, A + up_col[c] * lda, lda, m, up_col[c + 1] - up_col[c], up.s));
      GLA_CUDA(cudaEventRecord(up_events[c].e, up.s));
      up_ev[c] = up_events[c].e;
    }
  }
  GLA_TRY(dtau.alloc(k * sizeof(T), st.s));
  GLA_CUDA(cudaMemsetAsync(dtau.p, 0, k * sizeof(T), st.s));
  Event e0, e1;
  GLA_TRY(e0.create());
  GLA_TRY(e1.create());
  GLA_CUDA(cudaEventRecord(e0.e, st.s));
  // finished outer blocks go home on a second stream while the later far updates run (see QrHostSink)
  Stream cp;
  GLA_TRY(cp.create());
  QrHostSink<T> sink;
  sink.hA = A;
  sink.ldh = lda;
  sink.copy = cp.s;
  if (stream_up) {
    sink.up_chunks = (int)up_ev.size();
    sink.up_col = up_col.data();
    sink.up_ev = up_ev.data();
  }
  if (int rc = geqr_blocked_dev<T>(dA.p(), m, n, dA.ld, dtau.as<T>(), hint, st.s, &sink)) {
    if (stream_up) cudaStreamSynchronize(up.s);   // no chunk may still be in flight when the buffer is released
    cudaStreamSynchronize(cp.s);
    return rc;
  }
  GLA_CUDA(cudaEventRecord(e1.e, st.s));
  if (sink.copied_cols < n)
    GLA_TRY(d2h_matrix<T>(A + sink.copied_cols * lda, lda, dA.p() + sink.copied_cols * dA.ld, dA.ld, m, n - sink.copied_cols, st.s));
  GLA_CUDA(cudaStreamSynchronize(cp.s));
  if (stream_up) GLA_CUDA(cudaStreamSynchronize(up.s));
  GLA_CUDA(cudaMemcpyAsync(tau, dtau.p, k * sizeof(T), cudaMemcpyDeviceToHost, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  float ms = 0;
  GLA_CUDA(cudaEventElapsedTime(&ms, e0.e, e1.e));
  g_last_ms = ms;
  return 0;
}

template <class T>
int larft_host(const T* F, i64 m, i64 n, i64 ldf, const T* tau, T* Tm, i64 ldt) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (ldf < (m > 1 ? m : 1)) return -4;
  const i64 k = m < n ? m : n;
  if (k == 0) return 0;
  if (!F) return -1;
  if (!tau) return -5;
  if (!Tm) return -6;
  if (ldt < k) return -7;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dF, dT;
  DevBuf dtau;
  GLA_TRY(dF.upload(F, ldf, m, n, st.s));
  GLA_TRY(dtau.alloc(k * sizeof(T), st.s));
  GLA_CUDA(cudaMemcpyAsync(dtau.p, tau, k * sizeof(T), cudaMemcpyHostToDevice, st.s));
  dT.ld = round_up(k, 2);
  GLA_TRY(dT.buf.alloc((size_t)dT.ld * k * sizeof(T), st.s));
  GLA_TRY(larft_dev<T>(dF.p(), m, n, dF.ld, dtau.as<T>(), dT.p(), dT.ld, st.s));
  GLA_TRY(dT.download(Tm, ldt, k, k, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

template <class T>
int ormqr_host(const T* F, i64 mF, i64 nF, i64 ldf, const T* tau, T* A, i64 mA, i64 nA, i64 lda, int adjoint) {
  if (mF < 0) return -2;
  if (nF < 0) return -3;
  if (ldf < (mF > 1 ? mF : 1)) return -4;
  if (mA != mF) return -7;
  if (nA < 0) return -8;
  if (lda < (mA > 1 ? mA : 1)) return -9;
  const i64 k = mF < nF ? mF : nF;
  if (k == 0 || nA == 0 || mA == 0) return 0;
  if (!F) return -1;
  if (!tau) return -5;
  if (!A) return -6;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dF, dA;
  DevBuf dtau;
  GLA_TRY(dF.upload(F, ldf, mF, nF, st.s));
  GLA_TRY(dA.upload(A, lda, mA, nA, st.s));
  GLA_TRY(dtau.alloc(k * sizeof(T), st.s));
  GLA_CUDA(cudaMemcpyAsync(dtau.p, tau, k * sizeof(T), cudaMemcpyHostToDevice, st.s));
  GLA_TRY(ormqr_blocked_dev<T>(dF.p(), mF, nF, dF.ld, dtau.as<T>(), dA.p(), mA, nA, dA.ld, adjoint, st.s));
  GLA_TRY(dA.download(A, lda, mA, nA, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

template <class T>
int orgqr_thin_host(const T* F, i64 m, i64 n, i64 ldf, const T* tau, T* Q, i64 ldq) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (ldf < (m > 1 ? m : 1)) return -4;
  if (ldq < (m > 1 ? m : 1)) return -7;
  const i64 k = m < n ? m : n;
  if (k == 0) return 0;
  if (!F) return -1;
  if (!tau) return -5;
  if (!Q) return -6;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dF, dQ;
  DevBuf dtau;
  GLA_TRY(dF.upload(F, ldf, m, n, st.s));
  GLA_TRY(dtau.alloc(k * sizeof(T), st.s));
  GLA_CUDA(cudaMemcpyAsync(dtau.p, tau, k * sizeof(T), cudaMemcpyHostToDevice, st.s));
  dQ.ld = round_up(m, 16 / sizeof(T) > 2 ? 16 / sizeof(T) : 2);
  GLA_TRY(dQ.buf.alloc((size_t)dQ.ld * k * sizeof(T), st.s));
  GLA_TRY(orgqr_thin_dev<T>(dF.p(), m, n, dF.ld, dtau.as<T>(), dQ.p(), dQ.ld, st.s));
  GLA_TRY(dQ.download(Q, ldq, m, k, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

template <class T>
int reflector_apply_right_host(T* A, i64 m, i64 n, i64 lda, const T* x, i64 lenx, const T* tau) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (lenx != n) return -6;  // DimensionMismatch, src/qr.jl:21-27 (lenx is the 6th argument)
  if (m == 0 || n == 0) return 0;
  if (!A) return -1;
  if (!x) return -5;
  if (!tau) return -7;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dx;
  GLA_TRY(dA.upload(A, lda, m, n, st.s));
  GLA_TRY(dx.alloc(n * sizeof(T), st.s));
  GLA_CUDA(cudaMemcpyAsync(dx.p, x, n * sizeof(T), cudaMemcpyHostToDevice, st.s));
  GLA_TRY(reflector_apply_right_dev<T>(dA.p(), m, n, dA.ld, dx.as<T>(), *tau, st.s));
  GLA_TRY(dA.download(A, lda, m, n, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

template <class T>
int potrf_host(T* A, i64 n, i64 lda, i64 cutoff) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  if (!A) return -1;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dinfo;
  dA.ld = round_up(n, 16 / sizeof(T) > 2 ? 16 / sizeof(T) : 2);
  GLA_TRY(dA.buf.alloc((size_t)dA.ld * n * sizeof(T), st.s));
  GLA_TRY(h2d_lower<T>(dA.p(), dA.ld, A, lda, n, st.s));   // the strict upper triangle is neither read nor written on the device
  GLA_TRY(dinfo.alloc(sizeof(int), st.s));
  Event e0, e1;
  GLA_TRY(e0.create());
  GLA_TRY(e1.create());
  GLA_CUDA(cudaEventRecord(e0.e, st.s));
  // the columns of L that belong to a finished outer block go home on a second stream while the chain goes on (CholHostSink)
  Stream cp;
  GLA_TRY(cp.create());
  CholHostSink<T> sink;
  sink.hA = A;
  sink.ldh = lda;
  sink.copy = cp.s;
  if (int rc = potrf_recursive_L_dev<T>(dA.p(), n, dA.ld, cutoff, dinfo.as<int>(), st.s, &sink)) {
    cudaStreamSynchronize(cp.s);
    return rc;
  }
  GLA_CUDA(cudaEventRecord(e1.e, st.s));
  int info = 0;
  GLA_CUDA(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, st.s));
  // like the reference (DomainError out of sqrt at src/cholesky.jl:40 with A partially overwritten), a failed call
  // still returns the partially factorised lower triangle; the index of the minor goes out of band (gla_last_info)
  for (i64 j = sink.copied_cols; j < n; j += TRI_BC) {   // the columns that did not travel yet (lower trapezoid only)
    const i64 nc = n - j < TRI_BC ? n - j : TRI_BC;
    GLA_TRY(d2h_matrix<T>(A + j + j * lda, lda, dA.p() + j + j * dA.ld, dA.ld, n - j, nc, st.s));
  }
  GLA_CUDA(cudaStreamSynchronize(st.s));
  GLA_CUDA(cudaStreamSynchronize(cp.s));
  float ms = 0;
  GLA_CUDA(cudaEventElapsedTime(&ms, e0.e, e1.e));
  g_last_ms = ms;
  if (info != 0) {
    g_last_info = info;
    return GLA_ERR_NOT_POSDEF;
  }
  return 0;
}

template <class T>
int ldlt_host(T* A, i64 n, i64 lda, int uplo, i64 blocksize) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (uplo != 'L' && uplo != 'U') return -4;
  if (blocksize < 1) return -5;
  if (n == 0) return 0;
  if (!A) return -1;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dinfo;
  dA.ld = round_up(n, 16 / sizeof(T) > 2 ? 16 / sizeof(T) : 2);
  GLA_TRY(dA.buf.alloc((size_t)dA.ld * n * sizeof(T), st.s));
  if (uplo == 'U') GLA_TRY(h2d_upper<T>(dA.p(), dA.ld, A, lda, n, st.s));   // only the `uplo` triangle is referenced
  else GLA_TRY(h2d_lower<T>(dA.p(), dA.ld, A, lda, n, st.s));
  GLA_TRY(dinfo.alloc(sizeof(int), st.s));
  GLA_TRY(ldlt_dev<T>(dA.p(), n, dA.ld, uplo == 'U', dinfo.as<int>(), st.s));
  int info = 0;
  GLA_CUDA(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, st.s));
  if (uplo == 'U') GLA_TRY(d2h_upper<T>(A, lda, dA.p(), dA.ld, n, st.s));
  else GLA_TRY(d2h_lower<T>(A, lda, dA.p(), dA.ld, n, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  if (info != 0) {
    g_last_info = info;
    return GLA_ERR_SINGULAR;
  }
  return 0;
}

// ------------------------------------------------------------------ two-sided reductions, host pointers
template <class T>
int bidiagonalize_host(T* A, i64 m, i64 n, i64 lda, T* taul, T* taur) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (m == 0 || n == 0) return 0;
  if (!A) return -1;
  const i64 nl = m >= n ? n : m - 1, nr = m >= n ? n - 1 : m;
  if (nl > 0 && !taul) return -5;
  if (nr > 0 && !taur) return -6;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dt;
  GLA_TRY(dA.upload(A, lda, m, n, st.s));
  GLA_TRY(dt.alloc((size_t)(nl + nr + 2) * sizeof(T), st.s));
  GLA_CUDA(cudaMemsetAsync(dt.p, 0, (size_t)(nl + nr + 2) * sizeof(T), st.s));
  T* dl = dt.as<T>();
  T* dr = dl + nl + 1;
  Event e0, e1;
  GLA_TRY(e0.create());
  GLA_TRY(e1.create());
  GLA_CUDA(cudaEventRecord(e0.e, st.s));
  GLA_TRY(bidiagonalize_dev<T>(dA.p(), m, n, dA.ld, dl, dr, st.s));
  GLA_CUDA(cudaEventRecord(e1.e, st.s));
  GLA_TRY(dA.download(A, lda, m, n, st.s));
  if (nl > 0) GLA_CUDA(cudaMemcpyAsync(taul, dl, nl * sizeof(T), cudaMemcpyDeviceToHost, st.s));
  if (nr > 0) GLA_CUDA(cudaMemcpyAsync(taur, dr, nr * sizeof(T), cudaMemcpyDeviceToHost, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  float ms = 0;
  GLA_CUDA(cudaEventElapsedTime(&ms, e0.e, e1.e));
  g_last_ms = ms;
  return 0;
}

template <class T>
int square_reduction_host(T* A, i64 n, i64 lda, T* tau, int which, int upper) {   // which: 0 hessenberg, 1 symtri
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  if (!A) return -1;
  if (n > 1 && !tau) return which == 0 ? -4 : -5;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dA;
  DevBuf dt;
  GLA_TRY(dA.upload(A, lda, n, n, st.s));
  GLA_TRY(dt.alloc((size_t)n * sizeof(T), st.s));
  GLA_CUDA(cudaMemsetAsync(dt.p, 0, (size_t)n * sizeof(T), st.s));
  Event e0, e1;
  GLA_TRY(e0.create());
  GLA_TRY(e1.create());
  GLA_CUDA(cudaEventRecord(e0.e, st.s));
  if (which == 0) GLA_TRY(hessenberg_dev<T>(dA.p(), n, dA.ld, dt.as<T>(), st.s));
  else GLA_TRY(symtri_dev<T>(dA.p(), n, dA.ld, upper, dt.as<T>(), st.s));
  GLA_CUDA(cudaEventRecord(e1.e, st.s));
  GLA_TRY(dA.download(A, lda, n, n, st.s));
  if (n > 1) GLA_CUDA(cudaMemcpyAsync(tau, dt.p, (n - 1) * sizeof(T), cudaMemcpyDeviceToHost, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  float ms = 0;
  GLA_CUDA(cudaEventElapsedTime(&ms, e0.e, e1.e));
  g_last_ms = ms;
  return 0;
}

template <class T>
int herk_host(T* Cm, i64 n, i64 ldc, const T* A, i64 k, i64 lda, typename Sc<T>::real alpha) {
  if (n < 0) return -2;
  if (ldc < (n > 1 ? n : 1)) return -3;
  if (k < 0) return -5;
  if (lda < (n > 1 ? n : 1)) return -6;
  if (n == 0 || k == 0) return 0;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<T> dC, dA;
  GLA_TRY(dC.upload(Cm, ldc, n, n, st.s));
  GLA_TRY(dA.upload(A, lda, n, k, st.s));
  GLA_TRY(herk_lower_dev<T>(dC.p(), n, dC.ld, dA.p(), k, dA.ld, alpha, st.s));
  GLA_TRY(dC.download(Cm, ldc, n, n, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

int tsqr_host(const double* A, i64 m, i64 n, i64 lda, double* R, i64 ldr) {
  if (m < 0) return -2;
  if (n < 0 || n > 64) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (ldr < (n > 1 ? n : 1)) return -6;
  if (n == 0) return 0;
  Stream st;
  GLA_TRY(st.create());
  DevMatrix<double> dR;
  dR.ld = n;
  GLA_TRY(dR.buf.alloc((size_t)n * n * sizeof(double), st.s));
  // TSQR streams: row chunks of 2^20 rows go through a three-deep ring (H2D of chunk i+1 / i+2 under the reduction of chunk
  // i, which is ~6x shorter than its upload), every chunk leaves an n x n R in a stack, the stack is reduced at the end.  End
  // to end = the upload; 1.5 GB of device memory instead of the whole matrix.
  constexpr i64 CH = 1 << 20;
  if (m <= CH + CH / 2) {
    DevMatrix<double> dA;
    GLA_TRY(dA.upload(A, lda, m, n, st.s));
    GLA_TRY(tsqr_local_dev(dA.p(), m, n, dA.ld, dR.p(), n, st.s));
  } else {
    constexpr int NS = 3;
    const i64 nch = (m + CH - 1) / CH;
    Stream ss[NS];
    DevBuf buf[NS], stack;
    Event done_ev[NS];
    GLA_TRY(stack.alloc((size_t)nch * n * n * sizeof(double), st.s));
    GLA_CUDA(cudaStreamSynchronize(st.s));   // the stack exists before the chunk streams write into it
    for (int i = 0; i < NS; ++i) {
      GLA_TRY(ss[i].create());
      GLA_TRY(buf[i].alloc((size_t)CH * n * sizeof(double), ss[i].s));
      GLA_TRY(done_ev[i].create());
    }
    for (i64 c = 0; c < nch; ++c) {
      const int i = (int)(c % NS);
      const i64 r0 = c * CH, rows = (m - r0 < CH) ? m - r0 : CH;
      GLA_TRY(h2d_matrix<double>(buf[i].as<double>(), CH, A + r0, lda, rows, n, ss[i].s));
      GLA_TRY(tsqr_local_dev(buf[i].as<double>(), rows, n, CH, stack.as<double>() + c * n * n, n, ss[i].s));
    }
    for (int i = 0; i < NS; ++i) {
      GLA_CUDA(cudaEventRecord(done_ev[i].e, ss[i].s));
      GLA_CUDA(cudaStreamWaitEvent(st.s, done_ev[i].e, 0));
    }
    GLA_TRY(tsqr_combine_dev(stack.as<double>(), nch, n, dR.p(), n, st.s));
    GLA_TRY(dR.download(R, ldr, n, n, st.s));
    GLA_CUDA(cudaStreamSynchronize(st.s));
    for (int i = 0; i < NS; ++i) GLA_CUDA(cudaStreamSynchronize(ss[i].s));
    return 0;
  }
  GLA_TRY(dR.download(R, ldr, n, n, st.s));
  GLA_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

}  // namespace

extern "C" {

int gla_version(void) { return 1; }

int gla_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    check_cuda(e, __FILE__, __LINE__);
    return -1;
  }
  return n;
}

const char* gla_last_error_string(void) { return last_error(); }

int gla_set_device(int device) {
  GLA_CUDA(cudaSetDevice(device));
  return 0;
}

double gla_last_device_ms(void) { return g_last_ms; }
int64_t gla_last_info(void) { return g_last_info; }

// ---- batched
int gla_sgeqr_batched(float* A, int64_t m, int64_t n, int64_t batch, float* tau) {
  return geqr_batched_host<float>(A, m, n, batch, tau);
}
int gla_dgeqr_batched(double* A, int64_t m, int64_t n, int64_t batch, double* tau) {
  return geqr_batched_host<double>(A, m, n, batch, tau);
}
int gla_zgeqr_batched(void* A, int64_t m, int64_t n, int64_t batch, void* tau) {
  return geqr_batched_host<zd>(static_cast<zd*>(A), m, n, batch, static_cast<zd*>(tau));
}
int gla_sgeqr_batched_dev(float* dA, int64_t m, int64_t n, int64_t batch, float* dtau, void* stream) {
  return geqr_batched_dev<float>(dA, m, n, batch, dtau, static_cast<cudaStream_t>(stream));
}
int gla_dgeqr_batched_dev(double* dA, int64_t m, int64_t n, int64_t batch, double* dtau, void* stream) {
  return geqr_batched_dev<double>(dA, m, n, batch, dtau, static_cast<cudaStream_t>(stream));
}
int gla_zgeqr_batched_dev(void* dA, int64_t m, int64_t n, int64_t batch, void* dtau, void* stream) {
  return geqr_batched_dev<zd>(static_cast<zd*>(dA), m, n, batch, static_cast<zd*>(dtau),
                              static_cast<cudaStream_t>(stream));
}

// ---- blocked QR
#define ZP(p) static_cast<zd*>(p)
#define ZCP(p) static_cast<const zd*>(p)
#define STREAM(s) static_cast<cudaStream_t>(s)
int gla_sgeqr_blocked(float* A, int64_t m, int64_t n, int64_t lda, float* tau, int64_t hint) { return geqr_blocked_host<float>(A, m, n, lda, tau, hint); }
int gla_dgeqr_blocked(double* A, int64_t m, int64_t n, int64_t lda, double* tau, int64_t hint) { return geqr_blocked_host<double>(A, m, n, lda, tau, hint); }
int gla_zgeqr_blocked(void* A, int64_t m, int64_t n, int64_t lda, void* tau, int64_t hint) { return geqr_blocked_host<zd>(ZP(A), m, n, lda, ZP(tau), hint); }
int gla_sgeqr_blocked_dev(float* dA, int64_t m, int64_t n, int64_t lda, float* dtau, int64_t hint, void* stream) { return geqr_blocked_dev<float>(dA, m, n, lda, dtau, hint, STREAM(stream)); }
int gla_dgeqr_blocked_dev(double* dA, int64_t m, int64_t n, int64_t lda, double* dtau, int64_t hint, void* stream) { return geqr_blocked_dev<double>(dA, m, n, lda, dtau, hint, STREAM(stream)); }
int gla_zgeqr_blocked_dev(void* dA, int64_t m, int64_t n, int64_t lda, void* dtau, int64_t hint, void* stream) { return geqr_blocked_dev<zd>(ZP(dA), m, n, lda, ZP(dtau), hint, STREAM(stream)); }

// ---- T factor
int gla_slarft(const float* F, int64_t m, int64_t n, int64_t ldf, const float* tau, float* T, int64_t ldt) { return larft_host<float>(F, m, n, ldf, tau, T, ldt); }
int gla_dlarft(const double* F, int64_t m, int64_t n, int64_t ldf, const double* tau, double* T, int64_t ldt) { return larft_host<double>(F, m, n, ldf, tau, T, ldt); }
int gla_zlarft(const void* F, int64_t m, int64_t n, int64_t ldf, const void* tau, void* T, int64_t ldt) { return larft_host<zd>(ZCP(F), m, n, ldf, ZCP(tau), ZP(T), ldt); }

// ---- block reflector application
int gla_sormqr_blocked(const float* F, int64_t mF, int64_t nF, int64_t ldf, const float* tau, float* A, int64_t mA, int64_t nA, int64_t lda, int adjoint) { return ormqr_host<float>(F, mF, nF, ldf, tau, A, mA, nA, lda, adjoint); }
int gla_dormqr_blocked(const double* F, int64_t mF, int64_t nF, int64_t ldf, const double* tau, double* A, int64_t mA, int64_t nA, int64_t lda, int adjoint) { return ormqr_host<double>(F, mF, nF, ldf, tau, A, mA, nA, lda, adjoint); }
int gla_zormqr_blocked(const void* F, int64_t mF, int64_t nF, int64_t ldf, const void* tau, void* A, int64_t mA, int64_t nA, int64_t lda, int adjoint) { return ormqr_host<zd>(ZCP(F), mF, nF, ldf, ZCP(tau), ZP(A), mA, nA, lda, adjoint); }

// ---- thin Q
int gla_sorgqr_thin(const float* F, int64_t m, int64_t n, int64_t ldf, const float* tau, float* Q, int64_t ldq) { return orgqr_thin_host<float>(F, m, n, ldf, tau, Q, ldq); }
int gla_dorgqr_thin(const double* F, int64_t m, int64_t n, int64_t ldf, const double* tau, double* Q, int64_t ldq) { return orgqr_thin_host<double>(F, m, n, ldf, tau, Q, ldq); }
int gla_zorgqr_thin(const void* F, int64_t m, int64_t n, int64_t ldf, const void* tau, void* Q, int64_t ldq) { return orgqr_thin_host<zd>(ZCP(F), m, n, ldf, ZCP(tau), ZP(Q), ldq); }
int gla_sorgqr_thin_dev(const float* dF, int64_t m, int64_t n, int64_t ldf, const float* dtau, float* dQ, int64_t ldq, void* stream) { return orgqr_thin_dev<float>(dF, m, n, ldf, dtau, dQ, ldq, STREAM(stream)); }
int gla_dorgqr_thin_dev(const double* dF, int64_t m, int64_t n, int64_t ldf, const double* dtau, double* dQ, int64_t ldq, void* stream) { return orgqr_thin_dev<double>(dF, m, n, ldf, dtau, dQ, ldq, STREAM(stream)); }
int gla_zorgqr_thin_dev(const void* dF, int64_t m, int64_t n, int64_t ldf, const void* dtau, void* dQ, int64_t ldq, void* stream) { return orgqr_thin_dev<zd>(ZCP(dF), m, n, ldf, ZCP(dtau), ZP(dQ), ldq, STREAM(stream)); }

// ---- right reflector application
int gla_sreflector_apply_right(float* A, int64_t m, int64_t n, int64_t lda, const float* x, int64_t lenx, const float* tau) { return reflector_apply_right_host<float>(A, m, n, lda, x, lenx, tau); }
int gla_dreflector_apply_right(double* A, int64_t m, int64_t n, int64_t lda, const double* x, int64_t lenx, const double* tau) { return reflector_apply_right_host<double>(A, m, n, lda, x, lenx, tau); }
int gla_zreflector_apply_right(void* A, int64_t m, int64_t n, int64_t lda, const void* x, int64_t lenx, const void* tau) { return reflector_apply_right_host<zd>(ZP(A), m, n, lda, ZCP(x), lenx, ZCP(tau)); }

// ---- device-pointer twins of the T build, the block application and the right reflector application
int gla_slarft_dev(const float* dF, int64_t m, int64_t n, int64_t ldf, const float* dtau, float* dT, int64_t ldt, void* stream) { return larft_dev<float>(dF, m, n, ldf, dtau, dT, ldt, STREAM(stream)); }
int gla_dlarft_dev(const double* dF, int64_t m, int64_t n, int64_t ldf, const double* dtau, double* dT, int64_t ldt, void* stream) { return larft_dev<double>(dF, m, n, ldf, dtau, dT, ldt, STREAM(stream)); }
int gla_zlarft_dev(const void* dF, int64_t m, int64_t n, int64_t ldf, const void* dtau, void* dT, int64_t ldt, void* stream) { return larft_dev<zd>(ZCP(dF), m, n, ldf, ZCP(dtau), ZP(dT), ldt, STREAM(stream)); }
int gla_sormqr_blocked_dev(const float* dF, int64_t mF, int64_t nF, int64_t ldf, const float* dtau, float* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream) {
  if (mA != mF) return -7;
  return ormqr_blocked_dev<float>(dF, mF, nF, ldf, dtau, dA, mA, nA, lda, adjoint, STREAM(stream));
}
int gla_dormqr_blocked_dev(const double* dF, int64_t mF, int64_t nF, int64_t ldf, const double* dtau, double* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream) {
  if (mA != mF) return -7;
  return ormqr_blocked_dev<double>(dF, mF, nF, ldf, dtau, dA, mA, nA, lda, adjoint, STREAM(stream));
}
int gla_zormqr_blocked_dev(const void* dF, int64_t mF, int64_t nF, int64_t ldf, const void* dtau, void* dA, int64_t mA, int64_t nA, int64_t lda, int adjoint, void* stream) {
  if (mA != mF) return -7;
  return ormqr_blocked_dev<zd>(ZCP(dF), mF, nF, ldf, ZCP(dtau), ZP(dA), mA, nA, lda, adjoint, STREAM(stream));
}
int gla_sreflector_apply_right_dev(float* dA, int64_t m, int64_t n, int64_t lda, const float* dx, int64_t lenx, const float* tau, void* stream) {
  if (lenx != n) return -6;
  if (!tau) return -7;
  return reflector_apply_right_dev<float>(dA, m, n, lda, dx, *tau, STREAM(stream));
}
int gla_dreflector_apply_right_dev(double* dA, int64_t m, int64_t n, int64_t lda, const double* dx, int64_t lenx, const double* tau, void* stream) {
  if (lenx != n) return -6;
  if (!tau) return -7;
  return reflector_apply_right_dev<double>(dA, m, n, lda, dx, *tau, STREAM(stream));
}
int gla_zreflector_apply_right_dev(void* dA, int64_t m, int64_t n, int64_t lda, const void* dx, int64_t lenx, const void* tau, void* stream) {
  if (lenx != n) return -6;
  if (!tau) return -7;
  return reflector_apply_right_dev<zd>(ZP(dA), m, n, lda, ZCP(dx), *ZCP(tau), STREAM(stream));
}

// ---- TSQR
int gla_dtsqr_local_dev(const double* dA, int64_t m, int64_t n, int64_t lda, double* dR, int64_t ldr, void* stream) { return tsqr_local_dev(dA, m, n, lda, dR, ldr, STREAM(stream)); }
int gla_dtsqr_combine_dev(const double* dRs, int64_t count, int64_t n, double* dR, int64_t ldr, void* stream) { return tsqr_combine_dev(dRs, count, n, dR, ldr, STREAM(stream)); }
int gla_dtsqr(const double* A, int64_t m, int64_t n, int64_t lda, double* R, int64_t ldr) { return tsqr_host(A, m, n, lda, R, ldr); }
int gla_nccl_unique_id(void* id128) { return id128 ? nccl_unique_id(id128) : -1; }
int gla_nccl_comm_init(void** comm, int nranks, const void* id128, int rank) {
  if (!comm) return -1;
  if (nranks < 1) return -2;
  if (!id128) return -3;
  if (rank < 0 || rank >= nranks) return -4;
  return nccl_comm_init(comm, nranks, id128, rank);
}
int gla_nccl_comm_destroy(void* comm) { return comm ? nccl_comm_destroy(comm) : -1; }
int gla_dtsqr_allreduce_dev(void* comm, int nranks, const double* dRloc, int64_t n, double* dstack, double* dR, int64_t ldr,
                            void* stream) {
  return tsqr_allreduce_dev(comm, nranks, dRloc, n, dstack, dR, ldr, STREAM(stream));
}

// ---- recursive Cholesky
int gla_spotrf_recursive_L(float* A, int64_t n, int64_t lda, int64_t cutoff) { return potrf_host<float>(A, n, lda, cutoff); }
int gla_dpotrf_recursive_L(double* A, int64_t n, int64_t lda, int64_t cutoff) { return potrf_host<double>(A, n, lda, cutoff); }
int gla_zpotrf_recursive_L(void* A, int64_t n, int64_t lda, int64_t cutoff) { return potrf_host<zd>(ZP(A), n, lda, cutoff); }
int gla_spotrf_recursive_L_dev(float* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream) { return potrf_recursive_L_dev<float>(dA, n, lda, cutoff, dinfo, STREAM(stream)); }
int gla_dpotrf_recursive_L_dev(double* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream) { return potrf_recursive_L_dev<double>(dA, n, lda, cutoff, dinfo, STREAM(stream)); }
int gla_zpotrf_recursive_L_dev(void* dA, int64_t n, int64_t lda, int64_t cutoff, int* dinfo, void* stream) { return potrf_recursive_L_dev<zd>(ZP(dA), n, lda, cutoff, dinfo, STREAM(stream)); }

// cholUnblocked!(A, Val{:L}) (src/cholesky.jl:3-15) and cholBlocked!(A, Val{:L}, blocksize) (src/cholesky.jl:17-35): the
// lower Cholesky factor is unique, so both map onto the same device routine (the blocksize is a CPU tuning parameter)
int gla_spotrf_unblocked_L(float* A, int64_t n, int64_t lda) { return potrf_host<float>(A, n, lda, 1); }
int gla_dpotrf_unblocked_L(double* A, int64_t n, int64_t lda) { return potrf_host<double>(A, n, lda, 1); }
int gla_zpotrf_unblocked_L(void* A, int64_t n, int64_t lda) { return potrf_host<zd>(ZP(A), n, lda, 1); }
int gla_spotrf_blocked_L(float* A, int64_t n, int64_t lda, int64_t blocksize) { return blocksize < 1 ? -4 : potrf_host<float>(A, n, lda, 1); }
int gla_dpotrf_blocked_L(double* A, int64_t n, int64_t lda, int64_t blocksize) { return blocksize < 1 ? -4 : potrf_host<double>(A, n, lda, 1); }
int gla_zpotrf_blocked_L(void* A, int64_t n, int64_t lda, int64_t blocksize) { return blocksize < 1 ? -4 : potrf_host<zd>(ZP(A), n, lda, 1); }

// ---- LDL^H without pivoting (real element types; ComplexF64 / Quaternion stay on the reference path)
int gla_sldlt(float* A, int64_t n, int64_t lda, int uplo, int64_t blocksize) { return ldlt_host<float>(A, n, lda, uplo, blocksize); }
int gla_dldlt(double* A, int64_t n, int64_t lda, int uplo, int64_t blocksize) { return ldlt_host<double>(A, n, lda, uplo, blocksize); }
int gla_zldlt(void* A, int64_t n, int64_t lda, int uplo, int64_t blocksize) { return ldlt_host<zd>(ZP(A), n, lda, uplo, blocksize); }
int gla_zldlt_dev(void* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream) {
  if (uplo != 'L' && uplo != 'U') return -4;
  return ldlt_dev<zd>(ZP(dA), n, lda, uplo == 'U', dinfo, STREAM(stream));
}
int gla_sldlt_dev(float* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream) {
  if (uplo != 'L' && uplo != 'U') return -4;
  return ldlt_dev<float>(dA, n, lda, uplo == 'U', dinfo, STREAM(stream));
}
int gla_dldlt_dev(double* dA, int64_t n, int64_t lda, int uplo, int* dinfo, void* stream) {
  if (uplo != 'L' && uplo != 'U') return -4;
  return ldlt_dev<double>(dA, n, lda, uplo == 'U', dinfo, STREAM(stream));
}

// ---- two-sided reductions
#define GLA_TWOSIDED(P, T, CT)                                                                                        \
  int gla_##P##bidiagonalize(CT* A, int64_t m, int64_t n, int64_t lda, CT* taul, CT* taur) {                          \
    return bidiagonalize_host<T>((T*)A, m, n, lda, (T*)taul, (T*)taur);                                               \
  }                                                                                                                   \
  int gla_##P##bidiagonalize_dev(CT* dA, int64_t m, int64_t n, int64_t lda, CT* dtaul, CT* dtaur, void* stream) {     \
    return bidiagonalize_dev<T>((T*)dA, m, n, lda, (T*)dtaul, (T*)dtaur, STREAM(stream));                             \
  }                                                                                                                   \
  int gla_##P##hessenberg(CT* A, int64_t n, int64_t lda, CT* tau) {                                                   \
    return square_reduction_host<T>((T*)A, n, lda, (T*)tau, 0, 0);                                                    \
  }                                                                                                                   \
  int gla_##P##hessenberg_dev(CT* dA, int64_t n, int64_t lda, CT* dtau, void* stream) {                               \
    return hessenberg_dev<T>((T*)dA, n, lda, (T*)dtau, STREAM(stream));                                               \
  }                                                                                                                   \
  int gla_##P##symtri(CT* A, int64_t n, int64_t lda, int uplo, CT* tau) {                                             \
    if (uplo != 'L' && uplo != 'U') return -4;                                                                        \
    return square_reduction_host<T>((T*)A, n, lda, (T*)tau, 1, uplo == 'U');                                          \
  }                                                                                                                   \
  int gla_##P##symtri_dev(CT* dA, int64_t n, int64_t lda, int uplo, CT* dtau, void* stream) {                         \
    if (uplo != 'L' && uplo != 'U') return -4;                                                                        \
    return symtri_dev<T>((T*)dA, n, lda, uplo == 'U', (T*)dtau, STREAM(stream));                                      \
  }
GLA_TWOSIDED(s, float, float)
GLA_TWOSIDED(d, double, double)
GLA_TWOSIDED(z, zd, void)

// ---- workspace query
int64_t gla_workspace_query(int op, int elem_bytes, int64_t m, int64_t n) {
  if (elem_bytes != 4 && elem_bytes != 8 && elem_bytes != 16) return -2;
  if (m < 0) return -3;
  if (n < 0) return -4;
  const int64_t e = elem_bytes;
  switch (op) {
    case GLA_OP_GEQR_BLOCKED: return geqr_blocked_workspace_bytes(m, n, e);
    case GLA_OP_POTRF_L: return (round_up(n, 16) * n + (n + 63) / 64 * 64 * 64) * e + 256;   // mirror + diagonal blocks
    case GLA_OP_GEQR_BATCHED: return 0;
    case GLA_OP_TSQR: return (2 * 2 + 8) * (int64_t)sm_count() * n * n * e;   // per-warp R factors of level 0 + two levels of per-CTA ones
    case GLA_OP_LDLT: return (2 * round_up(n, 16) * n + (n + 63) / 64 * 64 * 64) * e + 512;   // mirror + Y = D U + diagonal blocks
    case GLA_OP_BIDIAGONALIZE:
    case GLA_OP_HESSENBERG:
    case GLA_OP_SYMTRI: {
      // barrier counter; per-CTA vector slabs when max(m, n) elements (two vectors for symtri) exceed the 200 KB shared-memory
      // budget; symtri: the two partial vectors and (uplo = 'U') the flipped copy; bidiagonalize with m < n: the A^H copy
      const int64_t len = m > n ? m : n;
      const int64_t nvec = op == GLA_OP_SYMTRI ? 2 : 1;
      int64_t bytes = 256;
      if ((32 + 512) * e + 16 + nvec * len * e > 200 * 1024) bytes += (int64_t)sm_count() * nvec * len * e;
      if (op == GLA_OP_SYMTRI) bytes += 2 * len * e + round_up(n, 2) * n * e;
      if (op == GLA_OP_BIDIAGONALIZE && m < n) bytes += round_up(n, 2) * m * e;
      return bytes;
    }
    default: return -1;
  }
}

// ---- Hermitian rank-k update
int gla_ssyrk_lower(float* C, int64_t n, int64_t ldc, const float* A, int64_t k, int64_t lda, float alpha) { return herk_host<float>(C, n, ldc, A, k, lda, alpha); }
int gla_dsyrk_lower(double* C, int64_t n, int64_t ldc, const double* A, int64_t k, int64_t lda, double alpha) { return herk_host<double>(C, n, ldc, A, k, lda, alpha); }
int gla_zherk_lower(void* C, int64_t n, int64_t ldc, const void* A, int64_t k, int64_t lda, double alpha) { return herk_host<zd>(ZP(C), n, ldc, ZCP(A), k, lda, alpha); }

}  // extern "C"
