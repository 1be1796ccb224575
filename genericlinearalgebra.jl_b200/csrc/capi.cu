// capi.cu -- the C ABI declared in include/gla_cuda.h.  Host-pointer entry points stage through
// device memory (H2D, kernels, D2H) and are synchronous; `_dev` twins are asynchronous on the
// caller's stream.  No C++ exception crosses this boundary and nothing here computes on the CPU.
#include "../../include/gla_cuda.h"
#include "gla_internal.cuh"

#include <new>

using namespace gla;

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    return check_cuda(e, __FILE__, __LINE__);
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
// Chunked three-stream pipeline: H2D of chunk i+1 and D2H of chunk i-1 overlap the kernel of chunk i
// (PCIe is full duplex), so the end-to-end time is max(H2D, D2H) rather than their sum.
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
  i64 chunk = (64ll << 20) / mat_bytes;
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  constexpr int NS = 3;
  Stream st[NS];
  DevBuf dA[NS], dtau[NS];
  int ns = (int)((batch + chunk - 1) / chunk < NS ? (batch + chunk - 1) / chunk : NS);
  for (int s = 0; s < ns; ++s) {
    GLA_TRY(st[s].create());
    GLA_TRY(dA[s].alloc(chunk * mat_bytes));
    GLA_TRY(dtau[s].alloc(chunk * k * sizeof(T)));
  }
  i64 done = 0;
  int it = 0;
  while (done < batch) {
    const int s = it % ns;
    const i64 nb = batch - done < chunk ? batch - done : chunk;
    T* hA = A + done * m * n;
    T* ht = tau + done * k;
    GLA_CUDA(cudaMemcpyAsync(dA[s].p, hA, nb * mat_bytes, cudaMemcpyHostToDevice, st[s].s));
    GLA_TRY(geqr_batched_dev<T>(dA[s].as<T>(), m, n, nb, dtau[s].as<T>(), st[s].s));
    GLA_CUDA(cudaMemcpyAsync(hA, dA[s].p, nb * mat_bytes, cudaMemcpyDeviceToHost, st[s].s));
    GLA_CUDA(cudaMemcpyAsync(ht, dtau[s].p, nb * k * sizeof(T), cudaMemcpyDeviceToHost, st[s].s));
    done += nb;
    ++it;
  }
  for (int s = 0; s < ns; ++s) GLA_CUDA(cudaStreamSynchronize(st[s].s));
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

}  // extern "C"
