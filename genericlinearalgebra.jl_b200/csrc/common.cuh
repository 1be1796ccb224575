// common.cuh -- scalar traits, error plumbing and small device helpers shared by all kernels.
// B200 / sm_100a only.  No CPU fallback lives anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace gla {

using i64 = int64_t;

// ------------------------------------------------------------------ complex<double>
struct __align__(16) zd {
  double x, y;
};

__host__ __device__ __forceinline__ zd make_zd(double a, double b) {
  zd r;
  r.x = a;
  r.y = b;
  return r;
}

// ------------------------------------------------------------------ scalar traits
template <class T>
struct Sc;

template <>
struct Sc<float> {
  using real = float;
  static constexpr bool is_complex = false;
  __host__ __device__ static float zero() { return 0.f; }
  __host__ __device__ static float one() { return 1.f; }
  __host__ __device__ static float from_real(float r) { return r; }
};
template <>
struct Sc<double> {
  using real = double;
  static constexpr bool is_complex = false;
  __host__ __device__ static double zero() { return 0.; }
  __host__ __device__ static double one() { return 1.; }
  __host__ __device__ static double from_real(double r) { return r; }
};
template <>
struct Sc<zd> {
  using real = double;
  static constexpr bool is_complex = true;
  __host__ __device__ static zd zero() { return make_zd(0., 0.); }
  __host__ __device__ static zd one() { return make_zd(1., 0.); }
  __host__ __device__ static zd from_real(double r) { return make_zd(r, 0.); }
};

// real types
__host__ __device__ __forceinline__ float cj(float a) { return a; }
__host__ __device__ __forceinline__ double cj(double a) { return a; }
__host__ __device__ __forceinline__ float re(float a) { return a; }
__host__ __device__ __forceinline__ double re(double a) { return a; }
__host__ __device__ __forceinline__ float abs2(float a) { return a * a; }
__host__ __device__ __forceinline__ double abs2(double a) { return a * a; }
// fma: acc + a*b
__device__ __forceinline__ float fmad(float a, float b, float acc) { return fmaf(a, b, acc); }
__device__ __forceinline__ double fmad(double a, double b, double acc) { return fma(a, b, acc); }
__host__ __device__ __forceinline__ float scale_real(float a, float r) { return a * r; }
__host__ __device__ __forceinline__ double scale_real(double a, double r) { return a * r; }
__host__ __device__ __forceinline__ bool is_zero(float a) { return a == 0.f; }
__host__ __device__ __forceinline__ bool is_zero(double a) { return a == 0.; }

// complex
__host__ __device__ __forceinline__ zd cj(zd a) { return make_zd(a.x, -a.y); }
__host__ __device__ __forceinline__ double re(zd a) { return a.x; }
__host__ __device__ __forceinline__ double abs2(zd a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ zd operator+(zd a, zd b) { return make_zd(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ zd operator-(zd a, zd b) { return make_zd(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ zd operator-(zd a) { return make_zd(-a.x, -a.y); }
__host__ __device__ __forceinline__ zd operator*(zd a, zd b) {
  return make_zd(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ zd& operator+=(zd& a, zd b) {
  a.x += b.x;
  a.y += b.y;
  return a;
}
__host__ __device__ __forceinline__ zd& operator-=(zd& a, zd b) {
  a.x -= b.x;
  a.y -= b.y;
  return a;
}
__device__ __forceinline__ zd fmad(zd a, zd b, zd acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}
__host__ __device__ __forceinline__ zd scale_real(zd a, double r) { return make_zd(a.x * r, a.y * r); }
__host__ __device__ __forceinline__ bool is_zero(zd a) { return a.x == 0. && a.y == 0.; }
// 1/z (Smith-free: inputs here are xi = alpha + nu with |xi| >= ||x||, never tiny relative to parts)
__host__ __device__ __forceinline__ zd zinv(zd a) {
  double d = 1.0 / (a.x * a.x + a.y * a.y);
  return make_zd(a.x * d, -a.y * d);
}
__host__ __device__ __forceinline__ float inv(float a) { return 1.f / a; }
__host__ __device__ __forceinline__ double inv(double a) { return 1. / a; }
__host__ __device__ __forceinline__ zd inv(zd a) { return zinv(a); }

// ------------------------------------------------------------------ L2-only loads
// Data handed from one kernel to the next at a FIXED address (per-panel T, Gram, split-K partials, Z, the C tile of
// an in-place update) is read with ld.global.cg: kernels of the two look-ahead streams share SMs, and a stale L1 /
// read-only-cache line of the previous contents of such a buffer was observed to survive the (other stream's) kernel
// boundary on a busy SM.  L2 is the point of coherence, so .cg loads cannot see stale data.
__device__ __forceinline__ float ldcg_t(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcg_t(const double* p) { return __ldcg(p); }
__device__ __forceinline__ zd ldcg_t(const zd* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return make_zd(v.x, v.y);
}

// ------------------------------------------------------------------ warp helpers
template <class R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ zd warp_sum(zd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

template <class T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
template <>
__device__ __forceinline__ zd shfl_xor_t<zd>(zd v, int m) {
  return make_zd(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

constexpr int GLA_ERR_DRIVER = 1900;   // driver entry point / tensor-map encode failure
constexpr int GLA_ERR_INTERNAL = 1901; // internal invariant violated

// ------------------------------------------------------------------ errors
void set_error(int code, const char* what, const char* file, int line);
int check_cuda(cudaError_t e, const char* file, int line);

#define GLA_CUDA(call)                                                   \
  do {                                                                   \
    int _rc = ::gla::check_cuda((call), __FILE__, __LINE__);             \
    if (_rc) return _rc;                                                 \
  } while (0)
#define GLA_TRY(call)       \
  do {                      \
    int _rc = (call);       \
    if (_rc) return _rc;    \
  } while (0)

inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }
inline i64 round_up(i64 a, i64 b) { return (a + b - 1) / b * b; }

int sm_count();  // of the current device (cached)
// stream-ordered allocation from the library-owned pool of the current device; release with cudaFreeAsync
int pool_malloc(void** p, size_t bytes, cudaStream_t st);
// High-priority side streams + six timing-free events for the look-ahead schedules (blocked QR, Cholesky), created once
// per (host thread, device) and reused by every call: work is stream-ordered, so consecutive asynchronous calls of one
// thread may share them.  Released when the thread exits.
struct AuxCtx {
  cudaStream_t hi = nullptr;
  cudaStream_t hi2 = nullptr;   // second high-priority stream: the T build of a QR panel beside the W product (qr_blocked.cu)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
int aux_ctx(AuxCtx** out);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size): the launch paths are host-bound for
// small problems (hundreds of dependent launches), so the per-launch driver call is worth avoiding
int ensure_dyn_smem(const void* func, int bytes);

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while its predecessor in the
// stream is still running; it must call pdl_wait() before its first global-memory access (the wait returns when the
// predecessor grids have completed and their writes are visible).  pdl_launch_dependents() at the top of a kernel lets
// the NEXT kernel's launch overlap this one.  The chains of small dependent kernels (T build, split sums, small products,
// Cholesky panels) are launch-latency bound; GLA_NO_PDL=1 switches the attribute off for A/B.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
long long pdl_max_n();   // largest problem order the drivers open a PdlScope for (GLA_PDL_MAXN, default 2048)
// The attribute is only set inside a PdlScope(true) of the calling thread: the Cholesky driver opens one always, the
// blocked-QR driver for n <= GLA_PDL_MAXN (default 2048), where the launch chain is what bounds the run.  Measured: blocked
// QR n = 1024 2.86 -> 2.69 ms, n = 2048 6.77 -> 6.49 ms (neutral from n = 4096 on), Cholesky n = 2048 1.33 -> 1.26 ms,
// n = 4096 3.13 -> 3.02 ms.  Where the trigger sits matters: at the TOP of the kernels the early-scheduled CTAs of the chain
// sat on SM slots that the concurrent bulk update of the look-ahead schedule needs (QR n = 16384 238 -> 243 ms, Cholesky
// n = 8192 11.0 -> 12.2 ms); the kernels now trigger after their main loop, right before the final stores.
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on);
  ~PdlScope();
};
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace gla
