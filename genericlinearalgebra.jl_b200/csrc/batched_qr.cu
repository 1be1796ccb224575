// batched_qr.cu -- K4: batched small Householder QR.
//
// Semantics per matrix = GenericLinearAlgebra.qrBlocked!/qrUnblocked! (reference src/qr.jl:86-146)
// with Julia's stdlib reflector!/reflectorApply! conventions (call sites src/qr.jl:96,102):
//   nu = copysign(||x||, Re x1); x1 <- -nu; x[2:] /= (x1+nu); tau = (x1+nu)/nu  (tau=0 for a zero
//   column, tau=2 for a length-1 column), trailing columns <- (I - conj(tau) v v^H) * columns.
//
// Two kernels:
//  * batched_qr32_reg_kernel<R>: 32x32 real matrices, ONE MATRIX PER WARP, register resident.
//    lane c owns column c (32 registers-worth of rows).  HBM -> smem with 128-bit coalesced loads
//    into a padded (conflict-free) staging tile, smem -> registers, 32 reflector steps where the
//    pivot column is published once through shared memory (broadcast LDS.128), results go back
//    through the same staging tile with 128-bit coalesced stores.  The column scaling by 1/xi is
//    deferred to one pass at the end and folded into the coefficients meanwhile, so each step is
//    two FMA sweeps (dot, axpy) plus one rsqrt and one reciprocal refined by Newton steps.
//  * batched_qr_smem_kernel<T>: any (m,n) whose matrix fits in shared memory, one CTA per matrix
//    (also used for ComplexF64).
#include "common.cuh"
#include "smallqr.cuh"

namespace gla {

// ---------------------------------------------------------------------------------- fast scalars
template <class R>
struct Fast;
template <>
struct Fast<double> {
  // 1/sqrt(x) to ~1 ulp: MUFU.RSQ64H seed (~2^-20) + 2 Newton steps in FP64 FMA
  static __device__ __forceinline__ double rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      double t = x * y;
      double e = fma(-t, y, 1.0);
      y = fma(0.5 * y, e, y);
    }
    return y;
  }
  // sqrt(x) given rs ~ 1/sqrt(x): one correction step -> (almost always) correctly rounded
  static __device__ __forceinline__ double sqrt_from_rsqrt(double x, double rs) {
    double g = x * rs;
    double r = fma(-g, g, x);
    return fma(r, 0.5 * rs, g);
  }
  static __device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      double e = fma(-x, y, 1.0);
      y = fma(y, e, y);
    }
    double e = fma(-x, y, 1.0);
    return fma(y, e, y);
  }
};
template <>
struct Fast<float> {
  static __device__ __forceinline__ float rsqrt(float x) {
    float y = rsqrtf(x);
    float t = x * y;
    float e = fmaf(-t, y, 1.0f);
    return fmaf(0.5f * y, e, y);
  }
  static __device__ __forceinline__ float sqrt_from_rsqrt(float x, float rs) {
    float g = x * rs;
    float r = fmaf(-g, g, x);
    return fmaf(r, 0.5f * rs, g);
  }
  static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
};

template <class R>
struct Vec16;
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int N = 2;
};
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int N = 4;
};

template <class R>
__device__ __forceinline__ void vec_to_arr(const double2& v, R* a) {
  a[0] = v.x;
  a[1] = v.y;
}
template <class R>
__device__ __forceinline__ void vec_to_arr(const float4& v, R* a) {
  a[0] = v.x;
  a[1] = v.y;
  a[2] = v.z;
  a[3] = v.w;
}
__device__ __forceinline__ double2 arr_to_vec(const double* a) { return make_double2(a[0], a[1]); }
__device__ __forceinline__ float4 arr_to_vec(const float* a) { return make_float4(a[0], a[1], a[2], a[3]); }

// ---------------------------------------------------------------------------------- 32x32, warp
template <class R>
struct Reg32Cfg {
  static constexpr int V = Vec16<R>::N;         // elements per 16-byte vector
  static constexpr int LD = 32 + 16 / sizeof(R);  // padded column stride (34 doubles / 36 floats):
                                                  // 16B aligned, LDS.128 of 8 lanes hit 8 distinct 16B slots
  static constexpr int TILE = 32 * LD;            // elements of staging tile per warp
};

template <class R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 384 / (WARPS * 32))
    batched_qr32_reg_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = Reg32Cfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVEC = 32 * 32 / V / 32;  // 16-byte vectors per lane per matrix

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::TILE;

  for (i64 mat = (i64)blockIdx.x * WARPS + warp; mat < batch; mat += (i64)gridDim.x * WARPS) {
    R* Ag = A + mat * 1024;
    // ---- HBM -> staging tile: 16-byte coalesced loads, column-padded stores
    {
      VT v[NVEC];
#pragma unroll
      for (int q = 0; q < NVEC; ++q) v[q] = __ldcs(reinterpret_cast<const VT*>(Ag) + lane + 32 * q);
#pragma unroll
      for (int q = 0; q < NVEC; ++q) {
        int e = (lane + 32 * q) * V;
        *reinterpret_cast<VT*>(sm + (e >> 5) * LD + (e & 31)) = v[q];
      }
    }
    __syncwarp();
    // ---- staging tile -> registers: lane c takes column c
    R a[32];
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      VT v = *reinterpret_cast<const VT*>(sm + lane * LD + i);
      vec_to_arr<R>(v, a + i);
    }
    __syncwarp();

    R my_tau = R(0), my_ixi = R(1);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int k0 = k & ~(V - 1);
      // lane k publishes its (current, un-normalised) column rows k0..31
      if (lane == k) {
#pragma unroll
        for (int i = k0; i < 32; i += V) *reinterpret_cast<VT*>(sm + k * LD + i) = arr_to_vec(a + i);
      }
      __syncwarp();
      const R* vk = sm + k * LD;
      // d = sum_{i>k} a_ik * a_ic  (own column c); lane k obtains its tail norm^2
      R d0 = R(0), d1 = R(0);
      R alpha = R(0);
#pragma unroll
      for (int i = k0; i < 32; i += V) {
        R y[V];
        VT v = *reinterpret_cast<const VT*>(vk + i);
        vec_to_arr<R>(v, y);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (i + j == k) alpha = y[j];
          if (i + j > k) {
            if ((i + j - k) & 1) d0 = fmad(y[j], a[i + j], d0);
            else d1 = fmad(y[j], a[i + j], d1);
          }
        }
      }
      asm volatile("" ::: "memory");  // the axpy sweep below re-reads the pivot column from smem
      const R d = d0 + d1;
      const R dk = __shfl_sync(0xffffffffu, d, k);
      const R n2 = fmad(alpha, alpha, dk);
      if (n2 != R(0)) {  // warp-uniform
        const R rs = Fast<R>::rsqrt(n2);
        const R nrm = Fast<R>::sqrt_from_rsqrt(n2, rs);
        const R nu = copysign(nrm, alpha);
        const R inv_nu = copysign(rs, alpha);
        const R xi = alpha + nu;
        // tau = xi / nu, correctly rounded (one residual step): a length-1 column must give
        // exactly 2 as Julia's division does
        const R tq = xi * inv_nu;
        const R tk = fmad(fmad(-tq, nu, xi), inv_nu, tq);
        const R ixi = Fast<R>::rcp(xi);
        // s = conj(tau) * (a_kc + v^H a_c[k+1:]) with v = a_k/xi  ->  tau*a_kc + d/nu
        const R s = fmad(d, inv_nu, tk * a[k]);
        const bool right = lane > k;
        const R t = right ? s * ixi : R(0);
        a[k] = (lane == k) ? -nu : (right ? a[k] - s : a[k]);
        if (lane == k) {
          my_tau = tk;
          my_ixi = ixi;
        }
        // re-read the pivot column from shared memory (broadcast) instead of holding 32 more
        // values per lane: keeps the kernel under 128 registers -> 16+ resident warps per SM
        const R nt = -t;
#pragma unroll
        for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
          R y[V];
          VT v = *reinterpret_cast<const VT*>(vk + i);
          vec_to_arr<R>(v, y);
#pragma unroll
          for (int j = 0; j < V; ++j)
            if (i + j > k) a[i + j] = fmad(nt, y[j], a[i + j]);
        }
      }
    }
    // deferred normalisation of the stored reflectors: rows below the diagonal *= 1/xi
#pragma unroll
    for (int i = 1; i < 32; ++i) a[i] = (i > lane) ? a[i] * my_ixi : a[i];

    // ---- registers -> staging tile -> HBM (16-byte coalesced, streaming)
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += V) *reinterpret_cast<VT*>(sm + lane * LD + i) = arr_to_vec(a + i);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NVEC; ++q) {
      int e = (lane + 32 * q) * V;
      VT v = *reinterpret_cast<const VT*>(sm + (e >> 5) * LD + (e & 31));
      __stcs(reinterpret_cast<VT*>(Ag) + lane + 32 * q, v);
    }
    tau[mat * 32 + lane] = my_tau;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------- generic, CTA
template <class T>
__global__ void __launch_bounds__(SMALLQR_THREADS)
    batched_qr_smem_kernel(T* __restrict__ A, T* __restrict__ tau, int m, int n, i64 batch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);
  const int k = m < n ? m : n;
  for (i64 mat = blockIdx.x; mat < batch; mat += gridDim.x) {
    T* Ag = A + mat * (i64)m * n;
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) sA[e] = Ag[e];
    __syncthreads();
    cta_qr_smem<T>(sA, m, n, m, tau + mat * k);
    __syncthreads();
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) Ag[e] = sA[e];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------- launchers
template <class T>
struct IsReal {
  static constexpr bool v = true;
};
template <>
struct IsReal<zd> {
  static constexpr bool v = false;
};

template <class R>
static int launch_reg32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  constexpr int WARPS = 4;
  size_t smem = (size_t)WARPS * Reg32Cfg<R>::TILE * sizeof(R);
  auto kern = batched_qr32_reg_kernel<R, WARPS>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  i64 need = (batch + WARPS - 1) / WARPS;
  i64 resident = (i64)sm_count() * occ;
  // whole waves of resident CTAs; each warp then strides over its share of the batch
  i64 grid = need < resident ? need : resident * ((need + resident - 1) / resident > 4 ? 4 : 1);
  if (grid > need) grid = need;
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template <class T>
int geqr_batched_dev(T* dA, i64 m, i64 n, i64 batch, T* dtau, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (batch < 0) return -4;
  if (m == 0 || n == 0 || batch == 0) return 0;
  if constexpr (IsReal<T>::v) {
    if (m == 32 && n == 32) return launch_reg32<T>(dA, dtau, batch, st);
  }
  size_t smem = (size_t)m * n * sizeof(T);
  if (smem > 96 * 1024) return -2;
  auto kern = batched_qr_smem_kernel<T>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  i64 grid = batch < (i64)sm_count() * 16 ? batch : (i64)sm_count() * 16;
  kern<<<(unsigned)grid, SMALLQR_THREADS, smem, st>>>(dA, dtau, (int)m, (int)n, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template int geqr_batched_dev<float>(float*, i64, i64, i64, float*, cudaStream_t);
template int geqr_batched_dev<double>(double*, i64, i64, i64, double*, cudaStream_t);
template int geqr_batched_dev<zd>(zd*, i64, i64, i64, zd*, cudaStream_t);

}  // namespace gla
