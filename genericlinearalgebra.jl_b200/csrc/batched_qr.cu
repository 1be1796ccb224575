// batched_qr.cu -- K4: batched small Householder QR.
//
// Semantics per matrix = GenericLinearAlgebra.qrBlocked!/qrUnblocked! (reference src/qr.jl:86-146)
// with Julia's stdlib reflector!/reflectorApply! conventions (call sites src/qr.jl:96,102):
//   nu = copysign(||x||, Re x1); x1 <- -nu; x[2:] /= (x1+nu); tau = (x1+nu)/nu  (tau=0 for a zero
//   column, tau=2 for a length-1 column), trailing columns <- (I - conj(tau) v v^H) * columns.
//
// Kernels:
//  * batched_qr32_ll4_kernel<R, 12, 1, 300, 2, true> -- 32x32 Float32/Float64 (195 M matrices/s in Float64): two matrices
//    per warp (one per half-warp), lane = column, left-looking in two 16-column halves, one padded work half tile S and one
//    incoming half tile P per matrix, P filled by cp.async with the next half while the current one is factorised, ONE CTA of
//    12 warps per SM whose warps meet at a barrier before every pair and leave it 300 cycles apart (instruction-cache
//    sharing without lock step), pivot column loaded once per step.  GLA_BATCHED_VARIANT=1 selects the round-1 form (pivot
//    column re-read by the axpy sweep) for A/B runs.  Earlier generations and the round-2 four-matrices-per-warp experiment
//    live in tools/retired/ and are not part of the library.
//  * batched_qr_warp_kernel<T>: n <= 32, m <= 64 (any element type, e.g. ComplexF64 32 x 32): one WARP per matrix, eight per CTA.
//  * batched_qr_smem_kernel<T>: any other (m,n) whose matrix fits in shared memory, one CTA per matrix.
// What bounds the 32x32 kernel: DESIGN.md section 8 (operand bandwidth of the FP64 pipe: a DFMA with three distinct
// register-pair operands issues every 3.5 cycles per SM sub-partition, tools/fp64_pattern.cu).
#include "common.cuh"
#include "smallqr.cuh"
#include "fastmath.cuh"

#include <stdlib.h>

namespace gla {

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

template <class R>
struct Seed;
template <>
struct Seed<double> {
  static __device__ __forceinline__ double rsqrt0(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
  }
  static __device__ __forceinline__ double rcp0(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
  }
  static constexpr int ITERS = 2;
};
template <>
struct Seed<float> {
  static __device__ __forceinline__ float rsqrt0(float x) { return rsqrtf(x); }
  static __device__ __forceinline__ float rcp0(float x) { return __frcp_rn(x); }
  static constexpr int ITERS = 1;
};

// The scalars of one reflector (Julia's reflector!, call site src/qr.jl:96) from alpha = x[1] and n2 = ||x||^2 > 0:
//   nu = copysign(sqrt(n2), alpha), inv_nu = 1/nu, xi = alpha + nu, r = 1/xi, tq = tau = xi/nu.
// Goldschmidt sqrt/rsqrt (g -> sqrt, hh -> 1/(2 sqrt)) plus a Newton reciprocal whose MUFU seed is taken from the first
// Goldschmidt iterate so that it overlaps the iterations (no IEEE sqrt / division subroutine on the chain).
template <class R>
__device__ __forceinline__ void reflector_chain(const R alpha, const R n2s, R& nu, R& inv_nu, R& r, R& tq) {
  const R y0 = Seed<R>::rsqrt0(n2s);
  R g = n2s * y0, hh = R(0.5) * y0;
  r = Seed<R>::rcp0(alpha + copysign(g, alpha));
#pragma unroll
  for (int it = 0; it < Seed<R>::ITERS; ++it) {
    const R e = fmad(-g, hh, R(0.5));
    g = fmad(g, e, g);
    hh = fmad(hh, e, hh);
  }
  nu = copysign(g, alpha);
  inv_nu = copysign(hh + hh, alpha);
  const R xi = alpha + nu;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const R e = fmad(-xi, r, R(1));
    r = fmad(r, e, r);
  }
  tq = xi * inv_nu;
}

// ====================================================================================== 32x32, two matrices per warp
// Two matrices per warp (one per half-warp), left-looking in two
// 16-column halves, lane = column, cp.async prefetch of the next
// half, ONE CTA of 12 warps per SM meeting at a barrier before every pair and leaving it 300 cycles apart.
template <class R>
struct Ll4Cfg {
  static constexpr int V = Vec16<R>::N;
  static constexpr int LD = 32 + 16 / sizeof(R);
  static constexpr int HALF = 16 * LD + 16 / sizeof(R);   // +16 B: the two half tiles of a warp sit 4 banks apart
  static constexpr int PER_WARP = 4 * HALF;               // S (work) and P (incoming) half tiles of both matrices of the pair
};

// Factorise 16 columns (K0 .. K0+15) right-looking inside the half-warp; lane c owns column K0 + c.
// ONE: the pivot column is loaded once per step into registers and used by the dot and the axpy sweep (half the broadcast
// LDS.128 of the round-1 form, which re-read it; same speed within 0.5 %, see profiles/r02_sweep_batched_q8_one_la.txt --
// the kernel is bound by the FP64 pipe's operand bandwidth, not by shared memory: profiles/r02_fp64_operand_pattern.txt).
template <class R, int K0, int NACC, bool ONE>
__device__ __forceinline__ void ll_factor_half_c(R (&a)[32], R* __restrict__ pub, const int c, R& tau_own, R& ixi_own) {
  using VT = typename Vec16<R>::type;
  constexpr int V = Vec16<R>::N;
  constexpr int LD = Ll4Cfg<R>::LD;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int k = K0 + kk;
    const bool own = c == kk;
    if (k == 31) {  // length-1 column: still reflected, x1 <- -x1, tau = 2 exactly (tau = 0 for a zero entry)
      const bool z = a[31] == R(0);
      tau_own = own ? (z ? R(0) : R(2)) : tau_own;
      a[31] = (own && !z) ? -a[31] : a[31];
      continue;
    }
    const int k0 = k & ~(V - 1);
    R* vk = pub + kk * LD;
    if (own) {
#pragma unroll
      for (int i = k0; i < 32; i += V) *reinterpret_cast<VT*>(vk + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    R alpha = R(0);
    R acc[4] = {R(0), R(0), R(0), R(0)};
    R yy[ONE ? 32 : 1];
    if (ONE) {
#pragma unroll
      for (int i = k0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), yy + i);
      alpha = yy[k];
#pragma unroll
      for (int r2 = k + 1; r2 < 32; ++r2) acc[(r2 - k) & (NACC - 1)] = fmad(yy[r2], a[r2], acc[(r2 - k) & (NACC - 1)]);
    } else {
#pragma unroll
      for (int i = k0; i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int r2 = i + j;
          if (r2 == k) alpha = y[j];
          if (r2 > k) acc[(r2 - k) & (NACC - 1)] = fmad(y[j], a[r2], acc[(r2 - k) & (NACC - 1)]);
        }
      }
    }
    const R d = NACC == 4 ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : NACC == 2 ? acc[0] + acc[1] : acc[0];
    const R dk = __shfl_sync(0xffffffffu, d, kk, 16);  // the owner's dot is the tail norm^2
    // NOTE (documented in include/gla_cuda.h): the sum of squares is unscaled here -- columns must satisfy
    // 1e-140 < ||x|| < 1e140 (1e-15 .. 1e15 in Float32).  A rescaling slow path behind a warp vote was measured at
    // 133 M matrices/s against 194 (register pressure: profiles/r02_sweep_batched_q8_one_la.txt); the generic
    // shared-memory kernel (any other shape, ComplexF64) does take the scaled norm like Julia's norm(x).
    const R n2 = fmad(alpha, alpha, dk);
    const bool zero = n2 == R(0);
    R nu, inv_nu, r, tq;
    reflector_chain<R>(alpha, zero ? R(1) : n2, nu, inv_nu, r, tq);
    // s = tau * (a_kc + v^T a_c[k+1:]) with v = x / xi  ->  tau * a_kc + d / nu
    const R s = fmad(d, inv_nu, tq * a[k]);
    const bool right = (c > kk) && !zero;
    const bool mine = own && !zero;
    const R nt = right ? -(s * r) : R(0);
    a[k] = mine ? -nu : (right ? a[k] - s : a[k]);
    tau_own = mine ? tq : tau_own;
    ixi_own = mine ? r : ixi_own;
    if (ONE) {
#pragma unroll
      for (int rr = k + 1; rr < 32; ++rr) a[rr] = fmad(nt, yy[rr], a[rr]);
    } else {
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int rr = i + j;
          if (rr > k) a[rr] = fmad(nt, y[j], a[rr]);
        }
      }
    }
  }
  // deferred normalisation of the stored reflector: rows below the diagonal *= 1/xi
#pragma unroll
  for (int i = K0 + 1; i < 32; ++i) a[i] = (i > K0 + c) ? a[i] * ixi_own : a[i];
}

template <class R, int WARPS, int SYNC, int STAG, int NACC = 2, bool ONE = true>
__global__ void __launch_bounds__(WARPS * 32, 12 / WARPS)
    batched_qr32_ll4_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = Ll4Cfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVH = 512 / V / 32;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, c = lane & 15;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::PER_WARP;
  R* S = sm + h * Cfg::HALF;
  R* Pw = sm + 2 * Cfg::HALF;
  R* P = Pw + h * Cfg::HALF;

  const i64 npairs = (batch + 1) >> 1;
  const i64 stride = (i64)gridDim.x * WARPS;
  auto prefetch_half = [&](i64 pr, int half) {
    const bool two = pr * 2 + 1 < batch;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const R* src = A + (pr * 2 + ((m == 1 && !two) ? 0 : m)) * 1024 + half * 512;
      const uint32_t nbytes = (m == 1 && !two) ? 0u : 16u;
#pragma unroll
      for (int u = 0; u < NVH; ++u) {
        const int e = (lane + 32 * u) * V;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Pw + m * Cfg::HALF + (e >> 5) * LD + (e & 31));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + e), "r"(nbytes) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto wait_prefetch = [&]() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  };
  auto store_half = [&](R* Ag, int m, int half, bool live) {
    if (!live) return;
#pragma unroll
    for (int u = 0; u < NVH; ++u) {
      const int e = (lane + 32 * u) * V;
      __stcs(reinterpret_cast<VT*>(Ag + m * 1024 + half * 512) + lane + 32 * u,
             *reinterpret_cast<const VT*>(sm + m * Cfg::HALF + (e >> 5) * LD + (e & 31)));
    }
  };

  if ((i64)blockIdx.x * WARPS + warp < npairs) prefetch_half((i64)blockIdx.x * WARPS + warp, 0);
  for (i64 base = (i64)blockIdx.x * WARPS; base < npairs; base += stride) {
    if (SYNC) __syncthreads();
    if (STAG > 0 && warp) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)warp * STAG) {
      }
    }
    const i64 pair = base + warp;
    if (pair >= npairs) {
      if (SYNC) continue;
      break;
    }
    const i64 mat0 = pair * 2;
    const bool both = mat0 + 1 < batch;
    R* Ag = A + mat0 * 1024;
    R tau_l = R(0), tau_r = R(0), ixi = R(1);
    R a[32];
    wait_prefetch();
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(P + c * LD + i), a + i);
    __syncwarp();
    prefetch_half(pair, 1);
    ll_factor_half_c<R, 0, NACC, ONE>(a, S, c, tau_l, ixi);
    __syncwarp();
    R b[32];
    wait_prefetch();
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      vec_to_arr<R>(*reinterpret_cast<const VT*>(P + c * LD + i), b + i);
      *reinterpret_cast<VT*>(S + c * LD + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    if (pair + stride < npairs) prefetch_half(pair + stride, 0);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const R tk = __shfl_sync(0xffffffffu, tau_l, k, 16);
      const R* vk = S + k * LD;
      R acc[4] = {b[k], R(0), R(0), R(0)};
      if (ONE) {
        R yy[32];
#pragma unroll
        for (int i = (k + 1) & ~(V - 1); i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), yy + i);
#pragma unroll
        for (int r = k + 1; r < 32; ++r) acc[(r - k) & (NACC - 1)] = fmad(yy[r], b[r], acc[(r - k) & (NACC - 1)]);
        const R ns = -(tk * (NACC == 4 ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : NACC == 2 ? acc[0] + acc[1] : acc[0]));
        b[k] += ns;
#pragma unroll
        for (int r = k + 1; r < 32; ++r) b[r] = fmad(ns, yy[r], b[r]);
        continue;
      }
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (i + j > k) acc[(i + j - k) & (NACC - 1)] = fmad(y[j], b[i + j], acc[(i + j - k) & (NACC - 1)]);
      }
      const R ns = -(tk * (NACC == 4 ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : NACC == 2 ? acc[0] + acc[1] : acc[0]));
      b[k] += ns;
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (i + j > k) b[i + j] = fmad(ns, y[j], b[i + j]);
      }
    }
    store_half(Ag, 0, 0, true);
    store_half(Ag, 1, 0, both);
    __syncwarp();
    ixi = R(1);
    ll_factor_half_c<R, 16, NACC, ONE>(b, S, c, tau_r, ixi);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += V) *reinterpret_cast<VT*>(S + c * LD + i) = arr_to_vec(b + i);
    __syncwarp();
    store_half(Ag, 0, 1, true);
    store_half(Ag, 1, 1, both);
    if (both || h == 0) {
      tau[(mat0 + h) * 32 + c] = tau_l;
      tau[(mat0 + h) * 32 + 16 + c] = tau_r;
    }
    __syncwarp();
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

template <class R, int WARPS, int SYNC, int STAG, int NACC = 2, bool ONE = true>
static int launch_ll4_32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  const size_t smem = (size_t)WARPS * Ll4Cfg<R>::PER_WARP * sizeof(R);
  auto kern = batched_qr32_ll4_kernel<R, WARPS, SYNC, STAG, NACC, ONE>;
  GLA_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem));
  const i64 npairs = (batch + 1) / 2;
  const i64 need = (npairs + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count();
  const i64 grid = need < resident ? need : resident;
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

// ====================================================================================== generic, one CTA per matrix
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

// ====================================================================================== generic small shapes, one WARP per matrix
// n <= 32 and a matrix of at most WQ_MAX_BYTES: WQ warps per CTA, each with its own matrix in its own slice of shared
// memory (leading dimension padded so that the lanes' columns start in different banks); grid-stride over the batch.
constexpr int WQ_WARPS = 8;
constexpr int WQ_MAX_BYTES = 24 * 1024;
template <class T>
__global__ void __launch_bounds__(WQ_WARPS * 32) batched_qr_warp_kernel(T* __restrict__ A, T* __restrict__ tau, int m, int n, int ld,
                                                                        i64 batch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* sA = reinterpret_cast<T*>(smem_raw) + (size_t)warp * ld * n;
  const int k = m < n ? m : n;
  for (i64 mat = (i64)blockIdx.x * WQ_WARPS + warp; mat < batch; mat += (i64)gridDim.x * WQ_WARPS) {
    T* Ag = A + mat * (i64)m * n;
    for (int e = lane; e < m * n; e += 32) {   // coalesced along the rows of the column-major matrix
      const int c = e / m, r = e - c * m;
      sA[c * ld + r] = Ag[e];
    }
    __syncwarp();
    warp_qr_smem<T>(sA, m, n, ld, tau + mat * k);
    __syncwarp();
    for (int e = lane; e < m * n; e += 32) {
      const int c = e / m, r = e - c * m;
      Ag[e] = sA[c * ld + r];
    }
    __syncwarp();
  }
}

// ====================================================================================== launchers
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
  // GLA_BATCHED_VARIANT: A/B switches, read once per process (default = measured best)
  static const int variant = [] {
    const char* e = getenv("GLA_BATCHED_VARIANT");
    return e ? atoi(e) : 0;
  }();
  switch (variant) {
    case 1: return launch_ll4_32<R, 12, 1, 300, 2, false>(dA, dtau, batch, st);   // round-1 form: pivot column re-read by the axpy sweep
    default: return launch_ll4_32<R, 12, 1, 300, 2, true>(dA, dtau, batch, st);
  }
}

template <class T>
int geqr_batched_dev(T* dA, i64 m, i64 n, i64 batch, T* dtau, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (batch < 0) return -4;
  if (m == 0 || n == 0 || batch == 0) return 0;
  if constexpr (IsReal<T>::v) {
    // the register kernels move 16-byte vectors (cp.async / LDS.128 / STG.128); a matrix stack that is not 16-byte aligned
    // (an offset view) goes through the generic shared-memory kernel below, which only uses element accesses
    if (m == 32 && n == 32 && (reinterpret_cast<uintptr_t>(dA) & 15) == 0 && (reinterpret_cast<uintptr_t>(dtau) & 15) == 0)
      return launch_reg32<T>(dA, dtau, batch, st);
  }
  {
    // small shapes: a warp per matrix (GLA_BATCHED_NO_WARP=1: one CTA per matrix as in round 1, for A/B)
    static const bool no_warp = getenv("GLA_BATCHED_NO_WARP") != nullptr;
    // odd leading dimension in 16-byte units: the 32 lanes' columns start in different bank groups
    int ld = (int)m;
    const int unit = 16 / (int)sizeof(T) > 1 ? 16 / (int)sizeof(T) : 1;
    ld = (int)round_up(ld, unit);
    if (((ld / unit) & 1) == 0) ld += unit;
    const size_t bytes = (size_t)ld * n * sizeof(T);
    if (!no_warp && n <= 32 && m <= 64 && bytes <= (size_t)WQ_MAX_BYTES) {
      auto kern = batched_qr_warp_kernel<T>;
      const size_t smem = bytes * WQ_WARPS;
      GLA_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem));
      const i64 ctas = (batch + WQ_WARPS - 1) / WQ_WARPS;
      const i64 grid = ctas < (i64)sm_count() * 8 ? ctas : (i64)sm_count() * 8;
      kern<<<(unsigned)grid, WQ_WARPS * 32, smem, st>>>(dA, dtau, (int)m, (int)n, ld, batch);
      GLA_CUDA(cudaGetLastError());
      return 0;
    }
  }
  size_t smem = (size_t)m * n * sizeof(T);
  if (smem > 96 * 1024) return -2;
  auto kern = batched_qr_smem_kernel<T>;
  GLA_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kern), 96 * 1024));
  i64 grid = batch < (i64)sm_count() * 16 ? batch : (i64)sm_count() * 16;
  kern<<<(unsigned)grid, SMALLQR_THREADS, smem, st>>>(dA, dtau, (int)m, (int)n, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template int geqr_batched_dev<float>(float*, i64, i64, i64, float*, cudaStream_t);
template int geqr_batched_dev<double>(double*, i64, i64, i64, double*, cudaStream_t);
template int geqr_batched_dev<zd>(zd*, i64, i64, i64, zd*, cudaStream_t);

}  // namespace gla
