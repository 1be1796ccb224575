// batched_qr.cu -- K4: batched small Householder QR.
//
// Semantics per matrix = GenericLinearAlgebra.qrBlocked!/qrUnblocked! (reference src/qr.jl:86-146)
// with Julia's stdlib reflector!/reflectorApply! conventions (call sites src/qr.jl:96,102):
//   nu = copysign(||x||, Re x1); x1 <- -nu; x[2:] /= (x1+nu); tau = (x1+nu)/nu  (tau=0 for a zero
//   column, tau=2 for a length-1 column), trailing columns <- (I - conj(tau) v v^H) * columns.
//
// Kernels (the 32x32 real case has a lineage; GLA_BATCHED_VARIANT selects the predecessors for A/B runs, see launch_reg32):
//  * batched_qr32_ll4_kernel<R, 12, 1, 300, 2, true>  -- DEFAULT for 32x32 Float32/Float64 (195 M matrices/s in Float64):
//    two matrices per warp (one per half-warp), lane = column, left-looking in two 16-column halves, one padded work
//    half tile S and one incoming half tile P per matrix, P filled by cp.async with the next half while the current one is
//    factorised, ONE CTA of 12 warps per SM whose warps meet at a barrier before every pair and leave it 300 cycles apart
//    (instruction-cache sharing without lock step), two dot accumulators, the finished left half stored after phase 2.
//  * batched_qr32_ll2_kernel (no prefetch tile: 152-167 M/s), batched_qr32_ll_kernel (full tile per matrix: 141),
//    batched_qr32_hw_kernel (both columns of a lane pair in registers: 125-132), batched_qr32_reg_kernel (one matrix per
//    warp: the first design).  All share the reflector conventions above and the deferred 1/xi normalisation; the scalar
//    chain per reflector is a Goldschmidt sqrt/rsqrt plus a Newton reciprocal whose MUFU seeds issue together.
//  * batched_qr_smem_kernel<T>: any (m,n) whose matrix fits in shared memory, one CTA per matrix (also ComplexF64).
// What bounds the default kernel and what was tried is in DESIGN.md section 8 (1) and profiles/r01_s5_sweep_*.txt.
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

// ---------------------------------------------------------------------------------- 32x32, warp
template <class R>
struct Reg32Cfg {
  static constexpr int V = Vec16<R>::N;           // elements per 16-byte vector
  static constexpr int LD = 32 + 16 / sizeof(R);  // padded column stride (34 doubles / 36 floats):
                                                  // 16B aligned, LDS.128 of 8 lanes hit 8 distinct 16B slots
  static constexpr int TILE = 32 * LD;            // elements of staging tile per warp
};

// 16-byte shared-memory load the compiler may neither hoist nor merge: keeps at most two chunks of
// the pivot column live (the register budget is what bounds the number of matrices in flight per SM,
// and the kernel is latency bound, so occupancy is throughput)
// `dep` is a fake input: it orders the load after the FMA that produced it, so the scheduler cannot
// cluster all loads of a sweep ahead of the arithmetic (which would keep the whole pivot column live).
__device__ __forceinline__ void lds16(const double* p, double* out, double dep) {
  asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];"
               : "=d"(out[0]), "=d"(out[1])
               : "r"((uint32_t)__cvta_generic_to_shared(p)), "d"(dep));
}
__device__ __forceinline__ void lds16(const float* p, float* out, float dep) {
  asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)), "f"(dep));
}
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

template <class R, int WARPS, int MINB>
__global__ void __maxnreg__(MINB)
    batched_qr32_reg_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = Reg32Cfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVEC = 32 * 32 / V / 32;  // 16-byte vectors per lane per matrix
  constexpr int CH = 4;                   // rows per chunk of the pivot column

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::TILE;

  for (i64 mat = (i64)blockIdx.x * WARPS + warp; mat < batch; mat += (i64)gridDim.x * WARPS) {
    R* Ag = A + mat * 1024;
    // ---- HBM -> staging tile: 16-byte coalesced streaming loads, column-padded stores
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      VT v[NVEC / 2];
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) v[q] = __ldcs(reinterpret_cast<const VT*>(Ag) + lane + 32 * (q + half * (NVEC / 2)));
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) {
        const int e = (lane + 32 * (q + half * (NVEC / 2))) * V;
        *reinterpret_cast<VT*>(sm + (e >> 5) * LD + (e & 31)) = v[q];
      }
    }
    __syncwarp();
    // ---- staging tile -> registers: lane c takes column c
    R a[32];
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(sm + lane * LD + i), a + i);
    __syncwarp();

    R my_tau = R(0), my_ixi = R(1);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      // ---- the pivot column (current, un-normalised) is broadcast from lane k with warp shuffles:
      // no shared-memory round trip, no divergent publish branch, no warp barrier
      R acc[4] = {R(0), R(0), R(0), R(0)};
      R x[32];
#pragma unroll
      for (int r = k; r < 32; ++r) x[r] = __shfl_sync(0xffffffffu, a[r], k);
      // d = sum_{i>k} a_ik * a_ic (own column c); lane k obtains its tail norm^2
#pragma unroll
      for (int r = k + 1; r < 32; ++r) acc[(r - k) & 3] = fmad(x[r], a[r], acc[(r - k) & 3]);
      const R alpha = x[k];
      const R d = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      const R dk = __shfl_sync(0xffffffffu, d, k);
      const R n2 = fmad(alpha, alpha, dk);
      const bool zero = n2 == R(0);  // zero column: tau = 0, nothing changes (branch-free: guarded selects)
      const R n2s = zero ? R(1) : n2;
      // Goldschmidt: g -> sqrt(n2), hh -> 1/(2 sqrt(n2)); the reciprocal of xi is seeded from the
      // approximate norm so its MUFU latency overlaps these iterations
      const R y0 = Seed<R>::rsqrt0(n2s);
      R g = n2s * y0, hh = R(0.5) * y0;
      R r = Seed<R>::rcp0(alpha + copysign(g, alpha));
#pragma unroll
      for (int it = 0; it < Seed<R>::ITERS; ++it) {
        const R e = fmad(-g, hh, R(0.5));
        g = fmad(g, e, g);
        hh = fmad(hh, e, hh);
      }
      const R nu = copysign(g, alpha);
      const R inv_nu = copysign(hh + hh, alpha);
      const R xi = alpha + nu;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const R e = fmad(-xi, r, R(1));
        r = fmad(r, e, r);
      }
      const R tq = xi * inv_nu;  // tau = xi / nu
      // s = conj(tau) * (a_kc + v^H a_c[k+1:]) with v = a_k/xi  ->  tau*a_kc + d/nu
      const R s = fmad(d, inv_nu, tq * a[k]);
      const bool right = (lane > k) && !zero;
      const R nt = right ? -(s * r) : R(0);
      {
        // off the critical path: correctly rounded tau (a length-1 column must give exactly 2, as
        // Julia's division does) and a last correction step for the stored norm
        const bool mine = (lane == k) && !zero;
        const R nuc = copysign(fmad(fmad(-g, g, n2s), hh, g), alpha);
        const R xic = alpha + nuc;
        const R tqc = xic * inv_nu;
        const R tk = fmad(fmad(-tqc, nuc, xic), inv_nu, tqc);
        my_tau = mine ? tk : my_tau;
        my_ixi = mine ? r : my_ixi;
        a[k] = mine ? -nuc : (right ? a[k] - s : a[k]);
      }
      // ---- axpy sweep: a_ic -= (s/xi) * a_ik
#pragma unroll
      for (int rr = k + 1; rr < 32; ++rr) a[rr] = fmad(nt, x[rr], a[rr]);
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
      const int e = (lane + 32 * q) * V;
      const VT v = *reinterpret_cast<const VT*>(sm + (e >> 5) * LD + (e & 31));
      __stcs(reinterpret_cast<VT*>(Ag) + lane + 32 * q, v);
    }
    tau[mat * 32 + lane] = my_tau;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------- 32x32, half-warp
// TWO matrices per warp: half-warp h owns matrix h, lane c of the half owns columns c and c+16
// (64 payload values per lane).  Compared with one matrix per warp this halves the per-matrix cost
// of the scalar chain (sqrt / reciprocal / tau are computed once per warp-step for two matrices) and
// of the pivot-column traffic, and keeps every lane busy for the first 16 steps (column c+16 is always
// to the right of the pivot).  Measured on B200: 125 M matrices/s vs 110 M for one matrix per warp.
template <class R>
struct HwCfg {
  static constexpr int V = Vec16<R>::N;
  static constexpr int LD = 32 + 16 / sizeof(R);
  static constexpr int TILE = 32 * LD + 16 / sizeof(R);   // +16 B: the two tiles of a warp sit 4 banks apart
  static constexpr int PER_WARP = 2 * TILE;
};

template <class R, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    batched_qr32_hw_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = HwCfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVEC = 2 * 1024 / V / 32;  // 16-byte vectors per lane per matrix PAIR

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, c = lane & 15;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::PER_WARP;
  R* my = sm + h * Cfg::TILE;  // tile of this half-warp's matrix

  const i64 npairs = (batch + 1) >> 1;
  for (i64 pair = (i64)blockIdx.x * WARPS + warp; pair < npairs; pair += (i64)gridDim.x * WARPS) {
    const i64 mat0 = pair * 2;
    const bool both = mat0 + 1 < batch;
    R* Ag = A + mat0 * 1024;
    // ---- HBM -> staging tiles (two matrices = 2048 contiguous elements), 16-byte coalesced streaming loads
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      VT v[NVEC / 2];
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) {
        const int p = lane + 32 * (q + half * (NVEC / 2));
        const bool ok = both || (p * V < 1024);
        v[q] = ok ? __ldcs(reinterpret_cast<const VT*>(Ag) + p) : VT{};
      }
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) {
        const int e = (lane + 32 * (q + half * (NVEC / 2))) * V;
        const int hm = e >> 10, col = (e & 1023) >> 5, row = e & 31;
        *reinterpret_cast<VT*>(sm + hm * Cfg::TILE + col * LD + row) = v[q];
      }
    }
    __syncwarp();
    R a0[32], a1[32];
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      vec_to_arr<R>(*reinterpret_cast<const VT*>(my + c * LD + i), a0 + i);
      vec_to_arr<R>(*reinterpret_cast<const VT*>(my + (c + 16) * LD + i), a1 + i);
    }
    __syncwarp();

    R tau0 = R(0), tau1 = R(0), ixi0 = R(1), ixi1 = R(1);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int k0 = k & ~(V - 1);
      const bool lo = k < 16;                 // compile-time after unrolling
      const bool own = c == (k & 15);
      // the owner of the pivot column publishes it (current, un-normalised) to its matrix' tile
      if (own) {
#pragma unroll
        for (int i = k0; i < 32; i += V)
          *reinterpret_cast<VT*>(my + k * LD + i) = lo ? arr_to_vec(a0 + i) : arr_to_vec(a1 + i);
      }
      __syncwarp();
      const R* vk = my + k * LD;
      // ---- dots with both owned columns; the owner obtains the tail norm^2
      R p0 = R(0), p1 = R(0), q0 = R(0), q1 = R(0), alpha = R(0);
#pragma unroll
      for (int i = k0; i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int r = i + j;
          if (r == k) alpha = y[j];
          if (r > k) {
            if ((r - k) & 1) {
              if (lo) p0 = fmad(y[j], a0[r], p0);
              q0 = fmad(y[j], a1[r], q0);
            } else {
              if (lo) p1 = fmad(y[j], a0[r], p1);
              q1 = fmad(y[j], a1[r], q1);
            }
          }
        }
      }
      const R d0 = p0 + p1, d1 = q0 + q1;
      const R dk = __shfl_sync(0xffffffffu, lo ? d0 : d1, k & 15, 16);
      const R n2 = fmad(alpha, alpha, dk);
      const bool zero = n2 == R(0);  // zero column: tau = 0, nothing changes (guarded selects, the halves may differ)
      const R n2s = zero ? R(1) : n2;
      // Goldschmidt: g -> sqrt(n2), hh -> 1/(2 sqrt(n2)); the reciprocal of xi is seeded from the
      // approximate norm so that its MUFU latency overlaps these iterations
      const R y0 = Seed<R>::rsqrt0(n2s);
      R g = n2s * y0, hh = R(0.5) * y0;
      R r = Seed<R>::rcp0(alpha + copysign(g, alpha));
#pragma unroll
      for (int it = 0; it < Seed<R>::ITERS; ++it) {
        const R e = fmad(-g, hh, R(0.5));
        g = fmad(g, e, g);
        hh = fmad(hh, e, hh);
      }
      const R nu = copysign(g, alpha);
      const R inv_nu = copysign(hh + hh, alpha);
      const R xi = alpha + nu;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const R e = fmad(-xi, r, R(1));
        r = fmad(r, e, r);
      }
      const R tq = xi * inv_nu;  // tau = xi / nu
      // off the critical path: correctly rounded tau (a length-1 column must give exactly 2, as Julia's
      // division does) and one correction step for the stored norm
      const R nuc = copysign(fmad(fmad(-g, g, n2s), hh, g), alpha);   // corrected (correctly rounded) norm
      const R xic = alpha + nuc;
      const R tqc = xic * inv_nu;
      const R tk = fmad(fmad(-tqc, nuc, xic), inv_nu, tqc);
      const R mnu = -nuc;
      // slot 1: column c+16
      R nt1;
      {
        const bool right = (c + 16 > k) && !zero;
        const R s = fmad(d1, inv_nu, tq * a1[k]);
        nt1 = right ? -(s * r) : R(0);
        const bool mine = !lo && own && !zero;
        a1[k] = mine ? mnu : (right ? a1[k] - s : a1[k]);
        tau1 = mine ? tk : tau1;
        ixi1 = mine ? r : ixi1;
      }
      R nt0 = R(0);
      if (lo) {
        const bool right = (c > k) && !zero;
        const R s = fmad(d0, inv_nu, tq * a0[k]);
        nt0 = right ? -(s * r) : R(0);
        const bool mine = own && !zero;
        a0[k] = mine ? mnu : (right ? a0[k] - s : a0[k]);
        tau0 = mine ? tk : tau0;
        ixi0 = mine ? r : ixi0;
      }
      asm volatile("" ::: "memory");
      // ---- axpy sweep: a_ic -= (s/xi) a_ik
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
        R yy[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), yy);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int rr = i + j;
          if (rr > k) {
            if (lo) a0[rr] = fmad(nt0, yy[j], a0[rr]);
            a1[rr] = fmad(nt1, yy[j], a1[rr]);
          }
        }
      }
    }
    // deferred normalisation of the stored reflectors
#pragma unroll
    for (int i = 1; i < 32; ++i) {
      a0[i] = (i > c) ? a0[i] * ixi0 : a0[i];
      a1[i] = (i > c + 16) ? a1[i] * ixi1 : a1[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      *reinterpret_cast<VT*>(my + c * LD + i) = arr_to_vec(a0 + i);
      *reinterpret_cast<VT*>(my + (c + 16) * LD + i) = arr_to_vec(a1 + i);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NVEC; ++q) {
      const int p = lane + 32 * q;
      const int e = p * V;
      const int hm = e >> 10, col = (e & 1023) >> 5, row = e & 31;
      if (both || hm == 0) {
        const VT v = *reinterpret_cast<const VT*>(sm + hm * Cfg::TILE + col * LD + row);
        __stcs(reinterpret_cast<VT*>(Ag) + p, v);
      }
    }
    if (both || h == 0) {
      tau[(mat0 + h) * 32 + c] = tau0;
      tau[(mat0 + h) * 32 + 16 + c] = tau1;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------- 32x32, left-looking halves
// Same lane mapping as the half-warp kernel (half-warp h owns matrix h of the pair) but the matrix is
// processed LEFT-LOOKING in two column halves, so a lane holds ONE column (32 values) at a time:
//   phase 1  lane c owns column c:      16 reflector steps on the 32 x 16 left half; the final columns
//            (R above the diagonal, -nu on it, normalised v below) go back to the staging tile
//   phase 2  lane c owns column 16+c:   the 16 reflectors are applied from the tile (v_k broadcast with
//            LDS.128, tau_k by shuffle) -- pure FMA streams, no scalar chain, every lane useful
//   phase 3  16 reflector steps on the trailing 16 x 16 block of the right half
// Half the register payload of the half-warp kernel (no spills, 3 CTAs/SM) and the pivot column stays in
// registers between the dot and the axpy sweep of a step.
template <class R, int K0>
__device__ __forceinline__ void ll_factor_half(R (&a)[32], R* __restrict__ my, const int c, R& tau_own, R& ixi_own) {
  using VT = typename Vec16<R>::type;
  constexpr int V = Vec16<R>::N;
  constexpr int LD = HwCfg<R>::LD;
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
    // the owner publishes the current (un-normalised) pivot column to its own tile column
    if (own) {
#pragma unroll
      for (int i = k0; i < 32; i += V) *reinterpret_cast<VT*>(my + k * LD + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    const R* vk = my + k * LD;
    R y[32];
#pragma unroll
    for (int i = k0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y + i);
    const R alpha = y[k];
    R acc[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
    for (int r = k + 1; r < 32; ++r) acc[(r - k) & 3] = fmad(y[r], a[r], acc[(r - k) & 3]);
    const R d = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    const R dk = __shfl_sync(0xffffffffu, d, kk, 16);  // the owner's dot is the tail norm^2
    const R n2 = fmad(alpha, alpha, dk);
    const bool zero = n2 == R(0);  // zero column: tau = 0, nothing changes (guarded selects, the halves may differ)
    const R n2s = zero ? R(1) : n2;
    // Goldschmidt: g -> sqrt(n2), hh -> 1/(2 sqrt(n2)); the reciprocal of xi is seeded from the approximate
    // norm so that its MUFU latency overlaps these iterations
    const R y0 = Seed<R>::rsqrt0(n2s);
    R g = n2s * y0, hh = R(0.5) * y0;
    R r = Seed<R>::rcp0(alpha + copysign(g, alpha));
#pragma unroll
    for (int it = 0; it < Seed<R>::ITERS; ++it) {
      const R e = fmad(-g, hh, R(0.5));
      g = fmad(g, e, g);
      hh = fmad(hh, e, hh);
    }
    const R nu = copysign(g, alpha);
    const R inv_nu = copysign(hh + hh, alpha);
    const R xi = alpha + nu;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const R e = fmad(-xi, r, R(1));
      r = fmad(r, e, r);
    }
    const R tq = xi * inv_nu;  // tau = xi / nu
    // s = conj(tau) (a_kc + v^H a_c[k+1:]) with v = a_k / xi  ->  tau a_kc + d / nu
    const R s = fmad(d, inv_nu, tq * a[k]);
    const bool right = (c > kk) && !zero;
    const bool mine = own && !zero;
    const R nt = right ? -(s * r) : R(0);
    a[k] = mine ? -nu : (right ? a[k] - s : a[k]);
    tau_own = mine ? tq : tau_own;
    ixi_own = mine ? r : ixi_own;
#pragma unroll
    for (int rr = k + 1; rr < 32; ++rr) a[rr] = fmad(nt, y[rr], a[rr]);
  }
  // deferred normalisation of the stored reflector (rows below the diagonal of the own column)
#pragma unroll
  for (int i = K0 + 1; i < 32; ++i) a[i] = (i > K0 + c) ? a[i] * ixi_own : a[i];
}

template <class R, int WARPS, int MINB, int MODE = 0, bool PF = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    batched_qr32_ll_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = HwCfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVEC = 2 * 1024 / V / 32;  // 16-byte vectors per lane per matrix PAIR

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, c = lane & 15;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::PER_WARP;
  R* my = sm + h * Cfg::TILE;  // tile of this half-warp's matrix

  const i64 npairs = (batch + 1) >> 1;
  for (i64 pair = (i64)blockIdx.x * WARPS + warp; pair < npairs; pair += (i64)gridDim.x * WARPS) {
    const i64 mat0 = pair * 2;
    const bool both = mat0 + 1 < batch;
    R* Ag = A + mat0 * 1024;
    if (PF) {  // pull the pair this warp handles next into L2 while this one is being factorised
      const i64 nxt = pair + (i64)gridDim.x * WARPS;
      if (nxt * 2 + 1 < batch) {
        const char* pn = reinterpret_cast<const char*>(A + nxt * 2048);
#pragma unroll
        for (int q = 0; q < (int)(2048 * sizeof(R) / 128 / 32); ++q)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + (size_t)(lane + 32 * q) * 128));
      }
    }
    // ---- HBM -> staging tiles (two matrices = 2048 contiguous elements), 16-byte coalesced streaming loads
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      VT v[NVEC / 2];
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) {
        const int p = lane + 32 * (q + part * (NVEC / 2));
        const bool ok = both || (p * V < 1024);
        v[q] = ok ? __ldcs(reinterpret_cast<const VT*>(Ag) + p) : VT{};
      }
#pragma unroll
      for (int q = 0; q < NVEC / 2; ++q) {
        const int e = (lane + 32 * (q + part * (NVEC / 2))) * V;
        const int hm = e >> 10, col = (e & 1023) >> 5, row = e & 31;
        *reinterpret_cast<VT*>(sm + hm * Cfg::TILE + col * LD + row) = v[q];
      }
    }
    __syncwarp();
    R a[32];
    R tau_l = R(0), tau_r = R(0), ixi = R(1);
    // ---- phase 1: left half
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(my + c * LD + i), a + i);
    if (MODE != 1) ll_factor_half<R, 0>(a, my, c, tau_l, ixi);
#pragma unroll
    for (int i = 0; i < 32; i += V) *reinterpret_cast<VT*>(my + c * LD + i) = arr_to_vec(a + i);
    __syncwarp();
    // ---- phase 2: the 16 reflectors applied to the right half
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(my + (c + 16) * LD + i), a + i);
#pragma unroll
    for (int k = 0; k < (MODE == 1 ? 0 : 16); ++k) {
      const R tk = __shfl_sync(0xffffffffu, tau_l, k, 16);
      const R* vk = my + k * LD;
      R acc[4] = {a[k], R(0), R(0), R(0)};
      R y[32];
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y + i);
#pragma unroll
      for (int r = k + 1; r < 32; ++r) acc[(r - k) & 3] = fmad(y[r], a[r], acc[(r - k) & 3]);
      const R ns = -(tk * ((acc[0] + acc[1]) + (acc[2] + acc[3])));
      a[k] += ns;
#pragma unroll
      for (int r = k + 1; r < 32; ++r) a[r] = fmad(ns, y[r], a[r]);
    }
    // ---- phase 3: trailing 16 x 16 block of the right half
    ixi = R(1);
    if (MODE != 1) ll_factor_half<R, 16>(a, my, c, tau_r, ixi);
#pragma unroll
    for (int i = 0; i < 32; i += V) *reinterpret_cast<VT*>(my + (c + 16) * LD + i) = arr_to_vec(a + i);
    __syncwarp();
    // ---- staging tiles -> HBM (16-byte coalesced, streaming)
#pragma unroll
    for (int q = 0; q < NVEC; ++q) {
      const int p = lane + 32 * q;
      const int e = p * V;
      const int hm = e >> 10, col = (e & 1023) >> 5, row = e & 31;
      if (both || hm == 0) {
        const VT v = *reinterpret_cast<const VT*>(sm + hm * Cfg::TILE + col * LD + row);
        __stcs(reinterpret_cast<VT*>(Ag) + p, v);
      }
    }
    if (both || h == 0) {
      tau[(mat0 + h) * 32 + c] = tau_l;
      tau[(mat0 + h) * 32 + 16 + c] = tau_r;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------- 32x32, left-looking, half tile
// Same algorithm and lane mapping as batched_qr32_ll_kernel, but only HALF a tile of shared memory per matrix and
// the pivot column re-read in 16-byte chunks instead of being held in 64 registers: the shared-memory and register
// footprints per matrix in flight are what bound this latency-bound kernel (24 matrices per SM before).
//   S (16 padded columns) is, in turn: staging of the left half -> pivot buffer of phase 1 -> staging of the right
//   half -> the finished left half (read by phase 2, stored to HBM) -> pivot buffer of phase 3 -> staging of the
//   finished right half.
// NACC: accumulators of the dot sweep (2 is 1 % faster than 4 in the staggered kernel and saves two DADD per step; 1 is
// slower again).  Publishing the pivot column with predicated stores instead of the `if (own)` branch was 7 % SLOWER.
template <class R, int K0, int NACC = 4>
__device__ __forceinline__ void ll_factor_half_c(R (&a)[32], R* __restrict__ pub, const int c, R& tau_own, R& ixi_own) {
  using VT = typename Vec16<R>::type;
  constexpr int V = Vec16<R>::N;
  constexpr int LD = HwCfg<R>::LD;
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
    R* vk = pub + kk * LD;   // column slot kk of the half tile
    if (own) {
#pragma unroll
      for (int i = k0; i < 32; i += V) *reinterpret_cast<VT*>(vk + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    R alpha = R(0);
    R acc[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
    for (int i = k0; i < 32; i += V) {
      R y[V];
      vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int r = i + j;
        if (r == k) alpha = y[j];
        if (r > k) acc[(r - k) & (NACC - 1)] = fmad(y[j], a[r], acc[(r - k) & (NACC - 1)]);
      }
    }
    const R d = NACC == 4 ? (acc[0] + acc[1]) + (acc[2] + acc[3]) : NACC == 2 ? acc[0] + acc[1] : acc[0];
    const R dk = __shfl_sync(0xffffffffu, d, kk, 16);  // the owner's dot is the tail norm^2
    const R n2 = fmad(alpha, alpha, dk);
    const bool zero = n2 == R(0);
    const R n2s = zero ? R(1) : n2;
    const R y0 = Seed<R>::rsqrt0(n2s);
    R g = n2s * y0, hh = R(0.5) * y0;
    R r = Seed<R>::rcp0(alpha + copysign(g, alpha));
#pragma unroll
    for (int it = 0; it < Seed<R>::ITERS; ++it) {
      const R e = fmad(-g, hh, R(0.5));
      g = fmad(g, e, g);
      hh = fmad(hh, e, hh);
    }
    const R nu = copysign(g, alpha);
    const R inv_nu = copysign(hh + hh, alpha);
    const R xi = alpha + nu;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const R e = fmad(-xi, r, R(1));
      r = fmad(r, e, r);
    }
    const R tq = xi * inv_nu;  // tau = xi / nu
    const R s = fmad(d, inv_nu, tq * a[k]);
    const bool right = (c > kk) && !zero;
    const bool mine = own && !zero;
    const R nt = right ? -(s * r) : R(0);
    a[k] = mine ? -nu : (right ? a[k] - s : a[k]);
    tau_own = mine ? tq : tau_own;
    ixi_own = mine ? r : ixi_own;
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
#pragma unroll
  for (int i = K0 + 1; i < 32; ++i) a[i] = (i > K0 + c) ? a[i] * ixi_own : a[i];
}

template <class R>
struct Ll2Cfg {
  static constexpr int V = Vec16<R>::N;
  static constexpr int LD = HwCfg<R>::LD;
  static constexpr int HALF = 16 * LD + 16 / sizeof(R);   // +16 B: the two half tiles of a warp sit 4 banks apart
  static constexpr int PER_WARP = 2 * HALF;
};

// SYNC = 1: the warps of a CTA meet at a barrier before every pair, so that they walk the 82 KB straight-line body
// together and share instruction-cache lines (the body is 2.5x the 32 KB L1.5 instruction cache and `no_instruction`
// was 18 % of the stall cycles with free-running warps).  Measured, 2^20 matrices: free-running 3 CTAs x 4 warps 151.9,
// barrier 3 x 4 159.6, 2 x 6 159.2, 1 x 12 167.0 M matrices/s; further barriers before phase 2 and 3: 153-158 (slower).
template <class R, int WARPS, int SYNC>
__device__ __forceinline__ void batched_qr32_ll2_body(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = Ll2Cfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVH = 512 / V / 32;  // 16-byte vectors per lane per HALF matrix

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, c = lane & 15;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::PER_WARP;
  R* S = sm + h * Cfg::HALF;  // half tile of this half-warp's matrix

  // coalesced HBM <-> half tile of matrix m, columns [16*half, 16*half + 16)
  auto load_half = [&](const R* Ag, int m, int half, bool live) {
    VT v[NVH];
#pragma unroll
    for (int u = 0; u < NVH; ++u)
      v[u] = live ? __ldcs(reinterpret_cast<const VT*>(Ag + m * 1024 + half * 512) + lane + 32 * u) : VT{};
#pragma unroll
    for (int u = 0; u < NVH; ++u) {
      const int e = (lane + 32 * u) * V;
      *reinterpret_cast<VT*>(sm + m * Cfg::HALF + (e >> 5) * LD + (e & 31)) = v[u];
    }
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

  const i64 npairs = (batch + 1) >> 1;
  for (i64 base = (i64)blockIdx.x * WARPS; base < npairs; base += (i64)gridDim.x * WARPS) {
    // uniform trip count per CTA; a warp without a pair runs the body on zeros with every load / store masked
    const bool any = base + warp < npairs;
    if (SYNC == 0 && !any) break;
    if (SYNC) __syncthreads();
    const i64 pair = any ? base + warp : npairs - 1;
    const i64 mat0 = pair * 2;
    const bool both = any && (mat0 + 1 < batch);
    R* Ag = A + mat0 * 1024;
    {  // pull the pair this warp handles next into L2 while this one is being factorised
      const i64 nxt = pair + (i64)gridDim.x * WARPS;
      if (nxt * 2 + 1 < batch) {
        const char* pn = reinterpret_cast<const char*>(A + nxt * 2048);
#pragma unroll
        for (int u = 0; u < (int)(2048 * sizeof(R) / 128 / 32); ++u)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + (size_t)(lane + 32 * u) * 128));
      }
    }
    R tau_l = R(0), tau_r = R(0), ixi = R(1);
    R a[32];
    // ---- phase 1: left half
    load_half(Ag, 0, 0, any);
    load_half(Ag, 1, 0, both);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(S + c * LD + i), a + i);
    __syncwarp();
    ll_factor_half_c<R, 0>(a, S, c, tau_l, ixi);
    __syncwarp();
    // ---- transition: right half in (through S), finished left half out (through S, where phase 2 reads it)
    R b[32];
    load_half(Ag, 0, 1, any);
    load_half(Ag, 1, 1, both);
    __syncwarp();
    // slot c of S holds column 16+c (staged) and receives column c (finished): a lane only touches its own slot here,
    // so the swap goes chunk by chunk and a and b are never both live in full
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      vec_to_arr<R>(*reinterpret_cast<const VT*>(S + c * LD + i), b + i);
      *reinterpret_cast<VT*>(S + c * LD + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    store_half(Ag, 0, 0, any);
    store_half(Ag, 1, 0, both);
    // ---- phase 2: the 16 reflectors applied to the right half (v_k broadcast from S, tau_k by shuffle)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const R tk = __shfl_sync(0xffffffffu, tau_l, k, 16);
      const R* vk = S + k * LD;
      R acc[4] = {b[k], R(0), R(0), R(0)};
#pragma unroll
      for (int i = (k + 1) & ~(V - 1); i < 32; i += V) {
        R y[V];
        vec_to_arr<R>(*reinterpret_cast<const VT*>(vk + i), y);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (i + j > k) acc[(i + j - k) & 3] = fmad(y[j], b[i + j], acc[(i + j - k) & 3]);
      }
      const R ns = -(tk * ((acc[0] + acc[1]) + (acc[2] + acc[3])));
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
    __syncwarp();   // all lanes are done with the left half in S; phase 3 reuses S as its pivot buffer
    // ---- phase 3: trailing 16 x 16 block of the right half
    ixi = R(1);
    ll_factor_half_c<R, 16>(b, S, c, tau_r, ixi);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += V) *reinterpret_cast<VT*>(S + c * LD + i) = arr_to_vec(b + i);
    __syncwarp();
    store_half(Ag, 0, 1, any);
    store_half(Ag, 1, 1, both);
    if (both || (any && h == 0)) {
      tau[(mat0 + h) * 32 + c] = tau_l;
      tau[(mat0 + h) * 32 + 16 + c] = tau_r;
    }
    __syncwarp();
  }
}

template <class R, int WARPS, int MINB, int SYNC = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    batched_qr32_ll2_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  batched_qr32_ll2_body<R, WARPS, SYNC>(A, tau, batch);
}

template <class R, int MINB, int WARPS = 4, int SYNC = 0>
static int launch_ll2_32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  const size_t smem = (size_t)WARPS * Ll2Cfg<R>::PER_WARP * sizeof(R);
  auto kern = batched_qr32_ll2_kernel<R, WARPS, MINB, SYNC>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  const i64 npairs = (batch + 1) / 2;
  const i64 need = (npairs + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count() * occ;
  const i64 grid = need < resident ? need : resident;   // one resident wave; warps stride over the pairs
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------- 32x32, left-looking, cp.async prefetch
// batched_qr32_ll2_kernel with a second half tile P per matrix that receives the NEXT half by cp.async (LDGSTS, 16 bytes
// per lane, the same coalesced pattern as the staging loads, no registers): the right half while phase 1 runs, the left
// half of the warp's next pair while phases 2 and 3 run.  With the warps of a CTA in step (the barrier per pair), the
// exposed global-load latency at the top of a pair and at the transition was 5 % of the stall samples of the ll2 kernel
// (ncu source page: the staging STS behind the LDGs), and nobody was left to cover it.
// STAG: after the barrier warp w waits w * STAG cycles, so the twelve warps walk the body as a train 11 * STAG cycles
// long instead of in lock step -- short enough to keep sharing instruction-cache lines (at 6000 cycles between the
// warps of a scheduler the gain of the barrier is gone: 161 M/s), long enough that the warps of a scheduler are not all
// inside their scalar chains, or all bursting LDS / DFMA, at the same time.  Measured (2^20 matrices, M matrices/s):
// no stagger 168; warps of a scheduler 750 .. 3000 cycles apart, the four schedulers together: 178-180; every warp
// 200 / 300 / 375 / 450 / 550 cycles behind its predecessor: 181.5 / 187.5 / 185.6 / 181.7 / 178.2.
// A barrier only every second / fourth pair: 178 / 180 (the train drifts apart).
// Tried and dropped: interleaved lanes (matrix = lane & 1, column = lane >> 1, so that the two publishing lanes share a
// quarter-warp): 171 M/s -- a broadcast LDS.128 with two distinct addresses inside every quarter-warp costs more than
// one address per half-warp.
template <class R>
struct Ll4Cfg {
  static constexpr int V = Vec16<R>::N;
  static constexpr int LD = HwCfg<R>::LD;
  static constexpr int HALF = Ll2Cfg<R>::HALF;
  static constexpr int PER_WARP = 4 * HALF;   // S (work) and P (incoming) half tiles of both matrices of the pair
};

template <class R, int WARPS, int SYNC, int STAG, int NACC = 2, bool LATE = true>
__global__ void __launch_bounds__(WARPS * 32, 12 / WARPS)
    batched_qr32_ll4_kernel(R* __restrict__ A, R* __restrict__ tau, i64 batch) {
  using Cfg = Ll4Cfg<R>;
  using VT = typename Vec16<R>::type;
  constexpr int V = Cfg::V;
  constexpr int LD = Cfg::LD;
  constexpr int NVH = 512 / V / 32;  // 16-byte vectors per lane per HALF matrix

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, c = lane & 15;
  R* sm = reinterpret_cast<R*>(smem_raw) + warp * Cfg::PER_WARP;
  R* S = sm + h * Cfg::HALF;          // work half tile of this half-warp's matrix
  R* Pw = sm + 2 * Cfg::HALF;         // incoming half tiles of the warp's two matrices
  R* P = Pw + h * Cfg::HALF;

  const i64 npairs = (batch + 1) >> 1;
  const i64 stride = (i64)gridDim.x * WARPS;
  // columns [16*half, 16*half + 16) of both matrices of pair pr -> P, asynchronously (zero fill for a missing matrix)
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
    if (STAG > 0 && warp) {   // warp w leaves the barrier w * STAG cycles after warp 0 (see the note above the kernel)
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
    // ---- phase 1: left half (prefetched into P)
    wait_prefetch();
#pragma unroll
    for (int i = 0; i < 32; i += V) vec_to_arr<R>(*reinterpret_cast<const VT*>(P + c * LD + i), a + i);
    __syncwarp();              // every lane has its column: P may be refilled
    prefetch_half(pair, 1);    // right half -> P while phase 1 runs
    ll_factor_half_c<R, 0, NACC>(a, S, c, tau_l, ixi);
    __syncwarp();
    // ---- transition: right half from P, finished left half into S (phase 2 reads it there) and out to HBM
    R b[32];
    wait_prefetch();
#pragma unroll
    for (int i = 0; i < 32; i += V) {
      vec_to_arr<R>(*reinterpret_cast<const VT*>(P + c * LD + i), b + i);
      *reinterpret_cast<VT*>(S + c * LD + i) = arr_to_vec(a + i);
    }
    __syncwarp();
    if (pair + stride < npairs) prefetch_half(pair + stride, 0);   // left half of the warp's next pair -> P
    if (!LATE) {
      store_half(Ag, 0, 0, true);
      store_half(Ag, 1, 0, both);
    }
    // ---- phase 2: the 16 reflectors applied to the right half (v_k broadcast from S, tau_k by shuffle)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const R tk = __shfl_sync(0xffffffffu, tau_l, k, 16);
      const R* vk = S + k * LD;
      R acc[4] = {b[k], R(0), R(0), R(0)};
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
    if (LATE) {     // the finished left half leaves for HBM after phase 2 (S is read-only until here): spreads the LSU burst
                    // of the transition (swap + prefetch issue + store): 189.6 -> 194.1 M/s; slice by slice inside phase 2: 193.4
                    // (issuing the two prefetches a few reflector steps INSIDE phases 1 and 2 instead: 190.8, slower)
      store_half(Ag, 0, 0, true);
      store_half(Ag, 1, 0, both);
    }
    __syncwarp();   // all lanes are done with the left half in S; phase 3 reuses S as its pivot buffer
    // ---- phase 3: trailing 16 x 16 block of the right half
    ixi = R(1);
    ll_factor_half_c<R, 16, NACC>(b, S, c, tau_r, ixi);
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

template <class R, int WARPS, int SYNC, int STAG, int NACC = 2, bool LATE = true>
static int launch_ll4_32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  const size_t smem = (size_t)WARPS * Ll4Cfg<R>::PER_WARP * sizeof(R);
  auto kern = batched_qr32_ll4_kernel<R, WARPS, SYNC, STAG, NACC, LATE>;
  // attribute + occupancy query once per device and instantiation (the host-pointer pipeline launches per chunk)
  static thread_local int cached_dev = -1, cached_occ = 0;
  int dev = 0;
  GLA_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int o = 0;
    GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, WARPS * 32, smem));
    cached_occ = o < 1 ? 1 : o;
    cached_dev = dev;
  }
  const int occ = cached_occ;
  const i64 npairs = (batch + 1) / 2;
  const i64 need = (npairs + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count() * occ;
  const i64 grid = need < resident ? need : resident;   // one resident wave; warps stride over the pairs
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template <class R>
static int launch_ll32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  constexpr int WARPS = 4;
  constexpr int MINB = sizeof(R) == 8 ? 3 : 4;
  const size_t smem = (size_t)WARPS * HwCfg<R>::PER_WARP * sizeof(R);
  static const int mode = [] {
    const char* e = getenv("GLA_BATCHED_MODE");
    return e ? atoi(e) : 0;
  }();
  // mode 1 = copy only (memory path ceiling of this structure, for profiling); default = L2 prefetch of the next pair
  auto kern = mode == 1 ? batched_qr32_ll_kernel<R, WARPS, MINB, 1, false>
              : mode == 2 ? batched_qr32_ll_kernel<R, WARPS, MINB, 0, false>
                          : batched_qr32_ll_kernel<R, WARPS, MINB, 0, true>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  const i64 npairs = (batch + 1) / 2;
  const i64 need = (npairs + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count() * occ;
  const i64 grid = need < resident ? need : resident;   // one resident wave; warps stride over the pairs
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template <class R>
static int launch_hw32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  constexpr int WARPS = 4;
  constexpr int MINB = sizeof(R) == 8 ? 2 : 4;
  const size_t smem = (size_t)WARPS * HwCfg<R>::PER_WARP * sizeof(R);
  auto kern = batched_qr32_hw_kernel<R, WARPS, MINB>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  const i64 npairs = (batch + 1) / 2;
  const i64 need = (npairs + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count() * occ;
  const i64 grid = need < resident ? need : resident;   // one resident wave; warps stride over the pairs
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
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
static int launch_hw32(R* dA, R* dtau, i64 batch, cudaStream_t st);

template <class R, int WARPS, int MAXNREG>
static int launch_reg32_cfg(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  const size_t smem = (size_t)WARPS * Reg32Cfg<R>::TILE * sizeof(R);
  auto kern = batched_qr32_reg_kernel<R, WARPS, MAXNREG>;
  GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) occ = 1;
  const i64 need = (batch + WARPS - 1) / WARPS;
  const i64 resident = (i64)sm_count() * occ;
  // one resident wave of CTAs; each warp strides over its share of the batch
  const i64 grid = need < resident ? need : resident;
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(dA, dtau, batch);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

template <class R>
static int launch_reg32(R* dA, R* dtau, i64 batch, cudaStream_t st) {
  // GLA_BATCHED_VARIANT selects the earlier kernels for A/B measurements (default = measured best)
  static const int variant = [] {
    const char* e = getenv("GLA_BATCHED_VARIANT");
    return e ? atoi(e) : 0;
  }();
  if (variant == 1) return launch_reg32_cfg<R, 4, 168>(dA, dtau, batch, st);
  if (variant == 2) return launch_reg32_cfg<R, 4, 200>(dA, dtau, batch, st);
  if (variant == 3) return launch_hw32<R>(dA, dtau, batch, st);
  if (variant == 4) return launch_ll2_32<R, 4>(dA, dtau, batch, st);   // half-tile kernel, <= 128 registers: 127 M/s (spills)
  if (variant == 7) return launch_ll32<R>(dA, dtau, batch, st);        // full-tile left-looking kernel: 141 M/s
  if (variant == 8) return launch_ll2_32<R, 3, 4, 0>(dA, dtau, batch, st);    // half tile, free-running 3 CTAs x 4 warps: 152 M/s
  if (variant == 9) return launch_ll2_32<R, 1, 12, 1>(dA, dtau, batch, st);   // half tile, 1 CTA x 12 warps, barrier per pair: 167 M/s
  if (variant == 10) return launch_ll2_32<R, 3, 4, 1>(dA, dtau, batch, st);   // half tile, 3 CTAs x 4 warps, barrier per pair: 160 M/s
  if (variant == 11) return launch_ll4_32<R, 12, 0, 0>(dA, dtau, batch, st);  // cp.async prefetch, free-running: 172 M/s
  if (variant == 12) return launch_ll4_32<R, 12, 1, 0>(dA, dtau, batch, st);  // cp.async prefetch, barrier, lock step: 168 M/s
  if (variant == 14) return launch_ll4_32<R, 12, 1, 300, 2, false>(dA, dtau, batch, st);   // left half stored at the transition: 189.6 M/s
  if (variant == 13) return launch_ll4_32<R, 12, 1, 300, 4, false>(dA, dtau, batch, st);   // four dot accumulators, early store: 187.5 M/s
  // default: half-tile left-looking kernel with cp.async prefetch of the next half, ONE CTA of 12 warps per SM (<= 168
  // registers) meeting at a barrier before every pair and leaving it 300 cycles apart, two dot accumulators, left half stored after phase 2: 194 M matrices/s
  return launch_ll4_32<R, 12, 1, 300>(dA, dtau, batch, st);
}

template <class T>
int geqr_batched_dev(T* dA, i64 m, i64 n, i64 batch, T* dtau, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (batch < 0) return -4;
  if (m == 0 || n == 0 || batch == 0) return 0;
  if constexpr (IsReal<T>::v) {
    // the register kernels move 16-byte vectors (cp.async / LDG.128 / STG.128); a matrix stack that is not 16-byte aligned
    // (an offset view) goes through the generic shared-memory kernel below, which only uses element accesses
    if (m == 32 && n == 32 && (reinterpret_cast<uintptr_t>(dA) & 15) == 0) return launch_reg32<T>(dA, dtau, batch, st);
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
