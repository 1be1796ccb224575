// smallqr.cuh -- CTA-level Householder QR of a matrix resident in shared memory.
// Used by the generic batched kernel, by the TSQR leaves/tree nodes and by the blocked driver
// for panels that fit one CTA.
//
// Per column k (reference qrUnblocked!, src/qr.jl:86-111; stdlib reflector!/reflectorApply!
// call sites src/qr.jl:96,102):
//   every warp recomputes ||A[k:,k]||^2 with a shuffle reduction (no block barrier for the norm),
//   warp w then owns columns k+1+w, k+1+w+W, ...: dot with the UN-normalised pivot column,
//   s = conj(tau)*(a_kc + conj(1/xi)*d), a_kc -= s, a_ic -= a_ik*(s/xi); the pivot column is
//   scaled by 1/xi one step later (after the single __syncthreads of the step) by all threads.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace gla {

constexpr int SMALLQR_THREADS = 256;

template <class R>
struct SafeRange;
template <>
struct SafeRange<double> {
  __host__ __device__ static double lo() { return 1e-280; }
  __host__ __device__ static double hi() { return 1e280; }
  __host__ __device__ static double fmax() { return 1.7976931348623157e308; }
};
template <>
struct SafeRange<float> {
  __host__ __device__ static float lo() { return 1e-30f; }
  __host__ __device__ static float hi() { return 1e30f; }
  __host__ __device__ static float fmax() { return 3.402823466e38f; }
};
__device__ __forceinline__ float absmax_part(float a) { return fabsf(a); }
__device__ __forceinline__ double absmax_part(double a) { return fabs(a); }
__device__ __forceinline__ double absmax_part(zd a) { return fmax(fabs(a.x), fabs(a.y)); }

template <class T>
struct ReflScalars {
  typename Sc<T>::real nu;  // copysign(norm, re(alpha))
  T tau;                    // xi / nu
  T ixi;                    // 1 / xi
  bool nonzero;
};

// scalars of LinearAlgebra.reflector! from alpha = x[1] and n2 = ||x||^2
template <class T>
__device__ __forceinline__ ReflScalars<T> reflector_scalars(T alpha, typename Sc<T>::real n2) {
  using R = typename Sc<T>::real;
  ReflScalars<T> r;
  r.nonzero = n2 != R(0);
  if (!r.nonzero) {
    r.nu = R(0);
    r.tau = Sc<T>::zero();
    r.ixi = Sc<T>::one();
    return r;
  }
  R nrm = sqrt(n2);
  r.nu = copysign(nrm, re(alpha));
  T xi = alpha + Sc<T>::from_real(r.nu);
  r.tau = scale_real(xi, R(1) / r.nu);
  r.ixi = inv(xi);
  return r;
}

// The same scalars from the MUFU seeds + FMA refinement of fastmath.cuh (no IEEE sqrt / division subroutine on the
// per-column critical path of the panel kernels): sqrt correctly rounded in almost all cases, reciprocals to <= 1 ulp.
template <class R>
__device__ __forceinline__ ReflScalars<R> reflector_scalars_fast(R alpha, R n2) {
  ReflScalars<R> r;
  r.nonzero = n2 != R(0);
  if (!r.nonzero || !(n2 >= SafeRange<R>::lo() && n2 <= SafeRange<R>::hi())) return reflector_scalars<R>(alpha, n2);
  const R rs = Fast<R>::rsqrt(n2);
  const R nrm = Fast<R>::sqrt_from_rsqrt(n2, rs);
  r.nu = copysign(nrm, alpha);
  const R xi = alpha + r.nu;
  r.ixi = Fast<R>::rcp(xi);
  r.tau = xi * copysign(rs, alpha);   // xi / nu
  // one Newton correction of tau against nu (rs is 1/sqrt(n2) to ~1 ulp, nu is the rounded sqrt)
  r.tau = fma(fma(-r.tau, r.nu, xi), copysign(rs, alpha), r.tau);
  return r;
}

// A: shared memory, column-major m x n, leading dimension ld.  tau: global (may be nullptr).
// All threads of the CTA must call; blockDim.x must be a multiple of 32.
template <class T>
__device__ void cta_qr_smem(T* sA, int m, int n, int ld, T* tau) {
  using R = typename Sc<T>::real;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int kmax = m < n ? m : n;
  T prev_ixi = Sc<T>::one();
  R prev_nu = R(0);
  bool prev_nonzero = false;
  for (int k = 0; k <= kmax; ++k) {
    __syncthreads();  // the single barrier of the step: updates of step k-1 are visible
    // deferred finish of pivot column k-1 (nobody reads it any more): diagonal <- -nu, rows below *= 1/xi
    if (prev_nonzero) {
      T* col = sA + (k - 1) * ld;
      for (int i = k + threadIdx.x; i < m; i += blockDim.x) col[i] = col[i] * prev_ixi;
      if (threadIdx.x == 0) col[k - 1] = Sc<T>::from_real(-prev_nu);
    }
    if (k == kmax) break;
    const T* ck = sA + k * ld;
    R part = R(0);
    for (int i = k + 1 + lane; i < m; i += 32) part += abs2(ck[i]);
    part = warp_sum(part);
    const T alpha = ck[k];
    R n2 = abs2(alpha) + part;
    // Julia's reflector! uses the scaled norm(x): when the plain sum of squares leaves the safe range (columns around
    // 1e-160 / 1e160, 1e-20 / 1e20 in Float32) redo it on x * 2^-e and carry the power of two separately (warp-uniform branch)
    R up = R(1);   // nu, xi are those of the scaled column times `up`
    if (!(n2 >= SafeRange<R>::lo() && n2 <= SafeRange<R>::hi())) {
      R amax = absmax_part(alpha);
      for (int i = k + 1 + lane; i < m; i += 32) amax = fmax(amax, absmax_part(ck[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (amax > R(0) && amax <= SafeRange<R>::fmax()) {
        int e;
        (void)frexp(amax, &e);
        const R sc = ldexp(R(1), -e);
        up = ldexp(R(1), e);
        R p2 = R(0);
        for (int i = k + 1 + lane; i < m; i += 32) p2 += abs2(scale_real(ck[i], sc));
        p2 = warp_sum(p2);
        n2 = abs2(scale_real(alpha, sc)) + p2;
      }
    }
    ReflScalars<T> rs = reflector_scalars<T>(up == R(1) ? alpha : scale_real(alpha, R(1) / up), n2);
    if (up != R(1)) {   // back to the unscaled column: nu * 2^e, 1/xi * 2^-e; tau is scale free
      rs.nu *= up;
      rs.ixi = scale_real(rs.ixi, R(1) / up);
    }
    prev_nonzero = rs.nonzero;
    prev_ixi = rs.ixi;
    prev_nu = rs.nu;
    if (tau && threadIdx.x == 0) tau[k] = rs.tau;
    if (rs.nonzero) {
      const T ctau = cj(rs.tau);
      // rescaled column: the dot runs on x * 2^-e (no overflow of x_i * c_i) against 1/xi of the scaled column
      const R isup = R(1) / up;
      const T cixi = cj(up == R(1) ? rs.ixi : scale_real(rs.ixi, up));
      for (int c = k + 1 + warp; c < n; c += nwarps) {
        T* cc = sA + c * ld;
        T d = Sc<T>::zero();
        if (up == R(1)) {
          for (int i = k + 1 + lane; i < m; i += 32) d = fmad(cj(ck[i]), cc[i], d);
        } else {
          for (int i = k + 1 + lane; i < m; i += 32) d = fmad(cj(scale_real(ck[i], isup)), cc[i], d);
        }
        d = warp_sum(d);
        const T s = ctau * (cc[k] + cixi * d);
        const T t = s * rs.ixi;
        for (int i = k + 1 + lane; i < m; i += 32) cc[i] = cc[i] - ck[i] * t;
        __syncwarp();  // every lane has read cc[k] before lane 0 overwrites it
        if (lane == 0) cc[k] = cc[k] - s;
      }
    }
  }
  __syncthreads();
}

// The same factorisation by ONE WARP on a matrix in (warp-private) shared memory: lane = column, so the dots of a step
// run over the rows inside each lane (no shuffle reduction per column, no block barrier); only the pivot column's norm is
// a warp reduction.  For the small batched shapes (n <= 32 columns per pass, any element type): eight independent
// matrices per CTA instead of one CTA per matrix.  Same arithmetic order per entry as cta_qr_smem except for the dots
// (serial over the rows here, a butterfly over 32 partial sums there).
template <class T>
__device__ void warp_qr_smem(T* sA, int m, int n, int ld, T* tau) {
  using R = typename Sc<T>::real;
  const int lane = threadIdx.x & 31;
  const int kmax = m < n ? m : n;
  for (int k = 0; k < kmax; ++k) {
    T* ck = sA + k * ld;
    R part = R(0);
    for (int i = k + 1 + lane; i < m; i += 32) part += abs2(ck[i]);
    part = warp_sum(part);
    const T alpha = ck[k];
    R n2 = abs2(alpha) + part;
    R up = R(1);
    if (!(n2 >= SafeRange<R>::lo() && n2 <= SafeRange<R>::hi())) {   // scaled norm, as in cta_qr_smem
      R amax = absmax_part(alpha);
      for (int i = k + 1 + lane; i < m; i += 32) amax = fmax(amax, absmax_part(ck[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (amax > R(0) && amax <= SafeRange<R>::fmax()) {
        int e;
        (void)frexp(amax, &e);
        const R sc = ldexp(R(1), -e);
        up = ldexp(R(1), e);
        R p2 = R(0);
        for (int i = k + 1 + lane; i < m; i += 32) p2 += abs2(scale_real(ck[i], sc));
        p2 = warp_sum(p2);
        n2 = abs2(scale_real(alpha, sc)) + p2;
      }
    }
    ReflScalars<T> rs = reflector_scalars<T>(up == R(1) ? alpha : scale_real(alpha, R(1) / up), n2);
    if (up != R(1)) {
      rs.nu *= up;
      rs.ixi = scale_real(rs.ixi, R(1) / up);
    }
    if (tau && lane == 0) tau[k] = rs.tau;
    if (rs.nonzero) {
      const T ctau = cj(rs.tau);
      const R isup = R(1) / up;
      const T cixi = cj(up == R(1) ? rs.ixi : scale_real(rs.ixi, up));
      for (int c = k + 1 + lane; c < n; c += 32) {
        T* cc = sA + c * ld;
        T d0 = Sc<T>::zero(), d1 = Sc<T>::zero();
        int i = k + 1;
        if (up == R(1)) {
          for (; i + 1 < m; i += 2) {
            d0 = fmad(cj(ck[i]), cc[i], d0);
            d1 = fmad(cj(ck[i + 1]), cc[i + 1], d1);
          }
          if (i < m) d0 = fmad(cj(ck[i]), cc[i], d0);
        } else {
          for (; i < m; ++i) d0 = fmad(cj(scale_real(ck[i], isup)), cc[i], d0);
        }
        const T s = ctau * (cc[k] + cixi * (d0 + d1));
        const T t = s * rs.ixi;
        for (i = k + 1; i < m; ++i) cc[i] = cc[i] - ck[i] * t;
        cc[k] = cc[k] - s;
      }
      __syncwarp();   // every lane has used the un-normalised pivot column
      for (int i = k + 1 + lane; i < m; i += 32) ck[i] = ck[i] * rs.ixi;
      if (lane == 0) ck[k] = Sc<T>::from_real(-rs.nu);
    }
    __syncwarp();
  }
}

}  // namespace gla
