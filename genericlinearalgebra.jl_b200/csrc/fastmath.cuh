// fastmath.cuh -- FP64/FP32 reciprocal, rsqrt and sqrt built from the MUFU seeds plus FMA refinement
// (no IEEE division / sqrt subroutine calls on the latency-critical scalar chains of the reflector kernels).
#pragma once
#include "common.cuh"

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

}  // namespace gla
