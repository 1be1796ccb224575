// gemm.cuh -- the one contraction every trailing update is built from:
//
//     C(i,j) = beta*C(i,j) + alpha * sum_k op(At(k,i)) * B(k,j)          ("TN": K is the contiguous
//                                                                          dimension of both operands)
//   At : K x M column-major (ldat)   B : K x N column-major (ldb)   C : M x N column-major (ldc)
//   op = conj when conj_a (ComplexF64 V^H A), identity otherwise.
//
// Float64 runs on the FP64 tensor pipe (mma.sync m8n8k4 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind)
// fed by TMA (cp.async.bulk.tensor, 128B-swizzled 16-double K slabs, mbarrier full/empty ring, one
// producer warp).  Float32 / ComplexF64, and Float64 operands that violate TMA's 16-byte alignment
// rules, run on a shared-memory tiled FMA kernel with the same interface.
//
// split-K: with nsplit > 1 the K range is cut into nsplit slices and slice z writes its own partial
// product to C + z*split_stride (beta forced to 0); the consumer sums the slices in a fixed order
// (deterministic, no floating-point atomics).
#pragma once
#include "common.cuh"

namespace gla {

template <class T>
struct GemmTN {
  const T* At = nullptr;
  i64 ldat = 0;
  const T* B = nullptr;
  i64 ldb = 0;
  T* C = nullptr;
  i64 ldc = 0;
  i64 M = 0, N = 0, K = 0;
  typename Sc<T>::real alpha = 1;
  int beta_one = 0;     // 0: C = alpha*acc, 1: C += alpha*acc
  int conj_a = 0;       // complex only
  int lower_only = 0;   // triangle mask: 1 = write only i >= j (lower), 2 = only i <= j (upper)
  int nsplit = 1;       // split-K slices (beta_one must be 0 when > 1)
  i64 split_stride = 0; // elements between partial outputs
  int yield_sms = 0;    // 1: a high-priority stream needs SMs while this product runs (look-ahead schedules): persistent kernels
                        //    bound the lifetime of their CTAs instead of holding every SM until the product is done
};

template <class T>
int gemm_tn(const GemmTN<T>& g, cudaStream_t st);

// out(M x N, ldo) = sum_z part[z]  (fixed order); part z at part + z*stride, ld = ldp
template <class T>
int sum_splits(T* out, i64 ldo, const T* part, i64 ldp, i64 stride, int nsplit, i64 M, i64 N, cudaStream_t st);

// pick a split count so that tiles*nsplit fills the GPU; K slices stay multiples of 16 and >= 64 (GLA_GEMM_MINK)
int choose_nsplit(i64 M, i64 N, i64 K, int bm, int bn);
// the same for the persistent tcgen05 Float32 kernel (128 x 128 tiles, one CTA per SM), at most max_split slices
int choose_nsplit_persistent(i64 M, i64 N, i64 K, int max_split);

}  // namespace gla
