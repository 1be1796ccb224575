// tsqr.cu -- K5: tall-skinny QR, R factor only, as a communication-avoiding reduction.
//
// Semantics: the R that qrBlocked! (reference src/qr.jl:113-146) leaves in the upper triangle of a
// tall m x n matrix, n <= 64, up to the row signs discussed in DESIGN.md ("TSQR sign"): every
// reduction node is a Householder QR with the reference's reflector! convention (stdlib reflector!,
// call site src/qr.jl:96: nu = copysign(norm, pivot), diagonal <- -nu, tau = (pivot+nu)/nu; left
// reflectorApply!, call site src/qr.jl:102), so each node's diagonal is -copysign(norm, pivot); the
// signs of the final R are those of the LAST node, which need not equal the signs sequential
// Householder on the whole matrix would give.
//
// tsqr_stream_kernel: TWO PERSISTENT CTAs PER SM, each streams over its contiguous row range in chunks of
// 32*NWARPS rows and folds every chunk into a running n x n R held in shared memory:
//     [ R ]          [ R' ]
//     [ C ]  = Q  *  [ 0  ]        (R upper triangular, C the dense chunk)
// Column step k only touches row k of R and the dense rows, so the flop count is the minimal 2 m n^2.
//   * warp w keeps rows 32w..32w+31 of the chunk IN REGISTERS, lane c owning columns c and c+32
//     (64 values x 2); the pivot column is published once per step through shared memory and read
//     back with broadcast LDS.128;
//   * one __syncthreads per step: every warp writes its 64 partial dots, all warps sum the NWARPS
//     partials in the same fixed order (bitwise identical scalars in every warp, deterministic);
//   * the next chunk is prefetched with cp.async into the (then free) staging tile while the current
//     one is being reduced; out-of-range rows / columns are zero-filled (zero rows do not change R);
//   * sqrt / reciprocal from MUFU seeds + FMA refinement (fastmath.cuh).
// The per-CTA R factors are stacked into a tall matrix and reduced by the same kernel until one is left.
// Multi-GPU: each rank reduces its row block (gla_dtsqr_local_dev), the ranks exchange their n x n R
// factors with ONE ncclAllGather (32 KiB each, NVLink) and every rank reduces the stack
// (gla_dtsqr_combine_dev); gla_dtsqr_allreduce_dev does exchange + reduction on a caller-owned communicator.
#include "fastmath.cuh"
#include "gla_internal.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <mutex>

namespace gla {

constexpr int TSQR_MAXN = 64;
constexpr int GLA_ERR_NCCL_CODE = 2000;  // GLA_ERR_NCCL in include/gla_cuda.h

// element (row, col) of the source: p + (row / blk_rows) * bstride + (row % blk_rows) + col * ld
struct TsqrSrc {
  const double* p;
  i64 ld, blk_rows, bstride;
};

// NWARPS warps, each holding RW rows of the chunk in registers
template <int NWARPS, int RW>
struct TsCfg {
  static constexpr int ROWS = RW * NWARPS;
  static constexpr int SMEM_DOUBLES = 64 * 64 + 2 * NWARPS * 64 + NWARPS * RW + NWARPS * 64 * RW;
  static constexpr size_t SMEM = (size_t)SMEM_DOUBLES * sizeof(double);
};

// staging tile of one warp: 64 columns x RW rows, column stride RW doubles (no padding); the 16-byte chunk q of
// column c sits at chunk q ^ (c & 7), so the LDS.128 of 8 consecutive lanes (one column each) hit 8 distinct banks
template <int RW>
__device__ __forceinline__ int ts_tile_idx(int c, int r) {
  return c * RW + ((((r >> 1) ^ (c & 7)) << 1) | (r & 1));
}

__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 8 : 0;  // src-size 0: the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// one column step.  H = 0: pivot in the lane's first column (k = kl), H = 1: in the second (k = 32 + kl)
template <int NWARPS, int RW, bool TWO, int H>
__device__ __forceinline__ void ts_step(double (&a0)[RW], double (&a1)[RW], const int kl, const int lane, const int warp,
                                        double* __restrict__ sR, double* __restrict__ part, double* __restrict__ pubw) {
  const int k = 32 * H + kl;
  if (lane == kl) {
#pragma unroll
    for (int r = 0; r < RW; r += 2)
      *reinterpret_cast<double2*>(pubw + r) = H == 0 ? make_double2(a0[r], a0[r + 1]) : make_double2(a1[r], a1[r + 1]);
  }
  __syncwarp();
  // row k of R: read BEFORE the step's barrier, warp 0 rewrites it after the barrier
  const double top0 = (H == 0) ? sR[k * 64 + lane] : 0.0;
  const double top1 = TWO ? sR[k * 64 + 32 + lane] : 0.0;
  double p0[4] = {0., 0., 0., 0.}, p1[4] = {0., 0., 0., 0.};
#pragma unroll
  for (int r = 0; r < RW; r += 4) {
    const double2 xa = *reinterpret_cast<const double2*>(pubw + r);
    const double2 xb = *reinterpret_cast<const double2*>(pubw + r + 2);
    if (H == 0) {
      p0[0] = fma(xa.x, a0[r], p0[0]);
      p0[1] = fma(xa.y, a0[r + 1], p0[1]);
      p0[2] = fma(xb.x, a0[r + 2], p0[2]);
      p0[3] = fma(xb.y, a0[r + 3], p0[3]);
    }
    if (TWO) {
      p1[0] = fma(xa.x, a1[r], p1[0]);
      p1[1] = fma(xa.y, a1[r + 1], p1[1]);
      p1[2] = fma(xb.x, a1[r + 2], p1[2]);
      p1[3] = fma(xb.y, a1[r + 3], p1[3]);
    }
  }
  double* pp = part + ((k & 1) * NWARPS + warp) * 64;
  if (H == 0) pp[lane] = (p0[0] + p0[1]) + (p0[2] + p0[3]);
  if (TWO) pp[32 + lane] = (p1[0] + p1[1]) + (p1[2] + p1[3]);
  __syncthreads();
  const double* pr = part + (k & 1) * NWARPS * 64;
  double d0 = 0., d1 = 0.;
#pragma unroll
  for (int w = 0; w < NWARPS; ++w) {
    if (H == 0) d0 += pr[w * 64 + lane];
    if (TWO) d1 += pr[w * 64 + 32 + lane];
  }
  const double dk = __shfl_sync(0xffffffffu, H == 0 ? d0 : d1, kl);
  const double alpha = __shfl_sync(0xffffffffu, H == 0 ? top0 : top1, kl);
  const double n2 = fma(alpha, alpha, dk);
  if (n2 != 0.0) {  // uniform over the CTA (identical scalars everywhere); zero column: tau = 0, nothing changes
    double nu, ixi, tau;
    if (n2 > 1e-280 && n2 < 1e280) {
      // Goldschmidt: g -> sqrt(n2), hh -> 1/(2 sqrt(n2)); the reciprocal of xi is seeded from the approximate
      // norm so that its MUFU latency overlaps these iterations (same scalar chain as the batched kernel)
      double y0, r;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(n2));
      double g = n2 * y0, hh = 0.5 * y0;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(alpha + copysign(g, alpha)));
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const double e = fma(-g, hh, 0.5);
        g = fma(g, e, g);
        hh = fma(hh, e, hh);
      }
      g = fma(fma(-g, g, n2), hh, g);  // last correction of the norm
      nu = copysign(g, alpha);
      const double xi = alpha + nu;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const double e = fma(-xi, r, 1.0);
        r = fma(r, e, r);
      }
      ixi = r;
      tau = xi * copysign(hh + hh, alpha);  // xi / nu
    } else {  // tiny / huge column (e.g. rounding residue of a rank-deficient block): IEEE sqrt and division
      nu = copysign(sqrt(n2), alpha);
      const double xi = alpha + nu;
      ixi = 1.0 / xi;
      tau = xi / nu;
    }
    // s = tau (R_kc + v^H a_c) with v = [1; x/xi]; R_kc -= s; dense part -= (s/xi) x
    double t0 = 0., t1 = 0., nt0 = top0, nt1 = top1;
    if (H == 0) {
      const bool right = lane > kl;
      const double s = tau * fma(d0, ixi, top0);
      t0 = right ? -(s * ixi) : 0.;
      nt0 = right ? top0 - s : (lane == kl ? -nu : top0);
    }
    if (TWO) {
      const bool right = (H == 0) || lane > kl;
      const double s = tau * fma(d1, ixi, top1);
      t1 = right ? -(s * ixi) : 0.;
      nt1 = right ? top1 - s : ((H == 1 && lane == kl) ? -nu : top1);
    }
    if (warp == 0) {
      if (H == 0) sR[k * 64 + lane] = nt0;
      if (TWO) sR[k * 64 + 32 + lane] = nt1;
    }
#pragma unroll
    for (int r = 0; r < RW; r += 4) {
      const double2 xa = *reinterpret_cast<const double2*>(pubw + r);
      const double2 xb = *reinterpret_cast<const double2*>(pubw + r + 2);
      if (H == 0) {
        a0[r] = fma(t0, xa.x, a0[r]);
        a0[r + 1] = fma(t0, xa.y, a0[r + 1]);
        a0[r + 2] = fma(t0, xb.x, a0[r + 2]);
        a0[r + 3] = fma(t0, xb.y, a0[r + 3]);
      }
      if (TWO) {
        a1[r] = fma(t1, xa.x, a1[r]);
        a1[r + 1] = fma(t1, xa.y, a1[r + 1]);
        a1[r + 2] = fma(t1, xb.x, a1[r + 2]);
        a1[r + 3] = fma(t1, xb.y, a1[r + 3]);
      }
    }
  }
  __syncwarp();  // every lane is done with the published column before the next owner overwrites it
}

// CTA b reduces rows [b*rows_per_cta, min(m, (b+1)*rows_per_cta)) to an n x n R, written (upper, zeros below) to
// Rout + b*out_row_step with leading dimension ldro.
template <int NWARPS, int RW, bool TWO>
__global__ void __launch_bounds__(NWARPS * 32, 2)
    tsqr_stream_kernel(const TsqrSrc src, const i64 m, const int n, const i64 rows_per_cta, double* __restrict__ Rout,
                       const i64 ldro, const i64 out_row_step) {
  using Cfg = TsCfg<NWARPS, RW>;
  extern __shared__ __align__(16) double ts_smem[];
  double* sR = ts_smem;                       // [64][64]   R(k, c) at k*64 + c
  double* part = sR + 64 * 64;                // [2][NWARPS][64] partial dots, double-buffered by step parity
  double* pub = part + 2 * NWARPS * 64;       // [NWARPS][RW] published pivot column
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tile = pub + NWARPS * RW + warp * 64 * RW;  // staging tile of this warp (ts_tile_idx)
  double* pubw = pub + warp * RW;

  for (int e = threadIdx.x; e < 64 * 64; e += NWARPS * 32) sR[e] = 0.0;
  const i64 row_begin = (i64)blockIdx.x * rows_per_cta;
  i64 row_end = row_begin + rows_per_cta;
  if (row_end > m) row_end = m;
  const i64 nchunks = (row_end - row_begin + Cfg::ROWS - 1) / Cfg::ROWS;
  const int ncols = TWO ? 64 : 32;
  constexpr int CPP = 32 / RW;                // columns per cp.async pass: RW lanes walk the rows of one column
  const int lr = lane % RW, lc = lane / RW;

  auto prefetch = [&](i64 chunk) {
    const i64 row = row_begin + chunk * Cfg::ROWS + RW * warp + lr;
    const bool rok = row < row_end;
    const i64 rr = rok ? row : row_begin;
    const double* base = src.p + (rr / src.blk_rows) * src.bstride + (rr % src.blk_rows);
#pragma unroll 8
    for (int c0 = 0; c0 < ncols; c0 += CPP) {
      const int c = c0 + lc;
      const bool ok = rok && c < n;
      cp_async8(tile + ts_tile_idx<RW>(c, lr), ok ? base + (i64)c * src.ld : src.p, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (nchunks > 0) prefetch(0);
  __syncthreads();  // sR zeroed
  const int k0max = n < 32 ? n : 32;
  const int k1max = n - 32;
  for (i64 chunk = 0; chunk < nchunks; ++chunk) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    double a0[RW], a1[RW];
#pragma unroll
    for (int r = 0; r < RW; r += 2) {
      const double2 v = *reinterpret_cast<const double2*>(tile + ts_tile_idx<RW>(lane, r));
      a0[r] = v.x;
      a0[r + 1] = v.y;
      if (TWO) {
        const double2 u = *reinterpret_cast<const double2*>(tile + ts_tile_idx<RW>(lane + 32, r));
        a1[r] = u.x;
        a1[r + 1] = u.y;
      } else {
        a1[r] = a1[r + 1] = 0.0;
      }
    }
    __syncwarp();  // the tile is free: stream the next chunk into it while this one is reduced
    if (chunk + 1 < nchunks) prefetch(chunk + 1);
    for (int kl = 0; kl < k0max; ++kl) ts_step<NWARPS, RW, TWO, 0>(a0, a1, kl, lane, warp, sR, part, pubw);
    if (TWO)
      for (int kl = 0; kl < k1max; ++kl) ts_step<NWARPS, RW, TWO, 1>(a0, a1, kl, lane, warp, sR, part, pubw);
  }
  __syncthreads();
  double* out = Rout + (i64)blockIdx.x * out_row_step;
  for (int e = threadIdx.x; e < n * n; e += NWARPS * 32) {
    const int j = e / n, i = e - j * n;
    out[(i64)j * ldro + i] = i <= j ? sR[i * 64 + j] : 0.0;
  }
}

// ------------------------------------------------------------------------------- level 0, one chain per WARP
// tsqr_warp_kernel: the first level of the reduction for long row ranges.  The CTA-per-chunk kernel above has two
// chains per SM and pays a block barrier + a cross-warp exchange of partial dots in every column step.  Here every WARP
// owns a chain of its own: a private R (packed upper triangle, 16.6 KB of shared memory) and 32-row chunks held in
// registers with the same lane = column pair layout, so a column step needs no barrier and no reduction at all (a lane's
// dot runs over the rows in its own registers; the pivot lane's dot with itself is the norm and travels by one shuffle).
// Eight independent chains per SM (two warps per sub-partition: the 128 payload registers + the temporaries of a step
// need ~250 registers) hide each other's scalar chains.  The next chunk is pulled into L2 by prefetch instructions while
// the current one is reduced; its loads are 16-byte vector loads straight into the registers.  Measured: 10.2 ms against
// 11.0 ms for the CTA-per-chunk kernel at 8,388,608 x 64 (a column step still costs ~2000 cycles of latency per warp).
// For n <= 32 half of the payload registers are dead: 24-row chunks and twelve warps per SM there (tw_config below).
// Measured and not taken: every lane storing its column into a [row pair][lane] buffer at the end of a step (16 conflict-free
// full-warp STS.128) instead of the owner lane publishing alone in a diverged branch -- 10.70 against 10.24 ms (the extra
// shared-memory store traffic costs more than the 16 single-lane stores).
// two configurations: 8 warps x 32-row chunks (two warps per sub-partition, 255 registers: 128 payload registers + the
// temporaries of a step) and 12 warps x 24-row chunks (three per sub-partition, 168 registers, 96 of them payload).
constexpr int TW_WARPS = 8;
constexpr int TW_WARPS_24 = 12;
constexpr int TW_RPACK = 64 * 65 / 2;             // packed upper triangle of a 64 x 64 R
constexpr int TW_WARP_DOUBLES = TW_RPACK + 32;    // + the published pivot column
constexpr size_t tw_smem(int nw) { return (size_t)nw * TW_WARP_DOUBLES * sizeof(double); }

__device__ __forceinline__ int tw_row(int k) { return k * 64 - (k * (k - 1)) / 2 - k; }   // R(k, c) at tw_row(k) + c, c >= k

template <bool TWO, int H, int RW>
__device__ __forceinline__ void tw_step(double (&a0)[RW], double (&a1)[RW], const int kl, const int lane,
                                        double* __restrict__ sR, double* __restrict__ pubw) {
  const int k = 32 * H + kl;
  if (lane == kl) {
#pragma unroll
    for (int r = 0; r < RW; r += 2)
      *reinterpret_cast<double2*>(pubw + r) = H == 0 ? make_double2(a0[r], a0[r + 1]) : make_double2(a1[r], a1[r + 1]);
  }
  __syncwarp();
  double* rk = sR + tw_row(k);
  const bool own0 = H == 0 && lane >= kl;                    // this lane holds R(k, lane)
  const bool own1 = TWO && (H == 0 || lane >= kl);           // ... and R(k, 32 + lane)
  const double top0 = own0 ? rk[lane] : 0.0;
  const double top1 = own1 ? rk[32 + lane] : 0.0;
  // the published column is read twice (dot sweep, update sweep) with broadcast LDS.128: keeping it in 64 registers next to
  // the 128 payload registers would not fit the 168-register budget of twelve warps per SM
  double p0[4] = {0., 0., 0., 0.}, p1[4] = {0., 0., 0., 0.};
#pragma unroll
  for (int r = 0; r < RW; r += 4) {
    const double2 xa = *reinterpret_cast<const double2*>(pubw + r);
    const double2 xb = *reinterpret_cast<const double2*>(pubw + r + 2);
    if (H == 0) {
      p0[0] = fma(xa.x, a0[r], p0[0]);
      p0[1] = fma(xa.y, a0[r + 1], p0[1]);
      p0[2] = fma(xb.x, a0[r + 2], p0[2]);
      p0[3] = fma(xb.y, a0[r + 3], p0[3]);
    }
    if (TWO) {
      p1[0] = fma(xa.x, a1[r], p1[0]);
      p1[1] = fma(xa.y, a1[r + 1], p1[1]);
      p1[2] = fma(xb.x, a1[r + 2], p1[2]);
      p1[3] = fma(xb.y, a1[r + 3], p1[3]);
    }
  }
  const double d0 = (p0[0] + p0[1]) + (p0[2] + p0[3]);
  const double d1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
  const double dk = __shfl_sync(0xffffffffu, H == 0 ? d0 : d1, kl);
  const double alpha = __shfl_sync(0xffffffffu, H == 0 ? top0 : top1, kl);
  const double n2 = fma(alpha, alpha, dk);
  if (n2 != 0.0) {  // warp-uniform; zero column: tau = 0, nothing changes
    double nu, ixi, tau;
    if (n2 > 1e-280 && n2 < 1e280) {   // same scalar chain as tsqr_stream_kernel (Goldschmidt from the MUFU seeds)
      double y0, r;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(n2));
      double g = n2 * y0, hh = 0.5 * y0;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(alpha + copysign(g, alpha)));
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const double e = fma(-g, hh, 0.5);
        g = fma(g, e, g);
        hh = fma(hh, e, hh);
      }
      g = fma(fma(-g, g, n2), hh, g);
      nu = copysign(g, alpha);
      const double xi = alpha + nu;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const double e = fma(-xi, r, 1.0);
        r = fma(r, e, r);
      }
      ixi = r;
      tau = xi * copysign(hh + hh, alpha);
    } else {
      nu = copysign(sqrt(n2), alpha);
      const double xi = alpha + nu;
      ixi = 1.0 / xi;
      tau = xi / nu;
    }
    double t0 = 0., t1 = 0.;
    if (H == 0) {
      const bool right = lane > kl;
      const double s = tau * fma(d0, ixi, top0);
      t0 = right ? -(s * ixi) : 0.;
      if (own0) rk[lane] = right ? top0 - s : -nu;
    }
    if (TWO) {
      const bool right = (H == 0) || lane > kl;
      const double s = tau * fma(d1, ixi, top1);
      t1 = right ? -(s * ixi) : 0.;
      if (own1) rk[32 + lane] = right ? top1 - s : -nu;
    }
#pragma unroll
    for (int r = 0; r < RW; r += 2) {
      const double2 xa = *reinterpret_cast<const double2*>(pubw + r);
      if (H == 0) {
        a0[r] = fma(t0, xa.x, a0[r]);
        a0[r + 1] = fma(t0, xa.y, a0[r + 1]);
      }
      if (TWO) {
        a1[r] = fma(t1, xa.x, a1[r]);
        a1[r + 1] = fma(t1, xa.y, a1[r + 1]);
      }
    }
  }
  __syncwarp();  // every lane has read the published column before the next owner overwrites it
}

template <bool TWO, int RW, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
    tsqr_warp_kernel(const double* __restrict__ A, const i64 lda, const i64 m, const int n, const i64 rows_per_warp,
                     double* __restrict__ Rout, const i64 ldro, const i64 out_row_step, const int vec_ok) {
  extern __shared__ __align__(16) double tw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sR = tw_smem + warp * TW_WARP_DOUBLES;
  double* pubw = sR + TW_RPACK;
  for (int e = lane; e < TW_RPACK; e += 32) sR[e] = 0.0;
  __syncwarp();
  const i64 gw = (i64)blockIdx.x * NW + warp;
  const i64 row_begin = gw * rows_per_warp;
  i64 row_end = row_begin + rows_per_warp;
  if (row_end > m) row_end = m;
  const bool c0ok = lane < n, c1ok = TWO && lane + 32 < n;
  const double* col0 = A + (i64)(c0ok ? lane : 0) * lda;
  const double* col1 = A + (i64)(c1ok ? lane + 32 : 0) * lda;
  const int k0max = n < 32 ? n : 32;
  const int k1max = n - 32;
  for (i64 row0 = row_begin; row0 < row_end; row0 += RW) {
    double a0[RW], a1[RW];
    const bool full = row0 + RW <= row_end;
    if (full && vec_ok) {
#pragma unroll
      for (int r = 0; r < RW; r += 2) {
        const double2 v = c0ok ? *reinterpret_cast<const double2*>(col0 + row0 + r) : make_double2(0., 0.);
        a0[r] = v.x;
        a0[r + 1] = v.y;
      }
#pragma unroll
      for (int r = 0; r < RW; r += 2) {
        const double2 v = c1ok ? *reinterpret_cast<const double2*>(col1 + row0 + r) : make_double2(0., 0.);
        a1[r] = v.x;
        a1[r + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const bool rok = row0 + r < row_end;
        a0[r] = (rok && c0ok) ? col0[row0 + r] : 0.0;
        a1[r] = (rok && c1ok) ? col1[row0 + r] : 0.0;
      }
    }
    if (row0 + RW < row_end) {   // pull the next chunk of this lane's two columns (8 * RW bytes each) into L2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(col0 + row0 + RW));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(col0 + row0 + RW + 16));
      if (TWO) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(col1 + row0 + RW));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(col1 + row0 + RW + 16));
      }
    }
    for (int kl = 0; kl < k0max; ++kl) tw_step<TWO, 0, RW>(a0, a1, kl, lane, sR, pubw);
    if (TWO)
      for (int kl = 0; kl < k1max; ++kl) tw_step<TWO, 1, RW>(a0, a1, kl, lane, sR, pubw);
  }
  __syncwarp();
  double* out = Rout + gw * out_row_step;
  for (int e = lane; e < n * n; e += 32) {
    const int j = e / n, i = e - j * n;
    out[(i64)j * ldro + i] = i <= j ? sR[tw_row(i) + j] : 0.0;
  }
}

constexpr int TS_WARPS = 4;   // warps per CTA
constexpr int TS_RW = 32;     // rows per warp -> 128-row chunks (measured: 11.0 ms vs 14.3 ms for 8 warps x 16 rows)
constexpr int TS_CTAS_PER_SM = 2;  // two independent CTAs per SM: one's scalar chain / barrier hides behind the other's FMA sweeps

template <int NW, int RW>
static int launch_stream_cfg(const TsqrSrc& src, i64 m, int n, i64 rows_per_cta, i64 grid, double* out, i64 ldro,
                             i64 out_row_step, cudaStream_t st) {
  using Cfg = TsCfg<NW, RW>;
  static_assert(Cfg::ROWS == TsCfg<TS_WARPS, TS_RW>::ROWS, "all configurations use the same chunk height");
  if (n > 32) {
    auto kern = tsqr_stream_kernel<NW, RW, true>;
    GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    kern<<<(unsigned)grid, NW * 32, Cfg::SMEM, st>>>(src, m, n, rows_per_cta, out, ldro, out_row_step);
  } else {
    auto kern = tsqr_stream_kernel<NW, RW, false>;
    GLA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    kern<<<(unsigned)grid, NW * 32, Cfg::SMEM, st>>>(src, m, n, rows_per_cta, out, ldro, out_row_step);
  }
  GLA_CUDA(cudaGetLastError());
  return 0;
}

static int launch_stream(const TsqrSrc& src, i64 m, int n, i64 rows_per_cta, i64 grid, double* out, i64 ldro,
                         i64 out_row_step, cudaStream_t st) {
  static const int cfg = [] {
    const char* e = getenv("GLA_TSQR_CFG");   // tuning knob: 1 = 8 warps x 16 rows (more, thinner warps)
    return e ? atoi(e) : 0;
  }();
  if (cfg == 1) return launch_stream_cfg<8, 16>(src, m, n, rows_per_cta, grid, out, ldro, out_row_step, st);
  return launch_stream_cfg<TS_WARPS, TS_RW>(src, m, n, rows_per_cta, grid, out, ldro, out_row_step, st);
}

// rows per chunk of the per-warp level 0: 32 (8 warps per SM) for n > 32, 24 (12 warps) for n <= 32.  Measured at
// 8,388,608 rows: n = 64 10.23 ms with 32 rows against 10.66 ms with 24 (the FP64 pipe, at the 2.9 cycles per DFMA of
// these operand patterns, is what bounds it: a third warp per sub-partition only adds scalar chains); n = 32 4.34 against
// 3.99 ms (half the payload registers are dead there, the third warp is free).  GLA_TSQR_WARP_ROWS = 24 | 32 forces one.
static int tw_config(int n) {
  static const int forced = [] {
    const char* e = getenv("GLA_TSQR_WARP_ROWS");
    const int v = e ? atoi(e) : 0;
    return (v == 24 || v == 32) ? v : 0;
  }();
  return forced ? forced : (n > 32 ? 32 : 24);
}

template <bool TWO, int RW, int NW>
static int launch_warp_cfg(const TsqrSrc& src, i64 m, int n, i64 rows_per_warp, double* wbuf, i64 tall, int vec_ok,
                           cudaStream_t st) {
  auto kern = tsqr_warp_kernel<TWO, RW, NW>;
  GLA_TRY(ensure_dyn_smem((const void*)kern, (int)tw_smem(NW)));
  kern<<<(unsigned)sm_count(), NW * 32, tw_smem(NW), st>>>(src.p, src.ld, m, n, rows_per_warp, wbuf, tall, n, vec_ok);
  return check_cuda(cudaGetLastError(), __FILE__, __LINE__);
}

static int launch_warp(const TsqrSrc& src, i64 m, int n, i64 rows_per_warp, double* wbuf, i64 tall, int vec_ok,
                       cudaStream_t st) {
  if (tw_config(n) == 32)
    return n > 32 ? launch_warp_cfg<true, 32, TW_WARPS>(src, m, n, rows_per_warp, wbuf, tall, vec_ok, st)
                  : launch_warp_cfg<false, 32, TW_WARPS>(src, m, n, rows_per_warp, wbuf, tall, vec_ok, st);
  return n > 32 ? launch_warp_cfg<true, 24, TW_WARPS_24>(src, m, n, rows_per_warp, wbuf, tall, vec_ok, st)
                : launch_warp_cfg<false, 24, TW_WARPS_24>(src, m, n, rows_per_warp, wbuf, tall, vec_ok, st);
}

// src (m rows) -> dR (n x n, ldr).  Level 0 uses one CTA per SM; the stacked per-CTA R factors (a tall
// (grid*n) x n matrix) are reduced by the same kernel, one chunk per CTA, until a single CTA remains.
static int run_reduction(TsqrSrc src, i64 m, int n, double* dR, i64 ldr, cudaStream_t st) {
  constexpr int ROWS = TsCfg<TS_WARPS, TS_RW>::ROWS;
  const i64 sms = (i64)sm_count() * TS_CTAS_PER_SM;  // resident CTAs
  double* buf[2] = {nullptr, nullptr};
  int rc = 0, cur = 0;
  bool first = true;
  double* wbuf = nullptr;
  // level 0 of a long plain column-major range: one chain per warp (tsqr_warp_kernel), 12 x #SM R factors out
  static const bool no_warp = getenv("GLA_TSQR_NO_WARP") != nullptr;   // A/B switch
  const int tw_rows = tw_config(n);
  const i64 nwarps = (i64)sm_count() * (tw_rows == 32 ? TW_WARPS : TW_WARPS_24);
  // (worth it from ~1500 rows per warp on: below that the extra tree level over its 8 x #SM R factors costs more than the
  //  level-0 gain -- 8-GPU shards of the 8,388,608-row config: 1.92 ms with it, 1.86 ms without)
  if (!no_warp && src.bstride == 0 && src.blk_rows >= m && m >= (i64)sm_count() * TW_WARPS * 1536) {
    const i64 rows_per_warp = round_up((m + nwarps - 1) / nwarps, tw_rows);
    const i64 used = (m + rows_per_warp - 1) / rows_per_warp;        // warps that own rows (the others write a zero R)
    const i64 tall = nwarps * n;
    rc = pool_malloc(reinterpret_cast<void**>(&wbuf), (size_t)tall * n * sizeof(double), st);
    if (!rc) {
      const int vec_ok = ((reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && (src.ld & 1) == 0) ? 1 : 0;
      (void)used;
      rc = launch_warp(src, m, n, rows_per_warp, wbuf, tall, vec_ok, st);
      if (!rc) rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
      src = TsqrSrc{wbuf, tall, tall, 0};
      m = tall;
      first = false;
    }
  }
  while (!rc) {
    i64 grid = (m + ROWS - 1) / ROWS;
    if (first && grid > sms) grid = sms;                 // persistent: every CTA streams many chunks
    if (!first && grid > 1) grid = (m + 2 * ROWS - 1) / (2 * ROWS);  // upper levels: 2 chunks (4 R factors) per CTA
    if (grid < 1) grid = 1;
    i64 rows_per_cta = round_up((m + grid - 1) / grid, ROWS);
    grid = (m + rows_per_cta - 1) / rows_per_cta;
    if (grid <= 1) {
      rc = launch_stream(src, m, n, rows_per_cta, 1, dR, ldr, 0, st);
      break;
    }
    const i64 tall = grid * n;
    if (!buf[cur]) rc = pool_malloc(reinterpret_cast<void**>(&buf[cur]), (size_t)sms * n * n * sizeof(double), st);
    if (rc) break;
    rc = launch_stream(src, m, n, rows_per_cta, grid, buf[cur], tall, n, st);
    src = TsqrSrc{buf[cur], tall, tall, 0};
    m = tall;
    cur ^= 1;
    first = false;
  }
  for (double* b : buf)
    if (b) cudaFreeAsync(b, st);
  if (wbuf) cudaFreeAsync(wbuf, st);
  return rc;
}

int tsqr_local_dev(const double* dA, i64 m, i64 n, i64 lda, double* dR, i64 ldr, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0 || n > TSQR_MAXN) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (ldr < (n > 1 ? n : 1)) return -6;
  if (n == 0) return 0;
  if (m == 0) {
    GLA_CUDA(cudaMemset2DAsync(dR, ldr * sizeof(double), 0, n * sizeof(double), n, st));
    return 0;
  }
  return run_reduction(TsqrSrc{dA, lda, m, 0}, m, (int)n, dR, ldr, st);
}

int tsqr_combine_dev(const double* dRs, i64 count, i64 n, double* dR, i64 ldr, cudaStream_t st) {
  if (count < 0) return -2;
  if (n < 0 || n > TSQR_MAXN) return -3;
  if (ldr < (n > 1 ? n : 1)) return -5;
  if (n == 0) return 0;
  if (count == 0) {
    GLA_CUDA(cudaMemset2DAsync(dR, ldr * sizeof(double), 0, n * sizeof(double), n, st));
    return 0;
  }
  // `count` stacked n x n blocks (each ld n): block b holds rows b*n .. b*n + n - 1 of the tall matrix
  return run_reduction(TsqrSrc{dRs, n, n, n * n}, count * n, (int)n, dR, ldr, st);
}

// ------------------------------------------------------------------------------- NCCL exchange
// libnccl.so.2 is resolved lazily with dlopen so that the library loads on hosts without NCCL (inside a
// torch process the soname resolves to the copy torch already mapped).
namespace {
struct NcclUniqueId128 {  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
  char internal[128];
};
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
}  // namespace
namespace {
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api;
}
int nccl_check(int r, const char* file, int line) {
  if (r == 0) return 0;
  set_error(GLA_ERR_NCCL_CODE + r, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error", file, line);
  return GLA_ERR_NCCL_CODE + r;
}
int nccl_ready() {
  if (nccl().ok) return 0;
  set_error(GLA_ERR_NCCL_CODE, "libnccl.so.2 not found or incomplete", __FILE__, __LINE__);
  return GLA_ERR_NCCL_CODE;
}
}  // namespace

int nccl_unique_id(void* id128) {
  GLA_TRY(nccl_ready());
  return nccl_check(nccl().GetUniqueId(id128), __FILE__, __LINE__);
}
int nccl_comm_init(void** comm, int nranks, const void* id128, int rank) {
  GLA_TRY(nccl_ready());
  NcclUniqueId128 id;
  memcpy(&id, id128, sizeof(id));
  return nccl_check(nccl().CommInitRank(comm, nranks, id, rank), __FILE__, __LINE__);
}
int nccl_comm_destroy(void* comm) {
  GLA_TRY(nccl_ready());
  return nccl_check(nccl().CommDestroy(comm), __FILE__, __LINE__);
}

// dRloc (n x n, contiguous, ld n) of this rank -> all-gather over `comm` into dstack (nranks x n x n) -> every
// rank reduces the stack to the same dR.
int tsqr_allreduce_dev(void* comm, int nranks, const double* dRloc, i64 n, double* dstack, double* dR, i64 ldr,
                       cudaStream_t st) {
  if (!comm) return -1;
  if (nranks < 1) return -2;
  if (n < 0 || n > TSQR_MAXN) return -4;
  if (ldr < (n > 1 ? n : 1)) return -7;
  if (n == 0) return 0;
  GLA_TRY(nccl_ready());
  const int ncclDouble = 8;  // ncclFloat64 (nccl.h: ncclDataType_t)
  GLA_TRY(nccl_check(nccl().AllGather(dRloc, dstack, (size_t)(n * n), ncclDouble, comm, st), __FILE__, __LINE__));
  return tsqr_combine_dev(dstack, nranks, n, dR, ldr, st);
}

}  // namespace gla
