// gemm.cu -- K3: the TN contraction behind every trailing update (see gemm.cuh).
//
// Float64 fast path (gemm_tn_dmma_kernel):
//   * operands arrive by TMA: one cp.async.bulk.tensor.2d per operand per stage, box = 16 doubles of
//     K (128 B, the swizzle span) x BM / BN rows, CU_TENSOR_MAP_SWIZZLE_128B, completion on an
//     mbarrier (complete_tx); a dedicated producer warp runs STAGES slabs ahead of the consumers and
//     re-arms a slot when every consumer warp has arrived on its `empty` barrier;
//   * consumers issue mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4) on 32x32 warp tiles, fragments
//     read straight from the swizzled tile with LDS.64.  MMA row g of block b is mapped to tile row
//     16*(b/2) + 2g + (b%2): with that interleave the 16 lanes of a half-warp hit 16 distinct 8-byte
//     slots of the 128 B swizzle span (conflict free), and each thread ends up owning 2 consecutive
//     rows x 4 consecutive columns of C, so the epilogue is 16-byte coalesced read-modify-write of
//     full 128 B lines with no shared-memory staging;
//   * out-of-range rows / K tails are zero-filled by TMA, the epilogue masks M/N edges.
#include "gemm.cuh"

#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace gla {

// ------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Two rules for shared memory that TMA (the async proxy) writes and the warps read (both learnt the hard way: without
// them the kernels were bitwise reproducible alone on the GPU and produced wrong tiles as soon as ANY other kernel ran
// concurrently -- the look-ahead stream of the blocked QR or an unrelated stream):
//  1. fragment loads are explicit ld.shared on 32-bit addresses.  A pointer derived from the aligned-up dynamic-smem
//     base is a GENERIC pointer to the compiler, which emits LD.E (generic) plus 64-bit address arithmetic;
//  2. a stage is handed back to TMA only after fence.proxy.async (CUTLASS: fence_view_async_shared()).  The mbarrier
//     arrive orders generic-proxy accesses only; without the cross-proxy fence ptxas schedules the arrive directly
//     behind the ISSUE of the last LDS (ahead of the DMMAs that consume them), and a refill can overwrite the slab
//     while those loads are still in flight.
__device__ __forceinline__ void release_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------- DMMA + TMA kernel
template <int BM, int BN, int STAGES>
struct DmmaCfg {
  static constexpr int WM = BM / 32, WN = BN / 32, NCW = WM * WN;
  static constexpr int THREADS = NCW * 32;
  static constexpr int STAGE_BYTES = (BM + BN) * 128;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
};

// 256 threads = 8 consumer warps (no separate producer warp: with 9 warps the register file only
// fits one CTA per SM); lane 0 of warp 0 also drives TMA, refilling the slot released one iteration
// earlier so it almost never waits.  <= 128 registers -> two CTAs per SM, whose prologues/epilogues
// overlap each other's main loops.
template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(DmmaCfg<BM, BN, STAGES>::THREADS, 2)
    gemm_tn_dmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        double* __restrict__ C, i64 ldc, int M, int N, int K, int klen, i64 split_stride,
                        double alpha, int beta_one, int lower_only, int vec_ok) {
  using Cfg = DmmaCfg<BM, BN, STAGES>;
  constexpr int WM = Cfg::WM, NCW = Cfg::NCW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + STAGES * BM * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (lower_only == 1 && n0 >= m0 + BM) return;  // tile strictly above the diagonal
  if (lower_only == 2 && m0 >= n0 + BN) return;  // tile strictly below the diagonal
  const int kbeg = blockIdx.z * klen;
  const int kend = (kbeg + klen < K) ? kbeg + klen : K;
  const int nk = (kend - kbeg + 15) >> 4;
  const bool producer = threadIdx.x == 0;

  if (producer) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();                // the operands and C are written by the kernels before this one
  if (producer) {
    // prologue: fill the ring
    for (int it = 0; it < STAGES && it < nk; ++it) {
      mbar_expect_tx(&full[it], Cfg::STAGE_BYTES);
      tma_load_2d(sA + it * BM * 128, &tmA, kbeg + it * 16, m0, &full[it]);
      tma_load_2d(sB + it * BN * 128, &tmB, kbeg + it * 16, n0, &full[it]);
    }
  }

  // ---- 32x32 warp tile = 4 x 4 DMMA blocks; thread owns rows i0,i0+1 x 4 consecutive columns per (P,Q)
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  double* Cz = C + (i64)blockIdx.z * split_stride;
  double acc[4][4][2];
  // C is folded into the accumulators up front (beta = 1), so its HBM latency hides behind the TMA
  // prologue instead of trailing the main loop.  sgn = +-1 makes acc = sgn*C so that alpha = +-1
  // needs no multiply at the end: C' = C + alpha*AB = alpha*(alpha*C + AB).
  const bool preload = beta_one && (alpha == 1.0 || alpha == -1.0);
#pragma unroll
  for (int P = 0; P < 2; ++P) {
    const int i0 = m0 + wm * 32 + 16 * P + 2 * g;
#pragma unroll
    for (int Q = 0; Q < 2; ++Q) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = cc >> 1, e = cc & 1;
        const int j = n0 + wn * 32 + 16 * Q + 4 * t + cc;
        double v0 = 0.0, v1 = 0.0;
        if (preload && j < N && i0 < M) {
          const double* p = Cz + (i64)j * ldc + i0;
          if (vec_ok && i0 + 1 < M) {
            const double2 o = __ldcg(reinterpret_cast<const double2*>(p));
            v0 = o.x;
            v1 = o.y;
          } else {
            v0 = __ldcg(p);
            if (i0 + 1 < M) v1 = __ldcg(p + 1);
          }
          v0 *= alpha;
          v1 *= alpha;
        }
        acc[2 * P][2 * Q + e][c] = v0;
        acc[2 * P + 1][2 * Q + e][c] = v1;
      }
    }
  }
  __syncthreads();  // barrier inits visible to all consumers

  // byte offset of tile row for DMMA block b, lane row g:  row = 16*(b/2) + 2g + (b%2)
  int rowoff[4], key[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int r = 16 * (b >> 1) + 2 * g + (b & 1);
    rowoff[b] = r * 128;
    key[b] = r & 7;
  }
  const int tlo = (t & 1) << 3, thi = t >> 1;
  const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);

  for (int it = 0; it < nk; ++it) {
    const int s = it % STAGES;
    // refill the slot that was consumed in iteration it-1 with the slab of iteration it-1+STAGES
    if (producer && it >= 1 && it - 1 + STAGES < nk) {
      const int ps = (it - 1) % STAGES;
      mbar_wait(&empty[ps], ((it - 1) / STAGES) & 1);
      mbar_expect_tx(&full[ps], Cfg::STAGE_BYTES);
      tma_load_2d(sA + ps * BM * 128, &tmA, kbeg + (it - 1 + STAGES) * 16, m0, &full[ps]);
      tma_load_2d(sB + ps * BN * 128, &tmB, kbeg + (it - 1 + STAGES) * 16, n0, &full[ps]);
    }
    __syncwarp();
    mbar_wait(&full[s], (it / STAGES) & 1);
    const uint32_t pa = sA_u32 + s * BM * 128 + wm * 32 * 128;
    const uint32_t pb = sB_u32 + s * BN * 128 + wn * 32 * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int off = rowoff[x] + ((((2 * j + thi) ^ key[x]) << 4) | tlo);
        a[x] = lds_f64(pa + off);
        b[x] = lds_f64(pb + off);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) dmma884(acc[x][y], a[x], b[y]);
    }
    release_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  pdl_launch_dependents();   // main loop done: the next kernel of the stream may be scheduled (its pdl_wait covers our stores)
  // ---- epilogue
#pragma unroll
  for (int P = 0; P < 2; ++P) {
    const int i0 = m0 + wm * 32 + 16 * P + 2 * g;
    if (i0 >= M) continue;
    const bool two = i0 + 1 < M;
#pragma unroll
    for (int Q = 0; Q < 2; ++Q) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {  // column within the 4: cc = 2c + e
        const int c = cc >> 1, e = cc & 1;
        const int j = n0 + wn * 32 + 16 * Q + 4 * t + cc;
        if (j >= N) continue;
        double v0 = alpha * acc[2 * P][2 * Q + e][c];
        double v1 = alpha * acc[2 * P + 1][2 * Q + e][c];
        double* p = Cz + (i64)j * ldc + i0;
        const bool w0 = !lower_only || (lower_only == 1 ? i0 >= j : i0 <= j);
        const bool w1 = two && (!lower_only || (lower_only == 1 ? i0 + 1 >= j : i0 + 1 <= j));
        const bool add = beta_one && !preload;
        if (vec_ok && w0 && w1) {
          double2* pv = reinterpret_cast<double2*>(p);
          if (add) {
            double2 o = __ldcg(pv);
            v0 += o.x;
            v1 += o.y;
          }
          *pv = make_double2(v0, v1);
        } else {
          if (w0) p[0] = add ? __ldcg(p) + v0 : v0;
          if (w1) p[1] = add ? __ldcg(p + 1) + v1 : v1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------- Float32 on the tensor pipe (3xTF32)
// Same TMA / mbarrier ring as the Float64 kernel, K slabs of 32 floats (the 128 B swizzle span); the warps issue
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (SASS HMMA.1688.F32.TF32) on 32 x 32 warp tiles.  TF32 keeps 10
// mantissa bits, so every operand is split IN REGISTERS into hi = rna_tf32(x) and lo = rna_tf32(x - hi) and a product is
// accumulated as lo*hi + hi*lo + hi*hi (small terms first): Float32-level accuracy (the dropped lo*lo term is 2^-22
// relative) at a third of the TF32 rate, with no extra pass over the operands in memory.  tcgen05 kind::tf32 would
// read the fragments from shared memory directly and so would need the split tiles materialised there; see DESIGN.md.
// Fragment addressing: tile row r holds 32 consecutive k as eight 16-byte chunks, chunk c at c ^ (r & 7).  MMA row g of
// a 16-row block is tile row base + g (+8), so r & 7 = g and the eight rows a quarter-warp... the 32 lanes of one
// LDS.32 (8 rows x 4 consecutive words of one chunk each) hit 32 distinct banks.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(DmmaCfg<BM, BN, STAGES>::THREADS, 2)
    gemm_tn_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          float* __restrict__ C, i64 ldc, int M, int N, int K, int klen, i64 split_stride, float alpha,
                          int beta_one, int lower_only) {
  using Cfg = DmmaCfg<BM, BN, STAGES>;   // same tile bookkeeping: 128 B per tile row and stage
  constexpr int WM = Cfg::WM, NCW = Cfg::NCW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + STAGES * BM * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (lower_only == 1 && n0 >= m0 + BM) return;
  if (lower_only == 2 && m0 >= n0 + BN) return;
  const int kbeg = blockIdx.z * klen;
  const int kend = (kbeg + klen < K) ? kbeg + klen : K;
  const int nk = (kend - kbeg + 31) >> 5;
  const bool producer = threadIdx.x == 0;

  if (producer) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();   // programmatic dependent launch (common.cuh): nothing global is touched before this point
  if (producer) {
    for (int it = 0; it < STAGES && it < nk; ++it) {
      mbar_expect_tx(&full[it], Cfg::STAGE_BYTES);
      tma_load_2d(sA + it * BM * 128, &tmA, kbeg + it * 32, m0, &full[it]);
      tma_load_2d(sB + it * BN * 128, &tmB, kbeg + it * 32, n0, &full[it]);
    }
  }
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  float acc[2][4][4];
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int Q = 0; Q < 4; ++Q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[P][Q][e] = 0.f;
  __syncthreads();  // barrier inits visible to all consumers

  const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
  for (int it = 0; it < nk; ++it) {
    const int s = it % STAGES;
    if (producer && it >= 1 && it - 1 + STAGES < nk) {
      const int ps = (it - 1) % STAGES;
      mbar_wait(&empty[ps], ((it - 1) / STAGES) & 1);
      mbar_expect_tx(&full[ps], Cfg::STAGE_BYTES);
      tma_load_2d(sA + ps * BM * 128, &tmA, kbeg + (it - 1 + STAGES) * 32, m0, &full[ps]);
      tma_load_2d(sB + ps * BN * 128, &tmB, kbeg + (it - 1 + STAGES) * 32, n0, &full[ps]);
    }
    __syncwarp();
    mbar_wait(&full[s], (it / STAGES) & 1);
    const uint32_t pa = sA_u32 + s * BM * 128 + (wm * 32 + g) * 128 + 4 * t;
    const uint32_t pb = sB_u32 + s * BN * 128 + (wn * 32 + g) * 128 + 4 * t;
    // The tensor pipe adds into its accumulator with truncation (a bias that grows linearly with the number of chained
    // MMAs: 1.3e-4 on the Gram probe of a 16384^2 factorisation, K up to 16384, against 2.5e-6 for FFMA).  So the MMAs of
    // ONE slab (12 per block) chain into a fresh accumulator, which is then added to the running sum with a rounded FADD.
    float part[2][4][4];
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int Q = 0; Q < 4; ++Q)
#pragma unroll
        for (int e = 0; e < 4; ++e) part[P][Q][e] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // four k-steps of 8 inside the slab
      const uint32_t c0 = (uint32_t)(((2 * j) ^ g) << 4), c1 = (uint32_t)(((2 * j + 1) ^ g) << 4);
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int P = 0; P < 2; ++P) {
        float v[4];
        v[0] = lds_f32(pa + (16 * P) * 128 + c0);        // (row g,     k = t)
        v[1] = lds_f32(pa + (16 * P + 8) * 128 + c0);    // (row g + 8, k = t)
        v[2] = lds_f32(pa + (16 * P) * 128 + c1);        // (row g,     k = t + 4)
        v[3] = lds_f32(pa + (16 * P + 8) * 128 + c1);    // (row g + 8, k = t + 4)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ah[P][e] = tf32_rna(v[e]);
          al[P][e] = tf32_rna(v[e] - __uint_as_float(ah[P][e]));
        }
      }
#pragma unroll
      for (int Q = 0; Q < 4; ++Q) {
        float v[2];
        v[0] = lds_f32(pb + (8 * Q) * 128 + c0);         // (k = t,     n = g)
        v[1] = lds_f32(pb + (8 * Q) * 128 + c1);         // (k = t + 4, n = g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          bh[Q][e] = tf32_rna(v[e]);
          bl[Q][e] = tf32_rna(v[e] - __uint_as_float(bh[Q][e]));
        }
      }
#pragma unroll
      for (int P = 0; P < 2; ++P)
#pragma unroll
        for (int Q = 0; Q < 4; ++Q) {
          mma_tf32(part[P][Q], al[P], bh[Q]);
          mma_tf32(part[P][Q], ah[P], bl[Q]);
          mma_tf32(part[P][Q], ah[P], bh[Q]);
        }
    }
#pragma unroll
    for (int P = 0; P < 2; ++P)
#pragma unroll
      for (int Q = 0; Q < 4; ++Q)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[P][Q][e] += part[P][Q][e];
    release_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  pdl_launch_dependents();   // main loop done: the next kernel of the stream may be scheduled (its pdl_wait covers our stores)
  // ---- epilogue: lane owns (i, j), (i, j + 1), (i + 8, j), (i + 8, j + 1) of every 16 x 8 block
  float* Cz = C + (i64)blockIdx.z * split_stride;
#pragma unroll
  for (int P = 0; P < 2; ++P)
#pragma unroll
    for (int Q = 0; Q < 4; ++Q)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = m0 + wm * 32 + 16 * P + g + 8 * (e >> 1);
        const int j = n0 + wn * 32 + 8 * Q + 2 * t + (e & 1);
        if (i >= M || j >= N) continue;
        if (lower_only == 1 && i < j) continue;
        if (lower_only == 2 && i > j) continue;
        float* p = Cz + (i64)j * ldc + i;
        const float v = alpha * acc[P][Q][e];
        *p = beta_one ? __ldcg(p) + v : v;
      }
}

// ------------------------------------------------------------------------------- Float32 on tcgen05 (UMMA, 3xTF32)
// The 5th-generation tensor core path for the Float32 contractions: tcgen05.mma.cta_group::1.kind::tf32 (SASS UTCHMMA),
// operands read by the tensor core straight from the 128B-swizzled tiles TMA delivers (K-major for both operands: the TN
// contraction is exactly UMMA's native K-major x K-major form), accumulators in tensor memory.  Persistent, warp
// specialised, one CTA per SM:
//     warp 0      TMA producer (one lane): 128 x 32-float boxes of At and B per stage into a 3-deep ring
//     warps 2-3   operand split: x -> hi = tf32(x) (round to nearest), lo = x - hi (exact), hi written back over the raw
//                 tile and lo into a sibling tile -- the same byte offsets, so the swizzle never has to be decoded; then
//                 fence.proxy.async (generic-proxy stores -> tensor-core reads) and an arrive on the stage's `ready` barrier
//     warp 1      MMA issuer (one lane): per stage 4 k-steps x {lo*hi, hi*lo, hi*hi} = 12 UMMAs of 128 x 128 x 8 into ONE
//                 of four 128-column TMEM buffers, tcgen05.commit -> `empty` (stage back to TMA) and -> `tmem_full`
//     warps 4-11  accumulation + epilogue: tcgen05.ld of the slab's partial product and a ROUNDED add into register
//                 accumulators (the tensor core adds into its accumulator with truncation, a bias that grows linearly with
//                 the chain length: every slab starts a fresh TMEM accumulator, exactly like the mma.sync kernel above);
//                 after the last slab of a tile the 128 x 128 block goes to global memory (alpha / beta, edge masks)
//                 while the MMA warp is already up to four slabs into the next tile.
namespace umma {

constexpr int BM = 128, BN = 128, BKF = 32;          // tile, floats of K per stage (one 128 B swizzle span)
constexpr int STAGES = 3;
constexpr int TBUF = 4;                              // TMEM accumulator buffers of BN columns
constexpr int THREADS = 384;
constexpr int SPLIT_WARP0 = 2, SPLIT_THREADS = 64, EPI_WARP0 = 4, EPI_THREADS = 256;
constexpr int TILE_BYTES = BM * 128;                 // 16 KB: 128 rows of 128 B
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A raw/hi, B raw/hi, A lo, B lo
constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO = 1 (ignored for swizzled
// K-major), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// mbarrier wait that cannot hang the device: a second without progress is a protocol bug -> trap (the launch fails loudly)
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 0xfff) == 0xfff) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ll) __trap();
    }
  }
}

struct TileIter {   // static round-robin over (m tile, n tile, K slice), skipping tiles outside the requested triangle
  int mt, nt, nz, lower_only;
  __device__ __forceinline__ bool decode(int t, int& m0, int& n0, int& z) const {
    const int per = mt * nt;
    z = t / per;
    const int r = t - z * per;
    n0 = (r / mt) * BN;
    m0 = (r % mt) * BM;
    if (lower_only == 1 && n0 >= m0 + BM) return false;
    if (lower_only == 2 && m0 >= n0 + BN) return false;
    return true;
  }
};

__global__ void __launch_bounds__(THREADS, 1)
    gemm_tn_umma_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                               float* __restrict__ C, i64 ldc, int M, int N, int K, int klen, int nsplit, i64 split_stride,
                               float alpha, int beta_one, int lower_only, long long* __restrict__ trace) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);   // TMA landed
  uint64_t* ready = full + STAGES;                                             // split done
  uint64_t* empty = ready + STAGES;                                            // MMAs of the stage retired
  uint64_t* tfull = empty + STAGES;                                            // TMEM buffer holds a slab product
  uint64_t* tempty = tfull + TBUF;                                             // TMEM buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + TBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TileIter ti{(M + BM - 1) / BM, (N + BN - 1) / BN, nsplit, lower_only};
  const int ntiles = ti.mt * ti.nt * nsplit;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], SPLIT_THREADS / 32);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < TBUF; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], EPI_THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // all of TMEM: TBUF x BN = 512 columns (one CTA per SM)
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TBUF * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0, z;
        if (!ti.decode(t, m0, n0, z)) continue;
        const int kbeg = z * klen, kend = min(kbeg + klen, K);
        const int nk = (kend - kbeg + BKF - 1) / BKF;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % STAGES;
          wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          if (trace && blockIdx.x == 0 && it < 256) trace[it] = clock64();
          unsigned char* st = base + s * STAGE_BYTES;
          mbar_expect_tx(&full[s], 2 * TILE_BYTES);
          tma_load_2d(st, &tmA, kbeg + kb * BKF, m0, &full[s]);
          tma_load_2d(st + TILE_BYTES, &tmB, kbeg + kb * BKF, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, sl = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0, z;
        if (!ti.decode(t, m0, n0, z)) continue;
        const int kbeg = z * klen, kend = min(kbeg + klen, K);
        const int nk = (kend - kbeg + BKF - 1) / BKF;
        for (int kb = 0; kb < nk; ++kb, ++it, ++sl) {
          const int s = it % STAGES, b = sl % TBUF;
          wait(&tempty[b], ((sl / TBUF) & 1) ^ 1);
          wait(&ready[s], (it / STAGES) & 1);
          if (trace && blockIdx.x == 0 && it < 256) trace[256 + it] = clock64();
          fence_after();
          const uint32_t st = smem_u32(base + s * STAGE_BYTES);
          const uint64_t a_hi = smem_desc(st), b_hi = smem_desc(st + TILE_BYTES);
          const uint64_t a_lo = smem_desc(st + 2 * TILE_BYTES), b_lo = smem_desc(st + 3 * TILE_BYTES);
          const uint32_t d = tmem_base + (uint32_t)(b * BN);
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {       // UMMA_K = 8 floats = 32 B: +2 in the descriptor's 16-byte units
            const uint64_t o = (uint64_t)(2 * k);
            mma_tf32_ss(d, a_lo + o, b_hi + o, k > 0 ? 1u : 0u);
            mma_tf32_ss(d, a_hi + o, b_lo + o, 1u);
            mma_tf32_ss(d, a_hi + o, b_hi + o, 1u);
          }
          commit(&empty[s]);    // the stage returns to TMA when these MMAs have read it
          commit(&tfull[b]);    // and the slab product is complete in TMEM
        }
      }
    }
  } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + SPLIT_THREADS / 32) {
    // ================================================================== operand split (hi / lo tiles)
    const int tid = threadIdx.x - SPLIT_WARP0 * 32;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, z;
      if (!ti.decode(t, m0, n0, z)) continue;
      const int kbeg = z * klen, kend = min(kbeg + klen, K);
      const int nk = (kend - kbeg + BKF - 1) / BKF;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % STAGES;
        wait(&full[s], (it / STAGES) & 1);
        if (trace && blockIdx.x == 0 && it < 256 && tid == 0) trace[512 + it] = clock64();
        const uint32_t st = smem_u32(base + s * STAGE_BYTES);
#pragma unroll 4
        for (int e = tid; e < 2 * TILE_BYTES / 16; e += SPLIT_THREADS) {
          const uint32_t a = st + (uint32_t)e * 16;
          uint32_t x0, x1, x2, x3;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(a));
          // (measured alternative: truncation split without the hi write-back, relying on the tensor core ignoring the 13 low
          //  bits -- 6 % faster, but errors 2x larger: 8.7e-7 instead of 4.2e-7 on the rank-k probe; not taken)
          const uint32_t h0 = (x0 + 0x1000u) & 0xffffe000u, h1 = (x1 + 0x1000u) & 0xffffe000u;   // round to nearest TF32
          const uint32_t h2 = (x2 + 0x1000u) & 0xffffe000u, h3 = (x3 + 0x1000u) & 0xffffe000u;   // (cvt.rna runs at 1/4 rate)
          const uint32_t l0 = __float_as_uint(__uint_as_float(x0) - __uint_as_float(h0));
          const uint32_t l1 = __float_as_uint(__uint_as_float(x1) - __uint_as_float(h1));
          const uint32_t l2 = __float_as_uint(__uint_as_float(x2) - __uint_as_float(h2));
          const uint32_t l3 = __float_as_uint(__uint_as_float(x3) - __uint_as_float(h3));
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2 * TILE_BYTES), "r"(l0), "r"(l1), "r"(l2), "r"(l3)
                       : "memory");
        }
        release_fence();          // generic-proxy stores -> async-proxy (tensor core) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
        if (trace && blockIdx.x == 0 && it < 256 && tid == 0) trace[768 + it] = clock64();
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_THREADS / 32) {
    // ================================================================== accumulate + epilogue
    const int ew = warp - EPI_WARP0;              // 0..7
    const int quarter = warp & 3;                 // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
    const int chalf = ew >> 2;                    // columns 64 * chalf .. + 63
    const uint32_t lane_addr = ((uint32_t)(quarter * 32) << 16) + (uint32_t)(chalf * 64);
    uint32_t sl = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, z;
      if (!ti.decode(t, m0, n0, z)) continue;
      const int kbeg = z * klen, kend = min(kbeg + klen, K);
      const int nk = (kend - kbeg + BKF - 1) / BKF;
      float acc[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc[c] = 0.f;
      for (int kb = 0; kb < nk; ++kb, ++sl) {
        const int b = sl % TBUF;
        wait(&tfull[b], (sl / TBUF) & 1);
        if (trace && blockIdx.x == 0 && sl < 256 && threadIdx.x == EPI_WARP0 * 32) trace[1024 + sl] = clock64();
        fence_after();
        const uint32_t ta = tmem_base + lane_addr + (uint32_t)(b * BN);
        float v[32];
        tmem_ld32(ta, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] += v[c];
        tmem_ld32(ta + 32, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[32 + c] += v[c];
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[b]);
      }
      const int i = m0 + quarter * 32 + lane;
      if (i < M) {
        float* Cz = C + (i64)z * split_stride + i;
        const bool b1 = beta_one && nsplit == 1;
        // read-modify-write in batches of 32 columns: all loads of a batch are in flight before its first store
        // (interleaved ld/st pairs serialise on the store -> load ordering: one DRAM round trip per column)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float cv[32];
          if (b1) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int j = n0 + chalf * 64 + h * 32 + c;
              const bool keep = j < N && !(lower_only == 1 && i < j) && !(lower_only == 2 && i > j);
              cv[c] = keep ? __ldcg(Cz + (i64)j * ldc) : 0.f;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int j = n0 + chalf * 64 + h * 32 + c;
            const bool keep = j < N && !(lower_only == 1 && i < j) && !(lower_only == 2 && i > j);
            const float r = alpha * acc[h * 32 + c];
            if (keep) Cz[(i64)j * ldc] = b1 ? cv[c] + r : r;
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TBUF * BN) : "memory");
  }
}

}  // namespace umma

// ------------------------------------------------------------------------------- ComplexF64 on the DMMA pipe
// C(i,j) = sum_k op(At(k,i)) B(k,j) for interleaved complex operands, computed as TWO real contractions over the
// 2K interleaved doubles of each operand row (the tile a TMA box delivers IS that real row):
//     conj:  Re C = sum_k' a[k'] b[k'],            Im C = sum_k' a[k'] b~[k'],   b~ = (b_im, -b_re) per pair
//     plain: Re C = sum_k' a[k'] (b_re, -b_im),    Im C = sum_k' a[k'] (b_im,  b_re)
// i.e. the B fragment of the "imaginary" product is the lane's PARTNER element inside the 16-byte complex number
// (address ^ 8) with a sign flip on odd lanes -- an XOR on the sign bit, no FP64 instruction.  4 real DMMA flops
// per complex FMA, the minimum without the 3M trick.  Warp tile 32 x 16 complex (Re and Im accumulators = the
// same 64 registers as the real kernel's 32 x 32 tile), CTA tile BM x BN complex, same TMA / mbarrier ring.
template <int BM, int BN, int STAGES>
struct ZdmmaCfg {
  static constexpr int WM = BM / 32, WN = BN / 16, NCW = WM * WN;
  static constexpr int THREADS = NCW * 32;
  static constexpr int STAGE_BYTES = (BM + BN) * 128;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;
};

__device__ __forceinline__ double xor_sign(double v, unsigned m) {
  return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
}

template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(ZdmmaCfg<BM, BN, STAGES>::THREADS, 2)
    gemm_tn_zdmma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         zd* __restrict__ C, i64 ldc, int M, int N, int K2, int klen2, i64 split_stride, double alpha,
                         int beta_one, int conj_a, int lower_only) {
  using Cfg = ZdmmaCfg<BM, BN, STAGES>;
  constexpr int WM = Cfg::WM, NCW = Cfg::NCW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + STAGES * BM * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (lower_only == 1 && n0 >= m0 + BM) return;
  if (lower_only == 2 && m0 >= n0 + BN) return;
  const int kbeg = blockIdx.z * klen2;                       // in doubles
  const int kend = (kbeg + klen2 < K2) ? kbeg + klen2 : K2;
  const int nk = (kend - kbeg + 15) >> 4;
  const bool producer = threadIdx.x == 0;

  if (producer) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();   // programmatic dependent launch (common.cuh): nothing global is touched before this point
  if (producer) {
    for (int it = 0; it < STAGES && it < nk; ++it) {
      mbar_expect_tx(&full[it], Cfg::STAGE_BYTES);
      tma_load_2d(sA + it * BM * 128, &tmA, kbeg + it * 16, m0, &full[it]);
      tma_load_2d(sB + it * BN * 128, &tmB, kbeg + it * 16, n0, &full[it]);
    }
  }

  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  zd* Cz = C + (i64)blockIdx.z * split_stride;
  // thread owns rows i0, i0+1 (per P) x columns 4t .. 4t+3 of the warp's 32 x 16 tile; column 4t + 2c + e lives in
  // accumulator [.][e][c] (same interleave as the real kernel)
  double are[4][2][2], aim[4][2][2];
  const bool preload = beta_one && (alpha == 1.0 || alpha == -1.0);
#pragma unroll
  for (int P = 0; P < 2; ++P) {
    const int i0 = m0 + wm * 32 + 16 * P + 2 * g;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = cc >> 1, e = cc & 1;
      const int j = n0 + wn * 16 + 4 * t + cc;
      zd v0 = make_zd(0., 0.), v1 = make_zd(0., 0.);
      if (preload && j < N && i0 < M) {
        const zd* p = Cz + (i64)j * ldc + i0;
        v0 = ldcg_t(p);
        if (i0 + 1 < M) v1 = ldcg_t(p + 1);
        v0 = scale_real(v0, alpha);
        v1 = scale_real(v1, alpha);
      }
      are[2 * P][e][c] = v0.x;
      aim[2 * P][e][c] = v0.y;
      are[2 * P + 1][e][c] = v1.x;
      aim[2 * P + 1][e][c] = v1.y;
    }
  }
  __syncthreads();

  int rowoff[4], key[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int r = 16 * (b >> 1) + 2 * g + (b & 1);
    rowoff[b] = r * 128;
    key[b] = r & 7;
  }
  const int tlo = (t & 1) << 3, thi = t >> 1;
  const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
  const unsigned SIGN = 0x80000000u;
  const unsigned mre = (t & 1) ? (conj_a ? 0u : SIGN) : 0u;   // flips the own element feeding Re C
  const unsigned mim = (t & 1) ? (conj_a ? SIGN : 0u) : 0u;   // flips the partner element feeding Im C

  for (int it = 0; it < nk; ++it) {
    const int s = it % STAGES;
    if (producer && it >= 1 && it - 1 + STAGES < nk) {
      const int ps = (it - 1) % STAGES;
      mbar_wait(&empty[ps], ((it - 1) / STAGES) & 1);
      mbar_expect_tx(&full[ps], Cfg::STAGE_BYTES);
      tma_load_2d(sA + ps * BM * 128, &tmA, kbeg + (it - 1 + STAGES) * 16, m0, &full[ps]);
      tma_load_2d(sB + ps * BN * 128, &tmB, kbeg + (it - 1 + STAGES) * 16, n0, &full[ps]);
    }
    __syncwarp();
    mbar_wait(&full[s], (it / STAGES) & 1);
    const uint32_t pa = sA_u32 + s * BM * 128 + wm * 32 * 128;
    const uint32_t pb = sB_u32 + s * BN * 128 + wn * 16 * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double a[4], bre[2], bim[2];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int off = rowoff[x] + ((((2 * j + thi) ^ key[x]) << 4) | tlo);
        a[x] = lds_f64(pa + off);
      }
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int off = rowoff[y] + ((((2 * j + thi) ^ key[y]) << 4) | tlo);
        bre[y] = xor_sign(lds_f64(pb + off), mre);
        bim[y] = xor_sign(lds_f64(pb + (off ^ 8)), mim);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) {
          dmma884(are[x][y], a[x], bre[y]);
          dmma884(aim[x][y], a[x], bim[y]);
        }
    }
    release_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  pdl_launch_dependents();   // main loop done: the next kernel of the stream may be scheduled (its pdl_wait covers our stores)
  // ---- epilogue: two consecutive complex rows = 32 contiguous bytes per column
#pragma unroll
  for (int P = 0; P < 2; ++P) {
    const int i0 = m0 + wm * 32 + 16 * P + 2 * g;
    if (i0 >= M) continue;
    const bool two = i0 + 1 < M;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = cc >> 1, e = cc & 1;
      const int j = n0 + wn * 16 + 4 * t + cc;
      if (j >= N) continue;
      zd v0 = make_zd(alpha * are[2 * P][e][c], alpha * aim[2 * P][e][c]);
      zd v1 = make_zd(alpha * are[2 * P + 1][e][c], alpha * aim[2 * P + 1][e][c]);
      zd* p = Cz + (i64)j * ldc + i0;
      const bool w0 = !lower_only || (lower_only == 1 ? i0 >= j : i0 <= j);
      const bool w1 = two && (!lower_only || (lower_only == 1 ? i0 + 1 >= j : i0 + 1 <= j));
      const bool add = beta_one && !preload;
      if (w0) p[0] = add ? ldcg_t(p) + v0 : v0;
      if (w1) p[1] = add ? ldcg_t(p + 1) + v1 : v1;
    }
  }
}

// ------------------------------------------------------------------------------- generic FMA kernel
// 64x64 tile, 16-deep K slabs, 256 threads, 4x4 outputs per thread.
template <class T>
__global__ void __launch_bounds__(256)
    gemm_tn_fma_kernel(const T* __restrict__ At, i64 ldat, const T* __restrict__ B, i64 ldb, T* __restrict__ C,
                       i64 ldc, int M, int N, int K, int klen, i64 split_stride, typename Sc<T>::real alpha,
                       int beta_one, int conj_a, int lower_only) {
  __shared__ T sA[16][64 + 1];
  __shared__ T sB[16][64 + 1];
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  if (lower_only == 1 && n0 >= m0 + 64) return;
  if (lower_only == 2 && m0 >= n0 + 64) return;
  const int kbeg = blockIdx.z * klen;
  const int kend = (kbeg + klen < K) ? kbeg + klen : K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> rows (M), ty -> cols (N)
  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = Sc<T>::zero();
  const int lk = threadIdx.x & 15, lr = threadIdx.x >> 4;  // loader: k index, row group
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = lr + 16 * r;
      const int k = k0 + lk;
      T va = Sc<T>::zero(), vb = Sc<T>::zero();
      if (k < kend && m0 + i < M) va = ldcg_t(At + (i64)(m0 + i) * ldat + k);
      if (k < kend && n0 + i < N) vb = ldcg_t(B + (i64)(n0 + i) * ldb + k);
      sA[lk][i] = conj_a ? cj(va) : va;
      sB[lk][i] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      T a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        a[x] = sA[k][tx + 16 * x];
        b[x] = sB[k][ty + 16 * x];
      }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmad(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  T* Cz = C + (i64)blockIdx.z * split_stride;
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const int j = n0 + ty + 16 * y;
    if (j >= N) continue;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int i = m0 + tx + 16 * x;
      if (i >= M) continue;
      if (lower_only == 1 && i < j) continue;
      if (lower_only == 2 && i > j) continue;
      T v = scale_real(acc[x][y], alpha);
      T* p = Cz + (i64)j * ldc + i;
      *p = beta_one ? ldcg_t(p) + v : v;
    }
  }
}

template <class T>
__global__ void sum_splits_kernel(T* __restrict__ out, i64 ldo, const T* __restrict__ part, i64 ldp, i64 stride,
                                  int nsplit, i64 M, i64 N) {
  const i64 total = M * N;
  pdl_launch_dependents();
  pdl_wait();
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 j = e / M, i = e - j * M;
    T s = ldcg_t(part + j * ldp + i);
    for (int z = 1; z < nsplit; ++z) s = s + ldcg_t(part + (i64)z * stride + j * ldp + i);
    out[j * ldo + i] = s;
  }
}

// ------------------------------------------------------------------------------- host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static std::mutex mu;
  static EncodeTiledFn fn = nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// K x R column-major operand (K contiguous), box = one 128-byte swizzle span of K (16 doubles / 32 floats) x rows
int make_map_t(CUtensorMap* tm, const void* base, i64 K, i64 R, i64 ld, int box_rows, int elem_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error(GLA_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return GLA_ERR_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d, K=%lld R=%lld ld=%lld)", (int)r,
             (long long)K, (long long)R, (long long)ld);
    set_error(GLA_ERR_DRIVER, msg, __FILE__, __LINE__);
    return GLA_ERR_DRIVER;
  }
  return 0;
}
int make_map(CUtensorMap* tm, const double* base, i64 K, i64 R, i64 ld, int box_rows) {
  return make_map_t(tm, base, K, R, ld, box_rows, 8);
}

template <int BM, int BN, int STAGES>
int launch_tf32x3(const GemmTN<float>& g, int klen, cudaStream_t st) {
  using Cfg = DmmaCfg<BM, BN, STAGES>;
  CUtensorMap tmA, tmB;
  GLA_TRY(make_map_t(&tmA, g.At, g.K, g.M, g.ldat, BM, 4));
  GLA_TRY(make_map_t(&tmB, g.B, g.K, g.N, g.ldb, BN, 4));
  auto kern = gemm_tn_tf32x3_kernel<BM, BN, STAGES>;
  GLA_TRY(ensure_dyn_smem((const void*)kern, (int)(Cfg::SMEM)));
  dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.N, BN), (unsigned)g.nsplit);
  GLA_CUDA(launch_pdl(kern, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM, st, tmA, tmB, g.C, g.ldc, (int)g.M, (int)g.N, (int)g.K,
                      klen, g.split_stride, g.alpha, g.nsplit > 1 ? 0 : g.beta_one, g.lower_only));
  return 0;
}

int launch_umma_tf32x3(const GemmTN<float>& g, int klen, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  GLA_TRY(make_map_t(&tmA, g.At, g.K, g.M, g.ldat, umma::BM, 4));
  GLA_TRY(make_map_t(&tmB, g.B, g.K, g.N, g.ldb, umma::BN, 4));
  auto kern = umma::gemm_tn_umma_tf32x3_kernel;
  GLA_TRY(ensure_dyn_smem((const void*)kern, umma::SMEM));
  const i64 tiles = (i64)ceil_div(g.M, umma::BM) * ceil_div(g.N, umma::BN) * g.nsplit;
  // persistent, one CTA per SM -- but a CTA retires after at most `tpc` tiles: a grid of immortal CTAs would keep every SM
  // until the whole product is done, and the high-priority panel chain of the look-ahead schedule (whose kernels need
  // SMs NOW) would wait behind it
  // (f32 qrBlocked! n = 16384: 136.5 ms with immortal CTAs, 131.9 / 127.9 / 125.1 ms with 8 / 4 / 2 tiles per CTA; a product that
  // runs alone prefers the immortal form: 3.97 against 4.48 ms for the wide block application)
  static const i64 tpc_env = [] { const char* e = getenv("GLA_UMMA_TILES_PER_CTA"); return e ? (i64)atoi(e) : -1ll; }();
  const i64 tpc = tpc_env >= 0 ? tpc_env : (g.yield_sms ? 2 : 0);
  i64 grid64 = tiles < sm_count() ? tiles : sm_count();
  if (tpc > 0 && (tiles + tpc - 1) / tpc > grid64) grid64 = (tiles + tpc - 1) / tpc;
  const unsigned grid = (unsigned)grid64;
  long long* trace = nullptr;
  static const char* trace_path = getenv("GLA_UMMA_TRACE");   // development aid: clock64 stamps of CTA 0's first 256 stages
  if (trace_path && g.K >= 8192) {
    GLA_CUDA(cudaMalloc(&trace, 1280 * sizeof(long long)));
    GLA_CUDA(cudaMemset(trace, 0, 1280 * sizeof(long long)));
  }
  kern<<<grid, umma::THREADS, umma::SMEM, st>>>(tmA, tmB, g.C, g.ldc, (int)g.M, (int)g.N, (int)g.K, klen, g.nsplit,
                                                g.split_stride, g.alpha, g.beta_one, g.lower_only, trace);
  GLA_CUDA(cudaGetLastError());
  if (trace) {
    static long long host[1280];
    GLA_CUDA(cudaStreamSynchronize(st));
    GLA_CUDA(cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    if (FILE* f = fopen(trace_path, "w")) {
      fprintf(f, "# M=%lld N=%lld K=%lld: stage, tma_issue, split_sees_full, split_done, mma_sees_ready, epilogue_sees_tfull (cycles from the first TMA)\n",
              (long long)g.M, (long long)g.N, (long long)g.K);
      for (int i = 0; i < 256; ++i)
        fprintf(f, "%d %lld %lld %lld %lld %lld\n", i, host[i] - host[0], host[512 + i] - host[0], host[768 + i] - host[0],
                host[256 + i] - host[0], host[1024 + i] - host[0]);
      fclose(f);
    }
  }
  return 0;
}

bool tma_ok_f32(const GemmTN<float>& g) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(g.At) && al(g.B) && (g.ldat & 3) == 0 && (g.ldb & 3) == 0 && g.K >= 1 && g.M < (1ll << 31) &&
         g.N < (1ll << 31) && g.K < (1ll << 31) && g.ldat * 4 < (1ll << 40) && g.ldb * 4 < (1ll << 40);
}

template <int BM, int BN, int STAGES>
int launch_dmma(const GemmTN<double>& g, int klen, cudaStream_t st) {
  using Cfg = DmmaCfg<BM, BN, STAGES>;
  CUtensorMap tmA, tmB;
  GLA_TRY(make_map(&tmA, g.At, g.K, g.M, g.ldat, BM));
  GLA_TRY(make_map(&tmB, g.B, g.K, g.N, g.ldb, BN));
  auto kern = gemm_tn_dmma_kernel<BM, BN, STAGES>;
  GLA_TRY(ensure_dyn_smem((const void*)kern, (int)(Cfg::SMEM)));
  dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.N, BN), (unsigned)g.nsplit);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && (g.ldc & 1) == 0 &&
                      ((g.split_stride & 1) == 0)) ? 1 : 0;
  GLA_CUDA(launch_pdl(kern, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM, st, tmA, tmB, g.C, g.ldc, (int)g.M, (int)g.N, (int)g.K,
                      klen, g.split_stride, g.alpha, g.nsplit > 1 ? 0 : g.beta_one, g.lower_only, vec_ok));
  return 0;
}

bool tma_ok(const GemmTN<double>& g) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(g.At) && al(g.B) && (g.ldat & 1) == 0 && (g.ldb & 1) == 0 && g.K >= 1 && g.M < (1ll << 31) &&
         g.N < (1ll << 31) && g.K < (1ll << 31) && g.ldat * 8 < (1ll << 40) && g.ldb * 8 < (1ll << 40);
}

}  // namespace

int choose_nsplit(i64 M, i64 N, i64 K, int bm, int bn) {
  const i64 tiles = (i64)ceil_div(M, bm) * ceil_div(N, bn);
  const i64 slots = 2 * (i64)sm_count();  // resident CTAs of the DMMA kernel
  static const i64 mink = [] { const char* e = getenv("GLA_GEMM_MINK"); return e ? (i64)atoi(e) : 64ll; }();   // shortest K slice (64: n = 1024 4.04 -> 3.71 ms against 256; larger sizes unchanged)
  const i64 max_by_k = K / mink > 0 ? K / mink : 1;
  if (tiles < slots) {  // under-filled grid: cut K until the machine is full
    i64 ns = (slots + tiles - 1) / tiles;
    if (ns > max_by_k) ns = max_by_k;
    if (ns > 64) ns = 64;
    return (int)(ns < 1 ? 1 : ns);
  }
  // a few waves: pick the split whose last wave is fullest (each extra slice costs one more pass over the
  // partial outputs, charged as 1.5 % per slice)
  if (tiles >= 8 * slots) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= 8 && ns <= max_by_k; ++ns) {
    const i64 ctas = tiles * ns;
    const double eff = (double)ctas / (double)((ctas + slots - 1) / slots * slots) - 0.015 * (ns - 1);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = ns;
    }
  }
  return best;
}

// split count for the PERSISTENT tcgen05 kernel (one CTA per SM, 128 x 128 tiles): the number of (tile, slice) work items
// should be a multiple of the SM count, or large against it (the static round-robin has no other tail balancing)
int choose_nsplit_persistent(i64 M, i64 N, i64 K, int max_split) {
  const i64 tiles = (i64)ceil_div(M, 128) * ceil_div(N, 128);
  const i64 slots = sm_count();
  i64 max_by_k = K / 1024 > 0 ? K / 1024 : 1;
  if (max_by_k > max_split) max_by_k = max_split;
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= max_by_k; ++ns) {
    const i64 items = tiles * ns;
    const double eff = (double)items / (double)((items + slots - 1) / slots * slots) - 0.01 * (ns - 1);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = ns;
    }
  }
  return best;
}

template <class T>
static int launch_fma(const GemmTN<T>& g, int klen, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(g.M, 64), (unsigned)ceil_div(g.N, 64), (unsigned)g.nsplit);
  gemm_tn_fma_kernel<T><<<grid, 256, 0, st>>>(g.At, g.ldat, g.B, g.ldb, g.C, g.ldc, (int)g.M, (int)g.N, (int)g.K, klen,
                                              g.split_stride, g.alpha, g.nsplit > 1 ? 0 : g.beta_one, g.conj_a,
                                              g.lower_only);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

static int slice_len(i64 K, int nsplit) {
  i64 klen = (K + nsplit - 1) / nsplit;
  klen = (klen + 15) / 16 * 16;
  return (int)klen;
}

template <>
int gemm_tn<double>(const GemmTN<double>& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  const int klen = slice_len(g.K > 0 ? g.K : 1, g.nsplit);
  static const bool force_fma = getenv("GLA_DGEMM_FMA") != nullptr;   // A/B switch for profiling
  if (!force_fma && g.K > 0 && tma_ok(g)) {
    // 8 consumer warps (32x32 each) + 1 TMA warp, 4 x 24 KB stages -> two CTAs per SM
    if (g.M <= 64) {
      // a short, wide product (the leaf solves / small updates of the Cholesky recursion) would put a handful of
      // 64 x 128 tiles on a handful of SMs: 64 x 32 tiles (two warps per CTA) spread it over four times as many
      if ((i64)ceil_div(g.N, 128) * g.nsplit * 2 < sm_count()) return launch_dmma<64, 32, 4>(g, klen, st);
      return launch_dmma<64, 128, 4>(g, klen, st);
    }
    return launch_dmma<128, 64, 4>(g, klen, st);
  }
  return launch_fma<double>(g, klen, st);
}
template <>
int gemm_tn<float>(const GemmTN<float>& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  i64 klen = ((g.K > 0 ? g.K : 1) + g.nsplit - 1) / g.nsplit;
  klen = (klen + 31) / 32 * 32;   // K slabs of 32 floats: slices must not share a slab
  static const bool force_fma = getenv("GLA_SGEMM_FMA") != nullptr;   // A/B switch
  // tcgen05 path for products with at least one full 128 x 128 tile; GLA_SGEMM_MMASYNC=1 keeps the mma.sync kernel (A/B)
  static const bool no_umma = getenv("GLA_SGEMM_MMASYNC") != nullptr;
  if (!force_fma && !no_umma && g.K > 0 && tma_ok_f32(g) && g.M >= 128 && g.N >= 128)
    return launch_umma_tf32x3(g, (int)klen, st);
  if (!force_fma && g.K > 0 && tma_ok_f32(g)) {
    if (g.M <= 64) {
      if ((i64)ceil_div(g.N, 128) * g.nsplit * 2 < sm_count()) return launch_tf32x3<64, 32, 4>(g, (int)klen, st);
      return launch_tf32x3<64, 128, 4>(g, (int)klen, st);
    }
    return launch_tf32x3<128, 64, 4>(g, (int)klen, st);
  }
  return launch_fma<float>(g, (int)klen, st);
}
template <int BM, int BN, int STAGES>
static int launch_zdmma(const GemmTN<zd>& g, int klen, cudaStream_t st) {
  using Cfg = ZdmmaCfg<BM, BN, STAGES>;
  CUtensorMap tmA, tmB;   // complex K x R operand viewed as real 2K x R, ld doubled
  GLA_TRY(make_map(&tmA, reinterpret_cast<const double*>(g.At), 2 * g.K, g.M, 2 * g.ldat, BM));
  GLA_TRY(make_map(&tmB, reinterpret_cast<const double*>(g.B), 2 * g.K, g.N, 2 * g.ldb, BN));
  auto kern = gemm_tn_zdmma_kernel<BM, BN, STAGES>;
  GLA_TRY(ensure_dyn_smem((const void*)kern, (int)(Cfg::SMEM)));
  dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.N, BN), (unsigned)g.nsplit);
  GLA_CUDA(launch_pdl(kern, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM, st, tmA, tmB, g.C, g.ldc, (int)g.M, (int)g.N,
                      (int)(2 * g.K), 2 * klen, g.split_stride, g.alpha, g.nsplit > 1 ? 0 : g.beta_one, g.conj_a, g.lower_only));
  return 0;
}

template <>
int gemm_tn<zd>(const GemmTN<zd>& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  const int klen = slice_len(g.K > 0 ? g.K : 1, g.nsplit);
  static const bool use_fma = getenv("GLA_ZGEMM_FMA") != nullptr;   // A/B switch for profiling
  if (!use_fma && g.K > 0 && g.M < (1ll << 31) && g.N < (1ll << 31) && g.K < (1ll << 30) && g.ldat < (1ll << 35) &&
      g.ldb < (1ll << 35)) {
    if (g.M <= 64) return launch_zdmma<64, 64, 4>(g, klen, st);
    return launch_zdmma<128, 32, 4>(g, klen, st);
  }
  return launch_fma<zd>(g, klen, st);
}

template <class T>
int sum_splits(T* out, i64 ldo, const T* part, i64 ldp, i64 stride, int nsplit, i64 M, i64 N, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  i64 total = M * N;
  int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  GLA_CUDA(launch_pdl(sum_splits_kernel<T>, dim3((unsigned)grid), dim3(256), (size_t)0, st, out, ldo, part, ldp, stride, nsplit, M, N));
  return 0;
}
template int sum_splits<float>(float*, i64, const float*, i64, i64, int, i64, i64, cudaStream_t);
template int sum_splits<double>(double*, i64, const double*, i64, i64, int, i64, i64, cudaStream_t);
template int sum_splits<zd>(zd*, i64, const zd*, i64, i64, int, i64, i64, cudaStream_t);

}  // namespace gla
