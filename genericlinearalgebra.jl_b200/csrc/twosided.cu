// twosided.cu -- two-sided Householder reductions, the step after QR in the reference's own SVD / eigen
// pipelines:
//   bidiagonalize!(A)              src/svd.jl:328-381          (left + right reflectorApply!, src/qr.jl:19-42)
//   _hessenberg!(A)                src/eigenGeneral.jl:18-31   (lmul!(H', .) + rmul!(., H), src/householder.jl:43-79)
//   symtriLower! / symtriUpper!    src/eigenSelfAdjoint.jl:450-564
//
// These are BLAS-2 reductions: every step streams the whole trailing matrix (HBM / L2 bound, 1 flop per 4 bytes), and the
// steps form a chain of length n.  B200 shape of the solution: ONE persistent kernel per reduction, launched
// cooperatively with one CTA of 512 threads per SM, the matrix resident in the 126 MB L2 for n <= ~3500 (f64); the
// steps are separated by a light grid barrier (one atomic + acquire spin per CTA, ~1 us) instead of ~5 kernel launches.
// Per step only TWO barriers are needed (three phases would need the reflector vector to travel between CTAs):
//   * every CTA recomputes the reflector of the step redundantly from the (read-only in this phase) column / row into
//     its own shared memory -- identical code on identical data, so all CTAs hold bitwise the same v, tau, nu;
//   * the in-place store of the scaled reflector is deferred to the NEXT phase (after the barrier nobody reads the
//     unscaled column any more), each CTA storing a slice of it from its private copy;
//   * left application (column dots): a warp owns a column -- no cross-CTA reduction;
//   * right application (row dots): a CTA owns a block of RB rows over ALL columns (RB x 512/RB thread layout,
//     coalesced along the rows), so y = A v is complete inside the CTA -- no partial sums in global memory.
// Vectors longer than the shared-memory budget live in a per-CTA private slab of global memory (same code path).
// The wide case of bidiagonalize! and the upper case of symtri! run the same kernels on A^H / on J A J (J = index
// reversal): that is exactly the arithmetic of the reference's second code path (see the drivers below).
#include <stdlib.h>

#include "gla_internal.cuh"
#include "smallqr.cuh"

namespace gla {
namespace {

constexpr int TS_THREADS = 512;
constexpr int TS_WARPS = TS_THREADS / 32;
constexpr int TS_RED = 2 * TS_WARPS;   // reduction scratch entries

// ------------------------------------------------------------------ grid barrier
struct GridBar {
  unsigned long long* count;   // monotone arrival counter (zeroed by the host before the launch)
  int* err;                    // set when a barrier timed out (cannot happen under a cooperative launch; never hang)
  unsigned long long target;
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// returns false (CTA-uniform) when the grid is broken; the caller leaves the kernel
__device__ __forceinline__ bool grid_barrier(GridBar& b, int* flag_sm) {
  __syncthreads();
  if (threadIdx.x == 0) {
    b.target += gridDim.x;
    __threadfence();
    atomicAdd(b.count, 1ull);
    const long long t0 = clock64();
    int ok = 1;
    while (ld_acquire_u64(b.count) < b.target) {
      if (ld_volatile_i32(b.err) != 0 || clock64() - t0 > (1ll << 32)) {
        atomicExch(b.err, 1);
        ok = 0;
        break;
      }
    }
    __threadfence();
    *flag_sm = ok;
  }
  __syncthreads();
  return *flag_sm != 0;
}

// ------------------------------------------------------------------ CTA reductions (every thread gets the result)
template <class V>
__device__ __forceinline__ V cta_sum(V v, V* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  V s = red[0];
#pragma unroll
  for (int w = 1; w < TS_WARPS; ++w) s = s + red[w];
  __syncthreads();
  return s;
}
template <class R>
__device__ __forceinline__ R cta_max(R v, R* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  R s = red[0];
#pragma unroll
  for (int w = 1; w < TS_WARPS; ++w) s = fmax(s, red[w]);
  __syncthreads();
  return s;
}

// ------------------------------------------------------------------ reflector, recomputed by every CTA
template <class T>
struct Refl {
  T tau;
  typename Sc<T>::real nu;
  bool nonzero;
};

// v <- reflector!(x) with v[0] = 1 (x = `len` elements at stride `inc`, conjugated first when conj_in); x itself is
// NOT written (see cta_writeback).  Scaled norm as in Julia's norm(x) when the plain sum of squares leaves the safe range.
template <class T>
__device__ Refl<T> cta_reflector(const T* x, i64 len, i64 inc, bool conj_in, T* v, T* red) {
  using R = typename Sc<T>::real;
  R* rred = reinterpret_cast<R*>(red);
  R part = R(0);
  for (i64 k = threadIdx.x; k < len; k += TS_THREADS) {
    T e = x[k * inc];
    if (conj_in) e = cj(e);
    v[k] = e;
    part += abs2(e);
  }
  R n2 = cta_sum<R>(part, rred);
  T alpha = v[0];
  R up = R(1);
  if (!(n2 >= SafeRange<R>::lo() && n2 <= SafeRange<R>::hi())) {
    R amax = R(0);
    for (i64 k = threadIdx.x; k < len; k += TS_THREADS) amax = fmax(amax, absmax_part(v[k]));
    amax = cta_max<R>(amax, rred);
    if (amax > R(0) && amax <= SafeRange<R>::fmax()) {
      int e;
      (void)frexp(amax, &e);
      const R sc = ldexp(R(1), -e);
      up = ldexp(R(1), e);
      R p2 = R(0);
      for (i64 k = threadIdx.x; k < len; k += TS_THREADS) p2 += abs2(scale_real(v[k], sc));
      n2 = cta_sum<R>(p2, rred);
      alpha = scale_real(alpha, sc);
    }
  }
  ReflScalars<T> rs = reflector_scalars<T>(alpha, n2);
  if (up != R(1)) {
    rs.nu *= up;
    rs.ixi = scale_real(rs.ixi, R(1) / up);
  }
  Refl<T> r;
  r.tau = rs.tau;
  r.nu = rs.nu;
  r.nonzero = rs.nonzero;
  __syncthreads();   // every thread has read alpha = v[0] before thread 0 overwrites it with the implicit 1 (racecheck finding)
  if (r.nonzero) {
    for (i64 k = threadIdx.x; k < len; k += TS_THREADS) v[k] = k == 0 ? Sc<T>::one() : v[k] * rs.ixi;
  }
  __syncthreads();
  return r;
}

// the deferred in-place store: this CTA's slice of x <- [-nu; v[1:]]  (a zero column keeps its -- conjugated -- entries)
template <class T>
__device__ void cta_writeback(T* x, i64 len, i64 inc, const T* v, const Refl<T>& r) {
  for (i64 k = (i64)blockIdx.x * TS_THREADS + threadIdx.x; k < len; k += (i64)gridDim.x * TS_THREADS) {
    T e = v[k];
    if (k == 0 && r.nonzero) e = Sc<T>::from_real(-r.nu);
    x[k * inc] = e;
  }
}

// ------------------------------------------------------------------ memory-level parallelism
// Every phase is a latency chain on L2 (~0.7 us per dependent round trip, 16 warps per SM): each lane keeps TS_U
// independent loads in flight, and a column is shared by `wpc` warps as soon as there are fewer columns than warps.
template <class T>
struct TsU {
  static constexpr int value = sizeof(T) == 16 ? 4 : 8;
};

// warps per column (power of two <= TS_WARPS): all warps of the grid stay busy, >= 64 rows per warp
__device__ __forceinline__ int warps_per_column(i64 rows, i64 cols) {
  const i64 nwarps = (i64)gridDim.x * TS_WARPS;
  int wpc = 1;
  while (wpc < TS_WARPS && cols * wpc * 2 <= nwarps && rows >= 128 * (i64)wpc) wpc <<= 1;
  return wpc;
}

// sum over the `wpc` warps that share a column (fixed order); every thread of the CTA must call
template <class T>
__device__ __forceinline__ T group_sum(T s, int wpc, int cslot, T* red) {
  s = warp_sum(s);
  if (wpc == 1) return s;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  T t = red[cslot * wpc];
  for (int w = 1; w < wpc; ++w) t = t + red[cslot * wpc + w];
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------ A <- (I - conj(tau) v v^H) A   (stdlib reflectorApply!)
template <class T>
__device__ void left_apply(T* A, i64 lda, i64 rows, i64 cols, const T* v, T tau, T* red) {
  constexpr int U = TsU<T>::value;
  const T ctau = cj(tau);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = warps_per_column(rows, cols);
  const int cpc = TS_WARPS / wpc, cslot = warp / wpc;
  const int tig = (warp % wpc) * 32 + lane, gs = wpc * 32;
  const bool single = rows <= (i64)gs * U;
  for (i64 base = (i64)blockIdx.x * cpc; base < cols; base += (i64)gridDim.x * cpc) {
    const i64 j = base + cslot;
    const bool ok = j < cols;
    T* a = A + (ok ? j : 0) * lda;
    T s0 = Sc<T>::zero(), s1 = Sc<T>::zero();
    T xs[U];
    if (ok && single) {   // one batch covers the column: keep it in registers for the update
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 k = tig + (i64)u * gs;
        xs[u] = k < rows ? a[k] : Sc<T>::zero();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 k = tig + (i64)u * gs;
        if (k < rows) {
          if (u & 1) s1 = fmad(cj(v[k]), xs[u], s1);
          else s0 = fmad(cj(v[k]), xs[u], s0);
        }
      }
    } else if (ok) {
      for (i64 k0 = tig; k0 < rows; k0 += (i64)gs * U) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          x[u] = k < rows ? a[k] : Sc<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          if (k < rows) {
            if (u & 1) s1 = fmad(cj(v[k]), x[u], s1);
            else s0 = fmad(cj(v[k]), x[u], s0);
          }
        }
      }
    }
    T s = group_sum<T>(s0 + s1, wpc, cslot, red);
    const T ms = -(ctau * s);
    if (ok && single) {   // the column is still in registers
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 k = tig + (i64)u * gs;
        if (k < rows) a[k] = fmad(v[k], ms, xs[u]);
      }
    } else if (ok) {
      for (i64 k0 = tig; k0 < rows; k0 += (i64)gs * U) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          x[u] = k < rows ? a[k] : Sc<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          if (k < rows) a[k] = fmad(v[k], ms, x[u]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ A <- A (I - tau v v^H)   (src/qr.jl:19-42, rmul!)
template <class T, int RB>
__device__ void right_apply_rb(T* A, i64 lda, i64 rows, i64 cols, const T* v, T tau, T* ysm) {
  constexpr int NCG = TS_THREADS / RB;
  constexpr int U = TsU<T>::value;
  const int rx = threadIdx.x % RB, cgp = threadIdx.x / RB;
  const i64 nblk = (rows + RB - 1) / RB;
  const bool single = cols <= (i64)NCG * U;
  for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
    const i64 r = b * RB + rx;
    const bool ok = r < rows;
    T* a = A + (ok ? r : 0);
    T s0 = Sc<T>::zero(), s1 = Sc<T>::zero();
    T xs[U];
    if (ok && single) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = cgp + (i64)u * NCG;
        xs[u] = j < cols ? a[j * lda] : Sc<T>::zero();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = cgp + (i64)u * NCG;
        if (j < cols) {
          if (u & 1) s1 = fmad(xs[u], v[j], s1);
          else s0 = fmad(xs[u], v[j], s0);
        }
      }
    } else if (ok) {
      for (i64 j0 = cgp; j0 < cols; j0 += (i64)NCG * U) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + (i64)u * NCG;
          x[u] = j < cols ? a[j * lda] : Sc<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + (i64)u * NCG;
          if (j < cols) {
            if (u & 1) s1 = fmad(x[u], v[j], s1);
            else s0 = fmad(x[u], v[j], s0);
          }
        }
      }
    }
    ysm[cgp * RB + rx] = s0 + s1;
    __syncthreads();
    T y = ysm[rx];
#pragma unroll 4
    for (int c = 1; c < NCG; ++c) y = y + ysm[c * RB + rx];
    __syncthreads();
    const T my = -(y * tau);
    if (ok && single) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = cgp + (i64)u * NCG;
        if (j < cols) a[j * lda] = fmad(my, cj(v[j]), xs[u]);
      }
    } else if (ok) {
      for (i64 j0 = cgp; j0 < cols; j0 += (i64)NCG * U) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + (i64)u * NCG;
          x[u] = j < cols ? a[j * lda] : Sc<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + (i64)u * NCG;
          if (j < cols) a[j * lda] = fmad(my, cj(v[j]), x[u]);
        }
      }
    }
  }
}
// rows per CTA block: the candidate with the fewest (rounds over the grid) x (columns per thread); ties -> wider rows
__device__ __forceinline__ int pick_row_block(i64 rows) {
  const i64 g = gridDim.x;
  int best = 32;
  i64 cost = ((rows + 31) / 32 + g - 1) / g * 32;
#pragma unroll
  for (int rb = 16; rb >= 4; rb >>= 1) {
    const i64 c = ((rows + rb - 1) / rb + g - 1) / g * rb;
    if (c < cost) {
      cost = c;
      best = rb;
    }
  }
  return best;
}
template <class T>
__device__ void right_apply(T* A, i64 lda, i64 rows, i64 cols, const T* v, T tau, T* ysm) {
  switch (pick_row_block(rows)) {
    case 32: right_apply_rb<T, 32>(A, lda, rows, cols, v, tau, ysm); break;
    case 16: right_apply_rb<T, 16>(A, lda, rows, cols, v, tau, ysm); break;
    case 8: right_apply_rb<T, 8>(A, lda, rows, cols, v, tau, ysm); break;
    default: right_apply_rb<T, 4>(A, lda, rows, cols, v, tau, ysm); break;
  }
}

// ------------------------------------------------------------------ u = Hermitian(At, :L) v in two partial vectors
// ucol[j] = re(At[j,j]) v_j + sum_{i>j} conj(At[i,j]) v_i   (`wpc` warps per column)
// urow[r] = sum_{j<r} At[r,j] v_j                            (a CTA per block of RB rows)
template <class T, int RB>
__device__ void symv_lower_rows(const T* At, i64 lda, i64 L, const T* v, T* urow, T* ysm) {
  constexpr int NCG = TS_THREADS / RB;
  constexpr int U = TsU<T>::value;
  const int rx = threadIdx.x % RB, cgp = threadIdx.x / RB;
  const i64 nblk = (L + RB - 1) / RB;
  // blocks are dealt from the bottom (longest rows first) so that the tail of the phase is made of short ones
  for (i64 bb = blockIdx.x; bb < nblk; bb += gridDim.x) {
    const i64 b = nblk - 1 - bb;
    const i64 r = b * RB + rx;
    const bool ok = r < L;
    const T* a = At + (ok ? r : 0);
    const i64 jend = ok ? r : 0;   // columns j < r
    T s0 = Sc<T>::zero(), s1 = Sc<T>::zero();
    for (i64 j0 = cgp; j0 < jend; j0 += (i64)NCG * U) {
      T x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + (i64)u * NCG;
        x[u] = j < jend ? a[j * lda] : Sc<T>::zero();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + (i64)u * NCG;
        if (j < jend) {
          if (u & 1) s1 = fmad(x[u], v[j], s1);
          else s0 = fmad(x[u], v[j], s0);
        }
      }
    }
    ysm[cgp * RB + rx] = s0 + s1;
    __syncthreads();
    if (cgp == 0 && ok) {
      T y = ysm[rx];
      for (int c = 1; c < NCG; ++c) y = y + ysm[c * RB + rx];
      urow[r] = y;
    }
    __syncthreads();
  }
}
template <class T>
__device__ void symv_lower(const T* At, i64 lda, i64 L, const T* v, T* ucol, T* urow, T* ysm, T* red) {
  constexpr int U = TsU<T>::value;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = warps_per_column(L, L);
  const int cpc = TS_WARPS / wpc, cslot = warp / wpc;
  const int tig = (warp % wpc) * 32 + lane, gs = wpc * 32;
  for (i64 base = (i64)blockIdx.x * cpc; base < L; base += (i64)gridDim.x * cpc) {
    const i64 j = base + cslot;
    const bool ok = j < L;
    const T* a = At + (ok ? j : 0) * lda;
    T s0 = Sc<T>::zero(), s1 = Sc<T>::zero();
    if (ok) {
      for (i64 k0 = j + 1 + tig; k0 < L; k0 += (i64)gs * U) {
        T x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          x[u] = k < L ? a[k] : Sc<T>::zero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 k = k0 + (i64)u * gs;
          if (k < L) {
            if (u & 1) s1 = fmad(cj(x[u]), v[k], s1);
            else s0 = fmad(cj(x[u]), v[k], s0);
          }
        }
      }
    }
    const T s = group_sum<T>(s0 + s1, wpc, cslot, red);
    if (ok && tig == 0) ucol[j] = s + scale_real(v[j], re(a[j]));
  }
  switch (pick_row_block(L)) {
    case 32: symv_lower_rows<T, 32>(At, lda, L, v, urow, ysm); break;
    case 16: symv_lower_rows<T, 16>(At, lda, L, v, urow, ysm); break;
    case 8: symv_lower_rows<T, 8>(At, lda, L, v, urow, ysm); break;
    default: symv_lower_rows<T, 4>(At, lda, L, v, urow, ysm); break;
  }
}

// At[i,j] += v_i conj(xi v_j) - v_i conj(u_j) - u_i conj(v_j), i >= j     (src/eigenSelfAdjoint.jl:483-490)
template <class T>
__device__ void rank2_lower(T* At, i64 lda, i64 L, const T* v, const T* u, typename Sc<T>::real xi) {
  constexpr int U = TsU<T>::value;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpc = warps_per_column(L, L);
  const int cpc = TS_WARPS / wpc, cslot = warp / wpc;
  const int tig = (warp % wpc) * 32 + lane, gs = wpc * 32;
  for (i64 base = (i64)blockIdx.x * cpc; base < L; base += (i64)gridDim.x * cpc) {
    const i64 j = base + cslot;
    if (j >= L) continue;
    T* a = At + j * lda;
    const T vj = v[j], uj = u[j];
    const T t1 = cj(scale_real(vj, xi)) - cj(uj);   // multiplies v_i
    const T t2 = -cj(vj);                            // multiplies u_i
    for (i64 k0 = j + tig; k0 < L; k0 += (i64)gs * U) {
      T x[U];
#pragma unroll
      for (int u_ = 0; u_ < U; ++u_) {
        const i64 k = k0 + (i64)u_ * gs;
        x[u_] = k < L ? a[k] : Sc<T>::zero();
      }
#pragma unroll
      for (int u_ = 0; u_ < U; ++u_) {
        const i64 k = k0 + (i64)u_ * gs;
        if (k < L) {
          T e = fmad(v[k], t1, x[u_]);
          e = fmad(u[k], t2, e);
          if (k == j) e = Sc<T>::from_real(re(e));
          a[k] = e;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the persistent kernels
template <class T>
struct TsArgs {
  T* A;
  i64 lda, m, n;
  T* tau1;       // taul (bidiagonalize), tau (hessenberg, symtri)
  T* tau2;       // taur
  T* gvec;       // per-CTA private vectors when they do not fit shared memory: gridDim.x * nvec * veclen
  i64 veclen;    // max(m, n)
  int vec_in_smem;
  T* ucol;       // symtri: the two partial vectors (n each)
  T* urow;
  unsigned long long* bar;
  int* err;
};

struct TsSmemHead {
  int flag;
  int pad[3];
};

template <class T>
__device__ __forceinline__ void ts_carve(const TsArgs<T>& a, unsigned char* smem, int nvec, int** flag, T** red, T** ysm, T** v0,
                                         T** v1) {
  *flag = &reinterpret_cast<TsSmemHead*>(smem)->flag;
  T* p = reinterpret_cast<T*>(smem + sizeof(TsSmemHead));
  *red = p;
  p += TS_RED;
  *ysm = p;
  p += TS_THREADS;
  if (a.vec_in_smem) {
    *v0 = p;
    *v1 = p + a.veclen;
  } else {
    *v0 = a.gvec + (i64)blockIdx.x * nvec * a.veclen;
    *v1 = *v0 + a.veclen;
  }
}

// m >= n: upper bidiagonal (src/svd.jl:334-346)
template <class T>
__global__ void __launch_bounds__(TS_THREADS, 1) bidiag_tall_kernel(TsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  int* flag;
  T *red, *ysm, *v, *unused;
  ts_carve<T>(a, ts_smem, 1, &flag, &red, &ysm, &v, &unused);
  GridBar bar{a.bar, a.err, 0ull};
  const i64 m = a.m, n = a.n, lda = a.lda;
  T* A = a.A;
  Refl<T> rrow;
  rrow.nonzero = false;
  bool have_row = false;
  for (i64 i = 0; i < n; ++i) {
    // ---- phase A: (store the row reflector of step i-1), column reflector, left application
    if (have_row) {
      cta_writeback<T>(A + (i - 1) + i * lda, n - i, lda, v, rrow);
      __syncthreads();
    }
    const Refl<T> rc = cta_reflector<T>(A + i + i * lda, m - i, 1, false, v, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.tau1[i] = rc.tau;
    if (rc.nonzero && i + 1 < n) left_apply<T>(A + i + (i + 1) * lda, lda, m - i, n - i - 1, v, rc.tau, red);
    if (!grid_barrier(bar, flag)) return;
    // ---- phase B: store the column reflector, row reflector, right application
    cta_writeback<T>(A + i + i * lda, m - i, 1, v, rc);
    __syncthreads();
    have_row = false;
    if (i + 1 < n) {
      rrow = cta_reflector<T>(A + i + (i + 1) * lda, n - i - 1, lda, true, v, red);
      have_row = true;
      if (blockIdx.x == 0 && threadIdx.x == 0) a.tau2[i] = rrow.tau;
      if (rrow.nonzero && i + 1 < m)
        right_apply<T>(A + (i + 1) + (i + 1) * lda, lda, m - i - 1, n - i - 1, v, rrow.tau, ysm);
      if (!grid_barrier(bar, flag)) return;
    }
  }
}

// src/eigenGeneral.jl:18-31
template <class T>
__global__ void __launch_bounds__(TS_THREADS, 1) hessenberg_kernel(TsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  int* flag;
  T *red, *ysm, *v, *unused;
  ts_carve<T>(a, ts_smem, 1, &flag, &red, &ysm, &v, &unused);
  GridBar bar{a.bar, a.err, 0ull};
  const i64 n = a.n, lda = a.lda;
  T* A = a.A;
  for (i64 i = 0; i + 1 < n; ++i) {
    const i64 len = n - i - 1;
    T* x = A + (i + 1) + i * lda;
    const Refl<T> rc = cta_reflector<T>(x, len, 1, false, v, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.tau1[i] = rc.tau;
    if (rc.nonzero) left_apply<T>(A + (i + 1) + (i + 1) * lda, lda, len, len, v, rc.tau, red);
    if (!grid_barrier(bar, flag)) return;
    cta_writeback<T>(x, len, 1, v, rc);
    if (rc.nonzero) right_apply<T>(A + (i + 1) * lda, lda, n, len, v, rc.tau, ysm);
    if (!grid_barrier(bar, flag)) return;
  }
}

// symtriLower!  src/eigenSelfAdjoint.jl:450-503
template <class T>
__global__ void __launch_bounds__(TS_THREADS, 1) symtri_lower_kernel(TsArgs<T> a) {
  using R = typename Sc<T>::real;
  extern __shared__ __align__(16) unsigned char ts_smem[];
  int* flag;
  T *red, *ysm, *v, *u;
  ts_carve<T>(a, ts_smem, 2, &flag, &red, &ysm, &v, &u);
  GridBar bar{a.bar, a.err, 0ull};
  const i64 n = a.n, lda = a.lda;
  T* A = a.A;
  if (Sc<T>::is_complex) {   // the imaginary parts of the diagonal are ignored (:458-460)
    for (i64 k = (i64)blockIdx.x * TS_THREADS + threadIdx.x; k < n; k += (i64)gridDim.x * TS_THREADS)
      A[k + k * lda] = Sc<T>::from_real(re(A[k + k * lda]));
    if (!grid_barrier(bar, flag)) return;
  }
  const i64 steps = n - 2 + (Sc<T>::is_complex ? 1 : 0);
  if (!Sc<T>::is_complex && n >= 2 && blockIdx.x == 0 && threadIdx.x == 0) a.tau1[n - 2] = Sc<T>::zero();
  for (i64 k = 0; k < steps; ++k) {
    const i64 L = n - k - 1;
    T* x = A + (k + 1) + k * lda;
    T* At = A + (k + 1) + (k + 1) * lda;
    const Refl<T> rc = cta_reflector<T>(x, L, 1, false, v, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.tau1[k] = rc.tau;
    if (rc.nonzero) symv_lower<T>(At, lda, L, v, a.ucol, a.urow, ysm, red);
    if (!grid_barrier(bar, flag)) return;
    cta_writeback<T>(x, L, 1, v, rc);
    if (rc.nonzero) {
      T part = Sc<T>::zero();
      for (i64 i = threadIdx.x; i < L; i += TS_THREADS) {
        const T ui = rc.tau * (a.ucol[i] + (i > 0 ? a.urow[i] : Sc<T>::zero()));
        u[i] = ui;
        part = fmad(cj(v[i]), ui, part);
      }
      const T dot = cta_sum<T>(part, red);   // its barriers also publish u
      const R xi = re(cj(rc.tau) * dot);
      rank2_lower<T>(At, lda, L, v, u, xi);
    }
    if (!grid_barrier(bar, flag)) return;
  }
}

// ------------------------------------------------------------------ layout helpers of the second code paths
// dst (n x m) = src^H (src m x n), or the plain transpose
template <class T>
__global__ void transpose_kernel(const T* __restrict__ src, i64 lds, i64 m, i64 n, T* __restrict__ dst, i64 ldd, int conj) {
  __shared__ T tile[32][33];
  const i64 i0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const i64 i = i0 + threadIdx.x, j = j0 + jj;
    if (i < m && j < n) tile[jj][threadIdx.x] = src[i + j * lds];
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const i64 j = j0 + threadIdx.x, i = i0 + ii;
    if (i < m && j < n) {
      const T e = tile[threadIdx.x][ii];
      dst[j + i * ldd] = conj ? cj(e) : e;
    }
  }
}
// lower triangle of dst <- upper triangle of src under the index reversal (dst[i,j] = src[n-1-i, n-1-j], i >= j), or back
template <class T>
__global__ void flip_triangle_kernel(const T* __restrict__ src, i64 lds, i64 n, T* __restrict__ dst, i64 ldd, int to_lower) {
  const i64 total = n * n;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 j = e / n, i = e - j * n;
    if (to_lower) {
      if (i >= j) dst[i + j * ldd] = src[(n - 1 - i) + (n - 1 - j) * lds];
    } else {
      if (i <= j) dst[i + j * ldd] = src[(n - 1 - i) + (n - 1 - j) * lds];
    }
  }
}

// ------------------------------------------------------------------ launch plumbing
template <class T>
struct TsWork {
  void* small = nullptr;   // barrier counter + error flag
  T* gvec = nullptr;
  T* uu = nullptr;
  cudaStream_t st = nullptr;
  void release() {
    if (small) cudaFreeAsync(small, st);
    if (gvec) cudaFreeAsync(gvec, st);
    if (uu) cudaFreeAsync(uu, st);
    small = nullptr;
    gvec = nullptr;
    uu = nullptr;
  }
};

constexpr int TS_SMEM_BUDGET = 200 * 1024;

template <class T>
int ts_prepare(TsArgs<T>& a, TsWork<T>& w, int nvec, bool need_u, int* smem_bytes, int* grid, const void* func, cudaStream_t st) {
  w.st = st;
  a.veclen = a.m > a.n ? a.m : a.n;
  const i64 fixed = (i64)sizeof(TsSmemHead) + (i64)(TS_RED + TS_THREADS) * sizeof(T);
  const i64 vec_bytes = (i64)nvec * a.veclen * sizeof(T);
  // GLA_TS_SMEM_BUDGET (bytes) shrinks the budget: tests use it to drive the global-slab path at small sizes
  const char* env = getenv("GLA_TS_SMEM_BUDGET");
  const i64 budget = env ? (i64)atoll(env) : (i64)TS_SMEM_BUDGET;
  a.vec_in_smem = fixed + vec_bytes <= (budget < TS_SMEM_BUDGET ? budget : (i64)TS_SMEM_BUDGET);
  *smem_bytes = (int)(fixed + (a.vec_in_smem ? vec_bytes : 0));
  GLA_TRY(ensure_dyn_smem(func, *smem_bytes));
  int per_sm = 0;
  GLA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, TS_THREADS, *smem_bytes));
  if (per_sm < 1) {
    set_error(GLA_ERR_INTERNAL, "two-sided kernel does not fit an SM", __FILE__, __LINE__);
    return GLA_ERR_INTERNAL;
  }
  *grid = sm_count();   // one CTA per SM: co-resident, as the cooperative launch requires
  GLA_TRY(pool_malloc(&w.small, 16, st));
  GLA_CUDA(cudaMemsetAsync(w.small, 0, 16, st));
  a.bar = static_cast<unsigned long long*>(w.small);
  a.err = reinterpret_cast<int*>(static_cast<unsigned char*>(w.small) + 8);
  a.gvec = nullptr;
  if (!a.vec_in_smem) {
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&w.gvec), (size_t)*grid * nvec * a.veclen * sizeof(T), st));
    a.gvec = w.gvec;
  }
  a.ucol = a.urow = nullptr;
  if (need_u) {
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&w.uu), (size_t)2 * a.veclen * sizeof(T), st));
    a.ucol = w.uu;
    a.urow = w.uu + a.veclen;
  }
  return 0;
}

template <class T>
int ts_launch(const void* func, TsArgs<T>& a, int grid, int smem_bytes, cudaStream_t st) {
  void* params[] = {&a};
  GLA_CUDA(cudaLaunchCooperativeKernel(func, dim3((unsigned)grid), dim3(TS_THREADS), params, (size_t)smem_bytes, st));
  return 0;
}

template <class T>
int bidiag_tall(T* dA, i64 m, i64 n, i64 lda, T* dtaul, T* dtaur, cudaStream_t st) {
  TsArgs<T> a{};
  a.A = dA;
  a.lda = lda;
  a.m = m;
  a.n = n;
  a.tau1 = dtaul;
  a.tau2 = dtaur;
  TsWork<T> w;
  int smem = 0, grid = 0;
  const void* f = reinterpret_cast<const void*>(&bidiag_tall_kernel<T>);
  int rc = ts_prepare<T>(a, w, 1, false, &smem, &grid, f, st);
  if (!rc) rc = ts_launch<T>(f, a, grid, smem, st);
  w.release();
  return rc;
}

}  // namespace

// bidiagonalize!(A): m >= n runs in place; m < n is the same reduction of A^H (the reference's second branch,
// src/svd.jl:358-371: the row reflector of conj(row i) IS the column reflector of A^H, a right application on A is the left
// application on A^H, and the stored reflectors of the two layouts are plain transposes of each other).
template <class T>
int bidiagonalize_dev(T* dA, i64 m, i64 n, i64 lda, T* dtaul, T* dtaur, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (m == 0 || n == 0) return 0;
  if (!dA) return -1;
  if (!dtaul && (m >= n ? n : m - 1) > 0) return -5;
  if (!dtaur && (m >= n ? n - 1 : m) > 0) return -6;
  if (m >= n) return bidiag_tall<T>(dA, m, n, lda, dtaul, dtaur, st);
  T* C = nullptr;
  const i64 ldc = round_up(n, 2);
  GLA_TRY(pool_malloc(reinterpret_cast<void**>(&C), (size_t)ldc * m * sizeof(T), st));
  const dim3 tb(32, 8), tg((unsigned)ceil_div(m, 32), (unsigned)ceil_div(n, 32));
  transpose_kernel<T><<<tg, tb, 0, st>>>(dA, lda, m, n, C, ldc, 1);
  int rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  if (!rc) rc = bidiag_tall<T>(C, n, m, ldc, dtaur, dtaul, st);
  if (!rc) {
    const dim3 tg2((unsigned)ceil_div(n, 32), (unsigned)ceil_div(m, 32));
    transpose_kernel<T><<<tg2, tb, 0, st>>>(C, ldc, n, m, dA, lda, 0);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  cudaFreeAsync(C, st);
  return rc;
}

template <class T>
int hessenberg_dev(T* dA, i64 n, i64 lda, T* dtau, cudaStream_t st) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n <= 1) return 0;
  if (!dA) return -1;
  if (!dtau) return -4;
  TsArgs<T> a{};
  a.A = dA;
  a.lda = lda;
  a.m = n;
  a.n = n;
  a.tau1 = dtau;
  TsWork<T> w;
  int smem = 0, grid = 0;
  const void* f = reinterpret_cast<const void*>(&hessenberg_kernel<T>);
  int rc = ts_prepare<T>(a, w, 1, false, &smem, &grid, f, st);
  if (!rc) rc = ts_launch<T>(f, a, grid, smem, st);
  w.release();
  return rc;
}

// symtri!(Hermitian(A, uplo)): the lower variant in place; the upper variant (src/eigenSelfAdjoint.jl:505-564 reflects
// about the LAST element of each column: `reflector!` on the reversed view) is the lower variant of J A J, J the index
// reversal -- element for element the same arithmetic -- run on a flipped copy of the upper triangle.
template <class T>
int symtri_dev(T* dA, i64 n, i64 lda, int upper, T* dtau, cudaStream_t st) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  if (!dA) return -1;
  if (!dtau && n > 1) return -5;
  T* B = nullptr;
  i64 ldb = lda;
  T* W = dA;
  int rc = 0;
  const unsigned fgrid = (unsigned)(ceil_div(n * n, 256) > 4096 ? 4096 : ceil_div(n * n, 256));
  if (upper) {
    ldb = round_up(n, 2);
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&B), (size_t)ldb * n * sizeof(T), st));
    flip_triangle_kernel<T><<<fgrid, 256, 0, st>>>(dA, lda, n, B, ldb, 1);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
    W = B;
  }
  TsArgs<T> a{};
  a.A = W;
  a.lda = ldb;
  a.m = n;
  a.n = n;
  a.tau1 = dtau;
  TsWork<T> w;
  int smem = 0, grid = 0;
  const void* f = reinterpret_cast<const void*>(&symtri_lower_kernel<T>);
  if (!rc) rc = ts_prepare<T>(a, w, 2, true, &smem, &grid, f, st);
  if (!rc) rc = ts_launch<T>(f, a, grid, smem, st);
  w.release();
  if (upper) {
    if (!rc) {
      flip_triangle_kernel<T><<<fgrid, 256, 0, st>>>(B, ldb, n, dA, lda, 0);
      rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
    }
    cudaFreeAsync(B, st);
  }
  return rc;
}

#define INST(T)                                                                          \
  template int bidiagonalize_dev<T>(T*, i64, i64, i64, T*, T*, cudaStream_t);            \
  template int hessenberg_dev<T>(T*, i64, i64, T*, cudaStream_t);                        \
  template int symtri_dev<T>(T*, i64, i64, int, T*, cudaStream_t);
INST(float)
INST(double)
INST(zd)

}  // namespace gla
