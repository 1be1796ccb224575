// qr_blocked.cu -- K1 (panel), K2 (compact-WY T) and the driver of the blocked Householder QR.
//
// Reference semantics: qrBlocked! src/qr.jl:113-146 = per panel  qrUnblocked! (src/qr.jl:86-111,
// stdlib reflector!/reflectorApply! call sites :96/:102)  ->  T build (src/qr.jl:64-83, with the conj
// the reference omits at :72)  ->  trailing update A2 <- (I - V T^H V^H) A2 (src/householder.jl:119-157).
//
// GPU structure: panels of NB = 64 columns, grouped into outer blocks of 256 (n <= 12288) or NBO = 384 columns whose
// reflectors hit the far trailing matrix in one K = 384 pass (apply_outer).  Per panel:
//   qr_panel_kernel   cooperative, P CTAs each holding a row slab of the panel in shared memory;
//                     ONE grid-wide reduction per column: every CTA publishes the partial dots
//                     d_c = sum_{i>j} conj(a_ij) a_ic of the un-normalised pivot column with every
//                     remaining panel column (c = j gives the tail norm^2), the owner of row j
//                     publishes that row; after the barrier every CTA reduces the partials in the
//                     same fixed order (bitwise identical, deterministic), derives nu, tau, 1/xi and
//                     applies  a_ic -= a_ij * (s_c/xi),  s_c = conj(tau) (a_jc + conj(1/xi) d_c).
//                     It also emits the clean reflector block Vc (unit diagonal, zeros above) and
//                     its transpose VcT, the K-contiguous operands of the trailing contractions.
//   gemm_tn (x3)      G = Vc^H Vc (split-K), W = Vc^H A2 (split-K), A2 -= Vc (T^H W): FP64 tensor
//                     pipe fed by TMA (gemm.cu)
//   larft_finish      T = (I + diag(tau) striu(G))^-1 diag(tau) in one CTA (shared memory, recursive doubling)
//   qr_panel_cluster_kernel   panels of <= 6656 rows (Float64; Float32 14208, ComplexF64 2384): one thread-block cluster of up to 16 CTAs, the per-column
//                     exchange through distributed shared memory (st.async onto the receiver's transaction barrier)
//                     instead of L2
// Q application and thin Q (ormqr_blocked_dev, orgqr_thin_dev) reuse the same outer-block machinery in both directions.
#include "gemm.cuh"
#include "gla_internal.cuh"
#include "smallqr.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>

namespace gla {

constexpr int NB = 64;             // panel width
constexpr int SB = 16;             // sub-panel width inside the panel kernel
constexpr int PANEL_THREADS = 256;
constexpr int TPS = PANEL_THREADS / SB;  // threads per sub-panel column (one half-warp)
constexpr int PANEL_MAX_CTAS = 64;
constexpr int PANEL_PMAX = 160;          // upper bound on CTAs of one panel launch (exchange buffer sizing)
constexpr int GW_MAX = SB * NB;          // entries of one sub-panel's [G | W] block

// ---- flag-in-data exchange between the CTAs of a panel launch ("LL" protocol): every exchanged real number
// travels as one aligned 16-byte {payload, tag} store, readers spin on the same 16 bytes until the tag of the
// step they wait for shows up.  One L2 round trip per exchange, no separate barrier, no fences; the tag is
// unique per (launch epoch, step), buffers are reused at distance >= 2 steps (see the WAR argument in DESIGN.md).
__device__ __forceinline__ void ll_store(ulonglong2* p, unsigned long long bits, unsigned long long tag) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(bits), "l"(tag) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const ulonglong2* p, unsigned long long tag) {
  unsigned long long v, t;
  do {
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v), "=l"(t) : "l"(p) : "memory");
  } while (t != tag);
  return v;
}
__device__ __forceinline__ unsigned long long to_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long to_bits(float v) { return (unsigned long long)__float_as_uint(v); }
template <class R>
__device__ __forceinline__ R from_bits(unsigned long long b);
template <>
__device__ __forceinline__ double from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }
template <>
__device__ __forceinline__ float from_bits<float>(unsigned long long b) { return __uint_as_float((unsigned)b); }

template <class T>
struct LLX {  // one T = NR tagged entries
  static constexpr int NR = Sc<T>::is_complex ? 2 : 1;
  using R = typename Sc<T>::real;
  static __device__ __forceinline__ void put(ulonglong2* base, i64 idx, T v, unsigned long long tag) {
    if constexpr (Sc<T>::is_complex) {
      ll_store(base + 2 * idx, to_bits(v.x), tag);
      ll_store(base + 2 * idx + 1, to_bits(v.y), tag);
    } else {
      ll_store(base + idx, to_bits(v), tag);
    }
  }
  // sum over pp = q, q+STEP, q+2*STEP, ... < P (ascending) of the T stored at base + pp*stride*NR
  template <int STEP>
  static __device__ __forceinline__ T gather(const ulonglong2* base, int q, int P, i64 stride, unsigned long long tag) {
    constexpr int MAXE = (PANEL_PMAX + STEP - 1) / STEP;  // entries per lane
    T sum = Sc<T>::zero();
    for (int b = 0; b < MAXE; b += 4) {   // batches of 4 T (4 or 8 loads in flight)
      if (q + b * STEP >= P) break;
      unsigned long long v[4 * NR], t[4 * NR];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = q + (b + u) * STEP;
          if (pp < P) {
#pragma unroll
            for (int c = 0; c < NR; ++c) {
              const ulonglong2* ptr = base + (i64)pp * stride * NR + c;
              asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[u * NR + c]), "=l"(t[u * NR + c]) : "l"(ptr) : "memory");
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int pp = q + (b + u) * STEP;
          if (pp < P) {
#pragma unroll
            for (int c = 0; c < NR; ++c) ok = ok && (t[u * NR + c] == tag);
          }
        }
      } while (!ok);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = q + (b + u) * STEP;
        if (pp < P) {
          if constexpr (Sc<T>::is_complex) sum = sum + make_zd(from_bits<double>(v[u * 2]), from_bits<double>(v[u * 2 + 1]));
          else sum = sum + from_bits<R>(v[u]);
        }
      }
    }
    return sum;
  }
  static __device__ __forceinline__ T get(const ulonglong2* base, i64 idx, unsigned long long tag) {
    if constexpr (Sc<T>::is_complex) {
      const double x = from_bits<double>(ll_load(base + 2 * idx, tag));
      const double y = from_bits<double>(ll_load(base + 2 * idx + 1, tag));
      return make_zd(x, y);
    } else {
      return from_bits<R>(ll_load(base + idx, tag));
    }
  }
};

template <class T>
struct PanelArgs {
  T* A;        // panel origin (row k0, col k0)
  i64 lda;
  int mk, nb;  // panel rows, panel columns (<= NB)
  T* tau;      // tau + k0
  T* Vc;       // mk x kk clean reflectors (ld = ldvc)
  i64 ldvc;
  T* VcT;      // kk x mk (ld = ldvct)
  i64 ldvct;
  ulonglong2* xd;  // [2][PANEL_PMAX][SB] T   per-column partial dots
  ulonglong2* xr;  // [2][SB] T               pivot-row entries
  ulonglong2* xw;  // [PANEL_PMAX][GW_MAX] T  partial [G | W] of a sub-panel
  ulonglong2* xz;  // [2][GW_MAX] T           reduced [G | W]
  unsigned long long epoch;
  int rows_per;  // rows per CTA
  int resident;  // slab lives in shared memory
  int lds;       // slab leading dimension when resident
};

// half-warp (16 lanes) butterfly sum; the mask names only the caller's half, so the two halves of a warp may
// run different trip counts around it
__device__ __forceinline__ float hw_shfl(float v, int o, unsigned m) { return __shfl_xor_sync(m, v, o); }
__device__ __forceinline__ double hw_shfl(double v, int o, unsigned m) { return __shfl_xor_sync(m, v, o); }
__device__ __forceinline__ zd hw_shfl(zd v, int o, unsigned m) {
  return make_zd(__shfl_xor_sync(m, v.x, o), __shfl_xor_sync(m, v.y, o));
}
__device__ __forceinline__ float hw_shfl_idx(float v, int src, unsigned m) { return __shfl_sync(m, v, src, TPS); }
__device__ __forceinline__ double hw_shfl_idx(double v, int src, unsigned m) { return __shfl_sync(m, v, src, TPS); }
__device__ __forceinline__ zd hw_shfl_idx(zd v, int src, unsigned m) {
  return make_zd(__shfl_sync(m, v.x, src, TPS), __shfl_sync(m, v.y, src, TPS));
}
// broadcast from lane `src` of the caller's half-warp
template <class T>
__device__ __forceinline__ T hw_bcast(T v, int src) {
  return hw_shfl_idx(v, src, 0xffffu << (threadIdx.x & 16));
}
template <class T>
__device__ __forceinline__ T hw_sum(T v) {
  const unsigned m = 0xffffu << (threadIdx.x & 16);
#pragma unroll
  for (int o = 1; o < TPS; o <<= 1) v = v + hw_shfl(v, o, m);
  return v;
}

// K1: panel factorisation.  P CTAs, CTA p owns the row slab [p*rows_per, ...) of the panel (in shared memory).
// The NB columns are processed in sub-panels of SB columns:
//   column step j (right-looking INSIDE the sub-panel only): local partial dots of the un-normalised pivot
//     column with the remaining sub-panel columns -> LL exchange (every CTA sums the partials in the same fixed
//     order, so all CTAs hold bitwise identical scalars) -> nu, tau, 1/xi -> rank-1 update of the sub-panel;
//   block step (once per sub-panel): [G | W] = Vs^H [Vs | A_rest] partials -> reduce-scatter + all-gather over
//     the LL buffers -> T_s = (I + diag(tau) striu(G))^-1 diag(tau) -> A_rest -= Vs (T_s^H W).
// Afterwards the panel is written back together with the clean reflector block Vc (unit diagonal, zeros
// above) and its transpose VcT, the K-contiguous operands of the trailing contractions.
template <class T>
__global__ void __launch_bounds__(PANEL_THREADS, 1) qr_panel_kernel(PanelArgs<T> a) {
  using R = typename Sc<T>::real;
  using X = LLX<T>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T sh_d[SB];
  __shared__ T sh_row[SB];
  __shared__ T sh_tau[SB];
  __shared__ T sh_gw[GW_MAX];            // reduced [G | W], entry (l, x) at x*SB + l
  __shared__ T sh_z[SB * (NB - SB)];     // Z = T_s^H W, entry (l, x) at x*SB + l
  __shared__ T sh_t[SB][SB + 1];         // T_s
  const int P = gridDim.x, p = blockIdx.x;
  const int r0 = p * a.rows_per;
  const int r1 = (r0 + a.rows_per < a.mk) ? r0 + a.rows_per : a.mk;
  const int rows = r1 > r0 ? r1 - r0 : 0;
  const int tid = threadIdx.x;
  const int cl = tid / TPS, q = tid % TPS;  // sub-panel column, row phase
  const int kk = a.mk < a.nb ? a.mk : a.nb;
  const unsigned long long tagbase = a.epoch << 12;

  // slab S(i_local, col)
  T* S;
  i64 ld;
  // 2-D thread mapping of the slab copies: rw threads walk the rows of a column, 256/rw columns at a time
  const int rw = rows > 128 ? 256 : (rows > 64 ? 128 : 64);
  const int ci0 = tid % rw, ccg = tid / rw, cncg = PANEL_THREADS / rw;
  if (a.resident) {
    S = reinterpret_cast<T*>(smem_raw);
    ld = a.lds;
    // every element is one cp.async: all loads of the slab are in flight at once
    // L2-only loads (see ldcg_t): the panel columns were just rewritten by the other stream's far update
    for (int col = ccg; col < a.nb; col += cncg) {
      int i = ci0;
      for (; i + 3 * rw < rows; i += 4 * rw) {   // four independent loads in flight per thread
        const T* src = a.A + (i64)col * a.lda + r0 + i;
        const T v0 = ldcg_t(src), v1 = ldcg_t(src + rw), v2 = ldcg_t(src + 2 * rw), v3 = ldcg_t(src + 3 * rw);
        T* dst = S + i + col * ld;
        dst[0] = v0; dst[rw] = v1; dst[2 * rw] = v2; dst[3 * rw] = v3;
      }
      for (; i < rows; i += rw) S[i + col * ld] = ldcg_t(a.A + (i64)col * a.lda + r0 + i);
    }
  } else {
    S = a.A + r0;
    ld = a.lda;
  }
  __syncthreads();

  T prev_ixi = Sc<T>::one();
  R prev_nu = R(0);
  bool prev_nonzero = false;
  // deferred finish of a factored pivot column: rows below the diagonal *= 1/xi, diagonal <- -nu
  auto finish = [&](int jc) {
    if (!prev_nonzero) return;
    T* col = S + (i64)jc * ld;
    int lo = jc + 1 - r0;
    if (lo < 0) lo = 0;
    for (int i = lo + tid; i < rows; i += PANEL_THREADS) col[i] = col[i] * prev_ixi;
    if (tid == 0 && jc >= r0 && jc < r1) col[jc - r0] = Sc<T>::from_real(-prev_nu);
  };

  const int nsub = (kk + SB - 1) / SB;
  for (int sp = 0; sp < nsub; ++sp) {
    const int sb0 = sp * SB;
    const int sbe = sb0 + SB < kk ? sb0 + SB : kk;
    const int c = sb0 + cl;
    for (int j = sb0; j < sbe; ++j) {
      if (j > sb0) finish(j - 1);
      const int jl = j - sb0;
      const int par = j & 1;
      const unsigned long long tag = tagbase + 1 + j;
      const T* piv = S + (i64)j * ld;
      int lo = j + 1 - r0;  // first local row strictly below the diagonal
      if (lo < 0) lo = 0;
      const bool active = cl >= jl && c < sbe;
      // ---- partial dots with the un-normalised pivot column
      T d = Sc<T>::zero();
      if (active) {
        const T* cc = S + (i64)c * ld;
        T d1 = Sc<T>::zero();
        int i = lo + q;
        for (; i + TPS < rows; i += 2 * TPS) {
          d = fmad(cj(piv[i]), cc[i], d);
          d1 = fmad(cj(piv[i + TPS]), cc[i + TPS], d1);
        }
        if (i < rows) d = fmad(cj(piv[i]), cc[i], d);
        d = d + d1;
      }
      d = hw_sum<T>(d);
      if (P > 1) {
        if (active && q == 0) X::put(a.xd, ((i64)par * PANEL_PMAX + p) * SB + cl, d, tag);
        if (active && q == 1 && j >= r0 && j < r1) X::put(a.xr, par * SB + cl, S[(j - r0) + (i64)c * ld], tag);
        // fixed-order reduction of the partials (identical in every CTA); lane TPS-1 of the column fetches the
        // pivot-row entry meanwhile, so both exchanges cost one round trip together
        T sum = Sc<T>::zero();
        if (active) sum = X::template gather<TPS>(a.xd + ((i64)par * PANEL_PMAX * SB + cl) * X::NR, q, P, (i64)SB, tag);
        T rowv = Sc<T>::zero();
        if (active && q == TPS - 1) rowv = X::get(a.xr, par * SB + cl, tag);
        sum = hw_sum<T>(sum);
        rowv = hw_bcast<T>(rowv, TPS - 1);
        if (active && q == 0) {
          sh_d[cl] = sum;
          sh_row[cl] = rowv;
        }
      } else if (active && q == 0) {
        sh_d[cl] = d;
        sh_row[cl] = S[j + (i64)c * ld];
      }
      __syncthreads();
      const T alpha = sh_row[jl];
      const R n2 = abs2(alpha) + re(sh_d[jl]);
      ReflScalars<T> rs;   // real types: MUFU seeds + FMA refinement instead of the IEEE sqrt / division subroutines (per-column critical path)
      if constexpr (Sc<T>::is_complex) rs = reflector_scalars<T>(alpha, n2);
      else rs = reflector_scalars_fast<T>(alpha, n2);
      prev_nonzero = rs.nonzero;
      prev_ixi = rs.ixi;
      prev_nu = rs.nu;
      if (tid == 0) {
        sh_tau[jl] = rs.tau;
        if (p == 0) a.tau[j] = rs.tau;
      }
      if (rs.nonzero && active && cl > jl) {
        const T s = cj(rs.tau) * (sh_row[cl] + cj(rs.ixi) * sh_d[cl]);
        const T t = s * rs.ixi;
        T* cc = S + (i64)c * ld;
        for (int i = lo + q; i < rows; i += TPS) cc[i] = cc[i] - piv[i] * t;
        if (q == 0 && j >= r0 && j < r1) cc[j - r0] = cc[j - r0] - s;
      }
      __syncthreads();
    }
    finish(sbe - 1);
    prev_nonzero = false;
    __syncthreads();

    // ---- block step: apply the sub-panel's reflectors to the remaining panel columns [sbe, nb)
    const int ws = sbe - sb0;
    const int nrest = a.nb - sbe;
    if (nrest <= 0) continue;
    const int nx = ws + nrest;  // columns of [G | W]
    // Vs(i_local, l): unit lower trapezoid of the sub-panel
    auto vs = [&](int i, int l) -> T {
      const int gi = r0 + i, col = sb0 + l;
      return gi > col ? S[i + (i64)col * ld] : (gi == col ? Sc<T>::one() : Sc<T>::zero());
    };
    // local partials: thread -> column x = tid / 4, reflectors l = 4*(tid%4) .. +3
    {
      const int lg = (tid & 3) * 4;
      for (int x = tid >> 2; x < nx; x += PANEL_THREADS / 4) {
        T acc[4] = {Sc<T>::zero(), Sc<T>::zero(), Sc<T>::zero(), Sc<T>::zero()};
        if (x < ws) {
          for (int i = 0; i < rows; ++i) {
            const T b = vs(i, x);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmad(cj(vs(i, lg + u)), b, acc[u]);
          }
        } else {
          const T* bc = S + (i64)(sbe + x - ws) * ld;
          int isp = sbe - r0;  // local rows below isp lie strictly below the sub-panel's triangle: plain loads
          isp = isp < 0 ? 0 : (isp > rows ? rows : isp);
          for (int i = 0; i < isp; ++i) {
            const T b = bc[i];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmad(cj(vs(i, lg + u)), b, acc[u]);
          }
          const T* v0 = S + (i64)(sb0 + lg) * ld;
#pragma unroll 4
          for (int i = isp; i < rows; ++i) {
            const T b = bc[i];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmad(cj(v0[i + (i64)u * ld]), b, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (lg + u < ws) {
            if (P > 1) X::put(a.xw, (i64)p * GW_MAX + x * SB + lg + u, acc[u], tagbase + 200 + sp);
            else sh_gw[x * SB + lg + u] = acc[u];
          }
        }
      }
    }
    if (P > 1) {
      // reduce-scatter: CTA p sums entries e = p, p+P, ... over all CTAs (fixed order), publishes them
      const int zpar = sp & 1;
      for (int e0 = p + cl * P; e0 < nx * SB; e0 += SB * P) {   // entry handled by this half-warp
        const int l = e0 % SB;
        T sum = Sc<T>::zero();
        if (l < ws) sum = X::template gather<TPS>(a.xw + (i64)e0 * X::NR, q, P, (i64)GW_MAX, tagbase + 200 + sp);
        sum = hw_sum<T>(sum);
        if (q == 0) X::put(a.xz, (i64)zpar * GW_MAX + e0, sum, tagbase + 300 + sp);
      }
      // all-gather
      for (int e = tid; e < nx * SB; e += PANEL_THREADS) sh_gw[e] = X::get(a.xz, (i64)zpar * GW_MAX + e, tagbase + 300 + sp);
    }
    __syncthreads();
    // T_s = (I + diag(tau) striu(G))^-1 diag(tau), one half-warp, row i per lane
    if (tid < SB) {
      const int i = tid;
      T xrow[SB];
#pragma unroll
      for (int l = 0; l < SB; ++l) xrow[l] = (l == i) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
      for (int jj = 1; jj < SB; ++jj) {
        T acc = Sc<T>::zero();
#pragma unroll
        for (int l = 0; l < SB; ++l)
          if (l < jj && l >= i && jj < ws) acc = fmad(xrow[l], sh_tau[l] * sh_gw[jj * SB + l], acc);
        if (i < jj && jj < ws) xrow[jj] = -acc;
      }
#pragma unroll
      for (int l = 0; l < SB; ++l) sh_t[i][l] = (i <= l && l < ws && i < ws) ? xrow[l] * sh_tau[l] : Sc<T>::zero();
    }
    __syncthreads();
    // Z = T_s^H W
    for (int e = tid; e < nrest * SB; e += PANEL_THREADS) {
      const int x = e / SB, l = e % SB;
      T acc = Sc<T>::zero();
      if (l < ws)
        for (int l2 = 0; l2 <= l; ++l2) acc = fmad(cj(sh_t[l2][l]), sh_gw[(ws + x) * SB + l2], acc);
      sh_z[x * SB + l] = acc;
    }
    __syncthreads();
    // A_rest -= Vs Z on the local rows
    {
      int xs_n = PANEL_THREADS / (rows > 0 ? rows : 1);
      if (xs_n < 1) xs_n = 1;
      if (xs_n > nrest) xs_n = nrest;
      for (int w = tid; w < rows * xs_n; w += PANEL_THREADS) {
        const int i = w % rows, xs = w / rows;
        const int xb = (int)((i64)nrest * xs / xs_n), xe = (int)((i64)nrest * (xs + 1) / xs_n);
        T v[SB];
#pragma unroll
        for (int l = 0; l < SB; ++l) v[l] = l < ws ? vs(i, l) : Sc<T>::zero();
        for (int x = xb; x < xe; ++x) {
          T* pa = S + i + (i64)(sbe + x) * ld;
          T acc = *pa;
#pragma unroll
          for (int l = 0; l < SB; ++l) acc = acc - v[l] * sh_z[x * SB + l];
          *pa = acc;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();

  // ---- write back: panel (if staged), clean reflectors Vc and VcT
  for (int col = ccg; col < a.nb; col += cncg)
    for (int i = ci0; i < rows; i += rw) {
      const T x = S[i + (i64)col * ld];
      if (a.resident) a.A[(i64)col * a.lda + r0 + i] = x;
      if (col < kk) {
        const int gi = r0 + i;
        a.Vc[(i64)col * a.ldvc + gi] = gi < col ? Sc<T>::zero() : (gi == col ? Sc<T>::one() : x);
      }
    }
  {
    const int col = tid % NB;  // VcT is contiguous along the reflector index
    if (col < kk)
      for (int i = tid / NB; i < rows; i += PANEL_THREADS / NB) {
        const int gi = r0 + i;
        const T v = gi < col ? Sc<T>::zero() : (gi == col ? Sc<T>::one() : S[i + (i64)col * ld]);
        a.VcT[(i64)gi * a.ldvct + col] = v;
      }
  }
}

// ------------------------------------------------------------------------------- K1c: cluster panel kernel
// Panels of at most 16 x 416 rows (Float64): ONE thread-block cluster, CTA r holds rows [r*ROWS, ...) of the panel in shared
// memory (row-major, padded), and the per-column exchange goes through DISTRIBUTED SHARED MEMORY instead of L2: every
// CTA sends its 2 x 64 partials (dots of the un-normalised pivot column with every column, and its contribution to
// pivot row j) into the same slot of every CTA's exchange buffer with st.async, which completes the bytes on a
// transaction barrier of the receiver; every CTA waits on its own barrier and sums the slots in rank order (bitwise
// identical scalars everywhere, deterministic).  The L2 flag exchange of qr_panel_kernel costs ~4000 cycles per column;
// remote stores + barrier.cluster (the first version of this kernel) 1000 - 1250; this one a one-way trip.  Plain
// right-looking steps (no 16-column sub-panels): at these slab heights the rank-1 update of the whole slab is cheaper
// than the bookkeeping of the blocked form.
// Semantics: qrUnblocked! (src/qr.jl:86-111) with stdlib reflector! / reflectorApply! (call sites :96, :102), then the
// clean reflector block Vc (unit diagonal, zeros above) and its transpose VcT like qr_panel_kernel.
constexpr int CL_MAX = 16;             // largest cluster (8 is the portable size; 16 needs the non-portable opt-in, see launch_panel)
constexpr int CP_THREADS = 512;
constexpr int CP_RG = CP_THREADS / NB;   // row groups: thread (c, rg) owns rows rg, rg + CP_RG, ... of column c
constexpr int CP_LD = NB + 1;          // padded row of the slab

template <class T>
struct ClusterPanelSmem {   // fixed part; the slab follows
  T xbuf[2][CL_MAX][2 * NB];   // [parity][source rank][dots | pivot-row entries]
  T part[CP_RG][NB];           // partial dots of the row groups
  T tot[2 * NB];               // reduced dots | pivot row
  unsigned long long xbar[2];  // one transaction barrier per parity: the exchange of a column is complete when the
                               // 2 x NB entries of every rank have landed in xbuf[parity]
};

// st.async: a store into the shared memory of a CTA of the cluster that completes `sizeof(T)` bytes on a transaction
// barrier of THAT CTA.  The receiver waits on its own barrier: one one-way trip instead of the all-to-all
// barrier.cluster round (clock64 trace: 1000 - 1250 of the 4600 cycles of a column went into cluster.sync()).
__device__ __forceinline__ unsigned cl_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cl_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_async(unsigned dst, float v, unsigned bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst),
               "r"(__float_as_uint(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cl_st_async(unsigned dst, double v, unsigned bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst),
               "l"(__double_as_longlong(v)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cl_st_async(unsigned dst, zd v, unsigned bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst),
               "l"(__double_as_longlong(v.x)), "l"(__double_as_longlong(v.y)), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cl_bar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_bar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cl_bar_wait(unsigned bar, unsigned parity) {   // false: watchdog expired
  for (unsigned spins = 0; spins < (1u << 20); ++spins) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

template <class T>
__global__ void __launch_bounds__(CP_THREADS, 1) qr_panel_cluster_kernel(PanelArgs<T> a) {
  using R = typename Sc<T>::real;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClusterPanelSmem<T>& sm = *reinterpret_cast<ClusterPanelSmem<T>*>(smem_raw);
  T* S = reinterpret_cast<T*>(smem_raw + sizeof(ClusterPanelSmem<T>));   // S[i * CP_LD + c]
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int c = tid & (NB - 1), rg = tid >> 6;   // column, row group (rows rg, rg + CP_RG, ...)
  const int r0 = rank * a.rows_per;
  const int r1 = (r0 + a.rows_per < a.mk) ? r0 + a.rows_per : a.mk;
  const int rows = r1 > r0 ? r1 - r0 : 0;
  const int nb = a.nb;
  const int kk = a.mk < nb ? a.mk : nb;

  pdl_wait();   // programmatic dependent launch (common.cuh)
  // ---- slab in: coalesced along the rows, L2-only loads (the far update of the other stream just rewrote the panel)
  for (int e = tid; e < rows * nb; e += CP_THREADS) {
    const int col = e / rows, i = e - col * rows;
    S[i * CP_LD + col] = ldcg_t(a.A + (i64)col * a.lda + r0 + i);
  }
  const unsigned xbar0 = cl_smem_addr(&sm.xbar[0]), xbar1 = cl_smem_addr(&sm.xbar[1]);
  if (tid == 0) {
    cl_bar_init(xbar0, 1);
    cl_bar_init(xbar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // every CTA of the cluster must be running (and its barriers initialised) before anybody writes into its shared
  // memory (compute-sanitizer: "address located in a block that might not have entered yet"); the barrier also orders
  // the slab loads
  cluster.sync();
  const unsigned xbytes = (unsigned)(CL * 2 * NB * sizeof(T));

  T pend_ixi = Sc<T>::one();   // deferred finish of the previous pivot column (rows below the diagonal *= 1/xi, diagonal <- -nu)
  R pend_nu = R(0);
  bool pend = false;
  auto finish = [&](int jc) {
    // by ALL threads, one row each: the four threads of column jc doing it alone was a serial 32-row loop at the head of
    // every column step (a clock64 trace of the 1024-row panel: 1400 of the 5300 cycles of a column)
    if (!pend) return;
    int lo = jc + 1 - r0;
    if (lo < 0) lo = 0;
    for (int i = lo + tid; i < rows; i += CP_THREADS) S[i * CP_LD + jc] = S[i * CP_LD + jc] * pend_ixi;
    if (tid == 0 && jc >= r0 && jc < r1) S[(jc - r0) * CP_LD + jc] = Sc<T>::from_real(-pend_nu);
  };

  for (int j = 0; j < kk; ++j) {
    if (j > 0) finish(j - 1);   // threads of column j-1 are idle from here on; nobody reads that column any more
    const int par = j & 1;
    int lo = j + 1 - r0;        // first local row strictly below the diagonal
    if (lo < 0) lo = 0;
    // ---- partial dots of the un-normalised pivot column with every column c >= j (c = j: tail norm^2)
    T d = Sc<T>::zero();
    if (c >= j && c < nb) {
      T d1 = Sc<T>::zero(), d2 = Sc<T>::zero(), d3 = Sc<T>::zero();
      int i = lo + rg;
      for (; i + 3 * CP_RG < rows; i += 4 * CP_RG) {   // four independent accumulators: the loads of a round are all in flight together
        const T p0 = S[i * CP_LD + j], p1 = S[(i + CP_RG) * CP_LD + j], p2 = S[(i + 2 * CP_RG) * CP_LD + j], p3 = S[(i + 3 * CP_RG) * CP_LD + j];
        const T q0 = S[i * CP_LD + c], q1 = S[(i + CP_RG) * CP_LD + c], q2 = S[(i + 2 * CP_RG) * CP_LD + c], q3 = S[(i + 3 * CP_RG) * CP_LD + c];
        d = fmad(cj(p0), q0, d);
        d1 = fmad(cj(p1), q1, d1);
        d2 = fmad(cj(p2), q2, d2);
        d3 = fmad(cj(p3), q3, d3);
      }
      for (; i < rows; i += CP_RG) d = fmad(cj(S[i * CP_LD + j]), S[i * CP_LD + c], d);
      d = (d + d1) + (d2 + d3);
    }
    sm.part[rg][c] = d;
    // the exchange of column j uses barrier `par` for the (j / 2)-th time.  No CTA can be two columns ahead: its stores
    // of column j + 2 come after its wait of column j + 1, which needs this CTA's stores of column j + 1, which come
    // after this CTA's reads of column j -- so two parities are enough for the buffers and for the barrier phases.
    if (tid == 0 && CL > 1) cl_bar_expect(par ? xbar1 : xbar0, xbytes);
    __syncthreads();
    {
      // ALL threads form the value of slot v (four copies of each: the dots of this CTA for v < NB, its pivot-row entries
      // beyond), and copy `grp` of a slot sends it to the ranks grp, grp + 4, ...: four st.async per thread instead of
      // sixteen by a quarter of the threads (an ncu capture of the 1024-row panel: 41 % of the kernel was this phase, with
      // twelve of the sixteen warps waiting at the barrier behind it)
      const int v = tid & (2 * NB - 1), grp = tid >> 7;
      T val;
      if (v < NB) {
        val = sm.part[0][v];
#pragma unroll
        for (int g2 = 1; g2 < CP_RG; ++g2) val = val + sm.part[g2][v];
      } else {
        val = (j >= r0 && j < r1 && v - NB < nb) ? S[(j - r0) * CP_LD + (v - NB)] : Sc<T>::zero();
      }
      if (CL > 1) {
        const unsigned slot = cl_smem_addr(&sm.xbuf[par][rank][v]), bar = par ? xbar1 : xbar0;
        for (int r = grp; r < CL; r += CP_THREADS / (2 * NB))   // same slot of every CTA's exchange buffer (distributed shared memory)
          cl_st_async(cl_mapa(slot, (unsigned)r), val, cl_mapa(bar, (unsigned)r));
        if (tid < 2 * NB) {
          if (!cl_bar_wait(bar, (unsigned)((j >> 1) & 1))) __trap();
          // fixed pairwise order over the ranks (the same in every CTA: bitwise identical sums), four levels instead of
          // a chain of fifteen dependent additions
          T x[CL_MAX];
#pragma unroll
          for (int r = 0; r < CL_MAX; ++r) x[r] = r < CL ? sm.xbuf[par][r][tid] : Sc<T>::zero();
#pragma unroll
          for (int st2 = 1; st2 < CL_MAX; st2 <<= 1)
#pragma unroll
            for (int r = 0; r + st2 < CL_MAX; r += 2 * st2) x[r] = x[r] + x[r + st2];
          sm.tot[tid] = x[0];
        }
      } else if (tid < 2 * NB) {
        sm.tot[tid] = val;   // (a panel of one CTA has nothing to exchange)
      }
    }
    __syncthreads();
    const T alpha = sm.tot[NB + j];
    const R n2 = abs2(alpha) + re(sm.tot[j]);
    ReflScalars<T> rs;
    if constexpr (Sc<T>::is_complex) rs = reflector_scalars<T>(alpha, n2);
    else rs = reflector_scalars_fast<T>(alpha, n2);
    pend = rs.nonzero;
    pend_ixi = rs.ixi;
    pend_nu = rs.nu;
    if (rank == 0 && tid == 0) a.tau[j] = rs.tau;
    if (rs.nonzero && c > j && c < nb) {
      // s = conj(tau) (a_jc + conj(1/xi) d_c);  a_jc -= s;  rows below -= x * (s / xi)
      const T s = cj(rs.tau) * (sm.tot[NB + c] + cj(rs.ixi) * sm.tot[c]);
      const T t = s * rs.ixi;
      int i = lo + rg;
      for (; i + 3 * CP_RG < rows; i += 4 * CP_RG) {   // column c != column j: loads first, then the four stores
        const T p0 = S[i * CP_LD + j], p1 = S[(i + CP_RG) * CP_LD + j], p2 = S[(i + 2 * CP_RG) * CP_LD + j], p3 = S[(i + 3 * CP_RG) * CP_LD + j];
        const T q0 = S[i * CP_LD + c], q1 = S[(i + CP_RG) * CP_LD + c], q2 = S[(i + 2 * CP_RG) * CP_LD + c], q3 = S[(i + 3 * CP_RG) * CP_LD + c];
        S[i * CP_LD + c] = q0 - p0 * t;
        S[(i + CP_RG) * CP_LD + c] = q1 - p1 * t;
        S[(i + 2 * CP_RG) * CP_LD + c] = q2 - p2 * t;
        S[(i + 3 * CP_RG) * CP_LD + c] = q3 - p3 * t;
      }
      for (; i < rows; i += CP_RG) S[i * CP_LD + c] = S[i * CP_LD + c] - S[i * CP_LD + j] * t;
      if (rg == 0 && j >= r0 && j < r1) S[(j - r0) * CP_LD + c] = S[(j - r0) * CP_LD + c] - s;
    }
    __syncthreads();
  }
  if (kk > 0) finish(kk - 1);
  __syncthreads();
  pdl_launch_dependents();   // only the stores are left

  // ---- panel, clean reflectors and their transpose out
  for (int e = tid; e < rows * nb; e += CP_THREADS) {
    const int col = e / rows, i = e - col * rows;
    const int gi = r0 + i;
    const T v = S[i * CP_LD + col];
    a.A[(i64)col * a.lda + gi] = v;
    if (col < kk) a.Vc[(i64)col * a.ldvc + gi] = gi < col ? Sc<T>::zero() : (gi == col ? Sc<T>::one() : v);
  }
  for (int e = tid; e < rows * kk; e += CP_THREADS) {
    const int i = e / kk, col = e - i * kk;
    const int gi = r0 + i;
    const T v = S[i * CP_LD + col];
    a.VcT[(i64)gi * a.ldvct + col] = gi < col ? Sc<T>::zero() : (gi == col ? Sc<T>::one() : v);
  }
}

// ------------------------------------------------------------------------------- clean V from factors
// Vc (mk x kk, ldvc) and VcT (kk x mk, ldvct) from the factored panel F (unit lower trapezoid)
template <class T>
__global__ void extract_v_kernel(const T* __restrict__ F, i64 ldf, int mk, int kk, T* __restrict__ Vc, i64 ldvc,
                                 T* __restrict__ VcT, i64 ldvct) {
  const i64 total = (i64)mk * kk;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const int col = (int)(e / mk), i = (int)(e - (i64)col * mk);
    T v = i < col ? Sc<T>::zero() : (i == col ? Sc<T>::one() : F[(i64)col * ldf + i]);
    Vc[(i64)col * ldvc + i] = v;
    VcT[(i64)i * ldvct + col] = v;
  }
}

// ------------------------------------------------------------------------------- K2: T from G
// G partials: nsplit slices of kk x kk (ld = kk, stride gstride).  T (kk x kk, ldt) upper.
//   U = diag(tau) striu(G);  X = (I + U)^-1;  T = X diag(tau)          (src/qr.jl:70-81)
template <class T>
__global__ void __launch_bounds__(256) larft_finish_kernel(const T* __restrict__ Gp, i64 gstride, int nsplit, int kk,
                                                           const T* __restrict__ tau, T* __restrict__ Tm, i64 ldt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef T Row[NB + 1];
  Row* U = reinterpret_cast<Row*>(smem_raw);
  Row* X = U + NB;
  Row* Pk = X + NB;
  T* st = reinterpret_cast<T*>(Pk + NB);
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  // this kernel sits on the panel chain of every QR: all loads of a thread (its 16 entries of G, one tau) are issued
  // before the first use -- one L2 round trip instead of seventeen dependent ones
  constexpr int PER = NB * NB / 256;
  T gv[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int e = tid + q * 256, j = e / NB, i = e - j * NB;
    gv[q] = (i < j && j < kk) ? ldcg_t(Gp + (i64)j * kk + i) : Sc<T>::zero();
  }
  if (tid < kk) st[tid] = ldcg_t(tau + tid);
  if (nsplit > 1) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int e = tid + q * 256, j = e / NB, i = e - j * NB;
      if (i < j && j < kk)
        for (int z = 1; z < nsplit; ++z) gv[q] = gv[q] + ldcg_t(Gp + (i64)z * gstride + (i64)j * kk + i);
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < PER; ++q) {   // identity padding beyond kk
    const int e = tid + q * 256, j = e / NB, i = e - j * NB;
    U[i][j] = (i < j && j < kk) ? st[i] * gv[q] : Sc<T>::zero();
    X[i][j] = (i == j) ? Sc<T>::one() : Sc<T>::zero();
  }
  __syncthreads();
  // X = (I + U)^-1 by recursive doubling over diagonal blocks of 1, 2, 4 .. 32: X12 = -X11 (U12 X22); twelve barriers with all
  // threads busy instead of 63 dependent column steps (39 -> ~10 us per panel; it sits on the panel chain of every QR)
  for (int b = 1; b < NB; b <<= 1) {
    const int elems = (NB / 2) * b;
    const int lb = 31 - __clz(b);
    for (int e = tid; e < elems; e += blockDim.x) {
      const int p = e >> (2 * lb), rem = e & (b * b - 1);
      const int base = p * 2 * b;
      const int i = base + (rem >> lb), c = base + b + (rem & (b - 1));
      T acc = Sc<T>::zero();
      for (int l = base + b; l <= c; ++l) acc = fmad(U[i][l], X[l][c], acc);
      Pk[i][c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < elems; e += blockDim.x) {
      const int p = e >> (2 * lb), rem = e & (b * b - 1);
      const int base = p * 2 * b;
      const int i = base + (rem >> lb), c = base + b + (rem & (b - 1));
      T acc = Sc<T>::zero();
      for (int l = i; l < base + b; ++l) acc = fmad(X[i][l], Pk[l][c], acc);
      X[i][c] = -acc;
    }
    __syncthreads();
  }
  for (int e = tid; e < kk * kk; e += blockDim.x) {
    const int j = e / kk, ii = e - j * kk;
    Tm[(i64)j * ldt + ii] = ii <= j ? X[ii][j] * st[j] : Sc<T>::zero();
  }
}

// ------------------------------------------------------------------------------- outer-block recurrence
// Two-level blocking: NBO/NB consecutive panels form one OUTER block whose reflectors are applied to the
// far trailing matrix in a single pass with K = NBO (K = NB would make that pass HBM-bound: 16 B of C
// traffic per 2*NB flops).  With V = [V_0 .. V_{nj-1}], W = V^H A2 and G = V^H V, the product
// Q_{nj-1}^H ... Q_0^H A2 = A2 - V Z follows from the block recurrence
//     Y_j = W_j - sum_{i<j} G_ji Z_i ,   Z_j = T_j^H Y_j ,
// which needs only the per-panel T_j (never the NBO x NBO T); apply_outer runs it as TN contractions on the tensor
// pipe (a scalar one-CTA-per-32-columns kernel did the same in round 1: 280 ms against 262 ms at n = 16384).

// ------------------------------------------------------------------------------- workspace
constexpr int NBO = 384;  // outer block: K of the far trailing contractions (measured at n=16384: 256 -> 258.8 ms, 384 -> 254.0 ms, 512 -> 258.3 ms)
// NBO is the CAPACITY (leading dimensions, buffer sizes); the factorisation steps through outer blocks of outer_width()
// columns: 256 up to n = 12288, 384 beyond.  Measured at the end of round 2 (faster panel chain than when 384 was chosen):
// 256 against 384 -- n = 1024 2.59 / 2.68 ms, n = 2048 6.20 / 6.51, n = 4096 15.9 / 16.6, n = 8192 49.4 / 51.0, n = 16384
// 243.4 / 238.2.  GLA_QR_NBO = 128 ... 384 (multiple of 64) forces a width.
static int outer_width(i64 n) {
  static const int forced = [] {
    const char* e = getenv("GLA_QR_NBO");
    const int v = e ? atoi(e) : 0;
    return (v >= 2 * NB && v <= NBO && v % NB == 0) ? v : 0;
  }();
  if (forced) return forced;
  return n <= 12288 ? 256 : NBO;
}

// Workspace of one factorisation.  The outer-block state (V, VT, per-panel T, Gram) is DOUBLE BUFFERED by outer
// block parity and every scratch array exists once per execution path, because the driver overlaps two paths:
//   path 0 (panel chain, high-priority stream): panels, per-panel T, updates inside the outer block
//   path 1 (far update, caller's stream):       Gram of the outer block, W = V^H A2, fix-up, A2 -= V Z
template <class T>
struct QrWork {
  T* V[2] = {nullptr, nullptr};    // m x NBO clean reflectors of an outer block (unit diagonal, zeros above)
  i64 ldv = 0;
  T* VT[2] = {nullptr, nullptr};   // NBO x m transpose (ld NBO)
  T* Tm[2] = {nullptr, nullptr};   // NBO/NB per-panel T factors, NB x NB each
  T* G[2] = {nullptr, nullptr};    // NBO x NBO Gram V^H V (ld NBO)
  T* Gs = nullptr;                 // NB x NB Gram scratch of build_T (path 0)
  T* Gp[2] = {nullptr, nullptr};   // split-K partials of a Gram, per path
  T* Wp[2] = {nullptr, nullptr};   // split-K partials of W = V^H A2, per path
  T* Z[2] = {nullptr, nullptr};    // NBO x nA, per path
  i64 wp_elems[2] = {0, 0};
  int yield_sms = 0;               // look-ahead active: the far-update products leave SMs to the panel chain (GemmTN::yield_sms)
  ulonglong2* xd = nullptr;        // LL exchange buffers of the panel kernel
  ulonglong2* xr = nullptr;
  ulonglong2* xw = nullptr;
  ulonglong2* xz = nullptr;
  unsigned long long epoch = 0;
  void* block = nullptr;
  cudaStream_t st = nullptr;
  int cur = 0;                     // outer buffer the panel chain is filling
  static constexpr int GP_SPLITS = 32;

  // nA0 / nA1: widest trailing block handled on path 0 / path 1 (0 = path unused); nbuf outer buffers
  // total_only != nullptr: size computation only (gla_workspace_query), nothing is allocated
  int alloc(i64 m, i64 nA0, i64 nA1, int nbuf, cudaStream_t stream, i64* total_only = nullptr) {
    st = stream;
    ldv = round_up(m, 4);   // 16-byte aligned columns for Float32 as well (TMA operands)
    auto al = [](i64 bytes) { return round_up(bytes, 256); };
    const i64 s_v = al(ldv * NBO * sizeof(T));
    const i64 s_vt = al((i64)NBO * m * sizeof(T));
    const i64 s_tm = al((i64)NBO * NB * sizeof(T));
    const i64 s_g = al((i64)NBO * NBO * sizeof(T));
    const i64 s_gs = al((i64)NB * NB * sizeof(T));
    const i64 s_gp = al((i64)GP_SPLITS * NBO * NBO * sizeof(T));
    i64 s_wp[2], s_z[2];
    const i64 nAs[2] = {nA0, nA1};
    for (int q = 0; q < 2; ++q) {
      i64 nA = nAs[q] < 1 ? 1 : nAs[q];
      // W partials: at most ~2*SMs tiles worth of split-K slices; 6 full W matrices, never less than what a
      // 64-slice split of a narrow (<= 4*NBO columns) W needs
      i64 wcols = 8 * nA;
      if (wcols < 64 * 4 * NBO && nA <= 4 * NBO) wcols = 64 * nA;
      wp_elems[q] = nAs[q] > 0 ? (i64)NBO * wcols : 0;
      s_wp[q] = al((wp_elems[q] > 0 ? wp_elems[q] : 1) * sizeof(T));
      s_z[q] = al((i64)NBO * nA * sizeof(T));
    }
    constexpr i64 NR = Sc<T>::is_complex ? 2 : 1;
    const i64 s_xd = al((i64)2 * PANEL_PMAX * SB * NR * 16);
    const i64 s_xr = al((i64)2 * SB * NR * 16);
    const i64 s_xw = al((i64)PANEL_PMAX * GW_MAX * NR * 16);
    const i64 s_xz = al((i64)2 * GW_MAX * NR * 16);
    const i64 total = nbuf * (s_v + s_vt + s_tm + s_g) + s_gs + 2 * s_gp + s_wp[0] + s_wp[1] + s_z[0] + s_z[1] +
                      s_xd + s_xr + s_xw + s_xz + 256;
    if (total_only) {
      *total_only = total;
      return 0;
    }
    GLA_TRY(pool_malloc(reinterpret_cast<void**>(&block), total, st));
    char* p = static_cast<char*>(block);
    for (int b = 0; b < 2; ++b) {
      const int src = b < nbuf ? b : 0;
      if (b < nbuf) {
        V[b] = reinterpret_cast<T*>(p); p += s_v;
        VT[b] = reinterpret_cast<T*>(p); p += s_vt;
        Tm[b] = reinterpret_cast<T*>(p); p += s_tm;
        G[b] = reinterpret_cast<T*>(p); p += s_g;
      } else {
        V[b] = V[src]; VT[b] = VT[src]; Tm[b] = Tm[src]; G[b] = G[src];
      }
    }
    Gs = reinterpret_cast<T*>(p); p += s_gs;
    for (int q = 0; q < 2; ++q) {
      Gp[q] = reinterpret_cast<T*>(p); p += s_gp;
      Wp[q] = reinterpret_cast<T*>(p); p += s_wp[q];
      Z[q] = reinterpret_cast<T*>(p); p += s_z[q];
    }
    xd = reinterpret_cast<ulonglong2*>(p); p += s_xd;
    xr = reinterpret_cast<ulonglong2*>(p); p += s_xr;
    xw = reinterpret_cast<ulonglong2*>(p); p += s_xw;
    xz = reinterpret_cast<ulonglong2*>(p); p += s_xz;
    // tags of a fresh workspace: zero never matches (epochs start at 1)
    GLA_CUDA(cudaMemsetAsync(xd, 0, s_xd + s_xr + s_xw + s_xz, st));
    epoch = 0;
    return 0;
  }
  void release() {
    if (block) cudaFreeAsync(block, st);
    block = nullptr;
  }
  // views of inner panel j (columns j0 = j*NB of the outer block, rows from j0) in the chain's buffer
  T* Vj(int j) const { return V[cur] + (i64)j * NB + (i64)j * NB * ldv; }
  T* VTj(int j) const { return VT[cur] + (i64)j * NB * NBO + (i64)j * NB; }
  T* Tj(int j) const { return Tm[cur] + (i64)j * NB * NB; }
};

// rows above panel j's diagonal block inside the outer block are zero in V / VT
template <class T>
__global__ void zero_top_kernel(T* __restrict__ V, i64 ldv, T* __restrict__ VT, int j0, int kk) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < j0 * kk; e += gridDim.x * blockDim.x) {
    const int col = e / j0, i = e - col * j0;
    V[(i64)(j0 + col) * ldv + i] = Sc<T>::zero();
    VT[(i64)i * NBO + j0 + col] = Sc<T>::zero();
  }
}

// the same for every panel of an outer block of nbo columns at once (one launch per outer block, off the panel chain)
template <class T>
__global__ void zero_top_block_kernel(T* __restrict__ V, i64 ldv, T* __restrict__ VT, int nbo) {
  const int total = nbo * nbo;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int col = e / nbo, i = e - col * nbo;
    if (i < (col / NB) * NB) {
      V[(i64)col * ldv + i] = Sc<T>::zero();
      VT[(i64)i * NBO + col] = Sc<T>::zero();
    }
  }
}

// ------------------------------------------------------------------------------- panel launch
template <class T>
static int launch_panel(T* A, i64 lda, i64 mk, int nb, T* tau, QrWork<T>& w, int j, cudaStream_t st, bool zero_top = true) {
  PanelArgs<T> a;
  a.A = A;
  a.lda = lda;
  a.mk = (int)mk;
  a.nb = nb;
  a.tau = tau;
  a.Vc = w.Vj(j);
  a.ldvc = w.ldv;
  a.VcT = w.VTj(j);
  a.ldvct = NBO;
  a.xd = w.xd;
  a.xr = w.xr;
  a.xw = w.xw;
  a.xz = w.xz;
  a.epoch = ++w.epoch;
  {  // panels that fit one thread-block cluster: exchange through distributed shared memory (qr_panel_cluster_kernel)
    static const bool no_cluster = getenv("GLA_PANEL_NO_CLUSTER") != nullptr;
    const i64 slab_cap = 226 * 1024 - (i64)sizeof(ClusterPanelSmem<T>);   // 227 KB per CTA minus the exchange buffers
    const i64 rows_max = slab_cap / ((i64)CP_LD * sizeof(T)) / 4 * 4;
    // measured (profiles/r02_qr_panel_cluster.txt): n = 1024 5.74 -> 4.30 ms, n = 2048 11.5 -> 10.4 ms, n = 4096 equal; beyond
    // 256 rows per CTA the rank-1 update of the whole slab per column costs more than the sub-panel form of qr_panel_kernel
    // (Float32 slabs cost half the shared-memory traffic per row: no cap below the capacity of 888 rows -- n = 8192 37.5 -> 34.3 ms,
    //  n = 16384 119.2 -> 113.7 ms against a cap of 320; Float64: 320 -> 416 is worth 0.2 - 0.6 % from n = 6144 on)
    static const i64 rows_cap = [] { const char* e = getenv("GLA_PANEL_CLUSTER_ROWS"); return e ? (i64)atoi(e) : (i64)(sizeof(T) == 4 ? 1024 : 416); }();
    const i64 rows_use = rows_max < rows_cap ? rows_max : rows_cap;
    auto kern = qr_panel_cluster_kernel<T>;
    // clusters of 16 CTAs (one GPC holds them): opt-in, and only if the occupancy query says such a cluster can be resident
    static const int cl_max = [&]() -> int {
      const char* e = getenv("GLA_PANEL_CLUSTER_MAX");
      int want = e ? atoi(e) : 16;
      if (want > CL_MAX) want = CL_MAX;
      if (want <= 8) return want < 1 ? 1 : want;
      if (ensure_dyn_smem((const void*)kern, (int)(sizeof(ClusterPanelSmem<T>) + slab_cap))) return 8;
      if (cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return 8;
      }
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)want);
      q.blockDim = dim3(CP_THREADS);
      q.dynamicSmemBytes = sizeof(ClusterPanelSmem<T>) + (size_t)slab_cap;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = (unsigned)want;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, (const void*)kern, &q) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        return 8;
      }
      return want;
    }();
    if (!no_cluster && mk <= cl_max * rows_use) {
      int CL = (int)((mk + 63) / 64);
      if (CL > cl_max) CL = cl_max;
      if (CL < 1) CL = 1;
      const i64 rows_per = round_up((mk + CL - 1) / CL, 4);
      CL = (int)((mk + rows_per - 1) / rows_per);
      a.rows_per = (int)rows_per;
      a.resident = 1;
      a.lds = CP_LD;
      const size_t smem = sizeof(ClusterPanelSmem<T>) + (size_t)rows_per * CP_LD * sizeof(T);
      GLA_TRY(ensure_dyn_smem((const void*)kern, (int)(sizeof(ClusterPanelSmem<T>) + slab_cap)));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)CL);
      cfg.blockDim = dim3(CP_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)CL;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      GLA_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
      if (j > 0 && zero_top) {
        const int kk = (int)(mk < nb ? mk : nb);
        zero_top_kernel<T><<<ceil_div((i64)j * NB * kk, 256), 256, 0, st>>>(w.V[w.cur], w.ldv, w.VT[w.cur], j * NB, kk);
        GLA_CUDA(cudaGetLastError());
      }
      return 0;
    }
  }
  const i64 smem_cap = 176 * 1024;
  const i64 max_rows = (smem_cap / ((i64)nb * sizeof(T)) - 4) / 16 * 16;  // rows that fit one CTA
  const int sms = sm_count();
  static const int min_rows = [] { const char* e = getenv("GLA_PANEL_MINROWS"); return e ? atoi(e) : 64; }();
  static const int max_ctas = [] { const char* e = getenv("GLA_PANEL_MAXCTAS"); return e ? atoi(e) : PANEL_MAX_CTAS; }();
  i64 rows_per = round_up((mk + max_ctas - 1) / max_ctas, 16);
  if (rows_per < min_rows) rows_per = min_rows;
  int resident = 1;
  if (rows_per > max_rows) {
    rows_per = max_rows;
    if ((mk + rows_per - 1) / rows_per > sms) {  // does not fit the chip: operate on global memory
      resident = 0;
      rows_per = round_up((mk + sms - 1) / sms, 16);
    }
  }
  const int P = (int)((mk + rows_per - 1) / rows_per);
  a.rows_per = (int)rows_per;
  a.resident = resident;
  a.lds = (int)(round_up(rows_per, 16) + 4);
  size_t smem = resident ? (size_t)a.lds * nb * sizeof(T) : 0;
  auto kern = qr_panel_kernel<T>;
  GLA_TRY(ensure_dyn_smem((const void*)kern, (int)((int)(smem_cap + 8 * 1024))));
  void* args[] = {&a};
  if (P > 1) {
    GLA_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(P), dim3(PANEL_THREADS), args, smem, st));
  } else {
    kern<<<1, PANEL_THREADS, smem, st>>>(a);
    GLA_CUDA(cudaGetLastError());
  }
  if (j > 0 && zero_top) {
    const int kk = (int)(mk < nb ? mk : nb);
    zero_top_kernel<T><<<ceil_div((i64)j * NB * kk, 256), 256, 0, st>>>(w.V[w.cur], w.ldv, w.VT[w.cur], j * NB, kk);
    GLA_CUDA(cudaGetLastError());
  }
  return 0;
}

// Gram of a clean reflector block: out (kk x kk, ldo) = Vc^H Vc, split-K partials summed in fixed order
template <class T>
static int gram(QrWork<T>& w, int path, const T* Vc, i64 ldvc, i64 mk, int kk, T* out, i64 ldo, cudaStream_t st) {
  GemmTN<T> g;
  g.At = Vc; g.ldat = ldvc;
  g.B = Vc; g.ldb = ldvc;
  g.C = w.Gp[path]; g.ldc = kk;
  g.M = kk; g.N = kk; g.K = mk;
  g.conj_a = 1;
  g.nsplit = choose_nsplit(kk, kk, mk, kk <= 64 ? 64 : 128, kk <= 64 ? 128 : 64);
  if (g.nsplit > QrWork<T>::GP_SPLITS) g.nsplit = QrWork<T>::GP_SPLITS;
  g.split_stride = (i64)kk * kk + (((i64)kk * kk) & 1);
  GLA_TRY(gemm_tn<T>(g, st));
  return sum_splits<T>(out, ldo, w.Gp[path], kk, g.split_stride, g.nsplit, kk, kk, st);
}

// T_j (kk x kk, ld NB) of the clean reflector block (Vc, ldvc) with tau (path 0 scratch)
template <class T>
static int build_T(QrWork<T>& w, const T* Vc, i64 ldvc, i64 mk, int kk, const T* tau, T* Tout, cudaStream_t st) {
  GLA_TRY(gram<T>(w, 0, Vc, ldvc, mk, kk, w.Gs, kk, st));
  const int smem = (3 * NB * (NB + 1) + NB) * (int)sizeof(T);
  GLA_TRY(ensure_dyn_smem((const void*)larft_finish_kernel<T>, (int)(smem)));
  GLA_CUDA(launch_pdl(larft_finish_kernel<T>, dim3(1), dim3(256), (size_t)smem, st, (const T*)w.Gs, (i64)0, 1, kk, tau, Tout, (i64)NB));
  return 0;
}

template <class T>
static int wsplit_for(i64 kk, i64 nA, i64 mk) {
  if (sizeof(T) == 4 && kk >= 128 && nA >= 128) return choose_nsplit_persistent(kk, nA, mk, 8);   // tcgen05 path (gemm.cu)
  return choose_nsplit(kk, nA, mk, kk <= 64 ? 64 : 128, kk <= 64 ? 128 : 64);
}

// A2 (mk x nA, lda) <- (I - Vc op(T) Vc^H) A2 for ONE panel (kk <= NB reflectors); path 0 scratch
template <class T>
static int apply_panel(QrWork<T>& w, const T* Vc, i64 ldvc, const T* VcT, i64 ldvct, const T* Tj, i64 mk, int kk,
                       T* A2, i64 lda, i64 nA, int adjoint, cudaStream_t st, cudaEvent_t t_ready = nullptr) {
  if (nA <= 0) return 0;
  GemmTN<T> g1;
  g1.At = Vc; g1.ldat = ldvc;
  g1.B = A2; g1.ldb = lda;
  g1.C = w.Wp[0]; g1.ldc = NB;
  g1.M = kk; g1.N = nA; g1.K = mk;
  g1.conj_a = 1;
  g1.nsplit = wsplit_for<T>(kk, nA, mk);
  g1.split_stride = (i64)NB * nA;
  if ((i64)g1.nsplit * NB * nA > w.wp_elems[0]) g1.nsplit = (int)(w.wp_elems[0] / ((i64)NB * nA));
  if (g1.nsplit < 1) {
    set_error(GLA_ERR_INTERNAL, "W workspace too small", __FILE__, __LINE__);
    return GLA_ERR_INTERNAL;
  }
  GLA_TRY(gemm_tn<T>(g1, st));
  if (!adjoint) {
    set_error(GLA_ERR_INTERNAL, "apply_panel is the adjoint (factorisation) form only; Q A goes through apply_outer_fwd", __FILE__,
              __LINE__);
    return GLA_ERR_INTERNAL;
  }
  // Z = T_j^H (sum of the split-K slices of W): T_j (column-major, zeros below the diagonal) is the K-contiguous
  // operand of the TN contraction as it stands
  if (g1.nsplit > 1) GLA_TRY(sum_splits<T>(w.Wp[0], NB, w.Wp[0], NB, g1.split_stride, g1.nsplit, kk, nA, st));
  if (t_ready) GLA_CUDA(cudaStreamWaitEvent(st, t_ready, 0));   // T_j was built on another stream meanwhile
  {
    GemmTN<T> gz;
    gz.At = Tj; gz.ldat = NB;
    gz.B = w.Wp[0]; gz.ldb = NB;
    gz.C = w.Z[0]; gz.ldc = NB;
    gz.M = kk; gz.N = nA; gz.K = kk;
    gz.conj_a = 1;
    GLA_TRY(gemm_tn<T>(gz, st));
  }
  GemmTN<T> g2;
  g2.At = VcT; g2.ldat = ldvct;
  g2.B = w.Z[0]; g2.ldb = NB;
  g2.C = A2; g2.ldc = lda;
  g2.M = mk; g2.N = nA; g2.K = kk;
  g2.alpha = -1;
  g2.beta_one = 1;
  return gemm_tn<T>(g2, st);
}

// A2 (mo x nA) <- Q_{nj-1}^H ... Q_0^H A2 for the whole outer block held in buffer b (kbig reflectors, Gram
// already in w.G[b]); path 1 scratch
template <class T>
static int outer_nsplit(const QrWork<T>& w, int kbig, i64 nA, i64 mo) {   // split-K slices of W = V^H A2 over nA columns
  int ns = wsplit_for<T>(kbig, nA, mo);
  if ((i64)ns * NBO * nA > w.wp_elems[1]) ns = (int)(w.wp_elems[1] / ((i64)NBO * nA));
  return ns;
}

// force_nsplit > 0: the slicing of a wider call this one is a piece of (same summation order, bitwise the same result)
template <class T>
static int apply_outer(QrWork<T>& w, int b, i64 mo, int kbig, T* A2, i64 lda, i64 nA, cudaStream_t st,
                       cudaEvent_t g_ready = nullptr, int force_nsplit = 0) {
  if (nA <= 0) {
    if (g_ready) GLA_CUDA(cudaStreamWaitEvent(st, g_ready, 0));
    return 0;
  }
  GemmTN<T> g1;
  g1.At = w.V[b]; g1.ldat = w.ldv;
  g1.B = A2; g1.ldb = lda;
  g1.C = w.Wp[1]; g1.ldc = NBO;
  g1.M = kbig; g1.N = nA; g1.K = mo;
  g1.conj_a = 1;
  g1.yield_sms = w.yield_sms;
  g1.nsplit = force_nsplit > 0 ? force_nsplit : outer_nsplit<T>(w, kbig, nA, mo);
  g1.split_stride = (i64)NBO * nA;
  if (g1.nsplit < 1) {
    set_error(GLA_ERR_INTERNAL, "W workspace too small", __FILE__, __LINE__);
    return GLA_ERR_INTERNAL;
  }
  GLA_TRY(gemm_tn<T>(g1, st));
  {
    // the block recurrence  Y_j = W_j - sum_{i<j} G_ji Z_i,  Z_j = T_j^H Y_j  as TN contractions on the tensor pipe:
    //   G_ji = conj(G(iNB.., jNB..))^T is the K-contiguous operand G + jNB*ldg (G is Hermitian), T_j^H likewise T_j
    if (g1.nsplit > 1)   // fixed-order sum of the split-K slices, in place into slice 0
      GLA_TRY(sum_splits<T>(w.Wp[1], NBO, w.Wp[1], NBO, g1.split_stride, g1.nsplit, kbig, nA, st));
    if (g_ready) GLA_CUDA(cudaStreamWaitEvent(st, g_ready, 0));   // the Gram of the outer block was formed on another stream meanwhile
    const int nj = (kbig + NB - 1) / NB;
    for (int j = 0; j < nj; ++j) {
      const int rows_j = (kbig - j * NB) < NB ? (kbig - j * NB) : NB;
      if (j > 0) {
        GemmTN<T> gy;
        gy.At = w.G[b] + (i64)j * NB * NBO; gy.ldat = NBO;
        gy.B = w.Z[1]; gy.ldb = NBO;
        gy.C = w.Wp[1] + j * NB; gy.ldc = NBO;
        gy.M = rows_j; gy.N = nA; gy.K = (i64)j * NB;
        gy.alpha = -1; gy.beta_one = 1; gy.conj_a = 1;
        GLA_TRY(gemm_tn<T>(gy, st));
      }
      GemmTN<T> gz;
      gz.At = w.Tm[b] + (i64)j * NB * NB; gz.ldat = NB;
      gz.B = w.Wp[1] + j * NB; gz.ldb = NBO;
      gz.C = w.Z[1] + j * NB; gz.ldc = NBO;
      gz.M = rows_j; gz.N = nA; gz.K = rows_j;
      gz.conj_a = 1;
      GLA_TRY(gemm_tn<T>(gz, st));
    }
  }
  GemmTN<T> g2;
  g2.At = w.VT[b]; g2.ldat = NBO;
  g2.B = w.Z[1]; g2.ldb = NBO;
  g2.C = A2; g2.ldc = lda;
  g2.M = mo; g2.N = nA; g2.K = kbig;
  g2.alpha = -1;
  g2.yield_sms = w.yield_sms;
  g2.beta_one = 1;
  return gemm_tn<T>(g2, st);
}

// ------------------------------------------------------------------------------- drivers
namespace {
struct AuxStream {  // high-priority side streams + the events of the look-ahead schedule (cached per thread and device)
  cudaStream_t s = nullptr;
  cudaStream_t s2 = nullptr;                              // T build of a panel, beside the W product of its application
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_panel = nullptr, ev_t = nullptr;         // panel factored (fork), T ready (join)
  int create() {
    AuxCtx* a = nullptr;
    GLA_TRY(aux_ctx(&a));
    s = a->hi;
    s2 = a->hi2;
    for (int i = 0; i < 3; ++i) ev[i] = a->ev[i];
    ev_panel = a->ev[4];
    ev_t = a->ev[5];
    return 0;
  }
};
}  // namespace

// device bytes geqr_blocked_dev takes from the stream-ordered pool for an m x n problem (gla_workspace_query)
i64 geqr_blocked_workspace_bytes(i64 m, i64 n, i64 elem_bytes) {
  if (m == 0 || n == 0) return 0;
  const int nbw = outer_width(n);
  const bool overlap = n > 2 * nbw && m > 2 * nbw;
  i64 total = 0;
  if (elem_bytes == 4) {
    QrWork<float> w;
    w.alloc(m, NBO, n > nbw ? n - nbw : 0, overlap ? 2 : 1, nullptr, &total);
  } else if (elem_bytes == 8) {
    QrWork<double> w;
    w.alloc(m, NBO, n > nbw ? n - nbw : 0, overlap ? 2 : 1, nullptr, &total);
  } else {
    QrWork<zd> w;
    w.alloc(m, NBO, n > nbw ? n - nbw : 0, overlap ? 2 : 1, nullptr, &total);
  }
  return total;
}

// Look-ahead schedule (one outer block ahead):
//   chain(o)  on the side stream : panels + per-panel T + updates inside outer block o   (needs far A(o-1))
//   far A(o)  on the caller's    : outer block o applied to the columns of outer block o+1
//   far B(o)  on the caller's    : ... and to everything right of it, concurrently with chain(o+1)
template <class T>
int geqr_blocked_dev(T* dA, i64 m, i64 n, i64 lda, T* dtau, i64 /*blocksize_hint*/, cudaStream_t st, QrHostSink<T>* sink) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (lda < (m > 1 ? m : 1)) return -4;
  if (m == 0 || n == 0) return 0;
  // One-outer-block look-ahead: the panel chain of block o+1 runs on a high-priority side stream concurrently with
  // the far update of block o.  (History, see DESIGN.md "Generic loads of TMA-written shared memory": until the GEMM
  // kernels read their fragments with explicit ld.shared, this concurrency exposed wrong tiles; the schedule itself
  // was never at fault.)  GLA_QR_NO_OVERLAP=1 selects the single-stream schedule for A/B measurements.
  static const bool no_overlap = getenv("GLA_QR_NO_OVERLAP") != nullptr;
  const int nbw = outer_width(n);   // columns per outer block (<= NBO, the capacity of the buffers)
  const bool overlap = !no_overlap && n > 2 * nbw && m > 2 * nbw;   // small problems: one stream, one buffer
  PdlScope pdl(n <= pdl_max_n());   // programmatic dependent launch of the small kernels where the launch chain bounds the run (common.cuh)
  QrWork<T> w;
  GLA_TRY(w.alloc(m, NBO, n > nbw ? n - nbw : 0, overlap ? 2 : 1, st));
  w.yield_sms = overlap ? 1 : 0;
  AuxStream aux;
  cudaStream_t sc = st;  // chain stream
  int rc = 0;
  // T fork: the Gram + larft_finish launches that build T_j do not depend on W = V_j^H A2 -- they run on a second
  // high-priority stream beside the W product and its split-K sum, and the chain waits for T_j right before Z = T_j^H W
  // (three launches off the critical path of every panel).  GLA_QR_NO_TFORK=1: everything on the chain stream.
  static const bool no_tfork = getenv("GLA_QR_NO_TFORK") != nullptr;
  const bool tfork = !no_tfork && n > NB && m > NB;
  if (overlap || tfork) rc = aux.create();
  if (overlap) {
    if (!rc) rc = check_cuda(cudaEventRecord(aux.ev[0], st), __FILE__, __LINE__);      // workspace ready
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(aux.s, aux.ev[0], 0), __FILE__, __LINE__);
    sc = aux.s;
  }
  // streamed upload (QrHostSink::up_*): stream `s` waits for every chunk that starts left of column c_end
  int up_waited_chain = 0, up_waited_far = 0;
  auto need_cols = [&](cudaStream_t s, int& waited, i64 c_end) -> int {
    if (!sink || !sink->up_chunks) return 0;
    while (waited < sink->up_chunks && sink->up_col[waited] < c_end) {
      GLA_CUDA(cudaStreamWaitEvent(s, sink->up_ev[waited], 0));
      ++waited;
    }
    return 0;
  };
  bool done = false;
  int ob = 0;
  for (i64 o0 = 0; !done && !rc; o0 += nbw, ++ob) {
    const i64 mo = m - o0, no = n - o0;
    const int nbo = (int)(no < nbw ? no : nbw);  // columns of this outer block
    const int b = overlap ? (ob & 1) : 0;
    w.cur = b;
    int kbig = 0;
    // ---- chain(o)
    // zeros above the diagonal blocks of V / VT for the whole outer block: one launch on the second side stream (the
    // panel kernels write only rows at and below their diagonal block; the first forked T build of the block follows it
    // on that stream, so the chain -- and through it the far update -- is ordered after it)
    if ((rc = need_cols(sc, up_waited_chain, o0 + nbo))) break;
    const bool zero_ahead = tfork && nbo > NB && mo > NB;
    if (zero_ahead) {
      if ((rc = check_cuda(cudaEventRecord(aux.ev_panel, sc), __FILE__, __LINE__))) break;   // V[b] is free from here on
      if ((rc = check_cuda(cudaStreamWaitEvent(aux.s2, aux.ev_panel, 0), __FILE__, __LINE__))) break;
      zero_top_block_kernel<T><<<ceil_div((i64)nbo * nbo, 256), 256, 0, aux.s2>>>(w.V[b], w.ldv, w.VT[b], nbo);
      if ((rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__))) break;
    }
    for (int j = 0; j * NB < nbo; ++j) {
      const i64 k0 = o0 + (i64)j * NB;
      const i64 mk = m - k0;
      if (mk <= 0) {  // no rows left (wide matrix): the previous panel was the last one
        done = true;
        break;
      }
      const int nb = (int)(n - k0 < NB ? n - k0 : NB);
      const int kk = (int)(mk < nb ? mk : nb);
      T* Ak = dA + k0 + k0 * lda;
      if ((rc = launch_panel<T>(Ak, lda, mk, nb, dtau + k0, w, j, sc, !zero_ahead))) break;
      kbig += kk;
      const i64 nin = (o0 + nbo) - (k0 + nb);  // remaining columns INSIDE the outer block
      const i64 nfar = n - (o0 + nbo);
      bool t_forked = false;
      if (nin > 0 || nfar > 0) {
        if (tfork && nin > 0) {
          if ((rc = check_cuda(cudaEventRecord(aux.ev_panel, sc), __FILE__, __LINE__))) break;
          if ((rc = check_cuda(cudaStreamWaitEvent(aux.s2, aux.ev_panel, 0), __FILE__, __LINE__))) break;
          if ((rc = build_T<T>(w, w.Vj(j), w.ldv, mk, kk, dtau + k0, w.Tj(j), aux.s2))) break;
          if ((rc = check_cuda(cudaEventRecord(aux.ev_t, aux.s2), __FILE__, __LINE__))) break;
          t_forked = true;
        } else {
          if ((rc = build_T<T>(w, w.Vj(j), w.ldv, mk, kk, dtau + k0, w.Tj(j), sc))) break;
        }
      }
      if (nin > 0) {
        if ((rc = apply_panel<T>(w, w.Vj(j), w.ldv, w.VTj(j), NBO, w.Tj(j), mk, kk, Ak + (i64)nb * lda, lda, nin, 1, sc,
                                 t_forked ? aux.ev_t : nullptr)))
          break;
      }
      if (!(mk > nb && n - k0 > nb)) {  // reference recursion stops: src/qr.jl:136
        done = true;
        // columns beyond this panel (wide case) still get this outer block's reflectors below
        break;
      }
    }
    if (rc) break;
    const i64 nfar = n - (o0 + nbo);
    if (o0 + nbo >= n || o0 + nbo >= m) done = true;
    if (nfar > 0 && kbig > 0) {
      if (overlap) {
        if ((rc = check_cuda(cudaEventRecord(aux.ev[1], sc), __FILE__, __LINE__))) break;       // chain(o) done
        if ((rc = check_cuda(cudaStreamWaitEvent(st, aux.ev[1], 0), __FILE__, __LINE__))) break;
        if (sink && sink->copy && sink->copied_cols == o0) {   // columns o0 .. o0 + nbo are final: start their way home
          if ((rc = check_cuda(cudaStreamWaitEvent(sink->copy, aux.ev[1], 0), __FILE__, __LINE__))) break;
          if ((rc = check_cuda(cudaMemcpy2DAsync(sink->hA + o0 * sink->ldh, sink->ldh * sizeof(T), dA + o0 * lda, lda * sizeof(T),
                                                 m * sizeof(T), nbo, cudaMemcpyDeviceToHost, sink->copy),
                               __FILE__, __LINE__)))
            break;
          sink->copied_cols = o0 + nbo;
        }
      }
      // ---- far update of outer block o: Gram once, then the next outer block's columns first
      //      (the Gram does not depend on W = V^H A2: with the T fork it runs on the second side stream beside that product)
      cudaEvent_t g_ready = nullptr;
      if (tfork && kbig > NB) {
        if ((rc = check_cuda(cudaEventRecord(aux.ev_panel, st), __FILE__, __LINE__))) break;
        if ((rc = check_cuda(cudaStreamWaitEvent(aux.s2, aux.ev_panel, 0), __FILE__, __LINE__))) break;
        if ((rc = gram<T>(w, 1, w.V[b], w.ldv, mo, kbig, w.G[b], NBO, aux.s2))) break;
        if ((rc = check_cuda(cudaEventRecord(aux.ev_t, aux.s2), __FILE__, __LINE__))) break;
        g_ready = aux.ev_t;
      } else if ((rc = gram<T>(w, 1, w.V[b], w.ldv, mo, kbig, w.G[b], NBO, st))) {
        break;
      }
      T* A2 = dA + o0 + (o0 + nbo) * lda;
      const i64 nA = (overlap && !done && nfar > nbw) ? nbw : nfar;
      if ((rc = need_cols(st, up_waited_far, o0 + nbo + nA))) break;
      if ((rc = apply_outer<T>(w, b, mo, kbig, A2, lda, nA, st, g_ready))) break;
      if (overlap && !done) {
        if ((rc = check_cuda(cudaEventRecord(aux.ev[2], st), __FILE__, __LINE__))) break;       // far A(o) done
        if ((rc = check_cuda(cudaStreamWaitEvent(sc, aux.ev[2], 0), __FILE__, __LINE__))) break;
      }
      // the rest of the far update -- cut along the chunk boundaries of a streamed upload while chunks are outstanding
      const int ns_rest = nfar > nA ? outer_nsplit<T>(w, kbig, nfar - nA, mo) : 0;   // pieces keep the slicing of the whole
      for (i64 c0 = nA; c0 < nfar && !rc;) {
        i64 c1 = nfar;
        if (sink && up_waited_far < sink->up_chunks) {
          const i64 abs0 = o0 + nbo + c0;   // first column of this piece
          int c = up_waited_far;
          while (c < sink->up_chunks && sink->up_col[c + 1] <= abs0) ++c;   // chunk that holds abs0 (or the first one right of it)
          if (c < sink->up_chunks && sink->up_col[c + 1] - (o0 + nbo) < c1) c1 = sink->up_col[c + 1] - (o0 + nbo);
          if (c1 <= c0) c1 = nfar;
          rc = need_cols(st, up_waited_far, o0 + nbo + c1);
          if (rc) break;
        }
        rc = apply_outer<T>(w, b, mo, kbig, A2 + c0 * lda, lda, c1 - c0, st, nullptr, ns_rest);
        c0 = c1;
      }
      if (rc) break;
    }
  }
  if (overlap) {  // join: everything issued on the side stream is ordered before what follows on `st`
    cudaError_t e = cudaEventRecord(aux.ev[1], aux.s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, aux.ev[1], 0);
    if (!rc) rc = check_cuda(e, __FILE__, __LINE__);
  }
  w.release();
  return rc;
}

// TH (rows x rows, ld NB) = T^H for a rows x rows upper T (ld NB): the K-contiguous operand of Z = T Y
template <class T>
__global__ void conj_transpose_t_kernel(const T* __restrict__ Tm, T* __restrict__ TH, int rows) {
  for (int e = threadIdx.x; e < rows * rows; e += blockDim.x) {
    const int i = e / rows, l = e - i * rows;   // TH(l, i) = conj(T(i, l))
    TH[(i64)i * NB + l] = cj(Tm[(i64)l * NB + i]);
  }
}

// A2 <- (I - V T V^H) A2 for ONE outer block (non-adjoint twin of apply_outer): the panels act LAST TO FIRST,
//     Y_j = W_j - sum_{i>j} G_ji Z_i ,   Z_j = T_j Y_j ,   A2 -= V Z          (lmul!(H, A, M), src/householder.jl:82-115)
template <class T>
static int apply_outer_fwd(QrWork<T>& w, int b, i64 mo, int kbig, T* A2, i64 lda, i64 nA, T* TH, cudaStream_t st) {
  if (nA <= 0) return 0;
  GemmTN<T> g1;
  g1.At = w.V[b]; g1.ldat = w.ldv;
  g1.B = A2; g1.ldb = lda;
  g1.C = w.Wp[1]; g1.ldc = NBO;
  g1.M = kbig; g1.N = nA; g1.K = mo;
  g1.conj_a = 1;
  g1.yield_sms = w.yield_sms;
  g1.nsplit = wsplit_for<T>(kbig, nA, mo);
  g1.split_stride = (i64)NBO * nA;
  if ((i64)g1.nsplit * NBO * nA > w.wp_elems[1]) g1.nsplit = (int)(w.wp_elems[1] / ((i64)NBO * nA));
  if (g1.nsplit < 1) {
    set_error(GLA_ERR_INTERNAL, "W workspace too small", __FILE__, __LINE__);
    return GLA_ERR_INTERNAL;
  }
  GLA_TRY(gemm_tn<T>(g1, st));
  if (g1.nsplit > 1) GLA_TRY(sum_splits<T>(w.Wp[1], NBO, w.Wp[1], NBO, g1.split_stride, g1.nsplit, kbig, nA, st));
  const int nj = (kbig + NB - 1) / NB;
  for (int j = nj - 1; j >= 0; --j) {
    const int rows_j = (kbig - j * NB) < NB ? (kbig - j * NB) : NB;
    conj_transpose_t_kernel<T><<<1, 256, 0, st>>>(w.Tm[b] + (i64)j * NB * NB, TH + (i64)j * NB * NB, rows_j);
    GLA_CUDA(cudaGetLastError());
    if (j < nj - 1) {
      GemmTN<T> gy;   // conj(At(l, a)) = G(jNB + a, (j+1)NB + l): G is Hermitian, so At = G + (j+1)NB + jNB * ldg
      gy.At = w.G[b] + (i64)(j + 1) * NB + (i64)j * NB * NBO; gy.ldat = NBO;
      gy.B = w.Z[1] + (j + 1) * NB; gy.ldb = NBO;
      gy.C = w.Wp[1] + j * NB; gy.ldc = NBO;
      gy.M = rows_j; gy.N = nA; gy.K = kbig - (i64)(j + 1) * NB;
      gy.alpha = -1; gy.beta_one = 1; gy.conj_a = 1;
      GLA_TRY(gemm_tn<T>(gy, st));
    }
    GemmTN<T> gz;
    gz.At = TH + (i64)j * NB * NB; gz.ldat = NB;
    gz.B = w.Wp[1] + j * NB; gz.ldb = NBO;
    gz.C = w.Z[1] + j * NB; gz.ldc = NBO;
    gz.M = rows_j; gz.N = nA; gz.K = rows_j;
    gz.conj_a = 1;
    GLA_TRY(gemm_tn<T>(gz, st));
  }
  GemmTN<T> g2;
  g2.At = w.VT[b]; g2.ldat = NBO;
  g2.B = w.Z[1]; g2.ldb = NBO;
  g2.C = A2; g2.ldc = lda;
  g2.M = mo; g2.N = nA; g2.K = kbig;
  g2.alpha = -1;
  g2.yield_sms = w.yield_sms;
  g2.beta_one = 1;
  return gemm_tn<T>(g2, st);
}

// A <- Q A (adjoint = 0) or Q^H A (adjoint = 1), Q = H_1 .. H_k from (F, tau).  Outer blocks of NBO = 384 reflectors:
// per block the clean V / V^T, the per-panel T_j and the Gram V^H V are rebuilt from the factors, then ONE K = 384 pass
// over A (W = V^H A, the block recurrence on 64-row strips of W, A -= V Z) -- the same wide form as the far update of the
// factorisation instead of one K = 64 pass per panel (16 B of A traffic per 128 flops: HBM bound).
template <class T>
int ormqr_blocked_dev(const T* dF, i64 mF, i64 nF, i64 ldf, const T* dtau, T* dA, i64 mA, i64 nA, i64 lda,
                      int adjoint, cudaStream_t st) {
  if (mF != mA) return -7;  // DimensionMismatch (src/householder.jl:87,129)
  if (mF < 0 || nF < 0 || nA < 0) return -2;
  if (ldf < (mF > 1 ? mF : 1)) return -4;
  if (lda < (mA > 1 ? mA : 1)) return -9;
  const i64 k = mF < nF ? mF : nF;
  if (k == 0 || nA == 0) return 0;
  QrWork<T> w;
  GLA_TRY(w.alloc(mF, 1, nA, 1, st));
  T* TH = nullptr;
  int rc = 0;
  if (!adjoint) rc = pool_malloc(reinterpret_cast<void**>(&TH), (size_t)NBO * NB * sizeof(T), st);
  const i64 nob = (k + NBO - 1) / NBO;
  w.cur = 0;
  for (i64 ib = 0; ib < nob && !rc; ++ib) {
    // Q^H A applies the blocks first to last, Q A last to first
    const i64 o0 = (adjoint ? ib : nob - 1 - ib) * NBO;
    const int kbig = (int)((k - o0) < NBO ? (k - o0) : NBO);
    const i64 mo = mF - o0;
    for (int j = 0; j * NB < kbig && !rc; ++j) {
      const i64 k0 = o0 + (i64)j * NB;
      const i64 mk = mF - k0;
      const int kk = (int)((k - k0) < NB ? (k - k0) : NB);
      const unsigned grid = (unsigned)(ceil_div(mk * kk, 256) > 2048 ? 2048 : ceil_div(mk * kk, 256));
      extract_v_kernel<T><<<grid, 256, 0, st>>>(dF + k0 + k0 * ldf, ldf, (int)mk, kk, w.Vj(j), w.ldv, w.VTj(j), NBO);
      if ((rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__))) break;
      if (j > 0) {
        zero_top_kernel<T><<<ceil_div((i64)j * NB * kk, 256), 256, 0, st>>>(w.V[0], w.ldv, w.VT[0], j * NB, kk);
        if ((rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__))) break;
      }
      rc = build_T<T>(w, w.Vj(j), w.ldv, mk, kk, dtau + k0, w.Tj(j), st);
    }
    if (rc) break;
    if (kbig > NB) rc = gram<T>(w, 1, w.V[0], w.ldv, mo, kbig, w.G[0], NBO, st);
    if (rc) break;
    rc = adjoint ? apply_outer<T>(w, 0, mo, kbig, dA + o0, lda, nA, st)
                 : apply_outer_fwd<T>(w, 0, mo, kbig, dA + o0, lda, nA, TH, st);
  }
  if (TH) cudaFreeAsync(TH, st);
  w.release();
  return rc;
}

// thin Q (m x k, k = min(m, n)): Q = H_1 .. H_k [I_k; 0]                 (used to verify ||Q^H Q - I||)
template <class T>
__global__ void eye_kernel(T* __restrict__ Q, i64 ldq, i64 m, i64 k) {
  const i64 total = m * k;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 j = e / m, i = e - j * m;
    Q[j * ldq + i] = i == j ? Sc<T>::one() : Sc<T>::zero();
  }
}
template <class T>
int orgqr_thin_dev(const T* dF, i64 m, i64 n, i64 ldf, const T* dtau, T* dQ, i64 ldq, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (ldf < (m > 1 ? m : 1)) return -4;
  if (ldq < (m > 1 ? m : 1)) return -7;
  const i64 k = m < n ? m : n;
  if (k == 0) return 0;
  const i64 total = m * k;
  eye_kernel<T><<<(unsigned)(ceil_div(total, 256) > 4096 ? 4096 : ceil_div(total, 256)), 256, 0, st>>>(dQ, ldq, m, k);
  GLA_CUDA(cudaGetLastError());
  return ormqr_blocked_dev<T>(dF, m, n, ldf, dtau, dQ, m, k, ldq, 0, st);
}

// ------------------------------------------------------------------------------- full-width T (API parity)
// getindex(::QR2, Tuple{:QBlocked}) builds T for ALL k = min(m,n) reflectors (src/qr.jl:66-69).  The
// factorisation itself never needs it (it works panel by panel); this exists for the drop-in API.
template <class T>
__global__ void clean_v_kernel(const T* __restrict__ F, i64 ldf, i64 m, i64 k, T* __restrict__ Vc, i64 ldv) {
  const i64 total = m * k;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 col = e / m, i = e - col * m;
    Vc[col * ldv + i] = i < col ? Sc<T>::zero() : (i == col ? Sc<T>::one() : F[col * ldf + i]);
  }
}

// one CTA: Tm holds G = V^H V on entry.  U = diag(tau) striu(G); X = (I+U)^-1 by the column
// recurrence X[:,j] = e_j - X[:,0:j] U[0:j,j]; T = X diag(tau).
template <class T>
__global__ void __launch_bounds__(1024)
    larft_big_kernel(T* __restrict__ Tm, i64 ldt, int k, const T* __restrict__ tau, T* __restrict__ X) {
  for (i64 e = threadIdx.x; e < (i64)k * k; e += blockDim.x) {
    const int j = (int)(e / k), i = (int)(e - (i64)j * k);
    Tm[(i64)j * ldt + i] = i < j ? tau[i] * Tm[(i64)j * ldt + i] : Sc<T>::zero();
    X[(i64)j * k + i] = (i == j) ? Sc<T>::one() : Sc<T>::zero();
  }
  __syncthreads();
  for (int j = 1; j < k; ++j) {
    for (int i = threadIdx.x; i < j; i += blockDim.x) {
      T acc = Sc<T>::zero();
      for (int l = i; l < j; ++l) acc = fmad(X[(i64)l * k + i], Tm[(i64)j * ldt + l], acc);
      X[(i64)j * k + i] = -acc;
    }
    __syncthreads();
  }
  for (i64 e = threadIdx.x; e < (i64)k * k; e += blockDim.x) {
    const int j = (int)(e / k), i = (int)(e - (i64)j * k);
    Tm[(i64)j * ldt + i] = i <= j ? X[(i64)j * k + i] * tau[j] : Sc<T>::zero();
  }
}

template <class T>
int larft_dev(const T* dF, i64 m, i64 n, i64 ldf, const T* dtau, T* dT, i64 ldt, cudaStream_t st) {
  if (m < 0) return -2;
  if (n < 0) return -3;
  const i64 k = m < n ? m : n;
  if (k == 0) return 0;
  if (ldt < k) return -7;
  const i64 ldv = round_up(m, 2);
  T* Vc = nullptr;
  T* X = nullptr;
  GLA_TRY(pool_malloc(reinterpret_cast<void**>(&Vc), (size_t)ldv * k * sizeof(T), st));
  int rc = pool_malloc(reinterpret_cast<void**>(&X), (size_t)k * k * sizeof(T), st);
  if (!rc) {
    const unsigned grid = (unsigned)(ceil_div(m * k, 256) > 4096 ? 4096 : ceil_div(m * k, 256));
    clean_v_kernel<T><<<grid, 256, 0, st>>>(dF, ldf, m, k, Vc, ldv);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  if (!rc) {
    GemmTN<T> g;
    g.At = Vc; g.ldat = ldv;
    g.B = Vc; g.ldb = ldv;
    g.C = dT; g.ldc = ldt;
    g.M = k; g.N = k; g.K = m;
    g.conj_a = 1;
    rc = gemm_tn<T>(g, st);
  }
  if (!rc) {
    larft_big_kernel<T><<<1, 1024, 0, st>>>(dT, ldt, (int)k, dtau, X);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  cudaFreeAsync(Vc, st);
  if (X) cudaFreeAsync(X, st);
  return rc;
}

// ------------------------------------------------------------------------------- right reflector apply
// A <- A (I - tau v v^H), v = [1; x[2:]]     (src/qr.jl:19-42), one thread per row
template <class T>
__global__ void reflector_apply_right_kernel(T* __restrict__ A, i64 m, i64 n, i64 lda, const T* __restrict__ x, T tau) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  T s = A[i];
  for (i64 j = 1; j < n; ++j) s = fmad(A[i + j * lda], x[j], s);
  s = s * tau;
  A[i] = A[i] - s;
  for (i64 j = 1; j < n; ++j) A[i + j * lda] = A[i + j * lda] - s * cj(x[j]);
}

template <class T>
int reflector_apply_right_dev(T* dA, i64 m, i64 n, i64 lda, const T* dx, T tau, cudaStream_t st) {
  if (m <= 0 || n <= 0) return 0;
  reflector_apply_right_kernel<T><<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(dA, m, n, lda, dx, tau);
  GLA_CUDA(cudaGetLastError());
  return 0;
}

#define INST(T)                                                                                             \
  template int geqr_blocked_dev<T>(T*, i64, i64, i64, T*, i64, cudaStream_t, QrHostSink<T>*);                               \
  template int ormqr_blocked_dev<T>(const T*, i64, i64, i64, const T*, T*, i64, i64, i64, int, cudaStream_t); \
  template int orgqr_thin_dev<T>(const T*, i64, i64, i64, const T*, T*, i64, cudaStream_t);                        \
  template int larft_dev<T>(const T*, i64, i64, i64, const T*, T*, i64, cudaStream_t);                      \
  template int reflector_apply_right_dev<T>(T*, i64, i64, i64, const T*, T, cudaStream_t);
INST(float)
INST(double)
INST(zd)

}  // namespace gla
