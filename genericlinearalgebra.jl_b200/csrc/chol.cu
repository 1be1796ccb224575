// chol.cu -- K6: recursive lower Cholesky and the Hermitian rank-k trailing update.
//
// Reference semantics: cholRecursive!(A, Val{:L}) src/cholesky.jl:37-55 =
//   chol(A11); A21 <- A21 L11^-H (rdiv!, :48); A22(lower) -= A21 A21^H (rankUpdate!, :51 ->
//   src/juliaBLAS.jl:89-112); chol(A22).  Only the lower triangle of A is read or written.
//
// GPU formulation: every contraction of the recursion is expressed in the one TN form the tensor-pipe
// kernel implements (K contiguous in both operands, gemm.cu) by factorising the MIRROR image
// W = (lower triangle of A)^H, i.e. computing the upper factor U = L^H with W = U^H U:
//     U11 = chol(W11)
//     U12 = U11^-H W12          recursive left solve; its updates are  W12_2 -= Ub^H Y1   (TN, conj A)
//     W22 -= U12^H U12          Hermitian rank-k update, upper tiles only                  (TN, conj A)
//     U22 = chol(W22)
// 64x64 diagonal blocks are factorised AND inverted by one CTA in shared memory; the base-case solve
// is then the TN product  Y = (Ud^-1)^H Wd, done in place (one M tile reads exactly the columns it
// writes).  At the end L = U^H is written back into the lower triangle of A; the strict upper
// triangle of A is never touched, exactly like the reference.
#include "fastmath.cuh"
#include "gemm.cuh"
#include "gla_internal.cuh"

#include <stdlib.h>

namespace gla {

constexpr int CB = 64;  // diagonal block

// ------------------------------------------------------------------------------- mirror copies
// W(i,j) = conj(A(j,i)) for i <= j  (upper triangle of the n x n mirror; strict lower zeroed)
template <class T>
__global__ void mirror_in_kernel(const T* __restrict__ A, i64 lda, T* __restrict__ W, i64 ldw, int n) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32;  // W tile origin (rows bi, cols bj)
  if (bi > bj + 31) {                                    // strictly lower tile of W: zero
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int i = bi + threadIdx.x, j = bj + r;
      if (i < n && j < n) W[(i64)j * ldw + i] = Sc<T>::zero();
    }
    return;
  }
  // read A tile rows bj.., cols bi.. (A(j,i)), coalesced along j
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = bj + threadIdx.x, i = bi + r;
    tile[r][threadIdx.x] = (i < n && j < n && j >= i) ? cj(A[(i64)i * lda + j]) : Sc<T>::zero();
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bi + threadIdx.x, j = bj + r;
    if (i < n && j < n) W[(i64)j * ldw + i] = tile[threadIdx.x][r];
  }
}

// A(i,j) = conj(U(j,i)) for i >= j; strict upper triangle of A untouched
template <class T>
__global__ void mirror_out_kernel(T* __restrict__ A, i64 lda, const T* __restrict__ U, i64 ldu, int n) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.x * 32, bj = blockIdx.y * 32;  // A tile origin
  if (bj > bi + 31) return;                              // strictly upper tile of A
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = bj + threadIdx.x, i = bi + r;          // U(j,i), coalesced along j
    tile[r][threadIdx.x] = (i < n && j < n) ? cj(U[(i64)i * ldu + j]) : Sc<T>::zero();
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bi + threadIdx.x, j = bj + r;
    if (i < n && j < n && i >= j) A[(i64)j * lda + i] = tile[threadIdx.x][r];
  }
}

// the same for the columns [c_lo, c_hi) of A only (c_lo a multiple of 32): grid.y counts column tiles from c_lo
template <class T>
__global__ void mirror_out_cols_kernel(T* __restrict__ A, i64 lda, const T* __restrict__ U, i64 ldu, int n, int c_lo, int c_hi) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.x * 32, bj = c_lo + blockIdx.y * 32;  // A tile origin
  if (bj > bi + 31 || bj >= c_hi) return;                        // strictly upper tile of A / beyond the range
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = bj + threadIdx.x, i = bi + r;
    tile[r][threadIdx.x] = (i < n && j < n) ? cj(U[(i64)i * ldu + j]) : Sc<T>::zero();
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bi + threadIdx.x, j = bj + r;
    if (i < n && j < c_hi && i >= j) A[(i64)j * lda + i] = tile[threadIdx.x][r];
  }
}

// P (k x n, ldp) = A^H for A n x k (lda)   (conjugate transpose, K-contiguous operand of a rank-k update)
template <class T>
__global__ void conj_transpose_kernel(const T* __restrict__ A, i64 lda, int n, int k, T* __restrict__ P, i64 ldp) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.x * 32, bl = blockIdx.y * 32;  // A tile: rows bi.., cols bl..
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bi + threadIdx.x, l = bl + r;
    tile[r][threadIdx.x] = (i < n && l < k) ? cj(A[(i64)l * lda + i]) : Sc<T>::zero();
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = bl + threadIdx.x, i = bi + r;
    if (i < n && l < k) P[(i64)i * ldp + l] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------- diagonal block
// One CTA: upper Cholesky of the nb x nb (<= 64) block at W (ld ldw) in place, plus its inverse into
// Uinv (CB x CB, ld CB, upper, zero below).  Non-positive pivot -> *info = offset + index + 1 (first wins).
//   factorisation: the block lives in REGISTERS, thread (ty, tx) of a 16 x 16 grid owns the cyclic 4 x 4 sub-grid
//     S(ty + 16a, tx + 16b); per column step the owners of row j publish the un-scaled row (double-buffered by
//     step parity, ONE __syncthreads per step), every thread derives d = sqrt(pivot) itself and applies the rank-1
//     update to its 16 entries;
//   inverse: back substitution X(i,c) = -(sum_{l=i+1..c} U(i,l) X(l,c)) / U(i,i), four lanes of ONE warp per column,
//     so the 63 dependent steps are separated by __syncwarp only.
template <class T>
__global__ void __launch_bounds__(256) potrf_block_kernel(T* __restrict__ W, i64 ldw, int nb, T* __restrict__ Uinv,
                                                          int* __restrict__ info, int offset) {
  using R = typename Sc<T>::real;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef T Row[CB + 1];
  Row* S = reinterpret_cast<Row*>(smem_raw);  // S[i][j]
  Row* X = S + CB;
  __shared__ T rowbuf[2][CB];
  __shared__ R dinv[CB];   // 1 / U(j,j)
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  T s[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, c = tx + 16 * b;
      s[a][b] = (i <= c && c < nb) ? W[(i64)c * ldw + i] : Sc<T>::zero();
    }
  int bad = 0;
  for (int j = 0; j < nb; ++j) {
    T* rb = rowbuf[j & 1];
    if (ty == (j & 15)) {
      const int ja = j >> 4;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const T v01 = ja == 0 ? s[0][b] : s[1][b];
        const T v23 = ja == 2 ? s[2][b] : s[3][b];
        rb[tx + 16 * b] = ja < 2 ? v01 : v23;
      }
    }
    __syncthreads();
    const R piv = re(rb[j]);
    if (!(piv > R(0))) {  // uniform: every thread reads the same pivot
      bad = j + 1;
      break;
    }
    // sqrt and reciprocal sqrt from the MUFU seed + FMA refinement (fastmath.cuh): every thread needs both, and the
    // IEEE sqrt / division subroutines would sit on the critical path of all 64 dependent steps
    const R rd = Fast<R>::rsqrt(piv);
    const R d = Fast<R>::sqrt_from_rsqrt(piv, rd);
    if (tid == 0) dinv[j] = rd;
    T ui[4], uc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ui[a] = cj(scale_real(rb[ty + 16 * a], rd));
#pragma unroll
    for (int b = 0; b < 4; ++b) uc[b] = scale_real(rb[tx + 16 * b], rd);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ty + 16 * a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = tx + 16 * b;
        if (i > j && c >= i) s[a][b] = s[a][b] - ui[a] * uc[b];                   // trailing upper triangle
        else if (i == j && c >= j) s[a][b] = c == j ? Sc<T>::from_real(d) : uc[b];  // row j of the factor
      }
    }
  }
  if (bad) {
    if (tid == 0) atomicCAS(info, 0, offset + bad);
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, c = tx + 16 * b;
      S[i][c] = s[a][b];
      X[i][c] = Sc<T>::zero();
      if (i <= c && c < nb) W[(i64)c * ldw + i] = s[a][b];
    }
  __syncthreads();
  // inverse of the upper factor, column c by lanes 4c .. 4c+3 of one warp
  const int c = tid >> 2, part = tid & 3;
  if (c < nb && part == 0) X[c][c] = Sc<T>::from_real(dinv[c]);
  __syncwarp();
  for (int i = nb - 2; i >= 0; --i) {
    T acc = Sc<T>::zero(), acc2 = Sc<T>::zero();
    if (c < nb && c > i) {
      int l = i + 1 + part;
      for (; l + 4 <= c; l += 8) {   // two independent accumulators hide the LDS -> FMA latency
        acc = fmad(S[i][l], X[l][c], acc);
        acc2 = fmad(S[i][l + 4], X[l + 4][c], acc2);
      }
      if (l <= c) acc = fmad(S[i][l], X[l][c], acc);
    }
    acc = acc + acc2;
    acc = acc + shfl_xor_t<T>(acc, 1);
    acc = acc + shfl_xor_t<T>(acc, 2);
    if (c < nb && c > i && part == 0) X[i][c] = scale_real(-acc, dinv[i]);
    __syncwarp();
  }
  __syncthreads();
  for (int e = tid; e < CB * CB; e += blockDim.x) {
    const int j = e / CB, i = e - j * CB;
    Uinv[(i64)j * CB + i] = (i <= j && j < nb) ? X[i][j] : Sc<T>::zero();
  }
}

// ------------------------------------------------------------------------------- fused row-panel kernel (real types)
// Right-looking step for the row panel U(r : r+nb, r : n) of the mirror (nb <= 64).  The 64 x 64 diagonal block is the
// sequential part of the whole factorisation (4096 dependent column steps at n = 4096), so it must not also pay kernel
// boundaries: EVERY CTA of the launch factorises the diagonal block redundantly (same arithmetic, so bitwise the same
// factor; no inter-CTA synchronisation) and solves ITS OWN 32-column slice of the row panel in the same elimination,
//     Y = U_D^-H * W(r : r+nb, slice)            (rdiv!, src/cholesky.jl:48: forward substitution, one row per step).
// kpend > 0: the panel has not yet received the update of the finished rows rp : rp+kpend right above it (the previous
// panel of its outer block and, for 128-row outer blocks, the whole previous outer block: kpend <= 192):
//     W(r : r+nb, r : n) -= U(rp : rp+kpend, r : r+nb)^H * U(rp : rp+kpend, r : n)     (rankUpdate!, :51)
// is applied first, in chunks of 64 rows, to the diagonal block by every CTA and to the slice by its owner.  The factor of the diagonal block
// goes to a scratch copy (Ud), not into W: other CTAs of the same launch may still be reading the block.
constexpr int SW = 32;  // slice width

template <class T>
struct PanelSmem {
  using R = typename Sc<T>::real;
  T P[CB][CB + 1];     // pending operand U(rp + k, r + c) as P[k][c]
  T Ys[CB][SW + 1];    // slice of the row panel
  T Ps[CB][SW + 1];    // pending operand of the slice
  T rowbuf[2][2 * (CB + SW)];   // rows j, j + 1 of the diagonal block and of the slice, by pair parity
  R dinv[CB];
};

// upper Cholesky of the block held in registers: thread (ty, tx) of a 16 x 16 grid owns the cyclic 4 x 4 sub-grid
// S(ty + 16a, tx + 16b).  Returns 0 or (index of the non-positive pivot) + 1 (uniform over the CTA).
// LDL = true: the unit upper factor of W = U^H D U instead (ldlt!, src/ldlt.jl: no square roots, D real and of either sign
// on the diagonal, U(j, c) = W(j, c) / d_j, trailing block -= conj(W(j, i)) W(j, c) / d_j); dinv[j] = 1 / d_j.
template <class T, bool LDL = false>
__device__ __forceinline__ int factor_block_regs(T (&s)[4][4], T (&ys)[4][2], const int nb, T (*rowbuf)[2 * (CB + SW)],
                                                 typename Sc<T>::real* dinv) {
  using R = typename Sc<T>::real;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  int j = 0;
  if constexpr (!LDL) {
    // TWO pivot rows per barrier: rows j and j + 1 are published as they stand; after the barrier every thread eliminates row j
    // from row j + 1 redundantly (the entries it needs: its four columns, its four rows, its two slice columns), so that both
    // reflectors of the pair are known without a second barrier / shared-memory round trip, and applies a rank-2 update.
    // Same arithmetic per entry as two single steps (row j + 1 is updated by row j exactly as the general update does it).
    for (; j + 1 < nb; j += 2) {
      T* rb0 = rowbuf[(j >> 1) & 1];
      T* rb1 = rb0 + (CB + SW);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jr = j + h;
        if (ty == (jr & 15)) {
          const int ja = jr >> 4;
          T* rb = h ? rb1 : rb0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const T v01 = ja == 0 ? s[0][b] : s[1][b];
            const T v23 = ja == 2 ? s[2][b] : s[3][b];
            rb[tx + 16 * b] = ja < 2 ? v01 : v23;
          }
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const T v01 = ja == 0 ? ys[0][b] : ys[1][b];
            const T v23 = ja == 2 ? ys[2][b] : ys[3][b];
            rb[CB + tx + 16 * b] = ja < 2 ? v01 : v23;
          }
        }
      }
      __syncthreads();
      const R piv0 = re(rb0[j]);
      if (!(piv0 > R(0))) return j + 1;
      const R rd0 = Fast<R>::rsqrt(piv0);
      T w0[4], uc0[4], us0[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) w0[a] = scale_real(rb0[ty + 16 * a], rd0);
#pragma unroll
      for (int b = 0; b < 4; ++b) uc0[b] = scale_real(rb0[tx + 16 * b], rd0);
#pragma unroll
      for (int b = 0; b < 2; ++b) us0[b] = scale_real(rb0[CB + tx + 16 * b], rd0);
      const T u01 = scale_real(rb0[j + 1], rd0);      // U(j, j + 1)
      const T m = cj(u01);
      const R piv1 = re(rb1[j + 1] - m * u01);
      if (!(piv1 > R(0))) return j + 2;
      const R rd1 = Fast<R>::rsqrt(piv1);
      if (tid == 0) {
        dinv[j] = rd0;
        dinv[j + 1] = rd1;
      }
      T ui0[4], ui1[4], uc1[4], us1[2];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        ui0[a] = (ty + 16 * a > j) ? cj(w0[a]) : Sc<T>::zero();
        ui1[a] = (ty + 16 * a > j + 1) ? cj(scale_real(rb1[ty + 16 * a] - m * w0[a], rd1)) : Sc<T>::zero();
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) uc1[b] = scale_real(rb1[tx + 16 * b] - m * uc0[b], rd1);
#pragma unroll
      for (int b = 0; b < 2; ++b) us1[b] = scale_real(rb1[CB + tx + 16 * b] - m * us0[b], rd1);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = (s[a][b] - ui0[a] * uc0[b]) - ui1[a] * uc1[b];
#pragma unroll
        for (int b = 0; b < 2; ++b) ys[a][b] = (ys[a][b] - ui0[a] * us0[b]) - ui1[a] * us1[b];
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jr = j + h;
        if (ty == (jr & 15)) {   // rows j, j + 1 of the factor (and of the solved slice)
          const int ja = jr >> 4;
          const R d = h ? Fast<R>::sqrt_from_rsqrt(piv1, rd1) : Fast<R>::sqrt_from_rsqrt(piv0, rd0);
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (a == ja) {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const int c = tx + 16 * b;
                s[a][b] = c == jr ? Sc<T>::from_real(d) : (h ? uc1[b] : uc0[b]);
              }
#pragma unroll
              for (int b = 0; b < 2; ++b) ys[a][b] = h ? us1[b] : us0[b];
            }
        }
      }
    }
  }
  int par = (j >> 1) & 1;   // the buffer the last pair did NOT use; single steps (LDL, odd tail) alternate from there
  for (; j < nb; ++j, par ^= 1) {
    T* rb = rowbuf[par];
    if (ty == (j & 15)) {
      const int ja = j >> 4;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const T v01 = ja == 0 ? s[0][b] : s[1][b];
        const T v23 = ja == 2 ? s[2][b] : s[3][b];
        rb[tx + 16 * b] = ja < 2 ? v01 : v23;
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {   // row j of the slice rides along: the solve is the same elimination
        const T v01 = ja == 0 ? ys[0][b] : ys[1][b];
        const T v23 = ja == 2 ? ys[2][b] : ys[3][b];
        rb[CB + tx + 16 * b] = ja < 2 ? v01 : v23;
      }
    }
    __syncthreads();
    const R piv = re(rb[j]);
    if (LDL ? !(piv != R(0)) : !(piv > R(0))) return j + 1;  // uniform: every thread reads the same pivot
    const R rd = LDL ? Fast<R>::rcp(piv) : Fast<R>::rsqrt(piv);
    const R d = LDL ? piv : Fast<R>::sqrt_from_rsqrt(piv, rd);
    if (tid == 0) dinv[j] = rd;
    // Entries strictly below the diagonal (c < i) are never read by anybody, so the rank-1 update only has to be
    // masked by ROW (i > j): four selects per step instead of a predicate per entry.
    T ui[4], uc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
      ui[a] = (ty + 16 * a > j) ? cj(LDL ? rb[ty + 16 * a] : scale_real(rb[ty + 16 * a], rd)) : Sc<T>::zero();
#pragma unroll
    for (int b = 0; b < 4; ++b) uc[b] = scale_real(rb[tx + 16 * b], rd);
    T us[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) us[b] = scale_real(rb[CB + tx + 16 * b], rd);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = s[a][b] - ui[a] * uc[b];
#pragma unroll
      for (int b = 0; b < 2; ++b) ys[a][b] = ys[a][b] - ui[a] * us[b];
    }
    if (ty == (j & 15)) {   // row j of the factor: U(j, c) = W(j, c) / d for c > j, d on the diagonal
      const int ja = j >> 4;
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (a == ja) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int c = tx + 16 * b;
            s[a][b] = c == j ? Sc<T>::from_real(d) : uc[b];
          }
          if (!LDL) {   // Cholesky: row j of the slice is final (U); LDL keeps Y = D U and scales at the store
#pragma unroll
            for (int b = 0; b < 2; ++b) ys[a][b] = us[b];
          }
        }
    }
  }
  return 0;
}

// LDL = true: W = U^H D U.  Yw (same shape as W) receives Y = D U for the row panel, the left operand of the trailing
// updates (W22 -= Y12^H U12); the pending update scales the rows of the pending operand by their d (read from the diagonal
// of the pending panel's factor block Udp).
template <class T, bool LDL = false>
__global__ void __launch_bounds__(256) chol_panel_kernel(T* __restrict__ W, i64 ldw, int n, int r, int nb, int rp, int kpend,
                                                         T* __restrict__ Ud, int* __restrict__ info,
                                                         T* __restrict__ Yw = nullptr, const T* __restrict__ Udp = nullptr) {
  using R = typename Sc<T>::real;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PanelSmem<T>& sm = *reinterpret_cast<PanelSmem<T>*>(smem_raw);
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int c0 = r + nb + SW * blockIdx.x;            // first column of this CTA's slice
  const int sw = min(SW, n - c0);                     // <= 0: no slice (the panel is the last one)
  pdl_wait();                // programmatic dependent launch (common.cuh): nothing global is touched before this wait
  // ---- loads: diagonal block (registers; identity padding beyond nb), pending operands, slice
  T s[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = ty + 16 * a, c = tx + 16 * b;
      s[a][b] = (i <= c && c < nb) ? W[(i64)(r + c) * ldw + r + i] : ((i == c && i >= nb) ? Sc<T>::one() : Sc<T>::zero());
    }
  const int k_ = tid & (CB - 1), cq = tid >> 6;   // staging element e = tid + 256 q  ->  row k_, column cq + 4 q
  {  // all global loads are issued before the first shared-memory store (one round trip, not one per element)
    T vy[CB * SW / 256];
#pragma unroll
    for (int q = 0; q < CB * SW / 256; ++q) {
      const int c = cq + 4 * q;
      vy[q] = (c < sw && k_ < nb) ? W[(i64)(c0 + c) * ldw + r + k_] : Sc<T>::zero();
    }
#pragma unroll
    for (int q = 0; q < CB * SW / 256; ++q) sm.Ys[k_][cq + 4 * q] = vy[q];
  }
  __syncthreads();
  // ---- pending rank-kpend update in chunks of 64 rows: diagonal block (upper 16 x 16 sub-blocks only) and slice
  T ys[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) ys[a][b] = sm.Ys[ty + 16 * a][tx + 16 * b];
  for (int kc = 0; kc < kpend; kc += CB) {
    const int kn = min(CB, kpend - kc);
    {
      T vp[CB * CB / 256], vs[CB * SW / 256];
#pragma unroll
      for (int q = 0; q < CB * CB / 256; ++q) {
        const int c = cq + 4 * q;
        vp[q] = (k_ < kn && c < nb) ? W[(i64)(r + c) * ldw + rp + kc + k_] : Sc<T>::zero();
      }
#pragma unroll
      for (int q = 0; q < CB * SW / 256; ++q) {
        const int c = cq + 4 * q;
        vs[q] = (c < sw && k_ < kn) ? W[(i64)(c0 + c) * ldw + rp + kc + k_] : Sc<T>::zero();
      }
      if (kc > 0) __syncthreads();   // the previous chunk has been consumed
#pragma unroll
      for (int q = 0; q < CB * CB / 256; ++q) sm.P[k_][cq + 4 * q] = vp[q];
#pragma unroll
      for (int q = 0; q < CB * SW / 256; ++q) sm.Ps[k_][cq + 4 * q] = vs[q];
    }
    __syncthreads();
    for (int k = 0; k < kn; ++k) {
      T ui[4], uc[4], us[2];
      const R dk = LDL ? re(Udp[(i64)(kc + k) * CB + kc + k]) : R(1);   // LDL: (D U)^H U, d of the pending row
#pragma unroll
      for (int a = 0; a < 4; ++a) ui[a] = cj(LDL ? scale_real(sm.P[k][ty + 16 * a], dk) : sm.P[k][ty + 16 * a]);
#pragma unroll
      for (int b = 0; b < 4; ++b) uc[b] = sm.P[k][tx + 16 * b];
#pragma unroll
      for (int b = 0; b < 2; ++b) us[b] = sm.Ps[k][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (a <= b) s[a][b] = s[a][b] - ui[a] * uc[b];
#pragma unroll
        for (int b = 0; b < 2; ++b) ys[a][b] = ys[a][b] - ui[a] * us[b];
      }
    }
  }
  // ---- factorisation of the diagonal block (registers), redundantly in every CTA
  // The slice solve Y = U_D^-H Ys is carried out INSIDE the elimination: row j of the slice is scaled and eliminated from
  // the rows below it together with row j of the diagonal block (same 64 steps, same barriers).  The earlier form -- explicit
  // inverse of U_D by recursive doubling, then a 64-term product per slice entry -- cost 17,100 + 6,900 of the kernel's
  // 65,300 cycles (clock64 trace, n = 4096).
  const int bad = factor_block_regs<T, LDL>(s, ys, nb, sm.rowbuf, sm.dinv);
  if (bad) {
    if (blockIdx.x == 0 && tid == 0) atomicCAS(info, 0, r + bad);
    return;
  }
  __syncthreads();   // dinv complete
  pdl_launch_dependents();   // only the stores are left: the next kernel of the chain may be scheduled (it waits for them)
  if (blockIdx.x == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 16 * a, c = tx + 16 * b;
        if (i <= c && c < nb) Ud[(i64)c * CB + i] = s[a][b];
      }
  }
  if (sw > 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int c = tx + 16 * b;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ty + 16 * a;
        if (c < sw && i < nb) {
          if (LDL) {   // ys = (D U)(i, c): keep it for the trailing updates, store U = D^-1 ys
            Yw[(i64)(c0 + c) * ldw + r + i] = ys[a][b];
            W[(i64)(c0 + c) * ldw + r + i] = scale_real(ys[a][b], sm.dinv[i]);
          } else {
            W[(i64)(c0 + c) * ldw + r + i] = ys[a][b];
          }
        }
      }
    }
  }
}

// the diagonal blocks factorised by chol_panel_kernel go back into the mirror (upper triangles)
template <class T>
__global__ void __launch_bounds__(256) put_diag_blocks_kernel(T* __restrict__ W, i64 ldw, int n, const T* __restrict__ Ud, int b0 = 0) {
  const int blk = b0 + blockIdx.x;   // diagonal block index
  const int r = blk * CB;
  const int nb = min(CB, n - r);
  const T* U = Ud + (i64)blk * CB * CB;
  for (int e = threadIdx.x; e < CB * CB; e += 256) {
    const int i = e & (CB - 1), c = e >> 6;
    if (i <= c && c < nb) W[(i64)(r + c) * ldw + r + i] = U[(i64)c * CB + i];
  }
}

// ------------------------------------------------------------------------------- recursion (host)
template <class T>
struct CholCtx {
  T* W;      // n x n mirror (upper)
  i64 ldw;
  T* Uinv;   // one CB x CB inverse per diagonal block
  int* info;
  cudaStream_t st;
  T* Y = nullptr;   // LDL only: D U, same shape as W
  T* dA = nullptr;  // user-layout matrix and host sink of the host-pointer path (see CholHostSink)
  i64 lda = 0;
  CholHostSink<T>* sink = nullptr;
};

static i64 split_point(i64 n) {
  // first half: a multiple of CB close to n/2 (the reference uses div(n,2); results agree up to rounding)
  i64 h = (n / 2 + CB - 1) / CB * CB;
  if (h >= n) h = n - CB > 0 ? (n - 1) / CB * CB : n / 2;
  if (h <= 0) h = n / 2;
  return h;
}

// Y = U11^-H * Wb where U11 is the k x k upper factor at (r0,r0) and Wb = W(r0:r0+k, c0:c0+nc), in place
template <class T>
static int solve_rec(CholCtx<T>& cx, i64 r0, i64 k, i64 c0, i64 nc) {
  if (k <= CB) {
    GemmTN<T> g;
    g.At = cx.Uinv + (r0 / CB) * CB * CB; g.ldat = CB;
    g.B = cx.W + r0 + c0 * cx.ldw; g.ldb = cx.ldw;
    g.C = cx.W + r0 + c0 * cx.ldw; g.ldc = cx.ldw;
    g.M = k; g.N = nc; g.K = k;
    g.conj_a = 1;
    return gemm_tn<T>(g, cx.st);
  }
  const i64 k1 = split_point(k);
  GLA_TRY(solve_rec<T>(cx, r0, k1, c0, nc));
  // W(r0+k1 : r0+k, cols) -= Ub^H Y1,  Ub = U(r0 : r0+k1, r0+k1 : r0+k)
  GemmTN<T> g;
  g.At = cx.W + r0 + (r0 + k1) * cx.ldw; g.ldat = cx.ldw;
  g.B = cx.W + r0 + c0 * cx.ldw; g.ldb = cx.ldw;
  g.C = cx.W + (r0 + k1) + c0 * cx.ldw; g.ldc = cx.ldw;
  g.M = k - k1; g.N = nc; g.K = k1;
  g.alpha = -1; g.beta_one = 1; g.conj_a = 1;
  GLA_TRY(gemm_tn<T>(g, cx.st));
  return solve_rec<T>(cx, r0 + k1, k - k1, c0, nc);
}

template <class T>
static int chol_rec(CholCtx<T>& cx, i64 r0, i64 n) {
  if (n <= CB) {
    const int smem = 2 * CB * (CB + 1) * (int)sizeof(T);
    GLA_TRY(ensure_dyn_smem((const void*)potrf_block_kernel<T>, (int)(smem)));
    potrf_block_kernel<T><<<1, 256, smem, cx.st>>>(cx.W + r0 + r0 * cx.ldw, cx.ldw, (int)n,
                                                   cx.Uinv + (r0 / CB) * CB * CB, cx.info, (int)r0);
    GLA_CUDA(cudaGetLastError());
    return 0;
  }
  const i64 n1 = split_point(n), n2 = n - n1;
  GLA_TRY(chol_rec<T>(cx, r0, n1));
  GLA_TRY(solve_rec<T>(cx, r0, n1, r0 + n1, n2));                 // U12 = U11^-H W12     (rdiv!, :48)
  GemmTN<T> g;                                                     // W22 -= U12^H U12     (rankUpdate!, :51)
  g.At = cx.W + r0 + (r0 + n1) * cx.ldw; g.ldat = cx.ldw;
  g.B = g.At; g.ldb = cx.ldw;
  g.C = cx.W + (r0 + n1) + (r0 + n1) * cx.ldw; g.ldc = cx.ldw;
  g.M = n2; g.N = n2; g.K = n1;
  g.alpha = -1; g.beta_one = 1; g.conj_a = 1; g.lower_only = 2;
  GLA_TRY(gemm_tn<T>(g, cx.st));
  return chol_rec<T>(cx, r0 + n1, n2);
}

// Right-looking driver on the fused panel kernel (Float32 / Float64): outer blocks of two 64-row panels; the second
// panel takes the update of the first inside its own kernel (K = 64), everything beyond the outer block gets ONE
// K = 128 trailing update on the tensor pipe.  96 launches at n = 4096 instead of ~450 for the recursion, none of them
// a single-CTA kernel.  The result equals the recursion's (and the reference's) up to rounding: the same U^H U = W.
template <class T>
static bool use_panel_path() {
  static const bool off = [] {
    const char* e = getenv("GLA_CHOL_RECURSIVE");
    return e && atoi(e) != 0;
  }();
  static const bool zoff = getenv("GLA_CHOL_Z_RECURSIVE") != nullptr;   // A/B: ComplexF64 on the round-1 recursion
  return !off && !(Sc<T>::is_complex && zoff);   // (ComplexF64 fits since the explicit inverse left the kernel: 140 KB of tiles)
}

template <class T, bool LDL = false>
static int chol_right_looking(CholCtx<T>& cx, i64 n) {
  PdlScope pdl(true);   // programmatic dependent launch of the chain's kernels (common.cuh): n = 2048 1.33 -> 1.26 ms, n = 4096 3.13 -> 3.02 ms, n = 8192 10.9 -> 10.7 ms
  {
    const int smem = (int)sizeof(PanelSmem<T>);
    GLA_TRY(ensure_dyn_smem((const void*)chol_panel_kernel<T, LDL>, smem));
    static const int dbg_skip = [] {   // timing experiments only: 1 = no trailing updates, 2 = no panel kernels
      const char* e = getenv("GLA_CHOL_DEBUG_SKIP");
      return e ? atoi(e) : 0;
    }();
    auto panel = [&](i64 r, i64 nb, i64 rp, i64 kpend) -> int {
      if (dbg_skip == 2) return 0;
      const i64 ncols = n - (r + nb);
      const unsigned grid = (unsigned)(ncols > 0 ? ceil_div(ncols, SW) : 1);
      return check_cuda(launch_pdl(chol_panel_kernel<T, LDL>, dim3(grid), dim3(256), (size_t)smem, cx.st, cx.W, cx.ldw, (int)n,
                                   (int)r, (int)nb, (int)rp, (int)kpend, cx.Uinv + (r / CB) * CB * CB, cx.info, cx.Y,
                                   (const T*)(cx.Uinv + (rp / CB) * CB * CB)),
                        __FILE__, __LINE__);
    };
    // Look-ahead: the sequential chain (panel kernels + the update of the NEXT outer block's 128 rows) runs on a cached
    // high-priority stream `sc`; the bulk of every trailing update (rows beyond the next outer block) runs on the caller's
    // stream concurrently with the next outer block's panel kernels.  GLA_CHOL_NO_OVERLAP=1: one stream.
    static const bool no_overlap = getenv("GLA_CHOL_NO_OVERLAP") != nullptr;
    const bool overlap = !no_overlap && n > 6 * CB;
    cudaStream_t sc = cx.st;
    AuxCtx* aux = nullptr;
    if (overlap) {
      GLA_TRY(aux_ctx(&aux));
      sc = aux->hi;
      GLA_CUDA(cudaEventRecord(aux->ev[0], cx.st));          // mirror ready
      GLA_CUDA(cudaStreamWaitEvent(sc, aux->ev[0], 0));
    }
    const cudaStream_t caller = cx.st;
    auto update = [&](i64 r0, i64 K, i64 c0, i64 mrows, cudaStream_t s) -> int {   // W(c0 : c0+mrows, c0 : n) -= U(r0 : r0+K, .)^H U(r0 : r0+K, .)
      GemmTN<T> g;                                                                  // (rankUpdate!, :51), upper part only
      g.At = (LDL ? cx.Y : cx.W) + r0 + c0 * cx.ldw; g.ldat = cx.ldw;   // LDL: (D U)^H U
      g.B = cx.W + r0 + c0 * cx.ldw; g.ldb = cx.ldw;
      g.C = cx.W + c0 + c0 * cx.ldw; g.ldc = cx.ldw;
      g.M = mrows; g.N = n - c0; g.K = K;
      g.alpha = -1; g.beta_one = 1; g.conj_a = 1; g.lower_only = 2;
      // the bulk update leaves SMs to the chain on the high-priority stream while the chain is what bounds the run
      // (Float32: n = 8192 7.36 against 7.73 ms; at n = 16384 the bulk dominates and immortal CTAs win, 31.0 against 33.4 ms)
      g.yield_sms = (overlap && s == caller && n <= 8192) ? 1 : 0;
      return gemm_tn<T>(g, s);
    };
    // outer block: 128 rows while the panel chain is what bounds the run (n <= 4096), 256 beyond that, where the K = 128
    // trailing updates would (each one reads and writes the whole trailing matrix once)
    const i64 OB = n > 6144 ? 4 * CB : 2 * CB;
    // (Folding the K = 128 update of the next outer block's rows into its two panel kernels as a longer pending update was
    // measured slower: 5.58 ms against 4.81 ms at n = 4096 -- the in-kernel FMA update costs more than the tensor-pipe launch.)
    bool bulk_pending = false;
    for (i64 r0 = 0; r0 < n; r0 += OB) {
      i64 done = 0;
      cx.st = sc;
      for (i64 p0 = r0; p0 < n && p0 < r0 + OB; p0 += 2 * CB) {   // pairs of panels; the second takes the first in-kernel
        if (done > 0 && dbg_skip != 1)                             // rows of this pair <- the pairs before it in the block
          GLA_TRY(update(r0, done, p0, n - p0 < 2 * CB ? n - p0 : 2 * CB, sc));
        const i64 nbA = n - p0 < CB ? n - p0 : CB;
        GLA_TRY(panel(p0, nbA, p0, 0));
        done += nbA;
        if (n - p0 > CB) {
          const i64 nbB = n - p0 - CB < CB ? n - p0 - CB : CB;
          GLA_TRY(panel(p0 + CB, nbB, p0, CB));
          done += nbB;
        }
      }
      cx.st = caller;
      const i64 t0 = r0 + done, nt = n - t0;
      if (nt <= 0 || dbg_skip == 1) continue;
      if (!overlap) {
        GLA_TRY(update(r0, done, t0, nt, caller));
        continue;
      }
      const i64 mu = nt < OB ? nt : OB;                   // rows of the next outer block: needed by the chain right away
      GLA_CUDA(cudaEventRecord(aux->ev[1], sc));          // panels of this outer block done
      if (!LDL && cx.sink && cx.sink->copy && cx.sink->copied_cols == r0) {
        // rows r0 .. t0 of U are final: their diagonal blocks go back into the mirror, the matching columns of L are
        // mirrored into the user layout and travel home while the chain goes on (nobody else touches these parts)
        cudaStream_t cs = cx.sink->copy;
        GLA_CUDA(cudaStreamWaitEvent(cs, aux->ev[1], 0));
        put_diag_blocks_kernel<T><<<(unsigned)(done / CB), 256, 0, cs>>>(cx.W, cx.ldw, (int)n, cx.Uinv, (int)(r0 / CB));
        dim3 tb(32, 8), grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(done, 32));
        mirror_out_cols_kernel<T><<<grid, tb, 0, cs>>>(cx.dA, cx.lda, cx.W, cx.ldw, (int)n, (int)r0, (int)t0);
        GLA_CUDA(cudaGetLastError());
        GLA_CUDA(cudaMemcpy2DAsync(cx.sink->hA + r0 + r0 * cx.sink->ldh, cx.sink->ldh * sizeof(T), cx.dA + r0 + r0 * cx.lda,
                                   cx.lda * sizeof(T), (n - r0) * sizeof(T), done, cudaMemcpyDeviceToHost, cs));
        cx.sink->copied_cols = t0;
      }
      if (bulk_pending) GLA_CUDA(cudaStreamWaitEvent(sc, aux->ev[2], 0));   // those rows carry the previous bulk update
      GLA_TRY(update(r0, done, t0, mu, sc));
      if (nt > mu) {
        GLA_CUDA(cudaStreamWaitEvent(caller, aux->ev[1], 0));
        GLA_TRY(update(r0, done, t0 + mu, nt - mu, caller));
        GLA_CUDA(cudaEventRecord(aux->ev[2], caller));
        bulk_pending = true;
      }
    }
    cx.st = caller;
    if (overlap) {
      GLA_CUDA(cudaEventRecord(aux->ev[3], sc));
      GLA_CUDA(cudaStreamWaitEvent(caller, aux->ev[3], 0));
    }
    const i64 b0 = (!LDL && cx.sink) ? cx.sink->copied_cols / CB : 0;   // blocks before b0 went back (and home) already
    if (ceil_div(n, CB) > b0)
      put_diag_blocks_kernel<T><<<(unsigned)(ceil_div(n, CB) - b0), 256, 0, cx.st>>>(cx.W, cx.ldw, (int)n, cx.Uinv, (int)b0);
    return check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
}

template <class T>
int potrf_recursive_L_dev(T* dA, i64 n, i64 lda, i64 /*cutoff*/, int* dinfo, cudaStream_t st, CholHostSink<T>* sink) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  if (!dA) return -1;
  if (!dinfo) return -5;
  CholCtx<T> cx;
  cx.ldw = round_up(n, 16);
  cx.st = st;
  cx.info = dinfo;
  const i64 nblk = (n + CB - 1) / CB;
  void* block = nullptr;
  const i64 wbytes = round_up(cx.ldw * n * sizeof(T), 256);
  GLA_TRY(pool_malloc(reinterpret_cast<void**>(&block), wbytes + nblk * CB * CB * sizeof(T), st));
  cx.W = static_cast<T*>(block);
  cx.Uinv = reinterpret_cast<T*>(static_cast<char*>(block) + wbytes);
  dim3 tb(32, 8), grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
  int rc = check_cuda(cudaMemsetAsync(dinfo, 0, sizeof(int), st), __FILE__, __LINE__);
  if (!rc) {
    mirror_in_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  cx.dA = dA;
  cx.lda = lda;
  cx.sink = use_panel_path<T>() ? sink : nullptr;
  if (!rc) rc = use_panel_path<T>() ? chol_right_looking<T>(cx, n) : chol_rec<T>(cx, 0, n);
  if (!rc) {
    const i64 c_lo = cx.sink ? cx.sink->copied_cols : 0;   // columns already mirrored (and on their way to the host)
    if (c_lo == 0) {
      mirror_out_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
    } else if (c_lo < n) {
      dim3 g2((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n - c_lo, 32));
      mirror_out_cols_kernel<T><<<g2, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n, (int)c_lo, (int)n);
    }
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  cudaFreeAsync(block, st);
  return rc;
}

// W(i,j) = A(i,j) for i <= j (upper triangle as it is; strict lower zeroed) and back: the 'U' variant of ldlt!
template <class T>
__global__ void upper_in_kernel(const T* __restrict__ A, i64 lda, T* __restrict__ W, i64 ldw, int n) {
  const i64 total = (i64)n * n;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const int j = (int)(e / n), i = (int)(e - (i64)j * n);
    W[(i64)j * ldw + i] = i <= j ? A[(i64)j * lda + i] : Sc<T>::zero();
  }
}
template <class T>
__global__ void upper_out_kernel(T* __restrict__ A, i64 lda, const T* __restrict__ W, i64 ldw, int n) {
  const i64 total = (i64)n * n;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const int j = (int)(e / n), i = (int)(e - (i64)j * n);
    if (i <= j) A[(i64)j * lda + i] = W[(i64)j * ldw + i];
  }
}

// ldlt!(Hermitian(A, uplo), blocksize) without pivoting (src/ldlt.jl:80-162): in place, D on the diagonal, the unit
// factor in the strict `uplo` triangle, the other triangle untouched.  Real element types on the fused panel kernel
// (the mirror holds the unit UPPER factor of W = U^H D U; 'L': W = tril(A)^H, 'U': W = triu(A)).  A zero pivot stops
// the factorisation: *dinfo = its 1-based index (the reference would divide by it).
template <class T>
int ldlt_dev(T* dA, i64 n, i64 lda, int upper, int* dinfo, cudaStream_t st) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  if (!dA) return -1;
  if (!dinfo) return -5;
  {   // ComplexF64: Hermitian input, the imaginary part of the diagonal is ignored (D is real)
    CholCtx<T> cx;
    cx.ldw = round_up(n, 16);
    cx.st = st;
    cx.info = dinfo;
    const i64 nblk = (n + CB - 1) / CB;
    void* block = nullptr;
    const i64 wbytes = round_up(cx.ldw * n * sizeof(T), 256);
    GLA_TRY(pool_malloc(&block, 2 * wbytes + nblk * CB * CB * sizeof(T), st));
    cx.W = static_cast<T*>(block);
    cx.Y = reinterpret_cast<T*>(static_cast<char*>(block) + wbytes);
    cx.Uinv = reinterpret_cast<T*>(static_cast<char*>(block) + 2 * wbytes);
    dim3 tb(32, 8), grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
    const unsigned g1 = (unsigned)(ceil_div((i64)n * n, 256) > 4096 ? 4096 : ceil_div((i64)n * n, 256));
    int rc = check_cuda(cudaMemsetAsync(dinfo, 0, sizeof(int), st), __FILE__, __LINE__);
    if (!rc) {
      if (upper) upper_in_kernel<T><<<g1, 256, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
      else mirror_in_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
      rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
    }
    if (!rc) rc = chol_right_looking<T, true>(cx, n);
    if (!rc) {
      if (upper) upper_out_kernel<T><<<g1, 256, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
      else mirror_out_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
      rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
    }
    cudaFreeAsync(block, st);
    return rc;
  }
}

// C(lower) += alpha * A A^H,  A n x k (lda)      rankUpdate!(Hermitian(C,:L), A, alpha)
template <class T>
int herk_lower_dev(T* dC, i64 n, i64 ldc, const T* dA, i64 k, i64 lda, typename Sc<T>::real alpha,
                   cudaStream_t st) {
  if (n < 0) return -2;
  if (k < 0) return -5;
  if (n == 0 || k == 0) return 0;
  const i64 ldp = round_up(k, 16);
  T* P = nullptr;
  GLA_TRY(pool_malloc(reinterpret_cast<void**>(&P), (size_t)ldp * n * sizeof(T), st));
  dim3 tb(32, 8), grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(k, 32));
  conj_transpose_kernel<T><<<grid, tb, 0, st>>>(dA, lda, (int)n, (int)k, P, ldp);
  int rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  if (!rc) {
    GemmTN<T> g;                       // (A A^H)(i,j) = sum_l conj(P(l,i)) P(l,j),  P = A^H
    g.At = P; g.ldat = ldp;
    g.B = P; g.ldb = ldp;
    g.C = dC; g.ldc = ldc;
    g.M = n; g.N = n; g.K = k;
    g.alpha = alpha; g.beta_one = 1; g.conj_a = 1; g.lower_only = 1;
    rc = gemm_tn<T>(g, st);
  }
  cudaFreeAsync(P, st);
  return rc;
}

#define INST(T)                                                                  \
  template int potrf_recursive_L_dev<T>(T*, i64, i64, i64, int*, cudaStream_t, CholHostSink<T>*);  \
  template int herk_lower_dev<T>(T*, i64, i64, const T*, i64, i64, typename Sc<T>::real, cudaStream_t); \
  template int ldlt_dev<T>(T*, i64, i64, int, int*, cudaStream_t);
INST(float)
INST(double)
INST(zd)

}  // namespace gla
