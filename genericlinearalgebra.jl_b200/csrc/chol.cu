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

// ------------------------------------------------------------------------------- recursion (host)
template <class T>
struct CholCtx {
  T* W;      // n x n mirror (upper)
  i64 ldw;
  T* Uinv;   // one CB x CB inverse per diagonal block
  int* info;
  cudaStream_t st;
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

template <class T>
int potrf_recursive_L_dev(T* dA, i64 n, i64 lda, i64 /*cutoff*/, int* dinfo, cudaStream_t st) {
  if (n < 0) return -2;
  if (lda < (n > 1 ? n : 1)) return -3;
  if (n == 0) return 0;
  CholCtx<T> cx;
  cx.ldw = round_up(n, 16);
  cx.st = st;
  cx.info = dinfo;
  const i64 nblk = (n + CB - 1) / CB;
  void* block = nullptr;
  const i64 wbytes = round_up(cx.ldw * n * sizeof(T), 256);
  GLA_CUDA(cudaMallocAsync(&block, wbytes + nblk * CB * CB * sizeof(T), st));
  cx.W = static_cast<T*>(block);
  cx.Uinv = reinterpret_cast<T*>(static_cast<char*>(block) + wbytes);
  GLA_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), st));
  dim3 tb(32, 8), grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
  mirror_in_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
  int rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  if (!rc) rc = chol_rec<T>(cx, 0, n);
  if (!rc) {
    mirror_out_kernel<T><<<grid, tb, 0, st>>>(dA, lda, cx.W, cx.ldw, (int)n);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  cudaFreeAsync(block, st);
  return rc;
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
  GLA_CUDA(cudaMallocAsync(&P, (size_t)ldp * n * sizeof(T), st));
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
  template int potrf_recursive_L_dev<T>(T*, i64, i64, i64, int*, cudaStream_t);  \
  template int herk_lower_dev<T>(T*, i64, i64, const T*, i64, i64, typename Sc<T>::real, cudaStream_t);
INST(float)
INST(double)
INST(zd)

}  // namespace gla
