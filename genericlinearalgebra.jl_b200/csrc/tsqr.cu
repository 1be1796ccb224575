// tsqr.cu -- K5: tall-skinny QR, R factor only, as a communication-avoiding tree.
//
// Semantics: the R that qrBlocked! (reference src/qr.jl:113-146) leaves in the upper triangle of a
// tall m x n matrix, n <= 64, up to the row signs discussed in DESIGN.md ("TSQR sign"): each tree
// node is a Householder QR with the reference's reflector! convention (smallqr.cuh), so every
// node's diagonal is -copysign(norm, pivot); the signs of the final R are those of the LAST node,
// which need not equal the signs sequential Householder on the whole matrix would give.
//
// Level 0: one CTA per chunk of ROWS0 rows (chunk staged in shared memory, Householder QR in place,
//          n x n R written out).  Level l>0: one CTA per group of FAN stacked R factors.
// Multi-GPU: each rank runs level 0.. on its row block (gla_dtsqr_local_dev), the ranks exchange
// their n x n R factors (NCCL all-gather, 32 KiB each) and every rank reduces the stack
// (gla_dtsqr_combine_dev).
#include "gla_internal.cuh"
#include "smallqr.cuh"

namespace gla {

constexpr int TSQR_MAXN = 64;

// chunk c = `nblk` consecutive blocks starting at block c*nblk; block b: rows_blk x n at src + b*bstride, ld
__global__ void __launch_bounds__(SMALLQR_THREADS)
    tsqr_node_kernel(const double* __restrict__ src, i64 ld, i64 bstride, i64 total_blocks, int rows_blk, i64 last_rows,
                     int nblk, int n, double* __restrict__ Rout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  const i64 c = blockIdx.x;
  const i64 b0 = c * nblk;
  i64 b1 = b0 + nblk;
  if (b1 > total_blocks) b1 = total_blocks;
  // rows of this chunk (only the globally last block may be short)
  i64 rows = (b1 - b0) * rows_blk;
  if (b1 == total_blocks) rows -= rows_blk - last_rows;
  const int lds = (int)(((i64)nblk * rows_blk) | 1);  // odd stride: conflict-free column walks
  const int R = (int)rows;
  for (i64 b = b0; b < b1; ++b) {
    const double* blk = src + b * bstride;
    const int rb = (b == total_blocks - 1) ? (int)last_rows : rows_blk;
    const int roff = (int)((b - b0) * rows_blk);
    for (int e = threadIdx.x; e < rb * n; e += blockDim.x) {
      const int j = e / rb, i = e - j * rb;
      sA[roff + i + j * lds] = blk[i + (i64)j * ld];
    }
  }
  __syncthreads();
  cta_qr_smem<double>(sA, R, n, lds, nullptr);
  double* out = Rout + c * (i64)n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int j = e / n, i = e - j * n;
    out[e] = (i <= j && i < R) ? sA[i + j * lds] : 0.0;
  }
}

static int run_tree(const double* src, i64 ld, i64 bstride, i64 total_blocks, int rows_blk, i64 last_rows, int n,
                    double* dR, i64 ldr, cudaStream_t st) {
  // workspace: ping-pong buffers of R stacks
  const int rows_target = n <= 16 ? 1024 : (n <= 32 ? 512 : 256);
  int nblk = rows_target / rows_blk;
  if (nblk < 2) nblk = 2;
  i64 chunks = (total_blocks + nblk - 1) / nblk;
  double* buf[2] = {nullptr, nullptr};
  GLA_CUDA(cudaMallocAsync(&buf[0], (size_t)chunks * n * n * sizeof(double), st));
  i64 chunks1 = (chunks + 3) / 4;
  int rc = check_cuda(cudaMallocAsync(&buf[1], (size_t)(chunks1 > 0 ? chunks1 : 1) * n * n * sizeof(double), st),
                      __FILE__, __LINE__);
  auto kern = tsqr_node_kernel;
  int cur = 0;
  if (!rc) rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), __FILE__, __LINE__);
  if (!rc) {
    size_t smem = (size_t)(((i64)nblk * rows_blk) | 1) * n * sizeof(double);
    kern<<<(unsigned)chunks, SMALLQR_THREADS, smem, st>>>(src, ld, bstride, total_blocks, rows_blk, last_rows, nblk, n,
                                                          buf[0]);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
  }
  i64 count = chunks;
  while (!rc && count > 1) {
    int fan = rows_target / n;
    if (fan < 2) fan = 2;
    const i64 next = (count + fan - 1) / fan;
    size_t smem = (size_t)(((i64)fan * n) | 1) * n * sizeof(double);
    kern<<<(unsigned)next, SMALLQR_THREADS, smem, st>>>(buf[cur], n, (i64)n * n, count, n, n, fan, n, buf[cur ^ 1]);
    rc = check_cuda(cudaGetLastError(), __FILE__, __LINE__);
    cur ^= 1;
    count = next;
  }
  if (!rc)
    rc = check_cuda(cudaMemcpy2DAsync(dR, ldr * sizeof(double), buf[cur], n * sizeof(double), n * sizeof(double), n,
                                      cudaMemcpyDeviceToDevice, st),
                    __FILE__, __LINE__);
  cudaFreeAsync(buf[0], st);
  if (buf[1]) cudaFreeAsync(buf[1], st);
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
  // level-0 blocks of 64 rows so that the chunk size (rows_target) adapts to n
  const int rows_blk = 64;
  const i64 total_blocks = (m + rows_blk - 1) / rows_blk;
  const i64 last_rows = m - (total_blocks - 1) * rows_blk;
  // the ping-pong sizing in run_tree assumes a fan-in >= 4 at the upper levels
  return run_tree(dA, lda, rows_blk, total_blocks, rows_blk, last_rows, (int)n, dR, ldr, st);
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
  return run_tree(dRs, n, n * n, count, (int)n, n, (int)n, dR, ldr, st);
}

}  // namespace gla
