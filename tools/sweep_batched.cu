// sweep_batched.cu -- torch-free timing harness for the batched 32x32 Float64 QR entry point (one process = one
// GLA_BATCHED_VARIANT).  Times gla_dgeqr_batched_dev with CUDA events on fresh input every repetition and dumps the
// first / last NDUMP matrices (input, factors, tau) so that tools/sweep_check.py can compare variants with the oracle.
//   make -C tools     (or: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/sweep_batched.cu -o tools/sweep_batched \
//        -Iinclude -Lgenericlinearalgebra.jl_b200/lib -lgla_cuda)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "gla_cuda.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void fill(double* a, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // sum of four uniforms: roughly normal, every matrix different
    const uint64_t h = mix(i), g = mix(i ^ 0xABCDEF1234567ull);
    const double u = ((h & 0xffffffffu) + (h >> 32) + (g & 0xffffffffu) + (g >> 32)) * (1.0 / 4294967296.0) - 2.0;
    a[i] = u * 1.7320508;
  }
}

int main(int argc, char** argv) {
  const int64_t batch = argc > 1 ? atoll(argv[1]) : (1 << 20);
  const int reps = argc > 2 ? atoi(argv[2]) : 7;
  const char* dump = argc > 3 ? argv[3] : nullptr;
  const int64_t NDUMP = std::min<int64_t>(1024, batch);
  const size_t n = (size_t)batch * 1024;
  double *src, *A, *tau;
  CK(cudaMalloc(&src, n * 8));
  CK(cudaMalloc(&A, n * 8));
  CK(cudaMalloc(&tau, (size_t)batch * 32 * 8));
  fill<<<148 * 8, 256>>>(src, n);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemcpyAsync(A, src, n * 8, cudaMemcpyDeviceToDevice, 0));
    CK(cudaMemsetAsync(tau, 0xff, (size_t)batch * 32 * 8, 0));
    CK(cudaEventRecord(e0, 0));
    const int rc = gla_dgeqr_batched_dev(A, 32, 32, batch, tau, nullptr);
    CK(cudaEventRecord(e1, 0));
    if (rc) { fprintf(stderr, "rc=%d %s\n", rc, gla_last_error_string()); return 3; }
    CK(cudaDeviceSynchronize());
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    if (r >= 2) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  const double best = ms[0], med = ms[ms.size() / 2];
  const char* v = getenv("GLA_BATCHED_VARIANT");
  printf("variant %s batch %lld: best %.3f ms median %.3f ms -> %.1f M matrices/s, %.0f GB/s algorithmic\n", v ? v : "default",
         (long long)batch, best, med, batch / best / 1e3, batch * 16640.0 / best / 1e6);
  if (dump) {
    FILE* f = fopen(dump, "wb");
    if (!f) return 4;
    std::vector<double> h((size_t)NDUMP * 1024), ht((size_t)NDUMP * 32);
    for (int part = 0; part < 2; ++part) {
      const int64_t m0 = part == 0 ? 0 : batch - NDUMP;
      CK(cudaMemcpy(h.data(), src + m0 * 1024, h.size() * 8, cudaMemcpyDeviceToHost));
      fwrite(h.data(), 8, h.size(), f);
      CK(cudaMemcpy(h.data(), A + m0 * 1024, h.size() * 8, cudaMemcpyDeviceToHost));
      fwrite(h.data(), 8, h.size(), f);
      CK(cudaMemcpy(ht.data(), tau + m0 * 32, ht.size() * 8, cudaMemcpyDeviceToHost));
      fwrite(ht.data(), 8, ht.size(), f);
    }
    fclose(f);
  }
  return 0;
}
