// fp64_peak.cu -- measures the FP64 roofline denominators on the box: DFMA (vector pipe) and
// DMMA m8n8k4 (mma.sync f64) peak, alone and mixed.  MEASURED_PEAKS.json has no FP64 entry.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8];
  double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i * 0.1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: dmma only, 1: mixed (odd warps dfma)
__global__ void dmma_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  int warp = threadIdx.x >> 5;
  if (MODE == 1 && (warp & 1)) {
    double x = 1.0000001, y = 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = fma(c[i], x, y);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      int grid = sms * (1024 / threads) ;
      cudaEventRecord(e0);
      dfma_kernel<<<grid, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 8 * iters * (double)grid * threads;
      if (rep) printf("DFMA  threads/SM=1024 (cta %4d): %.2f TFLOP/s (%.3f ms)\n", threads, fl / ms * 1e-9, ms);
    }
  }
  for (int wps : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      int threads = wps * 32;
      cudaEventRecord(e0);
      dmma_kernel<0><<<sms, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 256 * 8 * iters * (double)sms * wps;
      if (rep) printf("DMMA.884 warps/SM=%2d: %.2f TFLOP/s (%.3f ms)\n", wps, fl / ms * 1e-9, ms);
    }
  }
  for (int wps : {8, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      int threads = wps * 32;
      cudaEventRecord(e0);
      dmma_kernel<1><<<sms, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fl_mma = 2.0 * 256 * 8 * iters * (double)sms * (wps / 2);
      double fl_fma = 2.0 * 32 * 8 * iters * (double)sms * (wps / 2);
      if (rep) printf("MIXED warps/SM=%2d: DMMA %.2f + DFMA %.2f TFLOP/s (%.3f ms)\n", wps, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9, ms);
    }
  }
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
