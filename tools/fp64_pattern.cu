// fp64_pattern.cu -- DFMA throughput for the operand patterns of the reflector sweeps (3 distinct register-pair
// operands per instruction), against the reuse-friendly pattern tools/fp64_peak.cu measures.
//   dot : acc[j] = fma(y[i], a[i], acc[j])      (two fresh operands + one of NACC accumulators)
//   axpy: a[i]   = fma(nt,   y[i], a[i])        (one operand reused across the sweep)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NE = 24;

template <int MODE, int NACC>
__global__ void __launch_bounds__(512) k(double* out, const double* in, int iters) {
  double y[NE], a[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) { y[i] = in[i] + threadIdx.x * 1e-9; a[i] = in[32 + i]; }
  double nt = in[64];
  double tot = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      double acc[NACC];
#pragma unroll
      for (int j = 0; j < NACC; ++j) acc[j] = 0;
#pragma unroll
      for (int i = 0; i < NE; ++i) acc[i % NACC] = fma(y[i], a[i], acc[i % NACC]);
      double d = 0;
#pragma unroll
      for (int j = 0; j < NACC; ++j) d += acc[j];
      tot += d;
      if (MODE == 2) nt = d * 1e-30;
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < NE; ++i) a[i] = fma(nt, y[i], a[i]);
    }
    if (MODE == 0) { nt += 1e-30; }
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) tot += a[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <int MODE, int NACC>
void run(const char* name, double* out, double* in, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  for (int wps : {4, 8, 12, 16}) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k<MODE, NACC><<<sms, wps * 32>>>(out, in, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double nf = (MODE == 2 ? 2.0 * NE : 1.0 * NE) * iters * wps;  // DFMA warp instructions per SM
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-22s warps/SM=%2d: %.2f TFLOP/s, %.2f cycles per DFMA per SMSP at %d MHz nominal\n", name, wps,
           2.0 * 32 * nf * sms / best * 1e-9, best * 1e-3 * clk * 1e3 / (nf / 4), clk / 1000);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  cudaMalloc(&in, sizeof(double) * 128);
  double h[128];
  for (int i = 0; i < 128; ++i) h[i] = 1.0 / (i + 3);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0, 2>("dot  NACC=2", out, in, sms);
  run<0, 4>("dot  NACC=4", out, in, sms);
  run<1, 1>("axpy", out, in, sms);
  run<2, 2>("dot+axpy NACC=2", out, in, sms);
  run<2, 4>("dot+axpy NACC=4", out, in, sms);
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
