// fp64_lat.cu -- dependent-issue latency of DFMA / DMUL->DFMA / MUFU.RSQ64H / LDS.128 broadcast and how
// dependent chains from several warps of one SMSP overlap.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double* out, long long* cyc, int iters, int ilp) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double x = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  if (ilp == 1) for (int i = 0; i < iters; ++i) { a0 = fma(a0, x, y); }
  else if (ilp == 2) for (int i = 0; i < iters; ++i) { a0 = fma(a0, x, y); a1 = fma(a1, x, y); }
  else for (int i = 0; i < iters; ++i) { a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y); }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void mufu(double* out, long long* cyc, int iters) {
  double a = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y + 1.0; }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ldsb(double* out, long long* cyc, int iters) {
  __shared__ double2 s[64];
  s[threadIdx.x & 63] = make_double2(0, 0);
  __syncthreads();
  int idx = 0; double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { double vx, vy; asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"((unsigned)__cvta_generic_to_shared(&s[idx]))); idx = (int)vx; acc += vy; }
  long long t1 = clock64();
  out[threadIdx.x] = acc + idx;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  long long h; const int iters = 4096;
  for (int ilp : {1, 2, 4}) for (int warps : {1, 4, 8, 12, 16, 32}) {
    chain<<<1, warps * 32>>>(out, cyc, iters, ilp); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA chain ilp=%d warps/SM=%2d (%.1f per SMSP): %.2f cycles per dependent step, %.2f cyc/warp-instr/SMSP\n", ilp, warps, warps / 4.0, (double)h / iters, (double)h / iters / ilp / (warps > 4 ? warps / 4.0 : 1));
  }
  mufu<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RSQ64H + DADD dependent: %.2f cycles\n", (double)h / iters);
  ldsb<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.128 broadcast dependent (incl. cvt+add): %.2f cycles\n", (double)h / iters);
  return 0;
}
