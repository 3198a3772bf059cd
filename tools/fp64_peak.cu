// Microbenchmark: peak DFMA and DMMA (mma.sync.m8n8k4.f64) rates on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = a[i] * b + c;
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0); k_dfma<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("DFMA  warps/SM %2d: %.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", warps, 2.0 * sms * warps * 32 * 16.0 * iters / ms / 1e9,
                      sms * warps * 32 * 16.0 * iters / (ms * 1e-3) / sms / 1.965e9);
      cudaEventRecord(e0); k_dmma<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("DMMA  warps/SM %2d: %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, 2.0 * sms * warps * 8.0 * 256 * iters / ms / 1e9,
                      sms * warps * 8.0 * 256 * iters / (ms * 1e-3) / sms / 1.965e9);
    }
  }
  return 0;
}
