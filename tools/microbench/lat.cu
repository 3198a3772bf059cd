// Single-warp dependent-chain latencies on sm_100a (cycles per operation), the numbers the BCR kernels are
// designed against.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

constexpr int N = 512;

__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = double((i * 7 + 1) & 1023);
  __syncthreads();
  long long t0, t1;
  double x = seed + lane * 1e-3, y = 1.0 + 1e-9 * lane;
  int slot = 0;
  auto rec = [&](long long d, int n) { if (threadIdx.x == 0) cyc[slot] = d * 100 / n; ++slot; };
  // 0: DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); rec(t1 - t0, N);
  // 1: DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + y;
  t1 = clock64(); rec(t1 - t0, N);
  // 2: DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); rec(t1 - t0, N);
  // 3: dependent DMMA chain (accumulator)
  double c[2] = {x, y};
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma(c, 1e-3, 1e-3);
  t1 = clock64(); rec(t1 - t0, N);
  // 4: DMMA chain through the A operand (result -> next A)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double d[2] = {0.0, 0.0}; dmma(d, c[0], 1e-3); c[0] = d[0]; c[1] = d[1]; }
  t1 = clock64(); rec(t1 - t0, N);
  // 5: 4 independent DMMA chains (issue throughput of one warp), per DMMA
  double c1[2] = {x, y}, c2[2] = {y, x}, c3[2] = {x, x}, c4[2] = {y, y};
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { dmma(c1, 1e-3, 1e-3); dmma(c2, 1e-3, 1e-3); dmma(c3, 1e-3, 1e-3); dmma(c4, 1e-3, 1e-3); }
  t1 = clock64(); rec(t1 - t0, 4 * N);
  // 6: SHFL.64 chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64(); rec(t1 - t0, N);
  // 7: rcp.approx + 2 Newton steps chain
  x = 1.5 + 1e-3 * lane;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = rcp_nr(x) + 1.0;
  t1 = clock64(); rec(t1 - t0, N);
  // 8: rsqrt(double) chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64(); rec(t1 - t0, N);
  // 9: 1.0 / x chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
  t1 = clock64(); rec(t1 - t0, N);
  // 10: sqrt chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
  t1 = clock64(); rec(t1 - t0, N);
  // 11: dependent LDS chain
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = int(sm[idx]);
  t1 = clock64(); rec(t1 - t0, N);
  // 12: STS -> syncwarp -> LDS round trip
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); rec(t1 - t0, N);
  // 13: __syncthreads (whole block)
  __syncthreads();
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 128; ++i) __syncthreads();
  t1 = clock64(); rec(t1 - t0, 128);
  // 14: DMMA -> DADD -> DMMA chain (split accumulators as in the sweeps)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    double u[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
    dmma(u, c[0], 1e-3); dmma(v, c[1], 1e-3);
    c[0] = u[0] + v[0]; c[1] = u[1] + v[1];
  }
  t1 = clock64(); rec(t1 - t0, N);
  out[threadIdx.x] = x + c[0] + c[1] + c1[0] + c2[0] + c3[0] + c4[0] + idx;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, 32 * sizeof(long long));
  const char* names[] = {"DFMA chain", "DADD chain", "DMUL chain", "DMMA chain (accumulator)", "DMMA chain (A operand)",
                         "DMMA x4 independent (per DMMA)", "SHFL.64 chain", "rcp.approx + 2 NR (+ DADD)", "rsqrt(double) (+ DADD)",
                         "1.0 / x (+ DADD)", "sqrt (+ DADD)", "LDS chain (+ F2I)", "STS-syncwarp-LDS-syncwarp (+ DADD)",
                         "__syncthreads", "2 DMMA -> DADD chain"};
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, cyc, 1.0);
    cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 1.0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("---- %d threads per CTA (%s) ----\n", threads, cudaGetErrorString(e));
    for (int i = 0; i < 15; ++i) printf("%-40s %8.2f cycles\n", names[i], h[i] / 100.0);
  }
  return 0;
}
