// fp64 dependent-issue latencies on one warp (clock64 around a dependent chain).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* cyc, double x0, double a, double b) {
  double x = x0 + threadIdx.x * 1e-9;
  const int N = 4096;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, a, b);
    if (OP == 1) x = rsqrt(x) + b;
    if (OP == 2) x = sqrt(x) + b;
    if (OP == 3) x = 1.0 / x + b;
    if (OP == 4) x = x * a;
    if (OP == 5) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;
    if (OP == 6) x = __drcp_rn(x) + b;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = (t1 - t0);
}
__global__ void k_lds(double* out, long long* cyc, int stride) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + stride) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; ++i) j = idx[j];
  long long t1 = clock64();
  out[threadIdx.x] = j;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sync(long long* cyc) {
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "rsqrt+add", "sqrt+add", "div+add", "dmul", "shfl64+add", "drcp+add"};
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.5, 0.999, 0.7); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("%-12s %.1f cycles/op\n", names[OP], h / 4096.0);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    k_lds<<<1, 32>>>(out, cyc, 33); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("%-12s %.1f cycles/op\n", "lds chain", h / 4096.0);
    k_sync<<<1, 256>>>(cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("%-12s %.1f cycles/op\n", "bar.sync 256", h / 1024.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
