// Cycle-level timing + correctness of the CTA-level dense routines of bcr_dense.cuh on one CTA
// (M = 88, NB = 8: the config-4 super block).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   -Iphotometric-bundle-adjustment_b200/csrc -Iinclude tools/bcr_dense_bench.cu -o tools/bcr_dense_bench
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "bcr_dense.cuh"
using namespace pba;

template <int NB>
__global__ void __launch_bounds__(kBcrThreads) k_bench(int M, const double* A, const double* B, double* Lout, double* Uout,
                                                        double* wout, long long* t, int* fail) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;
  double* Ws = sm + M * ld;
  double* Dinv = Ws + M * ld;
  double* Pt = Dinv + M * NB;
  double* w = Pt + NB * bcr_ldp(M);
  cta_load(Ls, ld, A, M, false);
  cta_load(Ws, ld, B, M, false);
  for (int i = threadIdx.x; i < M; i += kBcrThreads) { Ws[i * ld + M] = B[i]; }
  cta_load_wait();
  __syncthreads();
  long long t0 = clock64();
  cta_cholesky<NB>(Ls, M, ld, Dinv, Pt, fail);
  __syncthreads();
  long long t1 = clock64();
  cta_trsm_lower<NB>(Ls, ld, Dinv, Ws, ld, M, M + 1);
  __syncthreads();
  long long t2 = clock64();
  for (int i = threadIdx.x; i < M; i += kBcrThreads) w[i] = Ws[i * ld + M];
  __syncthreads();
  long long t3 = clock64();
  cta_solve_lt<NB>(Ls, ld, Dinv, w, M);
  __syncthreads();
  long long t4 = clock64();
  cta_store(Lout, Ls, ld, M, true);
  cta_store(Uout, Ws, ld, M, false);
  for (int i = threadIdx.x; i < M; i += kBcrThreads) wout[i] = w[i];
  if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t4 - t3; }
}

int main() {
  const int M = 88, NB = 8;
  std::vector<double> G(M * M), A(M * M), B(M * M);
  srand(1);
  for (auto& v : G) v = rand() / double(RAND_MAX) - 0.5;
  for (auto& v : B) v = rand() / double(RAND_MAX) - 0.5;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      double s = i == j ? 1.0 : 0.0;
      for (int k = 0; k < M; ++k) s += G[i * M + k] * G[j * M + k];
      A[i * M + j] = s;
    }
  // host reference
  std::vector<double> L(A), U(B), y(M), x(M);
  for (int j = 0; j < M; ++j) {
    for (int k = 0; k < j; ++k) for (int i = j; i < M; ++i) L[i * M + j] -= L[i * M + k] * L[j * M + k];
    const double d = std::sqrt(L[j * M + j]);
    for (int i = j; i < M; ++i) L[i * M + j] /= d;
  }
  for (int c = 0; c < M; ++c)
    for (int i = 0; i < M; ++i) {
      double s = U[i * M + c];
      for (int k = 0; k < i; ++k) s -= L[i * M + k] * U[k * M + c];
      U[i * M + c] = s / L[i * M + i];
    }
  for (int i = 0; i < M; ++i) { double s = B[i]; for (int k = 0; k < i; ++k) s -= L[i * M + k] * y[k]; y[i] = s / L[i * M + i]; }
  for (int i = M - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < M; ++k) s -= L[k * M + i] * x[k]; x[i] = s / L[i * M + i]; }

  double *dA, *dB, *dL, *dU, *dw; long long* dt; int* df;
  cudaMalloc(&dA, 8 * M * M); cudaMalloc(&dB, 8 * M * M); cudaMalloc(&dL, 8 * M * M); cudaMalloc(&dU, 8 * M * M);
  cudaMalloc(&dw, 8 * M); cudaMalloc(&dt, 64); cudaMalloc(&df, 4);
  cudaMemset(df, 0, 4);
  cudaMemcpy(dA, A.data(), 8 * M * M, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), 8 * M * M, cudaMemcpyHostToDevice);
  const int smem = (2 * M * bcr_ld_odd(M) + M * NB + NB * bcr_ldp(M) + M + 8) * 8;
  cudaFuncSetAttribute(k_bench<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long t[3];
  for (int rep = 0; rep < 3; ++rep) {
    k_bench<NB><<<1, kBcrThreads, smem>>>(M, dA, dB, dL, dU, dw, dt, df);
    cudaMemcpy(t, dt, 24, cudaMemcpyDeviceToHost);
  }
  std::vector<double> hL(M * M), hU(M * M), hw(M);
  cudaMemcpy(hL.data(), dL, 8 * M * M, cudaMemcpyDeviceToHost);
  cudaMemcpy(hU.data(), dU, 8 * M * M, cudaMemcpyDeviceToHost);
  cudaMemcpy(hw.data(), dw, 8 * M, cudaMemcpyDeviceToHost);
  double eL = 0, eU = 0, ex = 0;
  for (int i = 0; i < M; ++i) for (int j = 0; j <= i; ++j) eL = fmax(eL, fabs(hL[i * M + j] - L[i * M + j]));
  for (int i = 0; i < M * M; ++i) eU = fmax(eU, fabs(hU[i] - U[i]));
  for (int i = 0; i < M; ++i) ex = fmax(ex, fabs(hw[i] - x[i]));
  int fail; cudaMemcpy(&fail, df, 4, cudaMemcpyDeviceToHost);
#ifdef BCR_DENSE_PROF
  {
    long long p[16];
    cudaMemcpyFromSymbol(p, g_prof, sizeof(p));
    printf("per call (thread 0), cycles: chol diag %lld wait %lld panel %lld wait %lld trailing %lld wait %lld | trsm diag %lld wait %lld update %lld wait %lld\n",
           p[0] / 3, p[1] / 3, p[2] / 3, p[3] / 3, p[4] / 3, p[5] / 3, p[8] / 3, p[9] / 3, p[10] / 3, p[11] / 3);
  }
#endif
  printf("cholesky %lld cyc (%.2f us)  trsm(89 cols) %lld cyc (%.2f us)  solve_lt %lld cyc (%.2f us) | err L %.2e U %.2e x %.2e fail %d  %s\n",
         t[0], t[0] / 1965.0, t[1], t[1] / 1965.0, t[2], t[2] / 1965.0, eL, eU, ex, fail, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
