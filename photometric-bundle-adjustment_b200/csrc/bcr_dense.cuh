// bcr_dense.cuh — CTA-level dense fp64 routines on shared-memory matrices used by the block
// cyclic reduction solver (bcr.cu): blocked Cholesky, triangular solves, 4x4-tile products,
// cp.async loads.  Included by bcr.cu and by tools/bcr_dense_bench.cu (cycle-level timing).
#pragma once
#include <stdint.h>

namespace pba {
namespace {

constexpr int kBcrThreads = 256;

// Debug only (tools/bcr_dense_bench.cu -DBCR_DENSE_PROF): per-phase cycles seen by thread 0.
#ifdef BCR_DENSE_PROF
__device__ long long g_prof[16];
#define PROF_BEGIN() long long _pt = clock64()
#define PROF(i) do { if (threadIdx.x == 0) { const long long _n = clock64(); g_prof[i] += _n - _pt; _pt = _n; } } while (0)
#else
#define PROF_BEGIN()
#define PROF(i)
#endif

// Debug only (make EXTRA=-DPBA_BCR_TIMING): clock64 stamps of CTA (0,0) of each kernel's first launch.
#ifdef PBA_BCR_TIMING
__device__ long long g_bcr_t[32];
#define BCR_STAMP(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_bcr_t[i] = clock64(); } while (0)
#else
#define BCR_STAMP(i)
#endif

// shared-memory leading dimensions.  The product kernel (k_bcr_reduce) wants 16-byte aligned
// rows (even) covering the 4-wide tiles; the factorisation kernels walk columns with one thread
// per row, which is conflict-free only for an odd stride.
__host__ __device__ inline int bcr_ld(int M) { return ((M + 3) / 4) * 4 + 2; }
__host__ __device__ inline int bcr_ld_odd(int M) { return M + 1 + (M & 1); }

// ---- CTA-level dense kernels on shared-memory matrices, blocked by NB = cd ----
// All matrices are row-major with leading dimension ld.  The only serial piece is the NB x NB
// diagonal-block factorisation (one thread, all in registers); the trailing matrix lives in
// registers (one or two TS x TS tiles per thread, TS = NB / 2) for the whole factorisation, so
// shared memory sees every entry of the factor exactly once.

// Factor the NB x NB block at D (lower Cholesky, in place) and write the inverse of the factor to
// Di [NB*NB] (lower, zeros above).  Executed by ONE thread, everything in registers.
template <int NB>
__device__ __forceinline__ void thread_factor_diag(double* D, int ld, double* Di, int* fail) {
  double L[NB][NB], X[NB][NB], id[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) L[r][c] = D[r * ld + c];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double d = L[j][j];
    if (!(d > 0.0)) { *fail = 1; d = 1.0; }
    const double inv = rsqrt(d);
    L[j][j] = d * inv;
    id[j] = inv;
#pragma unroll
    for (int r = j + 1; r < NB; ++r) L[r][j] *= inv;
#pragma unroll
    for (int c = j + 1; c < NB; ++c)
#pragma unroll
      for (int r = c; r < NB; ++r) L[r][c] -= L[r][j] * L[c][j];
  }
  // X = L^-1, column by column (the NB columns are independent chains)
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    X[c][c] = id[c];
#pragma unroll
    for (int r = c + 1; r < NB; ++r) {
      double t = 0.0;
#pragma unroll
      for (int q = c; q < r; ++q) t += L[r][q] * X[q][c];
      X[r][c] = -t * id[r];
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r)
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (c <= r) D[r * ld + c] = L[r][c];
      Di[r * NB + c] = c <= r ? X[r][c] : 0.0;
    }
}

// Shared scratch of cta_cholesky: the current panel, transposed (Pt[q][i - n0], q < NB).
__host__ __device__ inline int bcr_ldp(int M) { return ((M + 3) / 4) * 4 + 4; }
// largest M / NB cta_cholesky's static tile ownership covers
__host__ __device__ inline int bcr_max_blocks(int NB) { return NB == 8 ? 15 : 19; }

// A <- lower Cholesky factor of A (strict upper triangle untouched); Dinv[J] = inverse
// of the J-th diagonal block of the factor.  Pt: NB * bcr_ldp(M) doubles of scratch.
template <int NB>
__device__ void cta_cholesky(double* __restrict__ A, int M, int ld, double* __restrict__ Dinv, double* __restrict__ Pt,
                             int* fail) {
  constexpr int TS = NB / 2;  // tile size: tiles never straddle a block column
  // tiles per thread: T (T + 1) / 2 <= kTiles * 256 with T = 2 M / NB.  The shared-memory cap of
  // bcr_super_size() bounds M / NB by 14 (NB = 8) and 19 (NB = 6); see kBcrMaxBlocks.
  constexpr int kTiles = NB == 8 ? 2 : 3;
  const int tid = threadIdx.x;
  const int nbk = M / NB;
  const int T = 2 * nbk;
  const int ntile = T * (T + 1) / 2;
  const int ldp = bcr_ldp(M);
  // static tile ownership: lower-triangle tiles u = tr (tr + 1) / 2 + tc, tc <= tr
  int trs[kTiles], tcs[kTiles];
  double acc[kTiles][TS][TS];
#pragma unroll
  for (int k = 0; k < kTiles; ++k) {
    const int u = tid + k * kBcrThreads;
    int tr = -1, tc = 0;
    if (u < ntile) {
      tr = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
      while (tr * (tr + 1) / 2 > u) --tr;
      while ((tr + 1) * (tr + 2) / 2 <= u) ++tr;
      tc = u - tr * (tr + 1) / 2;
    }
    trs[k] = tr; tcs[k] = tc;
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
      for (int b = 0; b < TS; ++b) acc[k][a][b] = tr >= 0 ? A[(TS * tr + a) * ld + TS * tc + b] : 0.0;
  }
  PROF_BEGIN();
  for (int J = 0; J < nbk; ++J) {
    const int j0 = J * NB, n0 = j0 + NB;
    double* Di = Dinv + J * NB * NB;
    // block column J leaves the registers (J = 0: A still holds it)
    if (J > 0) {
#pragma unroll
      for (int k = 0; k < kTiles; ++k)
        if (trs[k] >= 0 && (tcs[k] >> 1) == J) {
#pragma unroll
          for (int a = 0; a < TS; ++a)
#pragma unroll
            for (int b = 0; b < TS; ++b) A[(TS * trs[k] + a) * ld + TS * tcs[k] + b] = acc[k][a][b];
        }
      __syncthreads();
    }
    if (tid == 0) thread_factor_diag<NB>(A + j0 * ld + j0, ld, Di, fail);
    PROF(0);
    __syncthreads();
    PROF(1);
    // panel: rows below the diagonal block  <-  row * L_D^-T = row * Di^T; kept in A (final)
    // and, transposed, in Pt for the trailing update
    for (int i = n0 + tid; i < M; i += kBcrThreads) {
      double v[NB], o[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) v[q] = A[i * ld + j0 + q];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q <= c; ++q) s += v[q] * Di[c * NB + q];
        o[c] = s;
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        A[i * ld + j0 + q] = o[q];
        Pt[q * ldp + (i - n0)] = o[q];
      }
    }
    PROF(2);
    __syncthreads();
    PROF(3);
    // trailing tiles (block columns > J) -= panel panel^T, in registers
#pragma unroll
    for (int k = 0; k < kTiles; ++k) {
      if (trs[k] < 0 || (tcs[k] >> 1) <= J) continue;
      const double* xp = Pt + (TS * trs[k] - n0);
      const double* yp = Pt + (TS * tcs[k] - n0);
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double x[TS], y[TS];
        if (TS == 4) {
          const double2 x0 = *reinterpret_cast<const double2*>(xp + q * ldp);
          const double2 x1 = *reinterpret_cast<const double2*>(xp + q * ldp + 2);
          const double2 y0 = *reinterpret_cast<const double2*>(yp + q * ldp);
          const double2 y1 = *reinterpret_cast<const double2*>(yp + q * ldp + 2);
          x[0] = x0.x; x[1] = x0.y; x[TS - 2] = x1.x; x[TS - 1] = x1.y;
          y[0] = y0.x; y[1] = y0.y; y[TS - 2] = y1.x; y[TS - 1] = y1.y;
        } else {
#pragma unroll
          for (int a = 0; a < TS; ++a) { x[a] = xp[q * ldp + a]; y[a] = yp[q * ldp + a]; }
        }
#pragma unroll
        for (int a = 0; a < TS; ++a)
#pragma unroll
          for (int b = 0; b < TS; ++b) acc[k][a][b] -= x[a] * y[b];
      }
    }
    PROF(4);
    // no barrier needed here: the next step's first barrier orders Pt reads before its rewrite
    PROF(5);
  }
  __syncthreads();
}

// W (M x ncols, shared) <- L^-1 W, blocked forward substitution.
template <int NB>
__device__ void cta_trsm_lower(const double* __restrict__ L, int ld, const double* __restrict__ Dinv,
                               double* __restrict__ W, int ldw, int M, int ncols) {
  const int tid = threadIdx.x;
  const int nbk = M / NB;
  PROF_BEGIN();
  for (int J = 0; J < nbk; ++J) {
    const int j0 = J * NB;
    const double* Di = Dinv + J * NB * NB;
    for (int c = tid; c < ncols; c += kBcrThreads) {
      double v[NB], o[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) v[q] = W[(j0 + q) * ldw + c];
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q <= r; ++q) s += Di[r * NB + q] * v[q];
        o[r] = s;
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) W[(j0 + q) * ldw + c] = o[q];
    }
    PROF(8);
    __syncthreads();
    PROF(9);
    // rows below: W[i][c] -= L[i][j0..] . W[j0..][c].  Thread (c = tid % 128, g = tid / 128) owns
    // column c and every other 4-row strip; the pivot-row values stay in registers and the four
    // rows of a strip are loaded, accumulated and stored as independent chains.
    const int n0 = j0 + NB;
    const int c = tid & 127, g = tid >> 7;
    for (int cc = c; cc < ncols; cc += 128) {
      double w[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) w[q] = W[(j0 + q) * ldw + cc];
      for (int i0 = n0 + 4 * g; i0 < M; i0 += 8) {
        double l[4][NB], s[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = min(i0 + a, M - 1);
#pragma unroll
          for (int q = 0; q < NB; ++q) l[a][q] = L[i * ld + j0 + q];
          s[a] = W[i * ldw + cc];
        }
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int a = 0; a < 4; ++a) s[a] -= l[a][q] * w[q];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (i0 + a < M) W[(i0 + a) * ldw + cc] = s[a];
      }
    }
    PROF(10);
    __syncthreads();
    PROF(11);
  }
}

// w (M, shared) <- L^-T w, blocked backward substitution.
template <int NB>
__device__ void cta_solve_lt(const double* L, int ld, const double* Dinv, double* w, int M) {
  const int tid = threadIdx.x;
  const int nbk = M / NB;
  for (int J = nbk - 1; J >= 0; --J) {
    const int j0 = J * NB;
    const double* Di = Dinv + J * NB * NB;
    if (tid < 32) {
      double v = 0.0;
      if (tid < NB) {
#pragma unroll
        for (int q = 0; q < NB; ++q) v += Di[q * NB + tid] * w[j0 + q];  // Di^T w_J
      }
      __syncwarp();
      if (tid < NB) w[j0 + tid] = v;
    }
    __syncthreads();
    for (int k = tid; k < j0; k += kBcrThreads) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < NB; ++q) s += L[(j0 + q) * ld + k] * w[j0 + q];
      w[k] -= s;
    }
    __syncthreads();
  }
}

// One 4x4 tile of X^T Y over k in [k0, k1):  acc[a][b] += sum_k X[k][4 tr + a] * Y[k][4 tc + b].
// X, Y: M x M in shared memory with an even leading dimension (16-byte LDS).  M is a multiple of
// cd (6 or 8); pad columns beyond M read finite junk only when M % 4 != 0 and are never stored.
__device__ __forceinline__ void tile_xty(const double* X, const double* Y, int ld, int tr, int tc, int k0, int k1,
                                         double (&acc)[4][4]) {
  const double* xp = X + 4 * tr;
  const double* yp = Y + 4 * tc;
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    const double2 x0 = *reinterpret_cast<const double2*>(xp + k * ld);
    const double2 x1 = *reinterpret_cast<const double2*>(xp + k * ld + 2);
    const double2 y0 = *reinterpret_cast<const double2*>(yp + k * ld);
    const double2 y1 = *reinterpret_cast<const double2*>(yp + k * ld + 2);
    const double xv[4] = {x0.x, x0.y, x1.x, x1.y};
    const double yv[4] = {y0.x, y0.y, y1.x, y1.y};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += xv[a] * yv[b];
  }
}

// global (M x M, dense) -> shared (leading dimension ld), optionally transposed, with cp.async
// (LDGSTS): every copy is in flight at once; call cta_load_wait() before reading.
__device__ __forceinline__ void cta_load(double* dst, int ld, const double* __restrict__ src, int M, bool transpose,
                                         bool lower_only = false) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned d0 = unsigned(__cvta_generic_to_shared(dst));
  for (int r = warp; r < M; r += kBcrThreads / 32) {
    const double* s = src + int64_t(r) * M;
    const int cend = lower_only ? r + 1 : M;  // a Cholesky factor: the strict upper triangle is never read
    for (int c = lane; c < cend; c += 32) {
      const unsigned da = d0 + unsigned((transpose ? c * ld + r : r * ld + c) * 8);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(da), "l"(s + c));
    }
  }
}
__device__ __forceinline__ void cta_load_wait() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::);
  __syncthreads();
}
__device__ __forceinline__ void cta_store(double* __restrict__ dst, const double* src, int ld, int M, bool lower_only) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < M; r += kBcrThreads / 32) {
    double* d = dst + int64_t(r) * M;
#pragma unroll 4
    for (int c = lane; c < M; c += 32) d[c] = (!lower_only || c <= r) ? src[r * ld + c] : 0.0;
  }
}

}  // namespace
}  // namespace pba
