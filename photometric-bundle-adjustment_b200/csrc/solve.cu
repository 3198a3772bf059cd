// solve.cu — solving the reduced camera system.
//
// Replaces Ceres' SparseSchurComplementSolver::SolveReducedLinearSystem
// (internal/ceres/schur_complement_solver.cc:319-356; sparse LDL^T on the CPU,
// internal/ceres/eigensparse.cc:56-106) with
//   * a tiled dense fp64 Cholesky for small RCS — the one true dense
//     contraction on the path, so its trailing update runs on the FP64 tensor
//     cores (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no fp64 kind);
//   * a block-Jacobi preconditioned conjugate-gradient solver on the
//     block-sparse RCS for large problems (the analogue of Ceres' SCHUR_JACOBI
//     PCG, conjugate_gradients_solver.cc:63-249), run as ONE cooperative
//     kernel so a solve costs one launch, not five per iteration.
#include <cooperative_groups.h>
#include <stdio.h>

#include <mutex>
#include <utility>
#include <vector>

#include "launch.h"
#include "pba_internal.h"

namespace cg = cooperative_groups;

namespace pba {

namespace {

constexpr int NB = 64;   // Cholesky tile
constexpr int LDS = 68;  // padded shared-memory row stride (doubles)

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// dense <- block list (both triangles), padded to a multiple of NB with identity.
__global__ void k_dense_init(int n, int ld, double* __restrict__ A, double* __restrict__ b_pad) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(ld) * ld) return;
  const int r = int(i / ld), c = int(i % ld);
  A[i] = (r == c && r >= n) ? 1.0 : 0.0;
  if (c == 0 && r >= n) b_pad[r] = 0.0;
}

__global__ void k_dense_fill(int cd, int64_t n_blocks, int ld, const int* __restrict__ blk_row,
                             const int* __restrict__ blk_col, const double* __restrict__ S, double* __restrict__ A) {
  const int64_t b = blockIdx.x;
  const int e = threadIdx.x;
  if (b >= n_blocks || e >= cd * cd) return;
  const int r = e / cd, c = e % cd;
  const int gr = blk_row[b] * cd + r, gc = blk_col[b] * cd + c;
  const double v = S[b * cd * cd + e];
  A[int64_t(gr) * ld + gc] = v;
  A[int64_t(gc) * ld + gr] = v;
}

// Factor the diagonal tile A_kk = L L^T in shared memory (one CTA).
__global__ void __launch_bounds__(256) k_chol_diag(int ld, int k, double* __restrict__ A, int* __restrict__ fail) {
  __shared__ double T[NB * (NB + 1)];
  double* Akk = A + (int64_t(k) * NB) * ld + k * NB;
  for (int i = threadIdx.x; i < NB * NB; i += 256) T[(i / NB) * (NB + 1) + i % NB] = Akk[int64_t(i / NB) * ld + i % NB];
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    if (threadIdx.x == 0) {
      const double d = T[j * (NB + 1) + j];
      if (!(d > 0.0)) { *fail = 1; T[j * (NB + 1) + j] = 1.0; }
      else T[j * (NB + 1) + j] = sqrt(d);
    }
    __syncthreads();
    const double dj = T[j * (NB + 1) + j];
    for (int i = j + 1 + threadIdx.x; i < NB; i += 256) T[i * (NB + 1) + j] /= dj;
    __syncthreads();
    // rank-1 update of the trailing lower triangle
    const int m = NB - j - 1;
    for (int idx = threadIdx.x; idx < m * m; idx += 256) {
      const int r = j + 1 + idx / m, c = j + 1 + idx % m;
      if (c <= r) T[r * (NB + 1) + c] -= T[r * (NB + 1) + j] * T[c * (NB + 1) + j];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NB * NB; i += 256) {
    const int r = i / NB, c = i % NB;
    Akk[int64_t(r) * ld + c] = c <= r ? T[r * (NB + 1) + c] : 0.0;
  }
}

// A_ik <- A_ik L_kk^-T for all row tiles i > k (one CTA of 64 threads per tile;
// each thread forward-substitutes one row).
__global__ void __launch_bounds__(NB) k_chol_trsm(int ld, int k, double* __restrict__ A) {
  __shared__ double L[NB * (NB + 1)];
  const int i = k + 1 + blockIdx.x;
  const double* Lkk = A + (int64_t(k) * NB) * ld + k * NB;
  for (int x = threadIdx.x; x < NB * NB; x += NB) L[(x / NB) * (NB + 1) + x % NB] = Lkk[int64_t(x / NB) * ld + x % NB];
  __syncthreads();
  double* row = A + (int64_t(i) * NB + threadIdx.x) * ld + k * NB;
  double x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = row[c];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double s = x[c];
#pragma unroll
    for (int j = 0; j < c; ++j) s -= x[j] * L[c * (NB + 1) + j];
    x[c] = s / L[c * (NB + 1) + c];
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = x[c];
}

// Trailing update A_ij -= L_ik L_jk^T for k < j <= i, one CTA (4 warps) per
// 64x64 tile, fp64 tensor cores (DMMA m8n8k4).  Warp w owns rows 16w..16w+15.
__global__ void __launch_bounds__(128) k_chol_syrk(int ld, int k, int nt, double* __restrict__ A) {
  extern __shared__ double syrk_sm[];
  double* Li = syrk_sm;
  double* Lj = syrk_sm + NB * LDS;
  // decode (i, j) with k < j <= i < nt from the linear tile index
  const int m = nt - k - 1;
  int t = blockIdx.x, ii = 0;
  while (t >= ii + 1) { t -= ii + 1; ++ii; }
  const int i = k + 1 + ii, j = k + 1 + t;
  (void)m;
  const double* Aik = A + (int64_t(i) * NB) * ld + k * NB;
  const double* Ajk = A + (int64_t(j) * NB) * ld + k * NB;
  for (int x = threadIdx.x; x < NB * NB; x += 128) {
    const int r = x / NB, c = x % NB;
    Li[r * LDS + c] = Aik[int64_t(r) * ld + c];
    Lj[r * LDS + c] = Ajk[int64_t(r) * ld + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  double acc[2][8][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
#pragma unroll 4
  for (int k0 = 0; k0 < NB; k0 += 4) {
    double af[2], bf[8];
#pragma unroll
    for (int a = 0; a < 2; ++a) af[a] = Li[(16 * warp + 8 * a + fr) * LDS + k0 + fc];
#pragma unroll
    for (int b = 0; b < 8; ++b) bf[b] = Lj[(8 * b + fr) * LDS + k0 + fc];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
  }
  double* Aij = A + (int64_t(i) * NB) * ld + j * NB;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = 16 * warp + 8 * a + fr, c = 8 * b + 2 * fc;
      double2* p = reinterpret_cast<double2*>(Aij + int64_t(r) * ld + c);
      double2 v = *p;
      v.x -= acc[a][b][0];
      v.y -= acc[a][b][1];
      *p = v;
    }
}

// Triangular solves with the tiled factor, one right-hand side.
// forward:  solve L_kk y_k = b_k (one CTA), then b_i -= L_ik y_k for i > k.
// backward: solve L_kk^T x_k = y_k, then y_i -= L_ki^T x_k for i < k.
__global__ void __launch_bounds__(NB) k_tri_diag(int ld, int k, int transpose, const double* __restrict__ A,
                                                  double* __restrict__ b) {
  __shared__ double L[NB * (NB + 1)];
  __shared__ double x[NB];
  const double* Lkk = A + (int64_t(k) * NB) * ld + k * NB;
  for (int i = threadIdx.x; i < NB * NB; i += NB) L[(i / NB) * (NB + 1) + i % NB] = Lkk[int64_t(i / NB) * ld + i % NB];
  x[threadIdx.x] = b[k * NB + threadIdx.x];
  __syncthreads();
  if (!transpose) {
    for (int j = 0; j < NB; ++j) {
      if (threadIdx.x == j) x[j] /= L[j * (NB + 1) + j];
      __syncthreads();
      if (threadIdx.x > j) x[threadIdx.x] -= L[threadIdx.x * (NB + 1) + j] * x[j];
      __syncthreads();
    }
  } else {
    for (int j = NB - 1; j >= 0; --j) {
      if (threadIdx.x == j) x[j] /= L[j * (NB + 1) + j];
      __syncthreads();
      if (threadIdx.x < j) x[threadIdx.x] -= L[j * (NB + 1) + threadIdx.x] * x[j];
      __syncthreads();
    }
  }
  b[k * NB + threadIdx.x] = x[threadIdx.x];
}

__global__ void __launch_bounds__(NB) k_tri_update(int ld, int k, int transpose, const double* __restrict__ A,
                                                    double* __restrict__ b) {
  __shared__ double xk[NB];
  xk[threadIdx.x] = b[k * NB + threadIdx.x];
  __syncthreads();
  double s = 0.0;
  if (!transpose) {
    const int i = k + 1 + blockIdx.x;
    const double* row = A + (int64_t(i) * NB + threadIdx.x) * ld + k * NB;
#pragma unroll 8
    for (int c = 0; c < NB; ++c) s += row[c] * xk[c];
    b[i * NB + threadIdx.x] -= s;
  } else {
    const int i = blockIdx.x;  // i < k
    const double* col = A + (int64_t(k) * NB) * ld + i * NB + threadIdx.x;
#pragma unroll 8
    for (int c = 0; c < NB; ++c) s += col[int64_t(c) * ld] * xk[c];
    b[i * NB + threadIdx.x] -= s;
  }
}

// ---------------------------------------------------- one-launch Cholesky ---
// The whole factorisation + both triangular solves as ONE cooperative kernel (r02d).  The multi-launch version
// above costs three launches per 64-wide tile column plus four per column for the solves (profiles/r02b: at the
// 900 unknowns of the EuRoC map, panel 0.99 ms + triangular solves 0.82 ms + TRSM 0.53 ms of a 3.4 ms LM step, all
// launch / barrier latency).  Here, per tile column k:
//   * the diagonal tile is factored by ONE CTA in shared memory as an unscaled elimination of [A_kk | I] with one
//     block barrier per pivot; it yields L_kk AND W_k = L_kk^-1 (rows scaled by 1 / sqrt(pivot) at the end);
//   * with W_k explicit, the panel is a GEMM, L_ik = A_ik W_k^T, on the FP64 tensor cores — no per-row substitution
//     chain — and the forward substitution rides along: y_k = W_k b_k, b_i -= L_ik y_k from the accumulators;
//   * the trailing update A_ij -= L_ik L_jk^T (DMMA) is spread over the grid; the CTA that owns tile (k+1, k+1)
//     factors it right behind its update, so a column costs two grid barriers, not three;
//   * the backward substitution is mat-vecs with W_k^T, one grid barrier per column.
// Everything other CTAs produced is read with ld.global.cg (L1 is not coherent across SMs).
constexpr int kCoopThreads = 256;
constexpr int TLD = NB + 1;  // row stride of the diagonal-phase tiles
constexpr size_t kCoopSmem = (2 * NB * LDS + 2 * NB) * sizeof(double);  // two operand tiles (>= two TLD tiles) + 2 vectors

struct CholCoopArgs {
  double* A;   // [ld][ld], lower tiles used
  double* b;   // [ld] right-hand side in, solution out
  double* W;   // [nt][64][64] inverses of the diagonal factor tiles
  double* y;   // [ld] forward-substituted right-hand side
  int ld, nt;
  int* fail;
};

__device__ __forceinline__ void coop_stage(double* dst, const double* __restrict__ src, int src_ld) {
  for (int x = threadIdx.x; x < NB * NB / 2; x += kCoopThreads) {
    const int r = x >> 5, c2 = x & 31;
    const double2 v = __ldcg(reinterpret_cast<const double2*>(src + int64_t(r) * src_ld) + c2);
    *reinterpret_cast<double2*>(dst + r * LDS + 2 * c2) = v;
  }
}

// acc = P Q^T for two 64 x 64 row-major tiles in shared memory (stride LDS); warp w owns rows 8 w .. 8 w + 7,
// lane (g, t) ends with C[8 w + g][8 b + 2 t + {0, 1}] in acc[b]
__device__ __forceinline__ void coop_pqt(const double* P, const double* Q, double (&acc)[8][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int b = 0; b < 8; ++b) { acc[b][0] = 0.0; acc[b][1] = 0.0; }
#pragma unroll 4
  for (int k0 = 0; k0 < NB; k0 += 4) {
    const double af = P[(8 * warp + fr) * LDS + k0 + fc];
#pragma unroll
    for (int b = 0; b < 8; ++b) dmma8x8x4(acc[b][0], acc[b][1], af, Q[(8 * b + fr) * LDS + k0 + fc]);
  }
}

#ifdef PBA_CHOL_TIMING
__device__ long long g_chol_t[12];
#define CT0() long long _ct = clock64()
#define CT(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _n = clock64(); g_chol_t[i] += _n - _ct; _ct = _n; } } while (0)
#define CTP(i) do { if (threadIdx.x == 252 && blockIdx.x == 0) { const long long _n = clock64(); g_chol_t[i] += _n - _cp; _cp = _n; } } while (0)
#else
#define CT0()
#define CT(i)
#define CTP(i)
#endif

__device__ __forceinline__ double coop_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// Diagonal tile k: A_kk = L L^T; L goes back to A, W = L^-1 to a.W.
// Unscaled Gaussian elimination on [A_kk | I] with the rows in REGISTERS: thread (r, ph) owns the 16 columns
// c = ph + 4 q of row r of ONE merged 64 x 64 array — column c holds the A part while c > j (kept symmetric, so
// "column j" is simply row j) and the I part once c <= j.  Per pivot j the four owners of row j publish it to
// shared memory (double-buffered: one block barrier per pivot), every row r > j subtracts (a_rj / a_jj) x row j,
// and the slot of column j — final, L_rj = a_rj / sqrt(a_jj) goes to the output tile — is re-used for the I part
// (-a_rj / a_jj).  At the end row r of the I part scaled by 1 / sqrt(pivot_r) is row r of W = L^-1.  Shared-memory
// traffic per pivot: 64 stores + broadcast loads (the first version kept both arrays in shared memory and
// read-modify-wrote them: ~0.4 M accesses per tile, 33 us; fully predicated 16-slot loops were slower still).
__device__ void coop_diag(const CholCoopArgs& a, int k, double* sm) {
  double* Lout = sm;             // [64][TLD]
  double* prow = sm + NB * TLD;  // [2][64]
  double* piv = prow + 2 * NB;   // [64]
  double* Akk = a.A + (int64_t(k) * NB) * a.ld + k * NB;
  const int r = threadIdx.x >> 2, ph = threadIdx.x & 3;
  double t[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int c = ph + 4 * q;  // the lower triangle is the reference copy: (r, c) or its mirror
    t[q] = __ldcg(c <= r ? Akk + int64_t(r) * a.ld + c : Akk + int64_t(c) * a.ld + r);
  }
#ifdef PBA_CHOL_TIMING
  long long _cp = clock64();
#endif
  for (int j = 0; j < NB; ++j) {
    double* pr = prow + (j & 1) * NB;
    if (r == j) {
#pragma unroll
      for (int q = 0; q < 16; ++q) pr[ph + 4 * q] = t[q];
    }
    CTP(8);
    __syncthreads();
    CTP(9);
    double d = pr[j];
    const bool bad = !(d > 0.0);
    if (bad) d = 1.0;
    if (threadIdx.x == 0) { piv[j] = d; if (bad) *a.fail = 1; }
    if (r > j) {
      // the only non-trivial operation on the per-pivot chain is this reciprocal (hardware approximation + two
      // Newton steps); the square roots wait until the end.  With 1 / sqrt(d) and 1 / d as library calls per pivot
      // a tile took 96 k cycles (in-kernel stamps), 90 % of the whole factorisation.
      // row j first, all 16 loads in flight (left to itself the compiler alternated load -> dependent FMA through
      // two registers: sixteen shared-memory latencies in series, ~1,000 cycles per pivot)
      double p[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) p[q] = pr[ph + 4 * q];
      const double f = pr[r] * coop_rcp(d);
      CTP(10);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int c = ph + 4 * q;
        const double u = fma(-f, p[q], t[q]);
        if (c == j) Lout[r * TLD + j] = t[q];  // a_rj of the j-th Schur complement: L_rj = a_rj / sqrt(pivot_j), scaled below
        t[q] = c == j ? -f : u;
      }
      CTP(11);
    }
  }
  __syncthreads();
  double* isv = prow;  // 1 / sqrt(pivot), per column
  if (threadIdx.x < NB) isv[threadIdx.x] = 1.0 / sqrt(piv[threadIdx.x]);
  __syncthreads();
  double* Wk = a.W + int64_t(k) * NB * NB;
  {
    const double isr = isv[r];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int c = ph + 4 * q;
      Wk[r * NB + c] = c < r ? t[q] * isr : (c == r ? isr : 0.0);
    }
  }
  for (int x = threadIdx.x; x < NB * NB; x += kCoopThreads) {
    const int rr = x >> 6, c = x & 63;
    Akk[int64_t(rr) * a.ld + c] = c < rr ? Lout[rr * TLD + c] * isv[c] : (c == rr ? piv[c] * isv[c] : 0.0);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCoopThreads) k_chol_coop(const CholCoopArgs a) {
  extern __shared__ __align__(16) double coop_sm[];
  cg::grid_group grid = cg::this_grid();
  double* P = coop_sm;
  double* Q = coop_sm + NB * LDS;
  double* vec = coop_sm + 2 * NB * LDS;  // [64] y_k / x_k, [64] scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int G = gridDim.x, cta = blockIdx.x;
  const int nt = a.nt;
  CT0();
  if (cta == 0) coop_diag(a, 0, coop_sm);
  __threadfence();
  grid.sync();
  CT(0);
  for (int k = 0; k < nt; ++k) {
    const int m = nt - k - 1;
    // ---- panel + forward substitution ----
    if (cta == 0 || cta < m) {
      for (int x = threadIdx.x; x < NB * NB / 2; x += kCoopThreads) {  // W_k: dense [64][64]
        const int r = x >> 5, c2 = x & 31;
        *reinterpret_cast<double2*>(Q + r * LDS + 2 * c2) = __ldcg(reinterpret_cast<const double2*>(a.W + int64_t(k) * NB * NB + r * NB) + c2);
      }
      if (threadIdx.x < NB) vec[NB + threadIdx.x] = __ldcg(a.b + k * NB + threadIdx.x);
      __syncthreads();
      {
        // y_k = W_k b_k: four lanes per row (a 64-long dependent chain per thread cost ~3.8 k cycles per column)
        const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
        double s = 0.0;
        for (int c = part; c <= r; c += 4) s += Q[r * LDS + c] * vec[NB + c];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          vec[r] = s;
          if (cta == 0) a.y[k * NB + r] = s;
        }
      }
      for (int i = k + 1 + cta; i < nt; i += G) {
        __syncthreads();  // y_k visible; P free again
        double* Aik = a.A + (int64_t(i) * NB) * a.ld + k * NB;
        coop_stage(P, Aik, a.ld);
        __syncthreads();
        double acc[8][2];
        coop_pqt(P, Q, acc);  // L_ik = A_ik W_k^T
        double part = 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          *reinterpret_cast<double2*>(Aik + int64_t(8 * warp + fr) * a.ld + 8 * b + 2 * fc) = make_double2(acc[b][0], acc[b][1]);
          part += acc[b][0] * vec[8 * b + 2 * fc] + acc[b][1] * vec[8 * b + 2 * fc + 1];
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (fc == 0) a.b[i * NB + 8 * warp + fr] = __ldcg(a.b + i * NB + 8 * warp + fr) - part;
      }
    }
    CT(1);
    __threadfence();
    grid.sync();
    CT(2);
    // ---- trailing update; CTA 0 takes tile (k+1, k+1) and factors it right away ----
    if (m > 0) {
      const int n_tiles = m * (m + 1) / 2;
      const int first = G > 1 ? (cta == 0 ? 0 : cta) : 0;            // tile 0 belongs to CTA 0; the others share 1..
      const int stride = G > 1 ? (cta == 0 ? n_tiles : G - 1) : 1;   // (a single-CTA grid walks all of them)
      for (int t = first; t < n_tiles; t += stride) {
        int ii = int((sqrtf(8.0f * float(t) + 1.0f) - 1.0f) * 0.5f);
        while (ii * (ii + 1) / 2 > t) --ii;
        while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
        const int jj = t - ii * (ii + 1) / 2;
        const int i = k + 1 + ii, j = k + 1 + jj;
        __syncthreads();
        coop_stage(P, a.A + (int64_t(i) * NB) * a.ld + k * NB, a.ld);
        coop_stage(Q, a.A + (int64_t(j) * NB) * a.ld + k * NB, a.ld);
        __syncthreads();
        double acc[8][2];
        coop_pqt(P, Q, acc);
        double* Aij = a.A + (int64_t(i) * NB) * a.ld + j * NB;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          double2* q = reinterpret_cast<double2*>(Aij + int64_t(8 * warp + fr) * a.ld + 8 * b + 2 * fc);
          double2 v = __ldcg(q);
          v.x -= acc[b][0]; v.y -= acc[b][1];
          *q = v;
        }
        if (G == 1 && t == 0) { __threadfence(); __syncthreads(); coop_diag(a, k + 1, coop_sm); }
      }
      CT(3);
      if (G > 1 && cta == 0) { __threadfence(); __syncthreads(); coop_diag(a, k + 1, coop_sm); }
      CT(4);
    }
    __threadfence();
    grid.sync();
    CT(5);
  }
  // ---- backward substitution: x_k = W_k^T y_k, y_i -= L_ki^T x_k for i < k ----
  for (int k = nt - 1; k >= 0; --k) {
    if (cta == 0 || cta < k) {
      __syncthreads();
      for (int x = threadIdx.x; x < NB * NB / 2; x += kCoopThreads) {
        const int r = x >> 5, c2 = x & 31;
        *reinterpret_cast<double2*>(Q + r * LDS + 2 * c2) = __ldcg(reinterpret_cast<const double2*>(a.W + int64_t(k) * NB * NB + r * NB) + c2);
      }
      if (threadIdx.x < NB) vec[NB + threadIdx.x] = __ldcg(a.y + k * NB + threadIdx.x);
      __syncthreads();
      {
        // x_k = W_k^T y_k, four lanes per column
        const int c = threadIdx.x >> 2, part = threadIdx.x & 3;
        double s = 0.0;
        for (int r = c + part; r < NB; r += 4) s += Q[r * LDS + c] * vec[NB + r];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          vec[c] = s;
          if (cta == 0) a.b[k * NB + c] = s;
        }
      }
      for (int i = cta; i < k; i += G) {
        __syncthreads();
        coop_stage(P, a.A + (int64_t(k) * NB) * a.ld + i * NB, a.ld);  // L_ki
        __syncthreads();
        // 4 partial sums per column, combined in fixed order
        const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
        double s = 0.0;
#pragma unroll 4
        for (int r = 16 * q; r < 16 * q + 16; ++r) s += P[r * LDS + c] * vec[r];
        __syncthreads();
        P[threadIdx.x] = s;  // P is consumed: reuse its first 256 doubles
        __syncthreads();
        if (threadIdx.x < NB) {
          const double tot = (P[threadIdx.x] + P[64 + threadIdx.x]) + (P[128 + threadIdx.x] + P[192 + threadIdx.x]);
          a.y[i * NB + threadIdx.x] = __ldcg(a.y + i * NB + threadIdx.x) - tot;
        }
      }
    }
    CT(6);
    __threadfence();
    grid.sync();
    CT(7);
  }
}

int chol_coop_grid(int device, int nt) {
  static std::mutex mu;
  static std::vector<std::pair<int, int>> cache;  // (device, CTAs that can be co-resident)
  int cap = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (auto& c : cache) if (c.first == device) cap = c.second;
    if (!cap) {
      int n_sm = 0, per_sm = 0;
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
      ensure_dynamic_smem((const void*)k_chol_coop, kCoopSmem, device);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_coop, kCoopThreads, kCoopSmem);
      cap = n_sm * (per_sm > 0 ? 1 : 0);  // one CTA per SM: a grid barrier costs less with fewer CTAs
      cache.emplace_back(device, cap);
    }
  }
  const int m = nt - 1;
  const int want = std::max(1, m * (m + 1) / 2 + 1);
  return std::max(1, std::min(cap, want));
}

// ------------------------------------------------------------------- PCG ---
// Inverse of the (SPD) diagonal blocks: the block-Jacobi preconditioner.
__global__ void k_block_inverse(int cd, int n_slots, const int* __restrict__ diag_blk, const double* __restrict__ S,
                                double* __restrict__ inv) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_slots) return;
  double M[64], X[64];
  const double* src = S + int64_t(diag_blk[a]) * cd * cd;
  for (int i = 0; i < cd * cd; ++i) M[i] = src[i];
  // Cholesky M = L L^T (in place, lower)
  for (int j = 0; j < cd; ++j) {
    double d = M[j * cd + j];
    for (int k = 0; k < j; ++k) d -= M[j * cd + k] * M[j * cd + k];
    d = sqrt(fmax(d, 1e-300));
    M[j * cd + j] = d;
    for (int i = j + 1; i < cd; ++i) {
      double s = M[i * cd + j];
      for (int k = 0; k < j; ++k) s -= M[i * cd + k] * M[j * cd + k];
      M[i * cd + j] = s / d;
    }
  }
  // X = M^-1 column by column
  for (int c = 0; c < cd; ++c) {
    double y[8];
    for (int i = 0; i < cd; ++i) {
      double s = i == c ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= M[i * cd + k] * y[k];
      y[i] = s / M[i * cd + i];
    }
    for (int i = cd - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < cd; ++k) s -= M[k * cd + i] * y[k];
      y[i] = s / M[i * cd + i];
    }
    for (int i = 0; i < cd; ++i) X[i * cd + c] = y[i];
  }
  for (int i = 0; i < cd * cd; ++i) inv[int64_t(a) * cd * cd + i] = X[i];
}

struct PcgArgs {
  int cd, n_slots, max_iter;
  double tol;
  const int* row_ptr;
  const int* row_blk;
  const int* row_col;
  const uint8_t* row_trans;
  const double* S;
  const double* inv;
  const double* b;
  double* x;
  double* r;
  double* z;
  double* p;
  double* Ap;
  double* part;     // [3][gridDim.x]
  double* out;      // [0] iterations, [1] final sqrt(rz/rz0)
};

__device__ __forceinline__ double grid_total(const double* part, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += part[i];
  return s;
}

// Deterministic: every per-CTA partial is produced by one CTA and summed in
// index order by everyone.
__device__ __forceinline__ void cta_partial(double v, double* dst, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < int(blockDim.x >> 5); ++i) s += sm[i];
    *dst = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_pcg(const PcgArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[8];
  const int cd = a.cd, n = a.n_slots * a.cd;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  const int gwarp = gtid >> 5, nwarp = gsz >> 5, lane = threadIdx.x & 31;
  const int fr = lane >> 2, fq = lane & 3;
  const int G = gridDim.x;

  // x = 0, r = b, z = M^-1 r, p = z, rz = r.z
  double loc = 0.0;
  for (int s = gwarp; s < a.n_slots; s += nwarp) {
    // block-Jacobi apply: z_s = inv_s r_s ; lanes (fr, fq): row fr, columns 2fq,2fq+1
    double v = 0.0;
    if (fr < cd) {
      const double* inv = a.inv + int64_t(s) * cd * cd + fr * cd;
      for (int c = 2 * fq; c < 2 * fq + 2 && c < cd; ++c) v += inv[c] * a.b[s * cd + c];
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (fq == 0 && fr < cd) {
      const int i = s * cd + fr;
      const double ri = a.b[i];
      a.x[i] = 0.0; a.r[i] = ri; a.z[i] = v; a.p[i] = v;
      loc += ri * v;
    }
  }
  cta_partial(loc, a.part + blockIdx.x, sm);
  grid.sync();
  double rz = grid_total(a.part, G);
  const double rz0 = rz;
  int it = 0;
  if (rz0 > 0.0) {
    for (it = 0; it < a.max_iter; ++it) {
      // Ap = S p (symmetric block rows), pAp
      loc = 0.0;
      for (int s = gwarp; s < a.n_slots; s += nwarp) {
        double v = 0.0;
        if (fr < cd) {
          for (int k = a.row_ptr[s]; k < a.row_ptr[s + 1]; ++k) {
            const double* blk = a.S + int64_t(a.row_blk[k]) * cd * cd;
            const double* pc = a.p + a.row_col[k] * cd;
            if (!a.row_trans[k]) {
              for (int c = 2 * fq; c < 2 * fq + 2 && c < cd; ++c) v += blk[fr * cd + c] * pc[c];
            } else {
              for (int c = 2 * fq; c < 2 * fq + 2 && c < cd; ++c) v += blk[c * cd + fr] * pc[c];
            }
          }
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (fq == 0 && fr < cd) {
          a.Ap[s * cd + fr] = v;
          loc += v * a.p[s * cd + fr];
        }
      }
      cta_partial(loc, a.part + G + blockIdx.x, sm);
      grid.sync();
      const double pAp = grid_total(a.part + G, G);
      const double alpha = rz / pAp;
      // x += alpha p ; r -= alpha Ap ; z = M^-1 r ; rz_new
      loc = 0.0;
      for (int s = gwarp; s < a.n_slots; s += nwarp) {
        double rn = 0.0;
        if (fq == 0 && fr < cd) {
          const int i = s * cd + fr;
          a.x[i] += alpha * a.p[i];
          rn = a.r[i] - alpha * a.Ap[i];
          a.r[i] = rn;
        }
        // broadcast the slot's new residual: lane 4c holds row c (all lanes shuffle)
        const int c0 = 2 * fq, c1 = 2 * fq + 1;
        const double rc0 = __shfl_sync(0xffffffffu, rn, 4 * (c0 < cd ? c0 : 0));
        const double rc1 = __shfl_sync(0xffffffffu, rn, 4 * (c1 < cd ? c1 : 0));
        double v = 0.0;
        if (fr < cd) {
          const double* inv = a.inv + int64_t(s) * cd * cd + fr * cd;
          if (c0 < cd) v += inv[c0] * rc0;
          if (c1 < cd) v += inv[c1] * rc1;
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (fq == 0 && fr < cd) {
          a.z[s * cd + fr] = v;
          loc += rn * v;
        }
      }
      cta_partial(loc, a.part + 2 * G + blockIdx.x, sm);
      grid.sync();
      const double rz_new = grid_total(a.part + 2 * G, G);
      if (!(rz_new > a.tol * a.tol * rz0)) { rz = rz_new; ++it; break; }
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int i = gtid; i < n; i += gsz) a.p[i] = a.z[i] + beta * a.p[i];
      grid.sync();
    }
  }
  if (gtid == 0) {
    a.out[0] = double(it);
    a.out[1] = rz0 > 0.0 ? sqrt(fabs(rz) / rz0) : 0.0;
  }
}

}  // namespace

// In-place tiled Cholesky + solve of the padded dense system (ld = multiple of 64).  `work` holds
// dense_work_size(ld) doubles (inverse diagonal tiles + the forward-substituted right-hand side).
// One cooperative launch (k_chol_coop); PBA_CHOL_V1=1 keeps the multi-launch version for A/B runs.
size_t dense_work_size(int ld) { return size_t(ld / NB) * NB * NB + size_t(ld); }

pba_status dense_cholesky_solve(Handle* h, double* A, double* b, int ld, int* fail_dev, double* work) {
  const int nt = ld / NB;
  static const bool force_v1 = getenv("PBA_CHOL_V1") != nullptr;
  int device = h->device;
  if (!force_v1 && work) {
    const int grid = chol_coop_grid(device, nt);
    if (grid > 0) {
      CholCoopArgs a;
      a.A = A; a.b = b; a.W = work; a.y = work + size_t(nt) * NB * NB; a.ld = ld; a.nt = nt; a.fail = fail_dev;
      void* args[] = {(void*)&a};
      PBA_CUDA_OK(ensure_dynamic_smem((const void*)k_chol_coop, kCoopSmem, device));
      h->stats.begin(K_CHOL_SYRK, h->stream);
      const cudaError_t e = cudaLaunchCooperativeKernel((void*)k_chol_coop, dim3(grid), dim3(kCoopThreads), args, kCoopSmem, h->stream);
      h->stats.end(h->stream);
      if (e != cudaSuccess) return map_cuda(e);
#ifdef PBA_CHOL_TIMING
      {
        static int calls = 0;
        if (++calls == 40) {
          cudaStreamSynchronize(h->stream);
          long long t[12];
          cudaMemcpyFromSymbol(t, g_chol_t, sizeof(t));
          const char* names[12] = {"diag0+sync", "panel", "sync1", "trailing(tile)", "diag(k+1)", "sync2", "backward", "sync3",
                                   "pivot: publish", "pivot: barrier", "pivot: d, rcp, loads", "pivot: update"};
          for (int i = 0; i < 12; ++i) fprintf(stderr, "[chol] %-16s %10.1f cycles/solve\n", names[i], double(t[i]) / calls);
          fprintf(stderr, "[chol] nt %d grid %d\n", nt, grid);
        }
      }
#endif
      return PBA_OK;
    }
  }
  constexpr size_t kSyrkSmem = 2 * NB * LDS * sizeof(double);  // 69,632 B: needs the opt-in limit
  for (int k = 0; k < nt; ++k) {
    PBA_LAUNCH(h, K_CHOL_PANEL, k_chol_diag, dim3(1), dim3(256), 0, ld, k, A, fail_dev);
    const int m = nt - k - 1;
    if (m > 0) {
      PBA_LAUNCH(h, K_CHOL_TRSM, k_chol_trsm, dim3(m), dim3(NB), 0, ld, k, A);
      PBA_LAUNCH(h, K_CHOL_SYRK, k_chol_syrk, dim3(m * (m + 1) / 2), dim3(128), kSyrkSmem, ld, k, nt, A);
    }
  }
  for (int k = 0; k < nt; ++k) {
    PBA_LAUNCH(h, K_CHOL_SOLVE, k_tri_diag, dim3(1), dim3(NB), 0, ld, k, 0, A, b);
    if (k + 1 < nt) { PBA_LAUNCH(h, K_CHOL_SOLVE, k_tri_update, dim3(nt - k - 1), dim3(NB), 0, ld, k, 0, A, b); }
  }
  for (int k = nt - 1; k >= 0; --k) {
    PBA_LAUNCH(h, K_CHOL_SOLVE, k_tri_diag, dim3(1), dim3(NB), 0, ld, k, 1, A, b);
    if (k > 0) { PBA_LAUNCH(h, K_CHOL_SOLVE, k_tri_update, dim3(k), dim3(NB), 0, ld, k, 1, A, b); }
  }
  return PBA_OK;
}

int dense_ld(int n) { return ((n + NB - 1) / NB) * NB; }

pba_status launch_cholesky_rcs(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim == 0) return PBA_OK;
  const int ld = dense_ld(z.dim);
  if (h->dense.n < size_t(ld) * ld + ld + dense_work_size(ld)) {
    PBA_CUDA_OK(h->dense.alloc(size_t(ld) * ld + ld + dense_work_size(ld)));
  }
  double* A = h->dense.p;
  double* b = A + size_t(ld) * ld;
  const double* S = h->rcs.p;
  const double* rhs = S + z.n_blocks * z.cd * z.cd;
  const int64_t tot = int64_t(ld) * ld;
  PBA_LAUNCH(h, K_DENSE_FILL, k_dense_init, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, z.dim, ld, A, b);
  PBA_LAUNCH(h, K_DENSE_FILL, k_dense_fill, dim3((unsigned)z.n_blocks), dim3(64), 0, z.cd, z.n_blocks, ld,
             h->d_blk_row.p, h->d_blk_col.p, S, A);
  PBA_CUDA_OK(cudaMemcpyAsync(b, rhs, sizeof(double) * z.dim, cudaMemcpyDeviceToDevice, h->stream));
  PBA_CUDA_OK(cudaMemsetAsync(h->chol_fail.p, 0, sizeof(int), h->stream));
  pba_status st = dense_cholesky_solve(h, A, b, ld, h->chol_fail.p, b + ld);
  if (st != PBA_OK) return st;
  PBA_CUDA_OK(cudaMemcpyAsync(h->y_cam.p, b, sizeof(double) * z.dim, cudaMemcpyDeviceToDevice, h->stream));
  return PBA_OK;
}

pba_status launch_pcg_rcs(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim == 0) return PBA_OK;
  const double* S = h->rcs.p;
  const double* rhs = S + z.n_blocks * z.cd * z.cd;
  PBA_LAUNCH(h, K_PCG, k_block_inverse, dim3((z.n_slots + 63) / 64), dim3(64), 0, z.cd, z.n_slots, h->d_diag_blk.p, S,
             h->blk_inv.p);
  int grid = h->pcg_grid;
  const int need = (z.n_slots * 32 + 255) / 256;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  PcgArgs a;
  a.cd = z.cd; a.n_slots = z.n_slots; a.max_iter = h->opt.pcg_max_iterations; a.tol = h->opt.pcg_tolerance;
  a.row_ptr = h->row_ptr.p; a.row_blk = h->row_blk.p; a.row_col = h->row_col.p; a.row_trans = h->row_trans.p;
  a.S = S; a.inv = h->blk_inv.p; a.b = rhs;
  double* ws = h->pcg_ws.p;
  a.x = h->y_cam.p; a.r = ws; a.z = ws + z.dim; a.p = ws + 2 * z.dim; a.Ap = ws + 3 * z.dim;
  a.part = ws + 4 * z.dim;
  a.out = h->scalars.p + S_PCG_ITERS;
  void* args[] = {(void*)&a};
  h->stats.begin(K_PCG, h->stream);
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_pcg, dim3(grid), dim3(256), args, 0, h->stream);
  h->stats.end(h->stream);
  if (e != cudaSuccess) return map_cuda(e);
  return PBA_OK;
}

int pcg_max_grid(int device) {
  // the occupancy query costs a good fraction of a millisecond; the answer is a property of the device
  static std::mutex mu;
  static std::vector<std::pair<int, int>> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& kv : cache) if (kv.first == device) return kv.second;
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg, 256, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  cache.emplace_back(device, sms * per_sm);
  return sms * per_sm;
}

}  // namespace pba

// ------------------------------------------------------ band Cholesky ------
// Exact solve for block-banded reduced camera systems (windowed covisibility:
// keyframe i only shares landmarks with i +- bw).  ONE CTA walks the block
// columns; the active (bw+1) x (bw+1) block window of the trailing matrix lives
// in shared memory as a ring (block (r,c) at [r % B][c % B]), so every column
// costs a handful of __syncthreads phases and no global round trip.  The
// forward substitution is fused into the factorisation; the backward
// substitution streams the stored factor columns back with a register prefetch.
// Replaces the sparse LDL^T of internal/ceres/eigensparse.cc:56-106 for this
// structure.  Work is O(n bw^2 cd^3): 2,000 keyframes, bw = 12 -> ~1.3e8 FMAs.
namespace pba {
namespace {

constexpr int kBandThreads = 256;

template <int CD>
__global__ void __launch_bounds__(kBandThreads) k_band_cholesky(int n_slots, int bw, const int* __restrict__ col_blk,
                                                                 const double* __restrict__ S,
                                                                 const double* __restrict__ rhs,
                                                                 double* __restrict__ Lband, double* __restrict__ x,
                                                                 int* __restrict__ fail) {
  constexpr int BS = CD * CD;
  constexpr int NT = kBandThreads;
  extern __shared__ double sm[];
  const int B = bw + 1;
  double* W = sm;                       // [B][B][BS] ring of the trailing window
  double* Cc = W + size_t(B) * B * BS;  // [B][BS] current block column (updates + original entries)
  double* Lc = Cc + size_t(B) * BS;     // [B][BS] factor column: slot 0 = L_kk^-1, slots i = L_{k+i,k}
  double* R = Lc + size_t(B) * BS;      // [B][CD] ring of the pending right-hand side (later: x window)
  double* yk = R + size_t(B) * CD;      // [CD]
  double* Lkk = yk + CD;                // [BS] diagonal factor (scratch)
  double* idg = Lkk + BS;               // [CD] 1 / diag(L_kk)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < B * B * BS; i += NT) W[i] = 0.0;
  for (int i = tid; i < B * CD; i += NT) R[i] = 0.0;

  // Per-thread element ownership is independent of the column: element idx = tid + q*NT
  // of the (bw+1) x BS column buffer <-> block i_q, entry e_q.
  constexpr int PF = (20 * BS + NT - 1) / NT;  // columns of up to 20 blocks (band_max_bw)
  int ei[PF], ee[PF], et[PF];
#pragma unroll
  for (int q = 0; q < PF; ++q) {
    const int idx = tid + q * NT;
    ei[q] = idx / BS;
    ee[q] = idx % BS;
    et[q] = (ee[q] % CD) * CD + ee[q] / CD;  // transposed entry
  }
  // Trailing-update pairs (i >= j >= 1) owned by this warp: p = warp, warp + 16, ...
  constexpr int NW = NT / 32;
  constexpr int MAXP = (19 * 20 / 2 + NW - 1) / NW;
  int pi[MAXP], pj[MAXP];
#pragma unroll
  for (int q = 0; q < MAXP; ++q) {
    int p = warp + q * NW, i = 1;
    while (p >= i) { p -= i; ++i; }
    pi[q] = i; pj[q] = 1 + p;
  }

  // Original entries of column k = row k of the upper block list, transposed; col_blk[k][i] =
  // block index of (k, k+i) or -1.  Values are prefetched one column ahead and their block
  // indices two columns ahead, so no global latency sits on the per-column critical path.
  double pf[PF];
  int tnext[PF];
  double rpf = 0.0;
  auto load_idx = [&](int kk) {
#pragma unroll
    for (int q = 0; q < PF; ++q) tnext[q] = (kk < n_slots && ei[q] <= bw) ? col_blk[kk * B + ei[q]] : -1;
  };
  auto load_val = [&](int kk) {
#pragma unroll
    for (int q = 0; q < PF; ++q) pf[q] = tnext[q] >= 0 ? S[size_t(tnext[q]) * BS + et[q]] : 0.0;
    if (tid < CD) rpf = kk < n_slots ? rhs[kk * CD + tid] : 0.0;
  };
  load_idx(0);
  load_val(0);
  load_idx(1);
  __syncthreads();

#ifdef PBA_BAND_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tt = clock64();
#define BAND_TICK(i) { long long _n = clock64(); tacc[i] += _n - tt; tt = _n; }
#else
#define BAND_TICK(i)
#endif
  int kb = 0;  // k % B
  for (int k = 0; k < n_slots; ++k) {
    const int nb = min(bw, n_slots - 1 - k);  // blocks below the diagonal in this column
    // A. column k <- accumulated updates of the ring + original entries
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      if (ei[q] <= nb) {
        int rb = kb + ei[q];
        rb = rb >= B ? rb - B : rb;
        Cc[tid + q * NT] = W[(size_t(rb) * B + kb) * BS + ee[q]] + pf[q];
      }
    }
    if (tid < CD) yk[tid] = R[kb * CD + tid] + rpf;
    __syncthreads();
    load_val(k + 1);  // uses the indices fetched during the previous column
    load_idx(k + 2);
    BAND_TICK(0)
    // B. warp 0: Cholesky of the diagonal block in lane 0's registers, inverse of the
    //    factor by columns across lanes, forward substitution y_k = L_kk^-1 (b_k + pending)
    if (tid < 32) {
      if (lane == 0) {
        double L[CD][CD];
#pragma unroll
        for (int r = 0; r < CD; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) L[r][c] = Cc[r * CD + c];
#pragma unroll
        for (int j = 0; j < CD; ++j) {
          double d = L[j][j];
          if (!(d > 0.0)) { *fail = 1; d = 1.0; }
          const double inv = rsqrt(d);
          L[j][j] = d * inv;
          idg[j] = inv;
#pragma unroll
          for (int r = j + 1; r < CD; ++r) L[r][j] *= inv;
#pragma unroll
          for (int c = j + 1; c < CD; ++c)
#pragma unroll
            for (int r = c; r < CD; ++r) L[r][c] -= L[r][j] * L[c][j];
        }
#pragma unroll
        for (int r = 0; r < CD; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) Lkk[r * CD + c] = L[r][c];
      }
      __syncwarp();
      if (lane < CD) {
        // column `lane` of M = L_kk^-1 (lower triangular)
        const int c = lane;
        double m[CD];
#pragma unroll
        for (int r = 0; r < CD; ++r) {
          double s = r == c ? 1.0 : 0.0;
#pragma unroll
          for (int q = 0; q < r; ++q) s -= (q >= c ? Lkk[r * CD + q] * m[q] : 0.0);
          m[r] = r >= c ? s * idg[r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < CD; ++r) Lc[r * CD + c] = m[r];
      }
      __syncwarp();
      if (lane < CD) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < CD; ++c) s += Lc[lane * CD + c] * yk[c];
        __syncwarp(0xffu >> (8 - CD));
        yk[lane] = s;
      }
    }
    __syncthreads();
    BAND_TICK(1)
    // C. L_ik = C_ik L_kk^-T = C_ik M^T.  CD = 8: one DMMA pair per block (this IS the
    //    Cholesky, the one place the path uses the FP64 tensor cores); CD = 6: dot products.
    if constexpr (CD == 8) {
      const int fr = lane >> 2, fc = lane & 3;
      for (int i = 1 + warp; i <= nb; i += NW) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4)
          dmma8x8x4(c0, c1, Cc[i * BS + fr * 8 + k0 + fc], Lc[fr * 8 + k0 + fc]);
        *reinterpret_cast<double2*>(Lc + i * BS + fr * 8 + 2 * fc) = make_double2(c0, c1);
      }
    } else {
      for (int idx = tid; idx < nb * BS; idx += NT) {
        const int i = 1 + idx / BS, e = idx % BS, r = e / CD, c = e % CD;
        const double* crow = Cc + i * BS + r * CD;
        const double* mrow = Lc + c * CD;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < CD; ++q) s += crow[q] * mrow[q];  // M is lower: mrow[q] = 0 for q > c
        Lc[i * BS + e] = s;
      }
    }
    __syncthreads();
    BAND_TICK(2)
    // D. store the factor column; pending rhs -= L_ik y_k; trailing window -= L_ik L_jk^T;
    //    recycle ring row / column k (disjoint from the updated blocks)
    {
      double* Lk = Lband + size_t(k) * B * BS;
#pragma unroll
      for (int q = 0; q < PF; ++q)
        if (ei[q] <= nb) Lk[tid + q * NT] = Lc[tid + q * NT];
      if (tid < CD) x[k * CD + tid] = yk[tid];
      if (tid < nb * CD) {
        const int i = 1 + tid / CD, r = tid % CD;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < CD; ++c) s += Lc[i * BS + r * CD + c] * yk[c];
        int rb = kb + i;
        rb = rb >= B ? rb - B : rb;
        R[rb * CD + r] -= s;
      }
      BAND_TICK(5)
      if constexpr (CD == 8) {
        const int fr = lane >> 2, fc = lane & 3;
        // the warp's pairs are independent: batches of 4 keep several loads / DMMAs in flight
        constexpr int PB = 6;
#pragma unroll 1
        for (int q0 = 0; q0 < MAXP; q0 += PB) {
          if (pi[q0] > nb) break;  // pairs are ordered by i: nothing further applies
          double2 wv[PB];
          double c0[PB], c1[PB], d0[PB], d1[PB];  // two independent accumulators per pair (k = 0..3 / 4..7)
          double2* wp[PB];
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            const int q = q0 + u;
            c0[u] = 0.0; c1[u] = 0.0; d0[u] = 0.0; d1[u] = 0.0;
            const bool on = q < MAXP && pi[q < MAXP ? q : 0] <= nb;
            const int ii = on ? pi[q] : 1, jj = on ? pj[q] : 1;
            int ri = kb + ii, rj = kb + jj;
            ri = ri >= B ? ri - B : ri;
            rj = rj >= B ? rj - B : rj;
            wp[u] = on ? reinterpret_cast<double2*>(W + (size_t(ri) * B + rj) * BS + fr * 8 + 2 * fc) : nullptr;
            wv[u] = on ? *wp[u] : make_double2(0.0, 0.0);
            dmma8x8x4(c0[u], c1[u], Lc[ii * BS + fr * 8 + fc], Lc[jj * BS + fr * 8 + fc]);
            dmma8x8x4(d0[u], d1[u], Lc[ii * BS + fr * 8 + 4 + fc], Lc[jj * BS + fr * 8 + 4 + fc]);
          }
#pragma unroll
          for (int u = 0; u < PB; ++u)
            if (wp[u]) *wp[u] = make_double2(wv[u].x - (c0[u] + d0[u]), wv[u].y - (c1[u] + d1[u]));
        }
      } else {
        constexpr int TS = CD / 2;
        const int npairs = nb * (nb + 1) / 2;
        for (int tix = tid; tix < npairs * 4; tix += NT) {
          int p = tix >> 2, i = 1;
          while (p >= i) { p -= i; ++i; }
          const int j = 1 + p;
          const int r0 = TS * ((tix >> 1) & 1), c0 = TS * (tix & 1);
          double acc[TS][TS];
#pragma unroll
          for (int a = 0; a < TS; ++a)
#pragma unroll
            for (int b = 0; b < TS; ++b) acc[a][b] = 0.0;
          const double* Li = Lc + i * BS + r0 * CD;
          const double* Lj = Lc + j * BS + c0 * CD;
#pragma unroll
          for (int q = 0; q < CD; ++q) {
            double av[TS], bv[TS];
#pragma unroll
            for (int a = 0; a < TS; ++a) { av[a] = Li[a * CD + q]; bv[a] = Lj[a * CD + q]; }
#pragma unroll
            for (int a = 0; a < TS; ++a)
#pragma unroll
              for (int b = 0; b < TS; ++b) acc[a][b] += av[a] * bv[b];
          }
          int ri = kb + i, rj = kb + j;
          ri = ri >= B ? ri - B : ri;
          rj = rj >= B ? rj - B : rj;
          double* Wb = W + (size_t(ri) * B + rj) * BS;
#pragma unroll
          for (int a = 0; a < TS; ++a)
#pragma unroll
            for (int b = 0; b < TS; ++b) Wb[(r0 + a) * CD + c0 + b] -= acc[a][b];
        }
      }
      BAND_TICK(3)
      for (int idx = tid; idx < B * BS; idx += NT) {
        const int o = idx / BS, e = idx % BS;
        W[(size_t(kb) * B + o) * BS + e] = 0.0;
        W[(size_t(o) * B + kb) * BS + e] = 0.0;
      }
      if (tid < CD) R[kb * CD + tid] = 0.0;
    }
    __syncthreads();
    kb = kb + 1 == B ? 0 : kb + 1;
    BAND_TICK(6)
  }

  // ---- backward substitution: x_k = M_k^T (y_k - sum_i L_{k+i,k}^T x_{k+i}) ----
  // x window reuses the R ring; the next factor column is prefetched into registers.
  double* X = R;
  auto prefetch = [&](int kk) {
    const double* Lk = Lband + size_t(kk) * B * BS;
    const int nbk = min(bw, n_slots - 1 - kk);
#pragma unroll
    for (int q = 0; q < PF; ++q) pf[q] = ei[q] <= nbk ? Lk[tid + q * NT] : 0.0;
    if (tid < CD) rpf = x[kk * CD + tid];
  };
  if (n_slots > 0) prefetch(n_slots - 1);
  kb = n_slots > 0 ? (n_slots - 1) % B : 0;
  for (int k = n_slots - 1; k >= 0; --k) {
    const int nb = min(bw, n_slots - 1 - k);
#pragma unroll
    for (int q = 0; q < PF; ++q)
      if (ei[q] <= nb) Lc[tid + q * NT] = pf[q];
    if (tid < CD) yk[tid] = rpf;
    __syncthreads();
    if (k > 0) prefetch(k - 1);
    if (tid < nb * CD) {
      const int i = 1 + tid / CD, c = tid % CD;
      int rb = kb + i;
      rb = rb >= B ? rb - B : rb;
      const double* xv = X + rb * CD;
      double part = 0.0;
#pragma unroll
      for (int r = 0; r < CD; ++r) part += Lc[i * BS + r * CD + c] * xv[r];
      Cc[tid] = part;  // [i-1][c]
    }
    __syncthreads();
    if (tid < 32) {
      double s = 0.0;
      if (lane < CD) {
        // two interleaved partial sums shorten the dependent chain
        double s0 = yk[lane], s1 = 0.0;
        int i = 0;
        for (; i + 1 < nb; i += 2) { s0 -= Cc[i * CD + lane]; s1 -= Cc[(i + 1) * CD + lane]; }
        if (i < nb) s0 -= Cc[i * CD + lane];
        s = s0 + s1;
      }
      // x_c = sum_r M[r][c] s_r  (M lower)
      double v = 0.0;
#pragma unroll
      for (int r = 0; r < CD; ++r) {
        const double sr = __shfl_sync(0xffffffffu, s, r);
        if (lane < CD) v += Lc[r * CD + lane] * sr;
      }
      if (lane < CD) { X[kb * CD + lane] = v; x[k * CD + lane] = v; }
    }
    __syncthreads();
    kb = kb == 0 ? B - 1 : kb - 1;
  }
  BAND_TICK(4)
#ifdef PBA_BAND_TIMING
  if (tid == 0) printf("band timing (cycles/column): A %lld  B %lld  C %lld  D1(store) %lld D2(pairs) %lld D3(zero+sync) %lld | backward/col %lld\n", tacc[0] / n_slots, tacc[1] / n_slots, tacc[2] / n_slots, tacc[5] / n_slots, tacc[3] / n_slots, tacc[6] / n_slots, tacc[4] / n_slots);
#endif
}

}  // namespace

size_t band_smem_bytes(int cd, int bw) {
  const size_t B = bw + 1, BS = size_t(cd) * cd;
  return (B * B * BS + 2 * B * BS + B * cd + 2 * cd + BS) * sizeof(double);
}

// Largest half-bandwidth (in blocks) the shared-memory window supports.
int band_max_bw(int cd) {
  int bw = 0;
  while (band_smem_bytes(cd, bw + 1) <= 200 * 1024 && (bw + 2) * cd * cd <= 20 * cd * cd) ++bw;
  return bw;
}

pba_status launch_band_rcs(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim == 0) return PBA_OK;
  const int bw = h->rcs_bandwidth;
  const size_t smem = band_smem_bytes(z.cd, bw);
  const double* S = h->rcs.p;
  const double* rhs = S + z.n_blocks * z.cd * z.cd;
  PBA_CUDA_OK(cudaMemsetAsync(h->chol_fail.p, 0, sizeof(int), h->stream));
  if (z.cd == 8) {
    PBA_LAUNCH(h, K_BAND_CHOL, k_band_cholesky<8>, dim3(1), dim3(kBandThreads), smem, z.n_slots, bw, h->d_col_blk.p, S, rhs,
               h->band_L.p, h->y_cam.p, h->chol_fail.p);
  } else {
    PBA_LAUNCH(h, K_BAND_CHOL, k_band_cholesky<6>, dim3(1), dim3(kBandThreads), smem, z.n_slots, bw, h->d_col_blk.p, S, rhs,
               h->band_L.p, h->y_cam.p, h->chol_fail.p);
  }
  return PBA_OK;
}

}  // namespace pba
