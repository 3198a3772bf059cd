// bcr2.cu — block cyclic reduction of the block-tridiagonal reduced camera system, second
// generation: the exact, parallel RCS solver on the FP64 tensor cores.
//
// Same algorithm as bcr.cu (super blocks of m = bandwidth keyframes; every level eliminates its odd
// super blocks in parallel; log2(S) levels), replacing Ceres' sequential sparse LDL^T
// (internal/ceres/eigensparse.cc:56-106).  What changed against the first generation (0.82 ms and
// 34 launches at 2,000 keyframes, no tensor-core work):
//   * ONE kernel per level does factor + both triangular solves (k_b2_fs): A_p = L L^T, then
//         [Uh | Vh | yh] = A_p^-1 [B_{p-1} | B_p^T | b_p]
//     with L, the right-hand sides and the solution never leaving the SM (the first generation wrote
//     L, U, V to HBM and read them back in two more launches per level);
//   * the dense work runs on DMMA (mma.sync.m8n8k4.f64): the Cholesky keeps the trailing matrix in
//     registers as 8x8 accumulator fragments (only the current block column goes through shared
//     memory), the triangular solves keep a whole 8-column strip of the right-hand side in one warp's
//     registers, so they need no block-wide barrier at all;
//   * storing Uh = A^-1 B instead of U = L^-1 B makes the back-substitution two mat-vecs per block
//         x_p = yh_p - Uh_p x_{p-1} - Vh_p x_{p+1}
//     (no triangular solve, no factor reload) and all sparse upper levels of it run in one CTA;
//   * the even blocks' update is one kernel of plain DMMA products
//         A'_e = A_e - B_e^T Uh_{e+1} - B_{e-1} Vh_{e-1},  B' = -B_{e+1} Uh_{e+1},  b' likewise.
// Every matrix is padded to Mp = a multiple of 8 (identity on the padding), so the 8x8 tiling does not
// depend on the camera block size (8 photometric, 6 geometric).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "launch.h"
#include "pba_internal.h"

namespace cg = cooperative_groups;

namespace pba {

namespace {

// factor + solve kernels: 8 warps with the full 255-register budget (a warp keeps up to 3 x NBK accumulator
// fragments of the right-hand side in registers); product kernel: 16 warps
constexpr int kB2Threads = 256;
constexpr int kB2Warps = kB2Threads / 32;
constexpr int kB2RedThreads = 512;
constexpr int kB2RedWarps = kB2RedThreads / 32;
constexpr int kB2MaxNbk = 14;  // Mp <= 112: two Mp x (Mp + 4) operand matrices of the reduce kernel must fit shared memory

// Debug only (make EXTRA=-DPBA_B2_TIMING): clock64 deltas of thread 0 of CTA (0, 0), accumulated per phase.
#ifdef PBA_B2_TIMING
__device__ long long g_b2_t[32];
#define B2_T0() long long _bt = clock64()
#define B2_T(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { const long long _n = clock64(); g_b2_t[i] += _n - _bt; _bt = _n; } } while (0)
#else
#define B2_T0()
#define B2_T(i)
#endif

struct B2Level {
  int n;          // super blocks at this level
  double* A;      // [n][Mp*Mp]     symmetric; consumers read the lower 8x8 tiles only
  double* B;      // [n-1][Mp*Mp]   B[p] = A[p+1][p]
  double* b;      // [n][Mp]
  double* Uh;     // [n/2][Mp*Mp]   per odd block p = 2q+1: A_p^-1 B[p-1]
  double* Vh;     // [n/2][Mp*Mp]   A_p^-1 B[p]^T
  double* yh;     // [n/2][Mp]      A_p^-1 b[p]
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cp8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Shared-memory leading dimensions: Mp + 4 (= 4 or 12 mod 16 doubles), so the A / B fragment loads of
// a DMMA (8 rows x 4 consecutive doubles, or 4 rows x 8 consecutive doubles) hit 32 distinct banks.
__host__ __device__ inline int b2_ld(int Mp) { return Mp + 4; }

// dense (Mp x Mp, row-major) global -> shared, rows 16-byte aligned on both sides
__device__ __forceinline__ void b2_stage(double* dst, int ld, const double* __restrict__ src, int Mp) {
  const int half = Mp >> 1;
  for (int i = threadIdx.x; i < Mp * half; i += blockDim.x) {
    const int r = i / half, c2 = i - r * half;
    cp16(dst + r * ld + 2 * c2, src + size_t(r) * Mp + 2 * c2);
  }
}

// ---- level 0 from the block-sparse RCS ----
__global__ void k_b2_build(int cd, int m, int M, int Mp, int64_t n_blocks, int dim, const int* __restrict__ blk_row,
                           const int* __restrict__ blk_col, const double* __restrict__ S, const double* __restrict__ rhs,
                           int n_super, double* __restrict__ A, double* __restrict__ B, double* __restrict__ b) {
  // four RCS blocks (or right-hand-side segments) per 256-thread CTA: one 64-thread CTA each was 24 k CTAs
  // = 15 us of pure launch overhead at 2,000 keyframes
  const int64_t blk = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 6);
  const int e = threadIdx.x & 63;
  const size_t MM = size_t(Mp) * Mp;
  if (blk < n_blocks) {
    if (e >= cd * cd) return;
    const int r = e / cd, c = e % cd;
    const int a = blk_row[blk], bb = blk_col[blk];  // a <= bb
    const int sa = a / m, sb = bb / m;
    const double v = S[blk * cd * cd + e];
    const int ra = (a % m) * cd + r, cb = (bb % m) * cd + c;
    if (sa == sb) {
      A[sa * MM + size_t(ra) * Mp + cb] = v;
      A[sa * MM + size_t(cb) * Mp + ra] = v;
    } else {  // sb == sa + 1: B[sa] = A[sb][sa] = block(a, bb)^T
      B[sa * MM + size_t(cb) * Mp + ra] = v;
    }
  } else {
    // right-hand side; identity on the padding (rows M..Mp-1 of every super block, and the rows of the
    // last super block beyond the system's dimension)
    const int s = int(blk - n_blocks);
    if (s >= n_super) return;
    for (int i = e; i < Mp; i += 64) {
      const int gi = s * M + i;
      const bool real = i < M && gi < dim;
      b[size_t(s) * Mp + i] = real ? rhs[gi] : 0.0;
      if (!real) A[s * MM + size_t(i) * Mp + i] = 1.0;
    }
  }
}

// x (padded super blocks) -> y_cam
__global__ void k_b2_unpad(int M, int Mp, int dim, const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) y[i] = x[size_t(i / M) * Mp + i % M];
}

// ---- 8x8 diagonal block, factored IN REGISTERS by one warp.  The tile sits in the DMMA accumulator layout
// (lane (g, t) holds A[g][2t], A[g][2t+1]); the warp runs Gaussian elimination on [A | I] with row operations
// and WITHOUT scaling the pivot rows: step j subtracts (a_gj / a_jj) x row j from the rows g > j, so that the
// identity becomes W~ = D^1/2 L^-1; the rows are scaled by 1 / sqrt(pivot) once, after the loop, which gives
// W = L^-1 — the only thing the panel and the triangular solves need (L^T itself is never formed).
// The dependent chain per pivot is: shuffle (pivot) -> reciprocal -> one FMA.  The reciprocal is the hardware
// approximation (rcp.approx.ftz.f64, ~20 bits) with two Newton steps (full double precision), four dependent
// FMAs; the products a_gj x (row j) it multiplies are formed meanwhile.  The first version of this function
// (scale row j by rsqrt(a_jj), then multiply, then FMA: library rsqrt + two more multiplies on the chain)
// measured ~2,400 cycles per tile, two thirds of the factorisation's critical path
// (profiles/r02a_b2_cholesky_phases.txt); the single-thread version of the first generation ~2,100.
// Returns W in (w0, w1) = W[g][2t], W[g][2t+1]. ----
__device__ __forceinline__ double b2_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// NOT inlined: these kernels run through their code once per launch — the first version (everything
// unrolled: 13,700 instructions = 219 KB for k_b2_fs<11>, more than the instruction cache) was bound by
// instruction fetch, 45 us a level whatever the amount of arithmetic.
__device__ __noinline__ double2 b2_diag_factor(double a0, double a1, int* fail) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double w0 = g == 2 * t ? 1.0 : 0.0;
  double w1 = g == 2 * t + 1 ? 1.0 : 0.0;
  double mypiv = 1.0;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jt = j >> 1;
    const double colj = (j & 1) ? a1 : a0;                      // column j of a lane whose t == jt
    double ajj = __shfl_sync(0xffffffffu, colj, 4 * j + jt);     // pivot
    const double arj = __shfl_sync(0xffffffffu, colj, 4 * g + jt);  // this row's entry in column j
    const double p0 = __shfl_sync(0xffffffffu, a0, 4 * j + t), p1 = __shfl_sync(0xffffffffu, a1, 4 * j + t);  // pivot row,
    const double q0 = __shfl_sync(0xffffffffu, w0, 4 * j + t), q1 = __shfl_sync(0xffffffffu, w1, 4 * j + t);  // my columns
    if (!(ajj > 0.0)) { bad = true; ajj = 1.0; }
    if (g == j) mypiv = ajj;
    const double r = b2_rcp(ajj);
    const double t0 = arj * p0, t1 = arj * p1, s0 = arj * q0, s1 = arj * q1;  // off the chain: no dependence on r
    if (g > j) { a0 = fma(-t0, r, a0); a1 = fma(-t1, r, a1); w0 = fma(-s0, r, w0); w1 = fma(-s1, r, w1); }
  }
  if (bad && lane == 0) *fail = 1;
  const double inv = rsqrt(mypiv);
  return make_double2(w0 * inv, w1 * inv);
}

// ---- A (Mp x Mp, shared, lower 8x8 tiles valid) <- the panels of its lower Cholesky factor (tiles (I, J),
// I > J); Dinv[J] = inverse of the J-th diagonal block of the factor (the diagonal tiles themselves are
// not stored: nothing reads them).  Right-looking, block size 8, and the DIAGONAL CHAIN HAS ITS OWN WARP:
//   * the trailing matrix lives in registers as DMMA accumulator fragments, spread over the first 7 warps
//     (tile u = I (I + 1) / 2 + K is owned by warp u % 7 until it is due);
//   * a sub-diagonal tile (I, K) is due — written back to shared memory — after the update with block column
//     K - 1 (it is then the input of panel K); a diagonal tile (I, I) one step EARLIER, after column I - 2;
//   * the last warp does nothing but the chain: in step J it scales the one panel tile the next diagonal block
//     needs, P_{J+1,J} = A_{J+1,J} W_J^T, applies the missing update A_{J+1,J+1} -= P P^T and factors the
//     tile in registers (b2_diag_factor) WHILE the other warps update their trailing tiles with column J.
// Two block-wide barriers per step (W_J + the due tiles are visible; panel J is complete).  The critical path
// of a step is barrier + one panel tile + barrier + 2 DMMAs + the 8x8 factor; in the first version the factor,
// the whole panel and the whole trailing update were serialised (~3,600 cycles a step, 40 k for 88 x 88). ----
template <int NBK>
__device__ void b2_cholesky(double* As, double* Dinv, int* fail) {
  constexpr int LD = 8 * NBK + 4;
  constexpr int NT = NBK * (NBK + 1) / 2;
  constexpr int NW = kB2Warps - 1;  // trailing warps; warp NW runs the diagonal chain
  constexpr int TPW = (NT + NW - 1) / NW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const bool chain = warp == NW;
  int ti[TPW], tk[TPW], due[TPW];
  double c[TPW][2];
#pragma unroll
  for (int s = 0; s < TPW; ++s) {
    const int u = warp + s * NW;
    int I = -1, K = 0;
    if (!chain && u < NT) {
      I = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
      while (I * (I + 1) / 2 > u) --I;
      while ((I + 1) * (I + 2) / 2 <= u) ++I;
      K = u - I * (I + 1) / 2;
    }
    ti[s] = I; tk[s] = K;
    due[s] = I < 0 ? -1 : (I == K ? I - 2 : K - 1);  // last block column this warp applies; < 0: never touched here
    c[s][0] = 0.0; c[s][1] = 0.0;
    if (due[s] >= 0) {
      const double2 v = *reinterpret_cast<const double2*>(As + (8 * I + g) * LD + 8 * K + 2 * t);
      c[s][0] = v.x; c[s][1] = v.y;
    }
  }
  B2_T0();
  if (chain) {
    const double2 v = *reinterpret_cast<const double2*>(As + g * LD + 2 * t);
    *reinterpret_cast<double2*>(Dinv + g * 8 + 2 * t) = b2_diag_factor(v.x, v.y, fail);
  }
  __syncthreads();
  B2_T(0);
#pragma unroll 1
  for (int J = 0; J + 1 < NBK; ++J) {
    // panel: P_I = A_IJ W_J^T  (W_J = inverse of the diagonal factor); tile (J + 1, J) belongs to the chain warp
    {
      const double w0 = Dinv[64 * J + g * 8 + t], w1 = Dinv[64 * J + g * 8 + 4 + t];
      for (int I = chain ? J + 1 : J + 2 + warp; I < (chain ? J + 2 : NBK); I += NW) {
        double* tile = As + (8 * I + g) * LD + 8 * J;
        const double a0 = tile[t], a1 = tile[4 + t];
        // two independent DMMAs and an add: a DMMA's result is ~160 cycles away (its four k-steps are dependent
        // FMAs), so chaining the second one on the first's accumulator would double the latency of the tile
        double x[2] = {0.0, 0.0}, x2[2] = {0.0, 0.0};
        dmma(x, a0, w0);
        dmma(x2, a1, w1);
        __syncwarp();
        *reinterpret_cast<double2*>(tile + 2 * t) = make_double2(x[0] + x2[0], x[1] + x2[1]);
      }
    }
    B2_T(1);
    // panel J complete: the trailing warps wait for it; the chain warp only needs its own tile, so it signals
    // and goes on (named barrier 1; barrier 0 = __syncthreads stays the step barrier)
    if (chain) {
      __syncwarp();
      asm volatile("bar.arrive 1, %0;" ::"n"(kB2Threads) : "memory");  // orders this warp's prior shared-memory writes
    } else {
      asm volatile("bar.sync 1, %0;" ::"n"(kB2Threads) : "memory");
    }
    B2_T(2);
    if (chain) {
      // next diagonal tile: every update but column J's is already in it (its owner wrote it back a step ago)
      const double2 v = *reinterpret_cast<const double2*>(As + (8 * (J + 1) + g) * LD + 8 * (J + 1) + 2 * t);
      double d[2] = {v.x, v.y}, d2[2] = {0.0, 0.0};
      const double* pa = As + (8 * (J + 1) + g) * LD + 8 * J;
      const double p0 = pa[t], p1 = pa[4 + t];
      dmma(d, -p0, p0);
      dmma(d2, -p1, p1);
      *reinterpret_cast<double2*>(Dinv + 64 * (J + 1) + g * 8 + 2 * t) = b2_diag_factor(d[0] + d2[0], d[1] + d2[1], fail);
    } else {
      // trailing tiles in registers; the ones that are due go back to shared memory
#pragma unroll
      for (int s = 0; s < TPW; ++s) {
        if (due[s] < J) continue;
        const double* pa = As + (8 * ti[s] + g) * LD + 8 * J;
        const double* pb = As + (8 * tk[s] + g) * LD + 8 * J;
        dmma(c[s], -pa[t], pb[t]);
        dmma(c[s], -pa[4 + t], pb[4 + t]);
        if (due[s] == J)
          *reinterpret_cast<double2*>(As + (8 * ti[s] + g) * LD + 8 * tk[s] + 2 * t) = make_double2(c[s][0], c[s][1]);
      }
    }
    B2_T(3);
    __syncthreads();
    B2_T(4);
  }
}

// accumulator layout (lane (g, t): X[g][2t], X[g][2t+1]) -> the two B-operand fragments of the same 8x8 tile
// (k-slice 0: X[t][g], k-slice 1: X[4+t][g]), with shuffles instead of a round trip through shared memory
__device__ __forceinline__ void b2_acc_to_b(double c0, double c1, int g, int t, double& b0, double& b1) {
  const int s0 = 4 * t + (g >> 1), s1 = 4 * (4 + t) + (g >> 1);
  const double l0 = __shfl_sync(0xffffffffu, c0, s0), h0 = __shfl_sync(0xffffffffu, c1, s0);
  const double l1 = __shfl_sync(0xffffffffu, c0, s1), h1 = __shfl_sync(0xffffffffu, c1, s1);
  b0 = (g & 1) ? h0 : l0;
  b1 = (g & 1) ? h1 : l1;
}

// ---- W (Mp x 8 nct, shared, leading dimension ldw) <- A^-1 W = L^-T L^-1 W, in place.  A warp owns CT column
// tiles (8 columns each) at a time and sweeps them forward and backward on its own (no block-wide barrier).
// The whole strip (NBK tiles per column tile) stays in REGISTERS as accumulator fragments; the loop over the
// block steps J is rolled, the loop over the strip's tiles is unrolled and predicated (I > J / I < J), so that
// every tile has a compile-time register and the code stays small (the fully unrolled first version — 2.4 k
// instructions per sweep and column-tile count, executed once — was bound by instruction fetch).  All the
// factor-fragment loads of a step are independent of its arithmetic; the dependent chain of a step is
//     X_J = W_J x_J (2 DMMAs) -> shuffle to operand layout -> update of the NEXT tile (2 DMMAs) -> shuffle,
// the updates of the other tiles fill the FP64 pipe behind it.  (The second version kept the strip in shared
// memory and re-read every tile around each update: ~870 cycles per block step, 19 k cycles per sweep pair,
// and 48 us for the 23 column tiles of a level-0 block on one SM against 13 us of FP64 pipe time.) ----
template <int NBK, int CT, int UNR = 1>
__device__ __forceinline__ void b2_solve_tiles(const double* As, const double* Dinv, double* Ws, int ldw, int nct) {
  constexpr int LD = 8 * NBK + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int ct0 = warp * CT; ct0 < nct; ct0 += kB2Warps * CT) {
    // a slot beyond the last tile computes on the last tile (the code stays branch-free) but never writes
    double* col[CT];
    bool on[CT];
    double x[CT][NBK][2];
    double n0[CT], n1[CT];  // the tile the next step finishes, picked out of the strip (runtime index)
#pragma unroll
    for (int q = 0; q < CT; ++q) {
      on[q] = ct0 + q < nct;
      col[q] = Ws + 8 * min(ct0 + q, nct - 1);
#pragma unroll
      for (int I = 0; I < NBK; ++I) {
        const double2 v = *reinterpret_cast<const double2*>(col[q] + (8 * I + g) * ldw + 2 * t);
        x[q][I][0] = v.x; x[q][I][1] = v.y;
      }
      n0[q] = x[q][0][0]; n1[q] = x[q][0][1];
    }
    // ---- forward: L y = w ----
    // UNR = NBK: the block steps are unrolled too — straight-line code without the tile selects (2 NBK per column
    // tile and step, ~70 % of the rolled loop's instructions).  Used where the sweeps are throughput-bound (dense
    // lower levels: every warp has tiles); ~3 k instructions for CT = 3, executed once.
#pragma unroll UNR
    for (int J = 0; J < NBK; ++J) {
      const double d0 = Dinv[64 * J + g * 8 + t], d1 = Dinv[64 * J + g * 8 + 4 + t];
      double xb0[CT], xb1[CT];
#pragma unroll
      for (int q = 0; q < CT; ++q) {
        double b0, b1;
        b2_acc_to_b(n0[q], n1[q], g, t, b0, b1);
        double y[2] = {0.0, 0.0}, y2[2] = {0.0, 0.0};  // independent DMMAs + add: see b2_cholesky
        dmma(y, d0, b0);
        dmma(y2, d1, b1);
        y[0] += y2[0]; y[1] += y2[1];
#pragma unroll
        for (int I = 0; I < NBK; ++I)
          if (I == J) { x[q][I][0] = y[0]; x[q][I][1] = y[1]; }
        b2_acc_to_b(y[0], y[1], g, t, xb0[q], xb1[q]);
      }
#pragma unroll
      for (int I = 1; I < NBK; ++I) {  // ascending: the next step's tile (J + 1) first
        if (I > J) {
          const double* la = As + (8 * I + g) * LD + 8 * J;
          const double a0 = -la[t], a1 = -la[4 + t];
#pragma unroll
          for (int q = 0; q < CT; ++q) {
            double u[2] = {0.0, 0.0};
            dmma(x[q][I], a0, xb0[q]);
            dmma(u, a1, xb1[q]);
            x[q][I][0] += u[0]; x[q][I][1] += u[1];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < CT; ++q) {
#pragma unroll
        for (int I = 1; I < NBK; ++I)
          if (I == J + 1) { n0[q] = x[q][I][0]; n1[q] = x[q][I][1]; }
      }
    }
    // ---- backward: L^T x = y ----
#pragma unroll
    for (int q = 0; q < CT; ++q) { n0[q] = x[q][NBK - 1][0]; n1[q] = x[q][NBK - 1][1]; }
#pragma unroll UNR
    for (int J = NBK - 1; J >= 0; --J) {
      const double d0 = Dinv[64 * J + t * 8 + g], d1 = Dinv[64 * J + (4 + t) * 8 + g];  // W_J^T
      double xb0[CT], xb1[CT];
#pragma unroll
      for (int q = 0; q < CT; ++q) {
        double b0, b1;
        b2_acc_to_b(n0[q], n1[q], g, t, b0, b1);
        double y[2] = {0.0, 0.0}, y2[2] = {0.0, 0.0};
        dmma(y, d0, b0);
        dmma(y2, d1, b1);
        y[0] += y2[0]; y[1] += y2[1];
        if (on[q]) *reinterpret_cast<double2*>(col[q] + (8 * J + g) * ldw + 2 * t) = make_double2(y[0], y[1]);
        b2_acc_to_b(y[0], y[1], g, t, xb0[q], xb1[q]);
      }
#pragma unroll
      for (int I = NBK - 2; I >= 0; --I) {  // descending: the next step's tile (J - 1) first
        if (I < J) {
          const double* la = As + (8 * J) * LD + 8 * I + g;  // element (g, k) = L[8 J + k][8 I + g]
          const double a0 = -la[t * LD], a1 = -la[(4 + t) * LD];
#pragma unroll
          for (int q = 0; q < CT; ++q) {
            double u[2] = {0.0, 0.0};
            dmma(x[q][I], a0, xb0[q]);
            dmma(u, a1, xb1[q]);
            x[q][I][0] += u[0]; x[q][I][1] += u[1];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < CT; ++q) {
#pragma unroll
        for (int I = 0; I < NBK - 1; ++I)
          if (I == J - 1) { n0[q] = x[q][I][0]; n1[q] = x[q][I][1]; }
      }
    }
    __syncwarp();
  }
}

// ---- the same solve for FEW column tiles per CTA (the sparse upper levels: the column tiles of an odd block are
// spread over up to 23 CTAs): a GROUP of WPT warps shares one column tile, warp wi of the group owns the strip
// tiles wi, wi + WPT, ... in registers.  Per block step the owner of tile J multiplies it by W_J and publishes
// X_J in shared memory (double-buffered, one named barrier of the group per step); every warp then updates its
// own tiles.  With one warp per column tile a block step costs ~770 cycles of in-order issue (the warp runs ~160
// instructions, most of them bookkeeping for the 10 other tiles); here the dependent chain is
//     select + shuffle -> 2 DMMAs -> add -> store | barrier | load -> 2 DMMAs -> add        (~250 cycles). ----
template <int NBK, int WPT>
__device__ __forceinline__ void b2_solve_group(const double* As, const double* Dinv, double* Ws, int ldw, int nct,
                                               double* xbuf_all) {
  constexpr int LD = 8 * NBK + 4;
  constexpr int SL = (NBK + WPT - 1) / WPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int grp = warp / WPT, wi = warp % WPT;
  if (grp >= nct) return;  // no block-wide barrier below
  double* col = Ws + 8 * grp + g * ldw + 2 * t;   // + 8 I ldw: this lane's accumulator pair of strip tile I
  double* xbuf = xbuf_all + grp * 128;
  const int bar_id = 2 + grp;
  // Everything that depends on the step is kept as a running pointer / counter: the straightforward index
  // arithmetic (tile -> address, clamping of the inactive slots) compiled to ~180 instructions per block step,
  // and a lone warp issues them at ~3.6 cycles each — the arithmetic itself is 4 DMMAs.
  double x[SL][2];
  int Is[SL];
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    const int I = wi + s * WPT;
    Is[s] = I < NBK ? I : NBK + 64;  // a slot without a tile: never active in either direction (I > J false, see below)
    x[s][0] = 0.0; x[s][1] = 0.0;
    if (I < NBK) {
      const double2 v = *reinterpret_cast<const double2*>(col + 8 * I * ldw);
      x[s][0] = v.x; x[s][1] = v.y;
    }
  }
  const double* dfw = Dinv + g * 8 + t;   // W_J   fragments: + 64 J, + 4
  const double* dbw = Dinv + t * 8 + g;   // W_J^T fragments: + 64 J, + 32
  double* xst = xbuf + g * 8 + 2 * t;     // publish (accumulator layout), + 64 parity
  const double* xld = xbuf + t * 8 + g;   // read as B operand, + 32 for the second k-slice
  int par = 0;
  // ---- forward: L y = w.  la[s] -> L tile (I_s, J), A-operand fragment of this lane; slots without a tile read
  // tile (NBK - 1, J) and discard the product ----
  {
    const double* la[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) la[s] = As + (8 * min(wi + s * WPT, NBK - 1) + g) * LD + t;
    int own = 0, sj = 0;  // owner warp of tile J (= J % WPT) and its slot (= J / WPT)
#pragma unroll 1
    for (int J = 0; J < NBK; ++J) {
      double a0[SL], a1[SL];
#pragma unroll
      for (int s = 0; s < SL; ++s) { a0[s] = la[s][0]; a1[s] = la[s][4]; la[s] += 8; }
      if (wi == own) {  // warp-uniform
        double c0 = x[0][0], c1 = x[0][1];
#pragma unroll
        for (int s = 1; s < SL; ++s)
          if (s == sj) { c0 = x[s][0]; c1 = x[s][1]; }
        double b0, b1;
        b2_acc_to_b(c0, c1, g, t, b0, b1);
        double y[2] = {0.0, 0.0}, y2[2] = {0.0, 0.0};
        dmma(y, dfw[0], b0);
        dmma(y2, dfw[4], b1);
        y[0] += y2[0]; y[1] += y2[1];
        *reinterpret_cast<double2*>(xst + par) = make_double2(y[0], y[1]);
#pragma unroll
        for (int s = 0; s < SL; ++s)
          if (s == sj) { x[s][0] = y[0]; x[s][1] = y[1]; }
        }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(WPT * 32) : "memory");
      const double xb0 = xld[par], xb1 = xld[par + 32];
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        if (Is[s] > J && Is[s] < NBK) {  // warp-uniform; an idle slot's DMMAs would still occupy the FP64 pipe
          double u[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
          dmma(u, a0[s], xb0);
          dmma(v, a1[s], xb1);
          x[s][0] -= u[0] + v[0]; x[s][1] -= u[1] + v[1];
        }
      }
      dfw += 64;
      par ^= 64;
      if (++own == WPT) { own = 0; ++sj; }
    }
  }
  // ---- backward: L^T x = y.  la[s] -> element (g, k) = L[8 J + k][8 I_s + g] ----
  {
    const double* la[SL];
#pragma unroll
    for (int s = 0; s < SL; ++s) la[s] = As + (8 * (NBK - 1) + t) * LD + 8 * min(wi + s * WPT, NBK - 1) + g;
    int own = (NBK - 1) % WPT, sj = (NBK - 1) / WPT;
    dbw += 64 * (NBK - 1);
    double* cst = col + 8 * (NBK - 1) * ldw;
#pragma unroll 1
    for (int J = NBK - 1; J >= 0; --J) {
      double a0[SL], a1[SL];
#pragma unroll
      for (int s = 0; s < SL; ++s) { a0[s] = la[s][0]; a1[s] = la[s][4 * LD]; la[s] -= 8 * LD; }
      if (wi == own) {
        double c0 = x[0][0], c1 = x[0][1];
#pragma unroll
        for (int s = 1; s < SL; ++s)
          if (s == sj) { c0 = x[s][0]; c1 = x[s][1]; }
        double b0, b1;
        b2_acc_to_b(c0, c1, g, t, b0, b1);
        double y[2] = {0.0, 0.0}, y2[2] = {0.0, 0.0};
        dmma(y, dbw[0], b0);
        dmma(y2, dbw[32], b1);
        y[0] += y2[0]; y[1] += y2[1];
        *reinterpret_cast<double2*>(xst + par) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2*>(cst) = make_double2(y[0], y[1]);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(WPT * 32) : "memory");
      const double xb0 = xld[par], xb1 = xld[par + 32];
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        if (Is[s] < J) {
          double u[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
          dmma(u, a0[s], xb0);
          dmma(v, a1[s], xb1);
          x[s][0] -= u[0] + v[0]; x[s][1] -= u[1] + v[1];
        }
      }
      dbw -= 64;
      cst -= 8 * ldw;
      par ^= 64;
      if (--own < 0) { own = WPT - 1; --sj; }
    }
  }
}

template <int NBK>
struct B2Cfg {
  static constexpr int Mp = 8 * NBK;
  static constexpr int LD = Mp + 4;
  static constexpr int NCT = 2 * NBK + 1;             // column tiles of [B_{p-1} | B_p^T | b_p 0..]
  static constexpr int CT = 3;        // column tiles a warp sweeps in lockstep (independent DMMA chains)
};

// ---- one level: eliminate the odd super blocks.  grid (n / 2, R): CTA (q, y) factors A_p (p = 2q + 1;
// every CTA of a row does so redundantly — there is nothing else for those SMs to do) and solves for
// the column tiles [y nct_cta, (y + 1) nct_cta) of the NCT = 2 NBK + 1. ----
template <int NBK>
__global__ void __launch_bounds__(kB2Threads, 1) k_b2_fs(B2Level lv, int nct_cta, int* __restrict__ fail) {
  using Cfg = B2Cfg<NBK>;
  constexpr int Mp = Cfg::Mp, LD = Cfg::LD, NCT = Cfg::NCT;
  extern __shared__ __align__(16) double b2_sm[];
  const int ldw = 8 * nct_cta + 4;
  double* As = b2_sm;                // [Mp][LD]
  double* Dinv = As + Mp * LD;       // [NBK][64]
  double* Ws = Dinv + NBK * 64;      // [Mp][ldw]
  double* xbuf = Ws + Mp * ldw;      // [8 groups][2][64]: b2_solve_group
  const int q = blockIdx.x, p = 2 * q + 1;
  const int ct_first = blockIdx.y * nct_cta;
  const int nct = min(nct_cta, NCT - ct_first);
  if (nct <= 0) return;
  const bool has_v = p + 1 < lv.n;
  const size_t MM = size_t(Mp) * Mp;
  b2_stage(As, LD, lv.A + p * MM, Mp);
  cp_commit();
  // right-hand sides of this CTA: global column gc of [B_{p-1} | B_p^T | b_p 0 0 0 0 0 0 0]
  {
    const double* Bl = lv.B + size_t(p - 1) * MM;
    const double* Br = lv.B + size_t(p) * MM;  // only dereferenced when has_v
    const double* bp = lv.b + size_t(p) * Mp;
    const int ncol = 8 * nct;
    // part 1: columns inside B_{p-1}: rows contiguous along the column index (16-byte copies)
    for (int i = threadIdx.x; i < Mp * (ncol >> 1); i += kB2Threads) {
      const int r = i / (ncol >> 1), c2 = i - r * (ncol >> 1);
      const int lc = 2 * c2, gc = 8 * ct_first + lc;
      double* dst = Ws + r * ldw + lc;
      if (gc < Mp) cp16(dst, Bl + size_t(r) * Mp + gc);
      else if (gc >= 2 * Mp || !has_v) {
        dst[0] = (gc == 2 * Mp) ? bp[r] : 0.0;
        dst[1] = 0.0;
      }
    }
    // part 2: columns inside B_p^T: W[r][Mp + j] = B_p[j][r]; consecutive threads take consecutive r
    // (conflict-free shared-memory writes; the strided global reads come from L2)
    if (has_v) {
      const int j0 = max(8 * ct_first, Mp) - Mp, j1 = min(8 * (ct_first + nct), 2 * Mp) - Mp;
      for (int i = threadIdx.x; i < (j1 - j0) * Mp; i += kB2Threads) {
        const int j = j0 + i / Mp, r = i % Mp;
        cp8(Ws + r * ldw + (Mp + j - 8 * ct_first), Br + size_t(j) * Mp + r);
      }
    }
    cp_commit();
  }
  cp_wait<1>();  // A has landed; the right-hand sides keep streaming in under the factorisation
  __syncthreads();
  b2_cholesky<NBK>(As, Dinv, fail);
  cp_wait<0>();
  __syncthreads();
  // one column tile per warp while there are warps to spare (a warp's DMMAs serialise on its scheduler's FP64
  // pipe: 264 per tile and sweep pair, 16 cycles each), three in lockstep otherwise
  if (nct == 1) b2_solve_group<NBK, 8>(As, Dinv, Ws, ldw, nct, xbuf);
  else if (nct == 2) b2_solve_group<NBK, 4>(As, Dinv, Ws, ldw, nct, xbuf);
  else if (nct <= 4) b2_solve_group<NBK, 2>(As, Dinv, Ws, ldw, nct, xbuf);
  else if (nct <= kB2Warps) b2_solve_tiles<NBK, 1, NBK>(As, Dinv, Ws, ldw, nct);
  else b2_solve_tiles<NBK, Cfg::CT, NBK>(As, Dinv, Ws, ldw, nct);
  __syncthreads();
  // solution columns -> Uh | Vh | yh (16-byte stores, rows contiguous)
  double* Uq = lv.Uh + size_t(q) * MM;
  double* Vq = lv.Vh + size_t(q) * MM;
  double* yq = lv.yh + size_t(q) * Mp;
  {
    const int ncol = 8 * nct, half = ncol >> 1;
    for (int i = threadIdx.x; i < Mp * half; i += kB2Threads) {
      const int r = i / half, lc = 2 * (i - r * half);
      const int gc = 8 * ct_first + lc;
      const double2 v = *reinterpret_cast<const double2*>(Ws + r * ldw + lc);
      if (gc < Mp) *reinterpret_cast<double2*>(Uq + size_t(r) * Mp + gc) = v;
      else if (gc < 2 * Mp) { if (has_v) *reinterpret_cast<double2*>(Vq + size_t(r) * Mp + (gc - Mp)) = v; }
      else if (gc == 2 * Mp) yq[r] = v.x;
    }
  }
}

// ---- the last block: x_0 = A^-1 b ----
template <int NBK>
__global__ void __launch_bounds__(kB2Threads, 1) k_b2_top(B2Level lv, double* __restrict__ x, int* __restrict__ fail) {
  using Cfg = B2Cfg<NBK>;
  constexpr int Mp = Cfg::Mp, LD = Cfg::LD;
  extern __shared__ __align__(16) double b2_sm[];
  constexpr int ldw = 12;
  double* As = b2_sm;
  double* Dinv = As + Mp * LD;
  double* Ws = Dinv + NBK * 64;  // [Mp][12]: column 0 = b
  double* xbuf = Ws + Mp * ldw;
  B2_T0();
  b2_stage(As, LD, lv.A, Mp);
  cp_commit();
  for (int i = threadIdx.x; i < Mp * 8; i += kB2Threads) Ws[(i >> 3) * ldw + (i & 7)] = (i & 7) == 0 ? lv.b[i >> 3] : 0.0;
  cp_wait<0>();
  __syncthreads();
  B2_T(8);
  b2_cholesky<NBK>(As, Dinv, fail);
  B2_T(9);
  b2_solve_group<NBK, 8>(As, Dinv, Ws, ldw, 1, xbuf);
  __syncthreads();
  B2_T(10);
  for (int i = threadIdx.x; i < Mp; i += kB2Threads) x[i] = Ws[i * ldw];
}

// ---- rows [8 I, 8 I + 8) of b' = b_e - B_e^T yh_r - B_{e-1} yh_l, by ONE warp, as a DMMA tile whose B operand is the
// vector in column 0 (fragments straight from L2, all loads of a product in flight).  The first version gave a warp
// 6-11 whole rows, one after the other, each a strided dot product with its own load latency: 7-14 us, the longest
// role of the update kernels at the sparse levels. ----
template <int NBK>
__device__ __forceinline__ void b2_rhs_tile(const B2Level& lv, const B2Level& nx, int pe, int I) {
  constexpr int Mp = 8 * NBK;
  const int e = 2 * pe;
  const bool has_l = e - 1 >= 0, has_r = e + 1 < lv.n;
  const size_t MM = size_t(Mp) * Mp;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int qr = e / 2, ql = e / 2 - 1;
  double c[2] = {t == 0 ? lv.b[size_t(e) * Mp + 8 * I + g] : 0.0, 0.0};
  double c2[2] = {0.0, 0.0};
  if (has_r) {
    const double* xa = lv.B + size_t(e) * MM + size_t(t) * Mp + 8 * I + g;  // (B_e^T)[8 I + g][k] = B_e[k][8 I + g]
    const double* yv = lv.yh + size_t(qr) * Mp + t;
    double a[Mp / 4], b[Mp / 4];
#pragma unroll
    for (int j = 0; j < Mp / 4; ++j) { a[j] = __ldg(xa + size_t(4 * j) * Mp); b[j] = g == 0 ? __ldg(yv + 4 * j) : 0.0; }
#pragma unroll
    for (int j = 0; j < Mp / 4; j += 2) {
      dmma(c, -a[j], b[j]);
      if (j + 1 < Mp / 4) dmma(c2, -a[j + 1], b[j + 1]);
    }
  }
  if (has_l) {
    const double* xa = lv.B + size_t(e - 1) * MM + size_t(8 * I + g) * Mp + t;  // B_{e-1}[8 I + g][k]
    const double* yv = lv.yh + size_t(ql) * Mp + t;
    double a[Mp / 4], b[Mp / 4];
#pragma unroll
    for (int j = 0; j < Mp / 4; ++j) { a[j] = __ldg(xa + 4 * j); b[j] = g == 0 ? __ldg(yv + 4 * j) : 0.0; }
#pragma unroll
    for (int j = 0; j < Mp / 4; j += 2) {
      dmma(c, -a[j], b[j]);
      if (j + 1 < Mp / 4) dmma(c2, -a[j + 1], b[j + 1]);
    }
  }
  if (t == 0) nx.b[size_t(pe) * Mp + 8 * I + g] = c[0] + c2[0];
}

// ---- even super blocks of a level -> next level.  grid (n_next, 3):
//   y == 0:  A' = A_e - B_e^T Uh_r - B_{e-1} Vh_l    (lower tiles; r / l = the odd neighbours' solutions)
//   y == 1:  B' = -B_{e+1} Uh_r                       (coupling across the eliminated block e + 1)
//   y == 2:  b' = b_e - B_e^T yh_r - B_{e-1} yh_l
// Operands are staged in shared memory two matrices at a time; a warp runs up to TW tiles in lockstep
// (independent accumulator chains), 2 NBK DMMAs per tile and operand pair. ----
template <int NBK>
__global__ void __launch_bounds__(kB2RedThreads, 1) k_b2_reduce(B2Level lv, B2Level nx, int PA, int PB) {
  constexpr int Mp = 8 * NBK, LD = Mp + 4;
  constexpr int NTL = NBK * (NBK + 1) / 2, NTF = NBK * NBK;
  constexpr int TWL = (NTL + kB2RedWarps - 1) / kB2RedWarps, TWF = (NTF + kB2RedWarps - 1) / kB2RedWarps;
  extern __shared__ __align__(16) double b2_sm[];
  double* Xs = b2_sm;            // [Mp][LD]
  double* Ys = Xs + Mp * LD;     // [Mp][LD]
  const int pe = blockIdx.x, e = 2 * pe;
  // blockIdx.y: [0, PA) = parts of A' (its lower tiles split into PA contiguous ranges), [PA, PA + PB) = parts of
  // B', PA + PB = b'.  At the sparse upper levels the products are FP64-bound on one SM (a million FMAs at 64 per
  // clock), so they are spread over many CTAs; every part stages the full operand matrices from L2.
  const int role = int(blockIdx.y) < PA ? 0 : (int(blockIdx.y) < PA + PB ? 1 : 2);
  const int part = role == 0 ? blockIdx.y : int(blockIdx.y) - PA;
  const bool has_l = e - 1 >= 0, has_r = e + 1 < lv.n;
  const size_t MM = size_t(Mp) * Mp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int qr = e / 2, ql = e / 2 - 1;  // odd-block indices of e + 1 and e - 1
  if (role == 2) {
    if (warp < NBK) b2_rhs_tile<NBK>(lv, nx, pe, warp);
    return;
  }
  if (role == 1) {
    if (e + 2 >= lv.n) return;  // no block beyond e + 1: no coupling at the next level
    b2_stage(Xs, LD, lv.B + size_t(e + 1) * MM, Mp);   // B_{e+1}: A operand, row-major
    b2_stage(Ys, LD, lv.Uh + size_t(qr) * MM, Mp);     // Uh_r:   B operand, k-major
    cp_commit();
    cp_wait<0>();
    __syncthreads();
    double c[TWF][2];
    int tI[TWF], tK[TWF];
    const int per = (NTF + PB - 1) / PB, u_lo = part * per, u_hi = min(NTF, u_lo + per);
#pragma unroll
    for (int s = 0; s < TWF; ++s) {
      const int u = u_lo + warp + s * kB2RedWarps;
      tI[s] = u < u_hi ? u / NBK : -1;
      tK[s] = u < u_hi ? u % NBK : 0;
      c[s][0] = c[s][1] = 0.0;
    }
#pragma unroll 2
    for (int kk = 0; kk < Mp; kk += 4) {
#pragma unroll
      for (int s = 0; s < TWF; ++s) {
        if (tI[s] < 0) continue;
        const double a = -Xs[(8 * tI[s] + g) * LD + kk + t];
        const double b = Ys[(kk + t) * LD + 8 * tK[s] + g];
        dmma(c[s], a, b);
      }
    }
    double* Bn = nx.B + size_t(pe) * MM;
#pragma unroll
    for (int s = 0; s < TWF; ++s)
      if (tI[s] >= 0)
        *reinterpret_cast<double2*>(Bn + size_t(8 * tI[s] + g) * Mp + 8 * tK[s] + 2 * t) = make_double2(c[s][0], c[s][1]);
    return;
  }
  // role 0: A'
  double c[TWL][2];
  int tI[TWL], tK[TWL];
  const double* Ae = lv.A + size_t(e) * MM;
  const int per = (NTL + PA - 1) / PA, u_lo = part * per, u_hi = min(NTL, u_lo + per);
#pragma unroll
  for (int s = 0; s < TWL; ++s) {
    const int u = u_lo + warp + s * kB2RedWarps;
    int I = -1, K = 0;
    if (u < u_hi) {
      I = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
      while (I * (I + 1) / 2 > u) --I;
      while ((I + 1) * (I + 2) / 2 <= u) ++I;
      K = u - I * (I + 1) / 2;
    }
    tI[s] = I; tK[s] = K;
    c[s][0] = c[s][1] = 0.0;
    if (I >= 0) {
      const double2 v = *reinterpret_cast<const double2*>(Ae + size_t(8 * I + g) * Mp + 8 * K + 2 * t);
      c[s][0] = v.x; c[s][1] = v.y;
    }
  }
  if (has_r) {
    b2_stage(Xs, LD, lv.B + size_t(e) * MM, Mp);    // B_e:  A operand TRANSPOSED (k-major)
    b2_stage(Ys, LD, lv.Uh + size_t(qr) * MM, Mp);  // Uh_r: B operand, k-major
    cp_commit();
    cp_wait<0>();
    __syncthreads();
#pragma unroll 2
    for (int kk = 0; kk < Mp; kk += 4) {
#pragma unroll
      for (int s = 0; s < TWL; ++s) {
        if (tI[s] < 0) continue;
        const double a = -Xs[(kk + t) * LD + 8 * tI[s] + g];
        const double b = Ys[(kk + t) * LD + 8 * tK[s] + g];
        dmma(c[s], a, b);
      }
    }
    __syncthreads();
  }
  if (has_l) {
    b2_stage(Xs, LD, lv.B + size_t(e - 1) * MM, Mp);  // B_{e-1}: A operand, row-major
    b2_stage(Ys, LD, lv.Vh + size_t(ql) * MM, Mp);    // Vh_l:    B operand, k-major
    cp_commit();
    cp_wait<0>();
    __syncthreads();
#pragma unroll 2
    for (int kk = 0; kk < Mp; kk += 4) {
#pragma unroll
      for (int s = 0; s < TWL; ++s) {
        if (tI[s] < 0) continue;
        const double a = -Xs[(8 * tI[s] + g) * LD + kk + t];
        const double b = Ys[(kk + t) * LD + 8 * tK[s] + g];
        dmma(c[s], a, b);
      }
    }
  }
  double* An = nx.A + size_t(pe) * MM;
#pragma unroll
  for (int s = 0; s < TWL; ++s)
    if (tI[s] >= 0)
      *reinterpret_cast<double2*>(An + size_t(8 * tI[s] + g) * Mp + 8 * tK[s] + 2 * t) = make_double2(c[s][0], c[s][1]);
}

// ---- the same update for the SPARSE upper levels (few even blocks): no shared-memory staging.  The staged kernel
// above spends its time bringing two full operand matrices in (twice, for the right and the left neighbour) before
// a few warps multiply for a microsecond: 12-14 us per level whatever the level's size.  Here a warp owns ONE
// output tile and loads its DMMA fragments straight from L2 (a fragment is 8 full 32-byte sectors), all 2 x 22 of
// a product in flight at once, then runs two independent accumulator chains.  Without the shared-memory reuse
// a level moves 187 tiles x 2 products x 11 KB per even block through L2, which is why the dense lower levels
// keep the staged kernel.  grid (n_next, ceil((66 + 121 + 11) / 8)) for 88 x 88 blocks. ----
template <int NBK>
__global__ void __launch_bounds__(kB2Threads, 1) k_b2_reduce_direct(B2Level lv, B2Level nx) {
  constexpr int Mp = 8 * NBK;
  constexpr int NTL = NBK * (NBK + 1) / 2, NTF = NBK * NBK;
  const int pe = blockIdx.x, e = 2 * pe;
  const bool has_l = e - 1 >= 0, has_r = e + 1 < lv.n;
  const size_t MM = size_t(Mp) * Mp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int qr = e / 2, ql = e / 2 - 1;  // odd-block indices of e + 1 and e - 1
  // [0, NTL): lower tiles of A';  [NTL, NTL + NTF): tiles of B';  then NBK row tiles of b'
  const int u = int(blockIdx.y) * kB2Warps + warp;
  if (u >= NTL + NTF + NBK) return;
  if (u >= NTL + NTF) { b2_rhs_tile<NBK>(lv, nx, pe, u - NTL - NTF); return; }
  // one product: c -= X(I-rows) * Y(K-cols) with the fragment addressing of the caller
  auto product = [&](double (&c)[2], const double* xa, size_t xa_k, const double* yb) {
    // xa + kk * xa_k: this lane's A-operand element of k-slice kk / 4 ; yb + kk * Mp: B-operand element
    double a[Mp / 4], b[Mp / 4];
#pragma unroll
    for (int j = 0; j < Mp / 4; ++j) { a[j] = __ldg(xa + size_t(4 * j) * xa_k); b[j] = __ldg(yb + size_t(4 * j) * Mp); }
    double c2[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < Mp / 4; j += 2) {
      dmma(c, -a[j], b[j]);
      if (j + 1 < Mp / 4) dmma(c2, -a[j + 1], b[j + 1]);
    }
    c[0] += c2[0]; c[1] += c2[1];
  };
  if (u < NTL) {
    int I = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
    while (I * (I + 1) / 2 > u) --I;
    while ((I + 1) * (I + 2) / 2 <= u) ++I;
    const int K = u - I * (I + 1) / 2;
    const double2 v = *reinterpret_cast<const double2*>(lv.A + size_t(e) * MM + size_t(8 * I + g) * Mp + 8 * K + 2 * t);
    double c[2] = {v.x, v.y};
    // B_e^T Uh_r: A-operand element (row g of tile I, k) = B_e[k][8 I + g]
    if (has_r) product(c, lv.B + size_t(e) * MM + size_t(t) * Mp + 8 * I + g, Mp, lv.Uh + size_t(qr) * MM + size_t(t) * Mp + 8 * K + g);
    // B_{e-1} Vh_l: A-operand element = B_{e-1}[8 I + g][k]
    if (has_l) product(c, lv.B + size_t(e - 1) * MM + size_t(8 * I + g) * Mp + t, 1, lv.Vh + size_t(ql) * MM + size_t(t) * Mp + 8 * K + g);
    *reinterpret_cast<double2*>(nx.A + size_t(pe) * MM + size_t(8 * I + g) * Mp + 8 * K + 2 * t) = make_double2(c[0], c[1]);
  } else {
    if (e + 2 >= lv.n) return;  // no block beyond e + 1: no coupling at the next level
    const int w = u - NTL, I = w / NBK, K = w % NBK;
    double c[2] = {0.0, 0.0};
    // B' = -B_{e+1} Uh_r
    product(c, lv.B + size_t(e + 1) * MM + size_t(8 * I + g) * Mp + t, 1, lv.Uh + size_t(qr) * MM + size_t(t) * Mp + 8 * K + g);
    *reinterpret_cast<double2*>(nx.B + size_t(pe) * MM + size_t(8 * I + g) * Mp + 8 * K + 2 * t) = make_double2(c[0], c[1]);
  }
}

// ---- back-substitution: x_p = yh_p - Uh_p x_{p-1} - Vh_p x_{p+1} for the odd blocks of every level, top down,
// in ONE cooperative kernel (a grid-wide barrier between levels instead of a launch per level; the first
// version's single-CTA kernel for the upper levels took 185 us, a launch per level 21 us each).  x is indexed
// by ORIGINAL super block (p << level).  A warp per row of a block, lanes along the row (coalesced), up to
// four rows of a warp in flight. ----
struct B2LevelPack { B2Level lv[24]; };

__global__ void __launch_bounds__(256) k_b2_back_all(int Mp, B2LevelPack pk, int n_levels, double* __restrict__ x) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const size_t MM = size_t(Mp) * Mp;
  for (int l = n_levels - 2; l >= 0; --l) {
    const B2Level& lv = pk.lv[l];
    const int rows = (lv.n / 2) * Mp;
    for (int r0 = gw; r0 < rows; r0 += 4 * nw) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = r0 + j * nw;
        if (row >= rows) continue;
        const int q = row / Mp, r = row - q * Mp, p = 2 * q + 1;
        const bool has_r = p + 1 < lv.n;
        const double* U = lv.Uh + q * MM + size_t(r) * Mp;
        const double* V = lv.Vh + q * MM + size_t(r) * Mp;
        const double* xl = x + (size_t(p - 1) << l) * Mp;
        const double* xr = x + (size_t(p + 1) << l) * Mp;
        for (int cc = lane; cc < Mp; cc += 32) {
          s[j] += U[cc] * xl[cc];
          if (has_r) s[j] += V[cc] * xr[cc];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        const int row = r0 + j * nw;
        if (lane == 0 && row < rows) {
          const int q = row / Mp, r = row - q * Mp, p = 2 * q + 1;
          x[(size_t(p) << l) * Mp + r] = lv.yh[size_t(q) * Mp + r] - s[j];
        }
      }
    }
    if (l > 0) grid.sync();
  }
}

size_t b2_fs_smem(int nbk, int nct_cta) {
  const int Mp = 8 * nbk;
  return (size_t(Mp) * b2_ld(Mp) + size_t(nbk) * 64 + size_t(Mp) * (8 * nct_cta + 4) + 8 * 128) * sizeof(double);
}

constexpr size_t kB2SmemCap = 225 * 1024;

template <int NBK>
pba_status b2_launch_level(Handle* h, const B2Level& lv, const B2Level& nx, int R_min, int n_sm) {
  // Column tiles of the triangular solves over R CTAs per odd block (each factors A_p itself): the solves are
  // FP64-bound on one SM (1.4 M FMAs at 64 per clock = 11 us for 88 x 88), so the SMs a sparse level leaves idle
  // take a share; R grows until the level fills the GPU once.
  const int NCT = 2 * NBK + 1, n_odd = lv.n / 2;
  int R = std::max(R_min, n_sm / std::max(1, n_odd));
  R = std::min(R, NCT);
  const int nct_cta = (NCT + R - 1) / R;
  R = (NCT + nct_cta - 1) / nct_cta;
  PBA_LAUNCH(h, K_BCR, k_b2_fs<NBK>, dim3(n_odd, R), dim3(kB2Threads), b2_fs_smem(NBK, nct_cta), lv, nct_cta,
             h->chol_fail.p);
  // products: parts per next-level block chosen so that a level is about one wave of CTAs
  int PA = 1, PB = 1;
  if (nx.n * 24 <= n_sm) { PA = 9; PB = 14; }
  else if (nx.n * 11 <= n_sm) { PA = 4; PB = 6; }
  else if (nx.n * 6 <= n_sm) { PA = 2; PB = 3; }
  static const bool no_direct = getenv("PBA_B2_NO_DIRECT") != nullptr;
  if (nx.n <= 12 && !no_direct) {
    constexpr int NCTA = (NBK * (NBK + 1) / 2 + NBK * NBK + NBK + kB2Warps - 1) / kB2Warps;
    PBA_LAUNCH(h, K_BCR, k_b2_reduce_direct<NBK>, dim3(nx.n, NCTA), dim3(kB2Threads), 0, lv, nx);
    return PBA_OK;
  }
  const size_t smem_r = size_t(2) * (8 * NBK) * b2_ld(8 * NBK) * sizeof(double);
  PBA_LAUNCH(h, K_BCR, k_b2_reduce<NBK>, dim3(nx.n, PA + PB + 1), dim3(kB2RedThreads), smem_r, lv, nx, PA, PB);
  return PBA_OK;
}
template <int NBK>
pba_status b2_launch_top(Handle* h, const B2Level& lv, double* x) {
  const size_t smem = (size_t(8 * NBK) * b2_ld(8 * NBK) + size_t(NBK) * 64 + size_t(8 * NBK) * 12 + 8 * 128) * sizeof(double);
  PBA_LAUNCH(h, K_BCR, k_b2_top<NBK>, dim3(1), dim3(kB2Threads), smem, lv, x, h->chol_fail.p);
  return PBA_OK;
}

}  // namespace

// Workspace of the second-generation solver (levels carved out of one buffer).  Called from bcr_setup.
pba_status bcr2_setup(Handle* h) {
  const Sizes& z = h->sz;
  h->b2_nbk = 0;
  if (!h->bcr_m) return PBA_OK;
  const int m = h->bcr_m, M = m * z.cd, Mp = (M + 7) / 8 * 8, nbk = Mp / 8;
  if (nbk > kB2MaxNbk) return PBA_OK;
  const int S = (z.n_slots + m - 1) / m;
  const size_t MM = size_t(Mp) * Mp;
  size_t total = 0;
  h->b2_off.clear();
  h->b2_n.clear();
  for (int n = S;; n = (n + 1) / 2) {
    h->b2_n.push_back(n);
    const int no = n / 2;
    const size_t sizes[6] = {n * MM, size_t(n > 1 ? n - 1 : 0) * MM, size_t(n) * Mp, no * MM, no * MM, size_t(no) * Mp};
    for (size_t sz : sizes) { h->b2_off.push_back(total); total += sz; }
    if (n == 1) break;
  }
  if (h->b2_n.size() > 24) return PBA_OK;
  h->b2_x_off = total;
  PBA_CUDA_OK(h->b2_ws.alloc(total + size_t(S) * Mp));
  // level 0 is rebuilt from the RCS blocks before every solve; the block pattern is static, so everything
  // k_b2_build never writes is zeroed here once (A and B of level 0 are contiguous)
  PBA_CUDA_OK(cudaMemsetAsync(h->b2_ws.p + h->b2_off[0], 0, sizeof(double) * (size_t(S) * MM + size_t(S > 1 ? S - 1 : 0) * MM), h->stream));
  h->b2_nbk = nbk;
  return PBA_OK;
}

pba_status launch_bcr2_rcs(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim == 0) return PBA_OK;
  const int m = h->bcr_m, M = m * z.cd, nbk = h->b2_nbk, Mp = 8 * nbk;
  const int S = h->b2_n[0];
  double* ws = h->b2_ws.p;
  const int nl = int(h->b2_n.size());
  B2LevelPack pk;
  for (int l = 0; l < nl; ++l) {
    const size_t* o = &h->b2_off[size_t(l) * 6];
    pk.lv[l] = B2Level{h->b2_n[l], ws + o[0], ws + o[1], ws + o[2], ws + o[3], ws + o[4], ws + o[5]};
  }
  const double* Sblk = h->rcs.p;
  const double* rhs = Sblk + z.n_blocks * z.cd * z.cd;
  double* x = ws + h->b2_x_off;
  PBA_CUDA_OK(cudaMemsetAsync(h->chol_fail.p, 0, sizeof(int), h->stream));
  PBA_LAUNCH(h, K_BCR, k_b2_build, dim3((unsigned)((z.n_blocks + S + 3) / 4)), dim3(256), 0, z.cd, m, M, Mp, z.n_blocks, z.dim,
             h->d_blk_row.p, h->d_blk_col.p, Sblk, rhs, S, pk.lv[0].A, pk.lv[0].B, pk.lv[0].b);
  // fewest CTAs per odd block of the factor + solve kernel whose share of the right-hand sides fits shared memory
  const int NCT = 2 * nbk + 1;
  int R = 1;
  while (b2_fs_smem(nbk, (NCT + R - 1) / R) > kB2SmemCap) ++R;
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
  pba_status st = PBA_OK;
  for (int l = 0; l + 1 < nl; ++l) {
#define B2_LEVEL_CALL(N) b2_launch_level<N>(h, pk.lv[l], pk.lv[l + 1], R, n_sm)
    switch (nbk) {
      case 1: st = B2_LEVEL_CALL(1); break;   case 2: st = B2_LEVEL_CALL(2); break;
      case 3: st = B2_LEVEL_CALL(3); break;   case 4: st = B2_LEVEL_CALL(4); break;
      case 5: st = B2_LEVEL_CALL(5); break;   case 6: st = B2_LEVEL_CALL(6); break;
      case 7: st = B2_LEVEL_CALL(7); break;   case 8: st = B2_LEVEL_CALL(8); break;
      case 9: st = B2_LEVEL_CALL(9); break;   case 10: st = B2_LEVEL_CALL(10); break;
      case 11: st = B2_LEVEL_CALL(11); break; case 12: st = B2_LEVEL_CALL(12); break;
      case 13: st = B2_LEVEL_CALL(13); break; case 14: st = B2_LEVEL_CALL(14); break;

      default: st = PBA_ERR_UNSUPPORTED;
    }
#undef B2_LEVEL_CALL
    if (st != PBA_OK) return st;
  }
#define B2_TOP_CALL(N) b2_launch_top<N>(h, pk.lv[nl - 1], x)
  switch (nbk) {
    case 1: st = B2_TOP_CALL(1); break;   case 2: st = B2_TOP_CALL(2); break;
    case 3: st = B2_TOP_CALL(3); break;   case 4: st = B2_TOP_CALL(4); break;
    case 5: st = B2_TOP_CALL(5); break;   case 6: st = B2_TOP_CALL(6); break;
    case 7: st = B2_TOP_CALL(7); break;   case 8: st = B2_TOP_CALL(8); break;
    case 9: st = B2_TOP_CALL(9); break;   case 10: st = B2_TOP_CALL(10); break;
    case 11: st = B2_TOP_CALL(11); break; case 12: st = B2_TOP_CALL(12); break;
    case 13: st = B2_TOP_CALL(13); break; case 14: st = B2_TOP_CALL(14); break;

    default: st = PBA_ERR_UNSUPPORTED;
  }
#undef B2_TOP_CALL
  if (st != PBA_OK) return st;
  // back-substitution: every level in one cooperative launch (one CTA per SM)
  if (nl >= 2) {
    int n_lev = nl;
    void* args[] = {(void*)&Mp, (void*)&pk, (void*)&n_lev, (void*)&x};
    h->stats.begin(K_BCR, h->stream);
    const cudaError_t e = cudaLaunchCooperativeKernel((void*)k_b2_back_all, dim3(n_sm), dim3(256), args, 0, h->stream);
    h->stats.end(h->stream);
    if (e != cudaSuccess) return map_cuda(e);
  }
  PBA_LAUNCH(h, K_BCR, k_b2_unpad, dim3((z.dim + 255) / 256), dim3(256), 0, M, Mp, z.dim, x, h->y_cam.p);
#ifdef PBA_B2_TIMING
  {
    static int calls = 0;
    if (++calls == 6) {
      cudaStreamSynchronize(h->stream);
      long long t[32];
      cudaMemcpyFromSymbol(t, g_b2_t, sizeof(t));
      // phases 0..4 are accumulated by every k_b2_fs (CTA 0) and k_b2_top call: per Cholesky = / (calls * (levels))
      const double nchol = double(calls) * nl;
      fprintf(stderr, "[b2] per Cholesky (cycles): diag+stores %.0f | barrier %.0f | panel %.0f | barrier %.0f | trailing %.0f ;  "
              "top kernel (per call): stage %.0f chol %.0f solve %.0f\n",
              t[0] / nchol, t[1] / nchol, t[2] / nchol, t[3] / nchol, t[4] / nchol, t[8] / double(calls), t[9] / double(calls),
              t[10] / double(calls));
    }
  }
#endif
  return PBA_OK;
}

}  // namespace pba
