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

// ---- 8x8 diagonal block, factored IN REGISTERS by the warp that owns the tile.  The tile sits in the DMMA
// accumulator layout (lane (g, t) holds A[g][2t], A[g][2t+1]); the warp runs the square-root form of
// Gaussian elimination on [A | I] with row operations: step j scales row j by 1 / sqrt(a_jj) and subtracts
// multiples of it from the rows below, so that A becomes L^T and the identity becomes W = L^-1 — the only
// thing the panel and the triangular solves need.  Per pivot the dependent chain is shuffle -> rsqrt ->
// multiply -> FMA (~110 cycles), every lane issues ~20 instructions; the single-thread version of the first
// generation (factor, then invert) took ~2,100 cycles per block and was the longest item of a level.
// Returns W in (w0, w1) = W[g][2t], W[g][2t+1]. ----
//
// NOT inlined: the factorisation calls it from every slot of its unrolled tile loop, and these kernels run
// through their code once per launch — the first version (everything unrolled: 13,700 instructions = 219 KB
// for k_b2_fs<11>, more than the instruction cache) was bound by instruction fetch, 45 us a level whatever
// the amount of arithmetic.
__device__ __noinline__ double2 b2_diag_factor(double a0, double a1, int* fail) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double w0 = g == 2 * t ? 1.0 : 0.0;
  double w1 = g == 2 * t + 1 ? 1.0 : 0.0;
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
    const double inv = rsqrt(ajj);
    const double m = arj * inv;
    const double u0 = p0 * inv, u1 = p1 * inv, v0 = q0 * inv, v1 = q1 * inv;
    if (g > j) { a0 -= m * u0; a1 -= m * u1; w0 -= m * v0; w1 -= m * v1; }
    else if (g == j) { a0 = u0; a1 = u1; w0 = v0; w1 = v1; }
  }
  if (bad && lane == 0) *fail = 1;
  return make_double2(w0, w1);
}

// ---- A (Mp x Mp, shared, lower 8x8 tiles valid) <- the panels of its lower Cholesky factor (tiles (I, J),
// I > J); Dinv[J] = inverse of the J-th diagonal block of the factor (the diagonal tiles themselves are
// not stored: nothing reads them).  Right-looking, block size 8.  The whole matrix lives in registers as
// DMMA accumulator fragments (tile u = I (I + 1) / 2 + K is owned by warp u % 8 for the whole
// factorisation); per block column J: the owners write the column's sub-diagonal tiles to shared memory
// while the owner of (J, J) factors it in registers and publishes W_J; barrier; the warps scale the panel
// P_I = A_IJ W_J^T (one DMMA pair per tile); barrier; every owned trailing tile takes C -= P_I P_K^T (one
// DMMA pair, A and B fragments straight from the panel). ----
template <int NBK>
__device__ void b2_cholesky(double* As, double* Dinv, int* fail) {
  constexpr int LD = 8 * NBK + 4;
  constexpr int NT = NBK * (NBK + 1) / 2;
  constexpr int TPW = (NT + kB2Warps - 1) / kB2Warps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  int ti[TPW], tk[TPW];
  double c[TPW][2];
#pragma unroll
  for (int s = 0; s < TPW; ++s) {
    const int u = warp + s * kB2Warps;
    int I = -1, K = 0;
    if (u < NT) {
      I = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
      while (I * (I + 1) / 2 > u) --I;
      while ((I + 1) * (I + 2) / 2 <= u) ++I;
      K = u - I * (I + 1) / 2;
    }
    ti[s] = I; tk[s] = K;
    c[s][0] = 0.0; c[s][1] = 0.0;
    if (I >= 0 && (K >= 1 || I == 0)) {  // the sub-diagonal tiles of block column 0 are the first panel: they stay in shared memory
      const double2 v = *reinterpret_cast<const double2*>(As + (8 * I + g) * LD + 8 * K + 2 * t);
      c[s][0] = v.x; c[s][1] = v.y;
    }
  }
  B2_T0();
#pragma unroll 1
  for (int J = 0; J < NBK; ++J) {
#pragma unroll
    for (int s = 0; s < TPW; ++s) {
      if (ti[s] < 0 || tk[s] != J) continue;
      if (ti[s] == J) {  // warp-uniform: this warp owns the diagonal tile
        *reinterpret_cast<double2*>(Dinv + 64 * J + g * 8 + 2 * t) = b2_diag_factor(c[s][0], c[s][1], fail);
      } else if (J > 0) {
        *reinterpret_cast<double2*>(As + (8 * ti[s] + g) * LD + 8 * J + 2 * t) = make_double2(c[s][0], c[s][1]);
      }
    }
    B2_T(0);
    __syncthreads();
    B2_T(1);
    // panel: P_I = A_IJ W^T  (W = inverse of the diagonal factor)
    {
      const double w0 = Dinv[64 * J + g * 8 + t], w1 = Dinv[64 * J + g * 8 + 4 + t];
      for (int I = J + 1 + warp; I < NBK; I += kB2Warps) {
        double* tile = As + (8 * I + g) * LD + 8 * J;
        const double a0 = tile[t], a1 = tile[4 + t];
        double x[2] = {0.0, 0.0};
        dmma(x, a0, w0);
        dmma(x, a1, w1);
        __syncwarp();
        *reinterpret_cast<double2*>(tile + 2 * t) = make_double2(x[0], x[1]);
      }
    }
    B2_T(2);
    __syncthreads();
    B2_T(3);
    // trailing tiles in registers
#pragma unroll
    for (int s = 0; s < TPW; ++s) {
      if (ti[s] < 0 || tk[s] <= J) continue;
      const double* pa = As + (8 * ti[s] + g) * LD + 8 * J;
      const double* pb = As + (8 * tk[s] + g) * LD + 8 * J;
      dmma(c[s], -pa[t], pb[t]);
      dmma(c[s], -pa[4 + t], pb[4 + t]);
    }
    B2_T(4);
  }
  __syncthreads();
}

// ---- W (Mp x 8 nct, shared, leading dimension ldw) <- A^-1 W = L^-T L^-1 W, in place.  A warp owns CT column
// tiles (8 columns each) at a time and sweeps them forward and backward on its own: the only
// synchronisation is __syncwarp (no block-wide barrier).  Tiles stay in shared memory in the accumulator
// layout (one 16-byte load / store per lane and update) and the loops are NOT unrolled over the block rows:
// the straight-line version with the strip in registers was 2.4 k instructions per sweep and column-tile
// count, executed once — instruction fetch, not arithmetic, set its time. ----
template <int NBK, int CT>
__device__ __forceinline__ void b2_solve_tiles(const double* As, const double* Dinv, double* Ws, int ldw, int nct) {
  constexpr int LD = 8 * NBK + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int ct0 = warp * CT; ct0 < nct; ct0 += kB2Warps * CT) {
    // a slot beyond the last tile reads the last tile (so the code stays branch-free up to the stores) but never writes
    double* col[CT];
    bool on[CT];
#pragma unroll
    for (int q = 0; q < CT; ++q) { on[q] = ct0 + q < nct; col[q] = Ws + 8 * min(ct0 + q, nct - 1); }
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      // dir 0: forward, L y = w (J ascending, rows below J updated); dir 1: backward, L^T x = y
#pragma unroll 1
      for (int jj = 0; jj < NBK; ++jj) {
        const int J = dir == 0 ? jj : NBK - 1 - jj;
        // diagonal block: X_J = W_J tile (forward) or W_J^T tile (backward)
        const double d0 = dir == 0 ? Dinv[64 * J + g * 8 + t] : Dinv[64 * J + t * 8 + g];
        const double d1 = dir == 0 ? Dinv[64 * J + g * 8 + 4 + t] : Dinv[64 * J + (4 + t) * 8 + g];
        double xb0[CT], xb1[CT];
        double x[CT][2];
#pragma unroll
        for (int q = 0; q < CT; ++q) {
          const double* tile = col[q] + 8 * J * ldw;
          const double b0 = tile[t * ldw + g], b1 = tile[(4 + t) * ldw + g];
          x[q][0] = 0.0; x[q][1] = 0.0;
          dmma(x[q], d0, b0);
          dmma(x[q], d1, b1);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CT; ++q)
          if (on[q]) *reinterpret_cast<double2*>(col[q] + (8 * J + g) * ldw + 2 * t) = make_double2(x[q][0], x[q][1]);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CT; ++q) {
          const double* tile = col[q] + 8 * J * ldw;
          xb0[q] = tile[t * ldw + g]; xb1[q] = tile[(4 + t) * ldw + g];
        }
        // other block rows: tile_I -= L_IJ X_J (forward, I > J) or (L_JI)^T X_J (backward, I < J)
        const int i_lo = dir == 0 ? J + 1 : 0, i_hi = dir == 0 ? NBK : J;
#pragma unroll 2
        for (int I = i_lo; I < i_hi; ++I) {
          double a0, a1;
          if (dir == 0) {
            const double* la = As + (8 * I + g) * LD + 8 * J;
            a0 = -la[t]; a1 = -la[4 + t];
          } else {
            const double* la = As + (8 * J) * LD + 8 * I + g;  // element (g, k) = L[8 J + k][8 I + g]
            a0 = -la[t * LD]; a1 = -la[(4 + t) * LD];
          }
#pragma unroll
          for (int q = 0; q < CT; ++q) {
            double2* cp = reinterpret_cast<double2*>(col[q] + (8 * I + g) * ldw + 2 * t);
            const double2 v = *cp;
            double c[2] = {v.x, v.y};
            dmma(c, a0, xb0[q]);
            dmma(c, a1, xb1[q]);
            if (on[q]) *cp = make_double2(c[0], c[1]);
          }
        }
        __syncwarp();
      }
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
  if (nct <= kB2Warps) b2_solve_tiles<NBK, 1>(As, Dinv, Ws, ldw, nct);
  else b2_solve_tiles<NBK, Cfg::CT>(As, Dinv, Ws, ldw, nct);
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
  B2_T0();
  b2_stage(As, LD, lv.A, Mp);
  cp_commit();
  for (int i = threadIdx.x; i < Mp * 8; i += kB2Threads) Ws[(i >> 3) * ldw + (i & 7)] = (i & 7) == 0 ? lv.b[i >> 3] : 0.0;
  cp_wait<0>();
  __syncthreads();
  B2_T(8);
  b2_cholesky<NBK>(As, Dinv, fail);
  B2_T(9);
  b2_solve_tiles<NBK, 1>(As, Dinv, Ws, ldw, 1);
  __syncthreads();
  B2_T(10);
  for (int i = threadIdx.x; i < Mp; i += kB2Threads) x[i] = Ws[i * ldw];
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
    // b': warp per row; lanes stride the 8 Mp-long dot products
    double* red = b2_sm;
    for (int i = threadIdx.x; i < 2 * Mp; i += kB2RedThreads) {
      const bool left = i >= Mp;
      red[i] = (left ? has_l : has_r) ? lv.yh[size_t(left ? ql : qr) * Mp + (left ? i - Mp : i)] : 0.0;
    }
    __syncthreads();
    const double* Be = has_r ? lv.B + size_t(e) * MM : nullptr;
    const double* Bl = has_l ? lv.B + size_t(e - 1) * MM : nullptr;
    for (int i = warp; i < Mp; i += kB2RedWarps) {
      double s = 0.0;
      for (int k = lane; k < Mp; k += 32) {
        if (has_r) s += Be[size_t(k) * Mp + i] * red[k];        // (B_e^T yh_r)[i]
        if (has_l) s += Bl[size_t(i) * Mp + k] * red[Mp + k];   // (B_{e-1} yh_l)[i]
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) nx.b[size_t(pe) * Mp + i] = lv.b[size_t(e) * Mp + i] - s;
    }
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
  return (size_t(Mp) * b2_ld(Mp) + size_t(nbk) * 64 + size_t(Mp) * (8 * nct_cta + 4)) * sizeof(double);
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
  const size_t smem_r = size_t(2) * (8 * NBK) * b2_ld(8 * NBK) * sizeof(double);
  PBA_LAUNCH(h, K_BCR, k_b2_reduce<NBK>, dim3(nx.n, PA + PB + 1), dim3(kB2RedThreads), smem_r, lv, nx, PA, PB);
  return PBA_OK;
}
template <int NBK>
pba_status b2_launch_top(Handle* h, const B2Level& lv, double* x) {
  const size_t smem = (size_t(8 * NBK) * b2_ld(8 * NBK) + size_t(NBK) * 64 + size_t(8 * NBK) * 12) * sizeof(double);
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
