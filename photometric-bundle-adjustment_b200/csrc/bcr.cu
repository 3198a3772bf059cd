// bcr.cu — block cyclic reduction (BCR) solver for block-banded reduced camera
// systems: the parallel, exact RCS solver.
//
// With windowed covisibility keyframe i only couples to i +- bw, so grouping m =
// bw consecutive keyframes into one "super block" (M = m * cd unknowns, 96 for
// the photometric 12-keyframe window) turns the RCS into an SPD BLOCK-
// TRIDIAGONAL system  A_p x_p + B_{p-1} x_{p-1} + B_p^T x_{p+1} = b_p,
// B_p = A[p+1][p].  Cyclic reduction eliminates every odd super block of a
// level in parallel (one CTA each):
//     A_p = L L^T,  U = L^-1 B_{p-1},  V = L^-1 B_p^T,  y = L^-1 b_p
// and the surviving even blocks pick up the Schur complements
//     A'_e = A_e - V_{e-1}^T V_{e-1} - U_{e+1}^T U_{e+1},
//     B'   = -V_o^T U_o  (coupling across the eliminated block o),
//     b'_e = b_e - V_{e-1}^T y_{e-1} - U_{e+1}^T y_{e+1}
// which is again SPD block tridiagonal with half the blocks.  log2(S) levels
// later one block remains; the back-substitution retraces the levels:
//     x_o = L^-T (y_o - U x_{o-1} - V x_{o+1}).
//
// This replaces Ceres' sequential sparse LDL^T (internal/ceres/eigensparse.cc:
// 56-106, the "serial hot spot" of SURVEY.md §3.1) with O(log S) dependent
// steps of dense M x M work — the only true dense fp64 contraction on the path.
// The sequential band factorisation in solve.cu takes 8.3 ms for 2,000
// keyframes on one SM; it also would not shrink when the landmarks are sharded
// over GPUs (every rank solves the same RCS), capping multi-GPU scaling.
#include <stdio.h>

#include "launch.h"
#include "pba_internal.h"
#include "bcr_dense.cuh"

namespace pba {

namespace {

struct BcrLevel {
  int n;            // super blocks at this level
  double* A;        // [n][M*M]
  double* B;        // [n-1][M*M]   B[p] = A[p+1][p]
  double* b;        // [n][M]
  double* L;        // [n/2][M*M]   per odd block: Cholesky factor
  double* U;        // [n/2][M*M]
  double* V;        // [n/2][M*M]
  double* y;        // [n/2][M]
  double* D;        // [n/2][M*NB]  inverses of the factor's diagonal blocks
};

// level-0 assembly from the block-sparse RCS
__global__ void k_bcr_build(int cd, int m, int M, int64_t n_blocks, int n_slots, const int* __restrict__ blk_row,
                            const int* __restrict__ blk_col, const double* __restrict__ S, const double* __restrict__ rhs,
                            int n_super, double* __restrict__ A, double* __restrict__ B, double* __restrict__ b) {
  const int64_t blk = blockIdx.x;
  const int e = threadIdx.x;
  if (blk < n_blocks) {
    if (e >= cd * cd) return;
    const int r = e / cd, c = e % cd;
    const int a = blk_row[blk], bb = blk_col[blk];
    const int sa = a / m, sb = bb / m;
    const double v = S[blk * cd * cd + e];
    const int ra = (a % m) * cd + r, cb = (bb % m) * cd + c;
    if (sa == sb) {
      A[(int64_t(sa) * M + ra) * M + cb] = v;
      A[(int64_t(sa) * M + cb) * M + ra] = v;
    } else {  // sb == sa + 1: coupling A[sb][sa] = block(a,bb)^T
      B[(int64_t(sa) * M + cb) * M + ra] = v;
    }
  } else {
    // right-hand side and identity padding of the last super block
    const int s = int(blk - n_blocks);
    if (s >= n_super) return;
    for (int i = e; i < M; i += blockDim.x) {
      const int g = s * M + i;
      const bool real = g < n_slots * cd;
      b[g] = real ? rhs[g] : 0.0;
      if (!real) A[(int64_t(s) * M + i) * M + i] = 1.0;
    }
  }
}

// Eliminate the odd super blocks of a level, in two launches so that neither has a long chain:
//   k_bcr_factor  one CTA per odd block:  A_p = L L^T (L and the inverses of its diagonal blocks stored)
//   k_bcr_solve   kSolveCols right-hand-side columns per CTA of  [U | V | y] = L^-1 [B[p-1] | B[p]^T | b[p]]
// (the triangular solves are independent per column: at the sparse upper levels they spread over
// up to 12 SMs per block instead of queueing behind the factorisation on one).
template <int NB>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_factor(int M, BcrLevel lv, int* __restrict__ fail) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;                      // [M][ld]
  double* Dinv = Ls + M * ld;           // [M/NB][NB*NB]
  double* Pt = Dinv + M * NB;           // [NB][bcr_ldp(M)]  Cholesky panel scratch
  const int q = blockIdx.x, p = 2 * q + 1;
  const int tid = threadIdx.x;
  BCR_STAMP(0);
  cta_load(Ls, ld, lv.A + int64_t(p) * M * M, M, false);
  cta_load_wait();
  BCR_STAMP(1);
  cta_cholesky<NB>(Ls, M, ld, Dinv, Pt, fail);
  BCR_STAMP(2);
  cta_store(lv.L + int64_t(q) * M * M, Ls, ld, M, true);
  for (int i = tid; i < M * NB; i += kBcrThreads) lv.D[int64_t(q) * M * NB + i] = Dinv[i];
  BCR_STAMP(3);
}

// kSolveCols columns per CTA: 16 at the sparse upper levels (12 CTAs per block, short chains),
// 32 while there are more blocks than SMs can hold at 12 CTAs each.
template <int NB, int kSolveCols>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_solve(int M, BcrLevel lv) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  constexpr int ldw = kSolveCols + 1;
  double* Ls = sm;                      // [M][ld]
  double* Dinv = Ls + M * ld;           // [M/NB][NB*NB]
  double* Ws = Dinv + M * NB;           // [M][ldw]
  const int q = blockIdx.x, p = 2 * q + 1;
  const int col0 = blockIdx.y * kSolveCols;  // first column of [U | V | y] handled here
  const int tid = threadIdx.x;
  const bool has_v = p + 1 < lv.n;
  const int ncols_all = 2 * M + 1;
  if (col0 >= ncols_all) return;
  // a slice entirely inside V of a block without right neighbour has nothing to do
  if (!has_v && col0 >= M && col0 + kSolveCols <= 2 * M) return;
  BCR_STAMP(4);
  cta_load(Ls, ld, lv.L + int64_t(q) * M * M, M, false, true);
  for (int i = tid; i < M * NB; i += kBcrThreads) Dinv[i] = lv.D[int64_t(q) * M * NB + i];
  {
    const double* Bl = lv.B + int64_t(p - 1) * M * M;
    const double* Br = has_v ? lv.B + int64_t(p) * M * M : nullptr;
    const double* bp = lv.b + int64_t(p) * M;
    // U columns are contiguous along c, V columns (B[p]^T) along i: index the threads accordingly
    for (int e = tid; e < M * kSolveCols; e += kBcrThreads) {
      const int cU = e % kSolveCols, iU = e / kSolveCols;  // c fastest
      const int iV = e % M, cV = e / M;                    // i fastest
      const int colU = col0 + cU, colV = col0 + cV;
      if (colU < M) Ws[iU * ldw + cU] = Bl[int64_t(iU) * M + colU];
      if (colV >= M && colV < 2 * M) Ws[iV * ldw + cV] = has_v ? Br[int64_t(colV - M) * M + iV] : 0.0;
      if (colU == 2 * M) Ws[iU * ldw + cU] = bp[iU];
      if (colU > 2 * M) Ws[iU * ldw + cU] = 0.0;
    }
  }
  cta_load_wait();
  BCR_STAMP(5);
  // forward substitution: thread (c, g) owns column c and the rows i = g (mod 16) of every update
  const int c = tid % kSolveCols, g = tid / kSolveCols;
  constexpr int kGroups = kBcrThreads / kSolveCols;
  const int nbk = M / NB;
  for (int J = 0; J < nbk; ++J) {
    const int j0 = J * NB, n0 = j0 + NB;
    const double* Di = Dinv + J * NB * NB;
    if (g == 0) {
      double v[NB], o[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) v[k] = Ws[(j0 + k) * ldw + c];
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k <= r; ++k) t += Di[r * NB + k] * v[k];
        o[r] = t;
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) Ws[(j0 + k) * ldw + c] = o[k];
    }
    __syncthreads();
    if (n0 + g < M) {
      double w[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) w[k] = Ws[(j0 + k) * ldw + c];
      for (int i = n0 + g; i < M; i += 2 * kGroups) {
        const int i2 = i + kGroups;
        const bool two = i2 < M;
        const int ib = two ? i2 : i;
        double la[NB], lb[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) { la[k] = Ls[i * ld + j0 + k]; lb[k] = Ls[ib * ld + j0 + k]; }
        double sa = Ws[i * ldw + c], sb = Ws[ib * ldw + c];
#pragma unroll
        for (int k = 0; k < NB; ++k) { sa -= la[k] * w[k]; sb -= lb[k] * w[k]; }
        Ws[i * ldw + c] = sa;
        if (two) Ws[i2 * ldw + c] = sb;
      }
    }
    __syncthreads();
  }
  BCR_STAMP(6);
  {
    double* Uq = lv.U + int64_t(q) * M * M;
    double* Vq = lv.V + int64_t(q) * M * M;
    double* yq = lv.y + int64_t(q) * M;
    for (int e = tid; e < M * kSolveCols; e += kBcrThreads) {
      const int cc = e % kSolveCols, i = e / kSolveCols;
      const int col = col0 + cc;
      const double v = Ws[i * ldw + cc];
      if (col < M) Uq[int64_t(i) * M + col] = v;
      else if (col < 2 * M) { if (has_v) Vq[int64_t(i) * M + (col - M)] = v; }
      else if (col == 2 * M) yq[i] = v;
    }
  }
  BCR_STAMP(7);
}

// Even super blocks of a level -> next level.  Nine CTAs per block (blockIdx.y) so the
// sparse upper levels still fill SMs and every CTA's dependent chain is short:
//   y in [0, 4):  A' = A_e - V_{e-1}^T V_{e-1} - U_{e+1}^T U_{e+1}: the lower-triangle 4x4 tiles,
//                 enumerated without holes, a quarter each; the k range of every tile is split over
//                 four thread groups and reduced through shared memory
//   y in [4, 8):  B' = -V_o^T U_o (coupling across the eliminated block o = e + 1), a quarter of
//                 the tiles each, k split over two thread groups
//   y == 8:       b' = b_e - V_{e-1}^T y_{e-1} - U_{e+1}^T y_{e+1}
// SPLIT = true is the latency-optimised form above (9 CTAs per block).  With more blocks than
// SMs the redundant operand loads of the 9 parts dominate instead, so the dense lower levels run
// SPLIT = false: one CTA for A' (every thread a tile, full k range), one for B', one for b'.
template <bool SPLIT>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_reduce(int M, BcrLevel lv, BcrLevel nx) {
  constexpr int kPartsA = SPLIT ? 4 : 1, kPartsB = SPLIT ? 4 : 1;
  constexpr int kGroupsA = SPLIT ? 4 : 1, kGroupsB = SPLIT ? 2 : 1;   // k-range split inside a CTA
  constexpr int kSlotsA = kBcrThreads / kGroupsA, kSlotsB = kBcrThreads / kGroupsB;
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld(M);
  double* Xs = sm;            // [M][ld]
  double* Ys = sm + M * ld;   // [M][ld]
  const int pe = blockIdx.x;  // position at the next level
  const int part = blockIdx.y;
  const int p = 2 * pe;       // even position at this level
  const int tid = threadIdx.x;
  double* An = nx.A + int64_t(pe) * M * M;
  const double* Ae = lv.A + int64_t(p) * M * M;
  const bool has_l = p - 1 >= 0, has_r = p + 1 < lv.n;
  const double* Vl = has_l ? lv.V + int64_t((p - 1) / 2) * M * M : nullptr;
  const double* Ur = has_r ? lv.U + int64_t((p + 1) / 2) * M * M : nullptr;
  const double* Vr = has_r ? lv.V + int64_t((p + 1) / 2) * M * M : nullptr;
  const int T = (M + 3) / 4;
  BCR_STAMP(8);
  if (part == kPartsA + kPartsB) {
    // b': thread (i, half) accumulates one of the two products for row i; coalesced over i
    double* red = sm;
    const int i = tid % 128, half = tid / 128;
    double s = 0.0;
    if (i < M && (half == 0 ? has_l : has_r)) {
      const double* Z = half == 0 ? Vl : Ur;
      const double* y = lv.y + int64_t(half == 0 ? (p - 1) / 2 : (p + 1) / 2) * M;
#pragma unroll 8
      for (int k = 0; k < M; ++k) s += Z[int64_t(k) * M + i] * y[k];
    }
    red[tid] = s;
    __syncthreads();
    if (tid < M) nx.b[int64_t(pe) * M + tid] = lv.b[int64_t(p) * M + tid] - red[tid] - red[128 + tid];
    return;
  }
  if (part >= kPartsA) {
    if (p + 2 >= lv.n) return;
    // B' = -Vr^T Ur: tiles [t0, t1) of the T x T grid
    const int nt = T * T, per = (nt + kPartsB - 1) / kPartsB;
    const int t0 = (part - kPartsA) * per, t1 = min(nt, t0 + per);
    cta_load(Xs, ld, Ur, M, false);
    cta_load(Ys, ld, Vr, M, false);
    cta_load_wait();
    const int g = tid / kSlotsB, ul = tid % kSlotsB;
    double* red = sm;  // [kSlotsB][16] partial tiles of group 1 (the operands are dead by then)
    double* Bn = nx.B + int64_t(pe) * M * M;
    for (int base = t0; base < t1; base += kSlotsB) {  // uniform trip count
      const int t = base + ul;
      const bool live = t < t1;
      const int tr = live ? t / T : 0, tc = live ? t % T : 0;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      if (live) tile_xty(Ys, Xs, ld, tr, tc, (M * g) / kGroupsB, (M * (g + 1)) / kGroupsB, acc);
      if (kGroupsB > 1) {
        __syncthreads();  // every read of Xs / Ys of this round is done
        if (g == 1) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) red[ul * 16 + 4 * a + b] = acc[a][b];
        }
        __syncthreads();
      }
      if (g == 0 && live) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int r = 4 * tr + a, c = 4 * tc + b;
            if (r < M && c < M)
              Bn[int64_t(r) * M + c] = -(acc[a][b] + (kGroupsB > 1 ? red[ul * 16 + 4 * a + b] : 0.0));
          }
      }
      if (kGroupsB > 1 && base + kSlotsB < t1) {  // another round: the partials overwrote the operands
        __syncthreads();
        cta_load(Xs, ld, Ur, M, false);
        cta_load(Ys, ld, Vr, M, false);
        cta_load_wait();
      }
    }
    return;
  }
  // A': lower-triangle tiles u = tr (tr + 1) / 2 + tc, tc <= tr
  const int nt = T * (T + 1) / 2, per = (nt + kPartsA - 1) / kPartsA;
  const int u0 = part * per, u1 = min(nt, u0 + per);
  BCR_STAMP(9);
  if (has_l) cta_load(Xs, ld, Vl, M, false);
  if (has_r) cta_load(Ys, ld, Ur, M, false);
  cta_load_wait();
  BCR_STAMP(10);
  const int g = tid / kSlotsA, ul = tid % kSlotsA;
  double* red = sm;  // [kGroupsA - 1][kSlotsA][16]
  for (int base = u0; base < u1; base += kSlotsA) {  // uniform trip count
    const int u = base + ul;
    const bool live = u < u1;
    int tr = 0, tc = 0;
    if (live) {
      tr = int((sqrtf(8.0f * float(u) + 1.0f) - 1.0f) * 0.5f);
      while (tr * (tr + 1) / 2 > u) --tr;
      while ((tr + 1) * (tr + 2) / 2 <= u) ++tr;
      tc = u - tr * (tr + 1) / 2;
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    if (live) {
      const int k0 = (M * g) / kGroupsA, k1 = (M * (g + 1)) / kGroupsA;
      if (has_l) tile_xty(Xs, Xs, ld, tr, tc, k0, k1, acc);
      if (has_r) tile_xty(Ys, Ys, ld, tr, tc, k0, k1, acc);
    }
    if (kGroupsA > 1) {
      __syncthreads();  // every read of Xs / Ys of this round is done
      if (g > 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) red[((g - 1) * kSlotsA + ul) * 16 + 4 * a + b] = acc[a][b];
      }
      __syncthreads();
    }
    if (g == 0 && live) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int r = 4 * tr + a, c = 4 * tc + b;
          if (r < M && c <= r) {
            double sum = acc[a][b];
            if (kGroupsA > 1) {
              const int e = ul * 16 + 4 * a + b;
#pragma unroll
              for (int gg = 0; gg < kGroupsA - 1; ++gg) sum += red[gg * kSlotsA * 16 + e];
            }
            const double v = Ae[int64_t(r) * M + c] - sum;
            An[int64_t(r) * M + c] = v;
            if (c < r) An[int64_t(c) * M + r] = v;
          }
        }
    }
    if (kGroupsA > 1 && base + kSlotsA < u1) {
      __syncthreads();
      if (has_l) cta_load(Xs, ld, Vl, M, false);
      if (has_r) cta_load(Ys, ld, Ur, M, false);
      cta_load_wait();
    }
  }
  BCR_STAMP(11);
}

// Last remaining block: x = A^-1 b (the survivor is always original block 0).
template <int NB>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_top(int M, BcrLevel lv, double* __restrict__ x,
                                                          int* __restrict__ fail) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;
  double* w = sm + M * ld;       // [M] as an M x 1 right-hand side (ldw = 1)
  double* Dinv = w + M;
  double* Pt = Dinv + M * NB;
  const int tid = threadIdx.x;
  cta_load(Ls, ld, lv.A, M, false);
  for (int i = tid; i < M; i += kBcrThreads) w[i] = lv.b[i];
  cta_load_wait();
  cta_cholesky<NB>(Ls, M, ld, Dinv, Pt, fail);
  cta_trsm_lower<NB>(Ls, ld, Dinv, w, 1, M, 1);
  cta_solve_lt<NB>(Ls, ld, Dinv, w, M);
  for (int i = tid; i < M; i += kBcrThreads) x[i] = w[i];
}

// Odd blocks of a level: x_o = L^-T (y_o - U x_{o-1} - V x_{o+1}); x indexed by ORIGINAL block (p << shift).
// UVS: U and V are staged in shared memory next to L (all three copies in flight at once);
// otherwise (three M x M matrices do not fit) their rows are read from global memory.
template <int NB, bool UVS>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_backsub(int M, BcrLevel lv, int shift, double* __restrict__ x) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;                               // [M][ld]
  double* Us = Ls + M * ld;                      // [M][ld]  (UVS only)
  double* Vs = Us + M * ld;                      // [M][ld]  (UVS only)
  double* w = UVS ? Vs + M * ld : Ls + M * ld;   // [M]
  double* xl = w + M;            // [M]
  double* xr = xl + M;           // [M]
  double* Dinv = xr + M;         // [M/NB][NB*NB]
  const int q = blockIdx.x, p = 2 * q + 1;
  const int tid = threadIdx.x;
  const bool has_r = p + 1 < lv.n;
  const double* U = lv.U + int64_t(q) * M * M;
  const double* V = lv.V + int64_t(q) * M * M;
  BCR_STAMP(16);
  for (int i = tid; i < M; i += kBcrThreads) {
    xl[i] = x[(int64_t(p - 1) << shift) * M + i];
    xr[i] = has_r ? x[(int64_t(p + 1) << shift) * M + i] : 0.0;
  }
  cta_load(Ls, ld, lv.L + int64_t(q) * M * M, M, false, true);
  if (UVS) {
    cta_load(Us, ld, U, M, false);
    if (has_r) cta_load(Vs, ld, V, M, false);
  }
  for (int i = tid; i < M * NB; i += kBcrThreads) Dinv[i] = lv.D[int64_t(q) * M * NB + i];
  cta_load_wait();
  BCR_STAMP(17);
  // w = y - U xl - V xr: thread (r, half) walks half of row r of U and V (odd ld: conflict-free
  // in shared memory; from global the 128-byte lines stay in L1 across the walk), one shuffle
  // joins the halves.  M <= 128.
  {
    const int r = tid >> 1, hf = tid & 1;
    const int c0 = hf ? M / 2 : 0, c1 = hf ? M : M / 2;
    double a0 = 0.0, a1 = 0.0;
    if (r < M) {
      const double* ur = UVS ? Us + r * ld : U + int64_t(r) * M;
      const double* vr = UVS ? Vs + r * ld : V + int64_t(r) * M;
#pragma unroll 4
      for (int c = c0; c < c1; ++c) a0 += ur[c] * xl[c];
      if (has_r) {
#pragma unroll 4
        for (int c = c0; c < c1; ++c) a1 += vr[c] * xr[c];
      }
    }
    a0 += a1;
    a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
    if (hf == 0 && r < M) w[r] = lv.y[int64_t(q) * M + r] - a0;
  }
  __syncthreads();
  BCR_STAMP(18);
  cta_solve_lt<NB>(Ls, ld, Dinv, w, M);
  BCR_STAMP(19);
  for (int i = tid; i < M; i += kBcrThreads) x[(int64_t(p) << shift) * M + i] = w[i];
}

}  // namespace

constexpr size_t kBcrSmemCap = 220 * 1024;

// per-kernel footprints (smaller than the two-matrix maximum so that two CTAs share an SM)
static size_t bcr_factor_smem(int M, int cd) {
  return (size_t(M) * bcr_ld_odd(M) + size_t(M) * cd + size_t(cd) * bcr_ldp(M) + 8) * sizeof(double);
}
static size_t bcr_backsub_smem(int M, int cd, bool uvs) {
  return (size_t(uvs ? 3 : 1) * M * bcr_ld_odd(M) + 3 * size_t(M) + size_t(M) * cd + 8) * sizeof(double);
}
static size_t bcr_solve_smem(int M, int cd, int cols) {
  return (size_t(M) * bcr_ld_odd(M) + size_t(M) * cd + size_t(M) * (cols + 1) + 8) * sizeof(double);
}

// two M x (M+1) matrices + vectors + the diagonal-block inverses
size_t bcr_smem_bytes(int M, int cd) { return (size_t(2) * M * (bcr_ld(M) > bcr_ld_odd(M) ? bcr_ld(M) : bcr_ld_odd(M)) + 4 * size_t(M) + size_t(M) * cd + size_t(cd) * bcr_ldp(M) + 8) * sizeof(double); }

// Super-block size (in keyframes) for a given half-bandwidth, or 0 when BCR does not apply.
int bcr_super_size(int cd, int bw, int n_slots) {
  const int m = bw < 1 ? 1 : bw;
  const int M = m * cd;
  if (bcr_smem_bytes(M, cd) > kBcrSmemCap || m > bcr_max_blocks(cd) || M > 128) return 0;  // two M x (M+1) fp64 matrices must fit shared memory
  if (n_slots < 2 * m) return 0;  // fewer than two super blocks: nothing to reduce
  return m;
}


pba_status bcr_setup(Handle* h) {
  const Sizes& z = h->sz;
  h->bcr_m = bcr_super_size(z.cd, h->rcs_bandwidth, z.n_slots);
  if (!h->bcr_m) return PBA_OK;
  const int m = h->bcr_m, M = m * z.cd;
  const int S = (z.n_slots + m - 1) / m;
  // carve every level out of one buffer
  size_t total = 0;
  std::vector<size_t> offA, offB, offb, offL, offU, offV, offy, offD;
  std::vector<int> ns;
  for (int n = S; ; n = (n + 1) / 2) {
    ns.push_back(n);
    offA.push_back(total); total += size_t(n) * M * M;
    offB.push_back(total); total += size_t(n > 1 ? n - 1 : 0) * M * M;
    offb.push_back(total); total += size_t(n) * M;
    const int no = n / 2;
    offL.push_back(total); total += size_t(no) * M * M;
    offU.push_back(total); total += size_t(no) * M * M;
    offV.push_back(total); total += size_t(no) * M * M;
    offy.push_back(total); total += size_t(no) * M;
    offD.push_back(total); total += size_t(no) * M * z.cd;
    if (n == 1) break;
  }
  PBA_CUDA_OK(h->bcr_ws.alloc(total + size_t(S) * M));
  h->bcr_levels = int(ns.size());
  h->bcr_n.assign(ns.begin(), ns.end());
  h->bcr_off.clear();
  for (size_t l = 0; l < ns.size(); ++l)
    for (size_t o : {offA[l], offB[l], offb[l], offL[l], offU[l], offV[l], offy[l], offD[l]}) h->bcr_off.push_back(o);
  h->bcr_x_off = total;
  // level 0 is rebuilt from the RCS blocks by k_bcr_build before every solve; the block pattern is
  // static, so the entries it never writes are zeroed here once (A and B of level 0 are contiguous)
  PBA_CUDA_OK(cudaMemsetAsync(h->bcr_ws.p + offA[0], 0, sizeof(double) * (size_t(S) * M * M + size_t(S > 1 ? S - 1 : 0) * M * M), h->stream));
  return bcr2_setup(h);
}

pba_status launch_bcr_rcs(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim == 0) return PBA_OK;
  const int m = h->bcr_m, M = m * z.cd;
  const int S = h->bcr_n[0];
  double* ws = h->bcr_ws.p;
  auto level = [&](int l) {
    BcrLevel v;
    v.n = h->bcr_n[l];
    const size_t* o = &h->bcr_off[size_t(l) * 8];
    v.A = ws + o[0]; v.B = ws + o[1]; v.b = ws + o[2]; v.L = ws + o[3]; v.U = ws + o[4]; v.V = ws + o[5]; v.y = ws + o[6];
    v.D = ws + o[7];
    return v;
  };
  const double* Sblk = h->rcs.p;
  const double* rhs = Sblk + z.n_blocks * z.cd * z.cd;
  double* x = ws + h->bcr_x_off;
  PBA_CUDA_OK(cudaMemsetAsync(h->chol_fail.p, 0, sizeof(int), h->stream));
  BcrLevel l0 = level(0);  // entries outside the static block pattern were zeroed once in bcr_setup
  PBA_LAUNCH(h, K_BCR, k_bcr_build, dim3((unsigned)(z.n_blocks + S)), dim3(64), 0, z.cd, m, M, z.n_blocks, z.n_slots,
             h->d_blk_row.p, h->d_blk_col.p, Sblk, rhs, S, l0.A, l0.B, l0.b);
  const size_t smem = bcr_smem_bytes(M, z.cd);
  const int nl = h->bcr_levels;
  const bool c8 = z.cd == 8;
  for (int l = 0; l + 1 < nl; ++l) {
    BcrLevel lv = level(l), nx = level(l + 1);
    const int n_odd = lv.n / 2;
    const bool wide = n_odd > 24;  // 12 CTAs per block would exceed two CTAs per SM
    const int sc = wide ? 32 : 16;
    const dim3 gs(n_odd, (2 * M + 1 + sc - 1) / sc);
    const size_t smem_f = bcr_factor_smem(M, z.cd), smem_s = bcr_solve_smem(M, z.cd, sc);
    if (c8) {
      PBA_LAUNCH(h, K_BCR, k_bcr_factor<8>, dim3(n_odd), dim3(kBcrThreads), smem_f, M, lv, h->chol_fail.p);
      if (wide) { PBA_LAUNCH(h, K_BCR, (k_bcr_solve<8, 32>), gs, dim3(kBcrThreads), smem_s, M, lv); }
      else { PBA_LAUNCH(h, K_BCR, (k_bcr_solve<8, 16>), gs, dim3(kBcrThreads), smem_s, M, lv); }
    } else {
      PBA_LAUNCH(h, K_BCR, k_bcr_factor<6>, dim3(n_odd), dim3(kBcrThreads), smem_f, M, lv, h->chol_fail.p);
      if (wide) { PBA_LAUNCH(h, K_BCR, (k_bcr_solve<6, 32>), gs, dim3(kBcrThreads), smem_s, M, lv); }
      else { PBA_LAUNCH(h, K_BCR, (k_bcr_solve<6, 16>), gs, dim3(kBcrThreads), smem_s, M, lv); }
    }
    if (nx.n > 40) { PBA_LAUNCH(h, K_BCR, k_bcr_reduce<false>, dim3(nx.n, 3), dim3(kBcrThreads), smem, M, lv, nx); }
    else { PBA_LAUNCH(h, K_BCR, k_bcr_reduce<true>, dim3(nx.n, 9), dim3(kBcrThreads), smem, M, lv, nx); }
  }
  if (c8) { PBA_LAUNCH(h, K_BCR, k_bcr_top<8>, dim3(1), dim3(kBcrThreads), smem, M, level(nl - 1), x, h->chol_fail.p); }
  else { PBA_LAUNCH(h, K_BCR, k_bcr_top<6>, dim3(1), dim3(kBcrThreads), smem, M, level(nl - 1), x, h->chol_fail.p); }
  for (int l = nl - 2; l >= 0; --l) {
    BcrLevel lv = level(l);
    const bool uvs = bcr_backsub_smem(M, z.cd, true) <= kBcrSmemCap;
    const size_t smem_b = bcr_backsub_smem(M, z.cd, uvs);
    if (c8) {
      if (uvs) { PBA_LAUNCH(h, K_BCR, (k_bcr_backsub<8, true>), dim3(lv.n / 2), dim3(kBcrThreads), smem_b, M, lv, l, x); }
      else { PBA_LAUNCH(h, K_BCR, (k_bcr_backsub<8, false>), dim3(lv.n / 2), dim3(kBcrThreads), smem_b, M, lv, l, x); }
    } else {
      if (uvs) { PBA_LAUNCH(h, K_BCR, (k_bcr_backsub<6, true>), dim3(lv.n / 2), dim3(kBcrThreads), smem_b, M, lv, l, x); }
      else { PBA_LAUNCH(h, K_BCR, (k_bcr_backsub<6, false>), dim3(lv.n / 2), dim3(kBcrThreads), smem_b, M, lv, l, x); }
    }
  }
  PBA_CUDA_OK(cudaMemcpyAsync(h->y_cam.p, x, sizeof(double) * z.dim, cudaMemcpyDeviceToDevice, h->stream));
#ifdef PBA_BCR_TIMING
  {
    static int calls = 0;
    if (++calls == 5) {
      cudaStreamSynchronize(h->stream);
      long long t[32];
      cudaMemcpyFromSymbol(t, g_bcr_t, sizeof(t));
      auto us = [&](int a, int b) { return double(t[b] - t[a]) / 1965.0; };
      fprintf(stderr, "[bcr] M=%d levels=%d | factor: load %.1f chol %.1f storeL %.1f | solve: load %.1f trsm %.1f store %.1f | reduce: b' %.1f load %.1f gemm %.1f | backsub: load %.1f matvec %.1f solve %.1f (us, last launch of each kernel, CTA 0)\n",
              M, nl, us(0, 1), us(1, 2), us(2, 3), us(4, 5), us(5, 6), us(6, 7), us(8, 9), us(9, 10), us(10, 11), us(16, 17), us(17, 18), us(18, 19));
    }
  }
#endif
  return PBA_OK;
}

}  // namespace pba
