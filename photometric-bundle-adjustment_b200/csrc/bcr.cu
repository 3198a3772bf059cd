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

namespace pba {

namespace {

constexpr int kBcrThreads = 256;

// shared-memory leading dimensions.  The product kernel (k_bcr_reduce) wants 16-byte aligned
// rows (even) covering the 4-wide tiles; the factorisation kernels walk columns with one thread
// per row, which is conflict-free only for an odd stride.
__host__ __device__ inline int bcr_ld(int M) { return ((M + 3) / 4) * 4 + 2; }
__host__ __device__ inline int bcr_ld_odd(int M) { return M + 1 + (M & 1); }

// ---- CTA-level dense kernels on shared-memory matrices, blocked by NB = cd ----
// All matrices are row-major with leading dimension ld.  The only serial piece is
// the NB x NB diagonal-block factorisation (warp 0); everything else is
// register-tiled, synchronised twice per block column.

// Factor the NB x NB block at D (lower Cholesky, in place) and write the inverse
// of the factor to Di [NB*NB] (lower).  Executed by warp 0; ends with __syncwarp.
template <int NB>
__device__ __forceinline__ void warp_factor_diag(double* D, int ld, double* Di, int* fail) {
  const int lane = threadIdx.x & 31;
  __shared__ double s_l[NB * NB];
  __shared__ double s_id[NB];
  if (lane == 0) {
    double L[NB][NB];
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
      s_id[j] = inv;
#pragma unroll
      for (int r = j + 1; r < NB; ++r) L[r][j] *= inv;
#pragma unroll
      for (int c = j + 1; c < NB; ++c)
#pragma unroll
        for (int r = c; r < NB; ++r) L[r][c] -= L[r][j] * L[c][j];
    }
#pragma unroll
    for (int r = 0; r < NB; ++r)
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        const double v = c <= r ? L[r][c] : 0.0;
        s_l[r * NB + c] = v;
        D[r * ld + c] = v;
      }
  }
  __syncwarp();
  if (lane < NB) {
    const int c = lane;  // column c of the inverse
    double m[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      double s = r == c ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < r; ++q) s -= (q >= c ? s_l[r * NB + q] * m[q] : 0.0);
      m[r] = r >= c ? s * s_id[r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) Di[r * NB + c] = m[r];
  }
  __syncwarp();
}

// A <- lower Cholesky factor of A (strict upper triangle untouched); Dinv[J] = inverse
// of the J-th diagonal block of the factor.
template <int NB>
__device__ void cta_cholesky(double* A, int M, int ld, double* Dinv, int* fail) {
  const int tid = threadIdx.x;
  const int nbk = M / NB;
  for (int J = 0; J < nbk; ++J) {
    const int j0 = J * NB;
    double* Di = Dinv + J * NB * NB;
    if (tid < 32) warp_factor_diag<NB>(A + j0 * ld + j0, ld, Di, fail);
    __syncthreads();
    // panel: rows below the diagonal block  <-  row * L_D^-T = row * Di^T
    for (int i = j0 + NB + tid; i < M; i += kBcrThreads) {
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
      for (int q = 0; q < NB; ++q) A[i * ld + j0 + q] = o[q];
    }
    __syncthreads();
    // trailing lower triangle -= panel panel^T, 4x4 register tiles
    const int n0 = j0 + NB, n = M - n0;
    const int T = (n + 3) / 4;
    for (int t = tid; t < T * T; t += kBcrThreads) {
      const int tr = t / T, tc = t % T;
      if (tc > tr) continue;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double x[4], y[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = n0 + 4 * tr + a, c = n0 + 4 * tc + a;
          x[a] = r < M ? A[r * ld + j0 + q] : 0.0;
          y[a] = c < M ? A[c * ld + j0 + q] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += x[a] * y[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int r = n0 + 4 * tr + a, c = n0 + 4 * tc + b;
          if (r < M && c <= r) A[r * ld + c] -= acc[a][b];
        }
    }
    __syncthreads();
  }
}

// W (M x ncols, shared) <- L^-1 W, blocked forward substitution.
template <int NB>
__device__ void cta_trsm_lower(const double* L, int ld, const double* Dinv, double* W, int ldw, int M, int ncols) {
  const int tid = threadIdx.x;
  const int nbk = M / NB;
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
    __syncthreads();
    // rows below: W[i][c] -= L[i][j0..] . W[j0..][c].  Thread (c = tid % 128, g = tid / 128) owns
    // column c and every other 4-row strip; the 8 pivot-row values stay in registers.
    const int n0 = j0 + NB;
    const int c = tid & 127, g = tid >> 7;
    for (int cc = c; cc < ncols; cc += 128) {
      double w[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) w[q] = W[(j0 + q) * ldw + cc];
      for (int i0 = n0 + 4 * g; i0 < M; i0 += 8) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = i0 + a;
          if (i < M) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < NB; ++q) s += L[i * ld + j0 + q] * w[q];
            W[i * ldw + cc] -= s;
          }
        }
      }
    }
    __syncthreads();
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

// out[r][c] = base[r][c] - sum_k X[k][r] Y[k][c]   (X, Y: M x M in shared memory with an even
// leading dimension -> 16-byte LDS; out/base global).  Only tile rows [tr0, tr1) are computed, so
// several CTAs can split one product.  lower_only: X == Y, compute c <= r and mirror.
__device__ void cta_xty_sub(const double* X, const double* Y, int ld, int M, double* __restrict__ out,
                            const double* __restrict__ base, bool lower_only, int tr0, int tr1) {
  const int T = (M + 3) / 4;
  const int ntile = (tr1 - tr0) * T;
  for (int t = threadIdx.x; t < ntile; t += kBcrThreads) {
    const int tr = tr0 + t / T, tc = t % T;
    if (lower_only && tc > tr) continue;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    // M is a multiple of cd (6 or 8); pad columns beyond M read finite junk only when M % 4 != 0,
    // and those accumulators are never stored
    const double* xp = X + 4 * tr;
    const double* yp = Y + 4 * tc;
#pragma unroll 2
    for (int k = 0; k < M; ++k) {
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
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * tr + a, c = 4 * tc + b;
        if (r < M && c < M) {
          const double v = (base ? base[int64_t(r) * M + c] : 0.0) - acc[a][b];
          out[int64_t(r) * M + c] = v;
          if (lower_only && c < r) out[int64_t(c) * M + r] = v;
        }
      }
  }
}

// global (M x M, dense) -> shared (leading dimension ld), optionally transposed, with cp.async
// (LDGSTS): every copy is in flight at once; call cta_load_wait() before reading.
__device__ __forceinline__ void cta_load(double* dst, int ld, const double* __restrict__ src, int M, bool transpose) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned d0 = unsigned(__cvta_generic_to_shared(dst));
  for (int r = warp; r < M; r += kBcrThreads / 32) {
    const double* s = src + int64_t(r) * M;
    for (int c = lane; c < M; c += 32) {
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

// Eliminate the odd super blocks of a level.  TWO CTAs per block (blockIdx.y), each factoring
// A_p itself (redundant, but the two triangular solves are the longer half of the work):
//   y = 0:  L (stored), U = L^-1 B[p-1], y = L^-1 b[p]
//   y = 1:  V = L^-1 B[p]^T  (only when block p + 1 exists)
template <int NB>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_eliminate(int M, BcrLevel lv, int* __restrict__ fail) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;                      // [M][ld]
  double* Ws = sm + M * ld;             // [M][ld]  right-hand sides (column M = y)
  double* Dinv = Ws + M * ld;           // [M/NB][NB*NB]
  const int q = blockIdx.x, p = 2 * q + 1;
  const int part = blockIdx.y;
  const int tid = threadIdx.x;
  if (part == 1 && p + 1 >= lv.n) return;
  cta_load(Ls, ld, lv.A + int64_t(p) * M * M, M, false);
  // the right-hand sides stream in while the factorisation runs
  if (part == 0) {
    cta_load(Ws, ld, lv.B + int64_t(p - 1) * M * M, M, false);
    for (int i = tid; i < M; i += kBcrThreads) Ws[i * ld + M] = lv.b[int64_t(p) * M + i];
  } else {
    cta_load(Ws, ld, lv.B + int64_t(p) * M * M, M, true);
  }
  cta_load_wait();
  cta_cholesky<NB>(Ls, M, ld, Dinv, fail);
  if (part == 0) {
    cta_store(lv.L + int64_t(q) * M * M, Ls, ld, M, true);
    for (int i = tid; i < M * NB; i += kBcrThreads) lv.D[int64_t(q) * M * NB + i] = Dinv[i];
    cta_trsm_lower<NB>(Ls, ld, Dinv, Ws, ld, M, M + 1);
    cta_store(lv.U + int64_t(q) * M * M, Ws, ld, M, false);
    for (int i = tid; i < M; i += kBcrThreads) lv.y[int64_t(q) * M + i] = Ws[i * ld + M];
  } else {
    cta_trsm_lower<NB>(Ls, ld, Dinv, Ws, ld, M, M);
    cta_store(lv.V + int64_t(q) * M * M, Ws, ld, M, false);
  }
}

// Even super blocks of a level -> next level.  THREE CTAs per block (blockIdx.y) so the
// sparse upper levels still fill SMs:
//   y = 0, 1:  A' = A_e - V_{e-1}^T V_{e-1} - U_{e+1}^T U_{e+1}  (tile rows split in halves; y = 0 also b')
//   y = 2:     B' = -V_o^T U_o  (coupling across the eliminated block o = e + 1)
__global__ void __launch_bounds__(kBcrThreads) k_bcr_reduce(int M, BcrLevel lv, BcrLevel nx) {
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
  if (part == 2) {
    if (p + 2 < lv.n) {
      cta_load(Xs, ld, Ur, M, false);
      cta_load(Ys, ld, Vr, M, false);
      cta_load_wait();
      cta_xty_sub(Ys, Xs, ld, M, nx.B + int64_t(pe) * M * M, nullptr, false, 0, T);
    }
    return;
  }
  if (part == 0) {
    // b' (vector work, straight from global)
    for (int i = tid; i < M; i += kBcrThreads) {
      double s = lv.b[int64_t(p) * M + i];
      if (has_l) {
        const double* y = lv.y + int64_t((p - 1) / 2) * M;
        for (int k = 0; k < M; ++k) s -= Vl[int64_t(k) * M + i] * y[k];
      }
      if (has_r) {
        const double* y = lv.y + int64_t((p + 1) / 2) * M;
        for (int k = 0; k < M; ++k) s -= Ur[int64_t(k) * M + i] * y[k];
      }
      nx.b[int64_t(pe) * M + i] = s;
    }
  }
  // the symmetric products are computed on the lower triangle: balance the two halves by area
  const int split = int(0.7071 * T + 0.5);
  const int tr0 = part == 0 ? 0 : split, tr1 = part == 0 ? split : T;
  if (has_l) cta_load(Xs, ld, Vl, M, false);
  if (has_r) cta_load(Ys, ld, Ur, M, false);
  cta_load_wait();
  // rows [4 tr0, 4 tr1) of the lower triangle (+ their mirror images); the two row ranges of
  // the two CTAs write disjoint entries, and each entry is final after one pass
  for (int t = tid; t < (tr1 - tr0) * T; t += kBcrThreads) {
    const int tr = tr0 + t / T, tc = t % T;
    if (tc > tr) continue;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 0 ? !has_l : !has_r) continue;
      const double* Z = pass == 0 ? Xs : Ys;
      const double* xp = Z + 4 * tr;
      const double* yp = Z + 4 * tc;
#pragma unroll 2
      for (int k = 0; k < M; ++k) {
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
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = 4 * tr + a, c = 4 * tc + b;
        if (r < M && c < M) {
          const double v = Ae[int64_t(r) * M + c] - acc[a][b];
          An[int64_t(r) * M + c] = v;
          if (c < r) An[int64_t(c) * M + r] = v;
        }
      }
  }
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
  const int tid = threadIdx.x;
  cta_load(Ls, ld, lv.A, M, false);
  for (int i = tid; i < M; i += kBcrThreads) w[i] = lv.b[i];
  cta_load_wait();
  cta_cholesky<NB>(Ls, M, ld, Dinv, fail);
  cta_trsm_lower<NB>(Ls, ld, Dinv, w, 1, M, 1);
  cta_solve_lt<NB>(Ls, ld, Dinv, w, M);
  for (int i = tid; i < M; i += kBcrThreads) x[i] = w[i];
}

// Odd blocks of a level: x_o = L^-T (y_o - U x_{o-1} - V x_{o+1}); x indexed by ORIGINAL block (p << shift).
template <int NB>
__global__ void __launch_bounds__(kBcrThreads) k_bcr_backsub(int M, BcrLevel lv, int shift, double* __restrict__ x) {
  extern __shared__ __align__(16) double sm[];
  const int ld = bcr_ld_odd(M);
  double* Ls = sm;               // [M][ld]
  double* w = sm + M * ld;       // [M]
  double* xl = w + M;            // [M]
  double* xr = xl + M;           // [M]
  double* Dinv = xr + M;         // [M/NB][NB*NB]
  const int q = blockIdx.x, p = 2 * q + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_r = p + 1 < lv.n;
  for (int i = tid; i < M; i += kBcrThreads) {
    xl[i] = x[(int64_t(p - 1) << shift) * M + i];
    xr[i] = has_r ? x[(int64_t(p + 1) << shift) * M + i] : 0.0;
  }
  cta_load(Ls, ld, lv.L + int64_t(q) * M * M, M, false);
  for (int i = tid; i < M * NB; i += kBcrThreads) Dinv[i] = lv.D[int64_t(q) * M * NB + i];
  cta_load_wait();
  // w = y - U xl - V xr: one warp per row, lanes stride the columns (coalesced)
  const double* U = lv.U + int64_t(q) * M * M;
  const double* V = lv.V + int64_t(q) * M * M;
  for (int r = warp; r < M; r += kBcrThreads / 32) {
    double s = 0.0;
    for (int c = lane; c < M; c += 32) {
      s += U[int64_t(r) * M + c] * xl[c];
      if (has_r) s += V[int64_t(r) * M + c] * xr[c];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) w[r] = lv.y[int64_t(q) * M + r] - s;
  }
  __syncthreads();
  cta_solve_lt<NB>(Ls, ld, Dinv, w, M);
  for (int i = tid; i < M; i += kBcrThreads) x[(int64_t(p) << shift) * M + i] = w[i];
}

}  // namespace

// two M x (M+1) matrices + vectors + the diagonal-block inverses
size_t bcr_smem_bytes(int M, int cd) { return (size_t(2) * M * (bcr_ld(M) > bcr_ld_odd(M) ? bcr_ld(M) : bcr_ld_odd(M)) + 4 * size_t(M) + size_t(M) * cd + 8) * sizeof(double); }

// Super-block size (in keyframes) for a given half-bandwidth, or 0 when BCR does not apply.
int bcr_super_size(int cd, int bw, int n_slots) {
  const int m = bw < 1 ? 1 : bw;
  const int M = m * cd;
  if (bcr_smem_bytes(M, cd) > 220 * 1024) return 0;  // two M x (M+1) fp64 matrices must fit shared memory
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
  const int smem = int(bcr_smem_bytes(M, z.cd));
  if (z.cd == 8) {
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_eliminate<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_top<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_backsub<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else {
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_eliminate<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_top<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_backsub<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  PBA_CUDA_OK(cudaFuncSetAttribute(k_bcr_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return PBA_OK;
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
  BcrLevel l0 = level(0);
  PBA_CUDA_OK(cudaMemsetAsync(l0.A, 0, sizeof(double) * (size_t(S) * M * M + size_t(S > 1 ? S - 1 : 0) * M * M), h->stream));
  PBA_LAUNCH(h, K_BCR, k_bcr_build, dim3((unsigned)(z.n_blocks + S)), dim3(64), 0, z.cd, m, M, z.n_blocks, z.n_slots,
             h->d_blk_row.p, h->d_blk_col.p, Sblk, rhs, S, l0.A, l0.B, l0.b);
  const size_t smem = bcr_smem_bytes(M, z.cd);
  const int nl = h->bcr_levels;
  const bool c8 = z.cd == 8;
  for (int l = 0; l + 1 < nl; ++l) {
    BcrLevel lv = level(l), nx = level(l + 1);
    if (c8) { PBA_LAUNCH(h, K_BCR, k_bcr_eliminate<8>, dim3(lv.n / 2, 2), dim3(kBcrThreads), smem, M, lv, h->chol_fail.p); }
    else { PBA_LAUNCH(h, K_BCR, k_bcr_eliminate<6>, dim3(lv.n / 2, 2), dim3(kBcrThreads), smem, M, lv, h->chol_fail.p); }
    PBA_LAUNCH(h, K_BCR, k_bcr_reduce, dim3(nx.n, 3), dim3(kBcrThreads), smem, M, lv, nx);
  }
  if (c8) { PBA_LAUNCH(h, K_BCR, k_bcr_top<8>, dim3(1), dim3(kBcrThreads), smem, M, level(nl - 1), x, h->chol_fail.p); }
  else { PBA_LAUNCH(h, K_BCR, k_bcr_top<6>, dim3(1), dim3(kBcrThreads), smem, M, level(nl - 1), x, h->chol_fail.p); }
  for (int l = nl - 2; l >= 0; --l) {
    BcrLevel lv = level(l);
    if (c8) { PBA_LAUNCH(h, K_BCR, k_bcr_backsub<8>, dim3(lv.n / 2), dim3(kBcrThreads), smem, M, lv, l, x); }
    else { PBA_LAUNCH(h, K_BCR, k_bcr_backsub<6>, dim3(lv.n / 2), dim3(kBcrThreads), smem, M, lv, l, x); }
  }
  PBA_CUDA_OK(cudaMemcpyAsync(h->y_cam.p, x, sizeof(double) * z.dim, cudaMemcpyDeviceToDevice, h->stream));
  return PBA_OK;
}

}  // namespace pba
