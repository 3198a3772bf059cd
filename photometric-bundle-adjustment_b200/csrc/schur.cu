// schur.cu — normal-equation assembly and Schur elimination of the inverse
// distances into the reduced camera system (RCS), back-substitution,
// retraction and gradient norms.
//
// Replaces Ceres' SchurEliminator::{Eliminate,BackSubstitute}
// (internal/ceres/schur_eliminator_impl.h:177-375), the Jacobi column scaling
// (trust_region_minimizer.cc:261-276), the LM diagonal
// (levenberg_marquardt_strategy.cc:76-88) and Program::Plus
// (local_parameterization_se3.hpp:44-51).
//
// No global atomics anywhere: every accumulation target has exactly one
// producer.  The work is split so that each stage is a plain reduction:
//   k_edge_gram    per (host,target) edge chunk:  [J_h J_t r]^T [J_h J_t r]
//   k_lm_gather    per landmark:                  row of F^T E, c_l, g_l
//   k_schur_syrk   per host group:                W^T diag(sigma^2/ete) W
//   k_rcs_reduce   per RCS block:                 sum of its partial blocks
// (the per-block source lists are static and built once in pba_create).
#include <type_traits>

#include "launch.h"
#include "pba_internal.h"

namespace pba {

namespace {

// ------------------------------------------------------------ edge Gram ----
// One CTA (3 warps) per edge chunk: G = M^T M for M = [J(0..C-1) | 0.. | r], a
// (R * n_chunk) x 16 matrix.  The sum over M's rows is separable, so the CTA
// streams ONE Jacobian row k at a time: 16 planes x 256 observations = sixteen
// contiguous 2 KB runs per stage (long DRAM bursts; staging 32 observations of
// all R rows instead reads 128 scattered 256 B pieces and ran at 31 % DRAM
// utilisation).  Stages are copied with cp.async into a double buffer
// (transposed: Mt[column][obs]) so the copy overlaps the previous stage's math.
// The products run on the FP64 tensor cores: one mma.sync.m8n8k4 (DMMA) adds four
// observations to an 8x8 tile of G, and because G = M^T M the A fragment
// (column-of-M x observation) and the B fragment (observation x column-of-M) of a
// column block are the SAME register, so a k-step of all three upper tiles costs
// 2 LDS.64 + 3 DMMA per lane (the scalar 8x8 register-tile version needed
// 16 LDS.128 + 128 DFMA per 64 observations and was issue-bound at 32 % of the
// FP64 pipe: 4.6 ms for 18 M observations).  Warp w takes k-steps w, w+3, ...;
// the three warps' fragments are summed through shared memory once per chunk.
// JR = interleaved planes [R][C+1][ld] (column C of every row is the residual).
constexpr int kGramStages = 2;  // measured: 3 x 256 -> 4.45 ms, 4 x 128 -> 4.22 ms, 2 x 256 -> 4.09 ms

__device__ __forceinline__ void gram_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int R, int C>
struct GramCfg {
  static constexpr int CD = C - 7;
  static constexpr int TOBS = 256;             // observations per stage
  static constexpr int RS = TOBS + 4;          // row stride of Mt: = 4 (mod 16) doubles -> conflict-free fragment loads
  static constexpr int DIR_STRIDE = 3 * CD * CD + 2 * CD;
};

template <int R, int C>
__global__ void __launch_bounds__(96) k_edge_gram(int64_t ld, const int* __restrict__ chunk_edge,
                                                   const int64_t* __restrict__ chunk_begin,
                                                   const int64_t* __restrict__ chunk_end,
                                                   const int* __restrict__ edge_h, const int* __restrict__ edge_t,
                                                   const int* __restrict__ slot, const double* __restrict__ JR,
                                                   const double* __restrict__ edge_M, double* __restrict__ part_dir) {
  using Cfg = GramCfg<R, C>;
  constexpr int CD = Cfg::CD, TOBS = Cfg::TOBS, RS = Cfg::RS, P = C + 1;
  // K1 stores only the host-pose, affine, inverse-distance and residual planes — the target-pose
  // columns are (host-pose columns) x M with one 6x6 M per edge (eval.cu, k_edge_prep) — so the
  // matrix staged here is [J_h(6) J_a(NAFF) 0.. | r]: 9 rows instead of 16, two DMMAs per k-step
  // instead of three, and the full Gram matrix is E^T G9 E at the end.
  constexpr bool RED = true;
  constexpr int NAFF = C - 13;      // affine columns: 2 photometric, 0 geometric
  constexpr int PL = C + 1 - 6;     // stored planes per row (pba_internal.h: stored_plane)
  constexpr int NR = RED ? 9 : 16;  // staged rows of Mt
  extern __shared__ __align__(16) double gram_sm[];  // Mt[kGramStages][NR * RS] ring + G[256]
  double* G = gram_sm + kGramStages * NR * RS;
  const int q = blockIdx.x;
  const int e = chunk_edge[q];
  const int64_t o0 = chunk_begin[q], o1 = chunk_end[q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fo = lane & 3;  // fragment coordinates: column of M (within a block), observation
  // upper tiles (0,0), (0,1), (1,1) of G; two accumulator sets keep six independent DMMA chains in flight
  double c00[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, c01[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, c11[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  // columns C..14 of M are structurally zero (nothing to do for the photometric 15-column rows)
  if (NAFF < 2) {  // staged rows 6 + NAFF .. 7 have no source plane
    for (int i = threadIdx.x; i < kGramStages * (2 - NAFF) * RS; i += 96) {
      const int b = i / ((2 - NAFF) * RS), j = i % ((2 - NAFF) * RS);
      gram_sm[b * NR * RS + (6 + NAFF) * RS + j] = 0.0;
    }
    __syncthreads();
  }

  // Stages start at an EVEN observation (o0 rounded down) so that every lane copies an aligned
  // pair with one 16-byte cp.async (half the LDGSTS instructions: the 8-byte version was
  // MIO-throttled); the observation before o0, if any, and everything from o1 on are zero-filled.
  const int64_t o0al = o0 & ~int64_t(1);
  const int n_t = int((o1 - o0al + TOBS - 1) / TOBS);  // obs tiles per row
  const int n_stage = n_t * R;                         // stage s = (row k = s / n_t, tile s % n_t)
  auto stage = [&](int st, int buf) {
    const int k = st / n_t;
    const int64_t base = o0al + int64_t(st % n_t) * TOBS;
    const int lo = int(o0 - base > 0 ? o0 - base : 0);             // first valid observation (0 or 1)
    const int hi = int(o1 - base < TOBS ? o1 - base : TOBS);       // one past the last valid observation
    double* Mt = gram_sm + buf * NR * RS;
    // warp w copies whole columns cc = w, w+3, ...: one address computation per column,
    // then 4 x (32 pairs) with immediate offsets
    for (int cc = warp; cc < 9; cc += 3) {
      // staged row cc <- stored plane 0..5 (host pose), 6.. (affine), PL - 1 (residual); the
      // inverse-distance plane is skipped, rows without a source plane stay zero
      if (cc >= 6 + NAFF && cc < 8) continue;
      const int sp = cc < 8 ? cc : PL - 1;
      const int c = cc;
      const double* src = JR + (int64_t(k) * PL + sp) * ld + base + 2 * lane;
      double* dst = Mt + c * RS + 2 * lane;
      const unsigned sa = unsigned(__cvta_generic_to_shared(dst));
#pragma unroll
      for (int j = 0; j < TOBS / 64; ++j) {
        const int a = j * 64 + 2 * lane;
        if (a >= lo && a + 1 < hi) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa + unsigned(j * 512)), "l"(src + j * 64));
        } else {
          if (a >= lo && a < hi) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa + unsigned(j * 512)), "l"(src + j * 64));
          } else {
            dst[j * 64] = 0.0;
          }
          if (a + 1 >= lo && a + 1 < hi) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa + unsigned(j * 512 + 8)), "l"(src + j * 64 + 1));
          } else {
            dst[j * 64 + 1] = 0.0;
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  // kGramStages-deep ring: stages st+1 .. st+kGramStages-1 are in flight while st is consumed
  for (int s0 = 0; s0 < kGramStages - 1; ++s0) {
    if (s0 < n_stage) stage(s0, s0);
    else asm volatile("cp.async.commit_group;\n" ::);
  }
  int buf = 0;
  for (int st = 0; st < n_stage; ++st) {
    const int nxt = st + kGramStages - 1;
    int nbuf = buf + kGramStages - 1;
    nbuf = nbuf >= kGramStages ? nbuf - kGramStages : nbuf;
    if (nxt < n_stage) stage(nxt, nbuf);
    else asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kGramStages - 1));
    __syncthreads();
    const int64_t base = o0al + int64_t(st % n_t) * TOBS;
    const int cnt = int(o1 - base < TOBS ? o1 - base : TOBS);
    const int nk = (cnt + 3) >> 2;  // k-steps of four observations (head and tail are zero-filled)
    const double* m0 = gram_sm + buf * NR * RS + fr * RS + fo;  // column block 0; block 1 = + 8 RS
    int ks = warp;
    if (RED) {
      // block 1 has one non-zero column (the residual, staged row 8): lanes of fragment row 0 carry it
      const double* m8 = gram_sm + buf * NR * RS + 8 * RS + fo;
      for (; ks + 3 < nk; ks += 6) {
        const double f0 = m0[4 * ks], g0 = m0[4 * ks + 12];
        const double r0 = m8[4 * ks], r1 = m8[4 * ks + 12];
        const double f1 = fr == 0 ? r0 : 0.0, g1 = fr == 0 ? r1 : 0.0;
        gram_dmma(c00[0], f0, f0); gram_dmma(c01[0], f0, f1);
        gram_dmma(c00[1], g0, g0); gram_dmma(c01[1], g0, g1);
      }
      if (ks < nk) {
        const double f0 = m0[4 * ks];
        const double r0 = m8[4 * ks];
        gram_dmma(c00[0], f0, f0); gram_dmma(c01[0], f0, fr == 0 ? r0 : 0.0);
      }
    } else {
      for (; ks + 3 < nk; ks += 6) {
        const double f0 = m0[4 * ks], f1 = m0[8 * RS + 4 * ks];
        const double g0 = m0[4 * ks + 12], g1 = m0[8 * RS + 4 * ks + 12];
        gram_dmma(c00[0], f0, f0); gram_dmma(c01[0], f0, f1); gram_dmma(c11[0], f1, f1);
        gram_dmma(c00[1], g0, g0); gram_dmma(c01[1], g0, g1); gram_dmma(c11[1], g1, g1);
      }
      if (ks < nk) {
        const double f0 = m0[4 * ks], f1 = m0[8 * RS + 4 * ks];
        gram_dmma(c00[0], f0, f0); gram_dmma(c01[0], f0, f1); gram_dmma(c11[0], f1, f1);
      }
    }
    __syncthreads();  // everyone is done with `buf` before a later stage overwrites it
    buf = buf + 1 == kGramStages ? 0 : buf + 1;
  }
  // sum the three warps' fragments: lane holds C[lane / 4][2 (lane % 4) + {0, 1}] of each tile
  {
    double* scratch = gram_sm;  // [3 warps][3 tiles][64]; the ring is dead after the last barrier
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int e = fr * 8 + 2 * fo + j;
      scratch[(warp * 3 + 0) * 64 + e] = c00[0][j] + c00[1][j];
      scratch[(warp * 3 + 1) * 64 + e] = c01[0][j] + c01[1][j];
      scratch[(warp * 3 + 2) * 64 + e] = c11[0][j] + c11[1][j];
    }
    __syncthreads();
    if (!RED) {
      for (int i = threadIdx.x; i < 192; i += 96) {
        const int t = i >> 6, e = i & 63, r = e >> 3, c = e & 7;
        const double v = scratch[t * 64 + e] + scratch[(3 + t) * 64 + e] + scratch[(6 + t) * 64 + e];
        const int ta = t == 2 ? 8 : 0, tb = t == 0 ? 0 : 8;
        G[(ta + r) * 16 + tb + c] = v;
        if (t == 1) G[(tb + c) * 16 + ta + r] = v;
      }
    } else {
      // G9 (9x9, symmetric; [8][8] = r^T r is not needed), E (9x16) and T1 = G9 E behind the scratch
      double* G9 = gram_sm + 9 * 64;
      double* E = G9 + 81;
      double* T1 = E + 9 * 16;
      for (int i = threadIdx.x; i < 81; i += 96) {
        const int r = i / 9, c = i % 9;
        double v = 0.0;
        if (r < 8 && c < 8) { const int e = r * 8 + c; v = scratch[e] + scratch[3 * 64 + e] + scratch[6 * 64 + e]; }
        else if (r < 8) { const int e = r * 8; v = scratch[64 + e] + scratch[4 * 64 + e] + scratch[7 * 64 + e]; }
        else if (c < 8) { const int e = c * 8; v = scratch[64 + e] + scratch[4 * 64 + e] + scratch[7 * 64 + e]; }
        G9[i] = v;
      }
      // columns of the full row: h(0..5) | t pose (6..11) = h x M | affine (12, 13) | rho (14, unused) | r (15)
      const double* Me = edge_M + 36 * int64_t(e);
      for (int i = threadIdx.x; i < 9 * 16; i += 96) {
        const int r = i >> 4, c = i & 15;
        double v = 0.0;
        if (c < 6) v = r == c ? 1.0 : 0.0;
        else if (c < 12) v = r < 6 ? Me[6 * r + (c - 6)] : 0.0;
        else if (c < 14) v = (c - 12 < NAFF && r == 6 + (c - 12)) ? 1.0 : 0.0;
        else if (c == 15) v = r == 8 ? 1.0 : 0.0;
        E[i] = v;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 9 * 16; i += 96) {
        const int r = i >> 4, c = i & 15;
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < 9; ++m) v += G9[r * 9 + m] * E[m * 16 + c];
        T1[i] = v;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 256; i += 96) {
        const int a = i >> 4, b = i & 15;
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < 9; ++m) v += E[m * 16 + a] * T1[m * 16 + b];
        G[i] = v;
      }
    }
  }
  __syncthreads();
  // partial layout: [HH | HT (or its transpose when slot_h > slot_t) | TT | gh | gt]
  const int hs = slot[edge_h[e]], ts = slot[edge_t[e]];
  const bool trans = hs > ts;
  double* out = part_dir + int64_t(q) * Cfg::DIR_STRIDE;
  for (int idx = threadIdx.x; idx < Cfg::DIR_STRIDE; idx += 96) {
    double v = 0.0;
    if (idx < CD * CD) {
      const int r = idx / CD, c = idx % CD;
      if (r < 6 && c < 6) v = G[r * 16 + c];
    } else if (idx < 2 * CD * CD) {
      const int r = (idx - CD * CD) / CD, c = (idx - CD * CD) % CD;
      if (!trans) { if (r < 6) v = G[r * 16 + 6 + c]; }
      else        { if (c < 6) v = G[c * 16 + 6 + r]; }
    } else if (idx < 3 * CD * CD) {
      const int r = (idx - 2 * CD * CD) / CD, c = (idx - 2 * CD * CD) % CD;
      v = G[(6 + r) * 16 + 6 + c];
    } else if (idx < 3 * CD * CD + CD) {
      const int r = idx - 3 * CD * CD;
      if (r < 6) v = G[r * 16 + 15];
    } else {
      const int r = idx - 3 * CD * CD - CD;
      v = G[(6 + r) * 16 + 15];
    }
    out[idx] = v;
  }
}

// -------------------------------------------------------- landmark gather --
// Per landmark: W row = [ F^T E per visible camera (8 wide each) | g_l c_l 0.. ],
// c_l = sum E^T E, g_l = sum E^T r over the landmark's observations.  A half-warp
// per landmark: lane q owns element q of the 16-double Schur record, so each
// record is one coalesced 128 B read and each camera slot one 64 B write.  Slots
// of cameras that do not see the landmark are zeroed once in pba_create and never
// written (the structure is static).
__global__ void __launch_bounds__(128) k_lm_gather(int n_lm, int cd, const int64_t* __restrict__ lm_ptr,
                                                    const int64_t* __restrict__ lm_pos,
                                                    const int* __restrict__ obs_col,
                                                    const int* __restrict__ lm_hostcol,
                                                    const int64_t* __restrict__ lm_w_off,
                                                    const int* __restrict__ lm_w_stride,
                                                    const double* __restrict__ orec, double* __restrict__ W,
                                                    double* __restrict__ lm_c, double* __restrict__ lm_g) {
  const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int q = threadIdx.x & 15;
  if (l >= n_lm) return;
  double* row = W + lm_w_off[l];
  const int stride = lm_w_stride[l];
  double acc = 0.0;  // lanes 0-5: host columns, 14: c_l, 15: g_l
  // observations four at a time: the position loads, then the record loads, are independent
  // (one dependent pair per observation kept the kernel latency-bound at 55 % of HBM)
  const int64_t k1 = lm_ptr[l + 1];
  for (int64_t k = lm_ptr[l]; k < k1; k += 4) {
    int64_t pos[4];
    double v[4];
    int col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pos[j] = k + j < k1 ? lm_pos[k + j] : -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = pos[j] >= 0 ? orec[16 * pos[j] + q] : 0.0;
      col[j] = pos[j] >= 0 ? obs_col[pos[j]] : -1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (pos[j] < 0) continue;
      if (q >= 6 && q < 6 + cd) {
        if (col[j] >= 0) row[8 * col[j] + (q - 6)] = v[j];
      } else {
        acc += v[j];
      }
    }
  }
  const int hc = lm_hostcol[l];
  if (q < 6) {
    if (hc >= 0) row[8 * hc + q] = acc;
  } else if (q == 14) {
    row[stride - 7] = acc;
    lm_c[l] = acc;
  } else if (q == 15) {
    row[stride - 8] = acc;
    lm_g[l] = acc;
  }
}

// Per landmark, per linear solve: Jacobi scale sigma_l (iteration 0 only),
// LM diagonal (refreshed after accepted steps only), ete^-1 and sigma^2/ete
// (trust_region_minimizer.cc:261-272, levenberg_marquardt_strategy.cc:76-88,
// schur_eliminator_impl.h:246-252).
__global__ void k_lm_scale(int n_lm, int init_scale, int jacobi, int refresh_diag, double radius, double min_diag,
                           double max_diag, const double* __restrict__ lm_c, double* __restrict__ lm_scale,
                           double* __restrict__ lm_diag, double* __restrict__ lm_iete, double* __restrict__ lm_s2) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  const double c = lm_c[l];
  if (init_scale) lm_scale[l] = jacobi ? 1.0 / (1.0 + sqrt(c)) : 1.0;
  const double s = lm_scale[l];
  const double cs = c * s * s;
  if (refresh_diag) lm_diag[l] = fmin(fmax(cs, min_diag), max_diag);
  const double ete = cs + lm_diag[l] / radius;
  const double ie = 1.0 / ete;
  lm_iete[l] = ie;
  lm_s2[l] = s * s * ie;
}

// ------------------------------------------------------------ Schur SYRK ---
// One CTA per host group: partial = W^T diag(s2) W over the group's landmarks,
// as cd x cd blocks for camera pairs (i <= j) plus the vectors W^T (s2 g).
// Landmark rows are staged in shared memory 'tile_l' at a time.  The products
// run on the FP64 tensor cores: a k-step is four landmarks, an 8x8 output tile
// one camera pair (or camera x the (g_l, c_l) block for the vectors), so a
// tile-step costs 2 LDS.64 + DMUL + DMMA per lane (the scalar 4x4 register
// tiles needed 8 LDS.64 per 16 DFMA and ran at a third of the FP64 peak).
// Warp w owns tiles w, w+8, ... (<= kSyrkTiles per pass) with static accumulators.
constexpr int kSyrkTiles = 12;
__host__ __device__ inline int syrk_row_stride(int stride) { return stride + 4; }  // = 4 or 12 (mod 16): conflict-free fragments

// One CTA per PASS of a group (96 tiles; `work` = (group, first tile) pairs, schur_syrk_work): with one CTA per
// group, a real map's few large host groups — the first keyframes host most landmarks and see a hundred cameras:
// 5,000 tiles = 54 passes over 500 landmark rows — ran on one SM each while the rest of the GPU idled (EuRoC map:
// 0.91 ms of a 2.2 ms LM step).
__global__ void __launch_bounds__(256) k_schur_syrk(int cd, int tile_l, const int* __restrict__ work,
                                                     const int* __restrict__ grp_lm_ptr,
                                                     const int* __restrict__ grp_cam_ptr,
                                                     const int64_t* __restrict__ grp_w_off,
                                                     const int64_t* __restrict__ grp_part_off,
                                                     const double* __restrict__ W, const double* __restrict__ lm_s2,
                                                     double* __restrict__ part_sch) {
  extern __shared__ double sm[];
  const int g = work[2 * blockIdx.x];
  const int l0 = grp_lm_ptr[g], l1 = grp_lm_ptr[g + 1];
  const int c = grp_cam_ptr[g + 1] - grp_cam_ptr[g];
  if (c == 0) return;
  const int stride = 8 * (c + 1);
  const int ss = syrk_row_stride(stride);
  const double* Wg = W + grp_w_off[g];
  double* out = part_sch + grp_part_off[g];
  const int P = c * (c + 1) / 2;
  const int n_tiles = P + c;  // camera pairs (i <= j), then camera i x the (g_l, c_l, 0..) block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fo = lane & 3;  // fragment coordinates: column within the block, landmark of the k-step

  {
    const int t0 = work[2 * blockIdx.x + 1];
    int oi[kSyrkTiles], oj[kSyrkTiles];
    double acc[kSyrkTiles][2];
#pragma unroll
    for (int tt = 0; tt < kSyrkTiles; ++tt) {
      const int t = t0 + warp + 8 * tt;
      oi[tt] = -1; oj[tt] = 0;
      if (t < n_tiles) {
        int i, j;
        if (t < P) {
          int p = t; i = 0;
          while (p >= c - i) { p -= c - i; ++i; }
          j = i + p;
        } else {
          i = t - P; j = c;
        }
        oi[tt] = 8 * i + fr; oj[tt] = 8 * j + fr;
      }
      acc[tt][0] = 0.0; acc[tt][1] = 0.0;
    }
    // double-buffered cp.async staging: tile lb + tile_l streams in while tile lb is multiplied
    const int buf_doubles = tile_l * ss + tile_l;
    auto stage = [&](int lb, int buf) {
      const int cnt = l1 - lb < tile_l ? l1 - lb : tile_l;
      const double* src = Wg + size_t(lb - l0) * stride;
      double* dst = sm + buf * buf_doubles;
      const int cpr = stride >> 1;  // 16-byte chunks per row (rows are 64-byte aligned on both sides)
      for (int i = threadIdx.x; i < cnt * cpr; i += 256) {
        const int l = i / cpr, x = i - l * cpr;
        const unsigned da = unsigned(__cvta_generic_to_shared(dst + l * ss + 2 * x));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(da), "l"(src + size_t(l) * stride + 2 * x));
      }
      for (int i = threadIdx.x; i < cnt; i += 256) {
        const unsigned da = unsigned(__cvta_generic_to_shared(dst + tile_l * ss + i));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(da), "l"(lm_s2 + lb + i));
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    __syncthreads();  // the previous pass is done with both buffers
    stage(l0, 0);
    int buf = 0;
    for (int lb = l0; lb < l1; lb += tile_l, buf ^= 1) {
      const int cnt = l1 - lb < tile_l ? l1 - lb : tile_l;
      if (lb + tile_l < l1) {
        stage(lb + tile_l, buf ^ 1);
        asm volatile("cp.async.wait_group 1;\n" ::);
      } else {
        asm volatile("cp.async.wait_group 0;\n" ::);
      }
      __syncthreads();
      const double* tile = sm + buf * buf_doubles;
      const double* s_s2 = tile + tile_l * ss;
      for (int k0 = 0; k0 < cnt; k0 += 4) {
        const int l = k0 + fo;
        const bool lv = l < cnt;
        const double* row = tile + l * ss;
        const double s2 = lv ? s_s2[l] : 0.0;
#pragma unroll
        for (int tt = 0; tt < kSyrkTiles; ++tt) {
          if (oi[tt] < 0) continue;  // warp-uniform
          const double a = lv ? s2 * row[oi[tt]] : 0.0;
          const double bq = lv ? row[oj[tt]] : 0.0;
          gram_dmma(acc[tt], a, bq);
        }
      }
      __syncthreads();  // everyone is done with `buf` before the next pass restages it
    }
    // lane holds C[fr][2 fo + {0, 1}] of every tile
#pragma unroll
    for (int tt = 0; tt < kSyrkTiles; ++tt) {
      const int t = t0 + warp + 8 * tt;
      if (t >= n_tiles) continue;
      if (t < P) {
        double* blk = out + size_t(t) * cd * cd;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          if (fr < cd && 2 * fo + jj < cd) blk[fr * cd + 2 * fo + jj] = acc[tt][jj];
      } else {
        double* vec = out + size_t(P) * cd * cd + size_t(t - P) * cd;
        if (fo == 0 && fr < cd) vec[fr] = acc[tt][0];
      }
    }
  }
}

// Row-paired variant for groups of at most kSyrkRowsMaxC cameras (windowed covisibility).  The tile
// rows i and c-1-i of the upper triangle (c + 3 tiles together, vector tile included) form a pair;
// TWO warps share a pair, one taking its even and one its odd tiles, so a group of 12 cameras keeps
// 12 warps busy with ~8 accumulator tiles each (one warp per pair: 6 of 8 warps busy, 121 registers,
// DMMA pipe 46 % busy with `barrier` / `short_scoreboard` on top of the stall list).  A warp keeps the
// two scaled A fragments of a k-step in registers and loads only the B fragment per tile: 1 LDS.64
// per DMMA instead of 2, which is what bounds the generic kernel above.
constexpr int kSyrkRowsMaxC = 12;
constexpr int kSyrkRowsThreads = 32 * 2 * ((kSyrkRowsMaxC + 1) / 2);                  // 384
constexpr int kSyrkRowTiles = (kSyrkRowsMaxC + 1 + 1) / 2;      // half of the longest row (row 0: j = 0..c)
constexpr int kSyrkRowTilesB = (kSyrkRowsMaxC / 2 + 2 + 1) / 2; // half of the second row (ib >= c / 2: at most c / 2 + 2 tiles)

__global__ void __launch_bounds__(kSyrkRowsThreads) k_schur_syrk_rows(int cd, int tile_l, const int* __restrict__ grp_lm_ptr,
                                                                       const int* __restrict__ grp_cam_ptr,
                                                                       const int64_t* __restrict__ grp_w_off,
                                                                       const int64_t* __restrict__ grp_part_off,
                                                                       const double* __restrict__ W,
                                                                       const double* __restrict__ lm_s2,
                                                                       double* __restrict__ part_sch) {
  extern __shared__ double sm[];
  const int g = blockIdx.x;
  const int l0 = grp_lm_ptr[g], l1 = grp_lm_ptr[g + 1];
  const int c = grp_cam_ptr[g + 1] - grp_cam_ptr[g];
  if (c == 0) return;
  const int stride = 8 * (c + 1);
  const int ss = syrk_row_stride(stride);
  const double* Wg = W + grp_w_off[g];
  double* out = part_sch + grp_part_off[g];
  const int P = c * (c + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fr = lane >> 2, fo = lane & 3;
  // pair pw: rows ia = pw (tiles j = ia..c) and ib = c - 1 - pw (tiles j = ib..c) when distinct;
  // this warp takes the tiles t of each row with t % 2 == half
  const int pw = warp >> 1, half = warp & 1;
  const int ia = pw, ib = c - 1 - pw;
  const bool has_a = ia < c && ia <= ib, has_b = ib > ia;
  double acca[kSyrkRowTiles][2], accb[kSyrkRowTilesB][2];
#pragma unroll
  for (int t = 0; t < kSyrkRowTiles; ++t) acca[t][0] = acca[t][1] = 0.0;
#pragma unroll
  for (int t = 0; t < kSyrkRowTilesB; ++t) accb[t][0] = accb[t][1] = 0.0;
  // tiles of row a / b owned by this warp: t = half, half + 2, ...
  const int na_full = has_a ? c - ia + 1 : 0, nb_full = has_b ? c - ib + 1 : 0;
  const int na = (na_full - half + 1) / 2, nb = (nb_full - half + 1) / 2;

  const int buf_doubles = tile_l * ss + tile_l;
  auto stage = [&](int lb, int buf) {
    const int cnt = l1 - lb < tile_l ? l1 - lb : tile_l;
    const double* src = Wg + size_t(lb - l0) * stride;
    double* dst = sm + buf * buf_doubles;
    const int cpr = stride >> 1;
    for (int i = threadIdx.x; i < cnt * cpr; i += kSyrkRowsThreads) {
      const int l = i / cpr, x = i - l * cpr;
      const unsigned da = unsigned(__cvta_generic_to_shared(dst + l * ss + 2 * x));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(da), "l"(src + size_t(l) * stride + 2 * x));
    }
    for (int i = threadIdx.x; i < cnt; i += kSyrkRowsThreads) {
      const unsigned da = unsigned(__cvta_generic_to_shared(dst + tile_l * ss + i));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(da), "l"(lm_s2 + lb + i));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  stage(l0, 0);
  int buf = 0;
  for (int lb = l0; lb < l1; lb += tile_l, buf ^= 1) {
    const int cnt = l1 - lb < tile_l ? l1 - lb : tile_l;
    if (lb + tile_l < l1) {
      stage(lb + tile_l, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    const double* tile = sm + buf * buf_doubles;
    const double* s_s2 = tile + tile_l * ss;
    if (has_a) {
      for (int k0 = 0; k0 < cnt; k0 += 4) {
        const int l = k0 + fo;
        const bool lv = l < cnt;
        const double* row = tile + l * ss + fr;
        const double s2 = lv ? s_s2[l] : 0.0;
        const double fa = lv ? s2 * row[8 * ia] : 0.0;
        const double fb = (lv && has_b) ? s2 * row[8 * ib] : 0.0;
        const double* ra = row + 8 * (ia + half);
        const double* rb = row + 8 * (ib + half);
#pragma unroll
        for (int t = 0; t < kSyrkRowTiles; ++t) {
          if (t < na) gram_dmma(acca[t], fa, lv ? ra[16 * t] : 0.0);
          if (t < kSyrkRowTilesB && t < nb) gram_dmma(accb[t < kSyrkRowTilesB ? t : 0], fb, lv ? rb[16 * t] : 0.0);
        }
      }
    }
    __syncthreads();
  }
  // lane holds C[fr][2 fo + {0, 1}] of every tile; pair (i, j) is block i c - i (i - 1) / 2 + (j - i)
  auto store_row = [&](int i, int n, auto& acc, auto ntile) {
    const int p0 = i * c - i * (i - 1) / 2;
#pragma unroll
    for (int t = 0; t < decltype(ntile)::value; ++t) {
      if (t >= n) continue;
      const int jt = half + 2 * t;  // tile index within the row: column block i + jt
      if (i + jt < c) {
        double* blk = out + size_t(p0 + jt) * cd * cd;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          if (fr < cd && 2 * fo + jj < cd) blk[fr * cd + 2 * fo + jj] = acc[t][jj];
      } else {
        double* vec = out + size_t(P) * cd * cd + size_t(i) * cd;
        if (fo == 0 && fr < cd) vec[fr] = acc[t][0];
      }
    }
  };
  if (has_a) store_row(ia, na, acca, std::integral_constant<int, kSyrkRowTiles>());
  if (has_b) store_row(ib, nb, accb, std::integral_constant<int, kSyrkRowTilesB>());
}

// ------------------------------------------------------------ RCS reduce ---
// One CTA per RCS block (cd*cd threads): S = sum(direct) - sum(schur).  The
// diagonal of the direct part (= squared camera column norms of J) is kept for
// Jacobi scaling and the LM diagonal.  Extra CTAs handle the right-hand side.
__global__ void k_rcs_reduce(int cd, int64_t n_blocks, int n_slots, const int64_t* __restrict__ dir_ptr,
                             const int64_t* __restrict__ dir_src, const int64_t* __restrict__ sch_ptr,
                             const int64_t* __restrict__ sch_src, const int64_t* __restrict__ vdir_ptr,
                             const int64_t* __restrict__ vdir_src, const int64_t* __restrict__ vsch_ptr,
                             const int64_t* __restrict__ vsch_src, const int* __restrict__ blk_row,
                             const int* __restrict__ blk_col, const double* __restrict__ part_dir,
                             const double* __restrict__ part_sch, double* __restrict__ S, double* __restrict__ rhs,
                             double* __restrict__ diagB, double* __restrict__ gcam, double* __restrict__ B_local,
                             double* __restrict__ g_local) {
  const int64_t b = blockIdx.x;
  const int e = threadIdx.x;
  if (b < n_blocks) {
    if (e >= cd * cd) return;
    double d = 0.0, s = 0.0;
    for (int64_t k = dir_ptr[b]; k < dir_ptr[b + 1]; ++k) d += part_dir[dir_src[k] + e];
    for (int64_t k = sch_ptr[b]; k < sch_ptr[b + 1]; ++k) s += part_sch[sch_src[k] + e];
    S[b * cd * cd + e] = d - s;
    B_local[b * cd * cd + e] = d;
    if (blk_row[b] == blk_col[b] && e / cd == e % cd) diagB[blk_row[b] * cd + e / cd] = d;
  } else {
    const int a = int(b - n_blocks);
    if (a >= n_slots || e >= cd) return;
    double d = 0.0, s = 0.0;
    for (int64_t k = vdir_ptr[a]; k < vdir_ptr[a + 1]; ++k) d += part_dir[vdir_src[k] + e];
    for (int64_t k = vsch_ptr[a]; k < vsch_ptr[a + 1]; ++k) s += part_sch[vsch_src[k] + e];
    rhs[a * cd + e] = d - s;
    gcam[a * cd + e] = d;
    g_local[a * cd + e] = d;
  }
}

// Camera column scales / LM diagonal (same rules as k_lm_scale).
__global__ void k_cam_scale(int dim, int init_scale, int jacobi, int refresh_diag, double radius, double min_diag,
                            double max_diag, const double* __restrict__ diagB, double* __restrict__ cam_scale,
                            double* __restrict__ cam_diag, double* __restrict__ cam_D2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dim) return;
  const double c = diagB[i];
  if (init_scale) cam_scale[i] = jacobi ? 1.0 / (1.0 + sqrt(c)) : 1.0;
  const double s = cam_scale[i];
  if (refresh_diag) cam_diag[i] = fmin(fmax(c * s * s, min_diag), max_diag);
  cam_D2[i] = cam_diag[i] / radius;
}

// S_ab <- diag(sigma_a) S_ab diag(sigma_b) (+ D^2 on the diagonal); rhs <- sigma rhs.
__global__ void k_rcs_scale(int cd, int64_t n_blocks, int n_slots, const int* __restrict__ blk_row,
                            const int* __restrict__ blk_col, const double* __restrict__ cam_scale,
                            const double* __restrict__ cam_D2, double* __restrict__ S, double* __restrict__ rhs) {
  const int64_t b = blockIdx.x;
  const int e = threadIdx.x;
  if (b < n_blocks) {
    if (e >= cd * cd) return;
    const int r = e / cd, c = e % cd;
    const int a = blk_row[b], bb = blk_col[b];
    double v = S[b * cd * cd + e] * cam_scale[a * cd + r] * cam_scale[bb * cd + c];
    if (a == bb && r == c) v += cam_D2[a * cd + r];
    S[b * cd * cd + e] = v;
  } else {
    const int a = int(b - n_blocks);
    if (a >= n_slots || e >= cd) return;
    rhs[a * cd + e] *= cam_scale[a * cd + e];
  }
}

// ------------------------------------------------------------- back-subst --
// y_l = ete^-1 sigma_l (g_l - sum_a w_la . (sigma_a y_a))  (schur_eliminator_impl.h:309-375)
// then the unscaled tangent step d_l = -sigma_l y_l.
//
// The same pass yields the landmark part of the model cost change
//   -(J d)^T (r + J d / 2) = -d^T g - d^T H d / 2,   H = J^T J, g = J^T r (unscaled)
// (trust_region_minimizer.cc:414-427) without touching J again:
//   landmark l:  -d_l (g_l + t_l + c_l d_l / 2),  t_l = sum_a w_la . d_a
// the camera part (-d_c.g_c - d_c^T B d_c / 2) comes from k_model_cost_cam.
// A half-warp per landmark (lanes stride the landmark's row of W: coalesced, all loads
// independent), 128 landmarks per CTA so the model-cost partials keep one slot per 128 landmarks.
__global__ void __launch_bounds__(256) k_backsub(int n_lm, int cd, const int* __restrict__ lm_group,
                                                  const int* __restrict__ grp_cam_ptr, const int* __restrict__ grp_cams,
                                                  const int64_t* __restrict__ lm_w_off,
                                                  const int* __restrict__ lm_w_stride, const int64_t* __restrict__ lm_ptr,
                                                  const double* __restrict__ W, const double* __restrict__ lm_scale,
                                                  const double* __restrict__ lm_iete, const double* __restrict__ d_cam,
                                                  double* __restrict__ d_rho, double* __restrict__ part_model) {
  __shared__ double sm[8];
  const int hw = threadIdx.x >> 4, q = threadIdx.x & 15;
  double mc = 0.0;
  for (int it = 0; it < 8; ++it) {
    const int l = blockIdx.x * 128 + it * 16 + hw;
    const bool live = l < n_lm && lm_ptr[l + 1] > lm_ptr[l];
    const double* row = nullptr;
    int stride = 0;
    // d_cam = -sigma y  =>  sum_a w_la . (sigma_a y_a) = -sum_a w_la . d_a
    double t = 0.0;
    if (live) {
      const int g = lm_group[l];
      const int c0 = grp_cam_ptr[g], c = grp_cam_ptr[g + 1] - c0;
      row = W + lm_w_off[l];
      stride = lm_w_stride[l];
#pragma unroll 4
      for (int e = q; e < 8 * c; e += 16) {
        const int j = e >> 3, i = e & 7;
        if (i < cd) t += row[e] * d_cam[grp_cams[c0 + j] * cd + i];
      }
    }
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (q == 0 && l < n_lm) {
      if (!live) {
        d_rho[l] = 0.0;
      } else {
        const double gl = row[stride - 8], cl = row[stride - 7];
        const double s = lm_scale[l];
        const double y = lm_iete[l] * s * (gl + t);
        const double dl = -s * y;
        d_rho[l] = dl;
        mc += -dl * (gl + t + 0.5 * cl * dl);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) mc += __shfl_down_sync(0xffffffffu, mc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += sm[i];
    part_model[blockIdx.x] = v;
  }
}

// The same computation as a STREAMING mat-vec (the default whenever every host group has at most 31 cameras):
// the rows of W are contiguous per landmark and all landmarks of a host group multiply the same vector
// (d_cam of the group's cameras), so
//   * a warp owns 32 consecutive landmarks; lane j loads landmark j's descriptors once (coalesced),
//   * the group's vector sits in registers, laid out like the row (lane q holds elements 2q, 2q+1 (+64 per step))
//     and is reloaded only when the group changes (~1,000 landmarks per group),
//   * rows are read as 16-byte loads, four landmarks (4 x S independent loads per lane) in flight.
// k_backsub above chases lm_ptr -> lm_group -> grp_cam_ptr -> grp_cams -> d_cam per landmark and element; ncu
// (profiles/r02a_backsub_details.txt) showed it at 41 % of the DRAM peak, 77 % of the warp cycles waiting on
// L1TEX scoreboards.  S = double2 steps per row: row length <= 64 S doubles.
template <int S>
__global__ void __launch_bounds__(128, S <= 2 ? 4 : 2) k_backsub_stream(int n_lm, int cd, const int* __restrict__ lm_group,
                                                         const int* __restrict__ grp_cam_ptr,
                                                         const int* __restrict__ grp_cams,
                                                         const int64_t* __restrict__ lm_w_off,
                                                         const int* __restrict__ lm_w_stride,
                                                         const int64_t* __restrict__ lm_ptr, const double* __restrict__ W,
                                                         const double* __restrict__ lm_scale,
                                                         const double* __restrict__ lm_iete,
                                                         const double* __restrict__ d_cam, double* __restrict__ d_rho,
                                                         double* __restrict__ part_model) {
  __shared__ double sm[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l0 = blockIdx.x * 128 + warp * 32;
  const int my_l = l0 + lane;
  const bool my_live = my_l < n_lm && lm_ptr[my_l + 1] > lm_ptr[my_l];
  const int my_g = my_live ? lm_group[my_l] : -1;
  const int64_t my_off = my_live ? lm_w_off[my_l] : 0;
  const int my_stride = my_live ? lm_w_stride[my_l] : 0;
  const double my_scale = my_live ? lm_scale[my_l] : 0.0, my_iete = my_live ? lm_iete[my_l] : 0.0;
  double mc = 0.0;
  int cur_g = -1, cur_c = 0;
  (void)cur_c;
  double2 dv[S];
#pragma unroll
  for (int s = 0; s < S; ++s) dv[s] = make_double2(0.0, 0.0);
  // Software pipeline over the eight batches of four landmarks: batch b + 1's loads are issued before batch b's
  // arithmetic (the group-change branch in the arithmetic keeps the compiler from hoisting them itself).
  struct Batch { double2 v[4][S], gc[4]; int g[4]; };  // gc (lane 0): [g_l, c_l] = the first two of the row's last eight
  auto load_batch = [&](int j0, Batch& B) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      B.g[u] = __shfl_sync(0xffffffffu, my_g, j0 + u);
      const int64_t off = __shfl_sync(0xffffffffu, my_off, j0 + u);
      const int st = __shfl_sync(0xffffffffu, my_stride, j0 + u);
      const double2* row = reinterpret_cast<const double2*>(W + off);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int e2 = lane + 32 * s;
        B.v[u][s] = (B.g[u] >= 0 && 2 * e2 < st) ? __ldcs(row + e2) : make_double2(0.0, 0.0);  // streamed once
      }
      B.gc[u] = (B.g[u] >= 0 && lane == 0) ? __ldcs(row + ((st - 8) >> 1)) : make_double2(0.0, 0.0);
    }
  };
  Batch cur;
  load_batch(0, cur);
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 4) {
    Batch nxt;
    if (j0 + 4 < 32) load_batch(j0 + 4, nxt);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + j0 + u;
      if (cur.g[u] < 0) {  // warp-uniform: no observations (or beyond the last landmark)
        if (lane == 0 && l < n_lm) d_rho[l] = 0.0;
      } else {
        if (cur.g[u] != cur_g) {
          cur_g = cur.g[u];
          const int c0 = grp_cam_ptr[cur_g];
          cur_c = grp_cam_ptr[cur_g + 1] - c0;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int e = 2 * (lane + 32 * s);
            const int j = e >> 3, i = e & 7;  // i is even: i and i + 1 belong to the same camera
            double a = 0.0, b = 0.0;
            if (j < cur_c) {
              const double* d = d_cam + grp_cams[c0 + j] * cd;
              if (i < cd) a = d[i];
              if (i + 1 < cd) b = d[i + 1];
            }
            dv[s] = make_double2(a, b);
          }
        }
        // d_cam = -sigma y  =>  sum_a w_la . (sigma_a y_a) = -sum_a w_la . d_a
        double t = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) t = fma(cur.v[u][s].x, dv[s].x, fma(cur.v[u][s].y, dv[s].y, t));
        const double gl = cur.gc[u].x, cl = cur.gc[u].y;  // lane 0 only
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const double sc = __shfl_sync(0xffffffffu, my_scale, j0 + u), ie = __shfl_sync(0xffffffffu, my_iete, j0 + u);
        if (lane == 0) {
          const double y = ie * sc * (gl + t);
          const double dl = -sc * y;
          d_rho[l] = dl;
          mc += -dl * (gl + t + 0.5 * cl * dl);
        }
      }
    }
    if (j0 + 4 < 32) cur = nxt;
  }
  if (lane == 0) sm[warp] = mc;
  __syncthreads();
  if (threadIdx.x == 0) {
    part_model[blockIdx.x] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
  }
}

// Camera part of the model cost change: EIGHT lanes per RCS block of the raw (unscaled, undamped, this rank's)
// direct part B (lane i takes row i: consecutive lanes read consecutive rows, the block is one contiguous run):
// -d_a^T B_ab d_b (x1/2 on the diagonal), plus -d_a . g_a from the diagonal block.  One thread per block read its
// 64 doubles alone, 512 bytes apart from its neighbour's: 17 us per launch for 24 k blocks.
__global__ void __launch_bounds__(128) k_model_cost_cam(int cd, int64_t n_blocks, const int* __restrict__ blk_row,
                                                         const int* __restrict__ blk_col,
                                                         const double* __restrict__ B_local,
                                                         const double* __restrict__ g_local,
                                                         const double* __restrict__ d_cam, double* __restrict__ part) {
  __shared__ double sm[4];
  const int64_t b = int64_t(blockIdx.x) * 16 + (threadIdx.x >> 3);
  const int i = threadIdx.x & 7;
  double mc = 0.0;
  if (b < n_blocks && i < cd) {
    const int a = blk_row[b], c = blk_col[b];
    const double* Bb = B_local + b * cd * cd + i * cd;
    const double* dc = d_cam + c * cd;
    double s = 0.0;
    for (int j = 0; j < cd; ++j) s += Bb[j] * dc[j];
    const double da = d_cam[a * cd + i];
    const double q = da * s;
    mc = a == c ? -da * g_local[a * cd + i] - 0.5 * q : -q;
  }
  for (int o = 16; o > 0; o >>= 1) mc += __shfl_down_sync(0xffffffffu, mc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mc;
  __syncthreads();
  if (threadIdx.x == 0) part[blockIdx.x] = sm[0] + sm[1] + sm[2] + sm[3];
}

__global__ void k_cam_step(int dim, const double* __restrict__ y, const double* __restrict__ scale,
                           double* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) d[i] = -y[i] * scale[i];
}

__device__ __forceinline__ double block_sum256(double v, double* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < int(blockDim.x >> 5); ++i) s += smem[i];
  __syncthreads();
  return s;
}
__device__ __forceinline__ double block_max256(double v, double* smem) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < int(blockDim.x >> 5); ++i) s = fmax(s, smem[i]);
  __syncthreads();
  return s;
}

// ------------------------------------------------------------ retraction ---
// candidate = Plus(x, delta): T exp(d) for poses, + for affine / rho.  Items
// [0, n_poses) are poses, [n_poses, n_poses + n_lm) landmarks.  Per-block
// partials: [0] ||x - cand||^2, [1] ||cand||^2 over the reduced program's
// ambient parameters (trust_region_minimizer.cc:706-726, :813-814).
// `own_poses` = this rank accounts for the (replicated) pose terms.
__global__ void __launch_bounds__(256) k_retract(int n_poses, int n_lm, int cd, int own_poses,
                                                  const int* __restrict__ slot, const uint8_t* __restrict__ aff_active,
                                                  const int64_t* __restrict__ lm_ptr, const double* __restrict__ poses,
                                                  const double* __restrict__ affine, const double* __restrict__ rho,
                                                  const double* __restrict__ d_cam, const double* __restrict__ d_rho,
                                                  double* __restrict__ poses_c, double* __restrict__ affine_c,
                                                  double* __restrict__ rho_c, double* __restrict__ part_step,
                                                  double* __restrict__ part_norm) {
  __shared__ double sm[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double st = 0.0, nm = 0.0;
  if (i < n_poses) {
    const int s = slot[i];
    double T[7], O[7];
    for (int k = 0; k < 7; ++k) T[k] = poses[7 * i + k];
    if (s >= 0) {
      se3_plus(T, d_cam + s * cd, O);
      if (own_poses)
        for (int k = 0; k < 7; ++k) { const double d = T[k] - O[k]; st += d * d; nm += O[k] * O[k]; }
    } else {
      for (int k = 0; k < 7; ++k) O[k] = T[k];
    }
    for (int k = 0; k < 7; ++k) poses_c[7 * i + k] = O[k];
    if (affine) {
      double a = affine[2 * i], b = affine[2 * i + 1];
      if (s >= 0 && aff_active[i]) {
        const double na = a + d_cam[s * cd + 6], nb = b + d_cam[s * cd + 7];
        if (own_poses) { st += (na - a) * (na - a) + (nb - b) * (nb - b); nm += na * na + nb * nb; }
        a = na; b = nb;
      }
      affine_c[2 * i] = a; affine_c[2 * i + 1] = b;
    }
  } else if (i < n_poses + n_lm) {
    const int l = i - n_poses;
    double r = rho[l];
    if (lm_ptr[l + 1] > lm_ptr[l]) {
      const double nr = r + d_rho[l];
      st += (nr - r) * (nr - r);
      nm += nr * nr;
      r = nr;
    }
    rho_c[l] = r;
  }
  const double a = block_sum256(st, sm);
  const double b = block_sum256(nm, sm);
  if (threadIdx.x == 0) { part_step[blockIdx.x] = a; part_norm[blockIdx.x] = b; }
}

// Gradient norms the way Ceres reports them: ||x - Plus(x, -g)|| in max and
// 2-norm (trust_region_minimizer.cc:279-298), g = J^T r of the unscaled J.
__global__ void __launch_bounds__(256) k_grad_norms(int n_poses, int n_lm, int cd, int own_poses,
                                                     const int* __restrict__ slot,
                                                     const uint8_t* __restrict__ aff_active,
                                                     const int64_t* __restrict__ lm_ptr,
                                                     const double* __restrict__ poses, const double* __restrict__ gcam,
                                                     const double* __restrict__ lm_g, double* __restrict__ part_max,
                                                     double* __restrict__ part_sq) {
  __shared__ double sm[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double mx = 0.0, sq = 0.0;
  if (i < n_poses) {
    const int s = slot[i];
    if (s >= 0 && own_poses) {
      double T[7], O[7], d[6];
      for (int k = 0; k < 7; ++k) T[k] = poses[7 * i + k];
      for (int k = 0; k < 6; ++k) d[k] = -gcam[s * cd + k];
      se3_plus(T, d, O);
      for (int k = 0; k < 7; ++k) { const double e = fabs(T[k] - O[k]); mx = fmax(mx, e); sq += e * e; }
      if (cd == 8 && aff_active[i])
        for (int k = 6; k < 8; ++k) { const double e = fabs(gcam[s * cd + k]); mx = fmax(mx, e); sq += e * e; }
    }
  } else if (i < n_poses + n_lm) {
    const int l = i - n_poses;
    if (lm_ptr[l + 1] > lm_ptr[l]) { const double e = fabs(lm_g[l]); mx = e; sq = e * e; }
  }
  const double a = block_max256(mx, sm);
  const double b = block_sum256(sq, sm);
  if (threadIdx.x == 0) { part_max[blockIdx.x] = a; part_sq[blockIdx.x] = b; }
}

__global__ void __launch_bounds__(1024) k_reduce2(const double* __restrict__ pa, const double* __restrict__ pb, int64_t n,
                                                   int a_is_max, double* __restrict__ oa, double* __restrict__ ob) {
  __shared__ double s[32], t[32];
  double va = 0.0, vb = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    va = a_is_max ? fmax(va, pa[i]) : va + pa[i];
    vb += pb[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double x = __shfl_down_sync(0xffffffffu, va, o);
    va = a_is_max ? fmax(va, x) : va + x;
    vb += __shfl_down_sync(0xffffffffu, vb, o);
  }
  if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5] = va; t[threadIdx.x >> 5] = vb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0;
    for (int i = 0; i < 32; ++i) { x = a_is_max ? fmax(x, s[i]) : x + s[i]; y += t[i]; }
    *oa = x; *ob = y;
  }
}

// world > 1: S_GMAX / S_GNORM2 / S_COST from the pose partials (every rank computes them from the all-reduced
// camera gradient) and the all-reduced tail [cost | sum g_l^2 | max |g_l| per rank].
__global__ void __launch_bounds__(1024) k_finish_norms(const double* __restrict__ pmax, const double* __restrict__ psq, int64_t n,
                                                        const double* __restrict__ tail, int world,
                                                        double* __restrict__ scalars) {
  __shared__ double s[32], t[32];
  double va = 0.0, vb = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) { va = fmax(va, pmax[i]); vb += psq[i]; }
  for (int i = threadIdx.x; i < world; i += 1024) va = fmax(va, tail[2 + i]);
  for (int o = 16; o > 0; o >>= 1) {
    va = fmax(va, __shfl_down_sync(0xffffffffu, va, o));
    vb += __shfl_down_sync(0xffffffffu, vb, o);
  }
  if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5] = va; t[threadIdx.x >> 5] = vb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0;
    for (int i = 0; i < 32; ++i) { x = fmax(x, s[i]); y += t[i]; }
    scalars[S_GMAX] = x;
    scalars[S_GNORM2] = y + tail[1];
    scalars[S_COST] = tail[0];
  }
}

constexpr auto k_edge_gram_photo = k_edge_gram<8, 15>;
constexpr auto k_edge_gram_geom = k_edge_gram<2, 13>;

}  // namespace

// After a Jacobian evaluation: direct normal-equation partials + landmark rows.
pba_status launch_post_jacobian(Handle* h) {
  const Sizes& z = h->sz;
  if (z.n_chunks > 0) {
    if (z.mode == PBA_MODE_PHOTOMETRIC) {
      constexpr size_t smem = (kGramStages * 9 * GramCfg<8, 15>::RS + 256) * sizeof(double);
      PBA_LAUNCH(h, K_EDGE_GRAM, k_edge_gram_photo, dim3(z.n_chunks), dim3(96), smem, z.ld, h->chunk_edge.p,
                 h->chunk_begin.p, h->chunk_end.p, h->edge_h.p, h->edge_t.p, h->d_slot.p, h->J.p, h->edge_M.p, h->part_dir.p);
    } else {
      constexpr size_t smem = (kGramStages * 9 * GramCfg<2, 13>::RS + 256) * sizeof(double);
      PBA_LAUNCH(h, K_EDGE_GRAM, k_edge_gram_geom, dim3(z.n_chunks), dim3(96), smem, z.ld, h->chunk_edge.p,
                 h->chunk_begin.p, h->chunk_end.p, h->edge_h.p, h->edge_t.p, h->d_slot.p, h->J.p, h->edge_M.p, h->part_dir.p);
    }
  }
  if (z.n_lm > 0) {
    PBA_LAUNCH(h, K_LM_GATHER, k_lm_gather, dim3((z.n_lm + 7) / 8), dim3(128), 0, z.n_lm, z.cd, h->lm_ptr.p,
               h->lm_pos.p, h->obs_col.p, h->lm_hostcol.p, h->lm_w_off.p, h->lm_w_stride.p, h->orec.p, h->W.p,
               h->lm_c.p, h->lm_g.p);
  }
  return PBA_OK;
}

void schur_syrk_work(const std::vector<int>& grp_cam_ptr, std::vector<int>* work) {
  work->clear();
  for (size_t g = 0; g + 1 < grp_cam_ptr.size(); ++g) {
    const int c = grp_cam_ptr[g + 1] - grp_cam_ptr[g];
    if (c == 0) continue;
    const int n_tiles = c * (c + 1) / 2 + c;
    for (int t0 = 0; t0 < n_tiles; t0 += 8 * kSyrkTiles) { work->push_back(int(g)); work->push_back(t0); }
  }
}

int schur_tile_l(int max_stride) {
  int t = 64;
  while (t > 1 && 2 * size_t(t) * (syrk_row_stride(max_stride) + 1) * sizeof(double) > 160 * 1024) t >>= 1;
  return t;
}

// Build the damped, Jacobi-scaled RCS for `radius` from the partials of the
// last Jacobian evaluation (SchurEliminator::Eliminate).
pba_status launch_build_rcs(Handle* h, double radius, bool refresh_diag, bool with_scalars) {
  const Sizes& z = h->sz;
  const int init_scale = h->scale_ready ? 0 : 1;
  const int jacobi = h->opt.jacobi_scaling;
  if (z.n_lm > 0) {
    PBA_LAUNCH(h, K_LM_SCALE, k_lm_scale, dim3((z.n_lm + 255) / 256), dim3(256), 0, z.n_lm, init_scale, jacobi,
               refresh_diag ? 1 : 0, radius, h->opt.min_lm_diagonal, h->opt.max_lm_diagonal, h->lm_c.p,
               h->lm_scale.p, h->lm_diag.p, h->lm_iete.p, h->lm_s2.p);
  }
  if (z.n_groups > 0) {
    const int tile_l = h->schur_tile_l;
    const size_t smem = 2 * (size_t(tile_l) * syrk_row_stride(h->max_w_stride) + tile_l) * sizeof(double);
    if (h->max_w_stride <= 8 * (kSyrkRowsMaxC + 1)) {
      PBA_LAUNCH(h, K_SCHUR_SYRK, k_schur_syrk_rows, dim3(z.n_groups), dim3(kSyrkRowsThreads), smem, z.cd, tile_l, h->grp_lm_ptr.p,
                 h->grp_cam_ptr.p, h->grp_w_off.p, h->grp_part_off.p, h->W.p, h->lm_s2.p, h->part_sch.p);
    } else {
      if (h->n_syrk_work > 0)
        PBA_LAUNCH(h, K_SCHUR_SYRK, k_schur_syrk, dim3(h->n_syrk_work), dim3(256), smem, z.cd, tile_l, h->syrk_work.p,
                   h->grp_lm_ptr.p, h->grp_cam_ptr.p, h->grp_w_off.p, h->grp_part_off.p, h->W.p, h->lm_s2.p, h->part_sch.p);
    }
  }
  double* S = h->rcs.p;
  double* rhs = S + z.n_blocks * z.cd * z.cd;
  double* diagB = rhs + z.dim;
  double* gcam = diagB + z.dim;
  const int64_t grid = z.n_blocks + z.n_slots;
  if (grid > 0) {
    PBA_LAUNCH(h, K_RCS_REDUCE, k_rcs_reduce, dim3((unsigned)grid), dim3(64), 0, z.cd, z.n_blocks, z.n_slots,
               h->blk_dir_ptr.p, h->blk_dir_src.p, h->blk_sch_ptr.p, h->blk_sch_src.p, h->vec_dir_ptr.p,
               h->vec_dir_src.p, h->vec_sch_ptr.p, h->vec_sch_src.p, h->d_blk_row.p, h->d_blk_col.p, h->part_dir.p,
               h->part_sch.p, S, rhs, diagB, gcam, h->rcs_B.p, h->rcs_B.p + z.n_blocks * z.cd * z.cd);
  }
  if (h->world > 1) {
    pba_status st = allreduce_rcs(h, with_scalars);
    if (st != PBA_OK) return st;
  }
  if (z.dim > 0) {
    PBA_LAUNCH(h, K_CAM_SCALE, k_cam_scale, dim3((z.dim + 255) / 256), dim3(256), 0, z.dim, init_scale, jacobi,
               refresh_diag ? 1 : 0, radius, h->opt.min_lm_diagonal, h->opt.max_lm_diagonal, diagB, h->cam_scale.p,
               h->cam_diag.p, h->cam_D2.p);
    PBA_LAUNCH(h, K_RCS_SCALE, k_rcs_scale, dim3((unsigned)grid), dim3(64), 0, z.cd, z.n_blocks, z.n_slots,
               h->d_blk_row.p, h->d_blk_col.p, h->cam_scale.p, h->cam_D2.p, S, rhs);
  }
  h->scale_ready = true;
  h->have_rcs = true;
  h->rcs_radius = radius;
  return PBA_OK;
}

// y_cam (scaled space) -> unscaled tangent steps for cameras and landmarks, and the
// model cost change of that step (scalar S_MODEL; this rank's share when sharded).
pba_status launch_backsub(Handle* h) {
  const Sizes& z = h->sz;
  if (z.dim > 0) {
    PBA_LAUNCH(h, K_BACKSUB, k_cam_step, dim3((z.dim + 255) / 256), dim3(256), 0, z.dim, h->y_cam.p, h->cam_scale.p,
               h->d_cam.p);
  }
  const int g_lm = (z.n_lm + 127) / 128;
  const int g_cam = int((z.n_blocks + 15) / 16);  // 16 RCS blocks per CTA (eight lanes each)
  double* part = h->red_ws.p;
  if (g_lm > 0) {
    // streaming mat-vec while a row fits 2 / 4 double2 steps per lane (<= 15 / 31 cameras per host group);
    // PBA_BACKSUB_V1=1 keeps the per-element kernel for A/B runs
    static const bool force_v1 = getenv("PBA_BACKSUB_V1") != nullptr;
    if (!force_v1 && h->max_w_stride <= 128) {
      PBA_LAUNCH(h, K_BACKSUB, k_backsub_stream<2>, dim3(g_lm), dim3(128), 0, z.n_lm, z.cd, h->lm_group.p, h->grp_cam_ptr.p,
                 h->grp_cams.p, h->lm_w_off.p, h->lm_w_stride.p, h->lm_ptr.p, h->W.p, h->lm_scale.p, h->lm_iete.p,
                 h->d_cam.p, h->d_rho.p, part);
    } else if (!force_v1 && h->max_w_stride <= 256) {
      PBA_LAUNCH(h, K_BACKSUB, k_backsub_stream<4>, dim3(g_lm), dim3(128), 0, z.n_lm, z.cd, h->lm_group.p, h->grp_cam_ptr.p,
                 h->grp_cams.p, h->lm_w_off.p, h->lm_w_stride.p, h->lm_ptr.p, h->W.p, h->lm_scale.p, h->lm_iete.p,
                 h->d_cam.p, h->d_rho.p, part);
    } else {
      PBA_LAUNCH(h, K_BACKSUB, k_backsub, dim3(g_lm), dim3(256), 0, z.n_lm, z.cd, h->lm_group.p, h->grp_cam_ptr.p,
                 h->grp_cams.p, h->lm_w_off.p, h->lm_w_stride.p, h->lm_ptr.p, h->W.p, h->lm_scale.p, h->lm_iete.p,
                 h->d_cam.p, h->d_rho.p, part);
    }
  }
  if (g_cam > 0) {
    PBA_LAUNCH(h, K_MODEL_COST, k_model_cost_cam, dim3(g_cam), dim3(128), 0, z.cd, z.n_blocks, h->d_blk_row.p,
               h->d_blk_col.p, h->rcs_B.p, h->rcs_B.p + z.n_blocks * z.cd * z.cd, h->d_cam.p, part + g_lm);
  }
  launch_reduce_sum(h, part, int64_t(g_lm) + g_cam, h->scalars.p + S_MODEL);
  return PBA_OK;
}

// candidate = Plus(x, d); scalars S_STEP2 = ||x - cand||^2, S_XNORM2 = ||cand||^2.
pba_status launch_retract(Handle* h) {
  const Sizes& z = h->sz;
  const int items = z.n_poses + z.n_lm;
  const int grid = (items + 255) / 256;
  double* pa = h->red_ws.p;
  double* pb = pa + grid;
  const bool photo = z.mode == PBA_MODE_PHOTOMETRIC;
  PBA_LAUNCH(h, K_RETRACT, k_retract, dim3(grid), dim3(256), 0, z.n_poses, z.n_lm, z.cd, h->rank == 0 ? 1 : 0,
             h->d_slot.p, h->d_affine_active.p, h->lm_ptr.p, h->poses.p, photo ? h->affine.p : nullptr, h->rho.p,
             h->d_cam.p, h->d_rho.p, h->poses_c.p, photo ? h->affine_c.p : nullptr, h->rho_c.p, pa, pb);
  PBA_LAUNCH(h, K_REDUCE_SUM, k_reduce2, dim3(1), dim3(1024), 0, pa, pb, int64_t(grid), 0, h->scalars.p + S_STEP2,
             h->scalars.p + S_XNORM2);
  return PBA_OK;
}

// world > 1, before the RCS all-reduce: this rank's landmark gradient norms into rcs_tail()
pba_status launch_landmark_gradient_norms(Handle* h) {
  const Sizes& z = h->sz;
  double* tail = h->rcs_tail();
  PBA_CUDA_OK(cudaMemsetAsync(tail + 1, 0, sizeof(double) * size_t(1 + h->world), h->stream));
  const int grid = (z.n_lm + 255) / 256;
  if (grid == 0) return PBA_OK;
  double* pa = h->red_ws.p;
  double* pb = pa + grid;
  PBA_LAUNCH(h, K_REDUCE_SUM, k_grad_norms, dim3(grid), dim3(256), 0, 0, z.n_lm, z.cd, 0, h->d_slot.p, h->d_affine_active.p,
             h->lm_ptr.p, h->poses.p, (const double*)nullptr, h->lm_g.p, pa, pb);
  PBA_LAUNCH(h, K_REDUCE_SUM, k_reduce2, dim3(1), dim3(1024), 0, pa, pb, int64_t(grid), 1, tail + 2 + h->rank, tail + 1);
  return PBA_OK;
}

pba_status launch_gradient_norms(Handle* h) {
  const Sizes& z = h->sz;
  if (h->world > 1) {
    // poses only (identical on every rank: the camera gradient is all-reduced), then the finishing kernel
    const int grid = (z.n_poses + 255) / 256;
    double* pa = h->red_ws.p;
    double* pb = pa + grid;
    const double* gcam = h->rcs.p + z.n_blocks * z.cd * z.cd + 2 * z.dim;
    if (grid > 0) {
      PBA_LAUNCH(h, K_REDUCE_SUM, k_grad_norms, dim3(grid), dim3(256), 0, z.n_poses, 0, z.cd, 1, h->d_slot.p,
                 h->d_affine_active.p, h->lm_ptr.p, h->poses.p, gcam, h->lm_g.p, pa, pb);
    }
    PBA_LAUNCH(h, K_REDUCE_SUM, k_finish_norms, dim3(1), dim3(1024), 0, pa, pb, int64_t(grid), h->rcs_tail(), h->world,
               h->scalars.p);
    return PBA_OK;
  }
  const int items = z.n_poses + z.n_lm;
  const int grid = (items + 255) / 256;
  double* pa = h->red_ws.p;
  double* pb = pa + grid;
  const double* gcam = h->rcs.p + z.n_blocks * z.cd * z.cd + 2 * z.dim;
  PBA_LAUNCH(h, K_REDUCE_SUM, k_grad_norms, dim3(grid), dim3(256), 0, z.n_poses, z.n_lm, z.cd, h->rank == 0 ? 1 : 0,
             h->d_slot.p, h->d_affine_active.p, h->lm_ptr.p, h->poses.p, gcam, h->lm_g.p, pa, pb);
  PBA_LAUNCH(h, K_REDUCE_SUM, k_reduce2, dim3(1), dim3(1024), 0, pa, pb, int64_t(grid), 1, h->scalars.p + S_GMAX,
             h->scalars.p + S_GNORM2);
  return PBA_OK;
}

}  // namespace pba
