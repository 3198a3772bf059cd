// eval.cu — residual / Jacobian evaluation kernels (K0, K1, K2).
//
// Replaces, for one shard of observations, Ceres' ProgramEvaluator::Evaluate
// (internal/ceres/program_evaluator.h:139-286) together with the reference
// functor (include/visnav/reprojection.h:82-112), the local-parameterisation
// multiply and the robust correction of ResidualBlock::Evaluate
// (internal/ceres/residual_block.cc:69-198).  Jacobians are closed-form
// (SURVEY.md §8(a)); the photometric residual is SURVEY.md §8(a-P).
//
// One thread per observation, observations in (host,target)-edge order, all
// outputs in SoA planes: every warp store is 256 contiguous bytes.
#include <stdlib.h>

#include "launch.h"
#include "pba_internal.h"

namespace pba {

namespace {

constexpr int kEvalThreads = 128;

struct EvalArgs {
  int64_t n;
  int64_t ld;  // plane stride of res / J
  int n_lm;
  // structure
  const int* obs_lm;
  const int* obs_edge;
  const int* edge_h;
  const int* edge_t;
  const int* pose_calib;
  const int* calib_model;
  const double* intr;
  const double* edge_T;  // [E][kEdgeStride] per-edge record (pba_internal.h)
  // landmark constants
  const double* lm_pat;
  const double* lm_uv;   // [n_lm][2] host pixel (bearing recomputation, pinhole)
  const uint8_t* lm_ok;
  const double* obs_uv;  // geometric [2][n]
  // keyframes as quads: one uint32 (I00 | I10<<8 | I01<<16 | I11<<24) per pixel
  const uint32_t* quads;
  int64_t image_stride;  // pixels per keyframe
  int width, height, pitch;
  // state
  const double* rho;
  // robust loss
  int use_huber;
  double huber;
  // outputs
  double* J;   // interleaved planes [R][C+1][ld]: J row columns 0..C-1, column C = residual
  double* orec;
  double* block_cost;
};

// K0: per-edge relative pose A = R_t^T R_h, t = R_t^T (t_h - t_t) and the
// target's affine brightness (exp(a), b).
//
// edge_M (Jacobian evaluations only): with the right-multiplied
// local parameterisation X_t = exp(-d_t) T_rel exp(d_h) P, and exp(-d_t) T_rel =
// T_rel exp(-Ad(T_rel^-1) d_t), so every Jacobian row satisfies
//     d r / d d_t = (d r / d d_h) M,   M = -Ad(T_rel^-1) = [[-A^T, A^T [t]x], [0, -A^T]]
// — one 6x6 matrix per EDGE.  K1 therefore does not store the six target-pose planes, and the
// Gram kernel derives the (h,t) and (t,t) blocks from the (h,h) block.
// 64 edges per CTA; every thread builds its edge's 256-byte record (and the 6x6 adjoint) in SHARED memory (row
// strides 33 / 37: conflict-free) and the CTA then copies both out as contiguous runs.  Writing the records straight
// from the threads — 32 + 36 stores per thread, each a 32-sector scatter — cost 11 us per launch for 20 k edges, twice
// per LM iteration and not shrinking with the number of GPUs.
constexpr int kPrepThreads = 64;
__global__ void __launch_bounds__(kPrepThreads) k_edge_prep(int n_edges, const int* __restrict__ edge_h,
                                                             const int* __restrict__ edge_t,
                                                             const double* __restrict__ poses, const double* __restrict__ affine,
                                                             const int* __restrict__ pose_calib, const int* __restrict__ calib_model,
                                                             const double* __restrict__ intr, double* __restrict__ edge_T,
                                                             double* __restrict__ edge_M) {
  __shared__ double s_T[kPrepThreads * (kEdgeStride + 1)];
  __shared__ double s_M[kPrepThreads * 37];
  const int e0 = blockIdx.x * kPrepThreads;
  const int e = e0 + threadIdx.x;
  if (e < n_edges) {
    const double* Th = poses + 7 * edge_h[e];
    const double* Tt = poses + 7 * edge_t[e];
    double Rh[9], Rt[9];
    quat_to_rot(Th, Rh);
    quat_to_rot(Tt, Rt);
    double* o = s_T + (kEdgeStride + 1) * threadIdx.x;
    {
      const int tc = pose_calib[edge_t[e]], hc = pose_calib[edge_h[e]];
      o[14] = double(tc);
      o[15] = double(calib_model[hc]);
      for (int q = 0; q < 8; ++q) o[16 + q] = intr[8 * tc + q];
      o[24] = 1.0 / intr[8 * hc]; o[25] = 1.0 / intr[8 * hc + 1]; o[26] = intr[8 * hc + 2]; o[27] = intr[8 * hc + 3];
      o[28] = o[29] = o[30] = o[31] = 0.0;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o[3 * i + j] = Rt[0 + i] * Rh[0 + j] + Rt[3 + i] * Rh[3 + j] + Rt[6 + i] * Rh[6 + j];
    const double dx = Th[4] - Tt[4], dy = Th[5] - Tt[5], dz = Th[6] - Tt[6];
    for (int i = 0; i < 3; ++i) o[9 + i] = Rt[0 + i] * dx + Rt[3 + i] * dy + Rt[6 + i] * dz;
    if (affine) {
      o[12] = exp(affine[2 * edge_t[e]]);
      o[13] = affine[2 * edge_t[e] + 1];
    } else {
      o[12] = 1.0;
      o[13] = 0.0;
    }
    if (edge_M) {
      double* m = s_M + 37 * threadIdx.x;
      const double tx = o[9], ty = o[10], tz = o[11];
      // S = [t]x
      const double S[9] = {0.0, -tz, ty, tz, 0.0, -tx, -ty, tx, 0.0};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          const double at = o[3 * j + i];  // (A^T)[i][j]
          double as = 0.0;                 // (A^T S)[i][j]
          for (int k = 0; k < 3; ++k) as += o[3 * k + i] * S[3 * k + j];
          m[6 * i + j] = -at;
          m[6 * i + 3 + j] = as;
          m[6 * (3 + i) + j] = 0.0;
          m[6 * (3 + i) + 3 + j] = -at;
        }
    }
  }
  __syncthreads();
  const int cnt = min(kPrepThreads, n_edges - e0);
  for (int x = threadIdx.x; x < cnt * kEdgeStride; x += kPrepThreads)
    edge_T[int64_t(e0) * kEdgeStride + x] = s_T[(x / kEdgeStride) * (kEdgeStride + 1) + x % kEdgeStride];
  if (edge_M)
    for (int x = threadIdx.x; x < cnt * 36; x += kPrepThreads) edge_M[int64_t(e0) * 36 + x] = s_M[(x / 36) * 37 + x % 36];
}

// 2x2 bilinear footprints: quad(x,y) = I(x,y) | I(x+1,y)<<8 | I(x,y+1)<<16 | I(x+1,y+1)<<24
// (clamped at the right/bottom border, which valid samples never touch).
__global__ void k_build_quads(int width, int height, int pitch, int64_t n_img, const uint8_t* __restrict__ img,
                              uint32_t* __restrict__ quads) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per = int64_t(width) * height;
  if (i >= per * n_img) return;
  const int64_t f = i / per;
  const int y = int((i % per) / width), x = int(i % width);
  const int x1 = x + 1 < width ? x + 1 : x, y1 = y + 1 < height ? y + 1 : y;
  const uint8_t* p = img + f * int64_t(pitch) * height;
  quads[i] = uint32_t(p[int64_t(y) * pitch + x]) | (uint32_t(p[int64_t(y) * pitch + x1]) << 8) |
             (uint32_t(p[int64_t(y1) * pitch + x]) << 16) | (uint32_t(p[int64_t(y1) * pitch + x1]) << 24);
}

// Landmark constants: unit host bearings (reprojection.h:106-107) and, for the
// photometric residual, the bilinear host intensities of the 8-pixel pattern.
// Photometric layout: SoA planes lm_pat[(k*4 + {bx,by,bz,I_h}) * n_lm + l].
__global__ void k_init_landmarks(int n_lm, int photo, const int* __restrict__ lm_host,
                                 const double* __restrict__ lm_uv, const int* __restrict__ pose_calib,
                                 const int* __restrict__ calib_model, const double* __restrict__ intr,
                                 const uint32_t* __restrict__ quads, int64_t image_stride, int width, int height,
                                 double* __restrict__ lm_pat, uint8_t* __restrict__ lm_ok) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lm) return;
  const int h = lm_host[l];
  const int c = pose_calib[h];
  const int model = calib_model[c];
  double in[8];
  for (int i = 0; i < 8; ++i) in[i] = intr[8 * c + i];
  const double u = lm_uv[2 * l], v = lm_uv[2 * l + 1];
  if (!photo) {
    double b[3];
    cam_bearing(model, in, u, v, b);
    lm_pat[4 * l + 0] = b[0]; lm_pat[4 * l + 1] = b[1]; lm_pat[4 * l + 2] = b[2]; lm_pat[4 * l + 3] = 0.0;
    lm_ok[l] = 1;
    return;
  }
  const uint32_t* img = quads + int64_t(h) * image_stride;
  bool ok = true;
  for (int k = 0; k < 8; ++k) {
    const double pu = u + kPatternDev[k][0], pv = v + kPatternDev[k][1];
    double b[3];
    cam_bearing(model, in, pu, pv, b);
    double I = 0.0;
    if (pu >= 0.0 && pv >= 0.0 && pu < double(width - 1) && pv < double(height - 1)) {
      const int x0 = int(floor(pu)), y0 = int(floor(pv));
      const double fx = pu - x0, fy = pv - y0;
      const uint32_t q = img[int64_t(y0) * width + x0];
      const double i00 = double(q & 0xffu), i10 = double((q >> 8) & 0xffu), i01 = double((q >> 16) & 0xffu),
                   i11 = double(q >> 24);
      I = (1.0 - fx) * (1.0 - fy) * i00 + fx * (1.0 - fy) * i10 + (1.0 - fx) * fy * i01 + fx * fy * i11;
    } else {
      ok = false;
    }
    double* o = lm_pat + int64_t(4 * k) * n_lm + l;
    o[0] = b[0]; o[int64_t(n_lm)] = b[1]; o[2 * int64_t(n_lm)] = b[2]; o[3 * int64_t(n_lm)] = I;
  }
  lm_ok[l] = ok;
}

__device__ __forceinline__ double block_sum(double v, double* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += smem[i];
  }
  return s;  // valid on thread 0
}

// HuberLoss::Evaluate (loss_function.cc:48-62) + Corrector (corrector.cc:82-86):
// returns rho(s)/2, *w = sqrt(rho'(s)).
__device__ __forceinline__ double huber(double s, int use_huber, double a, double* w) {
  *w = 1.0;
  if (!use_huber) return 0.5 * s;
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    const double rho1 = fmax(2.2250738585072014e-308, a / r);
    *w = sqrt(rho1);
    return 0.5 * (2.0 * a * r - b);
  }
  return 0.5 * s;
}

// ------------------------------------------------------------ photometric --
// K1 (WITH_J) / K2 (!WITH_J).  One thread per observation, 8 pattern pixels per
// thread; a warp's 32 consecutive observations make every store a full 256 B
// run of one SoA plane.
//
// The first version of this kernel (profiles/r01a_*) was bound by L1 wavefronts,
// not HBM: per observation it issued 32 single-byte image taps and 16 LDG.128
// of an AoS landmark record, every one of them a 32-line gather.  Now
//   * keyframes are stored as QUADS: one 32-bit word per pixel holding the 2x2
//     bilinear footprint (I00 I10 I01 I11), so a pixel costs ONE gather;
//   * landmark pattern constants are SoA planes [k][component][landmark]:
//     consecutive observations of an edge read near-consecutive landmarks;
//   * the block's Huber weight needs all 8 residuals before any Jacobian row
//     can be scaled, so pass 1 computes residuals (this IS the K2 code) and
//     keeps only the 8 quad words; pass 2 recomputes the warp + projection
//     Jacobian from registers (no second gather) and streams the weighted rows;
//   * the 128 B Schur record is transposed through shared memory per warp so
//     it leaves as coalesced 256 B stores.
// (A variant with 8 lanes per observation and shuffle reductions was measured
// slower — 7.7 ms vs 6.7 ms at 18M observations: 4x the load instructions for
// the per-edge constants.)
constexpr int kPhotoThreads = 128;
constexpr size_t kPhotoSmemBytes = 32 * kPhotoThreads * sizeof(double) + 16 * kPhotoThreads * 4;

constexpr size_t kPhotoSmemBytesRC = (8 * kPhotoThreads + 4 * 32 * 17) * sizeof(double) + 16 * kPhotoThreads * 4;

// Pinhole host point of pattern pixel k: X_h = b_k / rho with b_k = m_k / |m_k|,
// m_k = (mx0 + o_x / fx, my0 + o_y / fy, 1)  (camera_models.h:93-107 + reprojection.h:106-107).
__device__ __forceinline__ void pinhole_host_point(double mx0, double my0, double ifx, double ify, int k, double irho,
                                                   double& xh, double& yh, double& zh) {
  const double mx = fma(double(kPatternDev[k][0]), ifx, mx0);
  const double my = fma(double(kPatternDev[k][1]), ify, my0);
  const double sc = rsqrt(fma(mx, mx, fma(my, my, 1.0))) * irho;
  xh = mx * sc; yh = my * sc; zh = sc;
}

struct PhotoCtx {
  double A[9], tr[3], ea, bb, in[8], irho;
  int model;
};

__device__ __forceinline__ double quad_sample(uint32_t q, double fx, double fy) {
  const double i00 = double(q & 0xffu), i10 = double((q >> 8) & 0xffu), i01 = double((q >> 16) & 0xffu),
               i11 = double(q >> 24);
  return (1.0 - fx) * (1.0 - fy) * i00 + fx * (1.0 - fy) * i10 + (1.0 - fx) * fy * i01 + fx * fy * i11;
}

// MODEL >= 0: every camera of the problem uses that model (the usual case): the
// projection is branch-free, so the compiler can batch the gathers.  MODEL = -1:
// mixed models, runtime switch per observation.
//
// RC (pinhole only): the host bearings of the 8 pattern pixels are RECOMPUTED from the host pixel
// (b_k = m_k / |m_k|, m_k = ((u + o_x - cx) / fx, (v + o_y - cy) / fy, 1): two FMAs, a dot product and
// one rsqrt per pixel) instead of being read from the 24 lm_pat planes.  That removes 192 B per
// observation of L2 -> SM traffic, the 48 staging registers of phase 0 and the bearing part of the
// pass-1 -> pass-2 hand-over (pass 2 recomputes them as well), which is what capped the kernel at 3
// CTAs per SM (ncu, profiles/r01g_k1_*: 168 registers + 57 KB shared, 18.75 % occupancy, 65 % of the
// cycles without an eligible warp).  The host intensities I_h stay a table (they need the image).
// MINB = CTAs per SM the register allocation is bounded for, UNR = unroll factor of the pass-2 pixel loop.
template <bool WITH_J, int MODEL, bool RC, int MINB = (WITH_J ? 3 : 4), int UNR = 2>
__global__ void __launch_bounds__(kPhotoThreads, MINB) k_eval_photo(const EvalArgs a) {
  static_assert(!RC || MODEL == PBA_CAM_PINHOLE, "bearing recomputation is implemented for the pinhole model");
  __shared__ double s_red[kPhotoThreads / 32];
  // Dynamic shared memory (K1 only):
  //   table path (kPhotoSmemBytes = 40 KB):
  //     s_pat  [4 warps][32 values][32 lanes] doubles  pass-1 -> pass-2 hand-over: bx by bz Ih per pixel
  //     s_rec  [4 warps][32][17] doubles  Schur-record transposition, ALIASED on the warp's own s_pat
  //            block (dead once the warp has left the pixel loop)
  //   RC path (kPhotoSmemBytesRC = 33 KB): s_pat holds only I_h: [4 warps][8 values][32 lanes]; s_rec
  //            [4 warps][32][17] has its own region behind it
  //   both: s_quad [8][128] u32, s_off [8][128] int
  // [value][thread] layouts are conflict-free.  Keeping the hand-over in SHARED memory
  // matters: as a per-thread local array it spills through L2 to HBM (ncu,
  // profiles/r01b_*: 26.4 GB written per launch against 20.7 GB of outputs).
  extern __shared__ __align__(16) unsigned char k1_sm[];
  constexpr int kPatVals = RC ? 8 : 32;           // hand-over values per thread
  constexpr int kRecBase = RC ? 8 * kPhotoThreads : 0;  // doubles
  constexpr int kWordBase = RC ? 8 * kPhotoThreads + 4 * 32 * 17 : 32 * kPhotoThreads;  // doubles
  double* s_pat = reinterpret_cast<double*>(k1_sm) + (threadIdx.x >> 5) * (32 * kPatVals) + (threadIdx.x & 31);  // + 32 * value
  double* s_rec = reinterpret_cast<double*>(k1_sm) + kRecBase + (threadIdx.x >> 5) * (RC ? 32 * 17 : 1024);
  uint32_t* s_quad = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(k1_sm) + kWordBase);
  int* s_off = reinterpret_cast<int*>(s_quad + 8 * kPhotoThreads);
  const int tid = threadIdx.x;
  const int64_t i = int64_t(blockIdx.x) * kPhotoThreads + tid;
  const int lane = tid & 31, warp = tid >> 5;
  double cost = 0.0;
  double acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.0;
  if (i < a.n) {
    const int e = a.obs_edge[i];
    const int l = a.obs_lm[i];
    const int t = a.edge_t[e];
    const double2* T2 = reinterpret_cast<const double2*>(a.edge_T + kEdgeStride * int64_t(e));
    PhotoCtx c;
    c.model = MODEL >= 0 ? MODEL : a.calib_model[int(T2[7].x)];
    // ---- phase 0: all independent loads first (pattern constants, edge, intrinsics) ----
    double bx[RC ? 1 : 8], by[RC ? 1 : 8], bz[RC ? 1 : 8], Ih[8];
    double mx0 = 0.0, my0 = 0.0, ifx = 0.0, ify = 0.0;  // RC: normalised host pixel and 1 / f of the HOST camera
    {
      const double* pk = a.lm_pat + l;
      const int64_t nl = a.n_lm;
      if (RC) {
        const double2 huv = reinterpret_cast<const double2*>(a.lm_uv)[l];
        const double2 hif = T2[12], hcxy = T2[13];
#pragma unroll
        for (int k = 0; k < 8; ++k) Ih[k] = __ldg(pk + int64_t(4 * k + 3) * nl);
        ifx = hif.x; ify = hif.y;
        mx0 = (huv.x - hcxy.x) * ifx; my0 = (huv.y - hcxy.y) * ify;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          bx[RC ? 0 : k] = __ldg(pk + int64_t(4 * k) * nl);
          by[RC ? 0 : k] = __ldg(pk + int64_t(4 * k + 1) * nl);
          bz[RC ? 0 : k] = __ldg(pk + int64_t(4 * k + 2) * nl);
          Ih[k] = __ldg(pk + int64_t(4 * k + 3) * nl);
        }
      }
      const double2 t0 = T2[0], t1 = T2[1], t2 = T2[2], t3 = T2[3], t4 = T2[4], t5 = T2[5], t6 = T2[6];
      c.A[0] = t0.x; c.A[1] = t0.y; c.A[2] = t1.x; c.A[3] = t1.y; c.A[4] = t2.x; c.A[5] = t2.y; c.A[6] = t3.x;
      c.A[7] = t3.y; c.A[8] = t4.x; c.tr[0] = t4.y; c.tr[1] = t5.x; c.tr[2] = t5.y; c.ea = t6.x; c.bb = t6.y;
#pragma unroll
      for (int q = 0; q < (MODEL == PBA_CAM_PINHOLE ? 2 : 4); ++q) { const double2 v = T2[8 + q]; c.in[2 * q] = v.x; c.in[2 * q + 1] = v.y; }
    }
    c.irho = 1.0 / a.rho[l];
    const uint32_t* img = a.quads + int64_t(t) * a.image_stride;
    bool ok = a.lm_ok[l] != 0;

    // ---- phase 1: warp the 8 pixels, branch-free; then gather the 8 quads together ----
    double fx[8], fy[8];
    int off[8];
    int cell[8];  // (y0 << 16 | x0): pass 2 needs the integer cell, not the offset (no integer division by the pitch there)
    const double umax = double(a.width - 1), vmax = double(a.height - 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double xh, yh, zh;
      if (RC) {
        pinhole_host_point(mx0, my0, ifx, ify, k, c.irho, xh, yh, zh);
      } else {
        xh = bx[RC ? 0 : k] * c.irho; yh = by[RC ? 0 : k] * c.irho; zh = bz[RC ? 0 : k] * c.irho;
      }
      const double xt = c.A[0] * xh + c.A[1] * yh + c.A[2] * zh + c.tr[0];
      const double yt = c.A[3] * xh + c.A[4] * yh + c.A[5] * zh + c.tr[1];
      const double zt = c.A[6] * xh + c.A[7] * yh + c.A[8] * zh + c.tr[2];
      double uv[2];
      cam_project<false>(c.model, c.in, xt, yt, zt, uv, nullptr);
      const bool inb = uv[0] >= 0.0 && uv[1] >= 0.0 && uv[0] < umax && uv[1] < vmax;  // false for NaN
      ok &= inb;
      const double u = inb ? uv[0] : 0.0, v = inb ? uv[1] : 0.0;
      const double x0 = floor(u), y0 = floor(v);
      fx[k] = u - x0; fy[k] = v - y0;
      off[k] = int(y0) * a.pitch + int(x0);
      cell[k] = (int(y0) << 16) | int(x0);
    }
    uint32_t quad[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) quad[k] = __ldg(img + off[k]);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double rk = quad_sample(quad[k], fx[k], fy[k]) - (c.ea * Ih[k] + c.bb);
      s += rk * rk;
    }
    if (!ok) s = 0.0;  // invalid observation: r = 0, J = 0
    double w;
    cost = huber(s, a.use_huber, a.huber, &w);
    if (!ok) w = 0.0;

    if (WITH_J) {
      // invalid observation: the block is exactly zero.  Multiplying by w = 0 is not enough — a failed
      // projection can be Inf / NaN (z_t = 0, rho = 0) and 0 * NaN = NaN would reach the normal
      // equations — so pass 2 runs on a harmless point instead (X_h = 0, X_t = (0, 0, 1)).
      if (!ok) c.irho = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (RC) {
          s_pat[k * 32] = Ih[k];
        } else {
          s_pat[(4 * k + 0) * 32] = bx[RC ? 0 : k];
          s_pat[(4 * k + 1) * 32] = by[RC ? 0 : k];
          s_pat[(4 * k + 2) * 32] = bz[RC ? 0 : k];
          s_pat[(4 * k + 3) * 32] = Ih[k];
        }
        s_quad[k * kPhotoThreads + tid] = quad[k];
        s_off[k * kPhotoThreads + tid] = cell[k];
      }
      // ---- phase 2: one pixel at a time: recompute the warp with its projection Jacobian
      //      (no global loads), weight, stream the row out ----
      const int64_t n = a.ld;
#pragma unroll UNR
      for (int k = 0; k < 8; ++k) {
        double xh, yh, zh, Ihk;
        if (RC) {
          Ihk = s_pat[k * 32];
          pinhole_host_point(mx0, my0, ifx, ify, k, c.irho, xh, yh, zh);
        } else {
          const double bxk = s_pat[(4 * k + 0) * 32], byk = s_pat[(4 * k + 1) * 32], bzk = s_pat[(4 * k + 2) * 32];
          Ihk = s_pat[(4 * k + 3) * 32];
          xh = bxk * c.irho; yh = byk * c.irho; zh = bzk * c.irho;
        }
        const uint32_t q = s_quad[k * kPhotoThreads + tid];
        const int ofk = s_off[k * kPhotoThreads + tid];
        const double xt = c.A[0] * xh + c.A[1] * yh + c.A[2] * zh + c.tr[0];
        const double yt = c.A[3] * xh + c.A[4] * yh + c.A[5] * zh + c.tr[1];
        const double zt = c.A[6] * xh + c.A[7] * yh + c.A[8] * zh + c.tr[2];
        double uv[2], Jp[6];
        cam_project<true>(c.model, c.in, ok ? xt : 0.0, ok ? yt : 0.0, ok ? zt : 1.0, uv, Jp);
        // same integer cell as pass 1 (a recomputed floor could differ by one ulp of u)
        const int y0 = ofk >> 16, x0 = ofk & 0xffff;
        const double fxk = ok ? uv[0] - x0 : 0.0, fyk = ok ? uv[1] - y0 : 0.0;
        const double i00 = double(q & 0xffu), i10 = double((q >> 8) & 0xffu), i01 = double((q >> 16) & 0xffu),
                     i11 = double(q >> 24);
        const double I = (1.0 - fxk) * (1.0 - fyk) * i00 + fxk * (1.0 - fyk) * i10 + (1.0 - fxk) * fyk * i01 +
                         fxk * fyk * i11;
        const double rk = I - (c.ea * Ihk + c.bb);
        const double gx = (1.0 - fyk) * (i10 - i00) + fyk * (i11 - i01);
        const double gy = (1.0 - fxk) * (i01 - i00) + fxk * (i11 - i10);
        const double p0 = w * (gx * Jp[0] + gy * Jp[3]);
        const double p1 = w * (gx * Jp[1] + gy * Jp[4]);
        const double p2 = w * (gx * Jp[2] + gy * Jp[5]);
        // a = p A  (d r / d upsilon_h);   d r / d omega_h = -(a x X_h)
        const double ax = p0 * c.A[0] + p1 * c.A[3] + p2 * c.A[6];
        const double ay = p0 * c.A[1] + p1 * c.A[4] + p2 * c.A[7];
        const double az = p0 * c.A[2] + p1 * c.A[5] + p2 * c.A[8];
        double row[15];
        row[0] = ax; row[1] = ay; row[2] = az;
        row[3] = -(ay * zh - az * yh);
        row[4] = -(az * xh - ax * zh);
        row[5] = -(ax * yh - ay * xh);
        // target pose [-p | p x X_t] = row[0..5] x M of the edge: neither stored nor formed here
        // affine (a_t, b_t)
        row[12] = -w * c.ea * Ihk;
        row[13] = -w;
        // inverse distance: -(a . X_h) / rho
        row[14] = -(ax * xh + ay * yh + az * zh) * c.irho;
        const double rw = w * rk;
        // kPhotoPlanes planes per row: columns 6..11 (target pose) are NOT stored, they are
        // row[0..5] x M of the edge (k_edge_prep); stored_plane() compacts the rest, residual last
        double* Jk = a.J + (int64_t(k) * kPhotoPlanes) * n + i;
        Jk[int64_t(stored_plane(15)) * n] = rw;
#pragma unroll
        for (int qq = 0; qq < 15; ++qq)
          if (qq < 6 || qq >= 12) Jk[int64_t(stored_plane(qq)) * n] = row[qq];
        const double E = row[14];
#pragma unroll
        for (int qq = 0; qq < 14; ++qq)
          if (qq < 6 || qq >= 12) acc[qq] += E * row[qq];
        acc[14] += E * E;
        acc[15] += E * rw;
      }
      // target-pose part of the Schur record, (E^T J_h) M with M = [[-A^T, A^T [t]x], [0, -A^T]]:
      // u = A v_upsilon, E^T J_t = [-u | u x t - A v_omega]
      {
        const double u0 = c.A[0] * acc[0] + c.A[1] * acc[1] + c.A[2] * acc[2];
        const double u1 = c.A[3] * acc[0] + c.A[4] * acc[1] + c.A[5] * acc[2];
        const double u2 = c.A[6] * acc[0] + c.A[7] * acc[1] + c.A[8] * acc[2];
        const double w0 = c.A[0] * acc[3] + c.A[1] * acc[4] + c.A[2] * acc[5];
        const double w1 = c.A[3] * acc[3] + c.A[4] * acc[4] + c.A[5] * acc[5];
        const double w2 = c.A[6] * acc[3] + c.A[7] * acc[4] + c.A[8] * acc[5];
        acc[6] = -u0; acc[7] = -u1; acc[8] = -u2;
        acc[9] = (u1 * c.tr[2] - u2 * c.tr[1]) - w0;
        acc[10] = (u2 * c.tr[0] - u0 * c.tr[2]) - w1;
        acc[11] = (u0 * c.tr[1] - u1 * c.tr[0]) - w2;
      }
    }
  }
  if (WITH_J) {
    // Schur record [obs][16]: transpose the warp's 32 x 16 values through shared
    // memory (row stride 17) and store them as 16 coalesced 256 B runs.
    __syncwarp();  // table path: the warp is done with its s_pat block, which s_rec aliases
    double* rec = s_rec;
#pragma unroll
    for (int q = 0; q < 16; ++q) rec[lane * 17 + q] = acc[q];
    __syncwarp();
    const int64_t base = int64_t(blockIdx.x) * kPhotoThreads + warp * 32;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ob = 2 * j + (lane >> 4);
      if (base + ob < a.n) a.orec[16 * (base + ob) + (lane & 15)] = rec[ob * 17 + (lane & 15)];
    }
  }
  const double bs = block_sum(cost, s_red);
  if (threadIdx.x == 0) a.block_cost[blockIdx.x] = bs;
}

// K2 for the pinhole / recomputed-bearing path with the eight pattern pixels in TWO HALVES (project 4 -> gather 4 ->
// sample 4, twice): the live ranges of (fx, fy, offset, quad, I_h) halve, the kernel fits 5 / 6 CTAs per SM without
// spills (the template above needs 128 registers for its 8-deep batch; bounding IT to 5 / 6 CTAs spilled and ran
// slower, profiles/r02a_variants.txt).  Same arithmetic as k_eval_photo<false, pinhole, true>, pixel by pixel.
template <int MINB>
__global__ void __launch_bounds__(kPhotoThreads, MINB) k_cost_photo_halves(const EvalArgs a) {
  __shared__ double s_red[kPhotoThreads / 32];
  const int64_t i = int64_t(blockIdx.x) * kPhotoThreads + threadIdx.x;
  double cost = 0.0;
  if (i < a.n) {
    const int e = a.obs_edge[i];
    const int l = a.obs_lm[i];
    const int t = a.edge_t[e];
    const double2* T2 = reinterpret_cast<const double2*>(a.edge_T + kEdgeStride * int64_t(e));
    PhotoCtx c;
    c.model = PBA_CAM_PINHOLE;
    const double2 huv = reinterpret_cast<const double2*>(a.lm_uv)[l];
    const double2 hif = T2[12], hcxy = T2[13];
    const double2 t0 = T2[0], t1 = T2[1], t2 = T2[2], t3 = T2[3], t4 = T2[4], t5 = T2[5], t6 = T2[6];
    c.A[0] = t0.x; c.A[1] = t0.y; c.A[2] = t1.x; c.A[3] = t1.y; c.A[4] = t2.x; c.A[5] = t2.y; c.A[6] = t3.x;
    c.A[7] = t3.y; c.A[8] = t4.x; c.tr[0] = t4.y; c.tr[1] = t5.x; c.tr[2] = t5.y; c.ea = t6.x; c.bb = t6.y;
#pragma unroll
    for (int q = 0; q < 2; ++q) { const double2 v = T2[8 + q]; c.in[2 * q] = v.x; c.in[2 * q + 1] = v.y; }
    const double ifx = hif.x, ify = hif.y;
    const double mx0 = (huv.x - hcxy.x) * ifx, my0 = (huv.y - hcxy.y) * ify;
    c.irho = 1.0 / a.rho[l];
    const uint32_t* img = a.quads + int64_t(t) * a.image_stride;
    bool ok = a.lm_ok[l] != 0;
    const double umax = double(a.width - 1), vmax = double(a.height - 1);
    const double* pk = a.lm_pat + l;
    const int64_t nl = a.n_lm;
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double fx[4], fy[4], Ih[4];
      int off[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        Ih[k] = __ldg(pk + int64_t(4 * (4 * h + k) + 3) * nl);
        double xh, yh, zh;
        pinhole_host_point(mx0, my0, ifx, ify, 4 * h + k, c.irho, xh, yh, zh);
        const double xt = c.A[0] * xh + c.A[1] * yh + c.A[2] * zh + c.tr[0];
        const double yt = c.A[3] * xh + c.A[4] * yh + c.A[5] * zh + c.tr[1];
        const double zt = c.A[6] * xh + c.A[7] * yh + c.A[8] * zh + c.tr[2];
        double uv[2];
        cam_project<false>(PBA_CAM_PINHOLE, c.in, xt, yt, zt, uv, nullptr);
        const bool inb = uv[0] >= 0.0 && uv[1] >= 0.0 && uv[0] < umax && uv[1] < vmax;
        ok &= inb;
        const double u = inb ? uv[0] : 0.0, v = inb ? uv[1] : 0.0;
        const double x0 = floor(u), y0 = floor(v);
        fx[k] = u - x0; fy[k] = v - y0;
        off[k] = int(y0) * a.pitch + int(x0);
      }
      uint32_t quad[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) quad[k] = __ldg(img + off[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double rk = quad_sample(quad[k], fx[k], fy[k]) - (c.ea * Ih[k] + c.bb);
        s += rk * rk;
      }
    }
    if (!ok) s = 0.0;
    double w;
    cost = huber(s, a.use_huber, a.huber, &w);
  }
  const double bs = block_sum(cost, s_red);
  if (threadIdx.x == 0) a.block_cost[blockIdx.x] = bs;
}

// rc: recompute the pattern bearings (all cameras pinhole); see the kernel's comment.
// PBA_K1_VARIANT / PBA_K2_VARIANT pick the occupancy / unroll variants kept for A/B measurements.
template <bool WITH_J>
void (*photo_kernel(int model, bool rc))(const EvalArgs) {
  static const int var = [] { const char* e = getenv(WITH_J ? "PBA_K1_VARIANT" : "PBA_K2_VARIANT"); return e ? atoi(e) : 0; }();
  if (model == PBA_CAM_PINHOLE && rc) {
    if (WITH_J) {
      switch (var) {  // measured (profiles/r02a_variants.txt): 4 CTAs/SM unroll 2: 3.40 ms, unroll 1: 3.45, 5 CTAs (spills): 4.77,
                      // 3 CTAs (166 registers, no spill): 3.43; table path: 3.84
        case 1: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 4, 1>;
        case 2: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 5, 1>;
        case 3: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 3, 2>;
        default: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 4, 2>;
      }
    } else {
      switch (var) {  // measured (profiles/r02a_variants.txt): 4 CTAs/SM 1.055 ms, 5 (spills) 1.080, 6 1.131; table path 1.041
        case 1: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 5, 2>;
        case 2: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 6, 2>;
        case 3: return k_cost_photo_halves<5>;
        case 4: return k_cost_photo_halves<6>;
        case 5: return k_cost_photo_halves<8>;
        default: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, true, 4, 2>;
      }
    }
  }
  switch (model) {
    case PBA_CAM_PINHOLE: return k_eval_photo<WITH_J, PBA_CAM_PINHOLE, false>;
    case PBA_CAM_DS: return k_eval_photo<WITH_J, PBA_CAM_DS, false>;
    case PBA_CAM_KB4: return k_eval_photo<WITH_J, PBA_CAM_KB4, false>;
    case PBA_CAM_EUCM: return k_eval_photo<WITH_J, PBA_CAM_EUCM, false>;
  }
  return k_eval_photo<WITH_J, -1, false>;
}

// -------------------------------------------------------------- geometric --
// reprojection.h:82-112: r = z_t - pi_t(T_t^-1 T_h (b / rho)).  Both cameras
// use the HOST's model (the reference passes the host's model name for both,
// map_utils.h:363-364) with the target's intrinsic values.
// MINB: CTAs per SM the register allocation is bounded for (0 = unbounded: 90 registers, 5 CTAs per SM).  ncu
// (profiles/r02a_k1_geom_details.txt) shows the unbounded kernel latency-bound: IPC 0.63, 84 % of the cycles
// without an eligible warp at 31 % occupancy; PBA_K1G_VARIANT switches the bound for A/B runs.
template <bool WITH_J, int MINB = 0>
__global__ void __launch_bounds__(kEvalThreads, MINB > 0 ? MINB : 1) k_eval_geom(const EvalArgs a) {
  __shared__ double s_red[kEvalThreads / 32];
  const int64_t i = int64_t(blockIdx.x) * kEvalThreads + threadIdx.x;
  double cost = 0.0;
  if (i < a.n) {
    const int e = a.obs_edge[i];
    const int l = a.obs_lm[i];
    const double2* T2 = reinterpret_cast<const double2*>(a.edge_T + kEdgeStride * int64_t(e));
    double T[24];  // the record's first 24 doubles, 16 bytes per load
#pragma unroll
    for (int k = 0; k < 12; ++k) { const double2 v = T2[k]; T[2 * k] = v.x; T[2 * k + 1] = v.y; }
    const int model = int(T[15]);  // both cameras use the HOST's model (map_utils.h:363-364)
    double A[9], in[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = T[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = T[16 + k];
    const double rho = a.rho[l];
    const double irho = 1.0 / rho;
    const double4 bk = reinterpret_cast<const double4*>(a.lm_pat)[l];
    const double xh = bk.x * irho, yh = bk.y * irho, zh = bk.z * irho;
    const double xt = A[0] * xh + A[1] * yh + A[2] * zh + T[9];
    const double yt = A[3] * xh + A[4] * yh + A[5] * zh + T[10];
    const double zt = A[6] * xh + A[7] * yh + A[8] * zh + T[11];
    double uv[2], Jp[6];
    cam_project<WITH_J>(model, in, xt, yt, zt, uv, Jp);
    const double r0 = a.obs_uv[i] - uv[0];
    const double r1 = a.obs_uv[a.n + i] - uv[1];
    double w;
    cost = huber(r0 * r0 + r1 * r1, a.use_huber, a.huber, &w);
    if (WITH_J) {
      const int64_t n = a.ld;
      double acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        // d r / d X_t = -J_pi
        const double px = -w * Jp[3 * k], py = -w * Jp[3 * k + 1], pz = -w * Jp[3 * k + 2];
        const double ax = px * A[0] + py * A[3] + pz * A[6];
        const double ay = px * A[1] + py * A[4] + pz * A[7];
        const double az = px * A[2] + py * A[5] + pz * A[8];
        double row[13];
        row[0] = ax; row[1] = ay; row[2] = az;
        row[3] = -(ay * zh - az * yh);
        row[4] = -(az * xh - ax * zh);
        row[5] = -(ax * yh - ay * xh);
        // target pose [-p | p x X_t] = row[0..5] x M of the edge (k_edge_prep): neither stored nor formed
        row[12] = -(ax * xh + ay * yh + az * zh) * irho;
        const double rk = w * (k == 0 ? r0 : r1);
        // kGeomPlanes planes per row: host pose 0..5, inverse distance 6, residual 7
        double* Jk = a.J + (int64_t(k) * kGeomPlanes) * n + i;
        Jk[int64_t(stored_plane(13)) * n] = rk;
        Jk[int64_t(stored_plane(12)) * n] = row[12];
#pragma unroll
        for (int c = 0; c < 6; ++c) Jk[int64_t(c) * n] = row[c];
        const double E = row[12];
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] += E * row[c];
        acc[14] += E * E;
        acc[15] += E * rk;
      }
      // target-pose part of the Schur record, (E^T J_h) M: u = A v_upsilon, E^T J_t = [-u | u x t - A v_omega]
      {
        const double u0 = A[0] * acc[0] + A[1] * acc[1] + A[2] * acc[2];
        const double u1 = A[3] * acc[0] + A[4] * acc[1] + A[5] * acc[2];
        const double u2 = A[6] * acc[0] + A[7] * acc[1] + A[8] * acc[2];
        const double w0 = A[0] * acc[3] + A[1] * acc[4] + A[2] * acc[5];
        const double w1 = A[3] * acc[3] + A[4] * acc[4] + A[5] * acc[5];
        const double w2 = A[6] * acc[3] + A[7] * acc[4] + A[8] * acc[5];
        acc[6] = -u0; acc[7] = -u1; acc[8] = -u2;
        acc[9] = (u1 * T[11] - u2 * T[10]) - w0;
        acc[10] = (u2 * T[9] - u0 * T[11]) - w1;
        acc[11] = (u0 * T[10] - u1 * T[9]) - w2;
      }
      double2* o = reinterpret_cast<double2*>(a.orec + 16 * i);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = make_double2(acc[2 * c], acc[2 * c + 1]);
    }
  }
  const double bs = block_sum(cost, s_red);
  if (threadIdx.x == 0) a.block_cost[blockIdx.x] = bs;
}

// Second stage of every scalar reduction: one block, fixed order => the same
// bits on every run (no atomics anywhere).
__global__ void __launch_bounds__(1024) k_reduce_sum(const double* __restrict__ part, int64_t n, double* __restrict__ out) {
  __shared__ double s[32];
  double v = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) v += part[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += s[i];
    *out = t;
  }
}

// Long reductions (one partial per 128 observations: 140 k at config 4) go through kReduceMid
// CTAs first, each summing a fixed contiguous slice (one CTA alone took 37 us per reduction).
__global__ void __launch_bounds__(256) k_reduce_mid(const double* __restrict__ part, int64_t n, double* __restrict__ mid) {
  __shared__ double s[8];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = per * blockIdx.x, i1 = i0 + per < n ? i0 + per : n;
  double v = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) v += part[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s[i];
    mid[blockIdx.x] = t;
  }
}

// Test/diagnostic path: planes in edge order -> [obs][plane] in caller order.
// which = 0: residuals [obs][R]; 1: Jacobians [obs][R][C]  (source planes [R][C+1][ld]).
// edge_M != nullptr (photometric): the target-pose columns 6..11 are rebuilt as (columns 0..5) x M.
__global__ void k_unpermute(int64_t n, int64_t ld, int R, int C, int which, const int64_t* __restrict__ order,
                            const double* __restrict__ src, const int* __restrict__ obs_edge,
                            const double* __restrict__ edge_M, double* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t o = order[i];
  const int P = edge_M ? C + 1 - 6 : C + 1;  // planes per row
  for (int k = 0; k < R; ++k) {
    if (which == 0) {
      dst[o * R + k] = src[(int64_t(k) * P + (edge_M ? stored_plane(C) : C)) * ld + i];
    } else {
      for (int c = 0; c < C; ++c) {
        double v;
        if (edge_M && c >= 6 && c < 12) {
          const double* m = edge_M + 36 * int64_t(obs_edge[i]);
          v = 0.0;
          for (int q = 0; q < 6; ++q) v += src[(int64_t(k) * P + q) * ld + i] * m[6 * q + (c - 6)];
        } else {
          v = src[(int64_t(k) * P + (edge_M ? stored_plane(c) : c)) * ld + i];
        }
        dst[(o * R + k) * C + c] = v;
      }
    }
  }
}

// Selected blocks only: sel_pos[j] = edge-order position of the j-th requested block.
__global__ void k_gather_blocks(int64_t n_sel, int64_t ld, int R, int C, const int64_t* __restrict__ sel_pos,
                                const double* __restrict__ src, const int* __restrict__ obs_edge,
                                const double* __restrict__ edge_M, double* __restrict__ res, double* __restrict__ jac) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_sel) return;
  const int64_t i = sel_pos[j];
  const int P = C + 1 - 6;  // stored planes per row
  const double* m = edge_M + 36 * int64_t(obs_edge[i]);
  for (int k = 0; k < R; ++k) {
    const double* row = src + int64_t(k) * P * ld + i;
    if (res) res[j * R + k] = row[int64_t(stored_plane(C)) * ld];
    if (!jac) continue;
    double hcol[6];
    for (int q = 0; q < 6; ++q) hcol[q] = row[int64_t(q) * ld];
    for (int c = 0; c < C; ++c) {
      double v;
      if (c < 6) v = hcol[c];
      else if (c < 12) { v = 0.0; for (int q = 0; q < 6; ++q) v += hcol[q] * m[6 * q + (c - 6)]; }
      else v = row[int64_t(stored_plane(c)) * ld];
      jac[(j * R + k) * C + c] = v;
    }
  }
}

// Set-up: obs_edge[i] = e and obs_col[i] = edge_col[e] for the observations [edge_ptr[e], edge_ptr[e+1]) of edge e.
__global__ void k_expand_edges(int n_edges, const int64_t* __restrict__ edge_ptr, const int* __restrict__ edge_col,
                               int* __restrict__ obs_edge, int* __restrict__ obs_col) {
  const int e = blockIdx.x;
  if (e >= n_edges) return;
  const int col = edge_col[e];
  for (int64_t i = edge_ptr[e] + threadIdx.x; i < edge_ptr[e + 1]; i += blockDim.x) {
    obs_edge[i] = e;
    obs_col[i] = col;
  }
}

}  // namespace

pba_status launch_expand_edges(Handle* h, const int* edge_col_dev) {
  const Sizes& z = h->sz;
  if (z.n_edges == 0) return PBA_OK;
  PBA_LAUNCH(h, K_INIT_LM, k_expand_edges, dim3(z.n_edges), dim3(128), 0, z.n_edges, h->edge_ptr.p, edge_col_dev,
             h->obs_edge.p, h->obs_col.p);
  return PBA_OK;
}

// Upload-time conversion of the 8-bit keyframes (staged in `images_u8`, n_img
// keyframes of pitch*height bytes) into the quad layout at keyframe `first`.
// `stream` is the set-up thread's own stream (pba_create uploads the keyframes from a helper thread while
// the main thread orders the observations), so the launch is not routed through the handle's statistics.
pba_status launch_build_quads(Handle* h, cudaStream_t stream, const uint8_t* images_u8, int first, int n_img) {
  const Sizes& z = h->sz;
  const int64_t tot = int64_t(z.width) * z.height * n_img;
  if (tot == 0) return PBA_OK;
  k_build_quads<<<dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, stream>>>(z.width, z.height, z.pitch, int64_t(n_img), images_u8,
                                                                         h->quads.p + int64_t(first) * z.image_stride);
  return map_cuda(cudaGetLastError());
}

pba_status launch_init_landmarks(Handle* h) {
  const Sizes& z = h->sz;
  if (z.n_lm == 0) return PBA_OK;
  const int photo = z.mode == PBA_MODE_PHOTOMETRIC;
  PBA_LAUNCH(h, K_INIT_LM, k_init_landmarks, dim3((z.n_lm + 127) / 128), dim3(128), 0, z.n_lm, photo, h->lm_host.p,
             h->lm_uv.p, h->pose_calib.p, h->calib_model.p, h->intr.p, h->quads.p, z.image_stride, z.width,
             z.height, h->lm_pat.p, h->lm_ok.p);
  return PBA_OK;
}

int eval_grid(int64_t n) { return int((n + kEvalThreads - 1) / kEvalThreads); }

pba_status launch_evaluate(Handle* h, bool with_jacobian, const double* poses, const double* affine,
                           const double* rho, double* cost_out) {
  const Sizes& z = h->sz;
  const bool photo = z.mode == PBA_MODE_PHOTOMETRIC;
  if (z.n_edges > 0) {
    PBA_LAUNCH(h, K_EDGE_PREP, k_edge_prep, dim3((z.n_edges + kPrepThreads - 1) / kPrepThreads), dim3(kPrepThreads), 0, z.n_edges, h->edge_h.p,
               h->edge_t.p, poses, photo ? affine : nullptr, h->pose_calib.p, h->calib_model.p, h->intr.p, h->edge_T.p,
               with_jacobian ? h->edge_M.p : nullptr);
  }
  EvalArgs a;
  a.n = z.n_obs; a.ld = z.ld; a.n_lm = z.n_lm;
  a.obs_lm = h->obs_lm.p; a.obs_edge = h->obs_edge.p; a.edge_h = h->edge_h.p; a.edge_t = h->edge_t.p;
  a.pose_calib = h->pose_calib.p; a.calib_model = h->calib_model.p; a.intr = h->intr.p; a.edge_T = h->edge_T.p;
  a.lm_pat = h->lm_pat.p; a.lm_uv = h->lm_uv.p; a.lm_ok = h->lm_ok.p; a.obs_uv = h->obs_uv.p;
  a.quads = h->quads.p; a.image_stride = z.image_stride; a.width = z.width; a.height = z.height; a.pitch = z.width;  // quad rows are packed
  a.rho = rho; a.use_huber = h->opt.use_huber; a.huber = h->opt.huber_parameter;
  a.J = h->J.p; a.orec = h->orec.p; a.block_cost = h->red_ws.p;
  static_assert(kPhotoThreads == kEvalThreads, "block-cost workspace is sized for one CTA width");
  const int grid = eval_grid(z.n_obs);
  // pinhole everywhere: the pattern bearings are recomputed instead of read (PBA_K1_TABLE=1 keeps the table path,
  // for A/B measurements)
  static const bool force_table = getenv("PBA_K1_TABLE") != nullptr;
  const bool rc = h->uniform_model == PBA_CAM_PINHOLE && !force_table;
  if (grid > 0) {
    if (photo) {
      if (with_jacobian) {
        PBA_LAUNCH(h, K_RESJAC, photo_kernel<true>(h->uniform_model, rc), dim3(grid), dim3(kPhotoThreads),
                   rc ? kPhotoSmemBytesRC : kPhotoSmemBytes, a);
      }
      else { PBA_LAUNCH(h, K_COST, photo_kernel<false>(h->uniform_model, rc), dim3(grid), dim3(kPhotoThreads), 0, a); }
    } else {
      static const int gvar = [] { const char* e = getenv("PBA_K1G_VARIANT"); return e ? atoi(e) : 0; }();
      if (with_jacobian) {
        // measured at config 5 (profiles/r02b_variants_geom.txt): unbounded (104 registers) 0.745 ms, 6 CTAs/SM
        // (80 registers) 0.601 ms = 65.6 % of the copy peak, 8 CTAs/SM (64, spills) 0.651, 10 (48, spills) 0.831
        if (gvar == 1) { PBA_LAUNCH(h, K_RESJAC, (k_eval_geom<true, 0>), dim3(grid), dim3(kEvalThreads), 0, a); }
        else if (gvar == 2) { PBA_LAUNCH(h, K_RESJAC, (k_eval_geom<true, 8>), dim3(grid), dim3(kEvalThreads), 0, a); }
        else if (gvar == 3) { PBA_LAUNCH(h, K_RESJAC, (k_eval_geom<true, 10>), dim3(grid), dim3(kEvalThreads), 0, a); }
        else { PBA_LAUNCH(h, K_RESJAC, (k_eval_geom<true, 6>), dim3(grid), dim3(kEvalThreads), 0, a); }
      } else {
        if (gvar == 1) { PBA_LAUNCH(h, K_COST, (k_eval_geom<false, 6>), dim3(grid), dim3(kEvalThreads), 0, a); }
        else if (gvar == 2) { PBA_LAUNCH(h, K_COST, (k_eval_geom<false, 8>), dim3(grid), dim3(kEvalThreads), 0, a); }
        else { PBA_LAUNCH(h, K_COST, (k_eval_geom<false, 0>), dim3(grid), dim3(kEvalThreads), 0, a); }
      }
    }
  }
  launch_reduce_sum(h, h->red_ws.p, int64_t(grid), cost_out);
  return PBA_OK;
}

pba_status launch_unpermute(Handle* h, int which, double* dst) {
  const Sizes& z = h->sz;
  if (z.n_obs == 0) return PBA_OK;
  DevBuf<int64_t> order;
  PBA_CUDA_OK(order.upload(h->obs_order, h->stream));
  PBA_LAUNCH(h, K_UNPERMUTE, k_unpermute, dim3(int((z.n_obs + 127) / 128)), dim3(128), 0, z.n_obs, z.ld, z.R, z.C, which, order.p,
             h->J.p, h->obs_edge.p, h->edge_M.p, dst);
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

pba_status launch_gather_blocks(Handle* h, int64_t n_sel, const int64_t* sel_pos_dev, double* res_dev, double* jac_dev) {
  const Sizes& z = h->sz;
  if (n_sel == 0) return PBA_OK;
  PBA_LAUNCH(h, K_UNPERMUTE, k_gather_blocks, dim3(unsigned((n_sel + 127) / 128)), dim3(128), 0, n_sel, z.ld, z.R, z.C,
             sel_pos_dev, h->J.p, h->obs_edge.p, h->edge_M.p, res_dev, jac_dev);
  return PBA_OK;
}

void launch_reduce_sum(Handle* h, const double* part, int64_t n, double* out) {
  h->stats.begin(K_REDUCE_SUM, h->stream);
  if (n > 4096) {
    k_reduce_mid<<<kReduceMid, 256, 0, h->stream>>>(part, n, h->red_mid.p);
    k_reduce_sum<<<1, 1024, 0, h->stream>>>(h->red_mid.p, kReduceMid, out);
  } else {
    k_reduce_sum<<<1, 1024, 0, h->stream>>>(part, n, out);
  }
  h->stats.end(h->stream);
}

}  // namespace pba
