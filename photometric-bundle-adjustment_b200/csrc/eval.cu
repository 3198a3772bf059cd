// eval.cu — residual / Jacobian evaluation kernels (K0, K1, K2, K8).
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
#include "launch.h"
#include "pba_internal.h"

namespace pba {

namespace {

constexpr int kEvalThreads = 128;

struct EvalArgs {
  int64_t n;
  int n_lm;
  // structure
  const int* obs_lm;
  const int* obs_edge;
  const int* edge_h;
  const int* edge_t;
  const int* pose_calib;
  const int* calib_model;
  const double* intr;
  const double* edge_T;  // [E][16]: A(9) t(3) ea b
  // landmark constants
  const double* lm_pat;
  const uint8_t* lm_ok;
  const double* obs_uv;  // geometric [2][n]
  // images
  const uint8_t* images;
  int64_t image_stride;
  int width, height, pitch;
  // state
  const double* rho;
  // robust loss
  int use_huber;
  double huber;
  // outputs
  double* res;
  double* J;
  double* orec;
  double* block_cost;
};

// K0: per-edge relative pose A = R_t^T R_h, t = R_t^T (t_h - t_t) and the
// target's affine brightness (exp(a), b).
__global__ void k_edge_prep(int n_edges, const int* __restrict__ edge_h, const int* __restrict__ edge_t,
                            const double* __restrict__ poses, const double* __restrict__ affine,
                            double* __restrict__ edge_T) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const double* Th = poses + 7 * edge_h[e];
  const double* Tt = poses + 7 * edge_t[e];
  double Rh[9], Rt[9];
  quat_to_rot(Th, Rh);
  quat_to_rot(Tt, Rt);
  double* o = edge_T + 16 * e;
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
  o[14] = 0.0;
  o[15] = 0.0;
}

// Landmark constants: unit host bearings (reprojection.h:106-107) and, for the
// photometric residual, the bilinear host intensities of the 8-pixel pattern.
__global__ void k_init_landmarks(int n_lm, int photo, const int* __restrict__ lm_host,
                                 const double* __restrict__ lm_uv, const int* __restrict__ pose_calib,
                                 const int* __restrict__ calib_model, const double* __restrict__ intr,
                                 const uint8_t* __restrict__ images, int64_t image_stride, int width, int height,
                                 int pitch, double* __restrict__ lm_pat, uint8_t* __restrict__ lm_ok) {
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
  const uint8_t* img = images + int64_t(h) * image_stride;
  bool ok = true;
  for (int k = 0; k < 8; ++k) {
    const double pu = u + kPatternDev[k][0], pv = v + kPatternDev[k][1];
    double b[3];
    cam_bearing(model, in, pu, pv, b);
    double I = 0.0;
    if (pu >= 0.0 && pv >= 0.0 && pu < double(width - 1) && pv < double(height - 1)) {
      const int x0 = int(floor(pu)), y0 = int(floor(pv));
      const double fx = pu - x0, fy = pv - y0;
      const uint8_t* p = img + int64_t(y0) * pitch + x0;
      const double i00 = p[0], i10 = p[1], i01 = p[pitch], i11 = p[pitch + 1];
      I = (1.0 - fx) * (1.0 - fy) * i00 + fx * (1.0 - fy) * i10 + (1.0 - fx) * fy * i01 + fx * fy * i11;
    } else {
      ok = false;
    }
    double* o = lm_pat + (int64_t(l) * 8 + k) * 4;
    o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = I;
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
// K1 (WITH_J) / K2 (!WITH_J).  Per observation: 8 pattern pixels; pass 1
// warps each pixel into the target, samples intensity + analytic bilinear
// gradient and keeps r_k and p_k = grad^T dpi/dX; the block's Huber weight
// needs all 8 residuals, so pass 2 forms the weighted Jacobian rows.
template <bool WITH_J>
__global__ void __launch_bounds__(kEvalThreads) k_eval_photo(const EvalArgs a) {
  __shared__ double s_red[kEvalThreads / 32];
  const int64_t i = int64_t(blockIdx.x) * kEvalThreads + threadIdx.x;
  double cost = 0.0;
  if (i < a.n) {
    const int e = a.obs_edge[i];
    const int l = a.obs_lm[i];
    const int t = a.edge_t[e];
    const int tc = a.pose_calib[t];
    const int model = a.calib_model[tc];
    const double* T = a.edge_T + 16 * int64_t(e);
    double A[9], tr[3], in[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = T[k];
    tr[0] = T[9]; tr[1] = T[10]; tr[2] = T[11];
    const double ea = T[12], bb = T[13];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = a.intr[8 * tc + k];
    const double rho = a.rho[l];
    const double irho = 1.0 / rho;
    const uint8_t* img = a.images + int64_t(t) * a.image_stride;
    const double4* pat = reinterpret_cast<const double4*>(a.lm_pat) + int64_t(l) * 8;
    bool ok = a.lm_ok[l] != 0;

    double r[8];
    double p[WITH_J ? 8 : 1][3];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double4 bk = pat[k];
      const double xh = bk.x * irho, yh = bk.y * irho, zh = bk.z * irho;
      const double xt = A[0] * xh + A[1] * yh + A[2] * zh + tr[0];
      const double yt = A[3] * xh + A[4] * yh + A[5] * zh + tr[1];
      const double zt = A[6] * xh + A[7] * yh + A[8] * zh + tr[2];
      double uv[2], Jp[6];
      cam_project<WITH_J>(model, in, xt, yt, zt, uv, Jp);
      const double u = uv[0], v = uv[1];
      double rk = 0.0;
      if (u >= 0.0 && v >= 0.0 && u < double(a.width - 1) && v < double(a.height - 1)) {
        const int x0 = int(floor(u)), y0 = int(floor(v));
        const double fx = u - x0, fy = v - y0;
        const uint8_t* q = img + int64_t(y0) * a.pitch + x0;
        const double i00 = __ldg(q), i10 = __ldg(q + 1), i01 = __ldg(q + a.pitch), i11 = __ldg(q + a.pitch + 1);
        const double I = (1.0 - fx) * (1.0 - fy) * i00 + fx * (1.0 - fy) * i10 + (1.0 - fx) * fy * i01 + fx * fy * i11;
        rk = I - (ea * bk.w + bb);
        if (WITH_J) {
          const double gx = (1.0 - fy) * (i10 - i00) + fy * (i11 - i01);
          const double gy = (1.0 - fx) * (i01 - i00) + fx * (i11 - i10);
          p[k][0] = gx * Jp[0] + gy * Jp[3];
          p[k][1] = gx * Jp[1] + gy * Jp[4];
          p[k][2] = gx * Jp[2] + gy * Jp[5];
        }
      } else {
        ok = false;
        if (WITH_J) { p[k][0] = 0.0; p[k][1] = 0.0; p[k][2] = 0.0; }
      }
      r[k] = rk;
      s += rk * rk;
    }
    if (!ok) s = 0.0;  // invalid observation: r = 0, J = 0
    double w;
    cost = huber(s, a.use_huber, a.huber, &w);
    if (!ok) w = 0.0;
    if (WITH_J) {
      double acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.0;
      const int64_t n = a.n;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double4 bk = pat[k];
        const double xh = bk.x * irho, yh = bk.y * irho, zh = bk.z * irho;
        const double xt = A[0] * xh + A[1] * yh + A[2] * zh + tr[0];
        const double yt = A[3] * xh + A[4] * yh + A[5] * zh + tr[1];
        const double zt = A[6] * xh + A[7] * yh + A[8] * zh + tr[2];
        const double px = w * p[k][0], py = w * p[k][1], pz = w * p[k][2];
        // a = p A   (d r / d upsilon_h);   d r / d omega_h = -(a x X_h)
        const double ax = px * A[0] + py * A[3] + pz * A[6];
        const double ay = px * A[1] + py * A[4] + pz * A[7];
        const double az = px * A[2] + py * A[5] + pz * A[8];
        double row[15];
        row[0] = ax; row[1] = ay; row[2] = az;
        row[3] = -(ay * zh - az * yh);
        row[4] = -(az * xh - ax * zh);
        row[5] = -(ax * yh - ay * xh);
        // target pose: [-p | p x X_t]
        row[6] = -px; row[7] = -py; row[8] = -pz;
        row[9] = py * zt - pz * yt;
        row[10] = pz * xt - px * zt;
        row[11] = px * yt - py * xt;
        // affine (a_t, b_t)
        row[12] = -w * ea * bk.w;
        row[13] = -w;
        // inverse distance: -(a . X_h) / rho
        row[14] = -(ax * xh + ay * yh + az * zh) * irho;
        const double rk = w * r[k];
        a.res[int64_t(k) * n + i] = rk;
        double* Jk = a.J + (int64_t(k) * 15) * n + i;
#pragma unroll
        for (int c = 0; c < 15; ++c) Jk[int64_t(c) * n] = row[c];
        const double E = row[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) acc[c] += E * row[c];
        acc[14] += E * E;
        acc[15] += E * rk;
      }
      double2* o = reinterpret_cast<double2*>(a.orec + 16 * i);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = make_double2(acc[2 * c], acc[2 * c + 1]);
    }
  }
  const double bs = block_sum(cost, s_red);
  if (threadIdx.x == 0) a.block_cost[blockIdx.x] = bs;
}

// -------------------------------------------------------------- geometric --
// reprojection.h:82-112: r = z_t - pi_t(T_t^-1 T_h (b / rho)).  Both cameras
// use the HOST's model (the reference passes the host's model name for both,
// map_utils.h:363-364) with the target's intrinsic values.
template <bool WITH_J>
__global__ void __launch_bounds__(kEvalThreads) k_eval_geom(const EvalArgs a) {
  __shared__ double s_red[kEvalThreads / 32];
  const int64_t i = int64_t(blockIdx.x) * kEvalThreads + threadIdx.x;
  double cost = 0.0;
  if (i < a.n) {
    const int e = a.obs_edge[i];
    const int l = a.obs_lm[i];
    const int t = a.edge_t[e];
    const int tc = a.pose_calib[t];
    const int model = a.calib_model[a.pose_calib[a.edge_h[e]]];
    const double* T = a.edge_T + 16 * int64_t(e);
    double A[9], in[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = T[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = a.intr[8 * tc + k];
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
      const int64_t n = a.n;
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
        row[6] = -px; row[7] = -py; row[8] = -pz;
        row[9] = py * zt - pz * yt;
        row[10] = pz * xt - px * zt;
        row[11] = px * yt - py * xt;
        row[12] = -(ax * xh + ay * yh + az * zh) * irho;
        const double rk = w * (k == 0 ? r0 : r1);
        a.res[int64_t(k) * n + i] = rk;
        double* Jk = a.J + (int64_t(k) * 13) * n + i;
#pragma unroll
        for (int c = 0; c < 13; ++c) Jk[int64_t(c) * n] = row[c];
        const double E = row[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] += E * row[c];
        acc[14] += E * E;
        acc[15] += E * rk;
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

// K8: model_cost_change = -(J d)^T (r + J d / 2) (trust_region_minimizer.cc:414-427),
// with the unscaled Jacobian and the unscaled tangent step d.
template <int R, int C>
__global__ void __launch_bounds__(256) k_model_cost(int64_t n, const int* __restrict__ obs_lm,
                                                     const int* __restrict__ obs_edge, const int* __restrict__ edge_h,
                                                     const int* __restrict__ edge_t, const int* __restrict__ slot,
                                                     const double* __restrict__ J, const double* __restrict__ res,
                                                     const double* __restrict__ d_cam, const double* __restrict__ d_rho,
                                                     double* __restrict__ block_out) {
  __shared__ double s_red[8];
  constexpr int CD = C - 7;  // 6 geometric, 8 photometric
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    const int e = obs_edge[i];
    const int hs = slot[edge_h[e]], ts = slot[edge_t[e]];
    double d[C];
#pragma unroll
    for (int c = 0; c < 6; ++c) d[c] = hs >= 0 ? d_cam[hs * CD + c] : 0.0;
#pragma unroll
    for (int c = 0; c < CD; ++c) d[6 + c] = ts >= 0 ? d_cam[ts * CD + c] : 0.0;
    d[C - 1] = d_rho[obs_lm[i]];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double m = 0.0;
#pragma unroll
      for (int c = 0; c < C; ++c) m += J[(int64_t(k) * C + c) * n + i] * d[c];
      acc += -m * (res[int64_t(k) * n + i] + 0.5 * m);
    }
  }
  const double bs = block_sum(acc, s_red);
  if (threadIdx.x == 0) block_out[blockIdx.x] = bs;
}

// Test/diagnostic path: planes in edge order -> [obs][plane] in caller order.
__global__ void k_unpermute(int64_t n, int planes, const int64_t* __restrict__ order,
                            const double* __restrict__ src, double* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t o = order[i];
  for (int p = 0; p < planes; ++p) dst[o * planes + p] = src[int64_t(p) * n + i];
}

}  // namespace

pba_status launch_init_landmarks(Handle* h) {
  const Sizes& z = h->sz;
  if (z.n_lm == 0) return PBA_OK;
  const int photo = z.mode == PBA_MODE_PHOTOMETRIC;
  PBA_LAUNCH(h, K_INIT_LM, k_init_landmarks, dim3((z.n_lm + 127) / 128), dim3(128), 0, z.n_lm, photo, h->lm_host.p,
             h->lm_uv.p, h->pose_calib.p, h->calib_model.p, h->intr.p, h->images.p, z.image_stride, z.width,
             z.height, z.pitch, h->lm_pat.p, h->lm_ok.p);
  return PBA_OK;
}

int eval_grid(int64_t n) { return int((n + kEvalThreads - 1) / kEvalThreads); }

pba_status launch_evaluate(Handle* h, bool with_jacobian, const double* poses, const double* affine,
                           const double* rho, int cost_slot) {
  const Sizes& z = h->sz;
  const bool photo = z.mode == PBA_MODE_PHOTOMETRIC;
  if (z.n_edges > 0) {
    PBA_LAUNCH(h, K_EDGE_PREP, k_edge_prep, dim3((z.n_edges + 127) / 128), dim3(128), 0, z.n_edges, h->edge_h.p,
               h->edge_t.p, poses, photo ? affine : nullptr, h->edge_T.p);
  }
  EvalArgs a;
  a.n = z.n_obs; a.n_lm = z.n_lm;
  a.obs_lm = h->obs_lm.p; a.obs_edge = h->obs_edge.p; a.edge_h = h->edge_h.p; a.edge_t = h->edge_t.p;
  a.pose_calib = h->pose_calib.p; a.calib_model = h->calib_model.p; a.intr = h->intr.p; a.edge_T = h->edge_T.p;
  a.lm_pat = h->lm_pat.p; a.lm_ok = h->lm_ok.p; a.obs_uv = h->obs_uv.p;
  a.images = h->images.p; a.image_stride = z.image_stride; a.width = z.width; a.height = z.height; a.pitch = z.pitch;
  a.rho = rho; a.use_huber = h->opt.use_huber; a.huber = h->opt.huber_parameter;
  a.res = h->res.p; a.J = h->J.p; a.orec = h->orec.p; a.block_cost = h->red_ws.p;
  const int grid = eval_grid(z.n_obs);
  if (grid > 0) {
    if (photo) {
      if (with_jacobian) { PBA_LAUNCH(h, K_RESJAC, k_eval_photo<true>, dim3(grid), dim3(kEvalThreads), 0, a); }
      else { PBA_LAUNCH(h, K_COST, k_eval_photo<false>, dim3(grid), dim3(kEvalThreads), 0, a); }
    } else {
      if (with_jacobian) { PBA_LAUNCH(h, K_RESJAC, k_eval_geom<true>, dim3(grid), dim3(kEvalThreads), 0, a); }
      else { PBA_LAUNCH(h, K_COST, k_eval_geom<false>, dim3(grid), dim3(kEvalThreads), 0, a); }
    }
  }
  PBA_LAUNCH(h, K_REDUCE_SUM, k_reduce_sum, dim3(1), dim3(1024), 0, h->red_ws.p, int64_t(grid), h->scalars.p + cost_slot);
  return PBA_OK;
}

pba_status launch_model_cost(Handle* h) {
  const Sizes& z = h->sz;
  const int grid = int((z.n_obs + 255) / 256);
  if (grid > 0) {
    if (z.mode == PBA_MODE_PHOTOMETRIC) {
      PBA_LAUNCH(h, K_MODEL_COST, (k_model_cost<8, 15>), dim3(grid), dim3(256), 0, z.n_obs, h->obs_lm.p, h->obs_edge.p,
                 h->edge_h.p, h->edge_t.p, h->d_slot.p, h->J.p, h->res.p, h->d_cam.p, h->d_rho.p, h->red_ws.p);
    } else {
      PBA_LAUNCH(h, K_MODEL_COST, (k_model_cost<2, 13>), dim3(grid), dim3(256), 0, z.n_obs, h->obs_lm.p, h->obs_edge.p,
                 h->edge_h.p, h->edge_t.p, h->d_slot.p, h->J.p, h->res.p, h->d_cam.p, h->d_rho.p, h->red_ws.p);
    }
  }
  PBA_LAUNCH(h, K_REDUCE_SUM, k_reduce_sum, dim3(1), dim3(1024), 0, h->red_ws.p, int64_t(grid), h->scalars.p + S_MODEL);
  return PBA_OK;
}

pba_status launch_unpermute(Handle* h, const double* src_planes, int planes, double* dst) {
  const Sizes& z = h->sz;
  if (z.n_obs == 0) return PBA_OK;
  DevBuf<int64_t> order;
  PBA_CUDA_OK(order.upload(h->obs_order, h->stream));
  PBA_LAUNCH(h, K_UNPERMUTE, k_unpermute, dim3(int((z.n_obs + 127) / 128)), dim3(128), 0, z.n_obs, planes, order.p,
             src_planes, dst);
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

void launch_reduce_sum(Handle* h, const double* part, int64_t n, double* out) {
  h->stats.begin(K_REDUCE_SUM, h->stream);
  k_reduce_sum<<<1, 1024, 0, h->stream>>>(part, n, out);
  h->stats.end(h->stream);
}

}  // namespace pba
