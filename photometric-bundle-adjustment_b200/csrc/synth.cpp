// synth.cpp — deterministic synthetic BA scenes.
//
// A textured height-field wall z = 5 + 0.5 sin(0.8x) cos(0.6y) seen by a
// keyframe trajectory moving along +x; landmarks lie ON the surface so host and
// target patches are photo-consistent; every landmark is hosted in one keyframe
// and observed by a window of the following keyframes.  Used by tests and
// bench.py for both the CUDA engine and the CPU oracle / reference (same bytes
// go to both).  Host-only code (OpenMP), no CUDA: libpba_synth.so.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <random>
#include <vector>

#include "pba_math.h"
#include "pba_synth.h"
#include "synth_scene.h"

using namespace pba_scene;

namespace {

struct Rng {
  std::mt19937_64 g;
  explicit Rng(uint64_t seed) : g(seed) {}
  double uniform() { return (g() >> 11) * (1.0 / 9007199254740992.0); }
  // Box-Muller on our own uniforms: bit-reproducible across libstdc++ versions.
  double normal() {
    double u1 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    const double u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
};

}  // namespace

extern "C" {

void pba_synth_default_params(pba_synth_params* p, int mode, int n_kf, int n_pts, int model) {
  memset(p, 0, sizeof(*p));
  p->mode = mode; p->n_kf = n_kf; p->n_pts = n_pts; p->model = model;
  p->width = 752; p->height = 480;
  p->min_len = 8; p->max_len = 12;
  p->seed_pix = 1234; p->seed_vis = 99; p->seed_noise = 42;
  if (mode == PBA_MODE_GEOMETRIC) {
    p->pose_sigma = 0.01; p->rho_sigma = 0.05; p->pixel_sigma = 0.3;
  } else {
    p->pose_sigma = 0.002; p->rho_sigma = 0.02; p->pixel_sigma = 0.0;
    p->affine_a_sigma = 0.02; p->affine_b_sigma = 2.0;
  }
  const double pin[8] = {370.34, 370.34, 375.5, 239.5, 0, 0, 0, 0};
  const double ds[8] = {370.34, 370.34, 375.5, 239.5, -0.2, 0.55, 0, 0};
  const double kb4[8] = {379.045, 379.008, 375.5, 239.5, 0.00693023, -0.0013828, -0.000272596, -0.000452646};
  const double eucm[8] = {370.34, 370.34, 375.5, 239.5, 0.55, 1.05, 0, 0};
  const double* src = model == PBA_CAM_PINHOLE ? pin : model == PBA_CAM_DS ? ds : model == PBA_CAM_KB4 ? kb4 : eucm;
  memcpy(p->intrinsics, src, sizeof(pin));
}

// Visibility structure only: host index and window length per landmark.
static void visibility(const pba_synth_params* p, std::vector<int>& host, std::vector<int>& ntgt) {
  Rng rv(p->seed_vis);
  host.resize(p->n_pts); ntgt.resize(p->n_pts);
  const int max_host = p->n_kf - 2;  // at least one target after the host
  for (int l = 0; l < p->n_pts; ++l) {
    int h = int((int64_t(l) * (max_host + 1)) / p->n_pts);
    if (h > max_host) h = max_host;
    const int len = p->min_len + int(rv.uniform() * (p->max_len - p->min_len + 1));
    int nt = len - 1;
    if (h + nt > p->n_kf - 1) nt = p->n_kf - 1 - h;
    if (nt < 1) nt = 1;
    host[l] = h; ntgt[l] = nt;
  }
}

int64_t pba_synth_count_obs(const pba_synth_params* p) {
  std::vector<int> host, ntgt;
  visibility(p, host, ntgt);
  int64_t n = 0;
  for (int v : ntgt) n += v;
  return n;
}

// Fill the flat problem arrays (caller-allocated):
//   poses_gt/poses [n_kf*7], pose_fixed [n_kf], inv_depth_gt/inv_depth [n_pts],
//   lm_host [n_pts], lm_host_uv [n_pts*2], lm_obs_ptr [n_pts+1],
//   obs_target [n_obs], obs_uv [n_obs*2] (geometric; may be NULL),
//   affine [n_kf*2] (photometric; may be NULL).
int pba_synth_generate(const pba_synth_params* p, double* poses_gt, double* poses, uint8_t* pose_fixed,
                       double* inv_depth_gt, double* inv_depth, int32_t* lm_host, double* lm_host_uv,
                       int64_t* lm_obs_ptr, int32_t* obs_target, double* obs_uv, double* affine) {
  if (p->n_kf < 3 || p->n_pts < 1) return 1;
  std::vector<int> host, ntgt;
  visibility(p, host, ntgt);
  for (int i = 0; i < p->n_kf; ++i) kf_pose(i, poses_gt + 7 * i);

  Rng rp(p->seed_pix);
  // host pixels inside the central region so the 8-12 frame window stays in view
  const double u0 = 200.0 / 752.0 * p->width, u1 = 552.0 / 752.0 * p->width;
  const double v0 = 120.0 / 480.0 * p->height, v1 = 360.0 / 480.0 * p->height;
  lm_obs_ptr[0] = 0;
  std::vector<double> Xw(size_t(p->n_pts) * 3);
  for (int l = 0; l < p->n_pts; ++l) {
    const double u = u0 + (u1 - u0) * rp.uniform();
    const double v = v0 + (v1 - v0) * rp.uniform();
    lm_host_uv[2 * l] = u; lm_host_uv[2 * l + 1] = v;
    lm_host[l] = host[l];
    double b[3];
    pba::cam_bearing(p->model, p->intrinsics, u, v, b);
    const double s = ray_wall(poses_gt + 7 * host[l], b, &Xw[3 * l]);
    inv_depth_gt[l] = 1.0 / s;
    lm_obs_ptr[l + 1] = lm_obs_ptr[l] + ntgt[l];
  }
  Rng rn(p->seed_noise);
  // perturbed initial state; first two keyframes fixed (src/sfm.cpp:1903)
  for (int i = 0; i < p->n_kf; ++i) {
    pose_fixed[i] = i < 2;
    if (i < 2) {
      memcpy(poses + 7 * i, poses_gt + 7 * i, 56);
      if (affine) { affine[2 * i] = 0; affine[2 * i + 1] = 0; }
    } else {
      double d[6];
      for (int k = 0; k < 6; ++k) d[k] = p->pose_sigma * rn.normal();
      pba::se3_plus(poses_gt + 7 * i, d, poses + 7 * i);
      if (affine) {
        affine[2 * i] = p->affine_a_sigma * rn.normal();
        affine[2 * i + 1] = p->affine_b_sigma * rn.normal();
      }
    }
  }
  for (int l = 0; l < p->n_pts; ++l) inv_depth[l] = inv_depth_gt[l] / (1.0 + p->rho_sigma * rn.normal());
  for (int l = 0; l < p->n_pts; ++l) {
    for (int j = 0; j < ntgt[l]; ++j) {
      const int64_t o = lm_obs_ptr[l] + j;
      const int t = host[l] + 1 + j;
      obs_target[o] = t;
      if (obs_uv) {
        // true projection into the target + pixel noise
        const double* T = poses_gt + 7 * t;
        const double qc[4] = {-T[0], -T[1], -T[2], T[3]};
        const double d[3] = {Xw[3 * l] - T[4], Xw[3 * l + 1] - T[5], Xw[3 * l + 2] - T[6]};
        double Xt[3], uv[2];
        pba::quat_rotate(qc, d, Xt);
        pba::cam_project<false>(p->model, p->intrinsics, Xt[0], Xt[1], Xt[2], uv, nullptr);
        obs_uv[2 * o] = uv[0] + p->pixel_sigma * rn.normal();
        obs_uv[2 * o + 1] = uv[1] + p->pixel_sigma * rn.normal();
      }
    }
  }
  return 0;
}

// Render keyframes [first, first+count) of the ground-truth trajectory into
// `images` (count * pitch * height bytes, 8-bit grey).
int pba_synth_render(const pba_synth_params* p, int first, int count, int pitch, uint8_t* images) {
  const int w = p->width, h = p->height;
  // per-pixel unit bearings are the same for every keyframe
  std::vector<double> bear(size_t(w) * h * 3);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) pba::cam_bearing(p->model, p->intrinsics, x, y, &bear[(size_t(y) * w + x) * 3]);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int i = 0; i < count; ++i) {
    for (int y = 0; y < h; ++y) {
      double T[7];
      kf_pose(first + i, T);
      uint8_t* row = images + (size_t(i) * h + y) * pitch;
      for (int x = 0; x < w; ++x) {
        double X[3];
        ray_wall(T, &bear[(size_t(y) * w + x) * 3], X);
        double f = wall_texture(X[0], X[1]);
        f = f < 0.0 ? 0.0 : (f > 255.0 ? 255.0 : f);
        row[x] = uint8_t(floor(f + 0.5));
      }
    }
  }
  return 0;
}

}  // extern "C"
