// synth_scene.h — the synthetic scene's geometry, shared by the CPU generator
// (synth.cpp) and the GPU renderer (synth_gpu.cu).  SURVEY.md §8(d).
#pragma once
#include "pba_math.h"

namespace pba_scene {

PBA_HD double wall_height(double x, double y) { return 5.0 + 0.5 * sin(0.8 * x) * cos(0.6 * y); }

PBA_HD double wall_texture(double x, double y) {
  return 128.0 + 45.0 * sin(7.1 * x + 0.3) * cos(5.3 * y + 1.1) + 30.0 * sin(23.7 * x + 17.9 * y) +
         25.0 * sin(61.3 * x - 43.1 * y + 0.7);
}

// Ground-truth keyframe pose T_w_c (looking at +z).
PBA_HD void kf_pose(int i, double* T) {
  const double d[6] = {0, 0, 0, 0.02 * sin(0.07 * i), 0.02 * cos(0.05 * i), 0.02 * 0.5 * sin(0.03 * i)};
  double q[4], t[3];
  pba::se3_exp(d, q, t);
  T[0] = q[0]; T[1] = q[1]; T[2] = q[2]; T[3] = q[3];
  T[4] = 0.05 * i; T[5] = 0.02 * sin(0.1 * i); T[6] = 0.0;
}

// Intersect the camera ray through unit bearing b (camera frame) with the wall;
// returns the distance along the ray (fixed point, 8 iterations).
PBA_HD double ray_wall(const double* T, const double* b, double* Xw) {
  double d[3];
  pba::quat_rotate(T, b, d);
  const double* c = T + 4;
  double s = (5.0 - c[2]) / d[2];
  for (int it = 0; it < 8; ++it) {
    const double x = c[0] + s * d[0], y = c[1] + s * d[1];
    s = (wall_height(x, y) - c[2]) / d[2];
  }
  Xw[0] = c[0] + s * d[0]; Xw[1] = c[1] + s * d[1]; Xw[2] = c[2] + s * d[2];
  return s;
}

}  // namespace pba_scene
