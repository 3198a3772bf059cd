// pba_math.h — host/device fp64 geometry shared by the kernels and the host
// driver: SE(3) on unit quaternion + translation (semantics of the reference's
// Sophus 1.1.0, thirdparty/Sophus/sophus/se3.hpp / so3.hpp) and the camera
// models of include/visnav/camera_models.h with closed-form projection
// Jacobians.  Everything is analytic — there is no autodiff
// on the device.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pba.h"

#if defined(__CUDACC__)
#define PBA_HD __host__ __device__ __forceinline__
#else
#define PBA_HD inline
#endif

namespace pba {

// Pose layout (Sophus): q = (x, y, z, w) at [0..3], t at [4..6].

// R (row-major) of a unit quaternion.
PBA_HD void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double xx = x * x, yy = y * y, zz = z * z;
  const double xy = x * y, xz = x * z, yz = y * z;
  const double wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz);       R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz);       R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy);       R[7] = 2.0 * (yz + wx);       R[8] = 1.0 - 2.0 * (xx + yy);
}

// p' = q p q*  (so3.hpp:362-371: uv = 2 (v x p); p + w uv + v x uv).
PBA_HD void quat_rotate(const double* q, const double* p, double* out) {
  double ux = q[1] * p[2] - q[2] * p[1];
  double uy = q[2] * p[0] - q[0] * p[2];
  double uz = q[0] * p[1] - q[1] * p[0];
  ux += ux; uy += uy; uz += uz;
  out[0] = p[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  out[1] = p[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  out[2] = p[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

// a * b for quaternions stored (x,y,z,w)  (so3.hpp:329-345).
PBA_HD void quat_mul(const double* a, const double* b, double* o) {
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

// SE3::exp (se3.hpp:763-784, so3.hpp:585-621): delta = (upsilon, omega).
PBA_HD void se3_exp(const double* d, double* q, double* t) {
  const double ox = d[3], oy = d[4], oz = d[5];
  const double theta_sq = ox * ox + oy * oy + oz * oz;
  const double eps = 1e-10;
  double theta, imag, real;
  if (theta_sq < eps * eps) {
    theta = 0.0;
    const double t4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * t4;
  } else {
    theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = real;
  // V = I + a*Omega + b*Omega^2 (or R itself for tiny theta).
  double V[9];
  if (theta < eps) {
    quat_to_rot(q, V);
  } else {
    const double a = (1.0 - cos(theta)) / theta_sq;
    const double b = (theta - sin(theta)) / (theta_sq * theta);
    // Omega = hat(omega); Omega^2 = omega omega^T - theta^2 I
    V[0] = 1.0 + b * (ox * ox - theta_sq); V[1] = -a * oz + b * ox * oy;          V[2] = a * oy + b * ox * oz;
    V[3] = a * oz + b * ox * oy;           V[4] = 1.0 + b * (oy * oy - theta_sq); V[5] = -a * ox + b * oy * oz;
    V[6] = -a * oy + b * ox * oz;          V[7] = a * ox + b * oy * oz;           V[8] = 1.0 + b * (oz * oz - theta_sq);
  }
  t[0] = V[0] * d[0] + V[1] * d[1] + V[2] * d[2];
  t[1] = V[3] * d[0] + V[4] * d[1] + V[5] * d[2];
  t[2] = V[6] * d[0] + V[7] * d[1] + V[8] * d[2];
}

// LocalParameterizationSE3::Plus: T * exp(delta)
// (local_parameterization_se3.hpp:44-51; the SO3 product renormalises the
// quaternion, so3.hpp:482-489).
PBA_HD void se3_plus(const double* T, const double* d, double* out) {
  double qe[4], te[3], q[4], rt[3];
  se3_exp(d, qe, te);
  quat_mul(T, qe, q);
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  quat_rotate(T, te, rt);
  out[0] = q[0] / n; out[1] = q[1] / n; out[2] = q[2] / n; out[3] = q[3] / n;
  out[4] = T[4] + rt[0]; out[5] = T[5] + rt[1]; out[6] = T[6] + rt[2];
}

// ---------------------------------------------------------------- cameras --
// intr = [fx fy cx cy p1 p2 p3 p4] (camera_models.h:119-123).

// project (camera_models.h:75-91, :144-164, :226-249, :316-351).  No domain
// checks, like the reference.  If J != nullptr also d(u,v)/d(x,y,z), row-major 2x3.
template <bool WITH_J>
PBA_HD void cam_project(int model, const double* intr, double x, double y, double z,
                        double* uv, double* J) {
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  switch (model) {
    case PBA_CAM_PINHOLE: {
      const double iz = 1.0 / z;
      uv[0] = fx * x * iz + cx;
      uv[1] = fy * y * iz + cy;
      if (WITH_J) {
        J[0] = fx * iz; J[1] = 0.0;     J[2] = -fx * x * iz * iz;
        J[3] = 0.0;     J[4] = fy * iz; J[5] = -fy * y * iz * iz;
      }
    } break;
    case PBA_CAM_DS: {
      const double xi = intr[4], alpha = intr[5];
      const double d1 = sqrt(x * x + y * y + z * z);
      const double k = xi * d1 + z;
      const double d2 = sqrt(x * x + y * y + k * k);
      const double n = alpha * d2 + (1.0 - alpha) * k;
      const double in = 1.0 / n;
      uv[0] = fx * x * in + cx;
      uv[1] = fy * y * in + cy;
      if (WITH_J) {
        const double c = alpha * (1.0 + xi * k / d1) / d2 + (1.0 - alpha) * xi / d1;
        const double nx = x * c, ny = y * c;
        const double nz = (xi * z / d1 + 1.0) * (alpha * k / d2 + 1.0 - alpha);
        const double in2 = in * in;
        J[0] = fx * (in - x * nx * in2); J[1] = -fx * x * ny * in2;       J[2] = -fx * x * nz * in2;
        J[3] = -fy * y * nx * in2;       J[4] = fy * (in - y * ny * in2); J[5] = -fy * y * nz * in2;
      }
    } break;
    case PBA_CAM_KB4: {
      const double k1 = intr[4], k2 = intr[5], k3 = intr[6], k4 = intr[7];
      const double r2 = x * x + y * y;
      const double r = sqrt(r2);
      if (r == 0.0) {
        uv[0] = cx; uv[1] = cy;
        if (WITH_J) { J[0] = J[1] = J[2] = J[3] = J[4] = J[5] = 0.0; }
        break;
      }
      const double th = atan2(r, z);
      const double t2 = th * th;
      const double d = th + th * t2 * (k1 + t2 * (k2 + t2 * (k3 + t2 * k4)));
      const double ir = 1.0 / r;
      uv[0] = fx * d * x * ir + cx;
      uv[1] = fy * d * y * ir + cy;
      if (WITH_J) {
        const double dd = 1.0 + t2 * (3.0 * k1 + t2 * (5.0 * k2 + t2 * (7.0 * k3 + t2 * 9.0 * k4)));
        const double q = r2 + z * z;
        const double thx = x * z / (r * q), thy = y * z / (r * q), thz = -r / q;
        const double ir3 = ir * ir * ir;
        J[0] = fx * (dd * thx * x * ir + d * (ir - x * x * ir3));
        J[1] = fx * (dd * thy * x * ir - d * x * y * ir3);
        J[2] = fx * (dd * thz * x * ir);
        J[3] = fy * (dd * thx * y * ir - d * x * y * ir3);
        J[4] = fy * (dd * thy * y * ir + d * (ir - y * y * ir3));
        J[5] = fy * (dd * thz * y * ir);
      }
    } break;
    default: {  // PBA_CAM_EUCM
      const double alpha = intr[4], beta = intr[5];
      const double d = sqrt(beta * (x * x + y * y) + z * z);
      const double n = alpha * d + (1.0 - alpha) * z;
      const double in = 1.0 / n;
      uv[0] = fx * x * in + cx;
      uv[1] = fy * y * in + cy;
      if (WITH_J) {
        const double nx = alpha * beta * x / d, ny = alpha * beta * y / d;
        const double nz = alpha * z / d + (1.0 - alpha);
        const double in2 = in * in;
        J[0] = fx * (in - x * nx * in2); J[1] = -fx * x * ny * in2;       J[2] = -fx * x * nz * in2;
        J[3] = -fy * y * nx * in2;       J[4] = fy * (in - y * ny * in2); J[5] = -fy * y * nz * in2;
      }
    } break;
  }
}

// unproject (camera_models.h:93-107, :166-188, :251-277, :353-379).  Raw model
// output; callers normalise like reprojection.h:106-107 does.
PBA_HD void cam_unproject(int model, const double* intr, double u, double v, double* out) {
  const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const double mx = (u - cx) / fx, my = (v - cy) / fy;
  switch (model) {
    case PBA_CAM_PINHOLE: {
      const double n = sqrt(mx * mx + my * my + 1.0);
      out[0] = mx / n; out[1] = my / n; out[2] = 1.0 / n;
    } break;
    case PBA_CAM_DS: {
      const double xi = intr[4], alpha = intr[5];
      const double r2 = mx * mx + my * my;
      const double mz = (1.0 - alpha * alpha * r2) /
                        (alpha * sqrt(1.0 - (2.0 * alpha - 1.0) * r2) + 1.0 - alpha);
      const double f = (mz * xi + sqrt(mz * mz + (1.0 - xi * xi) * r2)) / (mz * mz + r2);
      out[0] = f * mx; out[1] = f * my; out[2] = f * mz - xi;
    } break;
    case PBA_CAM_KB4: {
      const double k1 = intr[4], k2 = intr[5], k3 = intr[6], k4 = intr[7];
      const double ru = sqrt(mx * mx + my * my);
      if (ru == 0.0) { out[0] = 0.0; out[1] = 0.0; out[2] = 1.0; break; }
      double th = 0.0;  // exactly 5 Newton steps from 0 (camera_models.h:372-375)
      for (int i = 0; i < 5; ++i) {
        const double t2 = th * th;
        const double f = th + th * t2 * (k1 + t2 * (k2 + t2 * (k3 + t2 * k4))) - ru;
        const double df = 1.0 + t2 * (3.0 * k1 + t2 * (5.0 * k2 + t2 * (7.0 * k3 + t2 * 9.0 * k4)));
        th = th - f / df;
      }
      const double s = sin(th);
      out[0] = s * mx / ru; out[1] = s * my / ru; out[2] = cos(th);
    } break;
    default: {  // EUCM
      const double alpha = intr[4], beta = intr[5];
      const double r2 = mx * mx + my * my;
      const double mz = (1.0 - beta * alpha * alpha * r2) /
                        (alpha * sqrt(1.0 - (2.0 * alpha - 1.0) * beta * r2) + (1.0 - alpha));
      const double n = sqrt(mx * mx + my * my + mz * mz);
      out[0] = mx / n; out[1] = my / n; out[2] = mz / n;
    } break;
  }
}

// Unit host bearing: normalize(unproject(z)) (reprojection.h:106-107).
PBA_HD void cam_bearing(int model, const double* intr, double u, double v, double* b) {
  cam_unproject(model, intr, u, v, b);
  const double n = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  b[0] /= n; b[1] /= n; b[2] /= n;
}

// DSO 8-pixel residual pattern.
#if defined(__CUDACC__)
__device__ __constant__ static const int kPatternDev[8][2] = {
    {0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};
#endif
static const int kPatternHost[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0},
                                       {0, 0},  {2, 0},   {-1, 1}, {0, 2}};

}  // namespace pba
