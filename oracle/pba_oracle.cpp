// oracle/pba_oracle.cpp — CPU restatement ("port") of the reference BA path.
//
// TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs as the CHECKER.  The product
// (libpba_b200.so) never links, loads or calls anything in oracle/.
//
// Parity status: PINNED.  This restatement is checked against the real
// reference (oracle/_ref/libpba_ref.so = unmodified visnav headers + vendored
// Ceres 2.0.0, built by oracle/ref/Makefile) in tests/test_oracle_vs_reference.py
// and against the golden vectors that library produced (tests/golden/, made by
// tests/golden/make_golden.py).  The reference repo itself ships no tests or
// golden vectors for this path.
//
// It deliberately follows the REFERENCE'S method, not the CUDA engine's:
// forward-mode dual numbers through quaternion SE(3) algebra (what Ceres
// AutoDiff + Sophus do), global 2x7 Jacobians multiplied by the local
// parameterisation Jacobian, then Ceres' robust correction, Jacobi scaling,
// LM damping, Schur elimination and trust-region logic.  The CUDA path uses
// closed-form Jacobians instead, so agreement is a real check.
//
// Restated reference locations (all under /root/reference):
//   functor ............ include/visnav/reprojection.h:82-112
//   camera models ...... include/visnav/camera_models.h:75-107,144-188,226-277,316-420
//   SE3 / SO3 .......... thirdparty/Sophus/sophus/se3.hpp:135-211,308-328,763-784
//                        thirdparty/Sophus/sophus/so3.hpp:297-303,329-371,482-489,585-621
//   local param ........ include/visnav/local_parameterization_se3.hpp:44-64
//   problem build ...... include/visnav/map_utils.h:327-383
//   residual block ..... thirdparty/ceres-solver/internal/ceres/residual_block.cc:69-198
//   Huber / corrector .. internal/ceres/loss_function.cc:48-62, corrector.cc:82-130
//   LM loop ............ internal/ceres/trust_region_minimizer.cc:67-826,
//                        levenberg_marquardt_strategy.cc:66-162,
//                        trust_region_step_evaluator.cc:52-112
//   Schur .............. internal/ceres/schur_eliminator_impl.h:177-375
//   photometric spec ... SURVEY.md §8(a-P) (the snapshot has no photometric code)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <limits>
#include <vector>

#include "pba.h"

namespace {

// ------------------------------------------------------------------ Jet ----
// Minimal forward-mode dual number (include/ceres/jet.h semantics).
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
template <int N> Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; const double iy = 1.0 / y.a; r.a = x.a * iy;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * iy; return r;
}
template <int N> Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N> Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N> Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N> Jet<N> operator-(double s, const Jet<N>& x) { return (-x) + s; }
template <int N> Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> Jet<N> operator/(const Jet<N>& x, double s) { return x * (1.0 / s); }
template <int N> Jet<N> operator/(double s, const Jet<N>& x) { return Jet<N>(s) / x; }
template <int N> Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N> bool operator==(const Jet<N>& x, double s) { return x.a == s; }
template <int N> Jet<N> jsqrt(const Jet<N>& x) { Jet<N> r; r.a = sqrt(x.a); const double d = 0.5 / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> Jet<N> jexp(const Jet<N>& x) { Jet<N> r; r.a = exp(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * r.a; return r; }
template <int N> Jet<N> jsin(const Jet<N>& x) { Jet<N> r; r.a = sin(x.a); const double c = cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * c; return r; }
template <int N> Jet<N> jcos(const Jet<N>& x) { Jet<N> r; r.a = cos(x.a); const double s = -sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> Jet<N> jatan2(const Jet<N>& y, const Jet<N>& x) {
  Jet<N> r; r.a = atan2(y.a, x.a); const double d = 1.0 / (x.a * x.a + y.a * y.a);
  for (int i = 0; i < N; ++i) r.v[i] = (x.a * y.v[i] - y.a * x.v[i]) * d; return r;
}
inline double jsqrt(double x) { return sqrt(x); }
inline double jexp(double x) { return exp(x); }
inline double jsin(double x) { return sin(x); }
inline double jcos(double x) { return cos(x); }
inline double jatan2(double y, double x) { return atan2(y, x); }
inline double scalar_of(double x) { return x; }
template <int N> double scalar_of(const Jet<N>& x) { return x.a; }

// ------------------------------------------------------------ SE3 (Sophus) -
template <class T> struct V3 { T x, y, z; };
template <class T> struct Q4 { T x, y, z, w; };
template <class T> struct SE3 { Q4<T> q; V3<T> t; };

template <class T> Q4<T> normalized(const Q4<T>& q) {  // so3.hpp:297-303
  const T n = jsqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Q4<T>{q.x / n, q.y / n, q.z / n, q.w / n};
}
template <class T> V3<T> rotate(const Q4<T>& q, const V3<T>& p) {  // so3.hpp:362-371
  V3<T> uv{q.y * p.z - q.z * p.y, q.z * p.x - q.x * p.z, q.x * p.y - q.y * p.x};
  uv.x = uv.x + uv.x; uv.y = uv.y + uv.y; uv.z = uv.z + uv.z;
  return V3<T>{p.x + q.w * uv.x + (q.y * uv.z - q.z * uv.y),
               p.y + q.w * uv.y + (q.z * uv.x - q.x * uv.z),
               p.z + q.w * uv.z + (q.x * uv.y - q.y * uv.x)};
}
template <class T> Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {  // so3.hpp:329-345 (+ctor normalise :482-489)
  Q4<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return normalized(r);
}
template <class T> SE3<T> inverse(const SE3<T>& a) {  // se3.hpp:206-211
  SE3<T> r;
  r.q = normalized(Q4<T>{-a.q.x, -a.q.y, -a.q.z, a.q.w});
  const V3<T> nt{-a.t.x, -a.t.y, -a.t.z};
  r.t = rotate(r.q, nt);
  return r;
}
template <class T> SE3<T> mul(const SE3<T>& a, const SE3<T>& b) {  // se3.hpp:308-314
  SE3<T> r;
  r.q = qmul(a.q, b.q);
  const V3<T> rb = rotate(a.q, b.t);
  r.t = V3<T>{a.t.x + rb.x, a.t.y + rb.y, a.t.z + rb.z};
  return r;
}
template <class T> V3<T> act(const SE3<T>& a, const V3<T>& p) {  // se3.hpp:325-328
  const V3<T> r = rotate(a.q, p);
  return V3<T>{r.x + a.t.x, r.y + a.t.y, r.z + a.t.z};
}
template <class T> SE3<T> map_se3(const T* s) {  // Eigen::Map: no normalisation
  return SE3<T>{Q4<T>{s[0], s[1], s[2], s[3]}, V3<T>{s[4], s[5], s[6]}};
}

// SE3::exp on doubles (se3.hpp:763-784, so3.hpp:585-621).
void se3_exp(const double* d, double* q, double* t) {
  const double ox = d[3], oy = d[4], oz = d[5];
  const double th2 = ox * ox + oy * oy + oz * oz;
  const double eps = 1e-10;
  double th, imag, real;
  if (th2 < eps * eps) {
    th = 0;
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    th = sqrt(th2);
    imag = sin(0.5 * th) / th;
    real = cos(0.5 * th);
  }
  q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = real;
  const double O[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  double O2[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += O[3 * i + k] * O[3 * k + j];
      O2[3 * i + j] = s;
    }
  if (th < eps) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];  // so3.matrix()
    V[0] = 1 - 2 * (y * y + z * z); V[1] = 2 * (x * y - w * z); V[2] = 2 * (x * z + w * y);
    V[3] = 2 * (x * y + w * z); V[4] = 1 - 2 * (x * x + z * z); V[5] = 2 * (y * z - w * x);
    V[6] = 2 * (x * z - w * y); V[7] = 2 * (y * z + w * x); V[8] = 1 - 2 * (x * x + y * y);
  } else {
    const double a = (1.0 - cos(th)) / th2, b = (th - sin(th)) / (th2 * th);
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
  }
  for (int i = 0; i < 3; ++i) t[i] = V[3 * i] * d[0] + V[3 * i + 1] * d[1] + V[3 * i + 2] * d[2];
}

// LocalParameterizationSE3::Plus (local_parameterization_se3.hpp:44-51).
void se3_plus(const double* T, const double* d, double* out) {
  double qe[4], te[3];
  se3_exp(d, qe, te);
  const SE3<double> a = map_se3(T);
  const SE3<double> e{Q4<double>{qe[0], qe[1], qe[2], qe[3]}, V3<double>{te[0], te[1], te[2]}};
  const SE3<double> r = mul(a, e);
  out[0] = r.q.x; out[1] = r.q.y; out[2] = r.q.z; out[3] = r.q.w;
  out[4] = r.t.x; out[5] = r.t.y; out[6] = r.t.z;
}

// SE3::Dx_this_mul_exp_x_at_0 (se3.hpp:135-204): 7x6 row-major.
void se3_plus_jacobian(const double* T, double* J) {
  const double qx = T[0], qy = T[1], qz = T[2], qw = T[3];
  memset(J, 0, 42 * sizeof(double));
  const double c0 = 0.5 * qw, c1 = 0.5 * qz, c3 = 0.5 * qy, c4 = 0.5 * qx;
  J[0 * 6 + 3] = c0;  J[0 * 6 + 4] = -c1; J[0 * 6 + 5] = c3;
  J[1 * 6 + 3] = c1;  J[1 * 6 + 4] = c0;  J[1 * 6 + 5] = -c4;
  J[2 * 6 + 3] = -c3; J[2 * 6 + 4] = c4;  J[2 * 6 + 5] = c0;
  J[3 * 6 + 3] = -c4; J[3 * 6 + 4] = -c3; J[3 * 6 + 5] = -c1;
  const double ww = qw * qw, xx = qx * qx, yy = qy * qy, zz = qz * qz;
  J[4 * 6 + 0] = ww + xx - yy - zz;       J[4 * 6 + 1] = 2 * (qx * qy - qw * qz); J[4 * 6 + 2] = 2 * (qw * qy + qx * qz);
  J[5 * 6 + 0] = 2 * (qw * qz + qx * qy); J[5 * 6 + 1] = ww - xx + yy - zz;       J[5 * 6 + 2] = 2 * (qy * qz - qw * qx);
  J[6 * 6 + 0] = 2 * (qx * qz - qw * qy); J[6 * 6 + 1] = 2 * (qw * qx + qy * qz); J[6 * 6 + 2] = ww - xx - yy + zz;
}

// ------------------------------------------------- camera models (visnav) --
template <class T> void project(int model, const T* p, const V3<T>& X, T* uv) {
  const T &fx = p[0], &fy = p[1], &cx = p[2], &cy = p[3];
  const T &x = X.x, &y = X.y, &z = X.z;
  if (model == PBA_CAM_PINHOLE) {  // camera_models.h:75-91
    uv[0] = fx * x / z + cx;
    uv[1] = fy * y / z + cy;
  } else if (model == PBA_CAM_DS) {  // :226-249
    const T &xi = p[4], &alpha = p[5];
    const T d1 = jsqrt(x * x + y * y + z * z);
    const T k = xi * d1 + z;
    const T d2 = jsqrt(x * x + y * y + k * k);
    const T denom = alpha * d2 + (1.0 - alpha) * (xi * d1 + z);
    uv[0] = fx * x / denom + cx;
    uv[1] = fy * y / denom + cy;
  } else if (model == PBA_CAM_KB4) {  // :316-351
    const T &k1 = p[4], &k2 = p[5], &k3 = p[6], &k4 = p[7];
    const T r = jsqrt(x * x + y * y);
    if (r == 0.0) { uv[0] = cx; uv[1] = cy; return; }
    const T th = jatan2(r, z);
    const T th2 = th * th, th3 = th2 * th;
    const T d = th + th3 * (k1 + th2 * (k2 + th2 * (k3 + th2 * k4)));
    uv[0] = fx * d * x / r + cx;
    uv[1] = fy * d * y / r + cy;
  } else {  // EUCM :144-164
    const T &alpha = p[4], &beta = p[5];
    const T d = jsqrt(beta * (x * x + y * y) + z * z);
    uv[0] = fx * x / (alpha * d + (1.0 - alpha) * z) + cx;
    uv[1] = fy * y / (alpha * d + (1.0 - alpha) * z) + cy;
  }
}

template <class T> V3<T> unproject(int model, const T* p, const T& u, const T& v) {
  const T &fx = p[0], &fy = p[1], &cx = p[2], &cy = p[3];
  const T mx = (u - cx) / fx, my = (v - cy) / fy;
  if (model == PBA_CAM_PINHOLE) {  // :93-107
    const T n = jsqrt(mx * mx + my * my + 1.0);
    return V3<T>{mx / n, my / n, T(1.0) / n};
  } else if (model == PBA_CAM_DS) {  // :251-277
    const T &xi = p[4], &alpha = p[5];
    const T r2 = mx * mx + my * my;
    const T mz = (1.0 - alpha * alpha * r2) / (alpha * jsqrt(1.0 - (2.0 * alpha - 1.0) * r2) + 1.0 - alpha);
    const T f = (mz * xi + jsqrt(mz * mz + (1.0 - xi * xi) * r2)) / (mz * mz + r2);
    return V3<T>{f * mx, f * my, f * mz - xi};
  } else if (model == PBA_CAM_KB4) {  // :353-379
    const T &k1 = p[4], &k2 = p[5], &k3 = p[6], &k4 = p[7];
    const T ru = jsqrt(mx * mx + my * my);
    if (ru == 0.0) return V3<T>{T(0.0), T(0.0), T(1.0)};
    T th = T(0.0);
    for (int i = 0; i < 5; ++i) {
      const T t2 = th * th, t3 = t2 * th;
      const T f = th + t3 * (k1 + t2 * (k2 + t2 * (k3 + t2 * k4))) - ru;
      const T df = 1.0 + t2 * (3.0 * k1 + t2 * (5.0 * k2 + t2 * (7.0 * k3 + t2 * 9.0 * k4)));
      th = th - f / df;
    }
    return V3<T>{jsin(th) * mx / ru, jsin(th) * my / ru, jcos(th)};
  } else {  // EUCM :166-188
    const T &alpha = p[4], &beta = p[5];
    const T r2 = mx * mx + my * my;
    const T mz = (1.0 - beta * alpha * alpha * r2) / (alpha * jsqrt(1.0 - (2.0 * alpha - 1.0) * beta * r2) + (1.0 - alpha));
    const T n = jsqrt(mx * mx + my * my + mz * mz);
    return V3<T>{mx / n, my / n, mz / n};
  }
}

template <class T> V3<T> unit(const V3<T>& a) {
  const T n = jsqrt(a.x * a.x + a.y * a.y + a.z * a.z);
  return V3<T>{a.x / n, a.y / n, a.z / n};
}

const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

struct Image { const uint8_t* ptr; int w, h, pitch; };

template <class T> bool bilinear(const Image& im, const T& u, const T& v, T* out) {
  const double us = scalar_of(u), vs = scalar_of(v);
  if (!(us >= 0.0) || !(vs >= 0.0) || !(us < double(im.w - 1)) || !(vs < double(im.h - 1))) return false;
  const int x0 = int(floor(us)), y0 = int(floor(vs));
  const T fx = u - double(x0), fy = v - double(y0);
  const uint8_t* p = im.ptr + size_t(y0) * im.pitch + x0;
  const double i00 = p[0], i10 = p[1], i01 = p[im.pitch], i11 = p[im.pitch + 1];
  *out = (1.0 - fx) * (1.0 - fy) * i00 + fx * (1.0 - fy) * i10 + (1.0 - fx) * fy * i01 + fx * fy * i11;
  return true;
}

// ------------------------------------------------------------- functors ----
// reprojection.h:82-112.  Parameter blocks (7 host, 7 target, 1 rho, 8 target
// intrinsics) -> Jet<23>, exactly Ceres' AutoDiffCostFunction<...,2,7,7,1,8>.
// NB both cameras use the HOST model name (reprojection.h:97-100).
template <class T>
void geometric_functor(const double* zt, const double* zh, const double* host_intr, int model,
                       const T* Th, const T* Tt, const T* rho, const T* tgt_intr, T* res) {
  T hi[8];
  for (int i = 0; i < 8; ++i) hi[i] = T(host_intr[i]);
  const V3<T> b = unit(unproject<T>(model, hi, T(zh[0]), T(zh[1])));
  const SE3<T> T_w_c1 = map_se3(Th), T_w_c2 = map_se3(Tt);
  const V3<T> Xh{b.x / rho[0], b.y / rho[0], b.z / rho[0]};
  const V3<T> Xt = act(mul(inverse(T_w_c2), T_w_c1), Xh);
  T uv[2];
  project<T>(model, tgt_intr, Xt, uv);
  res[0] = zt[0] - uv[0];
  res[1] = zt[1] - uv[1];
}

// SURVEY.md §8(a-P).  Parameter blocks (7 host, 7 target, 2 affine, 1 rho).
template <class T>
void photometric_functor(const double* zh, const double* I_h, bool host_valid, const Image& target,
                         const double* host_intr, int host_model, const double* tgt_intr, int tgt_model,
                         const T* Th, const T* Tt, const T* aff, const T* rho, T* res) {
  T hi[8], ti[8];
  for (int i = 0; i < 8; ++i) { hi[i] = T(host_intr[i]); ti[i] = T(tgt_intr[i]); }
  const SE3<T> T_w_c1 = map_se3(Th), T_w_c2 = map_se3(Tt);
  bool ok = host_valid;
  const T ea = jexp(aff[0]);
  for (int k = 0; k < 8 && ok; ++k) {
    const V3<T> b = unit(unproject<T>(host_model, hi, T(zh[0] + kPattern[k][0]), T(zh[1] + kPattern[k][1])));
    const V3<T> Xh{b.x / rho[0], b.y / rho[0], b.z / rho[0]};
    const V3<T> Xt = act(mul(inverse(T_w_c2), T_w_c1), Xh);
    T uv[2], It;
    project<T>(tgt_model, ti, Xt, uv);
    ok = bilinear(target, uv[0], uv[1], &It);
    if (ok) res[k] = It - (ea * I_h[k] + aff[1]);
  }
  if (!ok) for (int k = 0; k < 8; ++k) res[k] = T(0.0);
}

// -------------------------------------------------------------- problem ----
struct Prob {
  const pba_problem* p;
  bool photo;
  int R, C;           // residuals per block, local columns per block
  int cam_dim;        // 6 geometric, 8 photometric
  bool use_huber;
  double huber;
  std::vector<int> obs_lm;       // landmark of each obs
  std::vector<double> I_h;       // [n_lm*8]
  std::vector<uint8_t> host_ok;  // [n_lm]
};

Image image_of(const pba_problem* p, int pose) {
  return Image{p->image_ptrs ? p->image_ptrs[pose] : p->images + size_t(pose) * p->image_stride,
               p->width, p->height, p->pitch};
}

void init_prob(const pba_problem* p, bool use_huber, double huber, Prob* P) {
  P->p = p;
  P->photo = p->mode == PBA_MODE_PHOTOMETRIC;
  P->R = P->photo ? 8 : 2;
  P->C = P->photo ? 15 : 13;
  P->cam_dim = P->photo ? 8 : 6;
  P->use_huber = use_huber;
  P->huber = huber;
  P->obs_lm.resize(p->n_obs);
  for (int l = 0; l < p->n_landmarks; ++l)
    for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) P->obs_lm[o] = l;
  if (P->photo) {
    P->I_h.assign(size_t(p->n_landmarks) * 8, 0.0);
    P->host_ok.assign(p->n_landmarks, 1);
    for (int l = 0; l < p->n_landmarks; ++l) {
      const Image him = image_of(p, p->lm_host[l]);
      bool ok = true;
      for (int k = 0; k < 8 && ok; ++k)
        ok = bilinear<double>(him, p->lm_host_uv[2 * l] + kPattern[k][0], p->lm_host_uv[2 * l + 1] + kPattern[k][1],
                              &P->I_h[size_t(l) * 8 + k]);
      P->host_ok[l] = ok;
    }
  }
}

// One residual block: AutoDiff -> global Jacobians -> x local-param Jacobian ->
// robust correction (residual_block.cc:69-198).  State arrays may differ from
// the problem's (candidate point).  J may be NULL (cost-only).
// Returns the block's cost rho(s)/2.
double eval_block(const Prob& P, int64_t o, const double* poses, const double* affine, const double* rho_all,
                  double* res, double* J) {
  const pba_problem* p = P.p;
  const int l = P.obs_lm[o];
  const int h = p->lm_host[l], t = p->obs_target[o];
  const int hm = p->calib_model[p->pose_calib[h]];
  const double* hintr = p->intrinsics + 8 * p->pose_calib[h];
  const double* tintr = p->intrinsics + 8 * p->pose_calib[t];
  const int R = P.R, C = P.C;
  if (!J) {
    if (!P.photo) {
      geometric_functor<double>(p->obs_uv + 2 * o, p->lm_host_uv + 2 * l, hintr, hm, poses + 7 * h, poses + 7 * t,
                                rho_all + l, tintr, res);
    } else {
      photometric_functor<double>(p->lm_host_uv + 2 * l, &P.I_h[size_t(l) * 8], P.host_ok[l], image_of(p, t), hintr, hm,
                                  tintr, p->calib_model[p->pose_calib[t]], poses + 7 * h, poses + 7 * t, affine + 2 * t,
                                  rho_all + l, res);
    }
  } else {
    double Jg_h[8 * 7], Jg_t[8 * 7], J_aff[8 * 2], J_rho[8];
    if (!P.photo) {
      typedef Jet<23> JT;
      JT Th[7], Tt[7], rh[1], ti[8], r[2];
      for (int i = 0; i < 7; ++i) { Th[i] = JT(poses[7 * h + i], i); Tt[i] = JT(poses[7 * t + i], 7 + i); }
      rh[0] = JT(rho_all[l], 14);
      for (int i = 0; i < 8; ++i) ti[i] = JT(tintr[i], 15 + i);
      geometric_functor<JT>(p->obs_uv + 2 * o, p->lm_host_uv + 2 * l, hintr, hm, Th, Tt, rh, ti, r);
      for (int k = 0; k < 2; ++k) {
        res[k] = r[k].a;
        for (int i = 0; i < 7; ++i) { Jg_h[k * 7 + i] = r[k].v[i]; Jg_t[k * 7 + i] = r[k].v[7 + i]; }
        J_rho[k] = r[k].v[14];
      }
    } else {
      typedef Jet<17> JT;
      JT Th[7], Tt[7], af[2], rh[1], r[8];
      for (int i = 0; i < 7; ++i) { Th[i] = JT(poses[7 * h + i], i); Tt[i] = JT(poses[7 * t + i], 7 + i); }
      af[0] = JT(affine[2 * t], 14); af[1] = JT(affine[2 * t + 1], 15);
      rh[0] = JT(rho_all[l], 16);
      photometric_functor<JT>(p->lm_host_uv + 2 * l, &P.I_h[size_t(l) * 8], P.host_ok[l], image_of(p, t), hintr, hm, tintr,
                              p->calib_model[p->pose_calib[t]], Th, Tt, af, rh, r);
      for (int k = 0; k < 8; ++k) {
        res[k] = r[k].a;
        for (int i = 0; i < 7; ++i) { Jg_h[k * 7 + i] = r[k].v[i]; Jg_t[k * 7 + i] = r[k].v[7 + i]; }
        J_aff[k * 2] = r[k].v[14]; J_aff[k * 2 + 1] = r[k].v[15];
        J_rho[k] = r[k].v[16];
      }
    }
    // jacobians[i] = global_jacobians[i] * global_to_local (residual_block.cc:144-155)
    double Ph[42], Pt[42];
    se3_plus_jacobian(poses + 7 * h, Ph);
    se3_plus_jacobian(poses + 7 * t, Pt);
    for (int k = 0; k < R; ++k) {
      double* row = J + k * C;
      for (int c = 0; c < 6; ++c) {
        double sh = 0, st = 0;
        for (int i = 0; i < 7; ++i) { sh += Jg_h[k * 7 + i] * Ph[i * 6 + c]; st += Jg_t[k * 7 + i] * Pt[i * 6 + c]; }
        row[c] = sh; row[6 + c] = st;
      }
      if (P.photo) { row[12] = J_aff[k * 2]; row[13] = J_aff[k * 2 + 1]; }
      row[C - 1] = J_rho[k];
    }
  }
  double s = 0;
  for (int k = 0; k < R; ++k) s += res[k] * res[k];
  if (!P.use_huber) return 0.5 * s;
  // HuberLoss::Evaluate (loss_function.cc:48-62) with a = huber, b = a^2
  double rho0, rho1;
  const double b = P.huber * P.huber;
  if (s > b) {
    const double r = sqrt(s);
    rho0 = 2.0 * P.huber * r - b;
    rho1 = std::max(std::numeric_limits<double>::min(), P.huber / r);
  } else {
    rho0 = s; rho1 = 1.0;
  }
  // Corrector (corrector.cc:82-86,127-130): rho'' <= 0 for Huber => pure scaling by sqrt(rho').
  const double w = sqrt(rho1);
  if (J) for (int i = 0; i < R * C; ++i) J[i] *= w;
  for (int k = 0; k < R; ++k) res[k] *= w;
  return 0.5 * rho0;
}

// Whole-problem evaluation (program_evaluator.h:139-286).
double evaluate(const Prob& P, const double* poses, const double* affine, const double* rho, double* residuals,
                double* jacobians, int threads) {
  const int64_t n = P.p->n_obs;
  double cost = 0;
  (void)threads;
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(threads)
  for (int64_t o = 0; o < n; ++o) {
    double r[8];
    cost += eval_block(P, o, poses, affine, rho, residuals ? residuals + o * P.R : r,
                       jacobians ? jacobians + o * P.R * P.C : nullptr);
  }
  return cost;
}

// ------------------------------------------------------ dense Cholesky -----
bool cholesky_solve(int n, std::vector<double>& A, std::vector<double>& b) {
  // in-place lower Cholesky (row-major), then forward/back substitution
  for (int j = 0; j < n; ++j) {
    double d = A[size_t(j) * n + j];
    for (int k = 0; k < j; ++k) d -= A[size_t(j) * n + k] * A[size_t(j) * n + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    A[size_t(j) * n + j] = d;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < n; ++i) {
      double s = A[size_t(i) * n + j];
      for (int k = 0; k < j; ++k) s -= A[size_t(i) * n + k] * A[size_t(j) * n + k];
      A[size_t(i) * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[size_t(i) * n + k] * b[k];
    b[i] = s / A[size_t(i) * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[size_t(k) * n + i] * b[k];
    b[i] = s / A[size_t(i) * n + i];
  }
  return true;
}

// ---------------------------------------------------------------- solver ---
struct Layout {
  std::vector<int> slot;   // pose -> RCS camera slot or -1 (constant / unused)
  std::vector<uint8_t> affine_active;
  int n_slots = 0;
  int dim = 0;
};

void make_layout(const Prob& P, Layout* L) {
  const pba_problem* p = P.p;
  std::vector<uint8_t> used(p->n_poses, 0), is_target(p->n_poses, 0);
  for (int l = 0; l < p->n_landmarks; ++l)
    for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) {
      used[p->lm_host[l]] = 1; used[p->obs_target[o]] = 1; is_target[p->obs_target[o]] = 1;
    }
  L->slot.assign(p->n_poses, -1);
  L->affine_active.assign(p->n_poses, 0);
  L->n_slots = 0;
  for (int i = 0; i < p->n_poses; ++i) {
    const bool fixed = p->pose_fixed && p->pose_fixed[i];
    if (!fixed && used[i]) L->slot[i] = L->n_slots++;
    L->affine_active[i] = P.photo && !fixed && is_target[i];
  }
  L->dim = L->n_slots * P.cam_dim;
}

// Scaled-Jacobian Schur system for given D (schur_eliminator_impl.h:177-306).
// J: robustified, column-scaled local Jacobians [n_obs*R*C]; r: robustified residuals.
// Dcam [dim], Drho [n_lm].  Outputs dense S [dim*dim], rhs [dim]; ete_inv, g per landmark.
void schur_eliminate(const Prob& P, const Layout& L, const double* J, const double* r, const double* Dcam,
                     const double* Drho, std::vector<double>& S, std::vector<double>& rhs) {
  const pba_problem* p = P.p;
  const int n = L.dim, cd = P.cam_dim, R = P.R, C = P.C;
  S.assign(size_t(n) * n, 0.0);
  rhs.assign(n, 0.0);
  for (int i = 0; i < n; ++i) S[size_t(i) * n + i] = Dcam[i] * Dcam[i];
  std::vector<double> buf(n);
  std::vector<int> touched;
  for (int l = 0; l < p->n_landmarks; ++l) {
    const int64_t o0 = p->lm_obs_ptr[l], o1 = p->lm_obs_ptr[l + 1];
    if (o0 == o1) continue;
    const int h = p->lm_host[l], hs = L.slot[h];
    double ete = Drho[l] * Drho[l], g = 0;
    touched.clear();
    auto add_block = [&](int sa, const double* Ja, int wa, int sb, const double* Jb, int wb) {
      // S[sa,sb] += Ja^T Jb over the R rows (upper triangle: sa <= sb)
      for (int i = 0; i < wa; ++i)
        for (int j = 0; j < wb; ++j) {
          double s = 0;
          for (int k = 0; k < R; ++k) s += Ja[k * C + i] * Jb[k * C + j];
          S[size_t(sa * cd + i) * n + sb * cd + j] += s;
        }
    };
    for (int64_t o = o0; o < o1; ++o) {
      const double* Jo = J + o * R * C;
      const double* ro = r + o * R;
      const int t = p->obs_target[o], ts = L.slot[t];
      const int wt = cd;  // target block: 6 pose (+2 affine)
      for (int k = 0; k < R; ++k) { ete += Jo[k * C + C - 1] * Jo[k * C + C - 1]; g += Jo[k * C + C - 1] * ro[k]; }
      if (hs >= 0) {
        add_block(hs, Jo, 6, hs, Jo, 6);
        for (int i = 0; i < 6; ++i) {
          double e = 0, f = 0;
          for (int k = 0; k < R; ++k) { e += Jo[k * C + C - 1] * Jo[k * C + i]; f += Jo[k * C + i] * ro[k]; }
          if (buf[hs * cd + i] == 0.0 && e != 0.0) {}
          buf[hs * cd + i] += e;
          rhs[hs * cd + i] += f;
        }
        touched.push_back(hs);
      }
      if (ts >= 0) {
        add_block(ts, Jo + 6, wt, ts, Jo + 6, wt);
        for (int i = 0; i < wt; ++i) {
          double e = 0, f = 0;
          for (int k = 0; k < R; ++k) { e += Jo[k * C + C - 1] * Jo[k * C + 6 + i]; f += Jo[k * C + 6 + i] * ro[k]; }
          buf[ts * cd + i] += e;
          rhs[ts * cd + i] += f;
        }
        touched.push_back(ts);
      }
      if (hs >= 0 && ts >= 0) {
        if (hs < ts) add_block(hs, Jo, 6, ts, Jo + 6, wt);
        else {
          // S[ts,hs] += Jt^T Jh
          for (int i = 0; i < wt; ++i)
            for (int j = 0; j < 6; ++j) {
              double s = 0;
              for (int k = 0; k < R; ++k) s += Jo[k * C + 6 + i] * Jo[k * C + j];
              S[size_t(ts * cd + i) * n + hs * cd + j] += s;
            }
        }
      }
    }
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    const double inv = 1.0 / ete;
    for (size_t a = 0; a < touched.size(); ++a) {
      const int sa = touched[a];
      for (int i = 0; i < cd; ++i) rhs[sa * cd + i] -= buf[sa * cd + i] * inv * g;
      for (size_t b = a; b < touched.size(); ++b) {
        const int sb = touched[b];
        for (int i = 0; i < cd; ++i)
          for (int j = 0; j < cd; ++j) S[size_t(sa * cd + i) * n + sb * cd + j] -= buf[sa * cd + i] * inv * buf[sb * cd + j];
      }
    }
    for (int s : touched) for (int i = 0; i < cd; ++i) buf[s * cd + i] = 0.0;
  }
  // symmetrise (only block-upper part was accumulated)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const int bi = i / cd, bj = j / cd;
      if (bi != bj) S[size_t(i) * n + j] = S[size_t(j) * n + i];
    }
}

struct State {
  std::vector<double> poses, affine, rho;
};

// Plus for the whole parameter vector; delta layout: [cams dim | n_lm].
void plus_all(const Prob& P, const Layout& L, const State& x, const double* dcam, const double* drho, State* out) {
  const pba_problem* p = P.p;
  *out = x;
  for (int i = 0; i < p->n_poses; ++i) {
    const int s = L.slot[i];
    if (s < 0) continue;
    se3_plus(&x.poses[7 * i], dcam + s * P.cam_dim, &out->poses[7 * i]);
    if (P.photo && L.affine_active[i]) {
      out->affine[2 * i] = x.affine[2 * i] + dcam[s * P.cam_dim + 6];
      out->affine[2 * i + 1] = x.affine[2 * i + 1] + dcam[s * P.cam_dim + 7];
    }
  }
  for (int l = 0; l < p->n_landmarks; ++l)
    if (p->lm_obs_ptr[l + 1] > p->lm_obs_ptr[l]) out->rho[l] = x.rho[l] + drho[l];
}

// Ambient-space norms over the reduced program's parameter blocks.
double ambient_sqnorm(const Prob& P, const Layout& L, const State& a, const State* b, double* maxabs) {
  const pba_problem* p = P.p;
  double s = 0, m = 0;
  auto acc = [&](double v) { s += v * v; m = std::max(m, fabs(v)); };
  for (int i = 0; i < p->n_poses; ++i) {
    if (L.slot[i] < 0) continue;
    for (int k = 0; k < 7; ++k) acc(a.poses[7 * i + k] - (b ? b->poses[7 * i + k] : 0.0));
    if (P.photo && L.affine_active[i])
      for (int k = 0; k < 2; ++k) acc(a.affine[2 * i + k] - (b ? b->affine[2 * i + k] : 0.0));
  }
  for (int l = 0; l < p->n_landmarks; ++l)
    if (p->lm_obs_ptr[l + 1] > p->lm_obs_ptr[l]) acc(a.rho[l] - (b ? b->rho[l] : 0.0));
  if (maxabs) *maxabs = m;
  return s;
}

}  // namespace

#define ORACLE_API extern "C" __attribute__((visibility("default")))

ORACLE_API int pba_oracle_project(int model, const double* intr, int64_t n, const double* xyz, double* uv, double* J) {
  for (int64_t i = 0; i < n; ++i) {
    if (J) {
      typedef Jet<3> JT;
      JT pi[8], r[2];
      for (int k = 0; k < 8; ++k) pi[k] = JT(intr[k]);
      const V3<JT> X{JT(xyz[3 * i], 0), JT(xyz[3 * i + 1], 1), JT(xyz[3 * i + 2], 2)};
      project<JT>(model, pi, X, r);
      for (int a = 0; a < 2; ++a) { uv[2 * i + a] = r[a].a; for (int k = 0; k < 3; ++k) J[6 * i + 3 * a + k] = r[a].v[k]; }
    } else {
      const V3<double> X{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
      project<double>(model, intr, X, uv + 2 * i);
    }
  }
  return 0;
}

ORACLE_API int pba_oracle_unproject(int model, const double* intr, int64_t n, const double* uv, double* xyz) {
  for (int64_t i = 0; i < n; ++i) {
    const V3<double> b = unproject<double>(model, intr, uv[2 * i], uv[2 * i + 1]);
    xyz[3 * i] = b.x; xyz[3 * i + 1] = b.y; xyz[3 * i + 2] = b.z;
  }
  return 0;
}

ORACLE_API int pba_oracle_se3_plus(int64_t n, const double* poses7, const double* delta6, double* out7) {
  for (int64_t i = 0; i < n; ++i) se3_plus(poses7 + 7 * i, delta6 + 6 * i, out7 + 7 * i);
  return 0;
}

ORACLE_API int pba_oracle_se3_plus_jacobian(const double* pose7, double* J42) {
  se3_plus_jacobian(pose7, J42);
  return 0;
}

// Per-block robustified residuals [n_obs*R] / local Jacobians [n_obs*R*C] in
// the caller's observation order; *cost = sum rho(s)/2.
ORACLE_API int pba_oracle_eval(const pba_problem* p, int use_huber, double huber, int num_threads, double* residuals,
                               double* jacobians, double* cost) {
  Prob P;
  init_prob(p, use_huber != 0, huber, &P);
  const double c = evaluate(P, p->poses, p->affine, p->inv_depth, residuals, jacobians, num_threads > 0 ? num_threads : 1);
  if (cost) *cost = c;
  return 0;
}

// Dense reduced camera system for the CURRENT state, with Jacobi scaling
// computed from this Jacobian (i.e. what Ceres builds at iteration 0) and LM
// damping for `radius`.  S [dim*dim], rhs [dim]; dim = cam_dim * #free cameras.
ORACLE_API int pba_oracle_build_rcs(const pba_problem* p, int use_huber, double huber, double radius, int num_threads,
                                    int32_t* dim_out, double* S_out, double* rhs_out, double* scale_cam_out) {
  Prob P;
  init_prob(p, use_huber != 0, huber, &P);
  Layout L;
  make_layout(P, &L);
  if (dim_out) *dim_out = L.dim;
  if (!S_out) return 0;
  const int R = P.R, C = P.C, cd = P.cam_dim;
  std::vector<double> r(size_t(p->n_obs) * R), J(size_t(p->n_obs) * R * C);
  evaluate(P, p->poses, p->affine, p->inv_depth, r.data(), J.data(), num_threads > 0 ? num_threads : 1);
  std::vector<double> ncam(L.dim, 0.0), nrho(p->n_landmarks, 0.0);
  for (int64_t o = 0; o < p->n_obs; ++o) {
    const int l = P.obs_lm[o], hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
    for (int k = 0; k < R; ++k) {
      const double* row = &J[(o * R + k) * C];
      if (hs >= 0) for (int i = 0; i < 6; ++i) ncam[hs * cd + i] += row[i] * row[i];
      if (ts >= 0) for (int i = 0; i < cd; ++i) ncam[ts * cd + i] += row[6 + i] * row[6 + i];
      nrho[l] += row[C - 1] * row[C - 1];
    }
  }
  std::vector<double> scam(L.dim), srho(p->n_landmarks), Dcam(L.dim), Drho(p->n_landmarks);
  for (int i = 0; i < L.dim; ++i) scam[i] = 1.0 / (1.0 + sqrt(ncam[i]));
  for (int l = 0; l < p->n_landmarks; ++l) srho[l] = 1.0 / (1.0 + sqrt(nrho[l]));
  for (int64_t o = 0; o < p->n_obs; ++o) {
    const int l = P.obs_lm[o], hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
    for (int k = 0; k < R; ++k) {
      double* row = &J[(o * R + k) * C];
      for (int i = 0; i < 6; ++i) row[i] = hs >= 0 ? row[i] * scam[hs * cd + i] : 0.0;
      for (int i = 0; i < cd; ++i) row[6 + i] = ts >= 0 ? row[6 + i] * scam[ts * cd + i] : 0.0;
      row[C - 1] *= srho[l];
    }
  }
  for (int i = 0; i < L.dim; ++i) Dcam[i] = sqrt(std::min(std::max(ncam[i] * scam[i] * scam[i], 1e-6), 1e32) / radius);
  for (int l = 0; l < p->n_landmarks; ++l) Drho[l] = sqrt(std::min(std::max(nrho[l] * srho[l] * srho[l], 1e-6), 1e32) / radius);
  std::vector<double> S, rhs;
  schur_eliminate(P, L, J.data(), r.data(), Dcam.data(), Drho.data(), S, rhs);
  memcpy(S_out, S.data(), S.size() * sizeof(double));
  memcpy(rhs_out, rhs.data(), rhs.size() * sizeof(double));
  if (scale_cam_out) memcpy(scale_cam_out, scam.data(), scam.size() * sizeof(double));
  return 0;
}

// Full LM solve restating TrustRegionMinimizer + LevenbergMarquardtStrategy +
// SchurComplementSolver (dense Cholesky on the RCS instead of sparse LDL^T:
// both are exact solves).  Writes the best state back into the problem arrays.
ORACLE_API int pba_oracle_solve(pba_problem* p, const pba_options* opt, int num_threads, pba_summary* sum) {
  const auto t_start = std::chrono::steady_clock::now();
  auto now = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
  if (opt->optimize_intrinsics) return PBA_ERR_UNSUPPORTED;
  const int threads = num_threads > 0 ? num_threads : 1;
  Prob P;
  init_prob(p, opt->use_huber != 0, opt->huber_parameter, &P);
  Layout L;
  make_layout(P, &L);
  const int R = P.R, C = P.C, cd = P.cam_dim, n = L.dim, nl = p->n_landmarks;
  const int64_t no = p->n_obs;

  State x, cand;
  x.poses.assign(p->poses, p->poses + 7 * p->n_poses);
  x.affine.assign(2 * p->n_poses, 0.0);
  if (P.photo && p->affine) x.affine.assign(p->affine, p->affine + 2 * p->n_poses);
  x.rho.assign(p->inv_depth, p->inv_depth + nl);
  State best = x;

  std::vector<double> r(size_t(no) * R), J(size_t(no) * R * C), rc(size_t(no) * R);
  std::vector<double> scam(n, 1.0), srho(nl, 1.0), diag_cam(n), diag_rho(nl), Dcam(n), Drho(nl);
  std::vector<double> gcam(n), grho(nl), S, rhs, ycam(n), yrho(nl), dcam(n), drho(nl);

  pba_iteration* its = sum ? sum->iterations : nullptr;
  const int cap = sum ? sum->iterations_capacity : 0;
  if (sum) { memset(sum, 0, sizeof(*sum)); sum->iterations = its; sum->iterations_capacity = cap; }
  int n_it = 0;
  double t_push0 = -1.0, t_push_prev = 0.0;
  auto push = [&](pba_iteration& it) {
    const double t = now();
    if (t_push0 < 0.0) { t_push0 = t_push_prev = t; }
    it.iteration_time_in_seconds = t - t_push_prev;
    it.cumulative_time_in_seconds = t - t_push0;
    t_push_prev = t;
    if (its && n_it < cap) its[n_it] = it;
    ++n_it;
  };

  double t_jac = 0, t_res = 0, t_lin = 0;
  int n_jac = 0, n_res = 0, n_lin = 0;
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double x_cost = 0, x_norm = -1.0, minimum_cost = std::numeric_limits<double>::max();
  double grad_max = 0, grad_norm = 0;
  int termination = PBA_NO_CONVERGENCE;
  char message[256] = "";
  int num_successful = 0, num_unsuccessful = 0, consecutive_invalid = 0;

  // EvaluateGradientAndJacobian (trust_region_minimizer.cc:228-300)
  auto eval_grad_jac = [&](bool first) -> bool {
    const double t0 = now();
    x_cost = evaluate(P, x.poses.data(), x.affine.data(), x.rho.data(), r.data(), J.data(), threads);
    ++n_jac; ++n_res;
    if (!std::isfinite(x_cost)) return false;
    // gradient g = J^T r on the UNSCALED local Jacobian (program_evaluator.h:240-256)
    std::fill(gcam.begin(), gcam.end(), 0.0);
    std::fill(grho.begin(), grho.end(), 0.0);
    std::vector<double> ncam(n, 0.0), nrho(nl, 0.0);
    for (int64_t o = 0; o < no; ++o) {
      const int l = P.obs_lm[o], hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
      for (int k = 0; k < R; ++k) {
        const double* row = &J[(o * R + k) * C];
        const double rk = r[o * R + k];
        if (hs >= 0) for (int i = 0; i < 6; ++i) { gcam[hs * cd + i] += row[i] * rk; ncam[hs * cd + i] += row[i] * row[i]; }
        if (ts >= 0) for (int i = 0; i < cd; ++i) { gcam[ts * cd + i] += row[6 + i] * rk; ncam[ts * cd + i] += row[6 + i] * row[6 + i]; }
        grho[l] += row[C - 1] * rk;
        nrho[l] += row[C - 1] * row[C - 1];
      }
    }
    if (first && opt->jacobi_scaling) {
      for (int i = 0; i < n; ++i) scam[i] = 1.0 / (1.0 + sqrt(ncam[i]));
      for (int l = 0; l < nl; ++l) srho[l] = 1.0 / (1.0 + sqrt(nrho[l]));
    }
    // jacobian->ScaleColumns; squared column norms of the scaled Jacobian kept for the LM diagonal
    for (int64_t o = 0; o < no; ++o) {
      const int l = P.obs_lm[o], hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
      for (int k = 0; k < R; ++k) {
        double* row = &J[(o * R + k) * C];
        for (int i = 0; i < 6; ++i) row[i] = hs >= 0 ? row[i] * scam[hs * cd + i] : 0.0;
        for (int i = 0; i < cd; ++i) row[6 + i] = ts >= 0 ? row[6 + i] * scam[ts * cd + i] : 0.0;
        row[C - 1] *= srho[l];
      }
    }
    for (int i = 0; i < n; ++i) diag_cam[i] = ncam[i] * scam[i] * scam[i];
    for (int l = 0; l < nl; ++l) diag_rho[l] = nrho[l] * srho[l] * srho[l];
    // gradient norms via Plus(x, -g) (trust_region_minimizer.cc:279-298)
    std::vector<double> ng(n), ngr(nl);
    for (int i = 0; i < n; ++i) ng[i] = -gcam[i];
    for (int l = 0; l < nl; ++l) ngr[l] = -grho[l];
    // inactive affine columns have zero gradient already
    State proj;
    plus_all(P, L, x, ng.data(), ngr.data(), &proj);
    grad_norm = sqrt(ambient_sqnorm(P, L, x, &proj, &grad_max));
    t_jac += now() - t0;
    return true;
  };

  const double t_min0 = now();
  pba_iteration it;
  memset(&it, 0, sizeof(it));
  bool ok = eval_grad_jac(true);
  if (!ok) {
    termination = PBA_FAILURE;
    snprintf(message, sizeof(message), "Residual and Jacobian evaluation failed.");
  }
  const double initial_cost = x_cost;
  it.iteration = 0; it.cost = x_cost; it.step_is_valid = 1; it.step_is_successful = 1;
  it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm;
  double current_cost_se = x_cost;  // TrustRegionStepEvaluator::current_cost_ (monotonic steps)
  double min_iteration_cost = x_cost;  // SetSummaryFinalCost (solver_utils.h:52-57)

  while (ok) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (trust_region_minimizer.cc:311-359)
    if (it.step_is_successful) {
      ++num_successful;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; best = x; }
    } else {
      ++num_unsuccessful;
    }
    it.trust_region_radius = radius;
    min_iteration_cost = std::min(min_iteration_cost, it.cost);
    push(it);
    if (it.iteration >= opt->max_num_iterations) {
      snprintf(message, sizeof(message), "Maximum number of iterations reached. Number of iterations: %d.", it.iteration);
      termination = PBA_NO_CONVERGENCE; break;
    }
    if (it.gradient_max_norm <= opt->gradient_tolerance) {
      snprintf(message, sizeof(message), "Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt->gradient_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    if (radius <= opt->min_trust_region_radius) {
      snprintf(message, sizeof(message), "Minimum trust region radius reached.");
      termination = PBA_CONVERGENCE; break;
    }
    const double prev_gmax = it.gradient_max_norm, prev_gnorm = it.gradient_norm;
    const int next = it.iteration + 1;
    memset(&it, 0, sizeof(it));
    it.iteration = next;

    // ---- ComputeTrustRegionStep: LM strategy (levenberg_marquardt_strategy.cc:66-130)
    const double t_l0 = now();
    if (!reuse_diagonal) {
      for (int i = 0; i < n; ++i) Dcam[i] = std::min(std::max(diag_cam[i], opt->min_lm_diagonal), opt->max_lm_diagonal);
      for (int l = 0; l < nl; ++l) Drho[l] = std::min(std::max(diag_rho[l], opt->min_lm_diagonal), opt->max_lm_diagonal);
    }
    std::vector<double> lmc(n), lmr(nl);
    for (int i = 0; i < n; ++i) lmc[i] = sqrt(Dcam[i] / radius);
    for (int l = 0; l < nl; ++l) lmr[l] = sqrt(Drho[l] / radius);
    schur_eliminate(P, L, J.data(), r.data(), lmc.data(), lmr.data(), S, rhs);
    ycam = rhs;
    bool solved = n == 0 ? true : cholesky_solve(n, S, ycam);
    if (solved) for (int i = 0; i < n; ++i) if (!std::isfinite(ycam[i])) solved = false;
    // BackSubstitute (schur_eliminator_impl.h:309-375)
    if (solved) {
      for (int l = 0; l < nl; ++l) {
        double ete = lmr[l] * lmr[l], acc = 0;
        for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) {
          const int hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
          for (int k = 0; k < R; ++k) {
            const double* row = &J[(o * R + k) * C];
            double sj = r[o * R + k];
            if (hs >= 0) for (int i = 0; i < 6; ++i) sj -= row[i] * ycam[hs * cd + i];
            if (ts >= 0) for (int i = 0; i < cd; ++i) sj -= row[6 + i] * ycam[ts * cd + i];
            acc += row[C - 1] * sj;
            ete += row[C - 1] * row[C - 1];
          }
        }
        yrho[l] = p->lm_obs_ptr[l + 1] > p->lm_obs_ptr[l] ? acc / ete : 0.0;
      }
    }
    reuse_diagonal = true;
    ++n_lin;
    t_lin += now() - t_l0;
    it.linear_solver_iterations = 1;
    it.step_is_valid = 0;
    double model_cost_change = 0;
    if (solved) {
      // step = -y; model_cost_change = -(J s)^T (r + J s / 2) (trust_region_minimizer.cc:414-427)
      double mcc = 0;
#pragma omp parallel for schedule(static) reduction(+ : mcc) num_threads(threads)
      for (int64_t o = 0; o < no; ++o) {
        const int l = P.obs_lm[o], hs = L.slot[p->lm_host[l]], ts = L.slot[p->obs_target[o]];
        for (int k = 0; k < R; ++k) {
          const double* row = &J[(o * R + k) * C];
          double m = 0;
          if (hs >= 0) for (int i = 0; i < 6; ++i) m -= row[i] * ycam[hs * cd + i];
          if (ts >= 0) for (int i = 0; i < cd; ++i) m -= row[6 + i] * ycam[ts * cd + i];
          m -= row[C - 1] * yrho[l];
          mcc += -m * (r[o * R + k] + m / 2.0);
        }
      }
      model_cost_change = mcc;
      it.step_is_valid = model_cost_change > 0.0;
    }
    it.model_cost_change = model_cost_change;
    if (!it.step_is_valid) {
      // HandleInvalidStep (trust_region_minimizer.cc:450-485)
      if (++consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
        snprintf(message, sizeof(message), "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt->max_num_consecutive_invalid_steps);
        termination = PBA_FAILURE; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid
      it.cost = x_cost; it.cost_change = 0; it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;
      it.step_norm = 0; it.relative_decrease = 0; it.step_is_successful = 0;
      continue;
    }
    consecutive_invalid = 0;
    for (int i = 0; i < n; ++i) dcam[i] = -ycam[i] * scam[i];
    for (int l = 0; l < nl; ++l) drho[l] = -yrho[l] * srho[l];

    // ComputeCandidatePointAndEvaluateCost (:761-779)
    plus_all(P, L, x, dcam.data(), drho.data(), &cand);
    const double t_r0 = now();
    double cand_cost = evaluate(P, cand.poses.data(), cand.affine.data(), cand.rho.data(), nullptr, nullptr, threads);
    ++n_res;
    t_res += now() - t_r0;
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached (:706-726) — runs BEFORE the accept test
    it.step_norm = sqrt(ambient_sqnorm(P, L, x, &cand, nullptr));
    if (it.step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      snprintf(message, sizeof(message), "Parameter tolerance reached. Relative step_norm: %e <= %e.",
               it.step_norm / (x_norm + opt->parameter_tolerance), opt->parameter_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    // FunctionToleranceReached (:729-748)
    it.cost_change = x_cost - cand_cost;
    if (fabs(it.cost_change) <= opt->function_tolerance * x_cost) {
      snprintf(message, sizeof(message), "Function tolerance reached. |cost_change|/cost: %e <= %e",
               fabs(it.cost_change) / x_cost, opt->function_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    // IsStepSuccessful (:781-807) with the monotonic step evaluator
    it.relative_decrease = cand_cost >= std::numeric_limits<double>::max()
                               ? std::numeric_limits<double>::lowest()
                               : (current_cost_se - cand_cost) / model_cost_change;
    if (it.relative_decrease > opt->min_relative_decrease) {
      // HandleSuccessfulStep (:812-826)
      x = cand;
      x_norm = sqrt(ambient_sqnorm(P, L, x, nullptr, nullptr));
      if (!eval_grad_jac(false)) {
        termination = PBA_FAILURE;
        snprintf(message, sizeof(message), "Residual and Jacobian evaluation failed.");
        break;
      }
      it.cost = x_cost; it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm;
      it.step_is_successful = 1;
      radius = radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      current_cost_se = cand_cost;
    } else {
      it.step_is_successful = 0;
      it.cost = cand_cost;
      it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  const double t_min = now() - t_min0;

  // Solution is usable unless FAILURE (solver.cc:438-447): write back the best iterate.
  if (termination != PBA_FAILURE) {
    memcpy(p->poses, best.poses.data(), sizeof(double) * 7 * p->n_poses);
    if (P.photo && p->affine) memcpy(p->affine, best.affine.data(), sizeof(double) * 2 * p->n_poses);
    memcpy(p->inv_depth, best.rho.data(), sizeof(double) * nl);
  }
  if (sum) {
    sum->termination_type = termination;
    sum->num_iterations = std::min(n_it, cap > 0 ? cap : n_it);
    sum->num_successful_steps = num_successful;
    sum->num_unsuccessful_steps = num_unsuccessful;
    sum->num_residual_evaluations = n_res;
    sum->num_jacobian_evaluations = n_jac;
    sum->num_linear_solves = n_lin;
    sum->rcs_dim = n;
    sum->num_residual_blocks = no;
    sum->num_residuals = no * R;
    sum->num_effective_parameters = n + nl;
    sum->initial_cost = initial_cost;
    sum->final_cost = min_iteration_cost;
    sum->residual_evaluation_time_in_seconds = t_res;
    sum->jacobian_evaluation_time_in_seconds = t_jac;
    sum->linear_solver_time_in_seconds = t_lin;
    sum->minimizer_time_in_seconds = t_min;
    sum->total_time_in_seconds = now();
    snprintf(sum->message, sizeof(sum->message), "%s", message);
  }
  return PBA_OK;
}

// --------------------------------------------------------------------------
// SURVEY.md §8(f)-2/3 restated: Landmark::get_p (common_types.h:205-217),
// compute_projections (src/sfm.cpp:1956-1984, inlier observations),
// set_outlier_flags (src/sfm.cpp:1928-1952) and the keep/remove decision of
// remove_outlier_landmarks (src/sfm.cpp:2039-2091).  Slot layout as pba.h.
static V3<double> landmark_point(const pba_problem* p, int l) {
  const int h = p->lm_host[l];
  const int c = p->pose_calib[h];
  V3<double> b = unproject<double>(p->calib_model[c], p->intrinsics + 8 * c, p->lm_host_uv[2 * l], p->lm_host_uv[2 * l + 1]);
  const double n = sqrt(b.x * b.x + b.y * b.y + b.z * b.z);  // Eigen normalize(): v /= norm
  b = V3<double>{b.x / n, b.y / n, b.z / n};
  const double rho = p->inv_depth[l];
  const V3<double> X{b.x / rho, b.y / rho, b.z / rho};
  return act(map_se3<double>(p->poses + 7 * h), X);
}

ORACLE_API int pba_oracle_landmark_positions(const pba_problem* p, double* p_w) {
  for (int l = 0; l < p->n_landmarks; ++l) {
    const V3<double> w = landmark_point(p, l);
    p_w[3 * l] = w.x; p_w[3 * l + 1] = w.y; p_w[3 * l + 2] = w.z;
  }
  return 0;
}

ORACLE_API int pba_oracle_compute_projections(const pba_problem* p, const pba_projection_thresholds* thr,
                                              double* point_reprojected, double* point_3d_c, double* reprojection_error,
                                              uint32_t* outlier_flags, uint8_t* landmark_remove, int32_t* any_severe) {
  const int64_t ns = p->n_obs + p->n_landmarks;
  std::vector<uint32_t> fl(ns, 0);
  bool severe = false;
  for (int l = 0; l < p->n_landmarks; ++l) {
    const V3<double> w = landmark_point(p, l);
    const int64_t base = p->lm_obs_ptr[l];
    const int64_t cnt = p->lm_obs_ptr[l + 1] - base + 1;
    for (int64_t k = 0; k < cnt; ++k) {
      const int64_t s = base + l + k;
      const int pose = k == 0 ? p->lm_host[l] : p->obs_target[base + k - 1];
      const double* z = k == 0 ? p->lm_host_uv + 2 * l : p->obs_uv + 2 * (base + k - 1);
      const int c = p->pose_calib[pose];
      const V3<double> pc = act(inverse(map_se3<double>(p->poses + 7 * pose)), w);
      double uv[2];
      project<double>(p->calib_model[c], p->intrinsics + 8 * c, pc, uv);
      const double du = z[0] - uv[0], dv = z[1] - uv[1];
      const double e = sqrt(du * du + dv * dv);
      uint32_t f = PBA_OUTLIER_NONE;
      if (e > thr->reprojection_error_huge_pixel) f |= PBA_OUTLIER_REPROJECTION_ERROR_HUGE;
      if (e > thr->reprojection_error_normal_pixel) f |= PBA_OUTLIER_REPROJECTION_ERROR_NORMAL;
      if (sqrt(pc.x * pc.x + pc.y * pc.y + pc.z * pc.z) < thr->camera_center_distance_meter) f |= PBA_OUTLIER_CAMERA_DISTANCE;
      if (pc.z < thr->z_coordinate_meter) f |= PBA_OUTLIER_Z_COORDINATE;
      fl[s] = f;
      if (f & ~uint32_t(PBA_OUTLIER_REPROJECTION_ERROR_NORMAL)) severe = true;
      if (point_reprojected) { point_reprojected[2 * s] = uv[0]; point_reprojected[2 * s + 1] = uv[1]; }
      if (point_3d_c) { point_3d_c[3 * s] = pc.x; point_3d_c[3 * s + 1] = pc.y; point_3d_c[3 * s + 2] = pc.z; }
      if (reprojection_error) reprojection_error[s] = e;
    }
  }
  if (outlier_flags) memcpy(outlier_flags, fl.data(), sizeof(uint32_t) * ns);
  if (any_severe) *any_severe = severe ? 1 : 0;
  if (landmark_remove) {
    for (int l = 0; l < p->n_landmarks; ++l) {
      bool remove = false;
      for (int64_t s = p->lm_obs_ptr[l] + l; s < p->lm_obs_ptr[l + 1] + l + 1 && !remove; ++s) {
        const uint32_t f = fl[s];
        if (f & PBA_OUTLIER_REPROJECTION_ERROR_HUGE) remove = true;
        else if ((f & PBA_OUTLIER_REPROJECTION_ERROR_NORMAL) && !severe) remove = true;
        else if (f & PBA_OUTLIER_CAMERA_DISTANCE) remove = true;
        else if (f & PBA_OUTLIER_Z_COORDINATE) remove = true;
      }
      landmark_remove[l] = remove ? 1 : 0;
    }
  }
  return 0;
}

// ---- add_new_landmarks_between_cams (include/visnav/map_utils.h:121-195) ----
// Bearings (map_utils.h:150-160), relative pose T_c0_c1 = T_w_c0^-1 T_w_c1 (:166-170), the linear triangulation
// of opengv (thirdparty/opengv/src/triangulation/methods.cpp:36-64) and inv_depth = 1 / |p| (map_utils.h:190).
// opengv takes the right singular vector of the smallest singular value of the 4x4 DLT matrix A (Eigen
// JacobiSVD); here it is the eigenvector of the smallest eigenvalue of A^T A by cyclic Jacobi rotations — a
// different route to the same vector than the CUDA kernel's one-sided iteration on A's columns.
static void smallest_right_singular_vector(const double A[4][4], double v[4]) {
  double S[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[k][i] * A[k][j];
      S[i][j] = s;
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) { diag += fabs(S[i][i]); for (int j = i + 1; j < 4; ++j) off += fabs(S[i][j]); }
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (S[p][q] == 0.0) continue;
        const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 4; ++k) { const double a = S[k][p], b = S[k][q]; S[k][p] = c * a - sn * b; S[k][q] = sn * a + c * b; }
        for (int k = 0; k < 4; ++k) { const double a = S[p][k], b = S[q][k]; S[p][k] = c * a - sn * b; S[q][k] = sn * a + c * b; }
        for (int k = 0; k < 4; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - sn * b; V[k][q] = sn * a + c * b; }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (S[i][i] < S[best][best]) best = i;
  for (int k = 0; k < 4; ++k) v[k] = V[k][best];
}

ORACLE_API int pba_oracle_triangulate(int model0, const double* intr0, int model1, const double* intr1, const double* T_w_c0,
                                      const double* T_w_c1, int64_t n, const double* uv0, const double* uv1, double* p_c0,
                                      double* inv_depth) {
  const SE3<double> T0 = map_se3(T_w_c0), T1 = map_se3(T_w_c1);
  const SE3<double> T01 = mul(inverse(T0), T1);
  // rotation matrix of T01 and P2 = [R^T | -R^T t]
  double R[3][3];
  for (int c = 0; c < 3; ++c) {
    V3<double> e{c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
    const V3<double> r = rotate(T01.q, e);
    R[0][c] = r.x; R[1][c] = r.y; R[2][c] = r.z;
  }
  double P2[3][4];
  const double t[3] = {T01.t.x, T01.t.y, T01.t.z};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) P2[i][j] = R[j][i];
    P2[i][3] = -(R[0][i] * t[0] + R[1][i] * t[1] + R[2][i] * t[2]);
  }
  for (int64_t i = 0; i < n; ++i) {
    const V3<double> f1 = unit(unproject<double>(model0, intr0, uv0[2 * i], uv0[2 * i + 1]));
    const V3<double> f2 = unit(unproject<double>(model1, intr1, uv1[2 * i], uv1[2 * i + 1]));
    double A[4][4];
    for (int c = 0; c < 4; ++c) {
      const double p10 = c == 0, p11 = c == 1, p12 = c == 2;
      A[0][c] = f1.x * p12 - f1.z * p10;
      A[1][c] = f1.y * p12 - f1.z * p11;
      A[2][c] = f2.x * P2[2][c] - f2.z * P2[0][c];
      A[3][c] = f2.y * P2[2][c] - f2.z * P2[1][c];
    }
    double v[4];
    smallest_right_singular_vector(A, v);
    const double x = v[0] / v[3], y = v[1] / v[3], z = v[2] / v[3];
    if (p_c0) { p_c0[3 * i] = x; p_c0[3 * i + 1] = y; p_c0[3 * i + 2] = z; }
    inv_depth[i] = 1.0 / sqrt(x * x + y * y + z * z);
  }
  return 0;
}


// ===================================================================== front-end ====
// CPU restatement of the reference's descriptor / matching / epipolar functions, pinned by
// tests/golden/frontend_euroc.npz (made by tests/golden/make_golden_frontend.py from the reference's own
// functions, oracle/ref/frontend_harness.cpp) and, when oracle/_ref travelled, against that library live.
//   computeAngles ........ include/visnav/keypoints.h:182-212
//   computeDescriptors ... include/visnav/keypoints.h:214-245
//   matchSets / matchDescriptors ... include/visnav/keypoints.h:248-300
//   computeEssential / findInliersEssential ... include/visnav/matching_utils.h:50-79
namespace {
// the reference's four point tables (keypoints.h:54-131), one array per coordinate as there
const signed char kXa[256] = {
    8,4,-11,7,2,1,-2,-13,-13,10,-13,-11,7,-4,-13,-9,12,-3,-6,11,4,5,3,-8,-2,-13,-7,-4,-10,5,5,1,
    9,4,2,-4,-8,4,0,-13,-3,-6,8,0,7,-13,10,-6,10,-13,-13,3,5,-1,3,2,-13,-13,-13,-7,6,-9,-2,-12,
    3,-7,-3,2,-11,-1,5,-4,-9,-12,10,7,-7,-4,7,-7,-13,-3,7,-13,1,2,-4,-1,7,1,9,-1,-13,7,12,6,
    5,2,3,2,9,-8,-11,1,6,2,6,3,7,-11,-10,-5,-10,8,4,-10,4,-2,-5,7,-9,-5,8,-9,1,7,-2,11,
    -12,3,5,0,-9,0,-1,5,3,-13,-5,-4,6,-7,-13,1,4,-2,2,-2,4,-6,-3,7,4,-13,7,7,-7,-8,-13,2,
    10,-6,8,2,-11,-12,-11,5,-2,-1,-13,-10,-3,2,-9,-4,-4,-6,6,-13,11,7,-1,-4,-7,-13,-7,-8,-5,-13,1,1,
    9,5,-1,-9,-1,-13,8,2,7,-10,-10,4,3,-4,5,4,-9,0,-12,3,-10,8,-8,2,10,6,-7,-3,-1,-3,-8,4,
    2,6,3,11,-3,4,2,-10,-13,-13,6,0,-13,-9,-13,5,2,-1,9,11,3,-1,3,-13,5,8,7,-10,7,9,7,-1
};
const signed char kYa[256] = {
    -3,2,9,-12,-13,-7,-10,-13,-3,4,-8,7,7,-5,2,0,-6,6,-13,-13,7,-3,-7,-7,11,12,3,2,-12,-12,-6,0,
    11,7,-1,-12,-5,11,-8,-2,-2,9,12,9,-5,-6,7,-3,-9,8,0,3,7,7,-10,-4,0,-7,3,12,-10,-1,-5,5,
    -10,-7,-2,9,-13,6,-3,-13,-6,-10,2,12,-13,9,-1,6,11,7,-8,-7,-3,-6,3,-13,1,-1,1,-9,-13,7,-5,3,
    -13,-12,8,6,-12,4,12,12,-9,3,3,-3,8,-5,11,-8,5,-1,-6,12,-2,0,-8,-6,-13,-13,-8,-11,-8,-4,1,-6,
    -9,7,5,-4,12,7,2,11,5,-4,9,-7,5,6,6,-10,1,-2,-12,-13,1,-10,-13,5,-2,9,1,-8,-4,11,6,4,
    -5,-5,-3,-12,-2,-13,0,-3,-13,-8,-11,-2,9,-3,-13,6,12,-11,-3,11,11,-5,12,-8,1,-12,-2,5,-1,7,5,0,
    12,-8,11,-3,-10,1,-11,-13,-13,-10,-8,-6,12,2,-13,-13,9,3,1,2,-10,-13,-12,2,6,8,10,-9,-13,-7,-2,2,
    -5,-9,-1,-1,0,-11,-4,-6,7,12,0,-1,3,8,-6,-9,7,-6,5,-3,0,4,-6,0,8,9,-4,4,3,-7,0,-6
};
const signed char kXb[256] = {
    9,7,-8,12,2,1,-2,-11,-12,11,-8,-9,12,-3,-12,-7,12,-2,-4,12,5,10,6,-6,-1,-8,-5,-3,-6,6,7,4,
    11,4,4,-2,-7,9,1,-8,-2,-4,10,1,11,-11,12,-6,12,-8,-8,7,10,1,5,3,-13,-12,-11,-4,12,-7,0,-7,
    8,-4,-1,5,-5,0,5,-4,-9,-8,12,12,-6,-3,12,-5,-12,-2,12,-11,12,3,-2,1,8,3,12,-1,-10,10,12,7,
    6,2,4,12,10,-7,-4,2,7,3,11,8,9,-6,-5,-3,-9,12,6,-8,6,-2,-5,10,-8,-5,9,-9,1,9,-1,12,
    -6,7,10,2,-5,2,1,7,6,-8,-3,-3,8,-6,-5,3,8,2,12,0,9,-3,-1,12,5,-9,8,7,-7,-7,-12,3,
    12,-6,9,2,-10,-7,-10,11,-1,0,-12,-10,-2,3,-4,-3,-2,-4,6,-5,12,12,0,-3,-6,-8,-6,-6,-4,-8,5,10,
    10,10,1,-6,1,-8,10,3,12,-5,-8,8,8,-3,10,5,-4,3,-6,4,-10,12,-6,3,11,8,-6,-3,-1,-3,-8,12,
    3,11,7,12,-3,4,2,-8,-11,-11,11,1,-9,-6,-8,8,3,-1,11,12,3,0,4,-10,12,9,8,-10,12,10,12,0
};
const signed char kYb[256] = {
    5,-12,2,-13,12,6,-4,-8,-9,9,-9,12,6,0,-3,5,-1,12,-8,-8,1,-3,12,-2,-10,10,-3,7,11,-7,-1,-5,
    -13,12,4,7,-10,12,-13,2,3,-9,7,3,-10,0,1,12,-4,-12,-4,8,-7,-12,6,-10,5,12,8,7,8,-6,12,5,
    -13,5,-7,-11,-13,-1,2,12,6,-4,-3,12,5,4,2,1,5,-6,-7,-12,12,0,-13,9,-6,12,6,3,5,12,9,11,
    10,3,-6,-13,3,9,-6,-8,-4,-2,0,-8,3,-4,10,12,0,-6,-11,7,7,12,2,12,-8,-2,-13,0,-2,1,-4,-11,
    4,12,8,8,-13,12,7,-9,-8,9,-3,-12,0,12,-2,10,-4,-13,12,-6,3,-5,1,-11,-7,-5,6,6,1,-8,-8,9,
    3,7,-8,8,3,-9,-5,8,12,9,-5,11,-13,2,0,-10,-7,9,11,5,6,-2,7,-2,7,-13,-8,-9,5,10,-13,-13,
    -1,-9,-13,2,12,-10,-6,-6,-9,-7,-13,5,-13,-3,-12,-1,3,-9,1,-8,9,12,-5,7,-8,-12,5,9,5,4,3,12,
    11,-13,12,4,6,12,1,1,1,-13,-13,4,-2,-3,-2,10,-9,-1,-2,-8,5,10,5,5,11,-6,-12,9,4,-2,-2,-11
};

inline int px(const uint8_t* img, int pitch, int x, int y) { return img[int64_t(y) * pitch + x]; }
}  // namespace

ORACLE_API int pba_oracle_corner_descriptors(const uint8_t* image, int w, int h, int pitch, int n, const double* corners,
                                             int rotate_features, double* angles, uint8_t* descriptors) {
  (void)w; (void)h;
  for (int i = 0; i < n; ++i) {
    const int cx = int(corners[2 * i]), cy = int(corners[2 * i + 1]);
    double angle = 0.0;
    if (rotate_features) {
      double m01 = 0.0, m10 = 0.0;
      for (int x = -15; x <= 15; ++x) {
        const int yb = int(sqrt(double(15 * 15 - x * x)));
        for (int y = -yb; y <= yb; ++y) {
          const int v = px(image, pitch, cx + x, cy + y);
          m01 += y * v;
          m10 += x * v;
        }
      }
      angle = atan2(m01, m10);
    }
    angles[i] = angle;
    const double cs = cos(angle), sn = sin(angle);
    uint8_t* d = descriptors + 32 * int64_t(i);
    memset(d, 0, 32);
    for (int b = 0; b < 256; ++b) {
      const int xa = int(round(cs * kXa[b] - sn * kYa[b])), ya = int(round(sn * kXa[b] + cs * kYa[b]));
      const int xb = int(round(cs * kXb[b] - sn * kYb[b])), yb = int(round(sn * kXb[b] + cs * kYb[b]));
      if (px(image, pitch, cx + xa, cy + ya) < px(image, pitch, cx + xb, cy + yb)) d[b / 8] |= uint8_t(1u << (b % 8));
    }
  }
  return 0;
}

namespace {
int hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int k = 0; k < 32; ++k) d += __builtin_popcount(unsigned(a[k] ^ b[k]));
  return d;
}
void match_sets(int n1, const uint8_t* d1, int n2, const uint8_t* d2, int threshold, double ratio, std::vector<int>& out) {
  out.assign(n1, -1);
  for (int i = 0; i < n1; ++i) {
    int smallest = 256, second = 256, best = 0;
    for (int j = 0; j < n2; ++j) {
      const int dist = hamming256(d1 + 32 * int64_t(i), d2 + 32 * int64_t(j));
      if (dist < smallest) { second = smallest; smallest = dist; best = j; }
      else if (dist < second) second = dist;
    }
    if (smallest >= threshold) best = -1;
    if (second < smallest * ratio) best = -1;
    if (n2 == 0) best = -1;
    out[i] = best;
  }
}
}  // namespace

// matches [min(n1, n2)][2], ascending in the first index; returns the count
ORACLE_API int pba_oracle_match_descriptors(int n1, const uint8_t* d1, int n2, const uint8_t* d2, int threshold,
                                            double dist_2_best, int32_t* matches) {
  std::vector<int> pq, qp;
  match_sets(n1, d1, n2, d2, threshold, dist_2_best, pq);
  match_sets(n2, d2, n1, d1, threshold, dist_2_best, qp);
  int n = 0;
  for (int i = 0; i < n1; ++i)
    if (pq[i] != -1 && qp[pq[i]] == i) { matches[2 * n] = i; matches[2 * n + 1] = pq[i]; ++n; }
  return n;
}

ORACLE_API int pba_oracle_epipolar_inliers(int model0, const double* intr0, int model1, const double* intr1,
                                           const double* T_0_1, double threshold, int64_t n_matches, const int32_t* matches,
                                           const double* corners0, const double* corners1, double* E_out, uint8_t* inlier) {
  // E = [t / |t|]x R  (matching_utils.h:50-60)
  const Q4<double> q{T_0_1[0], T_0_1[1], T_0_1[2], T_0_1[3]};
  double R[3][3];
  for (int c = 0; c < 3; ++c) {
    V3<double> e{c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
    const V3<double> r = rotate(q, e);
    R[0][c] = r.x; R[1][c] = r.y; R[2][c] = r.z;
  }
  const double nt = sqrt(T_0_1[4] * T_0_1[4] + T_0_1[5] * T_0_1[5] + T_0_1[6] * T_0_1[6]);
  const double t[3] = {T_0_1[4] / nt, T_0_1[5] / nt, T_0_1[6] / nt};
  const double S[3][3] = {{0.0, -t[2], t[1]}, {t[2], 0.0, -t[0]}, {-t[1], t[0], 0.0}};
  double E[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) E[i][j] = S[i][0] * R[0][j] + S[i][1] * R[1][j] + S[i][2] * R[2][j];
  if (E_out)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) E_out[3 * i + j] = E[i][j];
  int n_in = 0;
  for (int64_t k = 0; k < n_matches; ++k) {
    const int i = matches[2 * k], j = matches[2 * k + 1];
    const V3<double> xl = unproject<double>(model0, intr0, corners0[2 * i], corners0[2 * i + 1]);
    const V3<double> xr = unproject<double>(model1, intr1, corners1[2 * j], corners1[2 * j + 1]);
    const double e0 = E[0][0] * xr.x + E[0][1] * xr.y + E[0][2] * xr.z;
    const double e1 = E[1][0] * xr.x + E[1][1] * xr.y + E[1][2] * xr.z;
    const double e2 = E[2][0] * xr.x + E[2][1] * xr.y + E[2][2] * xr.z;
    inlier[k] = fabs(xl.x * e0 + xl.y * e1 + xl.z * e2) <= threshold ? 1 : 0;
    n_in += inlier[k];
  }
  return n_in;
}

// TrackBuilder::Build / Filter / Export (include/visnav/tracks.h:53-160), restated with a plain sequential
// union-find.  Node id = feat_ptr[image] + feature; track id = smallest node id of the component (the reference's ids
// are the roots of ITS forest; the partition and the filter decisions are what is defined).
ORACLE_API int pba_oracle_build_tracks(int n_images, const int32_t* feat_ptr, int n_pairs, const int32_t* pairs,
                                       const int64_t* match_ptr, const int32_t* matches, int min_length, int32_t* track_of) {
  const int n = n_images > 0 ? feat_ptr[n_images] : 0;
  std::vector<int> parent(n), image(n);
  std::vector<char> touched(n, 0);
  for (int i = 0; i < n; ++i) parent[i] = i;
  for (int im = 0; im < n_images; ++im)
    for (int i = feat_ptr[im]; i < feat_ptr[im + 1]; ++i) image[i] = im;
  auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
  for (int k = 0; k < n_pairs; ++k)
    for (int64_t e = match_ptr[k]; e < match_ptr[k + 1]; ++e) {
      const int u = feat_ptr[pairs[2 * k]] + matches[2 * e], v = feat_ptr[pairs[2 * k + 1]] + matches[2 * e + 1];
      touched[u] = touched[v] = 1;
      const int ru = find(u), rv = find(v);
      if (ru != rv) parent[std::max(ru, rv)] = std::min(ru, rv);
    }
  // a component's nodes in ascending id are ascending in image: two equal consecutive images = a conflict
  std::vector<int> count(n, 0), last_image(n, -1);
  std::vector<char> bad(n, 0);
  for (int i = 0; i < n; ++i) {
    if (!touched[i]) continue;
    const int r = find(i);
    ++count[r];
    if (last_image[r] == image[i]) bad[r] = 1;
    last_image[r] = image[i];
  }
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    int t = -1;
    if (touched[i]) {
      const int r = find(i);
      if (count[r] >= min_length && !bad[r]) { t = r; kept += r == i; }
    }
    track_of[i] = t;
  }
  return kept;
}
