// projections.cu — SURVEY.md §8(f)-2/3: the callers either side of
// bundle_adjustment() in the reference's SfM loop, on the device.
//
//   Landmark::get_p            include/visnav/common_types.h:205-217
//   compute_projections        src/sfm.cpp:1956-2008  (inlier observations)
//   set_outlier_flags          src/sfm.cpp:1928-1952
//   remove_outlier_landmarks   src/sfm.cpp:2029-2100  (keep/remove decision only)
//
// Three kernels, all HBM-streaming:
//   k_landmark_points   thread per landmark -> p_w                      (get_p)
//   k_project_slots     thread per observation slot (host obs + CSR obs of a
//                       landmark are consecutive slots) -> p_c, reprojection,
//                       error, flags; warp-aggregated OR of "severe" flags
//   k_landmark_remove   thread per landmark -> remove decision
#include <string.h>

#include <vector>

#include "launch.h"
#include "pba_internal.h"

#define PBA_API extern "C" __attribute__((visibility("default")))

namespace pba {
namespace {

struct ProjView {
  int n_lm;
  int64_t n_slots;
  const double* poses;        // [n_poses*7]
  const int* pose_calib;      // [n_poses]
  const int* calib_model;     // [n_calib]
  const double* intr;         // [n_calib*8]
  const double* inv_depth;    // [n_lm]
  const int* lm_host;         // [n_lm]
  const double* lm_host_uv;   // [n_lm*2]
  const int64_t* lm_obs_ptr;  // [n_lm+1]
  const int* obs_target;      // [n_obs]
  const double* obs_uv;       // [n_obs*2]
};

struct Thr { double huge, normal, dist, z; };

// common_types.h:205-217: T_w_c * (unproject(z).normalized() / inv_depth)
__global__ void k_landmark_points(ProjView v, double* __restrict__ p_w) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= v.n_lm) return;
  const int h = v.lm_host[l];
  const int c = v.pose_calib[h];
  double intr[8], T[7], b[3], X[3], out[3];
#pragma unroll
  for (int k = 0; k < 8; ++k) intr[k] = v.intr[8 * c + k];
#pragma unroll
  for (int k = 0; k < 7; ++k) T[k] = v.poses[7 * h + k];
  cam_bearing(v.calib_model[c], intr, v.lm_host_uv[2 * l], v.lm_host_uv[2 * l + 1], b);
  const double rho = v.inv_depth[l];
  X[0] = b[0] / rho; X[1] = b[1] / rho; X[2] = b[2] / rho;
  quat_rotate(T, X, out);
  p_w[3 * l] = out[0] + T[4]; p_w[3 * l + 1] = out[1] + T[5]; p_w[3 * l + 2] = out[2] + T[6];
}

// Largest l with lm_obs_ptr[l] + l <= s.
__device__ __forceinline__ int slot_landmark(const int64_t* __restrict__ ptr, int n_lm, int64_t s) {
  int lo = 0, hi = n_lm - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ptr[mid] + mid <= s) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// src/sfm.cpp:1964-1984 + :1928-1952 for one observation slot.
__global__ void k_project_slots(ProjView v, Thr thr, const double* __restrict__ p_w, double* __restrict__ repro,
                                double* __restrict__ p3c, double* __restrict__ err, uint32_t* __restrict__ flags,
                                unsigned* __restrict__ any_severe) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool live = s < v.n_slots;
  bool severe = false;
  if (live) {
    const int l = slot_landmark(v.lm_obs_ptr, v.n_lm, s);
    const int64_t base = v.lm_obs_ptr[l];
    const int64_t k = s - (base + l);
    int pose;
    double zu, zv;
    if (k == 0) {
      pose = v.lm_host[l];
      zu = v.lm_host_uv[2 * l]; zv = v.lm_host_uv[2 * l + 1];
    } else {
      const int64_t o = base + k - 1;
      pose = v.obs_target[o];
      zu = v.obs_uv[2 * o]; zv = v.obs_uv[2 * o + 1];
    }
    const int c = v.pose_calib[pose];
    double intr[8], T[7];
#pragma unroll
    for (int i = 0; i < 8; ++i) intr[i] = v.intr[8 * c + i];
#pragma unroll
    for (int i = 0; i < 7; ++i) T[i] = v.poses[7 * pose + i];
    // T_w_c.inverse(): SO3 from the conjugate (re-normalised, so3.hpp:482-489), t' = -(R^-1 t)  (se3.hpp:206-211)
    const double qn = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
    const double qi[4] = {-T[0] / qn, -T[1] / qn, -T[2] / qn, T[3] / qn};
    const double nt[3] = {-T[4], -T[5], -T[6]};
    double ti[3], pc[3];
    quat_rotate(qi, nt, ti);
    const double pw[3] = {p_w[3 * l], p_w[3 * l + 1], p_w[3 * l + 2]};
    quat_rotate(qi, pw, pc);
    pc[0] += ti[0]; pc[1] += ti[1]; pc[2] += ti[2];
    double uv[2];
    cam_project<false>(v.calib_model[c], intr, pc[0], pc[1], pc[2], uv, nullptr);
    const double du = zu - uv[0], dv = zv - uv[1];
    const double e = sqrt(du * du + dv * dv);
    uint32_t f = 0;
    if (e > thr.huge) f |= PBA_OUTLIER_REPROJECTION_ERROR_HUGE;
    if (e > thr.normal) f |= PBA_OUTLIER_REPROJECTION_ERROR_NORMAL;
    if (sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]) < thr.dist) f |= PBA_OUTLIER_CAMERA_DISTANCE;
    if (pc[2] < thr.z) f |= PBA_OUTLIER_Z_COORDINATE;
    if (repro) { repro[2 * s] = uv[0]; repro[2 * s + 1] = uv[1]; }
    if (p3c) { p3c[3 * s] = pc[0]; p3c[3 * s + 1] = pc[1]; p3c[3 * s + 2] = pc[2]; }
    if (err) err[s] = e;
    flags[s] = f;
    severe = (f & ~uint32_t(PBA_OUTLIER_REPROJECTION_ERROR_NORMAL)) != 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, severe);
  if (m && (threadIdx.x & 31) == 0) atomicOr(any_severe, 1u);
}

// src/sfm.cpp:2056-2091: a landmark goes when any of its projections is a
// huge / camera-distance / z outlier, or a "normal" one while no severe
// outlier exists anywhere in the map.
__global__ void k_landmark_remove(ProjView v, const uint32_t* __restrict__ flags,
                                  const unsigned* __restrict__ any_severe, uint8_t* __restrict__ remove) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= v.n_lm) return;
  const uint32_t severe_mask =
      PBA_OUTLIER_REPROJECTION_ERROR_HUGE | PBA_OUTLIER_CAMERA_DISTANCE | PBA_OUTLIER_Z_COORDINATE;
  const uint32_t mask = *any_severe ? severe_mask : (severe_mask | PBA_OUTLIER_REPROJECTION_ERROR_NORMAL);
  const int64_t s0 = v.lm_obs_ptr[l] + l, s1 = v.lm_obs_ptr[l + 1] + l + 1;
  uint32_t acc = 0;
  for (int64_t s = s0; s < s1; ++s) acc |= flags[s];
  remove[l] = (acc & mask) ? 1 : 0;
}

bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

// Index sanity of the arrays the projection path touches.
bool valid_landmark_arrays(const pba_problem* p, bool with_obs) {
  if (!p || p->n_poses <= 0 || p->n_calib <= 0 || p->n_landmarks < 0) return false;
  if (!p->poses || !p->pose_calib || !p->calib_model || !p->intrinsics) return false;
  if (p->n_landmarks && (!p->inv_depth || !p->lm_host || !p->lm_host_uv)) return false;
  for (int i = 0; i < p->n_poses; ++i)
    if (p->pose_calib[i] < 0 || p->pose_calib[i] >= p->n_calib) return false;
  for (int i = 0; i < p->n_calib; ++i)
    if (p->calib_model[i] < 0 || p->calib_model[i] > PBA_CAM_EUCM) return false;
  bool ok = true;
#pragma omp parallel for reduction(&& : ok) schedule(static)
  for (int l = 0; l < p->n_landmarks; ++l) ok = ok && p->lm_host[l] >= 0 && p->lm_host[l] < p->n_poses;
  if (!ok || !with_obs) return ok;
  if (p->n_obs < 0 || !p->lm_obs_ptr || (p->n_obs && (!p->obs_target || !p->obs_uv))) return false;
  if (p->lm_obs_ptr[0] != 0 || p->lm_obs_ptr[p->n_landmarks] != p->n_obs) return false;
#pragma omp parallel for reduction(&& : ok) schedule(static)
  for (int l = 0; l < p->n_landmarks; ++l) ok = ok && p->lm_obs_ptr[l] <= p->lm_obs_ptr[l + 1];
#pragma omp parallel for reduction(&& : ok) schedule(static)
  for (int64_t o = 0; o < p->n_obs; ++o) ok = ok && p->obs_target[o] >= 0 && p->obs_target[o] < p->n_poses;
  return ok;
}

struct ProjBuffers {
  DevBuf<double> poses, intr, inv_depth, host_uv, obs_uv, p_w;
  DevBuf<int> pose_calib, calib_model, lm_host, obs_target;
  DevBuf<int64_t> obs_ptr;
};

template <class T>
cudaError_t put(DevBuf<T>& d, const T* src, size_t n, cudaStream_t s) {
  cudaError_t e = d.alloc(n);
  if (e != cudaSuccess || n == 0) return e;
  return cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
}

pba_status upload_view(const pba_problem* p, bool with_obs, cudaStream_t st, ProjBuffers* b, ProjView* v) {
  const size_t nl = size_t(p->n_landmarks);
  PBA_CUDA_OK(put(b->poses, (const double*)p->poses, size_t(p->n_poses) * 7, st));
  PBA_CUDA_OK(put(b->pose_calib, p->pose_calib, size_t(p->n_poses), st));
  PBA_CUDA_OK(put(b->calib_model, p->calib_model, size_t(p->n_calib), st));
  PBA_CUDA_OK(put(b->intr, p->intrinsics, size_t(p->n_calib) * 8, st));
  PBA_CUDA_OK(put(b->inv_depth, (const double*)p->inv_depth, nl, st));
  PBA_CUDA_OK(put(b->lm_host, p->lm_host, nl, st));
  PBA_CUDA_OK(put(b->host_uv, p->lm_host_uv, nl * 2, st));
  PBA_CUDA_OK(b->p_w.alloc(nl * 3));
  if (with_obs) {
    PBA_CUDA_OK(put(b->obs_ptr, p->lm_obs_ptr, nl + 1, st));
    PBA_CUDA_OK(put(b->obs_target, p->obs_target, size_t(p->n_obs), st));
    PBA_CUDA_OK(put(b->obs_uv, p->obs_uv, size_t(p->n_obs) * 2, st));
  }
  v->n_lm = p->n_landmarks;
  v->n_slots = with_obs ? p->n_obs + p->n_landmarks : 0;
  v->poses = b->poses.p; v->pose_calib = b->pose_calib.p; v->calib_model = b->calib_model.p; v->intr = b->intr.p;
  v->inv_depth = b->inv_depth.p; v->lm_host = b->lm_host.p; v->lm_host_uv = b->host_uv.p;
  v->lm_obs_ptr = b->obs_ptr.p; v->obs_target = b->obs_target.p; v->obs_uv = b->obs_uv.p;
  return PBA_OK;
}

}  // namespace
}  // namespace pba

using namespace pba;

namespace pba {
namespace {

// add_new_landmarks_between_cams (include/visnav/map_utils.h:121-195): a thread per shared track.
// v0 / v1 = unit bearings of the two corners; the point is opengv's linear triangulation in camera 0's frame
// (thirdparty/opengv/src/triangulation/methods.cpp:36-64): rows f_x P_2 - f_z P_0, f_y P_2 - f_z P_1 of the two
// projection matrices P1 = [I | 0], P2 = [R12^T | -R12^T t12], the null vector of that 4x4 matrix by SVD.  Here
// the SVD is a one-sided (Hestenes) Jacobi iteration on the columns, which is as accurate as Eigen's two-sided
// JacobiSVD for the small singular vector; the landmark's inverse distance is 1 / |p| (map_utils.h:190).
struct TriArgs {
  int64_t n;
  int model0, model1;
  double intr0[8], intr1[8];
  double Rt[9];   // R12^T
  double mt[3];   // -R12^T t12
};

__global__ void k_triangulate(TriArgs a, const double* __restrict__ uv0, const double* __restrict__ uv1,
                              double* __restrict__ p_out, double* __restrict__ inv_depth) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double f1[3], f2[3];
  cam_bearing(a.model0, a.intr0, uv0[2 * i], uv0[2 * i + 1], f1);
  cam_bearing(a.model1, a.intr1, uv1[2 * i], uv1[2 * i + 1], f2);
  // U = A (row-major 4x4), V = I
  double U[4][4], V[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double p10 = c == 0 ? 1.0 : 0.0, p11 = c == 1 ? 1.0 : 0.0, p12 = c == 2 ? 1.0 : 0.0;
    const double p20 = c < 3 ? a.Rt[c] : a.mt[0], p21 = c < 3 ? a.Rt[3 + c] : a.mt[1], p22 = c < 3 ? a.Rt[6 + c] : a.mt[2];
    U[0][c] = f1[0] * p12 - f1[2] * p10;
    U[1][c] = f1[1] * p12 - f1[2] * p11;
    U[2][c] = f2[0] * p22 - f2[2] * p20;
    U[3][c] = f2[1] * p22 - f2[2] * p21;
#pragma unroll
    for (int r = 0; r < 4; ++r) V[r][c] = r == c ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) { al += U[r][p] * U[r][p]; be += U[r][q] * U[r][q]; ga += U[r][p] * U[r][q]; }
        const double lim = 1e-16 * sqrt(al * be);
        if (fabs(ga) <= lim || ga == 0.0) continue;
        off = fmax(off, fabs(ga) / fmax(lim * 1e16, 1e-300));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double up = U[r][p], uq = U[r][q];
          U[r][p] = c * up - sn * uq; U[r][q] = sn * up + c * uq;
          const double vp = V[r][p], vq = V[r][q];
          V[r][p] = c * vp - sn * vq; V[r][q] = sn * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  int best = 0;
  double smin = 1e300;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double s2 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) s2 += U[r][c] * U[r][c];
    if (s2 < smin) { smin = s2; best = c; }
  }
  double v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) v[r] = best == 0 ? V[r][0] : best == 1 ? V[r][1] : best == 2 ? V[r][2] : V[r][3];
  const double px = v[0] / v[3], py = v[1] / v[3], pz = v[2] / v[3];
  if (p_out) { p_out[3 * i] = px; p_out[3 * i + 1] = py; p_out[3 * i + 2] = pz; }
  inv_depth[i] = 1.0 / sqrt(px * px + py * py + pz * pz);
}

}  // namespace
}  // namespace pba

PBA_API pba_status pba_triangulate_inverse_depth(int32_t model0, const double intr0[8], int32_t model1, const double intr1[8],
                                                 const double T_w_c0[7], const double T_w_c1[7], int64_t n,
                                                 const double* uv0, const double* uv1, int32_t device, double* p_c0,
                                                 double* inv_depth) {
  if (!intr0 || !intr1 || !T_w_c0 || !T_w_c1 || n < 0 || (n > 0 && (!uv0 || !uv1 || !inv_depth))) return PBA_ERR_INVALID_ARGUMENT;
  if (model0 < 0 || model0 > PBA_CAM_EUCM || model1 < 0 || model1 > PBA_CAM_EUCM) return PBA_ERR_UNSUPPORTED;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (n == 0) return PBA_OK;
  TriArgs a;
  a.n = n; a.model0 = model0; a.model1 = model1;
  memcpy(a.intr0, intr0, sizeof(a.intr0)); memcpy(a.intr1, intr1, sizeof(a.intr1));
  // T_c0_c1 = T_w_c0^-1 T_w_c1: R12 = R0^T R1, t12 = R0^T (t1 - t0); P2 = [R12^T | -R12^T t12]
  double R0[9], R1[9], R12[9], t12[3];
  quat_to_rot(T_w_c0, R0);
  quat_to_rot(T_w_c1, R1);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R12[3 * i + j] = R0[i] * R1[j] + R0[3 + i] * R1[3 + j] + R0[6 + i] * R1[6 + j];
  const double d[3] = {T_w_c1[4] - T_w_c0[4], T_w_c1[5] - T_w_c0[5], T_w_c1[6] - T_w_c0[6]};
  for (int i = 0; i < 3; ++i) t12[i] = R0[i] * d[0] + R0[3 + i] * d[1] + R0[6 + i] * d[2];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) a.Rt[3 * i + j] = R12[3 * j + i];
    a.mt[i] = -(R12[i] * t12[0] + R12[3 + i] * t12[1] + R12[6 + i] * t12[2]);
  }
  DevBuf<double> d0, d1, dp, dr;
  PBA_CUDA_OK(d0.alloc(size_t(2) * n)); PBA_CUDA_OK(d1.alloc(size_t(2) * n));
  PBA_CUDA_OK(dp.alloc(size_t(3) * n)); PBA_CUDA_OK(dr.alloc(size_t(n)));
  PBA_CUDA_OK(cudaMemcpy(d0.p, uv0, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(d1.p, uv1, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
  k_triangulate<<<unsigned((n + 127) / 128), 128>>>(a, d0.p, d1.p, dp.p, dr.p);
  PBA_CUDA_OK(cudaGetLastError());
  if (p_c0) PBA_CUDA_OK(cudaMemcpy(p_c0, dp.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
  PBA_CUDA_OK(cudaMemcpy(inv_depth, dr.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API void pba_projection_thresholds_init(pba_projection_thresholds* t) {
  if (!t) return;
  t->reprojection_error_huge_pixel = 40.0;   // src/sfm.cpp:256-257
  t->reprojection_error_normal_pixel = 3.0;  // src/sfm.cpp:254-255
  t->camera_center_distance_meter = 0.1;     // src/sfm.cpp:258-259
  t->z_coordinate_meter = 0.05;              // src/sfm.cpp:260-261
}

PBA_API pba_status pba_landmark_positions(const pba_problem* p, int32_t device, double* p_w) {
  if (!p_w || !valid_landmark_arrays(p, false)) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (p->n_landmarks == 0) return PBA_OK;
  ProjBuffers b;
  ProjView v;
  pba_status st = upload_view(p, false, 0, &b, &v);
  if (st != PBA_OK) return st;
  k_landmark_points<<<unsigned((v.n_lm + 127) / 128), 128>>>(v, b.p_w.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(p_w, b.p_w.p, sizeof(double) * 3 * size_t(v.n_lm), cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_compute_projections(const pba_problem* p, const pba_projection_thresholds* thresholds,
                                           int32_t device, double* point_reprojected, double* point_3d_c,
                                           double* reprojection_error, uint32_t* outlier_flags,
                                           uint8_t* landmark_remove, int32_t* any_severe_outliers) {
  if (!thresholds || !valid_landmark_arrays(p, true)) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (any_severe_outliers) *any_severe_outliers = 0;
  if (p->n_landmarks == 0) return PBA_OK;
  ProjBuffers b;
  ProjView v;
  pba_status st = upload_view(p, true, 0, &b, &v);
  if (st != PBA_OK) return st;
  const size_t ns = size_t(v.n_slots);
  DevBuf<double> repro, p3c, err;
  DevBuf<uint32_t> flags;
  DevBuf<unsigned> severe;
  DevBuf<uint8_t> remove;
  if (point_reprojected) PBA_CUDA_OK(repro.alloc(ns * 2));
  if (point_3d_c) PBA_CUDA_OK(p3c.alloc(ns * 3));
  if (reprojection_error) PBA_CUDA_OK(err.alloc(ns));
  PBA_CUDA_OK(flags.alloc(ns));
  PBA_CUDA_OK(severe.alloc(1));
  PBA_CUDA_OK(remove.alloc(size_t(v.n_lm)));
  PBA_CUDA_OK(cudaMemsetAsync(severe.p, 0, sizeof(unsigned), 0));
  const Thr thr = {thresholds->reprojection_error_huge_pixel, thresholds->reprojection_error_normal_pixel,
                   thresholds->camera_center_distance_meter, thresholds->z_coordinate_meter};
  k_landmark_points<<<unsigned((v.n_lm + 127) / 128), 128>>>(v, b.p_w.p);
  PBA_CUDA_OK(cudaGetLastError());
  k_project_slots<<<unsigned((v.n_slots + 127) / 128), 128>>>(v, thr, b.p_w.p, repro.p, p3c.p, err.p, flags.p, severe.p);
  PBA_CUDA_OK(cudaGetLastError());
  k_landmark_remove<<<unsigned((v.n_lm + 127) / 128), 128>>>(v, flags.p, severe.p, remove.p);
  PBA_CUDA_OK(cudaGetLastError());
  if (point_reprojected) PBA_CUDA_OK(cudaMemcpyAsync(point_reprojected, repro.p, sizeof(double) * 2 * ns, cudaMemcpyDeviceToHost, 0));
  if (point_3d_c) PBA_CUDA_OK(cudaMemcpyAsync(point_3d_c, p3c.p, sizeof(double) * 3 * ns, cudaMemcpyDeviceToHost, 0));
  if (reprojection_error) PBA_CUDA_OK(cudaMemcpyAsync(reprojection_error, err.p, sizeof(double) * ns, cudaMemcpyDeviceToHost, 0));
  if (outlier_flags) PBA_CUDA_OK(cudaMemcpyAsync(outlier_flags, flags.p, sizeof(uint32_t) * ns, cudaMemcpyDeviceToHost, 0));
  if (landmark_remove) PBA_CUDA_OK(cudaMemcpyAsync(landmark_remove, remove.p, size_t(v.n_lm), cudaMemcpyDeviceToHost, 0));
  unsigned sev = 0;
  PBA_CUDA_OK(cudaMemcpyAsync(&sev, severe.p, sizeof(unsigned), cudaMemcpyDeviceToHost, 0));
  PBA_CUDA_OK(cudaStreamSynchronize(0));
  if (any_severe_outliers) *any_severe_outliers = sev ? 1 : 0;
  return PBA_OK;
}
