// oracle/ref/ref_harness.cpp — C-ABI wrapper around the UNMODIFIED reference.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/ref/Makefile into
// oracle/_ref/libpba_ref.so; loaded only by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.  The product library
// never links or loads it.
//
// What runs here is the reference's own code, included from /root/reference:
//   * geometric BA: visnav::bundle_adjustment() (include/visnav/map_utils.h:322-399)
//     called unchanged on Corners/Cameras/Landmarks/Calibration containers
//     built from the flat pba_problem; and, for per-iteration traces, an
//     identical ceres::Problem assembled the way map_utils.h:327-383 does.
//   * per-block residuals + local Jacobians through the reference functor
//     (include/visnav/reprojection.h:74-118), vendored Ceres 2.0.0 AutoDiff,
//     Sophus::test::LocalParameterizationSE3 and ceres::HuberLoss, evaluated by
//     ceres::Problem::Evaluate (robust correction = internal/ceres/corrector.cc).
//   * photometric BA: the reference snapshot has no photometric functor
//     (SURVEY.md §0).  PhotometricCostFunctor below is the §8(a-P) spec written
//     in the style of BundleAdjustmentReprojectionCostFunctor and differentiated
//     by the same vendored AutoDiff — this is "the reference's Ceres AutoDiff
//     path on the same inputs" that BASELINE.json's north_star names.
#include <visnav/common_types.h>

#include <visnav/calibration.h>
#include <visnav/camera_models.h>
#include <visnav/local_parameterization_se3.hpp>
#include <visnav/map_utils.h>
#include <visnav/reprojection.h>
#include <visnav/aprilgrid.h>

#include <ceres/ceres.h>

#include <algorithm>
#include <bitset>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "pba.h"

namespace {

const char* model_name(int m) {
  switch (m) {
    case PBA_CAM_PINHOLE: return "pinhole";
    case PBA_CAM_DS: return "ds";
    case PBA_CAM_KB4: return "kb4";
    case PBA_CAM_EUCM: return "eucm";
  }
  return "unknown";
}

// DSO 8-pixel residual pattern.
const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0},
                            {0, 0},  {2, 0},   {-1, 1}, {0, 2}};

inline double scalar_part(const double& x) { return x; }
template <int N>
inline double scalar_part(const ceres::Jet<double, N>& x) { return x.a; }

struct Image {
  const uint8_t* ptr;
  int w, h, pitch;
};

// Bilinear interpolation of an 8-bit image, differentiable w.r.t. (u,v):
// the integer cell comes from the scalar part (what ceres::floor on a Jet
// yields, include/ceres/jet.h:516), the fractional weights keep derivatives.
template <class T>
bool bilinear(const Image& im, const T& u, const T& v, T* out) {
  const double us = scalar_part(u), vs = scalar_part(v);
  if (!(us >= 0.0) || !(vs >= 0.0) || !(us < double(im.w - 1)) ||
      !(vs < double(im.h - 1)))
    return false;
  const int x0 = int(std::floor(us)), y0 = int(std::floor(vs));
  const T fx = u - T(double(x0));
  const T fy = v - T(double(y0));
  const uint8_t* p = im.ptr + size_t(y0) * im.pitch + x0;
  const double i00 = p[0], i10 = p[1], i01 = p[im.pitch], i11 = p[im.pitch + 1];
  *out = (T(1.0) - fx) * (T(1.0) - fy) * i00 + fx * (T(1.0) - fy) * i10 +
         (T(1.0) - fx) * fy * i01 + fx * fy * i11;
  return true;
}

// Photometric analogue of BundleAdjustmentReprojectionCostFunctor
// (reprojection.h:74-118): same host-ray / inverse-distance / two-pose
// geometry, residual = I_t(u_k) - (exp(a_t) I_h(z_h+o_k) + b_t), k = 0..7.
struct PhotometricCostFunctor {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  PhotometricCostFunctor(const Eigen::Vector2d& p_2d_ref,
                         const double* host_intensity, bool host_valid,
                         Image target, double* ref_intrinsics,
                         const std::string& ref_model, double* tgt_intrinsics,
                         const std::string& tgt_model)
      : p_2d_ref(p_2d_ref), host_valid(host_valid), target(target),
        ref_intrinsics(ref_intrinsics), ref_model(ref_model),
        tgt_intrinsics(tgt_intrinsics), tgt_model(tgt_model) {
    for (int k = 0; k < 8; ++k) I_h[k] = host_intensity[k];
  }

  template <class T>
  bool operator()(T const* const sT_w_c1, T const* const sT_w_c2,
                  T const* const sAffine, T const* const inv_depth,
                  T* sResiduals) const {
    Eigen::Map<Sophus::SE3<T> const> const T_w_c1(sT_w_c1);
    Eigen::Map<Sophus::SE3<T> const> const T_w_c2(sT_w_c2);
    T ref_intr_[8], tgt_intr_[8];
    for (int i = 0; i < 8; i++) {
      ref_intr_[i] = T(ref_intrinsics[i]);
      tgt_intr_[i] = T(tgt_intrinsics[i]);
    }
    const std::shared_ptr<visnav::AbstractCamera<T>> cam1 =
        visnav::AbstractCamera<T>::from_data(ref_model, ref_intr_);
    const std::shared_ptr<visnav::AbstractCamera<T>> cam2 =
        visnav::AbstractCamera<T>::from_data(tgt_model, tgt_intr_);

    bool ok = host_valid;
    const T ea = exp(sAffine[0]);
    for (int k = 0; k < 8 && ok; ++k) {
      Eigen::Matrix<T, 2, 1> p;
      p[0] = T(p_2d_ref[0] + kPattern[k][0]);
      p[1] = T(p_2d_ref[1] + kPattern[k][1]);
      Eigen::Matrix<T, 3, 1> bearing = cam1->unproject(p);
      bearing.normalize();
      const Eigen::Matrix<T, 2, 1> uv =
          cam2->project(T_w_c2.inverse() * T_w_c1 * (bearing / inv_depth[0]));
      T I_t;
      ok = bilinear(target, uv[0], uv[1], &I_t);
      if (ok) sResiduals[k] = I_t - (ea * I_h[k] + sAffine[1]);
    }
    if (!ok) {
      for (int k = 0; k < 8; ++k) sResiduals[k] = T(0.0);
    }
    return true;
  }

  Eigen::Vector2d p_2d_ref;
  double I_h[8];
  bool host_valid;
  Image target;
  double* ref_intrinsics;
  std::string ref_model;
  double* tgt_intrinsics;
  std::string tgt_model;
};

struct Built {
  std::vector<std::array<double, 7>> poses;
  std::vector<std::array<double, 2>> affine;
  std::vector<std::array<double, 8>> intr;
  std::vector<double> rho;
  std::vector<ceres::ResidualBlockId> blocks;
  std::unique_ptr<ceres::Problem> problem;
};

Image image_of(const pba_problem* p, int pose) {
  Image im;
  im.ptr = p->image_ptrs ? p->image_ptrs[pose]
                         : p->images + size_t(pose) * p->image_stride;
  im.w = p->width;
  im.h = p->height;
  im.pitch = p->pitch;
  return im;
}

// Assemble the ceres::Problem exactly the way map_utils.h:327-375 does
// (parameter blocks, LocalParameterizationSE3, constant blocks, one residual
// block per (landmark, non-host observation), HuberLoss per block).
void build_problem(const pba_problem* p, bool use_huber, double huber,
                   bool honour_fixed, Built* b) {
  b->poses.resize(p->n_poses);
  b->affine.resize(p->n_poses);
  b->intr.resize(p->n_calib);
  b->rho.assign(p->inv_depth, p->inv_depth + p->n_landmarks);
  for (int i = 0; i < p->n_poses; ++i) {
    std::memcpy(b->poses[i].data(), p->poses + 7 * i, 7 * sizeof(double));
    if (p->affine) std::memcpy(b->affine[i].data(), p->affine + 2 * i, 16);
  }
  for (int i = 0; i < p->n_calib; ++i)
    std::memcpy(b->intr[i].data(), p->intrinsics + 8 * i, 64);

  b->problem.reset(new ceres::Problem);
  ceres::Problem& problem = *b->problem;
  const bool photo = p->mode == PBA_MODE_PHOTOMETRIC;
  for (int i = 0; i < p->n_poses; ++i) {
    problem.AddParameterBlock(b->poses[i].data(), 7,
                              new Sophus::test::LocalParameterizationSE3);
    if (photo) problem.AddParameterBlock(b->affine[i].data(), 2);
    if (honour_fixed && p->pose_fixed && p->pose_fixed[i]) {
      problem.SetParameterBlockConstant(b->poses[i].data());
      if (photo) problem.SetParameterBlockConstant(b->affine[i].data());
    }
  }
  if (!photo) {
    for (int i = 0; i < p->n_calib; ++i) {
      problem.AddParameterBlock(b->intr[i].data(), 8);
      problem.SetParameterBlockConstant(b->intr[i].data());
    }
  }
  b->blocks.reserve(p->n_obs);
  for (int l = 0; l < p->n_landmarks; ++l) {
    const int h = p->lm_host[l];
    const int hc = p->pose_calib[h];
    const Eigen::Vector2d zh(p->lm_host_uv[2 * l], p->lm_host_uv[2 * l + 1]);
    double I_h[8] = {0};
    bool host_valid = true;
    if (photo) {
      const Image him = image_of(p, h);
      for (int k = 0; k < 8 && host_valid; ++k)
        host_valid = bilinear<double>(him, zh[0] + kPattern[k][0],
                                      zh[1] + kPattern[k][1], &I_h[k]);
    }
    for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) {
      const int t = p->obs_target[o];
      const int tc = p->pose_calib[t];
      ceres::LossFunction* loss = use_huber ? new ceres::HuberLoss(huber) : nullptr;
      if (!photo) {
        const Eigen::Vector2d zt(p->obs_uv[2 * o], p->obs_uv[2 * o + 1]);
        auto* functor = new visnav::BundleAdjustmentReprojectionCostFunctor(
            zt, zh, b->intr[hc].data(), model_name(p->calib_model[hc]));
        // NB: the reference builds BOTH cameras from the host's model name
        // (reprojection.h:97-100, map_utils.h:363-364).
        auto* cost = new ceres::AutoDiffCostFunction<
            visnav::BundleAdjustmentReprojectionCostFunctor, 2, 7, 7, 1, 8>(functor);
        b->blocks.push_back(problem.AddResidualBlock(
            cost, loss, b->poses[h].data(), b->poses[t].data(), &b->rho[l],
            b->intr[tc].data()));
      } else {
        auto* functor = new PhotometricCostFunctor(
            zh, I_h, host_valid, image_of(p, t), b->intr[hc].data(),
            model_name(p->calib_model[hc]), b->intr[tc].data(),
            model_name(p->calib_model[tc]));
        auto* cost = new ceres::AutoDiffCostFunction<PhotometricCostFunctor, 8,
                                                     7, 7, 2, 1>(functor);
        b->blocks.push_back(problem.AddResidualBlock(
            cost, loss, b->poses[h].data(), b->poses[t].data(),
            b->affine[t].data(), &b->rho[l]));
      }
    }
  }
}

void fill_summary(const ceres::Solver::Summary& s, pba_summary* out) {
  if (!out) return;
  pba_iteration* it = out->iterations;
  const int cap = out->iterations_capacity;
  out->termination_type =
      s.termination_type == ceres::CONVERGENCE
          ? PBA_CONVERGENCE
          : (s.termination_type == ceres::NO_CONVERGENCE ? PBA_NO_CONVERGENCE
                                                         : PBA_FAILURE);
  out->num_iterations = int(s.iterations.size());
  out->num_successful_steps = s.num_successful_steps;
  out->num_unsuccessful_steps = s.num_unsuccessful_steps;
  out->num_residual_evaluations = s.num_residual_evaluations;
  out->num_jacobian_evaluations = s.num_jacobian_evaluations;
  out->num_linear_solves = s.num_linear_solves;
  out->num_residual_blocks = s.num_residual_blocks_reduced;
  out->num_residuals = s.num_residuals_reduced;
  out->num_effective_parameters = s.num_effective_parameters_reduced;
  out->gpu_kernel_launches = 0;
  out->initial_cost = s.initial_cost;
  out->final_cost = s.final_cost;
  out->setup_time_in_seconds = s.preprocessor_time_in_seconds;
  out->residual_evaluation_time_in_seconds = s.residual_evaluation_time_in_seconds;
  out->jacobian_evaluation_time_in_seconds = s.jacobian_evaluation_time_in_seconds;
  out->linear_solver_time_in_seconds = s.linear_solver_time_in_seconds;
  out->minimizer_time_in_seconds = s.minimizer_time_in_seconds;
  out->total_time_in_seconds = s.total_time_in_seconds;
  std::snprintf(out->message, sizeof(out->message), "%s", s.message.c_str());
  for (int i = 0; i < int(s.iterations.size()) && i < cap && it; ++i) {
    const ceres::IterationSummary& a = s.iterations[i];
    it[i].iteration = a.iteration;
    it[i].step_is_valid = a.step_is_valid;
    it[i].step_is_successful = a.step_is_successful;
    it[i].linear_solver_iterations = a.linear_solver_iterations;
    it[i].cost = a.cost;
    it[i].cost_change = a.cost_change;
    it[i].gradient_max_norm = a.gradient_max_norm;
    it[i].gradient_norm = a.gradient_norm;
    it[i].step_norm = a.step_norm;
    it[i].relative_decrease = a.relative_decrease;
    it[i].trust_region_radius = a.trust_region_radius;
    it[i].iteration_time_in_seconds = a.iteration_time_in_seconds;
    it[i].cumulative_time_in_seconds = a.cumulative_time_in_seconds;
    it[i].model_cost_change = 0.0;
  }
}

// The reference's own map containers (common_types.h:160-230) filled from the
// flat problem; FrameCamId(frame = pose index, cam = calibration index).
struct RefMap {
  visnav::Corners corners;
  visnav::Cameras cameras;
  visnav::Landmarks landmarks;
  visnav::Calibration calib;
  std::set<visnav::FrameCamId> fixed;
  std::vector<visnav::FrameCamId> fcid;
};

int build_map(const pba_problem* p, RefMap* m) {
  using namespace visnav;
  int max_cam = 0;
  for (int i = 0; i < p->n_poses; ++i) max_cam = std::max(max_cam, p->pose_calib[i]);
  if (max_cam >= p->n_calib) return 2;
  for (int i = 0; i < p->n_calib; ++i) {
    m->calib.intrinsics.push_back(AbstractCamera<double>::from_data(
        model_name(p->calib_model[i]), p->intrinsics + 8 * i));
    m->calib.T_i_c.push_back(Sophus::SE3d());
  }
  m->fcid.resize(p->n_poses);
  for (int i = 0; i < p->n_poses; ++i) {
    m->fcid[i] = FrameCamId(i, size_t(p->pose_calib[i]));
    Camera cam;
    std::memcpy(cam.T_w_c.data(), p->poses + 7 * i, 7 * sizeof(double));
    m->cameras[m->fcid[i]] = cam;
    m->corners[m->fcid[i]];
    if (p->pose_fixed && p->pose_fixed[i]) m->fixed.insert(m->fcid[i]);
  }
  for (int l = 0; l < p->n_landmarks; ++l) {
    Landmark lm;
    lm.inv_depth = p->inv_depth[l];
    const int h = p->lm_host[l];
    auto& hc = m->corners[m->fcid[h]].corners;
    lm.obs[m->fcid[h]] = FeatureId(hc.size());
    hc.emplace_back(p->lm_host_uv[2 * l], p->lm_host_uv[2 * l + 1]);
    for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) {
      const int t = p->obs_target[o];
      if (!(m->fcid[h] < m->fcid[t])) return 3;  // host must be obs.begin()
      auto& tc = m->corners[m->fcid[t]].corners;
      lm.obs[m->fcid[t]] = FeatureId(tc.size());
      tc.emplace_back(p->obs_uv[2 * o], p->obs_uv[2 * o + 1]);
    }
    m->landmarks[TrackId(l)] = lm;
  }
  return 0;
}

}  // namespace

#define REF_API __attribute__((visibility("default")))

extern "C" {

REF_API int pba_ref_hardware_threads() { return int(std::thread::hardware_concurrency()); }

// Per-block robustified residuals [n_obs*R] and local Jacobians [n_obs*R*C]
// (row-major per block; columns host pose 6 | target pose 6 | [affine 2] | rho)
// in the caller's observation order, via ceres::Problem::Evaluate.  All poses
// are treated as variable so every block has all its columns.
REF_API int pba_ref_eval(const pba_problem* p, int use_huber, double huber,
                 int num_threads, double* residuals, double* jacobians,
                 double* cost) {
  Built b;
  build_problem(p, use_huber != 0, huber, /*honour_fixed=*/false, &b);
  const bool photo = p->mode == PBA_MODE_PHOTOMETRIC;
  const int R = photo ? 8 : 2, C = photo ? 15 : 13;

  ceres::Problem::EvaluateOptions eo;
  eo.apply_loss_function = true;
  eo.num_threads = num_threads > 0 ? num_threads : int(std::thread::hardware_concurrency());
  eo.residual_blocks = b.blocks;
  // column layout: poses (6 each), then affines (2 each), then rho (1 each)
  for (int i = 0; i < p->n_poses; ++i) eo.parameter_blocks.push_back(b.poses[i].data());
  if (photo)
    for (int i = 0; i < p->n_poses; ++i) eo.parameter_blocks.push_back(b.affine[i].data());
  for (int l = 0; l < p->n_landmarks; ++l) eo.parameter_blocks.push_back(&b.rho[l]);
  const int64_t aff0 = 6LL * p->n_poses;
  const int64_t rho0 = aff0 + (photo ? 2LL * p->n_poses : 0);

  double c = 0;
  std::vector<double> res;
  ceres::CRSMatrix J;
  if (!b.problem->Evaluate(eo, &c, &res, nullptr, jacobians ? &J : nullptr)) return 1;
  if (cost) *cost = c;
  if (residuals) std::memcpy(residuals, res.data(), res.size() * sizeof(double));
  if (jacobians) {
    std::memset(jacobians, 0, sizeof(double) * size_t(p->n_obs) * R * C);
    int64_t o = 0;
    for (int l = 0; l < p->n_landmarks; ++l) {
      const int h = p->lm_host[l];
      for (int64_t oo = p->lm_obs_ptr[l]; oo < p->lm_obs_ptr[l + 1]; ++oo, ++o) {
        const int t = p->obs_target[oo];
        for (int r = 0; r < R; ++r) {
          const int64_t row = o * R + r;
          double* out = jacobians + row * C;
          for (int k = J.rows[row]; k < J.rows[row + 1]; ++k) {
            const int64_t col = J.cols[k];
            const double v = J.values[k];
            if (col < aff0) {
              const int pose = int(col / 6), comp = int(col % 6);
              if (pose == h) out[comp] += v;
              else if (pose == t) out[6 + comp] += v;
            } else if (col < rho0) {
              out[12 + (col - aff0) % 2] += v;
            } else {
              out[C - 1] += v;
            }
          }
        }
      }
    }
  }
  return 0;
}

// Full solve.  use_reference_entry=1 (geometric only): build the reference's
// containers and call visnav::bundle_adjustment() UNCHANGED; the summary then
// only carries wall time (the reference returns void).  use_reference_entry=0:
// harness-built Problem identical to map_utils.h:327-383 with the full
// Solver::Summary captured.  Poses / inverse distances / affine are written
// back into the pba_problem arrays, as Ceres does in place.
REF_API int pba_ref_solve(pba_problem* p, const pba_options* opt, int num_threads,
                  int use_reference_entry, pba_summary* summary) {
  const bool photo = p->mode == PBA_MODE_PHOTOMETRIC;
  if (use_reference_entry && !photo) {
    using namespace visnav;
    RefMap m;
    if (int rc = build_map(p, &m)) return rc;
    Corners& corners = m.corners;
    Cameras& cameras = m.cameras;
    Landmarks& landmarks = m.landmarks;
    Calibration& calib = m.calib;
    std::set<FrameCamId>& fixed = m.fixed;
    std::vector<FrameCamId>& fcid = m.fcid;
    BundleAdjustmentOptions ba;
    ba.verbosity_level = opt->verbosity_level;
    ba.optimize_intrinsics = opt->optimize_intrinsics != 0;
    ba.use_huber = opt->use_huber != 0;
    ba.huber_parameter = opt->huber_parameter;
    ba.max_num_iterations = opt->max_num_iterations;
    const auto t0 = std::chrono::steady_clock::now();
    bundle_adjustment(corners, ba, fixed, calib, cameras, landmarks);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < p->n_poses; ++i)
      std::memcpy(p->poses + 7 * i, cameras.at(fcid[i]).T_w_c.data(), 7 * sizeof(double));
    for (int l = 0; l < p->n_landmarks; ++l) p->inv_depth[l] = landmarks.at(TrackId(l)).inv_depth;
    if (summary) {
      pba_iteration* it = summary->iterations;
      const int cap = summary->iterations_capacity;
      std::memset(summary, 0, sizeof(*summary));
      summary->iterations = it;
      summary->iterations_capacity = cap;
      summary->total_time_in_seconds = dt;
      summary->termination_type = -1;  // not observable through the void API
    }
    return 0;
  }

  Built b;
  build_problem(p, opt->use_huber != 0, opt->huber_parameter, /*honour_fixed=*/true, &b);
  // map_utils.h:378-383
  ceres::Solver::Options ceres_options;
  ceres_options.max_num_iterations = opt->max_num_iterations;
  ceres_options.linear_solver_type = ceres::SPARSE_SCHUR;
  ceres_options.num_threads =
      num_threads > 0 ? num_threads : int(std::thread::hardware_concurrency());
  ceres::Solver::Summary s;
  ceres::Solve(ceres_options, b.problem.get(), &s);
  if (opt->verbosity_level == 1) std::printf("%s\n", s.BriefReport().c_str());
  if (opt->verbosity_level == 2) std::printf("%s\n", s.FullReport().c_str());
  for (int i = 0; i < p->n_poses; ++i) {
    std::memcpy(p->poses + 7 * i, b.poses[i].data(), 7 * sizeof(double));
    if (photo && p->affine) std::memcpy(p->affine + 2 * i, b.affine[i].data(), 16);
  }
  std::memcpy(p->inv_depth, b.rho.data(), sizeof(double) * p->n_landmarks);
  fill_summary(s, summary);
  return 0;
}

// Camera models of the reference (camera_models.h) on double, for pinning the
// oracle's / kernels' project & unproject.
REF_API int pba_ref_project(int model, const double* intr, int64_t n, const double* xyz, double* uv) {
  auto cam = visnav::AbstractCamera<double>::from_data(model_name(model), intr);
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector2d r = cam->project(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    uv[2 * i] = r[0];
    uv[2 * i + 1] = r[1];
  }
  return 0;
}

// d(uv)/d(xyz) by the vendored AutoDiff Jet (SURVEY.md §8(a) J_pi check).
REF_API int pba_ref_project_jacobian(int model, const double* intr, int64_t n, const double* xyz, double* J) {
  using JetT = ceres::Jet<double, 3>;
  JetT ji[8];
  for (int i = 0; i < 8; ++i) ji[i] = JetT(intr[i]);
  auto cam = visnav::AbstractCamera<JetT>::from_data(model_name(model), ji);
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Matrix<JetT, 3, 1> X;
    for (int k = 0; k < 3; ++k) X[k] = JetT(xyz[3 * i + k], k);
    Eigen::Matrix<JetT, 2, 1> r = cam->project(X);
    for (int a = 0; a < 2; ++a)
      for (int k = 0; k < 3; ++k) J[6 * i + 3 * a + k] = r[a].v[k];
  }
  return 0;
}

REF_API int pba_ref_unproject(int model, const double* intr, int64_t n, const double* uv, double* xyz) {
  auto cam = visnav::AbstractCamera<double>::from_data(model_name(model), intr);
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector3d r = cam->unproject(Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]));
    xyz[3 * i] = r[0];
    xyz[3 * i + 1] = r[1];
    xyz[3 * i + 2] = r[2];
  }
  return 0;
}

// LocalParameterizationSE3::Plus / ComputeJacobian (local_parameterization_se3.hpp:44-64).
REF_API int pba_ref_se3_plus(int64_t n, const double* poses7, const double* delta6, double* out7) {
  Sophus::test::LocalParameterizationSE3 lp;
  for (int64_t i = 0; i < n; ++i) lp.Plus(poses7 + 7 * i, delta6 + 6 * i, out7 + 7 * i);
  return 0;
}
REF_API int pba_ref_se3_plus_jacobian(const double* pose7, double* J42) {
  Sophus::test::LocalParameterizationSE3 lp;
  lp.ComputeJacobian(pose7, J42);
  return 0;
}


// SURVEY.md §8(f)-2/3.  Landmark::get_p (common_types.h:205-217), SE3::inverse
// and AbstractCamera::project are the reference's own code, called on its own
// containers.  compute_projections / set_outlier_flags themselves live in
// src/sfm.cpp (GUI translation unit with pangolin::Var globals — unbuildable
// here), so the loop of src/sfm.cpp:1960-1984 and the four tests of
// src/sfm.cpp:1928-1952 are repeated around those calls.
REF_API int pba_ref_landmark_positions(const pba_problem* p, double* p_w) {
  RefMap m;
  if (int rc = build_map(p, &m)) return rc;
  for (int l = 0; l < p->n_landmarks; ++l) {
    const Eigen::Vector3d w =
        m.landmarks.at(visnav::TrackId(l)).get_p(m.cameras, m.calib, m.corners);
    std::memcpy(p_w + 3 * l, w.data(), 3 * sizeof(double));
  }
  return 0;
}

REF_API int pba_ref_compute_projections(const pba_problem* p,
                                        const pba_projection_thresholds* thr,
                                        double* point_reprojected, double* point_3d_c,
                                        double* reprojection_error,
                                        uint32_t* outlier_flags) {
  using namespace visnav;
  RefMap m;
  if (int rc = build_map(p, &m)) return rc;
  for (int l = 0; l < p->n_landmarks; ++l) {
    const Landmark& lm = m.landmarks.at(TrackId(l));
    const int64_t base = p->lm_obs_ptr[l];
    const int64_t cnt = p->lm_obs_ptr[l + 1] - base + 1;
    for (int64_t k = 0; k < cnt; ++k) {
      const int64_t s = base + l + k;
      const int pose = k == 0 ? p->lm_host[l] : p->obs_target[base + k - 1];
      const FrameCamId& fcid = m.fcid[pose];
      const Eigen::Vector2d p_2d_corner =
          m.corners.at(fcid).corners[lm.obs.at(fcid)];
      const Eigen::Vector3d p_c = m.cameras.at(fcid).T_w_c.inverse() *
                                  lm.get_p(m.cameras, m.calib, m.corners);
      const Eigen::Vector2d p_2d_repoj =
          m.calib.intrinsics.at(fcid.cam_id)->project(p_c);
      const double e = (p_2d_corner - p_2d_repoj).norm();
      uint32_t f = OutlierNone;
      if (e > thr->reprojection_error_huge_pixel) f |= OutlierReprojectionErrorHuge;
      if (e > thr->reprojection_error_normal_pixel) f |= OutlierReprojectionErrorNormal;
      if (p_c.norm() < thr->camera_center_distance_meter) f |= OutlierCameraDistance;
      if (p_c.z() < thr->z_coordinate_meter) f |= OutlierZCoordinate;
      if (point_reprojected) std::memcpy(point_reprojected + 2 * s, p_2d_repoj.data(), 16);
      if (point_3d_c) std::memcpy(point_3d_c + 3 * s, p_c.data(), 24);
      if (reprojection_error) reprojection_error[s] = e;
      if (outlier_flags) outlier_flags[s] = f;
    }
  }
  return 0;
}

// the map archive.  save_map_file / load_map_file are the reference's own
// functions (include/visnav/map_utils.h:58-116, cereal binary); the containers come from the flat
// problem plus deterministic filler for the fields bundle adjustment does not touch (corner
// angles, descriptors, matches, feature tracks, one outlier track / outlier observation).
REF_API int pba_ref_save_map(const pba_problem* p, const char* path) {
  using namespace visnav;
  RefMap m;
  if (int rc = build_map(p, &m)) return rc;
  Matches matches;
  FeatureTracks tracks, outlier_tracks;
  for (auto& kv : m.corners) {
    KeypointsData& kd = kv.second;
    const size_t n = kd.corners.size();
    kd.corner_angles.resize(n);
    kd.corner_descriptors.resize(n);
    for (size_t i = 0; i < n; ++i) {
      kd.corner_angles[i] = 0.001 * double(i) - 0.5 * double(kv.first.frame_id);
      std::bitset<256> d;
      for (int b = 0; b < 256; ++b) d[b] = ((i * 2654435761u + size_t(kv.first.frame_id) * 40503u + b * 7919u) >> 7) & 1u;
      kd.corner_descriptors[i] = d;
    }
  }
  for (int i = 0; i + 1 < p->n_poses; ++i) {
    MatchData md;
    md.T_i_j = m.cameras.at(m.fcid[i]).T_w_c.inverse() * m.cameras.at(m.fcid[i + 1]).T_w_c;
    const int n = int(std::min(m.corners.at(m.fcid[i]).corners.size(), m.corners.at(m.fcid[i + 1]).corners.size()));
    for (int k = 0; k < std::min(n, 7); ++k) md.matches.emplace_back(k, n - 1 - k);
    for (int k = 0; k < std::min(n, 4); ++k) md.inliers.emplace_back(k, n - 1 - k);
    matches[std::make_pair(m.fcid[i], m.fcid[i + 1])] = md;
  }
  for (auto& kv : m.landmarks) tracks[kv.first] = kv.second.obs;
  if (!m.landmarks.empty()) {
    auto best = m.landmarks.begin();
    for (auto it = m.landmarks.begin(); it != m.landmarks.end(); ++it)
      if (it->second.obs.size() > best->second.obs.size()) best = it;
    Landmark& first = best->second;  // the landmark with the most observations
    outlier_tracks[TrackId(1000000)] = first.obs;
    if (first.obs.size() > 2) {  // move the last observation to outlier_obs
      auto last = std::prev(first.obs.end());
      first.outlier_obs[last->first] = last->second;
      first.obs.erase(last);
    }
  }
  save_map_file(path, m.corners, matches, tracks, outlier_tracks, m.cameras, m.landmarks);
  return 0;
}

// load_map_file(in) -> save_map_file(out): what the reference understood of a file it did not write
REF_API int pba_ref_map_roundtrip(const char* in_path, const char* out_path, int64_t* counts) {
  using namespace visnav;
  Corners corners;
  Matches matches;
  FeatureTracks tracks, outlier_tracks;
  Cameras cameras;
  Landmarks landmarks;
  load_map_file(in_path, corners, matches, tracks, outlier_tracks, cameras, landmarks);
  if (counts) {
    counts[0] = int64_t(corners.size()); counts[1] = int64_t(matches.size()); counts[2] = int64_t(tracks.size());
    counts[3] = int64_t(outlier_tracks.size()); counts[4] = int64_t(cameras.size()); counts[5] = int64_t(landmarks.size());
  }
  save_map_file(out_path, corners, matches, tracks, outlier_tracks, cameras, landmarks);
  return 0;
}

}  // extern "C"

// add_new_landmarks_between_cams (include/visnav/map_utils.h:121-195) through the reference's OWN function:
// two cameras, n feature tracks shared by them, no landmarks yet -> the reference triangulates every track
// (opengv::triangulation::triangulate, compiled from thirdparty/opengv/src by oracle/ref/Makefile) and sets
// inv_depth = 1 / |p|.  p_c0 (optional) comes from opengv's triangulate called the same way.
extern "C" REF_API int pba_ref_add_new_landmarks(int model0, const double* intr0, int model1, const double* intr1,
                                      const double* T_w_c0, const double* T_w_c1, int64_t n, const double* uv0,
                                      const double* uv1, double* p_c0, double* inv_depth) {
  using namespace visnav;
  Calibration calib;
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(model_name(model0), intr0));
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(model_name(model1), intr1));
  calib.T_i_c.push_back(Sophus::SE3d());
  calib.T_i_c.push_back(Sophus::SE3d());
  const FrameCamId fcid0(0, 0), fcid1(1, 1);
  Cameras cameras;
  Camera c0, c1;
  std::memcpy(c0.T_w_c.data(), T_w_c0, 7 * sizeof(double));
  std::memcpy(c1.T_w_c.data(), T_w_c1, 7 * sizeof(double));
  cameras[fcid0] = c0;
  cameras[fcid1] = c1;
  Corners corners;
  FeatureTracks tracks;
  for (int64_t i = 0; i < n; ++i) {
    corners[fcid0].corners.emplace_back(uv0[2 * i], uv0[2 * i + 1]);
    corners[fcid1].corners.emplace_back(uv1[2 * i], uv1[2 * i + 1]);
    FeatureTrack t;
    t[fcid0] = FeatureId(i);
    t[fcid1] = FeatureId(i);
    tracks[TrackId(i)] = t;
  }
  Landmarks landmarks;
  const int added = add_new_landmarks_between_cams(fcid0, fcid1, calib, corners, tracks, cameras, landmarks);
  if (added != int(n)) return 2;
  for (int64_t i = 0; i < n; ++i) inv_depth[i] = landmarks.at(TrackId(i)).inv_depth;
  if (p_c0) {
    opengv::bearingVectors_t b0, b1;
    for (int64_t i = 0; i < n; ++i) {
      Eigen::Vector3d v0 = calib.intrinsics[0]->unproject(corners[fcid0].corners[i]);
      Eigen::Vector3d v1 = calib.intrinsics[1]->unproject(corners[fcid1].corners[i]);
      b0.push_back(v0.normalized());
      b1.push_back(v1.normalized());
    }
    const Sophus::SE3d T01 = c0.T_w_c.inverse() * c1.T_w_c;
    opengv::relative_pose::CentralRelativeAdapter adapter(b0, b1, T01.translation(), T01.rotationMatrix());
    for (int64_t i = 0; i < n; ++i) {
      const Eigen::Vector3d p = opengv::triangulation::triangulate(adapter, i);
      p_c0[3 * i] = p[0]; p_c0[3 * i + 1] = p[1]; p_c0[3 * i + 2] = p[2];
    }
  }
  return 0;
}

// Camera calibration exactly as the reference's calibration application sets it up (src/calibration.cpp:373-425,
// a GUI translation unit that cannot be compiled here): one ReprojectionCostFunctor (include/visnav/reprojection.h:
// 47-71, the reference's own) per detected AprilGrid corner (include/visnav/aprilgrid.h), parameter blocks T_w_i per
// frame, T_i_c per camera (camera 0's constant) and the intrinsics per camera, LocalParameterizationSE3, the same
// solver options.  It exists so that BASELINE config 1 (data/euroc_V1) can get the stereo extrinsics the snapshot
// only provides through that application (tools/euroc/calibrate.py).  In/out arrays are updated in place.
extern "C" REF_API int pba_ref_calibrate(int model, int n_frames, int n_cams, double* intr /*[n_cams*8]*/,
                                         double* T_i_c /*[n_cams*7]*/, double* T_w_i /*[n_frames*7]*/, int64_t n_obs,
                                         const int32_t* obs_frame, const int32_t* obs_cam, const int32_t* obs_corner,
                                         const double* obs_uv, int num_threads, double* initial_cost, double* final_cost) {
  using namespace visnav;
  const AprilGrid grid;
  const std::string name = model_name(model);
  std::vector<Sophus::SE3d, Eigen::aligned_allocator<Sophus::SE3d>> Twi(n_frames), Tic(n_cams);
  for (int i = 0; i < n_frames; ++i) std::memcpy(Twi[i].data(), T_w_i + 7 * i, 7 * sizeof(double));
  for (int i = 0; i < n_cams; ++i) std::memcpy(Tic[i].data(), T_i_c + 7 * i, 7 * sizeof(double));
  ceres::Problem problem;
  for (int i = 0; i < n_frames; ++i)
    problem.AddParameterBlock(Twi[i].data(), Sophus::SE3d::num_parameters, new Sophus::test::LocalParameterizationSE3);
  for (int i = 0; i < n_cams; ++i) {
    problem.AddParameterBlock(Tic[i].data(), Sophus::SE3d::num_parameters, new Sophus::test::LocalParameterizationSE3);
    if (i == 0) problem.SetParameterBlockConstant(Tic[i].data());
  }
  for (int64_t k = 0; k < n_obs; ++k) {
    if (obs_corner[k] < 0 || size_t(obs_corner[k]) >= grid.aprilgrid_corner_pos_3d.size()) return 2;
    const Eigen::Vector2d p_2d(obs_uv[2 * k], obs_uv[2 * k + 1]);
    const Eigen::Vector3d& p_3d = grid.aprilgrid_corner_pos_3d[obs_corner[k]];
    auto* functor = new ReprojectionCostFunctor(p_2d, p_3d, name);
    auto* cost = new ceres::AutoDiffCostFunction<ReprojectionCostFunctor, 2, Sophus::SE3d::num_parameters,
                                                 Sophus::SE3d::num_parameters, 8>(functor);
    problem.AddResidualBlock(cost, nullptr, Twi[obs_frame[k]].data(), Tic[obs_cam[k]].data(), intr + 8 * obs_cam[k]);
  }
  ceres::Solver::Options options;
  options.gradient_tolerance = 0.01 * Sophus::Constants<double>::epsilon();
  options.function_tolerance = 0.01 * Sophus::Constants<double>::epsilon();
  options.linear_solver_type = ceres::SPARSE_NORMAL_CHOLESKY;
  options.num_threads = num_threads > 0 ? num_threads : int(std::thread::hardware_concurrency());
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  for (int i = 0; i < n_frames; ++i) std::memcpy(T_w_i + 7 * i, Twi[i].data(), 7 * sizeof(double));
  for (int i = 0; i < n_cams; ++i) std::memcpy(T_i_c + 7 * i, Tic[i].data(), 7 * sizeof(double));
  if (initial_cost) *initial_cost = summary.initial_cost;
  if (final_cost) *final_cost = summary.final_cost;
  return summary.IsSolutionUsable() ? 0 : 3;
}

