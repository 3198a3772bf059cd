// visnav_b200/bundle_adjustment.h — host-side drop-in for the reference's
//
//   void visnav::bundle_adjustment(const Corners& feature_corners,
//                                  const BundleAdjustmentOptions& options,
//                                  const std::set<FrameCamId>& fixed_cameras,
//                                  Calibration& calib_cam, Cameras& cameras,
//                                  Landmarks& landmarks)
//   (reference: include/visnav/map_utils.h:322-399, called from src/sfm.cpp:1912)
//
// Header-only C++14 shim over the C ABI (include/pba.h, libpba_b200.so).  It is a
// template over the reference's container types, so it compiles unchanged against
// the reference's own include/visnav/common_types.h / calibration.h (that is how
// oracle/ref/dropin_harness.cpp uses it) and needs neither Eigen, Sophus nor Ceres
// itself.  Same semantics as the reference (map_utils.h:327-392):
//   * one pose block per camera in `cameras`, constant for `fixed_cameras`;
//   * intrinsics constant (optimize_intrinsics = true is rejected: the reference
//     marks it as not working, map_utils.h:339);
//   * per landmark the host is obs.begin() (smallest FrameCamId), one residual
//     block per further observation, Huber loss per block; outlier_obs ignored;
//   * both cameras of a block use the HOST camera's model name with the target's
//     intrinsic values (reprojection.h:97-100, map_utils.h:363-364);
//   * poses (T_w_c) and inverse distances are updated in place; on solver FAILURE
//     the inputs are left untouched (solver.cc:438-447);
//   * unknown camera model names abort, like AbstractCamera::from_data
//     (camera_models.h:469-473);
//   * verbosity 1 / 2 print a one-line / full report (map_utils.h:384-392).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "pba.h"

namespace visnav_b200 {

inline int camera_model_id(const std::string& name) {
  if (name == "pinhole") return PBA_CAM_PINHOLE;
  if (name == "ds") return PBA_CAM_DS;
  if (name == "kb4") return PBA_CAM_KB4;
  if (name == "eucm") return PBA_CAM_EUCM;
  std::fprintf(stderr, "Camera model %s is not implemented.\n", name.c_str());
  std::abort();
}

// Optional photometric inputs (the reference keeps the images in the sfm.cpp global
// `images`, src/sfm.cpp:120-121).  image(fcid) must return a pointer
// to the 8-bit grey image of that camera; affine holds (a, b) per camera in `cameras`
// iteration order and is updated in place.
struct PhotometricInputs {
  std::vector<const uint8_t*> images;  // in `cameras` iteration order
  int width = 0, height = 0, pitch = 0;
  std::vector<double>* affine = nullptr;  // [2 * cameras.size()], may be null (zeros)
};

template <class CornersT, class OptionsT, class FixedSetT, class CalibrationT, class CamerasT, class LandmarksT>
pba_status bundle_adjustment(const CornersT& feature_corners, const OptionsT& options, const FixedSetT& fixed_cameras,
                             CalibrationT& calib_cam, CamerasT& cameras, LandmarksT& landmarks,
                             pba_summary* summary = nullptr, const PhotometricInputs* photo = nullptr,
                             const pba_options* engine_options = nullptr) {
  using FrameCamIdT = typename CamerasT::key_type;
  // ---- flatten the containers into the SoA problem of include/pba.h ----
  std::map<FrameCamIdT, int> pose_index;
  std::vector<double> poses;
  std::vector<uint8_t> pose_fixed;
  std::vector<int32_t> pose_calib;
  poses.reserve(cameras.size() * 7);
  for (auto& kv : cameras) {
    pose_index[kv.first] = int(pose_fixed.size());
    const double* T = kv.second.T_w_c.data();  // Sophus layout qx qy qz qw tx ty tz
    poses.insert(poses.end(), T, T + 7);
    pose_fixed.push_back(fixed_cameras.count(kv.first) > 0);
    pose_calib.push_back(int32_t(kv.first.cam_id));
  }
  const int n_calib = int(calib_cam.intrinsics.size());
  std::vector<int32_t> calib_model(n_calib);
  std::vector<double> intrinsics(size_t(n_calib) * 8);
  for (int i = 0; i < n_calib; ++i) {
    calib_model[i] = camera_model_id(calib_cam.intrinsics[i]->name());
    std::memcpy(&intrinsics[size_t(i) * 8], calib_cam.intrinsics[i]->data(), 8 * sizeof(double));
  }
  std::vector<double> inv_depth, host_uv, obs_uv;
  std::vector<int32_t> lm_host, obs_target;
  std::vector<int64_t> obs_ptr(1, 0);
  std::vector<typename LandmarksT::mapped_type*> lm_ref;
  for (auto& kv : landmarks) {
    auto& lm = kv.second;
    if (lm.obs.empty()) continue;
    const auto host = lm.obs.begin();
    const auto& zh = feature_corners.at(host->first).corners[host->second];
    lm_ref.push_back(&lm);
    inv_depth.push_back(lm.inv_depth);
    lm_host.push_back(pose_index.at(host->first));
    host_uv.push_back(zh[0]);
    host_uv.push_back(zh[1]);
    for (auto it = std::next(lm.obs.begin()); it != lm.obs.end(); ++it) {
      const auto& zt = feature_corners.at(it->first).corners[it->second];
      obs_target.push_back(pose_index.at(it->first));
      obs_uv.push_back(zt[0]);
      obs_uv.push_back(zt[1]);
    }
    obs_ptr.push_back(int64_t(obs_target.size()));
  }

  pba_problem p;
  std::memset(&p, 0, sizeof(p));
  p.mode = photo ? PBA_MODE_PHOTOMETRIC : PBA_MODE_GEOMETRIC;
  p.n_poses = int32_t(pose_fixed.size());
  p.n_calib = n_calib;
  p.n_landmarks = int32_t(inv_depth.size());
  p.n_obs = int64_t(obs_target.size());
  p.poses = poses.data();
  p.pose_fixed = pose_fixed.data();
  p.pose_calib = pose_calib.data();
  p.calib_model = calib_model.data();
  p.intrinsics = intrinsics.data();
  p.inv_depth = inv_depth.data();
  p.lm_host = lm_host.data();
  p.lm_host_uv = host_uv.data();
  p.lm_obs_ptr = obs_ptr.data();
  p.obs_target = obs_target.data();
  p.obs_uv = obs_uv.data();
  // With the host = obs.begin() rule every host index is smaller than its targets;
  // the geometric functor also evaluates the target with the host's model name.
  for (int i = 0; i < n_calib; ++i) (void)i;
  std::vector<double> zero_affine;
  if (photo) {
    p.image_ptrs = photo->images.data();
    p.width = photo->width;
    p.height = photo->height;
    p.pitch = photo->pitch;
    p.obs_uv = nullptr;
    if (photo->affine) {
      p.affine = photo->affine->data();
    } else {
      zero_affine.assign(size_t(p.n_poses) * 2, 0.0);
      p.affine = zero_affine.data();
    }
  }

  pba_options o;
  if (engine_options) o = *engine_options; else pba_options_init(&o);
  o.verbosity_level = options.verbosity_level;
  o.optimize_intrinsics = options.optimize_intrinsics ? 1 : 0;
  o.use_huber = options.use_huber ? 1 : 0;
  o.huber_parameter = options.huber_parameter;
  o.max_num_iterations = options.max_num_iterations;

  const pba_status st = pba_solve(&p, &o, summary);
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::bundle_adjustment: %s\n", pba_status_string(st));
    return st;
  }
  // ---- write back in place, like Ceres does through the raw parameter pointers ----
  int i = 0;
  for (auto& kv : cameras) std::memcpy(kv.second.T_w_c.data(), &poses[size_t(i++) * 7], 7 * sizeof(double));
  for (size_t l = 0; l < lm_ref.size(); ++l) lm_ref[l]->inv_depth = inv_depth[l];
  return PBA_OK;
}

}  // namespace visnav_b200
