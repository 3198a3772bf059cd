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
#include <memory>
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

// The reference's map containers flattened into the SoA problem of include/pba.h
// (one instance per call; `problem` points into the vectors held here).
// Flattening rule = map_utils.h:327-375: cameras in `Cameras` iteration order,
// landmarks in `Landmarks` iteration order (those with an empty `obs` skipped),
// host = obs.begin(), then one observation per further `obs` entry.
template <class CornersT, class CalibrationT, class CamerasT, class LandmarksT>
struct FlatMap {
  using FrameCamIdT = typename CamerasT::key_type;
  using TrackIdT = typename LandmarksT::key_type;
  using LandmarkT = typename LandmarksT::mapped_type;

  std::vector<FrameCamIdT> pose_fcid;  // pose index -> FrameCamId
  std::map<FrameCamIdT, int> pose_index;
  std::vector<double> poses, intrinsics, inv_depth, host_uv, obs_uv;
  std::vector<uint8_t> pose_fixed;
  std::vector<int32_t> pose_calib, calib_model, lm_host, obs_target;
  std::vector<int64_t> obs_ptr;
  std::vector<LandmarkT*> lm_ref;     // landmark index -> the reference's Landmark
  std::vector<TrackIdT> lm_track;     // landmark index -> TrackId
  pba_problem problem;

  template <class FixedSetT>
  FlatMap(const CornersT& feature_corners, const FixedSetT& fixed_cameras, CalibrationT& calib_cam, CamerasT& cameras,
          LandmarksT& landmarks)
      : obs_ptr(1, 0) {
    poses.reserve(cameras.size() * 7);
    for (auto& kv : cameras) {
      pose_index[kv.first] = int(pose_fixed.size());
      pose_fcid.push_back(kv.first);
      const double* T = kv.second.T_w_c.data();  // Sophus layout qx qy qz qw tx ty tz
      poses.insert(poses.end(), T, T + 7);
      pose_fixed.push_back(fixed_cameras.count(kv.first) > 0);
      pose_calib.push_back(int32_t(kv.first.cam_id));
    }
    const int n_calib = int(calib_cam.intrinsics.size());
    calib_model.resize(n_calib);
    intrinsics.resize(size_t(n_calib) * 8);
    for (int i = 0; i < n_calib; ++i) {
      calib_model[i] = camera_model_id(calib_cam.intrinsics[i]->name());
      std::memcpy(&intrinsics[size_t(i) * 8], calib_cam.intrinsics[i]->data(), 8 * sizeof(double));
    }
    for (auto& kv : landmarks) {
      auto& lm = kv.second;
      if (lm.obs.empty()) continue;
      const auto host = lm.obs.begin();
      const auto& zh = feature_corners.at(host->first).corners[host->second];
      lm_ref.push_back(&lm);
      lm_track.push_back(kv.first);
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
    std::memset(&problem, 0, sizeof(problem));
    problem.mode = PBA_MODE_GEOMETRIC;
    problem.n_poses = int32_t(pose_fixed.size());
    problem.n_calib = n_calib;
    problem.n_landmarks = int32_t(inv_depth.size());
    problem.n_obs = int64_t(obs_target.size());
    problem.poses = poses.data();
    problem.pose_fixed = pose_fixed.data();
    problem.pose_calib = pose_calib.data();
    problem.calib_model = calib_model.data();
    problem.intrinsics = intrinsics.data();
    problem.inv_depth = inv_depth.data();
    problem.lm_host = lm_host.data();
    problem.lm_host_uv = host_uv.data();
    problem.lm_obs_ptr = obs_ptr.data();
    problem.obs_target = obs_target.data();
    problem.obs_uv = obs_uv.data();
  }
  FlatMap(const FlatMap&) = delete;
  FlatMap& operator=(const FlatMap&) = delete;

  // In-place update of T_w_c / inv_depth, like Ceres through the raw parameter pointers.
  void write_back(CamerasT& cameras) {
    int i = 0;
    for (auto& kv : cameras) std::memcpy(kv.second.T_w_c.data(), &poses[size_t(i++) * 7], 7 * sizeof(double));
    for (size_t l = 0; l < lm_ref.size(); ++l) lm_ref[l]->inv_depth = inv_depth[l];
  }
};

template <class CornersT, class OptionsT, class FixedSetT, class CalibrationT, class CamerasT, class LandmarksT>
pba_status bundle_adjustment(const CornersT& feature_corners, const OptionsT& options, const FixedSetT& fixed_cameras,
                             CalibrationT& calib_cam, CamerasT& cameras, LandmarksT& landmarks,
                             pba_summary* summary = nullptr, const PhotometricInputs* photo = nullptr,
                             const pba_options* engine_options = nullptr) {
  FlatMap<CornersT, CalibrationT, CamerasT, LandmarksT> flat(feature_corners, fixed_cameras, calib_cam, cameras, landmarks);
  pba_problem& p = flat.problem;
  std::vector<double> zero_affine;
  if (photo) {
    p.mode = PBA_MODE_PHOTOMETRIC;
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
  flat.write_back(cameras);
  return PBA_OK;
}

// Drop-in for the inlier half of compute_projections() (src/sfm.cpp:1956-1984)
// including set_outlier_flags() (src/sfm.cpp:1928-1952): fills the reference's
// `ImageProjections` / `TrackProjections` (common_types.h:288-316) from one GPU
// pass.  `ProjectedLandmarkT` is the reference's ProjectedLandmark.  If
// `tracks_to_remove` is given it receives the TrackIds remove_outlier_landmarks()
// (src/sfm.cpp:2039-2091) would erase.  The already-rejected `outlier_obs`
// (src/sfm.cpp:1986-2005, drawing only) stay with the caller.
template <class ProjectedLandmarkT, class CornersT, class CalibrationT, class CamerasT, class LandmarksT,
          class ImageProjectionsT, class TrackProjectionsT>
pba_status compute_projections(const CornersT& feature_corners, CalibrationT& calib_cam, CamerasT& cameras,
                               LandmarksT& landmarks, const pba_projection_thresholds& thresholds,
                               ImageProjectionsT& image_projections, TrackProjectionsT& track_projections,
                               std::vector<typename LandmarksT::key_type>* tracks_to_remove = nullptr,
                               int device = 0) {
  image_projections.clear();
  track_projections.clear();
  if (tracks_to_remove) tracks_to_remove->clear();
  std::map<typename CamerasT::key_type, int> no_fixed;
  FlatMap<CornersT, CalibrationT, CamerasT, LandmarksT> flat(feature_corners, no_fixed, calib_cam, cameras, landmarks);
  const pba_problem& p = flat.problem;
  const size_t ns = size_t(p.n_obs) + size_t(p.n_landmarks);
  std::vector<double> repro(ns * 2), p3c(ns * 3), err(ns);
  std::vector<uint32_t> flags(ns);
  std::vector<uint8_t> remove(size_t(p.n_landmarks));
  const pba_status st = pba_compute_projections(&p, &thresholds, device, repro.data(), p3c.data(), err.data(),
                                                flags.data(), remove.data(), nullptr);
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::compute_projections: %s\n", pba_status_string(st));
    return st;
  }
  for (int l = 0; l < p.n_landmarks; ++l) {
    const int64_t base = p.lm_obs_ptr[l];
    const int64_t cnt = p.lm_obs_ptr[l + 1] - base + 1;
    for (int64_t k = 0; k < cnt; ++k) {
      const size_t s = size_t(base + l + k);
      const int pose = k == 0 ? p.lm_host[l] : p.obs_target[base + k - 1];
      const double* z = k == 0 ? &flat.host_uv[size_t(l) * 2] : &flat.obs_uv[size_t(base + k - 1) * 2];
      std::shared_ptr<ProjectedLandmarkT> proj(new ProjectedLandmarkT);
      proj->track_id = flat.lm_track[l];
      proj->point_measured[0] = z[0]; proj->point_measured[1] = z[1];
      proj->point_reprojected[0] = repro[2 * s]; proj->point_reprojected[1] = repro[2 * s + 1];
      proj->point_3d_c[0] = p3c[3 * s]; proj->point_3d_c[1] = p3c[3 * s + 1]; proj->point_3d_c[2] = p3c[3 * s + 2];
      proj->reprojection_error = err[s];
      proj->outlier_flags = flags[s];
      const auto& fcid = flat.pose_fcid[pose];
      image_projections[fcid].obs.push_back(proj);
      track_projections[flat.lm_track[l]][fcid] = proj;
    }
    if (tracks_to_remove && remove[l]) tracks_to_remove->push_back(flat.lm_track[l]);
  }
  return PBA_OK;
}

// Drop-in for the reference's
//
//   int add_new_landmarks_between_cams(const FrameCamId& fcid0, const FrameCamId& fcid1,
//                                      const Calibration&, const Corners&, const FeatureTracks&,
//                                      const Cameras&, Landmarks&)
//   (include/visnav/map_utils.h:121-195; called by initialize_scene_from_stereo_pair and by the
//   "add landmarks" stage of the SfM loop, right before optimize())
//
// Same semantics: every feature track seen in BOTH images that is not a landmark yet is triangulated in
// camera 0's frame from the two unit bearings (opengv's linear method) and becomes a landmark with inverse
// distance 1 / |p| and the observations of all cameras already in the map.  The bearings + triangulation of
// all new tracks run in one GPU pass (pba_triangulate_inverse_depth).  Returns the number of new landmarks,
// or -1 if the GPU call failed.
template <class FrameCamIdT, class CalibrationT, class CornersT, class FeatureTracksT, class CamerasT, class LandmarksT>
int add_new_landmarks_between_cams(const FrameCamIdT& fcid0, const FrameCamIdT& fcid1, const CalibrationT& calib_cam,
                                   const CornersT& feature_corners, const FeatureTracksT& feature_tracks,
                                   const CamerasT& cameras, LandmarksT& landmarks, int device = 0) {
  using TrackIdT = typename FeatureTracksT::key_type;
  std::vector<TrackIdT> new_track_ids;
  std::vector<double> uv0, uv1;
  const auto& corners0 = feature_corners.at(fcid0).corners;
  const auto& corners1 = feature_corners.at(fcid1).corners;
  for (const auto& kv : feature_tracks) {  // GetTracksInImages (tracks.h:175-197) + "not a landmark yet"
    const auto it0 = kv.second.find(fcid0), it1 = kv.second.find(fcid1);
    if (it0 == kv.second.end() || it1 == kv.second.end() || landmarks.count(kv.first) > 0) continue;
    const auto& z0 = corners0[it0->second];
    const auto& z1 = corners1[it1->second];
    uv0.push_back(z0[0]); uv0.push_back(z0[1]);
    uv1.push_back(z1[0]); uv1.push_back(z1[1]);
    new_track_ids.push_back(kv.first);
  }
  if (new_track_ids.empty()) return 0;
  const auto& cam0 = calib_cam.intrinsics[fcid0.cam_id];
  const auto& cam1 = calib_cam.intrinsics[fcid1.cam_id];
  const int64_t n = int64_t(new_track_ids.size());
  std::vector<double> rho(size_t(n), 0.0);
  const pba_status st = pba_triangulate_inverse_depth(
      camera_model_id(cam0->name()), cam0->data(), camera_model_id(cam1->name()), cam1->data(),
      cameras.at(fcid0).T_w_c.data(), cameras.at(fcid1).T_w_c.data(), n, uv0.data(), uv1.data(), device, nullptr,
      rho.data());
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::add_new_landmarks_between_cams: %s\n", pba_status_string(st));
    return -1;
  }
  for (int64_t i = 0; i < n; ++i) {
    typename LandmarksT::mapped_type lm;
    lm.inv_depth = rho[size_t(i)];
    for (const auto& track_kv : feature_tracks.at(new_track_ids[size_t(i)]))
      if (cameras.count(track_kv.first) > 0) lm.obs[track_kv.first] = track_kv.second;
    landmarks[new_track_ids[size_t(i)]] = lm;
  }
  return int(n);
}

}  // namespace visnav_b200
