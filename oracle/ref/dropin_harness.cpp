// oracle/ref/dropin_harness.cpp — drop-in proof (TEST INFRASTRUCTURE).
//
// Builds the REFERENCE'S OWN containers (include/visnav/common_types.h,
// calibration.h, camera_models.h from /root/reference) from a flat problem and
// runs, in the same process, either
//   * the unmodified reference visnav::bundle_adjustment() (Ceres on the CPU), or
//   * visnav_b200::bundle_adjustment() (include/visnav_b200/bundle_adjustment.h)
//     -> pba_solve -> libpba_b200.so (CUDA engine)
// with the SAME argument list, and writes the optimised state back.  This is the
// call a maintainer would swap in src/sfm.cpp:1912.
#include <visnav/common_types.h>

#include <visnav/calibration.h>
#include <visnav/camera_models.h>
#include <visnav/keypoints.h>
#include <visnav/map_utils.h>
#include <visnav/matching_utils.h>
#include <visnav/tracks.h>

#include <algorithm>
#include <cstring>
#include <set>
#include <vector>

#include "pba.h"
#include "visnav_b200/bundle_adjustment.h"
#include "visnav_b200/frontend.h"

namespace {

struct RefMap {
  visnav::Corners corners;
  visnav::Cameras cameras;
  visnav::Landmarks landmarks;
  visnav::Calibration calib;
  std::set<visnav::FrameCamId> fixed;
  std::vector<visnav::FrameCamId> fcid;
};

// The reference's own containers from the flat problem (FrameCamId = (pose index, calibration index)).
int build_map(const pba_problem* p, RefMap* m) {
  using namespace visnav;
  for (int i = 0; i < p->n_calib; ++i) {
    const char* names[] = {"pinhole", "ds", "kb4", "eucm"};
    m->calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[p->calib_model[i]], p->intrinsics + 8 * i));
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
      if (!(m->fcid[h] < m->fcid[t])) return 11;  // host must be obs.begin()
      auto& tc = m->corners[m->fcid[t]].corners;
      lm.obs[m->fcid[t]] = FeatureId(tc.size());
      tc.emplace_back(p->obs_uv[2 * o], p->obs_uv[2 * o + 1]);
    }
    m->landmarks[TrackId(l)] = lm;
  }
  return 0;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int pba_dropin_solve(pba_problem* p, const pba_options* opt,
                                                                        int use_b200, pba_summary* summary) {
  using namespace visnav;
  if (p->mode != PBA_MODE_GEOMETRIC) return 10;
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

  int rc = 0;
  if (use_b200) {
    // the drop-in: identical argument list, CUDA engine behind it
    rc = int(visnav_b200::bundle_adjustment(corners, ba, fixed, calib, cameras, landmarks, summary));
  } else {
    bundle_adjustment(corners, ba, fixed, calib, cameras, landmarks);
  }
  for (int i = 0; i < p->n_poses; ++i)
    std::memcpy(p->poses + 7 * i, cameras.at(fcid[i]).T_w_c.data(), 7 * sizeof(double));
  for (int l = 0; l < p->n_landmarks; ++l) p->inv_depth[l] = landmarks.at(TrackId(l)).inv_depth;
  return rc;
}


// compute_projections() drop-in proof.  Fills the
// reference's ImageProjections / TrackProjections either
//   * use_b200 = 0: the way src/sfm.cpp:1960-1984 + :1928-1952 does, with the
//     reference's own Landmark::get_p / SE3::inverse / project (sfm.cpp itself
//     is a GUI translation unit and cannot be compiled here), or
//   * use_b200 = 1: visnav_b200::compute_projections -> pba_compute_projections
//     (CUDA),
// then reads BOTH back the same way: per track in std::map<FrameCamId> order
// (= slot order when every landmark's targets are ascending).  n_image_obs
// receives image_projections[fcid].obs.size() per pose.
extern "C" __attribute__((visibility("default"))) int pba_dropin_compute_projections(
    const pba_problem* p, const pba_projection_thresholds* thr, int use_b200, double* point_measured,
    double* point_reprojected, double* point_3d_c, double* reprojection_error, uint32_t* outlier_flags,
    int64_t* n_image_obs, uint8_t* landmark_remove) {
  using namespace visnav;
  RefMap m;
  if (int rc = build_map(p, &m)) return rc;
  ImageProjections image_projections;
  TrackProjections track_projections;
  if (use_b200) {
    std::vector<TrackId> gone;
    const pba_status st = visnav_b200::compute_projections<ProjectedLandmark>(
        m.corners, m.calib, m.cameras, m.landmarks, *thr, image_projections, track_projections, &gone);
    if (st != PBA_OK) return int(st);
    if (landmark_remove) {
      std::memset(landmark_remove, 0, size_t(p->n_landmarks));
      for (TrackId t : gone) landmark_remove[t] = 1;
    }
  } else {
    for (const auto& kv_lm : m.landmarks) {
      for (const auto& kv_obs : kv_lm.second.obs) {
        const FrameCamId& fcid = kv_obs.first;
        const Eigen::Vector2d p_2d_corner = m.corners.at(fcid).corners[kv_obs.second];
        const Eigen::Vector3d p_c =
            m.cameras.at(fcid).T_w_c.inverse() * kv_lm.second.get_p(m.cameras, m.calib, m.corners);
        const Eigen::Vector2d p_2d_repoj = m.calib.intrinsics.at(fcid.cam_id)->project(p_c);
        ProjectedLandmarkPtr proj_lm(new ProjectedLandmark);
        proj_lm->track_id = kv_lm.first;
        proj_lm->point_measured = p_2d_corner;
        proj_lm->point_reprojected = p_2d_repoj;
        proj_lm->point_3d_c = p_c;
        proj_lm->reprojection_error = (p_2d_corner - p_2d_repoj).norm();
        if (proj_lm->reprojection_error > thr->reprojection_error_huge_pixel)
          proj_lm->outlier_flags |= OutlierReprojectionErrorHuge;
        if (proj_lm->reprojection_error > thr->reprojection_error_normal_pixel)
          proj_lm->outlier_flags |= OutlierReprojectionErrorNormal;
        if (proj_lm->point_3d_c.norm() < thr->camera_center_distance_meter)
          proj_lm->outlier_flags |= OutlierCameraDistance;
        if (proj_lm->point_3d_c.z() < thr->z_coordinate_meter) proj_lm->outlier_flags |= OutlierZCoordinate;
        image_projections[fcid].obs.push_back(proj_lm);
        track_projections[kv_lm.first][fcid] = proj_lm;
      }
    }
    if (landmark_remove) std::memset(landmark_remove, 0, size_t(p->n_landmarks));
  }
  for (int l = 0; l < p->n_landmarks; ++l) {
    const auto it = track_projections.find(TrackId(l));
    if (it == track_projections.end()) return 20;
    int64_t s = p->lm_obs_ptr[l] + l;
    if (int64_t(it->second.size()) != p->lm_obs_ptr[l + 1] - p->lm_obs_ptr[l] + 1) return 21;
    for (const auto& kv : it->second) {
      const ProjectedLandmark& q = *kv.second;
      if (q.track_id != TrackId(l)) return 22;
      std::memcpy(point_measured + 2 * s, q.point_measured.data(), 16);
      std::memcpy(point_reprojected + 2 * s, q.point_reprojected.data(), 16);
      std::memcpy(point_3d_c + 3 * s, q.point_3d_c.data(), 24);
      reprojection_error[s] = q.reprojection_error;
      outlier_flags[s] = q.outlier_flags;
      ++s;
    }
  }
  for (int i = 0; i < p->n_poses; ++i) {
    const auto it = image_projections.find(m.fcid[i]);
    n_image_obs[i] = it == image_projections.end() ? 0 : int64_t(it->second.obs.size());
  }
  return 0;
}

// add_new_landmarks_between_cams() drop-in proof: the reference's containers (two cameras,
// n shared feature tracks, every third track already a landmark) through either the reference's own function
// (use_b200 = 0; opengv triangulation compiled from the vendored sources) or visnav_b200's (CUDA).  Outputs per
// track: inv_depth (or -1 where no landmark was added / the old one kept) and the number of observations stored.
extern "C" __attribute__((visibility("default"))) int pba_dropin_add_new_landmarks(
    int model0, const double* intr0, int model1, const double* intr1, const double* T_w_c0, const double* T_w_c1,
    int64_t n, const double* uv0, const double* uv1, int use_b200, double* inv_depth, int32_t* n_obs, int32_t* added) {
  using namespace visnav;
  static const char* const names[] = {"pinhole", "ds", "kb4", "eucm"};
  Calibration calib;
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[model0], intr0));
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[model1], intr1));
  calib.T_i_c.push_back(Sophus::SE3d());
  calib.T_i_c.push_back(Sophus::SE3d());
  const FrameCamId fcid0(0, 0), fcid1(0, 1), other(7, 0);  // `other` is in the tracks but not in the map
  Cameras cameras;
  Camera c0, c1;
  std::memcpy(c0.T_w_c.data(), T_w_c0, 7 * sizeof(double));
  std::memcpy(c1.T_w_c.data(), T_w_c1, 7 * sizeof(double));
  cameras[fcid0] = c0;
  cameras[fcid1] = c1;
  Corners corners;
  FeatureTracks tracks;
  Landmarks landmarks;
  for (int64_t i = 0; i < n; ++i) {
    corners[fcid0].corners.emplace_back(uv0[2 * i], uv0[2 * i + 1]);
    corners[fcid1].corners.emplace_back(uv1[2 * i], uv1[2 * i + 1]);
    FeatureTrack t;
    t[fcid0] = FeatureId(i);
    if (i % 5 != 4) t[fcid1] = FeatureId(i);  // every fifth track is not shared
    t[other] = FeatureId(0);
    tracks[TrackId(i)] = t;
    if (i % 3 == 0) {  // already in the map: must be left alone
      Landmark lm;
      lm.inv_depth = -1.0;
      lm.obs[fcid0] = FeatureId(i);
      landmarks[TrackId(i)] = lm;
    }
  }
  const int cnt = use_b200 ? visnav_b200::add_new_landmarks_between_cams(fcid0, fcid1, calib, corners, tracks, cameras, landmarks)
                           : add_new_landmarks_between_cams(fcid0, fcid1, calib, corners, tracks, cameras, landmarks);
  if (cnt < 0) return 30;
  *added = cnt;
  for (int64_t i = 0; i < n; ++i) {
    const auto it = landmarks.find(TrackId(i));
    inv_depth[i] = it == landmarks.end() ? -2.0 : it->second.inv_depth;
    n_obs[i] = it == landmarks.end() ? 0 : int32_t(it->second.obs.size());
  }
  return 0;
}


// Front-end drop-in proof: one stereo pair through the reference's containers (KeypointsData,
// MatchData, Corners, Matches, the calibration's camera objects) with either the reference's own functions
// (use_b200 = 0: computeAngles, computeDescriptors, matchDescriptors, computeEssential, findInliersEssential —
// keypoints.h:182-300, matching_utils.h:50-79, exactly the calls of src/sfm.cpp:1197-1250) or visnav_b200's
// (CUDA).  Outputs: angles / descriptors of both images, the matches (sorted) and the epipolar inliers.
extern "C" __attribute__((visibility("default"))) int pba_dropin_frontend_stereo(
    const uint8_t* image0, const uint8_t* image1, int w, int h, int pitch, int n0, const double* corners0, int n1,
    const double* corners1, int model0, const double* intr0, int model1, const double* intr1, const double* T_0_1,
    int threshold, double dist_2_best, double epipolar_threshold, int use_b200, double* angles0, uint8_t* desc0,
    double* angles1, uint8_t* desc1, int32_t* matches, int32_t* n_matches, int32_t* inliers, int32_t* n_inliers) {
  using namespace visnav;
  static const char* const names[] = {"pinhole", "ds", "kb4", "eucm"};
  const FrameCamId fcid0(0, 0), fcid1(0, 1);
  pangolin::ManagedImage<uint8_t> img0(const_cast<uint8_t*>(image0), w, h, pitch), img1(const_cast<uint8_t*>(image1), w, h, pitch);
  Corners feature_corners;
  for (int i = 0; i < n0; ++i) feature_corners[fcid0].corners.emplace_back(corners0[2 * i], corners0[2 * i + 1]);
  for (int i = 0; i < n1; ++i) feature_corners[fcid1].corners.emplace_back(corners1[2 * i], corners1[2 * i + 1]);
  KeypointsData& kd0 = feature_corners[fcid0];
  KeypointsData& kd1 = feature_corners[fcid1];
  Calibration calib;
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[model0], intr0));
  calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[model1], intr1));
  const Sophus::SE3d T01(Eigen::Quaterniond(T_0_1[3], T_0_1[0], T_0_1[1], T_0_1[2]), Eigen::Vector3d(T_0_1[4], T_0_1[5], T_0_1[6]));
  MatchData md;
  md.T_i_j = T01;
  if (use_b200) {
    if (visnav_b200::computeAnglesAndDescriptors(img0, kd0, true) != PBA_OK) return 40;
    if (visnav_b200::computeAnglesAndDescriptors(img1, kd1, true) != PBA_OK) return 41;
    Matches feature_matches;
    const std::vector<std::pair<FrameCamId, FrameCamId>> pairs = {{fcid0, fcid1}};
    if (visnav_b200::matchImagePairs(feature_corners, pairs, threshold, dist_2_best, feature_matches) != PBA_OK) return 42;
    md.matches = feature_matches.at(std::make_pair(fcid0, fcid1)).matches;
    std::vector<std::pair<int, int>> again;  // the two-set form gives the same list
    if (visnav_b200::matchDescriptors(kd0.corner_descriptors, kd1.corner_descriptors, again, threshold, dist_2_best) != PBA_OK) return 43;
    if (again.size() != md.matches.size()) return 44;
    for (size_t k = 0; k < again.size(); ++k)
      if (again[k].first != md.matches[k].first || again[k].second != md.matches[k].second) return 45;
    if (visnav_b200::findInliersEssential(kd0, kd1, calib.intrinsics[0], calib.intrinsics[1], T01, epipolar_threshold, md) != PBA_OK) return 46;
  } else {
    computeAngles(img0, kd0, true);
    computeDescriptors(img0, kd0);
    computeAngles(img1, kd1, true);
    computeDescriptors(img1, kd1);
    matchDescriptors(kd0.corner_descriptors, kd1.corner_descriptors, md.matches, threshold, dist_2_best);
    std::sort(md.matches.begin(), md.matches.end());
    Eigen::Matrix3d E;
    computeEssential(T01, E);
    findInliersEssential(kd0, kd1, calib.intrinsics[0], calib.intrinsics[1], E, epipolar_threshold, md);
  }
  auto dump = [](const KeypointsData& kd, double* ang, uint8_t* d) {
    std::memset(d, 0, kd.corners.size() * 32);
    for (size_t i = 0; i < kd.corners.size(); ++i) {
      ang[i] = kd.corner_angles[i];
      for (size_t b = 0; b < 256; ++b)
        if (kd.corner_descriptors[i][b]) d[32 * i + b / 8] |= uint8_t(1u << (b % 8));
    }
  };
  dump(kd0, angles0, desc0);
  dump(kd1, angles1, desc1);
  *n_matches = int32_t(md.matches.size());
  for (size_t k = 0; k < md.matches.size(); ++k) { matches[2 * k] = md.matches[k].first; matches[2 * k + 1] = md.matches[k].second; }
  *n_inliers = int32_t(md.inliers.size());
  for (size_t k = 0; k < md.inliers.size(); ++k) { inliers[2 * k] = md.inliers[k].first; inliers[2 * k + 1] = md.inliers[k].second; }
  return 0;
}

// build_tracks() drop-in proof (src/sfm.cpp:1511-1520): the reference's Corners / Matches / FeatureTracks through
// either the reference's TrackBuilder (use_b200 = 0) or visnav_b200::buildTracks (CUDA).  Image i is
// FrameCamId(i / 2, i % 2) (stereo frames).  Output per node (image offset + feature): the smallest node of its
// track, -1 for none — comparable whatever the TrackIds are.
extern "C" __attribute__((visibility("default"))) int pba_dropin_build_tracks(
    int n_images, const int32_t* feat_ptr, int n_pairs, const int32_t* pairs, const int64_t* match_ptr,
    const int32_t* matches, int min_length, int use_b200, int32_t* track_of, int32_t* n_tracks) {
  using namespace visnav;
  std::vector<FrameCamId> ids;
  Corners feature_corners;
  for (int i = 0; i < n_images; ++i) {
    ids.emplace_back(i / 2, size_t(i % 2));
    feature_corners[ids.back()].corners.resize(size_t(feat_ptr[i + 1] - feat_ptr[i]));
  }
  Matches feature_matches;
  for (int k = 0; k < n_pairs; ++k) {
    MatchData md;
    for (int64_t e = match_ptr[k]; e < match_ptr[k + 1]; ++e) md.inliers.emplace_back(matches[2 * e], matches[2 * e + 1]);
    feature_matches[std::make_pair(ids[pairs[2 * k]], ids[pairs[2 * k + 1]])] = md;
  }
  FeatureTracks tracks;
  if (use_b200) {
    if (visnav_b200::buildTracks(feature_matches, feature_corners, size_t(min_length), tracks) != PBA_OK) return 50;
  } else {
    TrackBuilder tb;
    tb.Build(feature_matches);
    tb.Filter(size_t(min_length));
    tb.Export(tracks);
  }
  const int n = feat_ptr[n_images];
  for (int i = 0; i < n; ++i) track_of[i] = -1;
  for (const auto& kv : tracks) {
    int lo = n;
    for (const auto& f : kv.second) lo = std::min(lo, int(feat_ptr[f.first.frame_id * 2 + int(f.first.cam_id)] + f.second));
    for (const auto& f : kv.second) track_of[feat_ptr[f.first.frame_id * 2 + int(f.first.cam_id)] + f.second] = lo;
  }
  *n_tracks = int32_t(tracks.size());
  return 0;
}
