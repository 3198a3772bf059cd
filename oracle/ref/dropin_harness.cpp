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
#include <visnav/map_utils.h>

#include <cstring>
#include <set>
#include <vector>

#include "pba.h"
#include "visnav_b200/bundle_adjustment.h"

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

