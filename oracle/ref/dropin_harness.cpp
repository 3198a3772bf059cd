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

extern "C" __attribute__((visibility("default"))) int pba_dropin_solve(pba_problem* p, const pba_options* opt,
                                                                        int use_b200, pba_summary* summary) {
  using namespace visnav;
  if (p->mode != PBA_MODE_GEOMETRIC) return 10;
  Corners corners;
  Cameras cameras;
  Landmarks landmarks;
  Calibration calib;
  std::set<FrameCamId> fixed;
  for (int i = 0; i < p->n_calib; ++i) {
    const char* names[] = {"pinhole", "ds", "kb4", "eucm"};
    calib.intrinsics.push_back(AbstractCamera<double>::from_data(names[p->calib_model[i]], p->intrinsics + 8 * i));
    calib.T_i_c.push_back(Sophus::SE3d());
  }
  std::vector<FrameCamId> fcid(p->n_poses);
  for (int i = 0; i < p->n_poses; ++i) {
    fcid[i] = FrameCamId(i, size_t(p->pose_calib[i]));
    Camera cam;
    std::memcpy(cam.T_w_c.data(), p->poses + 7 * i, 7 * sizeof(double));
    cameras[fcid[i]] = cam;
    corners[fcid[i]];
    if (p->pose_fixed && p->pose_fixed[i]) fixed.insert(fcid[i]);
  }
  for (int l = 0; l < p->n_landmarks; ++l) {
    Landmark lm;
    lm.inv_depth = p->inv_depth[l];
    const int h = p->lm_host[l];
    auto& hc = corners[fcid[h]].corners;
    lm.obs[fcid[h]] = FeatureId(hc.size());
    hc.emplace_back(p->lm_host_uv[2 * l], p->lm_host_uv[2 * l + 1]);
    for (int64_t o = p->lm_obs_ptr[l]; o < p->lm_obs_ptr[l + 1]; ++o) {
      const int t = p->obs_target[o];
      if (!(fcid[h] < fcid[t])) return 11;  // host must be obs.begin()
      auto& tc = corners[fcid[t]].corners;
      lm.obs[fcid[t]] = FeatureId(tc.size());
      tc.emplace_back(p->obs_uv[2 * o], p->obs_uv[2 * o + 1]);
    }
    landmarks[TrackId(l)] = lm;
  }
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
