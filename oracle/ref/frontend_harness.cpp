// oracle/ref/frontend_harness.cpp — C-ABI wrapper around the UNMODIFIED reference front-end functions
// (SURVEY.md §8(f)-1): visnav::computeAngles / computeDescriptors / matchDescriptors
// (include/visnav/keypoints.h:182-300), visnav::computeEssential / findInliersEssential
// (include/visnav/matching_utils.h:50-79).
//
// and visnav::TrackBuilder (include/visnav/tracks.h:53-160).
// TEST INFRASTRUCTURE ONLY: built by oracle/ref/Makefile into oracle/_ref/libpba_ref_frontend.so, loaded only
// by tests/ (golden-vector generation and live parity).  Pangolin and OpenCV are replaced by the two shims
// under oracle/ref/shim (an image view and never-called declarations); nothing of the reference is copied.
#include <visnav/common_types.h>

#include <visnav/camera_models.h>
#include <visnav/keypoints.h>
#include <visnav/matching_utils.h>
#include <visnav/tracks.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define FE_API extern "C" __attribute__((visibility("default")))

namespace {
const char* model_name(int m) {
  switch (m) {
    case 0: return "pinhole";
    case 1: return "ds";
    case 2: return "kb4";
    case 3: return "eucm";
  }
  return "unknown";
}
visnav::KeypointsData corners_of(int n, const double* xy) {
  visnav::KeypointsData kd;
  for (int i = 0; i < n; ++i) kd.corners.emplace_back(xy[2 * i], xy[2 * i + 1]);
  return kd;
}
std::vector<std::bitset<256>> bits_of(int n, const uint8_t* d) {
  std::vector<std::bitset<256>> out(n);
  for (int i = 0; i < n; ++i)
    for (int b = 0; b < 256; ++b) out[i][b] = (d[32 * i + b / 8] >> (b % 8)) & 1;
  return out;
}
}  // namespace

// angles [n], descriptors [n][32]: bit d of the reference's std::bitset<256> -> byte d / 8, bit d % 8
FE_API void pba_ref_corner_descriptors(const uint8_t* image, int w, int h, int pitch, int n, const double* corners,
                                       int rotate_features, double* angles, uint8_t* descriptors) {
  pangolin::ManagedImage<uint8_t> img(const_cast<uint8_t*>(image), w, h, pitch);
  visnav::KeypointsData kd = corners_of(n, corners);
  visnav::computeAngles(img, kd, rotate_features != 0);
  visnav::computeDescriptors(img, kd);
  std::memset(descriptors, 0, size_t(32) * n);
  for (int i = 0; i < n; ++i) {
    angles[i] = kd.corner_angles[i];
    for (int b = 0; b < 256; ++b)
      if (kd.corner_descriptors[i][b]) descriptors[32 * i + b / 8] |= uint8_t(1u << (b % 8));
  }
}

// matches [min(n1, n2)][2], sorted by the first index (the reference's order is its unordered_map's); returns the count
FE_API int pba_ref_match_descriptors(int n1, const uint8_t* d1, int n2, const uint8_t* d2, int threshold,
                                     double dist_2_best, int32_t* matches) {
  std::vector<std::pair<int, int>> m;
  visnav::matchDescriptors(bits_of(n1, d1), bits_of(n2, d2), m, threshold, dist_2_best);
  std::sort(m.begin(), m.end());
  for (size_t k = 0; k < m.size(); ++k) { matches[2 * k] = m[k].first; matches[2 * k + 1] = m[k].second; }
  return int(m.size());
}

// E (row-major 3x3) from T_0_1 = [qx qy qz qw tx ty tz]
FE_API void pba_ref_compute_essential(const double* T, double* E) {
  const Sophus::SE3d T_0_1(Eigen::Quaterniond(T[3], T[0], T[1], T[2]), Eigen::Vector3d(T[4], T[5], T[6]));
  Eigen::Matrix3d Em;
  visnav::computeEssential(T_0_1, Em);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) E[3 * i + j] = Em(i, j);
}

// inlier[k] = 1 when match k passes the reference's epipolar test; returns the number of inliers
FE_API int pba_ref_epipolar_inliers(int n_matches, const int32_t* matches, int n1, const double* corners1, int n2,
                                    const double* corners2, int model1, const double* intr1, int model2,
                                    const double* intr2, const double* E, double threshold, uint8_t* inlier) {
  visnav::KeypointsData kd1 = corners_of(n1, corners1), kd2 = corners_of(n2, corners2);
  Eigen::Matrix<double, 8, 1> p1, p2;
  for (int i = 0; i < 8; ++i) { p1[i] = intr1[i]; p2[i] = intr2[i]; }
  std::shared_ptr<visnav::AbstractCamera<double>> c1 = visnav::AbstractCamera<double>::from_data(model_name(model1), p1.data());
  std::shared_ptr<visnav::AbstractCamera<double>> c2 = visnav::AbstractCamera<double>::from_data(model_name(model2), p2.data());
  Eigen::Matrix3d Em;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Em(i, j) = E[3 * i + j];
  visnav::MatchData md;
  for (int k = 0; k < n_matches; ++k) md.matches.emplace_back(matches[2 * k], matches[2 * k + 1]);
  visnav::findInliersEssential(kd1, kd2, c1, c2, Em, threshold, md);
  std::memset(inlier, 0, n_matches);
  size_t q = 0;  // the reference keeps the matches' order
  for (int k = 0; k < n_matches && q < md.inliers.size(); ++k)
    if (md.matches[k] == md.inliers[q]) { inlier[k] = 1; ++q; }
  return int(md.inliers.size());
}

// The reference's TrackBuilder on the inlier matches of a list of image pairs (image i = FrameCamId(i, 0)).
// track_of [n_nodes]: the SMALLEST node id of the exported track a node belongs to (the reference's own TrackIds are
// the roots of its union-find forest), -1 for nodes in no exported track.  Returns the number of exported tracks.
FE_API int pba_ref_build_tracks(int n_images, const int32_t* feat_ptr, int n_pairs, const int32_t* pairs,
                                const int64_t* match_ptr, const int32_t* matches, int min_length, int32_t* track_of) {
  using namespace visnav;
  Matches feature_matches;
  for (int k = 0; k < n_pairs; ++k) {
    MatchData md;
    for (int64_t e = match_ptr[k]; e < match_ptr[k + 1]; ++e) md.inliers.emplace_back(matches[2 * e], matches[2 * e + 1]);
    feature_matches[std::make_pair(FrameCamId(pairs[2 * k], 0), FrameCamId(pairs[2 * k + 1], 0))] = md;
  }
  TrackBuilder tb;
  tb.Build(feature_matches);
  tb.Filter(size_t(min_length));
  FeatureTracks tracks;
  tb.Export(tracks);
  const int n = n_images > 0 ? feat_ptr[n_images] : 0;
  for (int i = 0; i < n; ++i) track_of[i] = -1;
  for (const auto& kv : tracks) {
    int lo = n;
    for (const auto& f : kv.second) lo = std::min(lo, int(feat_ptr[f.first.frame_id] + f.second));
    for (const auto& f : kv.second) track_of[feat_ptr[f.first.frame_id] + f.second] = lo;
  }
  return int(tracks.size());
}
