// visnav_b200/frontend.h — host-side drop-ins for the reference's feature front-end, the steps
// of the SfM pipeline between corner detection and the BA problem:
//
//   void computeAngles(const pangolin::ManagedImage<uint8_t>&, KeypointsData&, bool rotate_features)
//   void computeDescriptors(const pangolin::ManagedImage<uint8_t>&, KeypointsData&)       keypoints.h:182-245
//   void matchDescriptors(const std::vector<std::bitset<256>>&, const std::vector<std::bitset<256>>&,
//                         std::vector<std::pair<int, int>>& matches, int threshold, double dist_2_best)
//                                                                                         keypoints.h:282-300
//   void computeEssential(const Sophus::SE3d& T_0_1, Eigen::Matrix3d& E)
//   void findInliersEssential(kd1, kd2, cam1, cam2, E, epipolar_error_threshold, MatchData&)
//                                                                              matching_utils.h:50-79
//   TrackBuilder::Build / Filter / Export                                      tracks.h:53-160
//
// Header-only C++14 templates over the reference's own types (they compile unchanged against
// include/visnav/common_types.h, keypoints.h and camera_models.h: oracle/ref/dropin_harness.cpp), on top of the C ABI
// (include/pba.h: pba_corner_descriptors, pba_match_descriptors, pba_epipolar_inliers, pba_build_tracks).  Corner DETECTION
// (detectKeypoints: cv::goodFeaturesToTrack, keypoints.h:133-151) stays where it is.  Every function returns a
// pba_status (the reference's return void); on an error the outputs are left cleared.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include "pba.h"
#include "visnav_b200/bundle_adjustment.h"  // camera_model_id

namespace visnav_b200 {

namespace detail {
// std::bitset<256> <-> the ABI's 32 bytes (bit d -> byte d / 8, bit d % 8)
template <class DescriptorT>
void pack_descriptors(const std::vector<DescriptorT>& in, std::vector<uint8_t>& out) {
  out.assign(in.size() * 32, 0);
  for (size_t i = 0; i < in.size(); ++i)
    for (size_t b = 0; b < 256; ++b)
      if (in[i][b]) out[32 * i + b / 8] |= uint8_t(1u << (b % 8));
}
}  // namespace detail

// computeAngles() + computeDescriptors() of one image: fills kd.corner_angles and kd.corner_descriptors from
// kd.corners.  ImageT is pangolin::ManagedImage<uint8_t> (ptr, w, h, pitch in bytes).
template <class ImageT, class KeypointsDataT>
pba_status computeAnglesAndDescriptors(const ImageT& img_raw, KeypointsDataT& kd, bool rotate_features, int device = 0) {
  const int32_t n = int32_t(kd.corners.size());
  kd.corner_angles.assign(size_t(n), 0.0);
  kd.corner_descriptors.clear();
  kd.corner_descriptors.resize(size_t(n));
  if (n == 0) return PBA_OK;
  std::vector<double> xy(size_t(n) * 2);
  for (int32_t i = 0; i < n; ++i) { xy[2 * i] = kd.corners[size_t(i)][0]; xy[2 * i + 1] = kd.corners[size_t(i)][1]; }
  const int32_t ptr[2] = {0, n};
  std::vector<uint8_t> desc(size_t(n) * 32);
  const pba_status st = pba_corner_descriptors(img_raw.ptr, 1, int64_t(img_raw.pitch) * int64_t(img_raw.h), int32_t(img_raw.w),
                                               int32_t(img_raw.h), int32_t(img_raw.pitch), ptr, xy.data(),
                                               rotate_features ? 1 : 0, device, kd.corner_angles.data(), desc.data());
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::computeAnglesAndDescriptors: %s\n", pba_status_string(st));
    return st;
  }
  for (int32_t i = 0; i < n; ++i)
    for (size_t b = 0; b < 256; ++b) kd.corner_descriptors[size_t(i)][b] = (desc[32 * size_t(i) + b / 8] >> (b % 8)) & 1;
  return PBA_OK;
}

// matchDescriptors(): same arguments as the reference.  The matches come out ascending in the first index (the
// reference's order is that of its unordered_map).
template <class DescriptorT>
pba_status matchDescriptors(const std::vector<DescriptorT>& corner_descriptors_1,
                            const std::vector<DescriptorT>& corner_descriptors_2,
                            std::vector<std::pair<int, int>>& matches, int threshold, double dist_2_best, int device = 0) {
  matches.clear();
  std::vector<uint8_t> d1, d2;
  detail::pack_descriptors(corner_descriptors_1, d1);
  detail::pack_descriptors(corner_descriptors_2, d2);
  d1.insert(d1.end(), d2.begin(), d2.end());
  const int32_t n1 = int32_t(corner_descriptors_1.size()), n2 = int32_t(corner_descriptors_2.size());
  const int32_t set_ptr[3] = {0, n1, n1 + n2}, pair[2] = {0, 1};
  int64_t match_ptr[2] = {0, 0};
  const int64_t cap = n1 < n2 ? n1 : n2;
  std::vector<int32_t> out(size_t(cap) * 2 + 2);
  const pba_status st = pba_match_descriptors(2, set_ptr, d1.data(), 1, pair, threshold, dist_2_best, device, match_ptr,
                                              out.data(), cap);
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::matchDescriptors: %s\n", pba_status_string(st));
    return st;
  }
  for (int64_t k = 0; k < match_ptr[1]; ++k) matches.emplace_back(out[2 * k], out[2 * k + 1]);
  return PBA_OK;
}

// match_stereo / match_all in ONE device call (src/sfm.cpp:1216-1262, 1286-1330 call matchDescriptors once per
// image pair): `pairs` lists (FrameCamId, FrameCamId); feature_matches[pair].matches is filled for every pair.
// MatchesT is the reference's `Matches` (common_types.h: map of pair -> MatchData).
template <class CornersT, class PairT, class MatchesT>
pba_status matchImagePairs(const CornersT& feature_corners, const std::vector<PairT>& pairs, int threshold,
                           double dist_2_best, MatchesT& feature_matches, int device = 0) {
  using FrameCamIdT = typename CornersT::key_type;
  std::vector<FrameCamIdT> ids;
  std::vector<int32_t> set_ptr(1, 0);
  std::vector<uint8_t> desc, one;
  for (const auto& kv : feature_corners) {
    ids.push_back(kv.first);
    detail::pack_descriptors(kv.second.corner_descriptors, one);
    desc.insert(desc.end(), one.begin(), one.end());
    set_ptr.push_back(int32_t(desc.size() / 32));
  }
  auto index_of = [&](const FrameCamIdT& id) {
    for (size_t i = 0; i < ids.size(); ++i)
      if (!(ids[i] < id) && !(id < ids[i])) return int32_t(i);
    return int32_t(-1);
  };
  std::vector<int32_t> pr;
  int64_t cap = 0;
  for (const auto& p : pairs) {
    const int32_t a = index_of(p.first), b = index_of(p.second);
    if (a < 0 || b < 0) return PBA_ERR_INVALID_ARGUMENT;
    pr.push_back(a); pr.push_back(b);
    const int32_t na = set_ptr[a + 1] - set_ptr[a], nb = set_ptr[b + 1] - set_ptr[b];
    cap += na < nb ? na : nb;
  }
  std::vector<int64_t> match_ptr(pairs.size() + 1, 0);
  std::vector<int32_t> out(size_t(cap) * 2 + 2);
  const pba_status st = pba_match_descriptors(int32_t(ids.size()), set_ptr.data(), desc.data(), int32_t(pairs.size()),
                                              pr.data(), threshold, dist_2_best, device, match_ptr.data(), out.data(), cap);
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::matchImagePairs: %s\n", pba_status_string(st));
    return st;
  }
  for (size_t k = 0; k < pairs.size(); ++k) {
    auto& md = feature_matches[std::make_pair(pairs[k].first, pairs[k].second)];
    md.matches.clear();
    for (int64_t q = match_ptr[k]; q < match_ptr[k + 1]; ++q) md.matches.emplace_back(out[2 * q], out[2 * q + 1]);
  }
  return PBA_OK;
}

// computeEssential() + findInliersEssential() as match_stereo uses them (src/sfm.cpp:1223-1250): the essential
// matrix of T_0_1 and md.inliers = the matches with |x_L^T E x_R| <= epipolar_error_threshold, in the matches'
// order.  SE3T is Sophus::SE3d (data() = qx qy qz qw tx ty tz); cam1 / cam2 are the reference's camera pointers
// (name(), data()).  E_out (9 doubles, row-major) is optional.
template <class KeypointsDataT, class CameraPtrT, class SE3T, class MatchDataT>
pba_status findInliersEssential(const KeypointsDataT& kd1, const KeypointsDataT& kd2, const CameraPtrT& cam1,
                                const CameraPtrT& cam2, const SE3T& T_0_1, double epipolar_error_threshold,
                                MatchDataT& md, double* E_out = nullptr, int device = 0) {
  md.inliers.clear();
  const int64_t n = int64_t(md.matches.size());
  std::vector<int32_t> m(size_t(n) * 2 + 2);
  for (int64_t k = 0; k < n; ++k) { m[2 * k] = int32_t(md.matches[size_t(k)].first); m[2 * k + 1] = int32_t(md.matches[size_t(k)].second); }
  std::vector<double> c0(kd1.corners.size() * 2 + 2), c1(kd2.corners.size() * 2 + 2);
  for (size_t i = 0; i < kd1.corners.size(); ++i) { c0[2 * i] = kd1.corners[i][0]; c0[2 * i + 1] = kd1.corners[i][1]; }
  for (size_t i = 0; i < kd2.corners.size(); ++i) { c1[2 * i] = kd2.corners[i][0]; c1[2 * i + 1] = kd2.corners[i][1]; }
  for (int64_t k = 0; k < n; ++k)
    if (m[2 * k] < 0 || size_t(m[2 * k]) >= kd1.corners.size() || m[2 * k + 1] < 0 || size_t(m[2 * k + 1]) >= kd2.corners.size())
      return PBA_ERR_INVALID_ARGUMENT;
  std::vector<uint8_t> inl(size_t(n) + 1);
  const pba_status st = pba_epipolar_inliers(camera_model_id(cam1->name()), cam1->data(), camera_model_id(cam2->name()),
                                             cam2->data(), T_0_1.data(), epipolar_error_threshold, n, m.data(), c0.data(),
                                             c1.data(), device, E_out, inl.data());
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::findInliersEssential: %s\n", pba_status_string(st));
    return st;
  }
  for (int64_t k = 0; k < n; ++k)
    if (inl[size_t(k)]) md.inliers.emplace_back(md.matches[size_t(k)].first, md.matches[size_t(k)].second);
  return PBA_OK;
}

// TrackBuilder::Build + Filter + Export as build_tracks() uses them (src/sfm.cpp:1511-1520; include/visnav/tracks.h:53-160):
// feature tracks = connected components of the inlier-match graph, minus the tracks with two features of one image
// or fewer than min_length features.  feature_corners supplies the images and their feature counts.  TrackIds are
// the smallest node (image, feature) of the track in FrameCamId order — the reference's are the roots of its
// union-find forest; both are opaque keys.
template <class MatchesT, class CornersT, class FeatureTracksT>
pba_status buildTracks(const MatchesT& feature_matches, const CornersT& feature_corners, size_t min_length,
                       FeatureTracksT& feature_tracks, int device = 0) {
  using FrameCamIdT = typename CornersT::key_type;
  using TrackIdT = typename FeatureTracksT::key_type;
  feature_tracks.clear();
  std::vector<FrameCamIdT> ids;
  for (const auto& kv : feature_corners) ids.push_back(kv.first);
  std::sort(ids.begin(), ids.end());
  std::vector<int32_t> feat_ptr(1, 0);
  for (const auto& id : ids) feat_ptr.push_back(feat_ptr.back() + int32_t(feature_corners.at(id).corners.size()));
  auto index_of = [&](const FrameCamIdT& id) {
    const auto it = std::lower_bound(ids.begin(), ids.end(), id);
    return (it != ids.end() && !(id < *it)) ? int32_t(it - ids.begin()) : int32_t(-1);
  };
  std::vector<int32_t> pairs, matches;
  std::vector<int64_t> match_ptr(1, 0);
  for (const auto& kv : feature_matches) {
    const int32_t a = index_of(kv.first.first), b = index_of(kv.first.second);
    if (a < 0 || b < 0) return PBA_ERR_INVALID_ARGUMENT;
    pairs.push_back(a); pairs.push_back(b);
    for (const auto& m : kv.second.inliers) { matches.push_back(int32_t(m.first)); matches.push_back(int32_t(m.second)); }
    match_ptr.push_back(int64_t(matches.size() / 2));
  }
  matches.push_back(0); matches.push_back(0);  // never an empty buffer
  pairs.push_back(0); pairs.push_back(0);
  std::vector<int32_t> track_of(size_t(feat_ptr.back()) + 1, -1);
  const pba_status st = pba_build_tracks(int32_t(ids.size()), feat_ptr.data(), int32_t(match_ptr.size() - 1), pairs.data(),
                                         match_ptr.data(), matches.data(), int32_t(min_length), device, track_of.data(), nullptr);
  if (st != PBA_OK) {
    std::fprintf(stderr, "visnav_b200::buildTracks: %s\n", pba_status_string(st));
    return st;
  }
  for (size_t im = 0; im < ids.size(); ++im)
    for (int32_t f = 0; f < feat_ptr[im + 1] - feat_ptr[im]; ++f) {
      const int32_t t = track_of[size_t(feat_ptr[im] + f)];
      if (t >= 0) feature_tracks[TrackIdT(t)].emplace(ids[im], f);
    }
  return PBA_OK;
}

}  // namespace visnav_b200
