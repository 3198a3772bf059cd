// Shim for include/visnav/keypoints.h:42 (OpenCV's C++ library is not in this image).  Only
// visnav::detectKeypoints() uses OpenCV (cv::goodFeaturesToTrack); the harness never calls it — corner
// detection is done offline with the Python cv2 module of the same OpenCV (tests/golden/make_golden_frontend.py)
// — so these declarations exist solely to let the header compile.
#pragma once
#include <cstdlib>
#include <vector>
#define CV_8U 0
namespace cv {
struct Point2f { float x, y; };
struct Mat {
  Mat(int, int, int, void*) {}
};
inline void goodFeaturesToTrack(const Mat&, std::vector<Point2f>&, int, double, double) { std::abort(); }
}  // namespace cv
