"""Generates tests/golden/frontend_euroc.npz: the reference's OWN front-end functions (include/visnav/keypoints.h
computeAngles / computeDescriptors / matchDescriptors, include/visnav/matching_utils.h computeEssential /
findInliersEssential; compiled unmodified into oracle/_ref/libpba_ref_frontend.so by oracle/ref/Makefile) on REAL
EuRoC images: the first stereo pairs of the bundled data/euroc_V1 sequence, taken from the committed fixture
tests/golden/euroc_v1_photo.npz so that the GPU test needs nothing else.

Corners come from cv2.goodFeaturesToTrack(image, 1500, 0.01, 8) filtered with InBounds(x, y, 19) exactly as
visnav::detectKeypoints does (keypoints.h:133-151; same OpenCV function, Python binding) and are stored in the
fixture, so the test does not depend on OpenCV.

    python tests/golden/make_golden_frontend.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cv2  # noqa: E402

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402

N_IMAGES = 6            # three stereo pairs
NUM_FEATURES = 1500     # src/sfm.cpp:197
MATCH_MAX_DIST = 70     # src/sfm.cpp:200
MATCH_NEXT_BEST = 1.2   # src/sfm.cpp:201-202
EPIPOLAR_THRESHOLD = 1e-3  # src/sfm.cpp:1249
PAIRS = [(0, 1), (2, 3), (4, 5), (0, 2), (1, 3), (0, 4), (3, 4)]  # stereo pairs first, then pairs across time


def detect(img):
    pts = cv2.goodFeaturesToTrack(img, NUM_FEATURES, 0.01, 8).reshape(-1, 2)  # float32, like cv::Point2f
    h, w = img.shape
    b = np.float32(19)
    keep = (b <= pts[:, 0]) & (pts[:, 0] < np.float32(w) - b) & (b <= pts[:, 1]) & (pts[:, 1] < np.float32(h) - b)
    return pts[keep].astype(np.float64)


def main():
    assert of.have_ref_frontend(), "build oracle/_ref first (make ref)"
    src = np.load(os.path.join(HERE, "euroc_v1_photo.npz"))
    images = src["images"][:N_IMAGES]
    calib = pb.load_calibration(os.path.join(ROOT, "tools", "euroc", "opt_calib.json"))
    out = {"image_index": np.arange(N_IMAGES), "pairs": np.array(PAIRS, np.int32), "threshold": MATCH_MAX_DIST,
           "dist_2_best": MATCH_NEXT_BEST, "epipolar_threshold": EPIPOLAR_THRESHOLD,
           "calib_model": calib.calib_model, "intrinsics": calib.intrinsics, "T_0_1": calib.T_i_c[1]}
    assert np.allclose(calib.T_i_c[0], [0, 0, 0, 1, 0, 0, 0])  # T_0_1 = T_i_c[0]^-1 T_i_c[1] (src/sfm.cpp:1223)
    corners, descs = [], []
    for i in range(N_IMAGES):
        c = detect(images[i])
        ang, d = of.corner_descriptors("ref", images[i], c, True)
        ang0, d0 = of.corner_descriptors("ref", images[i], c, False)
        corners.append(c)
        descs.append(d)
        out["corners_%d" % i] = c
        out["angles_%d" % i] = ang
        out["descriptors_%d" % i] = d
        out["descriptors_norot_%d" % i] = d0
        assert (ang0 == 0).all()
        print("image %d: %d corners" % (i, len(c)))
    for k, (a, b) in enumerate(PAIRS):
        m = of.match_descriptors("ref", descs[a], descs[b], MATCH_MAX_DIST, MATCH_NEXT_BEST)
        out["matches_%d" % k] = m
        msg = "pair (%d, %d): %d matches" % (a, b, len(m))
        if b == a + 1 and a % 2 == 0:  # a stereo pair: match_stereo's epipolar test (src/sfm.cpp:1248-1250)
            E, inl = of.epipolar_inliers("ref", calib.calib_model[0], calib.intrinsics[0], calib.calib_model[1],
                                         calib.intrinsics[1], calib.T_i_c[1], m, corners[a], corners[b], EPIPOLAR_THRESHOLD)
            out["E"] = E
            out["inliers_%d" % k] = inl
            msg += ", %d epipolar inliers" % inl.sum()
        print(msg)
    np.savez_compressed(os.path.join(HERE, "frontend_euroc.npz"), **out)


if __name__ == "__main__":
    main()
