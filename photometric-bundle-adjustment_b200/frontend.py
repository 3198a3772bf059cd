"""Feature front-end on the device: host-side mirror of the reference's
computeAngles / computeDescriptors / matchDescriptors (include/visnav/keypoints.h:182-300) and
computeEssential / findInliersEssential (include/visnav/matching_utils.h:50-79) over the C ABI
(include/pba.h: pba_corner_descriptors, pba_match_descriptors, pba_epipolar_inliers).

Corner detection stays with the caller, as in the reference (cv::goodFeaturesToTrack, keypoints.h:133-151).
A descriptor is 32 bytes: bit d of the reference's std::bitset<256> is bit d % 8 of byte d / 8.
"""
import ctypes as C

import numpy as np

from . import _ffi

EDGE_THRESHOLD = 19  # keypoints.h:50


def _lib():
    lib = _ffi.load_lib()
    if not getattr(lib, "_frontend_bound", False):
        i32p, i64p, u8p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
        lib.pba_corner_descriptors.argtypes = [u8p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, i32p, dp,
                                               C.c_int32, C.c_int32, dp, u8p]
        lib.pba_match_descriptors.argtypes = [C.c_int32, i32p, u8p, C.c_int32, i32p, C.c_int32, C.c_double, C.c_int32,
                                              i64p, i32p, C.c_int64]
        lib.pba_epipolar_inliers.argtypes = [C.c_int32, dp, C.c_int32, dp, dp, C.c_double, C.c_int64, i32p, dp, dp,
                                             C.c_int32, dp, u8p]
        lib.pba_build_tracks.argtypes = [C.c_int32, i32p, C.c_int32, i32p, i64p, i32p, C.c_int32, C.c_int32, i32p, i32p]
        lib._frontend_bound = True
    return lib


def corner_descriptors(images, corners, rotate_features=True, device=0):
    """images [n, h, w] uint8; corners: list of [k_i, 2] arrays (x, y), one per image.
    Returns (angles, descriptors): lists of [k_i] float64 and [k_i, 32] uint8 (detectKeypointsAndDescriptors minus
    the detection, keypoints.h:247-253)."""
    images = np.ascontiguousarray(images, np.uint8)
    if images.ndim == 2:
        images = images[None]
    n, h, w = images.shape
    assert len(corners) == n
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum([len(c) for c in corners])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float64).reshape(-1, 2) for c in corners])
                                if n else np.zeros((0, 2)), np.float64)
    total = int(ptr[-1])
    angles = np.zeros(total)
    desc = np.zeros((total, 32), np.uint8)
    _ffi.check(_lib().pba_corner_descriptors(_ffi.ptr(images, C.c_uint8), n, h * w, w, h, w, _ffi.ptr(ptr, C.c_int32),
                                             _ffi.ptr(flat, C.c_double), int(bool(rotate_features)), device,
                                             _ffi.ptr(angles, C.c_double), _ffi.ptr(desc, C.c_uint8)),
               "pba_corner_descriptors")
    return ([angles[ptr[i]:ptr[i + 1]] for i in range(n)], [desc[ptr[i]:ptr[i + 1]] for i in range(n)])


def match_descriptors(descriptor_sets, pairs, threshold=70, dist_2_best=1.2, device=0):
    """descriptor_sets: list of [k_i, 32] uint8 (one per image); pairs: [m, 2] set indices.
    Returns a list of [q, 2] int32 arrays (index in the first set, index in the second set), ascending in the first
    index — matchDescriptors (keypoints.h:282-300) with the caller's defaults (src/sfm.cpp:200-202)."""
    pairs = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
    ns = len(descriptor_sets)
    ptr = np.zeros(ns + 1, np.int32)
    ptr[1:] = np.cumsum([len(d) for d in descriptor_sets])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in descriptor_sets])
                                if ns else np.zeros((0, 32), np.uint8), np.uint8)
    sizes = np.diff(ptr)
    cap = int(np.minimum(sizes[pairs[:, 0]], sizes[pairs[:, 1]]).sum()) if len(pairs) else 0
    mptr = np.zeros(len(pairs) + 1, np.int64)
    matches = np.zeros((max(cap, 1), 2), np.int32)
    _ffi.check(_lib().pba_match_descriptors(ns, _ffi.ptr(ptr, C.c_int32), _ffi.ptr(flat, C.c_uint8), len(pairs),
                                            _ffi.ptr(pairs, C.c_int32), int(threshold), float(dist_2_best), device,
                                            _ffi.ptr(mptr, C.c_int64), _ffi.ptr(matches, C.c_int32), cap),
               "pba_match_descriptors")
    return [matches[mptr[k]:mptr[k + 1]].copy() for k in range(len(pairs))]


def epipolar_inliers(model0, intr0, model1, intr1, T_0_1, matches, corners0, corners1, threshold=1e-3, device=0):
    """findInliersEssential (matching_utils.h:62-79) with E = computeEssential(T_0_1) (:50-60).
    Returns (E [3, 3], inlier mask [len(matches)] bool); the caller's threshold is 1e-3 (src/sfm.cpp:1249)."""
    m0 = _ffi.CAM_NAMES[model0] if isinstance(model0, str) else int(model0)
    m1 = _ffi.CAM_NAMES[model1] if isinstance(model1, str) else int(model1)
    i0 = np.ascontiguousarray(intr0, np.float64).reshape(8)
    i1 = np.ascontiguousarray(intr1, np.float64).reshape(8)
    T = np.ascontiguousarray(T_0_1, np.float64).reshape(7)
    matches = np.ascontiguousarray(np.asarray(matches, np.int32).reshape(-1, 2))
    c0 = np.ascontiguousarray(corners0, np.float64).reshape(-1, 2)
    c1 = np.ascontiguousarray(corners1, np.float64).reshape(-1, 2)
    if len(matches):
        assert matches[:, 0].max() < len(c0) and matches[:, 1].max() < len(c1)
    E = np.zeros(9)
    inl = np.zeros(max(len(matches), 1), np.uint8)
    _ffi.check(_lib().pba_epipolar_inliers(m0, _ffi.ptr(i0, C.c_double), m1, _ffi.ptr(i1, C.c_double),
                                           _ffi.ptr(T, C.c_double), float(threshold), len(matches),
                                           _ffi.ptr(matches, C.c_int32), _ffi.ptr(c0, C.c_double),
                                           _ffi.ptr(c1, C.c_double), device, _ffi.ptr(E, C.c_double),
                                           _ffi.ptr(inl, C.c_uint8)), "pba_epipolar_inliers")
    return E.reshape(3, 3), inl[:len(matches)].astype(bool)


def build_tracks(feature_counts, pairs, matches, min_length=3, device=0):
    """TrackBuilder Build + Filter + Export (tracks.h:53-160; caller build_tracks, src/sfm.cpp:1511-1520, default minimum
    length 3, src/sfm.cpp:214).  feature_counts: features per image; pairs [m, 2] image indices; matches: list of [q, 2]
    arrays (the inlier matches of each pair).  Returns (track_of, n_tracks): track_of is a list of int32 arrays, one per
    image — the id of the track each feature belongs to (= the smallest node id of the track, node id = image offset +
    feature), -1 for none."""
    counts = np.asarray(feature_counts, np.int64)
    fp = np.zeros(len(counts) + 1, np.int32)
    fp[1:] = np.cumsum(counts)
    pairs = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
    assert len(matches) == len(pairs)
    mp = np.zeros(len(pairs) + 1, np.int64)
    mp[1:] = np.cumsum([len(m) for m in matches])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32).reshape(-1, 2) for m in matches])
                                if len(matches) else np.zeros((0, 2), np.int32), np.int32)
    if len(flat) == 0:
        flat = np.zeros((1, 2), np.int32)
    out = np.full(max(int(fp[-1]), 1), -1, np.int32)
    nt = C.c_int32(0)
    _ffi.check(_lib().pba_build_tracks(len(counts), _ffi.ptr(fp, C.c_int32), len(pairs), _ffi.ptr(pairs, C.c_int32),
                                       _ffi.ptr(mp, C.c_int64), _ffi.ptr(flat, C.c_int32), int(min_length), device,
                                       _ffi.ptr(out, C.c_int32), C.byref(nt)), "pba_build_tracks")
    return [out[fp[i]:fp[i + 1]].copy() for i in range(len(counts))], int(nt.value)
