"""Feature front-end: orientation + rotated-BRIEF descriptors, mutual descriptor matching and
the stereo pairs' epipolar inlier test — the steps of the reference's SfM pipeline between corner detection and the
BA problem (include/visnav/keypoints.h:182-300, include/visnav/matching_utils.h:50-79, callers
src/sfm.cpp:1191-1330).

Golden vectors: tests/golden/frontend_euroc.npz, produced by the reference's OWN functions (compiled unmodified,
oracle/ref/frontend_harness.cpp) on real EuRoC stereo images (tests/golden/make_golden_frontend.py).
CPU tests pin the oracle restatement to them (and to the reference library live, when it is present);
GPU tests compare the CUDA kernels (through the C ABI) with both.  Bar: descriptors, matches and inlier flags
BIT-EXACT; the orientation angle (fp64 atan2) and E within 1e-14.
"""
import os

import numpy as np
import pytest

import oracle_ffi as of
import pba_b200 as pb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANGLE_TOL = 1e-14


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "frontend_euroc.npz"))


@pytest.fixture(scope="module")
def images(g):
    return np.load(os.path.join(GOLDEN, "euroc_v1_photo.npz"))["images"][: len(g["image_index"])]


def random_descriptors(rng, n, pool=None, flip=0):
    """n descriptors; with `pool`, noisy copies of pool rows (so that real matches, ties and near-ties exist)."""
    if pool is None or len(pool) == 0:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d = pool[rng.integers(0, len(pool), n)].copy()
    for i in range(n):
        for _ in range(int(rng.integers(0, flip + 1))):
            b = int(rng.integers(0, 256))
            d[i, b // 8] ^= np.uint8(1 << (b % 8))
    return d


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_descriptors_match_the_reference_golden(g, images):
    for i in range(len(images)):
        ang, d = of.corner_descriptors("oracle", images[i], g["corners_%d" % i], True)
        assert np.abs(ang - g["angles_%d" % i]).max() <= ANGLE_TOL
        assert np.array_equal(d, g["descriptors_%d" % i])
        ang0, d0 = of.corner_descriptors("oracle", images[i], g["corners_%d" % i], False)
        assert (ang0 == 0).all() and np.array_equal(d0, g["descriptors_norot_%d" % i])


def test_oracle_matches_and_inliers_match_the_reference_golden(g):
    for k, (a, b) in enumerate(g["pairs"]):
        m = of.match_descriptors("oracle", g["descriptors_%d" % a], g["descriptors_%d" % b], int(g["threshold"]),
                                 float(g["dist_2_best"]))
        assert np.array_equal(m, g["matches_%d" % k])
        if "inliers_%d" % k in g.files:
            E, inl = of.epipolar_inliers("oracle", g["calib_model"][0], g["intrinsics"][0], g["calib_model"][1],
                                         g["intrinsics"][1], g["T_0_1"], m, g["corners_%d" % a], g["corners_%d" % b],
                                         float(g["epipolar_threshold"]))
            assert np.abs(E - g["E"]).max() <= 1e-15
            assert np.array_equal(inl, g["inliers_%d" % k])
            assert 0 < inl.sum() < len(inl)


@pytest.mark.skipif(not of.have_ref_frontend(), reason="oracle/_ref/libpba_ref_frontend.so not built")
def test_oracle_matching_equals_the_reference_on_ties_and_edge_cases():
    """Crafted sets: duplicated descriptors (distance-0 ties: the lowest index wins and the ratio test 0 < 0 * 1.2
    keeps the match), near-ties, an empty set, thresholds at the distance itself."""
    rng = np.random.default_rng(5)
    pool = random_descriptors(rng, 40)
    for n1, n2, flip, thr, ratio in [(60, 70, 0, 70, 1.2), (130, 90, 3, 70, 1.2), (257, 300, 40, 70, 1.2),
                                     (50, 0, 0, 70, 1.2), (0, 50, 0, 70, 1.2), (64, 64, 60, 40, 1.0),
                                     (64, 64, 60, 256, 4.0), (33, 47, 20, 1, 1.2)]:
        d1, d2 = random_descriptors(rng, n1, pool, flip), random_descriptors(rng, n2, pool, flip)
        mo = of.match_descriptors("oracle", d1, d2, thr, ratio)
        mr = of.match_descriptors("ref", d1, d2, thr, ratio)
        assert np.array_equal(mo, mr), (n1, n2, flip, thr, ratio)


def test_corners_outside_the_reference_bounds_are_rejected(images):
    """detectKeypoints only keeps corners InBounds(x, y, 19) (keypoints.h:146-150); the C ABI refuses anything else
    before touching the device."""
    for bad in ([18.9, 100.0], [100.0, 480 - 19.0], [752 - 19.0, 50.0], [np.nan, 50.0]):
        with pytest.raises(RuntimeError, match="INVALID_ARGUMENT"):
            pb.corner_descriptors(images[:1], [np.array([[100.0, 100.0], bad])])


@pytest.mark.skipif(pb.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_front_end_has_no_cpu_fallback(g, images):
    with pytest.raises(RuntimeError, match="PBA_ERR_NO_DEVICE"):
        pb.corner_descriptors(images[:1], [g["corners_0"]])
    with pytest.raises(RuntimeError, match="PBA_ERR_NO_DEVICE"):
        pb.match_descriptors([g["descriptors_0"], g["descriptors_1"]], [(0, 1)])
    with pytest.raises(RuntimeError, match="PBA_ERR_NO_DEVICE"):
        pb.epipolar_inliers(1, g["intrinsics"][0], 1, g["intrinsics"][1], g["T_0_1"], g["matches_0"], g["corners_0"],
                            g["corners_1"])


# ------------------------------------------------------------------------------------------------ GPU: the kernels
@pytest.mark.gpu
def test_cuda_descriptors_bit_exact_on_real_images(g, images):
    corners = [g["corners_%d" % i] for i in range(len(images))]
    ang, desc = pb.corner_descriptors(images, corners, True)            # all images in one call
    for i in range(len(images)):
        assert np.abs(ang[i] - g["angles_%d" % i]).max() <= ANGLE_TOL
        assert np.array_equal(desc[i], g["descriptors_%d" % i])
    ang0, desc0 = pb.corner_descriptors(images, corners, False)
    for i in range(len(images)):
        assert (ang0[i] == 0).all() and np.array_equal(desc0[i], g["descriptors_norot_%d" % i])
    a1, d1 = pb.corner_descriptors(images[3], [corners[3]], True)        # one image, 2-D input
    assert np.array_equal(d1[0], g["descriptors_3"])


@pytest.mark.gpu
def test_cuda_matches_and_inliers_bit_exact_on_real_images(g):
    descs = [g["descriptors_%d" % i] for i in range(len(g["image_index"]))]
    ms = pb.match_descriptors(descs, g["pairs"], int(g["threshold"]), float(g["dist_2_best"]))  # all pairs, one call
    for k, (a, b) in enumerate(g["pairs"]):
        assert np.array_equal(ms[k], g["matches_%d" % k]), k
        if "inliers_%d" % k in g.files:
            E, inl = pb.epipolar_inliers(int(g["calib_model"][0]), g["intrinsics"][0], int(g["calib_model"][1]),
                                         g["intrinsics"][1], g["T_0_1"], ms[k], g["corners_%d" % a], g["corners_%d" % b],
                                         float(g["epipolar_threshold"]))
            assert np.abs(E - g["E"]).max() <= 1e-14
            assert np.array_equal(inl, g["inliers_%d" % k])


@pytest.mark.gpu
def test_cuda_matching_equals_the_oracle_on_ties_ragged_and_large_sets():
    """Empty sets, sizes around the CTA (128) and staging (512) granularities, duplicates (ties), many pairs in one
    call including a set matched against itself."""
    rng = np.random.default_rng(17)
    pool = random_descriptors(rng, 64)
    sizes = [0, 1, 127, 128, 129, 511, 512, 513, 1500, 1300, 700]
    sets = [random_descriptors(rng, n, pool, 30) for n in sizes]
    pairs = [(a, b) for a in range(len(sizes)) for b in range(len(sizes)) if (a + 2 * b) % 3 == 0]
    for thr, ratio in [(70, 1.2), (256, 1.0)]:
        ms = pb.match_descriptors(sets, pairs, thr, ratio)
        for k, (a, b) in enumerate(pairs):
            mo = of.match_descriptors("oracle", sets[a], sets[b], thr, ratio)
            assert np.array_equal(ms[k], mo), (sizes[a], sizes[b], thr, ratio)
    assert pb.match_descriptors(sets, []) == []


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4", "eucm"])
def test_cuda_epipolar_inliers_equal_the_oracle_for_every_camera_model(model):
    rng = np.random.default_rng(23)
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 4, 10, model)
    intr = prob.intrinsics[0]
    mid = int(prob.calib_model[0])
    c0 = np.stack([rng.uniform(30, 720, 400), rng.uniform(30, 450, 400)], 1)
    c1 = c0 + rng.normal(0, 1.5, c0.shape) + [-20.0, 0.0]
    m = np.stack([rng.permutation(400)[:300], rng.permutation(400)[:300]], 1).astype(np.int32)
    m[:150, 1] = m[:150, 0]                                             # half of them true correspondences
    T = np.array([0.004, -0.002, 0.001, 0.99999, 0.11, 0.001, -0.002])
    T[:4] /= np.linalg.norm(T[:4])
    for thr in (1e-3, 5e-3):
        Eg, ig = pb.epipolar_inliers(mid, intr, mid, intr, T, m, c0, c1, thr)
        Eo, io = of.epipolar_inliers("oracle", mid, intr, mid, intr, T, m, c0, c1, thr)
        assert np.abs(Eg - Eo).max() <= 1e-15
        assert np.array_equal(ig, io)
        assert 0 < ig.sum() < len(ig)


@pytest.mark.gpu
def test_reference_containers_drive_both_front_ends(g, images):
    """Drop-in proof: one stereo pair through the reference's own containers (KeypointsData, Corners, Matches,
    MatchData, the calibration's camera objects), once with the reference's functions — the calls of
    src/sfm.cpp:1197-1250 — and once with include/visnav_b200/frontend.h on the CUDA kernels; identical outputs."""
    import ctypes as C
    base = os.path.dirname(of.REF_SO)
    path = os.path.join(base, "libpba_dropin_v4.so" if of.REF_SO.endswith("_v4.so") else "libpba_dropin.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpba_dropin.so not built on this box")
    lib = C.CDLL(path)
    u8, d, i32 = C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.pba_dropin_frontend_stereo.argtypes = [u8, u8, C.c_int, C.c_int, C.c_int, C.c_int, d, C.c_int, d, C.c_int, d, C.c_int, d,
                                               d, C.c_int, C.c_double, C.c_double, C.c_int, d, u8, d, u8, i32, i32, i32, i32]
    c0, c1 = np.ascontiguousarray(g["corners_2"]), np.ascontiguousarray(g["corners_3"])
    im0, im1 = np.ascontiguousarray(images[2]), np.ascontiguousarray(images[3])
    h, w = im0.shape
    intr = np.ascontiguousarray(g["intrinsics"])
    T = np.ascontiguousarray(g["T_0_1"])
    res = []
    for use_b200 in (0, 1):
        a0, a1 = np.zeros(len(c0)), np.zeros(len(c1))
        d0, d1 = np.zeros((len(c0), 32), np.uint8), np.zeros((len(c1), 32), np.uint8)
        m, inl = np.zeros((len(c0), 2), np.int32), np.zeros((len(c0), 2), np.int32)
        nm, ni = C.c_int32(), C.c_int32()
        P = pb._ffi.ptr
        rc = lib.pba_dropin_frontend_stereo(P(im0, C.c_uint8), P(im1, C.c_uint8), w, h, w, len(c0), P(c0, C.c_double), len(c1),
                                            P(c1, C.c_double), int(g["calib_model"][0]), P(intr[0], C.c_double),
                                            int(g["calib_model"][1]), P(intr[1], C.c_double), P(T, C.c_double),
                                            int(g["threshold"]), float(g["dist_2_best"]), float(g["epipolar_threshold"]),
                                            use_b200, P(a0, C.c_double), P(d0, C.c_uint8), P(a1, C.c_double), P(d1, C.c_uint8),
                                            P(m, C.c_int32), C.byref(nm), P(inl, C.c_int32), C.byref(ni))
        assert rc == 0, rc
        res.append((a0, d0, a1, d1, m[:nm.value].copy(), inl[:ni.value].copy()))
    ref, gpu = res
    assert np.abs(ref[0] - gpu[0]).max() <= ANGLE_TOL and np.abs(ref[2] - gpu[2]).max() <= ANGLE_TOL
    assert np.array_equal(ref[1], gpu[1]) and np.array_equal(ref[3], gpu[3])
    assert np.array_equal(ref[4], gpu[4]) and np.array_equal(ref[5], gpu[5])
    assert np.array_equal(gpu[4], g["matches_1"])                       # pair (2, 3) of the fixture
    assert np.array_equal(gpu[5], g["matches_1"][g["inliers_1"]])


# ------------------------------------------------------------------------------------------------ feature tracks
def random_match_graph(rng, n_images, n_feat, n_pairs, density):
    """Random pairwise matches (each a partial injective map, as matchDescriptors produces) between random image pairs."""
    counts = rng.integers(max(1, n_feat // 2), n_feat + 1, n_images)
    all_pairs = [(a, b) for a in range(n_images) for b in range(a + 1, n_images)]
    sel = rng.choice(len(all_pairs), min(n_pairs, len(all_pairs)), replace=False)
    pairs, matches = [], []
    for s in sel:
        a, b = all_pairs[s]
        q = int(density * min(counts[a], counts[b]))
        fa = rng.choice(counts[a], q, replace=False)
        fb = rng.choice(counts[b], q, replace=False)
        pairs.append((a, b))
        matches.append(np.stack([np.sort(fa), fb], 1).astype(np.int32))
    return counts, np.array(pairs, np.int32).reshape(-1, 2), matches


def real_match_graph(g):
    """The fixture's pairs: epipolar inliers for the stereo pairs, all matches for the pairs across time."""
    n = len(g["image_index"])
    counts = [len(g["corners_%d" % i]) for i in range(n)]
    matches = [g["matches_%d" % k][g["inliers_%d" % k]] if "inliers_%d" % k in g.files else g["matches_%d" % k]
               for k in range(len(g["pairs"]))]
    return counts, g["pairs"], matches


@pytest.mark.skipif(not of.have_ref_frontend(), reason="oracle/_ref/libpba_ref_frontend.so not built")
def test_oracle_tracks_equal_the_reference_track_builder(g):
    """Partition into tracks, conflict removal (two features of one image) and the minimum length, against the
    reference's TrackBuilder (tracks.h:53-160) on the real match graph and on random ones with many conflicts."""
    rng = np.random.default_rng(31)
    cases = [real_match_graph(g) + (3,), real_match_graph(g) + (2,)]
    for n_img, n_feat, n_pairs, dens, ml in [(6, 40, 10, 0.5, 2), (12, 60, 30, 0.3, 3), (20, 200, 60, 0.15, 3),
                                             (8, 30, 28, 0.9, 4), (5, 10, 0, 0.5, 2)]:
        cases.append(random_match_graph(rng, n_img, n_feat, n_pairs, dens) + (ml,))
    kept_any = dropped_any = False
    for counts, pairs, matches, ml in cases:
        to, no = of.build_tracks("oracle", counts, pairs, matches, ml)
        tr, nr = of.build_tracks("ref", counts, pairs, matches, ml)
        assert no == nr
        for a, b in zip(to, tr):
            assert np.array_equal(a, b)
        kept_any |= no > 0
        touched = set()
        for (a, b), m in zip(pairs, matches):
            touched |= {(int(a), int(x)) for x in m[:, 0]} | {(int(b), int(x)) for x in m[:, 1]}
        dropped_any |= any(to[i][f] < 0 for i, f in touched)
    assert kept_any and dropped_any


@pytest.mark.gpu
def test_cuda_tracks_equal_the_oracle(g):
    rng = np.random.default_rng(37)
    cases = [real_match_graph(g) + (3,)]
    for n_img, n_feat, n_pairs, dens, ml in [(6, 40, 10, 0.5, 2), (20, 200, 60, 0.15, 3), (8, 30, 28, 0.9, 4),
                                             (5, 10, 0, 0.5, 2), (60, 1500, 400, 0.2, 3), (3, 5000, 3, 0.6, 2)]:
        cases.append(random_match_graph(rng, n_img, n_feat, n_pairs, dens) + (ml,))
    # a long chain: image i's feature 0 matched to image i+1's feature 0 (one component of diameter 199)
    chain_counts = [3] * 200
    chain_pairs = [(i, i + 1) for i in range(199)]
    cases.append((chain_counts, np.array(chain_pairs, np.int32), [np.array([[0, 0]], np.int32)] * 199, 3))
    for counts, pairs, matches, ml in cases:
        tg, ng = pb.build_tracks(counts, pairs, matches, ml)
        to, no = of.build_tracks("oracle", counts, pairs, matches, ml)
        assert ng == no
        for a, b in zip(tg, to):
            assert np.array_equal(a, b)
    tg, ng = pb.build_tracks(chain_counts, chain_pairs, [np.array([[0, 0]], np.int32)] * 199, 3)
    assert ng == 1 and all(t[0] == 0 and t[1] == -1 for t in tg)


@pytest.mark.gpu
def test_reference_containers_drive_both_track_builders(g):
    """Drop-in proof for build_tracks() (src/sfm.cpp:1511-1520): the reference's Corners / Matches / FeatureTracks
    through its own TrackBuilder and through visnav_b200::buildTracks on the CUDA kernels; same tracks."""
    import ctypes as C
    base = os.path.dirname(of.REF_SO)
    path = os.path.join(base, "libpba_dropin_v4.so" if of.REF_SO.endswith("_v4.so") else "libpba_dropin.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpba_dropin.so not built on this box")
    lib = C.CDLL(path)
    i32, i64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.pba_dropin_build_tracks.argtypes = [C.c_int, i32, C.c_int, i32, i64, i32, C.c_int, C.c_int, i32, i32]
    rng = np.random.default_rng(41)
    for counts, pairs, matches, ml in [real_match_graph(g) + (3,), random_match_graph(rng, 16, 120, 50, 0.25) + (3,)]:
        fp = np.zeros(len(counts) + 1, np.int32)
        fp[1:] = np.cumsum(counts)
        pr = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        mp = np.zeros(len(pr) + 1, np.int64)
        mp[1:] = np.cumsum([len(m) for m in matches])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32).reshape(-1, 2) for m in matches]))
        res = []
        for use_b200 in (0, 1):
            out = np.full(int(fp[-1]), -1, np.int32)
            nt = C.c_int32()
            P = pb._ffi.ptr
            rc = lib.pba_dropin_build_tracks(len(counts), P(fp, C.c_int32), len(pr), P(pr, C.c_int32), P(mp, C.c_int64),
                                             P(flat, C.c_int32), ml, use_b200, P(out, C.c_int32), C.byref(nt))
            assert rc == 0, rc
            res.append((out, nt.value))
        assert res[0][1] == res[1][1] and res[0][1] > 0
        assert np.array_equal(res[0][0], res[1][0])
