"""the reference's map archive (save_map_file / load_map_file,
include/visnav/map_utils.h:58-116, cereal binary; serialization.h:116-205).

tests/golden/map_geom_ds.cereal was written by the reference's OWN save_map_file
(tests/golden/make_golden.py --map) from the geom_ds fixture.  The Python reader must
recover the fixture's problem from it, the writer must re-emit it byte for byte, and —
when oracle/_ref is present — the reference's own load_map_file must understand a file
written here."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util as gu
import oracle_ffi as of
import pba_b200 as pb
from pba_b200 import map_io

MAP = os.path.join(gu.GOLDEN_DIR, "map_geom_ds.cereal")
FIX = os.path.join(gu.GOLDEN_DIR, "geom_ds.npz")


def test_reader_recovers_the_fixture_problem():
    prob, _ = gu.load(FIX)
    m = pb.load_map_file(MAP)
    assert len(m.cameras) == prob.n_poses and len(m.landmarks) == prob.n_landmarks
    fc = sorted(m.cameras)
    assert np.array_equal(np.array([m.cameras[k] for k in fc]), prob.poses)      # bit-exact T_w_c
    assert len(m.feature_tracks) == prob.n_landmarks and list(m.outlier_tracks) == [1000000]
    assert len(m.feature_matches) == prob.n_poses - 1
    # one landmark carries an outlier observation (harness filler); put it back and flatten
    moved = [(t, lm) for t, lm in m.landmarks.items() if lm.outlier_obs]
    assert len(moved) == 1
    moved[0][1].obs.update(moved[0][1].outlier_obs)
    p2, tids = map_io.map_to_problem(m, prob.calib_model, prob.intrinsics)
    order = np.argsort(tids)   # TrackId = fixture landmark index; unordered_map order in the file
    assert sorted(tids) == list(range(prob.n_landmarks))
    assert np.array_equal(p2.inv_depth[order], prob.inv_depth)
    assert np.array_equal(p2.lm_host[order], prob.lm_host)
    assert np.array_equal(p2.lm_host_uv[order], prob.lm_host_uv)
    cnt = np.diff(p2.lm_obs_ptr)[order]
    assert np.array_equal(cnt, np.diff(prob.lm_obs_ptr))
    for j, l in enumerate(order[:20]):
        a = slice(p2.lm_obs_ptr[l], p2.lm_obs_ptr[l + 1])
        b = slice(prob.lm_obs_ptr[j], prob.lm_obs_ptr[j + 1])
        assert np.array_equal(p2.obs_target[a], prob.obs_target[b]) and np.array_equal(p2.obs_uv[a], prob.obs_uv[b])


def test_writer_re_emits_the_reference_file_byte_for_byte():
    data = open(MAP, "rb").read()
    m = map_io.loads(data)
    kd = next(iter(m.feature_corners.values()))
    assert kd.corner_descriptors.shape[1] == 32 and kd.corner_angles.size == kd.corners.shape[0]
    assert map_io.dumps(m) == data


def test_corrupt_archives_are_rejected():
    data = open(MAP, "rb").read()
    with pytest.raises(ValueError):
        map_io.loads(data[:-5])
    with pytest.raises(ValueError):
        map_io.loads(data + b"\0")


def test_optimised_state_round_trips_through_the_map(tmp_path):
    prob, _ = gu.load(FIX)
    m = pb.load_map_file(MAP)
    p2, tids = map_io.map_to_problem(m, prob.calib_model, prob.intrinsics, fixed_cameras=[(0, 0), (1, 0)])
    assert p2.pose_fixed.sum() == 2
    p2.poses[2:, 4:] += 0.01          # stand-in for an optimisation result
    p2.inv_depth *= 1.01
    map_io.update_map_from_problem(m, p2, tids)
    out = tmp_path / "opt.cereal"
    pb.save_map_file(out, m)
    m2 = pb.load_map_file(out)
    assert np.array_equal(np.array([m2.cameras[k] for k in sorted(m2.cameras)]), p2.poses)
    assert all(m2.landmarks[t].inv_depth == p2.inv_depth[l] for l, t in enumerate(tids))


@pytest.mark.skipif(not of.have_ref(), reason="oracle/_ref not built")
def test_reference_load_map_file_reads_what_we_write(tmp_path):
    """Python-written archive -> the reference's load_map_file -> the reference's save_map_file: same content."""
    m = pb.load_map_file(MAP)
    for lm in m.landmarks.values():
        lm.inv_depth *= 0.5
    mine = tmp_path / "mine.cereal"
    back = tmp_path / "back.cereal"
    pb.save_map_file(mine, m)
    lib = of.ref()
    lib.pba_ref_map_roundtrip.argtypes = [C.c_char_p, C.c_char_p, pb._ffi.c_i64_p]
    counts = np.zeros(6, np.int64)
    assert lib.pba_ref_map_roundtrip(str(mine).encode(), str(back).encode(), pb._ffi.ptr(counts, C.c_int64)) == 0
    assert list(counts) == [len(m.feature_corners), len(m.feature_matches), len(m.feature_tracks),
                            len(m.outlier_tracks), len(m.cameras), len(m.landmarks)]
    m2 = pb.load_map_file(back)
    assert m2.cameras.keys() == m.cameras.keys() and all(np.array_equal(m2.cameras[k], m.cameras[k]) for k in m.cameras)
    assert {t: (l.inv_depth, l.obs, l.outlier_obs) for t, l in m2.landmarks.items()} == \
        {t: (l.inv_depth, l.obs, l.outlier_obs) for t, l in m.landmarks.items()}
    assert m2.feature_tracks == m.feature_tracks and m2.outlier_tracks == m.outlier_tracks
    for k, kd in m.feature_corners.items():
        k2 = m2.feature_corners[k]
        assert np.array_equal(kd.corners, k2.corners) and np.array_equal(kd.corner_angles, k2.corner_angles)
        assert np.array_equal(kd.corner_descriptors, k2.corner_descriptors)
    for k, md in m.feature_matches.items():
        d2 = m2.feature_matches[k]
        assert np.array_equal(md.matches, d2.matches) and np.array_equal(md.inliers, d2.inliers)
        assert np.abs(md.T_i_j - d2.T_i_j).max() < 1e-15   # SE3d re-normalises the quaternion on load
