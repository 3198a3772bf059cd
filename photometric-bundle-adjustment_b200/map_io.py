"""The reference's map archive (`map.cereal`) — SURVEY.md §8(f)-4.

`save_map_file` / `load_map_file` (include/visnav/map_utils.h:58-116) write six containers with
cereal's BinaryOutputArchive, in this order: feature_corners, feature_matches, feature_tracks,
outlier_tracks, cameras, landmarks.  This module reads and writes that byte format so that maps
optimised by the B200 engine round-trip into the reference GUI and vice versa.

Encoding (cereal portable-size binary archive, little endian, serialization.h:116-205):
  container          u64 count, then the items (maps: key then value per item)
  vector<arithmetic> u64 count, then the raw values
  FrameCamId         i64 frame_id, u64 cam_id                      (serialization.h:200-203)
  Eigen fixed matrix its coefficients, row-major walk              (serialization.h:57-67)
  Sophus::SE3d       px py pz qx qy qz qw (7 f64)                  (serialization.h:150-159)
  std::bitset<256>   u8 type (3 = bits), then 32 bytes, bit i -> byte i/8, mask 0x80 >> (i%8)
  KeypointsData      corners, corner_angles, corner_descriptors    (serialization.h:183-187)
  MatchData          T_i_j, inliers, matches                       (serialization.h:171-174)
  Camera             T_w_c;   Landmark  inv_depth, obs, outlier_obs (serialization.h:189-198)
Host-side data format only: no GPU work here.
"""
import dataclasses
import io
import struct
from typing import Dict, List, Tuple

import numpy as np

FrameCamId = Tuple[int, int]  # (frame_id, cam_id); ordering = the reference's operator< (common_types.h:87-90)


@dataclasses.dataclass
class KeypointsData:
    corners: np.ndarray             # [n, 2] f64
    corner_angles: np.ndarray       # [n] f64 (may be empty)
    corner_descriptors: np.ndarray  # [n, 32] u8: 256-bit descriptors in the archive's byte order (may be empty)


@dataclasses.dataclass
class MatchData:
    T_i_j: np.ndarray    # [7] qx qy qz qw tx ty tz (Sophus data order, as pba_problem.poses)
    matches: np.ndarray  # [m, 2] i32
    inliers: np.ndarray  # [k, 2] i32


@dataclasses.dataclass
class MapLandmark:
    inv_depth: float
    obs: Dict[FrameCamId, int]
    outlier_obs: Dict[FrameCamId, int]


@dataclasses.dataclass
class Map:
    feature_corners: Dict[FrameCamId, KeypointsData]
    feature_matches: Dict[Tuple[FrameCamId, FrameCamId], MatchData]
    feature_tracks: Dict[int, Dict[FrameCamId, int]]
    outlier_tracks: Dict[int, Dict[FrameCamId, int]]
    cameras: Dict[FrameCamId, np.ndarray]  # T_w_c as [7] qx qy qz qw tx ty tz
    landmarks: Dict[int, MapLandmark]


# ------------------------------------------------------------------ reading --
class _Reader:
    def __init__(self, data):
        self.b = memoryview(data)
        self.o = 0

    def take(self, fmt):
        if self.o + struct.calcsize("<" + fmt) > len(self.b):
            raise ValueError("map archive truncated")
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def array(self, dtype, count):
        n = np.dtype(dtype).itemsize * count
        a = np.frombuffer(self.b[self.o:self.o + n], dtype=dtype).copy()
        if a.size != count:
            raise ValueError("map archive truncated")
        self.o += n
        return a

    def size(self):
        n = self.take("Q")
        if n > len(self.b):
            raise ValueError("map archive corrupt (container size %d)" % n)
        return n

    def fcid(self):
        return (self.take("q"), self.take("Q"))

    def se3(self):
        p = self.array("<f8", 7)  # px py pz qx qy qz qw
        return np.array([p[3], p[4], p[5], p[6], p[0], p[1], p[2]])

    def track(self):
        return {self.fcid(): self.take("i") for _ in range(self.size())}


def loads(data) -> Map:
    r = _Reader(data)
    corners = {}
    for _ in range(r.size()):
        k = r.fcid()
        c = r.array("<f8", 2 * r.size()).reshape(-1, 2)
        ang = r.array("<f8", r.size())
        nd = r.size()
        desc = np.zeros((nd, 32), np.uint8)
        for i in range(nd):
            if r.take("B") != 3:
                raise ValueError("unsupported bitset encoding in map archive")
            desc[i] = r.array("u1", 32)
        corners[k] = KeypointsData(c, ang, desc)
    matches = {}
    for _ in range(r.size()):
        k = (r.fcid(), r.fcid())
        T = r.se3()
        inl = r.array("<i4", 2 * r.size()).reshape(-1, 2)
        mat = r.array("<i4", 2 * r.size()).reshape(-1, 2)
        matches[k] = MatchData(T, mat, inl)
    tracks = {r.take("q"): r.track() for _ in range(r.size())}
    outlier = {r.take("q"): r.track() for _ in range(r.size())}
    cameras = {}
    for _ in range(r.size()):
        k = r.fcid()
        cameras[k] = r.se3()
    landmarks = {}
    for _ in range(r.size()):
        tid = r.take("q")
        rho = r.take("d")
        landmarks[tid] = MapLandmark(rho, r.track(), r.track())
    if r.o != len(r.b):
        raise ValueError("map archive has %d trailing bytes" % (len(r.b) - r.o))
    return Map(corners, matches, tracks, outlier, cameras, landmarks)


def load_map_file(path) -> Map:
    """visnav::load_map_file (map_utils.h:88-116)."""
    with open(path, "rb") as f:
        return loads(f.read())


# ------------------------------------------------------------------ writing --
def _fcid(w, k):
    w.write(struct.pack("<qQ", int(k[0]), int(k[1])))


def _se3(w, T):
    T = np.asarray(T, np.float64).reshape(7)
    w.write(np.array([T[4], T[5], T[6], T[0], T[1], T[2], T[3]], "<f8").tobytes())


def _track(w, t):
    w.write(struct.pack("<Q", len(t)))
    for k in sorted(t):  # std::map iteration order
        _fcid(w, k)
        w.write(struct.pack("<i", int(t[k])))


def dumps(m: Map) -> bytes:
    w = io.BytesIO()
    w.write(struct.pack("<Q", len(m.feature_corners)))
    for k, kd in m.feature_corners.items():
        _fcid(w, k)
        c = np.ascontiguousarray(kd.corners, "<f8").reshape(-1, 2)
        w.write(struct.pack("<Q", c.shape[0])); w.write(c.tobytes())
        a = np.ascontiguousarray(kd.corner_angles, "<f8").reshape(-1)
        w.write(struct.pack("<Q", a.size)); w.write(a.tobytes())
        d = np.ascontiguousarray(kd.corner_descriptors, np.uint8).reshape(-1, 32)
        w.write(struct.pack("<Q", d.shape[0]))
        for row in d:
            w.write(b"\x03"); w.write(row.tobytes())
    w.write(struct.pack("<Q", len(m.feature_matches)))
    for k, md in m.feature_matches.items():
        _fcid(w, k[0]); _fcid(w, k[1])
        _se3(w, md.T_i_j)
        for arr in (md.inliers, md.matches):
            p = np.ascontiguousarray(arr, "<i4").reshape(-1, 2)
            w.write(struct.pack("<Q", p.shape[0])); w.write(p.tobytes())
    for tracks in (m.feature_tracks, m.outlier_tracks):
        w.write(struct.pack("<Q", len(tracks)))
        for tid, t in tracks.items():
            w.write(struct.pack("<q", int(tid)))
            _track(w, t)
    w.write(struct.pack("<Q", len(m.cameras)))
    for k in sorted(m.cameras):
        _fcid(w, k)
        _se3(w, m.cameras[k])
    w.write(struct.pack("<Q", len(m.landmarks)))
    for tid, lm in m.landmarks.items():
        w.write(struct.pack("<qd", int(tid), float(lm.inv_depth)))
        _track(w, lm.obs)
        _track(w, lm.outlier_obs)
    return w.getvalue()


def save_map_file(path, m: Map):
    """visnav::save_map_file (map_utils.h:58-86)."""
    with open(path, "wb") as f:
        f.write(dumps(m))


# ------------------------------------------------------- map <-> flat problem --
def map_to_problem(m: Map, calib_model, intrinsics, fixed_cameras=(), mode=0):
    """Flatten the map the way visnav::bundle_adjustment does (map_utils.h:327-375): cameras in
    FrameCamId order, host = first observation of a landmark, one observation per further one.
    Returns (Problem, track ids in landmark order)."""
    from .problem import Problem
    fcids = sorted(m.cameras)
    index = {k: i for i, k in enumerate(fcids)}
    poses = np.array([m.cameras[k] for k in fcids]).reshape(-1, 7)
    fixed = np.array([1 if k in set(fixed_cameras) else 0 for k in fcids], np.uint8)
    pose_calib = np.array([k[1] for k in fcids], np.int32)
    rho, host, host_uv, ptr, target, uv, tids = [], [], [], [0], [], [], []
    for tid, lm in m.landmarks.items():
        obs = sorted(lm.obs.items())
        if not obs:
            continue
        tids.append(tid)
        rho.append(lm.inv_depth)
        (hk, hf) = obs[0]
        host.append(index[hk])
        host_uv.append(m.feature_corners[hk].corners[hf])
        for k, fid in obs[1:]:
            target.append(index[k])
            uv.append(m.feature_corners[k].corners[fid])
        ptr.append(len(target))
    prob = Problem(mode, poses, fixed, pose_calib, calib_model, intrinsics, np.array(rho), np.array(host, np.int32),
                   np.array(host_uv).reshape(-1, 2), np.array(ptr, np.int64), np.array(target, np.int32),
                   np.array(uv).reshape(-1, 2))
    return prob, tids


def update_map_from_problem(m: Map, prob, tids):
    """Write optimised poses / inverse distances back into the map (in place)."""
    for i, k in enumerate(sorted(m.cameras)):
        m.cameras[k] = np.array(prob.poses[i])
    for l, tid in enumerate(tids):
        m.landmarks[tid].inv_depth = float(prob.inv_depth[l])
