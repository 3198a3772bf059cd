"""Calibration container + the reference's JSON formats (host-side mirror).

Mirrors `visnav::Calibration` (include/visnav/calibration.h:82-93: per-camera
extrinsics T_i_c and intrinsics) and the two cereal JSON schemas the reference
reads/writes (include/visnav/serialization.h):

  * data/euroc_calib/calibration-double-sphere.json — `LoadCalibration<double,
    DoubleSphereCamera>` (serialization.h:92-113, :171-174): intrinsics as
    {fx, fy, cx, cy, xi, alpha};
  * opt_calib.json written by the calibration app and loaded by sfm
    (src/sfm.cpp:933-957) — `Calibration` with generic cameras
    {cam_type, fx, fy, cx, cy, p1..p4, width, height} (serialization.h:116-143).

SE3 is stored as {px, py, pz, qx, qy, qz, qw} (serialization.h:156-164); in memory
we keep Sophus' parameter order qx qy qz qw tx ty tz (the layout of pba_problem.poses).
"""
import json

import numpy as np

from . import _ffi

_NAMES = {v: k for k, v in _ffi.CAM_NAMES.items()}


class Calibration:
    def __init__(self, T_i_c, models, intrinsics, widths=None, heights=None):
        self.T_i_c = np.ascontiguousarray(T_i_c, np.float64).reshape(-1, 7)
        self.models = [m if isinstance(m, str) else _NAMES[int(m)] for m in models]
        self.intrinsics = np.ascontiguousarray(intrinsics, np.float64).reshape(-1, 8)
        n = len(self.models)
        self.widths = list(widths) if widths is not None else [0] * n
        self.heights = list(heights) if heights is not None else [0] * n
        for m in self.models:
            if m not in _ffi.CAM_NAMES:
                # AbstractCamera::from_data aborts on unknown names (camera_models.h:469-473)
                raise ValueError("Camera model %s is not implemented." % m)

    @property
    def calib_model(self):
        """PBA_CAM_* ids, the `calib_model` array of pba_problem."""
        return np.array([_ffi.CAM_NAMES[m] for m in self.models], np.int32)


def _se3_from_json(d):
    return [d["qx"], d["qy"], d["qz"], d["qw"], d["px"], d["py"], d["pz"]]


def _se3_to_json(T):
    return {"px": T[4], "py": T[5], "pz": T[6], "qx": T[0], "qy": T[1], "qz": T[2], "qw": T[3]}


def load_calibration(path):
    """Reads either schema; double-sphere-only files load as model "ds"."""
    with open(path) as f:
        root = json.load(f)
    v = root["value0"]
    T = [_se3_from_json(d) for d in v["cam.T_i_c"]]
    models, intr, widths, heights = [], [], [], []
    for c in v["cam.intrinsics"]:
        if "cam_type" in c:
            models.append(c["cam_type"])
            intr.append([c["fx"], c["fy"], c["cx"], c["cy"], c["p1"], c["p2"], c["p3"], c["p4"]])
            widths.append(int(c.get("width", 0)))
            heights.append(int(c.get("height", 0)))
        else:  # LoadCalibration<double, DoubleSphereCamera>
            models.append("ds")
            intr.append([c["fx"], c["fy"], c["cx"], c["cy"], c["xi"], c["alpha"], 0.0, 0.0])
            widths.append(0)
            heights.append(0)
    return Calibration(T, models, intr, widths, heights)


def save_calibration(path, calib):
    """Writes the generic-camera schema (what src/calibration.cpp:430-438 produces)."""
    cams = []
    for m, p, w, h in zip(calib.models, calib.intrinsics, calib.widths, calib.heights):
        cams.append({"cam_type": m, "fx": p[0], "fy": p[1], "cx": p[2], "cy": p[3], "p1": p[4], "p2": p[5],
                     "p3": p[6], "p4": p[7], "width": int(w), "height": int(h)})
    root = {"value0": {"cam.T_i_c": [_se3_to_json(T) for T in calib.T_i_c.tolist()], "cam.intrinsics": cams}}
    with open(path, "w") as f:
        json.dump(root, f, indent=4)


def initialize_from_double_sphere(model, ds_intrinsics):
    """AbstractCamera::initialize (camera_models.h:477-519): start another model from a
    double-sphere calibration."""
    p = np.array(ds_intrinsics, np.float64).copy()
    if model == "ds":
        return p
    p[4:] = 0.0
    if model == "eucm":
        p[4], p[5] = 0.5, 1.0
    elif model not in ("pinhole", "kb4"):
        raise ValueError("Camera model %s is not implemented." % model)
    return p
