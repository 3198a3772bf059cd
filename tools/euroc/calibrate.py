"""Stereo calibration for BASELINE config 1 (data/euroc_V1), exactly the way the reference's calibration
application does it (src/calibration.cpp:245-425), but head-less.

The snapshot ships the calibration *exercise* inputs — data/euroc_calib/{detected_corners.json, init_poses.json,
calibration-double-sphere.json (an initial guess: identity extrinsics, xi = 0, alpha = 0.5)} — and expects
`opt_calib.json` to be produced by its GUI application.  This script loads the same three files, runs the same
optimisation through oracle/_ref (pba_ref_calibrate: the reference's own ReprojectionCostFunctor + AprilGrid +
LocalParameterizationSE3 + vendored Ceres with the application's solver options) and writes opt_calib.json in the
schema the reference's sfm reads (src/sfm.cpp:934-950, src/calibration.cpp:430-438).

    python tools/euroc/calibrate.py [--model ds] [--out tools/euroc/opt_calib.json]
Needs /root/reference (it is an offline input generator, like tests/golden/make_golden*.py).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402
from pba_b200 import _ffi  # noqa: E402
from pba_b200.calibration import Calibration, initialize_from_double_sphere, save_calibration  # noqa: E402

CALIB_DIR = "/root/reference/data/euroc_calib"


def se3(j):
    return [j["qx"], j["qy"], j["qz"], j["qw"], j["px"], j["py"], j["pz"]]


def load_inputs(path=CALIB_DIR):
    corners = json.load(open(os.path.join(path, "detected_corners.json")))["value0"]
    poses = json.load(open(os.path.join(path, "init_poses.json")))["value0"]
    calib = json.load(open(os.path.join(path, "calibration-double-sphere.json")))["value0"]
    n_frames = len(corners) // 2
    obs = []  # (frame, cam, corner id, u, v)
    for e in corners:
        f, c = e["key"]["first"], e["key"]["second"]
        for uv, cid in zip(e["value"]["value0"], e["value"]["value1"]):
            obs.append((f, c, cid, uv["value0"], uv["value1"]))
    T_w_i = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], np.float64), (n_frames, 1))
    for e in poses:  # vec_T_w_i[frame] = init pose of camera 0 (src/calibration.cpp:319-323)
        if e["key"]["second"] == 0:
            T_w_i[e["key"]["first"]] = se3(e["value"]["value0"])
    intr_ds = np.array([[c["fx"], c["fy"], c["cx"], c["cy"], c["xi"], c["alpha"], 0, 0] for c in calib["cam.intrinsics"]])
    T_i_c = np.array([se3(t) for t in calib["cam.T_i_c"]], np.float64)
    return n_frames, np.array(obs, np.float64), T_w_i, intr_ds, T_i_c


def calibrate(model="ds"):
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    n_frames, obs, T_w_i, intr_ds, T_i_c = load_inputs()
    intr = np.ascontiguousarray([initialize_from_double_sphere(model, p) for p in intr_ds])
    lib = of.ref()
    d, i32 = _ffi.c_double_p, _ffi.c_i32_p
    lib.pba_ref_calibrate.argtypes = [C.c_int, C.c_int, C.c_int, d, d, d, C.c_int64, i32, i32, i32, d, C.c_int, d, d]
    frame = np.ascontiguousarray(obs[:, 0], np.int32)
    cam = np.ascontiguousarray(obs[:, 1], np.int32)
    cid = np.ascontiguousarray(obs[:, 2], np.int32)
    uv = np.ascontiguousarray(obs[:, 3:5])
    T_i_c = np.ascontiguousarray(T_i_c)
    T_w_i = np.ascontiguousarray(T_w_i)
    c0, c1 = C.c_double(), C.c_double()
    rc = lib.pba_ref_calibrate(_ffi.CAM_NAMES[model], n_frames, 2, _ffi.ptr(intr, C.c_double), _ffi.ptr(T_i_c, C.c_double),
                               _ffi.ptr(T_w_i, C.c_double), len(obs), _ffi.ptr(frame, C.c_int32), _ffi.ptr(cam, C.c_int32),
                               _ffi.ptr(cid, C.c_int32), _ffi.ptr(uv, C.c_double), 0, C.byref(c0), C.byref(c1))
    assert rc == 0, rc
    rms = np.sqrt(2.0 * c1.value / len(obs))
    return Calibration(T_i_c, [model, model], intr, [752, 752], [480, 480]), c0.value, c1.value, rms, len(obs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ds")
    ap.add_argument("--out", default=os.path.join(ROOT, "tools", "euroc", "opt_calib.json"))
    a = ap.parse_args()
    cal, c0, c1, rms, n = calibrate(a.model)
    save_calibration(a.out, cal)
    print("calibration (%s): %d corners, cost %.4e -> %.4e, rms reprojection error %.3f px" % (a.model, n, c0, c1, rms))
    print("intrinsics:\n", cal.intrinsics)
    print("T_i_c[1] (qx qy qz qw tx ty tz):", cal.T_i_c[1], " baseline %.4f m" % np.linalg.norm(cal.T_i_c[1][4:]))
    print("wrote", a.out)
