"""Load a tests/golden/*.npz fixture (made by tests/golden/make_golden.py from the
real reference) back into a pba_b200.Problem."""
import glob
import os

import numpy as np

import pba_b200 as pb

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "geom_*.npz")) + glob.glob(os.path.join(GOLDEN_DIR, "photo_*.npz")))


def projection_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "projections_*.npz")))


def load(path):
    g = np.load(path)
    mode = int(g["mode"])
    prob = pb.Problem(mode, g["poses"].copy(), g["pose_fixed"], g["pose_calib"], g["calib_model"], g["intrinsics"],
                      g["inv_depth"].copy(), g["lm_host"], g["lm_host_uv"], g["lm_obs_ptr"], g["obs_target"],
                      g["obs_uv"] if "obs_uv" in g else None, g["images"] if "images" in g else None,
                      g["affine"].copy() if "affine" in g else None)
    return prob, g
