"""Generates tests/golden/triangulate_<model>.npz from the REAL reference: the reference's own
add_new_landmarks_between_cams() (include/visnav/map_utils.h:121-195) + the vendored opengv triangulation, driven
by oracle/_ref (built by oracle/ref/Makefile).  Inputs: a stereo-like camera pair (different intrinsics per camera,
the reference's getTestProjections() values), 3D points in front of both, their corner pixels with 0.3 px noise.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402
from pba_b200 import _ffi  # noqa: E402
import ctypes as C  # noqa: E402

INTR = {
    "pinhole": [0.5 * 805, 0.5 * 800, 505, 509, 0, 0, 0, 0],
    "eucm": [0.5 * 500, 0.5 * 500, 319.5, 239.5, 0.51231234, 0.9, 0, 0],
    "ds": [0.5 * 805, 0.5 * 800, 505, 509, 0.5 * -0.150694, 0.5 * 1.48785, 0, 0],
    "kb4": [379.045, 379.008, 505.512, 509.969, 0.00693023, -0.0013828, -0.000272596, -0.000452646],
}


def make_inputs(model, n=257, seed=5):
    rng = np.random.default_rng(seed)
    mid = _ffi.CAM_NAMES[model]
    intr0 = np.array(INTR[model], np.float64)
    intr1 = intr0.copy()
    intr1[:4] *= [1.01, 0.99, 1.002, 0.997]
    # camera 0 somewhere, camera 1 = a 11 cm stereo baseline with a small rotation
    T0 = np.array([0.02, -0.01, 0.03, 0.0, 0.3, -0.2, 0.1]); T0[3] = np.sqrt(1 - (T0[:3] ** 2).sum())
    d = np.array([0.11, 0.004, -0.002, 0.004, -0.006, 0.003])
    T1 = np.zeros(7)
    of.oracle().pba_oracle_se3_plus(1, _ffi.ptr(T0, C.c_double), _ffi.ptr(d, C.c_double), _ffi.ptr(T1, C.c_double))
    Xc0 = np.c_[rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(1.5, 8.0, n)]
    # world points and their projections through the oracle's camera model
    def rot(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    Xw = Xc0 @ rot(T0[:4]).T + T0[4:]
    Xc1 = (Xw - T1[4:]) @ rot(T1[:4])
    uv0, uv1 = np.zeros((n, 2)), np.zeros((n, 2))
    of.oracle().pba_oracle_project(mid, _ffi.ptr(intr0, C.c_double), n, _ffi.ptr(np.ascontiguousarray(Xc0), C.c_double),
                                   _ffi.ptr(uv0, C.c_double), None)
    of.oracle().pba_oracle_project(mid, _ffi.ptr(intr1, C.c_double), n, _ffi.ptr(np.ascontiguousarray(Xc1), C.c_double),
                                   _ffi.ptr(uv1, C.c_double), None)
    uv0 += rng.normal(0, 0.3, uv0.shape)
    uv1 += rng.normal(0, 0.3, uv1.shape)
    return mid, intr0, intr1, T0, T1, uv0, uv1, Xc0


def main():
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    for model in INTR:
        mid, intr0, intr1, T0, T1, uv0, uv1, Xc0 = make_inputs(model)
        p, rho = of.triangulate("ref", mid, intr0, mid, intr1, T0, T1, uv0, uv1)
        np.savez_compressed(os.path.join(HERE, "triangulate_%s.npz" % model), model=mid, intr0=intr0, intr1=intr1,
                            T_w_c0=T0, T_w_c1=T1, uv0=uv0, uv1=uv1, ref_p_c0=p, ref_inv_depth=rho, true_p_c0=Xc0)
        print(model, "max |p - truth| %.3e" % np.abs(p - Xc0).max(), "rho range", rho.min(), rho.max())


if __name__ == "__main__":
    main()
