"""Test-only ctypes bindings for the checkers under oracle/.

  * oracle/libpba_oracle.so  — the CPU restatement ("port"), always built.
  * oracle/_ref/libpba_ref.so — the unmodified reference + vendored Ceres 2.0.0
    (built here from /root/reference by oracle/ref/Makefile; travels to the GPU
    box prebuilt, may be absent).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

import pba_b200
from pba_b200 import _ffi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(_ROOT, "oracle", "libpba_oracle.so")


def _pick_ref():
    """AVX-512 build when the host supports it (so the CPU baseline is not handicapped), else AVX2."""
    base = os.path.join(_ROOT, "oracle", "_ref")
    v4 = os.path.join(base, "libpba_ref_v4.so")
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    if os.path.exists(v4) and all(f in flags for f in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd")):
        return v4
    return os.path.join(base, "libpba_ref.so")


REF_SO = _pick_ref()

_d = _ffi.c_double_p


def _bind_common(lib, prefix):
    P = C.POINTER(_ffi.pba_problem)
    getattr(lib, prefix + "_eval").argtypes = [P, C.c_int, C.c_double, C.c_int, _d, _d, _d]
    getattr(lib, prefix + "_project").restype = C.c_int
    getattr(lib, prefix + "_unproject").argtypes = [C.c_int, _d, C.c_int64, _d, _d]
    getattr(lib, prefix + "_se3_plus").argtypes = [C.c_int64, _d, _d, _d]
    getattr(lib, prefix + "_se3_plus_jacobian").argtypes = [_d, _d]


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(ORACLE_SO)
        _bind_common(lib, "pba_oracle")
        lib.pba_oracle_project.argtypes = [C.c_int, _d, C.c_int64, _d, _d, _d]
        lib.pba_oracle_solve.argtypes = [C.POINTER(_ffi.pba_problem), C.POINTER(_ffi.pba_options), C.c_int,
                                         C.POINTER(_ffi.pba_summary)]
        lib.pba_oracle_build_rcs.argtypes = [C.POINTER(_ffi.pba_problem), C.c_int, C.c_double, C.c_double, C.c_int,
                                             _ffi.c_i32_p, _d, _d, _d]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        _bind_common(lib, "pba_ref")
        lib.pba_ref_project.argtypes = [C.c_int, _d, C.c_int64, _d, _d]
        lib.pba_ref_project_jacobian.argtypes = [C.c_int, _d, C.c_int64, _d, _d]
        lib.pba_ref_solve.argtypes = [C.POINTER(_ffi.pba_problem), C.POINTER(_ffi.pba_options), C.c_int, C.c_int,
                                      C.POINTER(_ffi.pba_summary)]
        lib.pba_ref_hardware_threads.restype = C.c_int
        _ref = lib
    return _ref


def default_options(**kw):
    """pba_options with BundleAdjustmentOptions + Ceres defaults (no CUDA library needed)."""
    o = _ffi.pba_options()
    o.verbosity_level, o.optimize_intrinsics, o.use_huber, o.huber_parameter, o.max_num_iterations = 0, 0, 1, 1.0, 20
    o.solver, o.cholesky_max_dim, o.pcg_max_iterations, o.pcg_tolerance = 0, 16384, 500, 1e-10
    o.initial_trust_region_radius, o.max_trust_region_radius, o.min_trust_region_radius = 1e4, 1e16, 1e-32
    o.min_relative_decrease, o.min_lm_diagonal, o.max_lm_diagonal = 1e-3, 1e-6, 1e32
    o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance = 1e-6, 1e-10, 1e-8
    o.max_num_consecutive_invalid_steps, o.jacobi_scaling, o.device, o.profile, o.num_gpus = 5, 1, 0, 0, 1
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# The vendored Ceres 2.0.0 counts Jacobian non-zeros in an `int` (internal/ceres/block_sparse_matrix.cc:80,
# "Check failed: num_nonzeros_ >= 0"): a problem with more than 2^31 - 1 Jacobian entries aborts.  BASELINE
# config 4 (17,959,243 photometric blocks x 8 x 15 = 2.155e9 entries) is 0.36 % over that limit, so the
# reference itself can only run a prefix of it.
CERES_MAX_NONZEROS = 2 ** 31 - 1


def largest_reference_prefix(prob):
    """(n_kf, problem): the longest keyframe prefix of `prob` whose Jacobian the reference's Ceres can hold."""
    per_block = prob.res_per_obs * prob.cols_per_obs
    if prob.n_obs * per_block <= CERES_MAX_NONZEROS:
        return prob.n_poses, prob
    lo, hi = 2, prob.n_poses
    while hi - lo > 1:  # observations grow monotonically with the prefix length
        mid = (lo + hi) // 2
        last = np.searchsorted(prob.lm_host, mid, side="left")  # landmarks hosted before `mid` (upper bound)
        if int(prob.lm_obs_ptr[last]) * per_block <= CERES_MAX_NONZEROS:
            lo = mid
        else:
            hi = mid
    return lo, prob.prefix_keyframes(lo)


def evaluate(lib_kind, prob, use_huber=True, huber=1.0, threads=0, jac=True):
    """(cost, residuals [n_obs,R], jacobians [n_obs,R,C]) from the oracle port or the reference."""
    lib = oracle() if lib_kind == "oracle" else ref()
    fn = lib.pba_oracle_eval if lib_kind == "oracle" else lib.pba_ref_eval
    r = np.zeros((prob.n_obs, prob.res_per_obs))
    J = np.zeros((prob.n_obs, prob.res_per_obs, prob.cols_per_obs)) if jac else None
    cost = C.c_double()
    pc = prob.c
    rc = fn(C.byref(pc), int(use_huber), float(huber), threads if threads else (os.cpu_count() or 1),
            _ffi.ptr(r, C.c_double), _ffi.ptr(J, C.c_double), C.byref(cost))
    assert rc == 0, rc
    return cost.value, r, J


def solve(lib_kind, prob, opts=None, threads=0, use_reference_entry=False):
    """Runs the checker's full solve IN PLACE on prob; returns a pba_b200.Summary."""
    opts = opts or default_options()
    s = pba_b200.Summary(max(64, opts.max_num_iterations))
    pc = prob.c
    threads = threads if threads else (os.cpu_count() or 1)
    if lib_kind == "oracle":
        rc = oracle().pba_oracle_solve(C.byref(pc), C.byref(opts), threads, C.byref(s.c))
    else:
        rc = ref().pba_ref_solve(C.byref(pc), C.byref(opts), threads, int(use_reference_entry), C.byref(s.c))
    assert rc == 0, rc
    return s


def build_rcs(prob, use_huber=True, huber=1.0, radius=1e4, threads=0):
    lib = oracle()
    pc = prob.c
    dim = C.c_int32()
    threads = threads if threads else (os.cpu_count() or 1)
    lib.pba_oracle_build_rcs(C.byref(pc), int(use_huber), float(huber), float(radius), threads, C.byref(dim),
                             None, None, None)
    S = np.zeros((dim.value, dim.value))
    rhs = np.zeros(dim.value)
    scale = np.zeros(dim.value)
    lib.pba_oracle_build_rcs(C.byref(pc), int(use_huber), float(huber), float(radius), threads, C.byref(dim),
                             _ffi.ptr(S, C.c_double), _ffi.ptr(rhs, C.c_double), _ffi.ptr(scale, C.c_double))
    return S, rhs, scale


def landmark_positions(lib_kind, prob):
    """Landmark::get_p for every landmark from the oracle port or the reference's own method."""
    lib = oracle() if lib_kind == "oracle" else ref()
    fn = lib.pba_oracle_landmark_positions if lib_kind == "oracle" else lib.pba_ref_landmark_positions
    fn.argtypes = [C.POINTER(_ffi.pba_problem), _d]
    out = np.zeros((prob.n_landmarks, 3))
    pc = prob.c
    rc = fn(C.byref(pc), _ffi.ptr(out, C.c_double))
    assert rc == 0, rc
    return out


def compute_projections(lib_kind, prob, thresholds=None):
    """compute_projections + set_outlier_flags (+ removal decision for the oracle port).
    Returns a dict of flat arrays in the slot layout of include/pba.h."""
    thresholds = thresholds or pba_b200.ProjectionThresholds()
    t = thresholds.to_c()
    nl, ns = prob.n_landmarks, prob.n_obs + prob.n_landmarks
    out = dict(point_reprojected=np.zeros((ns, 2)), point_3d_c=np.zeros((ns, 3)), reprojection_error=np.zeros(ns),
               outlier_flags=np.zeros(ns, np.uint32))
    pc = prob.c
    P, T = C.POINTER(_ffi.pba_problem), C.POINTER(_ffi.pba_projection_thresholds)
    if lib_kind == "oracle":
        fn = oracle().pba_oracle_compute_projections
        fn.argtypes = [P, T, _d, _d, _d, _ffi.c_u32_p, _ffi.c_u8_p, _ffi.c_i32_p]
        out["landmark_remove"] = np.zeros(nl, np.uint8)
        severe = C.c_int32(0)
        rc = fn(C.byref(pc), C.byref(t), _ffi.ptr(out["point_reprojected"], C.c_double),
                _ffi.ptr(out["point_3d_c"], C.c_double), _ffi.ptr(out["reprojection_error"], C.c_double),
                _ffi.ptr(out["outlier_flags"], C.c_uint32), _ffi.ptr(out["landmark_remove"], C.c_uint8),
                C.byref(severe))
        out["any_severe_outliers"] = bool(severe.value)
    else:
        fn = ref().pba_ref_compute_projections
        fn.argtypes = [P, T, _d, _d, _d, _ffi.c_u32_p]
        rc = fn(C.byref(pc), C.byref(t), _ffi.ptr(out["point_reprojected"], C.c_double),
                _ffi.ptr(out["point_3d_c"], C.c_double), _ffi.ptr(out["reprojection_error"], C.c_double),
                _ffi.ptr(out["outlier_flags"], C.c_uint32))
    assert rc == 0, rc
    return out


def triangulate(lib_kind, model0, intr0, model1, intr1, T_w_c0, T_w_c1, uv0, uv1):
    """add_new_landmarks_between_cams: (p in camera 0's frame [n,3], inverse distance [n]) from the oracle port or
    from the reference's own function (+ opengv's triangulate for p)."""
    lib = oracle() if lib_kind == "oracle" else ref()
    fn = lib.pba_oracle_triangulate if lib_kind == "oracle" else lib.pba_ref_add_new_landmarks
    fn.argtypes = [C.c_int, _d, C.c_int, _d, _d, _d, C.c_int64, _d, _d, _d, _d]
    fn.restype = C.c_int
    i0, i1 = np.ascontiguousarray(intr0, np.float64), np.ascontiguousarray(intr1, np.float64)
    T0, T1 = np.ascontiguousarray(T_w_c0, np.float64), np.ascontiguousarray(T_w_c1, np.float64)
    uv0, uv1 = np.ascontiguousarray(uv0, np.float64).reshape(-1, 2), np.ascontiguousarray(uv1, np.float64).reshape(-1, 2)
    n = uv0.shape[0]
    p, rho = np.zeros((n, 3)), np.zeros(n)
    rc = fn(int(model0), _ffi.ptr(i0, C.c_double), int(model1), _ffi.ptr(i1, C.c_double), _ffi.ptr(T0, C.c_double),
            _ffi.ptr(T1, C.c_double), n, _ffi.ptr(uv0, C.c_double), _ffi.ptr(uv1, C.c_double), _ffi.ptr(p, C.c_double),
            _ffi.ptr(rho, C.c_double))
    assert rc == 0, rc
    return p, rho


# ---- front-end checkers: oracle restatement and the reference's own functions ----
REF_FRONTEND_SO = os.path.join(_ROOT, "oracle", "_ref", "libpba_ref_frontend.so")
_ref_fe = None


def have_ref_frontend():
    return os.path.exists(REF_FRONTEND_SO)


def _frontend_lib(kind):
    global _ref_fe
    if kind == "oracle":
        return oracle(), "pba_oracle"
    if _ref_fe is None:
        _ref_fe = C.CDLL(REF_FRONTEND_SO)
    return _ref_fe, "pba_ref"


def corner_descriptors(kind, image, corners, rotate_features=True):
    """computeAngles + computeDescriptors of one image: (angles [n], descriptors [n, 32])."""
    lib, pre = _frontend_lib(kind)
    image = np.ascontiguousarray(image, np.uint8)
    corners = np.ascontiguousarray(corners, np.float64).reshape(-1, 2)
    h, w = image.shape
    n = len(corners)
    ang = np.zeros(n)
    desc = np.zeros((n, 32), np.uint8)
    f = getattr(lib, pre + "_corner_descriptors")
    f.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, _d, C.c_int, _d, C.POINTER(C.c_uint8)]
    f(_ffi.ptr(image, C.c_uint8), w, h, w, n, _ffi.ptr(corners, C.c_double), int(bool(rotate_features)),
      _ffi.ptr(ang, C.c_double), _ffi.ptr(desc, C.c_uint8))
    return ang, desc


def match_descriptors(kind, d1, d2, threshold=70, dist_2_best=1.2):
    """matchDescriptors: [q, 2] int32, ascending in the first index."""
    lib, pre = _frontend_lib(kind)
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    out = np.zeros((max(min(len(d1), len(d2)), 1), 2), np.int32)
    f = getattr(lib, pre + "_match_descriptors")
    f.argtypes = [C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_double, C.POINTER(C.c_int32)]
    f.restype = C.c_int
    n = f(len(d1), _ffi.ptr(d1, C.c_uint8), len(d2), _ffi.ptr(d2, C.c_uint8), int(threshold), float(dist_2_best),
          _ffi.ptr(out, C.c_int32))
    return out[:n].copy()


def epipolar_inliers(kind, model0, intr0, model1, intr1, T_0_1, matches, corners0, corners1, threshold=1e-3):
    """(E [3, 3], inlier mask) from computeEssential + findInliersEssential."""
    lib, pre = _frontend_lib(kind)
    i0 = np.ascontiguousarray(intr0, np.float64).reshape(8)
    i1 = np.ascontiguousarray(intr1, np.float64).reshape(8)
    T = np.ascontiguousarray(T_0_1, np.float64).reshape(7)
    matches = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    c0 = np.ascontiguousarray(corners0, np.float64).reshape(-1, 2)
    c1 = np.ascontiguousarray(corners1, np.float64).reshape(-1, 2)
    E = np.zeros(9)
    inl = np.zeros(max(len(matches), 1), np.uint8)
    u8 = C.POINTER(C.c_uint8)
    if kind == "oracle":
        f = lib.pba_oracle_epipolar_inliers
        f.argtypes = [C.c_int, _d, C.c_int, _d, _d, C.c_double, C.c_int64, C.POINTER(C.c_int32), _d, _d, _d, u8]
        f(int(model0), _ffi.ptr(i0, C.c_double), int(model1), _ffi.ptr(i1, C.c_double), _ffi.ptr(T, C.c_double),
          float(threshold), len(matches), _ffi.ptr(matches, C.c_int32), _ffi.ptr(c0, C.c_double), _ffi.ptr(c1, C.c_double),
          _ffi.ptr(E, C.c_double), _ffi.ptr(inl, C.c_uint8))
    else:
        lib.pba_ref_compute_essential.argtypes = [_d, _d]
        lib.pba_ref_compute_essential(_ffi.ptr(T, C.c_double), _ffi.ptr(E, C.c_double))
        f = lib.pba_ref_epipolar_inliers
        f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, _d, C.c_int, _d, C.c_int, _d, C.c_int, _d, _d, C.c_double, u8]
        f(len(matches), _ffi.ptr(matches, C.c_int32), len(c0), _ffi.ptr(c0, C.c_double), len(c1), _ffi.ptr(c1, C.c_double),
          int(model0), _ffi.ptr(i0, C.c_double), int(model1), _ffi.ptr(i1, C.c_double), _ffi.ptr(E, C.c_double),
          float(threshold), _ffi.ptr(inl, C.c_uint8))
    return E.reshape(3, 3), inl[:len(matches)].astype(bool)


def build_tracks(kind, feature_counts, pairs, matches, min_length=3):
    """(track_of per image, number of tracks): TrackBuilder Build + Filter + Export, canonical ids (smallest node)."""
    lib, pre = _frontend_lib(kind)
    counts = np.asarray(feature_counts, np.int64)
    fp = np.zeros(len(counts) + 1, np.int32)
    fp[1:] = np.cumsum(counts)
    pairs = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
    mp = np.zeros(len(pairs) + 1, np.int64)
    mp[1:] = np.cumsum([len(m) for m in matches])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32).reshape(-1, 2) for m in matches] + [np.zeros((1, 2), np.int32)]))
    out = np.full(max(int(fp[-1]), 1), -1, np.int32)
    f = getattr(lib, pre + "_build_tracks")
    f.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                  C.c_int, C.POINTER(C.c_int32)]
    f.restype = C.c_int
    nt = f(len(counts), _ffi.ptr(fp, C.c_int32), len(pairs), _ffi.ptr(pairs, C.c_int32), _ffi.ptr(mp, C.c_int64),
           _ffi.ptr(flat, C.c_int32), int(min_length), _ffi.ptr(out, C.c_int32))
    return [out[fp[i]:fp[i + 1]].copy() for i in range(len(counts))], int(nt)
