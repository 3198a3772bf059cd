"""ctypes mirror of include/pba.h (the C ABI) and of the synth generator's structs.

Only plain pointers and sizes cross the boundary; numpy arrays are passed as
host buffers.  No torch types appear here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PBA_OK = 0
STATUS_NAMES = {0: "PBA_OK", 1: "PBA_ERR_INVALID_ARGUMENT", 2: "PBA_ERR_NO_DEVICE", 3: "PBA_ERR_CUDA",
                4: "PBA_ERR_UNSUPPORTED", 5: "PBA_ERR_NUMERICAL_FAILURE", 6: "PBA_ERR_NCCL",
                7: "PBA_ERR_OUT_OF_MEMORY"}
MODE_GEOMETRIC, MODE_PHOTOMETRIC = 0, 1
CAM_PINHOLE, CAM_DS, CAM_KB4, CAM_EUCM = 0, 1, 2, 3
CAM_NAMES = {"pinhole": CAM_PINHOLE, "ds": CAM_DS, "kb4": CAM_KB4, "eucm": CAM_EUCM}
SOLVER_AUTO, SOLVER_CHOLESKY, SOLVER_PCG, SOLVER_BAND, SOLVER_BCR = 0, 1, 2, 3, 4
CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2
NCCL_ID_BYTES = 128

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u32_p = C.POINTER(C.c_uint32)


class pba_problem(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("n_poses", C.c_int32), ("n_calib", C.c_int32), ("n_landmarks", C.c_int32),
        ("n_obs", C.c_int64),
        ("poses", c_double_p), ("pose_fixed", c_u8_p), ("pose_calib", c_i32_p), ("calib_model", c_i32_p),
        ("intrinsics", c_double_p),
        ("inv_depth", c_double_p), ("lm_host", c_i32_p), ("lm_host_uv", c_double_p), ("lm_obs_ptr", c_i64_p),
        ("obs_target", c_i32_p), ("obs_uv", c_double_p),
        ("images", c_u8_p), ("image_ptrs", C.POINTER(c_u8_p)), ("image_stride", C.c_int64),
        ("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32),
        ("affine", c_double_p),
    ]


class pba_options(C.Structure):
    _fields_ = [
        ("verbosity_level", C.c_int32), ("optimize_intrinsics", C.c_int32), ("use_huber", C.c_int32),
        ("huber_parameter", C.c_double), ("max_num_iterations", C.c_int32),
        ("solver", C.c_int32), ("cholesky_max_dim", C.c_int32), ("pcg_max_iterations", C.c_int32),
        ("pcg_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
        ("device", C.c_int32), ("profile", C.c_int32), ("num_gpus", C.c_int32), ("reserved0_", C.c_int32),
    ]


class pba_iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
        ("linear_solver_iterations", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
        ("gradient_norm", C.c_double), ("step_norm", C.c_double), ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double), ("model_cost_change", C.c_double),
        ("iteration_time_in_seconds", C.c_double), ("cumulative_time_in_seconds", C.c_double),
    ]


class pba_summary(C.Structure):
    _fields_ = [
        ("termination_type", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32), ("num_residual_evaluations", C.c_int32),
        ("num_jacobian_evaluations", C.c_int32), ("num_linear_solves", C.c_int32), ("rcs_dim", C.c_int32),
        ("linear_solver", C.c_int32), ("num_inexact_linear_solves", C.c_int32),
        ("rcs_blocks", C.c_int64), ("num_residual_blocks", C.c_int64), ("num_residuals", C.c_int64),
        ("num_effective_parameters", C.c_int64), ("gpu_kernel_launches", C.c_int64),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("setup_time_in_seconds", C.c_double),
        ("residual_evaluation_time_in_seconds", C.c_double), ("jacobian_evaluation_time_in_seconds", C.c_double),
        ("linear_solver_time_in_seconds", C.c_double), ("minimizer_time_in_seconds", C.c_double),
        ("total_time_in_seconds", C.c_double),
        ("iterations", C.POINTER(pba_iteration)), ("iterations_capacity", C.c_int32),
        ("message", C.c_char * 256),
    ]


class pba_projection_thresholds(C.Structure):
    _fields_ = [("reprojection_error_huge_pixel", C.c_double), ("reprojection_error_normal_pixel", C.c_double),
                ("camera_center_distance_meter", C.c_double), ("z_coordinate_meter", C.c_double)]


class pba_kernel_stat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double)]


class pba_synth_params(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("n_kf", C.c_int32), ("n_pts", C.c_int32), ("model", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("min_len", C.c_int32), ("max_len", C.c_int32),
        ("seed_pix", C.c_uint64), ("seed_vis", C.c_uint64), ("seed_noise", C.c_uint64),
        ("pose_sigma", C.c_double), ("rho_sigma", C.c_double), ("pixel_sigma", C.c_double),
        ("affine_a_sigma", C.c_double), ("affine_b_sigma", C.c_double),
        ("intrinsics", C.c_double * 8),
    ]


def ptr(a, ctype):
    """Pointer to a C-contiguous numpy array (or NULL for None)."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


# Every symbol include/pba.h declares (checked by tests/test_abi.py).
PBA_SYMBOLS = [
    "pba_abi_version", "pba_status_string", "pba_device_count", "pba_options_init", "pba_solve", "pba_multi_gpu_init", "pba_analyze_structure", "pba_collective_kind",
    "pba_create", "pba_destroy", "pba_trim_device_cache", "pba_set_stream", "pba_synchronize", "pba_evaluate", "pba_get_residuals",
    "pba_get_jacobians", "pba_get_blocks", "pba_build_rcs", "pba_get_rcs_dim", "pba_get_rcs", "pba_solve_rcs", "pba_minimize",
    "pba_lm_iterate",
    "pba_set_state", "pba_get_state", "pba_get_sizes", "pba_reset_kernel_stats", "pba_set_profile", "pba_get_kernel_stats",
    "pba_nccl_unique_id", "pba_comm_init", "pba_camera_project", "pba_camera_unproject", "pba_se3_plus",
    "pba_cholesky_solve", "pba_projection_thresholds_init", "pba_landmark_positions", "pba_compute_projections", "pba_triangulate_inverse_depth",
    "pba_corner_descriptors", "pba_match_descriptors", "pba_epipolar_inliers", "pba_build_tracks",
]

_lib = None
_synth = None


class ExtensionMissing(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libpba_b200.so")


def load_lib():
    """Load the CUDA engine.  Fails loudly: there is no fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ExtensionMissing(
            "%s not built — run `make lib` (or __graft_entry__.build()). "
            "The B200 engine has no CPU or PyTorch fallback." % path)
    lib = C.CDLL(path)
    lib.pba_abi_version.restype = C.c_int32
    lib.pba_status_string.restype = C.c_char_p
    lib.pba_status_string.argtypes = [C.c_int]
    lib.pba_device_count.restype = C.c_int32
    lib.pba_options_init.argtypes = [C.POINTER(pba_options)]
    lib.pba_options_init.restype = None
    H = C.c_void_p
    sigs = {
        "pba_solve": [C.POINTER(pba_problem), C.POINTER(pba_options), C.POINTER(pba_summary)],
        "pba_multi_gpu_init": [C.c_int32, C.c_int32],
        "pba_collective_kind": [C.c_void_p],
        "pba_analyze_structure": [C.POINTER(pba_problem), C.POINTER(pba_options), c_i32_p, c_i32_p, c_i32_p, c_i32_p,
                                  c_i64_p],
        "pba_create": [C.POINTER(pba_problem), C.POINTER(pba_options), C.c_int32, C.c_int32, C.POINTER(H)],
        "pba_set_stream": [H, C.c_void_p],
        "pba_synchronize": [H],
        "pba_evaluate": [H, C.c_int32, c_double_p],
        "pba_get_residuals": [H, c_double_p],
        "pba_get_jacobians": [H, c_double_p],
        "pba_get_blocks": [H, C.c_int64, c_i64_p, c_double_p, c_double_p],
        "pba_build_rcs": [H, C.c_double],
        "pba_get_rcs_dim": [H, c_i32_p],
        "pba_get_rcs": [H, c_double_p, c_double_p],
        "pba_solve_rcs": [H, C.c_int32, c_double_p, c_i32_p],
        "pba_minimize": [H, C.POINTER(pba_summary)],
        "pba_lm_iterate": [H, C.c_double, C.c_int32, C.POINTER(pba_iteration)],
        "pba_set_state": [H, c_double_p, c_double_p, c_double_p],
        "pba_get_state": [H, c_double_p, c_double_p, c_double_p],
        "pba_get_sizes": [H, c_i64_p, c_i32_p, c_i64_p],
        "pba_reset_kernel_stats": [H],
        "pba_set_profile": [H, C.c_int32],
        "pba_nccl_unique_id": [c_u8_p],
        "pba_comm_init": [H, c_u8_p],
        "pba_camera_project": [C.c_int32, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p],
        "pba_camera_unproject": [C.c_int32, c_double_p, C.c_int64, c_double_p, c_double_p],
        "pba_se3_plus": [C.c_int64, c_double_p, c_double_p, c_double_p],
        "pba_cholesky_solve": [C.c_int32, c_double_p, c_double_p, c_double_p],
        "pba_landmark_positions": [C.POINTER(pba_problem), C.c_int32, c_double_p],
        "pba_triangulate_inverse_depth": [C.c_int32, c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p, C.c_int64,
                                          c_double_p, c_double_p, C.c_int32, c_double_p, c_double_p],
        "pba_compute_projections": [C.POINTER(pba_problem), C.POINTER(pba_projection_thresholds), C.c_int32,
                                    c_double_p, c_double_p, c_double_p, c_u32_p, c_u8_p, c_i32_p],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.pba_projection_thresholds_init.argtypes = [C.POINTER(pba_projection_thresholds)]
    lib.pba_projection_thresholds_init.restype = None
    lib.pba_destroy.argtypes = [H]
    lib.pba_destroy.restype = None
    lib.pba_trim_device_cache.argtypes = []
    lib.pba_trim_device_cache.restype = None
    lib.pba_get_kernel_stats.argtypes = [H, C.POINTER(pba_kernel_stat), C.c_int32]
    lib.pba_get_kernel_stats.restype = C.c_int32
    lib.pba_synth_render_gpu.argtypes = [C.POINTER(pba_synth_params), C.c_int, C.c_int, C.c_int, c_u8_p]
    lib.pba_synth_render_gpu.restype = C.c_int
    _lib = lib
    return lib


def load_synth():
    global _synth
    if _synth is not None:
        return _synth
    path = os.path.join(_HERE, "libpba_synth.so")
    if not os.path.exists(path):
        raise ExtensionMissing("%s not built — run `make synth`" % path)
    s = C.CDLL(path)
    s.pba_synth_default_params.argtypes = [C.POINTER(pba_synth_params), C.c_int, C.c_int, C.c_int, C.c_int]
    s.pba_synth_default_params.restype = None
    s.pba_synth_count_obs.argtypes = [C.POINTER(pba_synth_params)]
    s.pba_synth_count_obs.restype = C.c_int64
    s.pba_synth_generate.argtypes = [C.POINTER(pba_synth_params), c_double_p, c_double_p, c_u8_p, c_double_p,
                                     c_double_p, c_i32_p, c_double_p, c_i64_p, c_i32_p, c_double_p, c_double_p]
    s.pba_synth_generate.restype = C.c_int
    s.pba_synth_render.argtypes = [C.POINTER(pba_synth_params), C.c_int, C.c_int, C.c_int, c_u8_p]
    s.pba_synth_render.restype = C.c_int
    _synth = s
    return s


def check(status, what=""):
    if status != PBA_OK:
        raise RuntimeError("%s failed: %s" % (what or "pba call", STATUS_NAMES.get(status, status)))
