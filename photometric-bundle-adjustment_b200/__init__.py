"""photometric-bundle-adjustment_b200 — B200-native bundle adjustment behind the
reference's `bundle_adjustment()` boundary (include/visnav/map_utils.h:322).

Python here is only the host-side mirror of that interface over the C ABI
(include/pba.h, libpba_b200.so).  All compute is hand-written CUDA for sm_100a;
there is no CPU / PyTorch fallback — calls raise if the extension is missing or
no GPU is present.
"""
from . import _ffi
from ._ffi import (CAM_DS, CAM_EUCM, CAM_KB4, CAM_PINHOLE, CONVERGENCE, FAILURE, MODE_GEOMETRIC,
                   MODE_PHOTOMETRIC, NO_CONVERGENCE, SOLVER_AUTO, SOLVER_BAND, SOLVER_BCR, SOLVER_CHOLESKY, SOLVER_PCG, ExtensionMissing)
from .calibration import Calibration, initialize_from_double_sphere, load_calibration, save_calibration
from .engine import BundleAdjustmentOptions, Engine, Summary, analyze_structure, bundle_adjustment, device_count, multi_gpu_init
from .frontend import build_tracks, corner_descriptors, epipolar_inliers, match_descriptors
from .map_io import Map, load_map_file, save_map_file
from .problem import Problem, partition_landmarks
from .projections import (ProjectionThresholds, Projections, compute_projections, landmark_positions,
                          triangulate_inverse_depth)
from .synth import make_grid_scene, make_scene

__all__ = [
    "Calibration", "load_calibration", "save_calibration", "initialize_from_double_sphere",
    "BundleAdjustmentOptions", "Engine", "Summary", "bundle_adjustment", "device_count", "multi_gpu_init", "analyze_structure", "triangulate_inverse_depth", "Problem",
    "partition_landmarks", "make_scene", "make_grid_scene", "MODE_GEOMETRIC", "MODE_PHOTOMETRIC", "CAM_PINHOLE", "CAM_DS",
    "CAM_KB4", "CAM_EUCM", "SOLVER_AUTO", "SOLVER_CHOLESKY", "SOLVER_PCG", "SOLVER_BAND", "SOLVER_BCR", "CONVERGENCE", "NO_CONVERGENCE",
    "FAILURE", "ExtensionMissing", "ProjectionThresholds", "Projections", "compute_projections",
    "landmark_positions", "corner_descriptors", "match_descriptors", "epipolar_inliers", "build_tracks", "Map", "load_map_file", "save_map_file",
]
