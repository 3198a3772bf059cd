/* pba.h — C ABI of the B200-native bundle-adjustment engine (libpba_b200.so).
 *
 * This is the drop-in boundary for the reference's single hot entry point
 *
 *   void visnav::bundle_adjustment(const Corners&, const BundleAdjustmentOptions&,
 *                                  const std::set<FrameCamId>& fixed_cameras,
 *                                  Calibration&, Cameras&, Landmarks&)
 *   (reference: include/visnav/map_utils.h:322-399, caller src/sfm.cpp:1903-1913)
 *
 * A host shim (include/visnav_b200/bundle_adjustment.h)
 * flattens the reference containers into the SoA `pba_problem` below, calls
 * `pba_solve`, and writes poses / inverse distances back in place — the same
 * in/out contract Ceres has through raw parameter pointers (map_utils.h:331,
 * :373).  Only plain pointers and sizes cross this boundary; no C++/torch
 * types.  Every function returns a pba_status; nothing throws.
 *
 * There is no CPU fallback: every compute entry point returns
 * PBA_ERR_NO_DEVICE when no CUDA device is present.
 */
#ifndef PBA_H_
#define PBA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBA_ABI_VERSION 2

typedef enum pba_status {
  PBA_OK = 0,
  PBA_ERR_INVALID_ARGUMENT = 1,
  PBA_ERR_NO_DEVICE = 2,
  PBA_ERR_CUDA = 3,
  PBA_ERR_UNSUPPORTED = 4, /* e.g. optimize_intrinsics=true (map_utils.h:339) */
  PBA_ERR_NUMERICAL_FAILURE = 5,
  PBA_ERR_NCCL = 6,
  PBA_ERR_OUT_OF_MEMORY = 7
} pba_status;

/* Residual family.  GEOMETRIC is the reference's
 * BundleAdjustmentReprojectionCostFunctor (reprojection.h:74-118): 2 residuals,
 * local Jacobian 2 x (6 host pose | 6 target pose | 1 inverse distance).
 * PHOTOMETRIC is SURVEY.md §8(a-P): 8 residuals (DSO pattern), local Jacobian
 * 8 x (6 host pose | 6 target pose | 2 target affine (a,b) | 1 inverse distance). */
enum { PBA_MODE_GEOMETRIC = 0, PBA_MODE_PHOTOMETRIC = 1 };

/* Camera models, reference camera_models.h:47-420; names as in
 * AbstractCamera::from_data (camera_models.h:452-474). */
enum { PBA_CAM_PINHOLE = 0, PBA_CAM_DS = 1, PBA_CAM_KB4 = 2, PBA_CAM_EUCM = 3 };

/* RCS solvers: dense tiled Cholesky (FP64 tensor cores), block-Jacobi PCG, and two exact
 * solvers for block-banded systems (windowed covisibility): BAND = sequential block-banded
 * Cholesky on one SM, BCR = parallel block cyclic reduction on dense super blocks.
 * AUTO = BCR when applicable (BAND for chains of <= 64 free keyframes), else BAND, else CHOLESKY while
 * dim <= cholesky_max_dim, else PCG.  The free cameras are renumbered by reverse Cuthill-McKee when their own
 * order is not banded (loop closures, unordered maps: pba_analyze_structure), so the banded solvers cover
 * maps whose keyframes arrive in any order; PCG is the last resort, and a PCG solve that stops above
 * pcg_tolerance counts as an invalid step (pba_summary.num_inexact_linear_solves). */
enum { PBA_SOLVER_AUTO = 0, PBA_SOLVER_CHOLESKY = 1, PBA_SOLVER_PCG = 2, PBA_SOLVER_BAND = 3, PBA_SOLVER_BCR = 4 };

/* Ceres termination types (include/ceres/types.h) kept so reports line up. */
enum { PBA_CONVERGENCE = 0, PBA_NO_CONVERGENCE = 1, PBA_FAILURE = 2 };

#define PBA_PATTERN_SIZE 8
#define PBA_GEOM_RES 2
#define PBA_GEOM_COLS 13  /* 6 + 6 + 1 */
#define PBA_PHOTO_RES 8
#define PBA_PHOTO_COLS 15 /* 6 + 6 + 2 + 1 */

/* Flat (SoA) bundle-adjustment problem.  All arrays are HOST memory owned by
 * the caller.  "pose index" = position of the keyframe in `poses`
 * (the reference's Cameras map iterates in FrameCamId order, common_types.h:87).
 * in/out arrays are updated in place on success, exactly like Ceres does
 * through T_w_c.data() / &landmark.inv_depth (map_utils.h:331,373). */
typedef struct pba_problem {
  int32_t mode;     /* PBA_MODE_* */
  int32_t n_poses;  /* keyframes / cameras in the map */
  int32_t n_calib;  /* intrinsics blocks (Calibration::intrinsics.size()) */
  int32_t n_landmarks;
  int64_t n_obs;    /* residual blocks = sum over landmarks of (|obs| - 1) */

  double* poses;              /* [n_poses*7] Sophus order qx qy qz qw tx ty tz; in/out */
  const uint8_t* pose_fixed;  /* [n_poses] 1 = constant (fixed_cameras) */
  const int32_t* pose_calib;  /* [n_poses] index into intrinsics (FrameCamId::cam_id) */
  const int32_t* calib_model; /* [n_calib] PBA_CAM_* */
  const double* intrinsics;   /* [n_calib*8] fx fy cx cy p1..p4 (camera_models.h:119-123) */

  double* inv_depth;          /* [n_landmarks] inverse distance along host ray; in/out */
  const int32_t* lm_host;     /* [n_landmarks] pose index of host = obs.begin() (map_utils.h:351) */
  const double* lm_host_uv;   /* [n_landmarks*2] host pixel z_h */
  const int64_t* lm_obs_ptr;  /* [n_landmarks+1] CSR offsets into obs_* (non-host obs only) */
  const int32_t* obs_target;  /* [n_obs] pose index of the target keyframe */
  const double* obs_uv;       /* [n_obs*2] target pixel z_t (GEOMETRIC only, else NULL) */

  /* PHOTOMETRIC only: one 8-bit grey image per pose, all the same size.
   * Either image_ptrs[i] (if non-NULL) or images + i*image_stride. */
  const uint8_t* images;
  const uint8_t* const* image_ptrs;
  int64_t image_stride;       /* bytes between consecutive images in `images` */
  int32_t width, height, pitch;
  double* affine;             /* [n_poses*2] (a,b) per keyframe; in/out */
} pba_problem;

/* First five fields = BundleAdjustmentOptions (map_utils.h:304-319), same
 * defaults via pba_options_init.  The rest default to Ceres 2.0.0's
 * Solver::Options values (include/ceres/solver.h:277-322). */
typedef struct pba_options {
  int32_t verbosity_level;     /* 0 silent, 1 brief report, 2 full report */
  int32_t optimize_intrinsics; /* must be 0 (reference marks it broken, map_utils.h:339) */
  int32_t use_huber;
  double huber_parameter;
  int32_t max_num_iterations;

  int32_t solver;              /* PBA_SOLVER_* (see above) */
  int32_t cholesky_max_dim;    /* default 16384 (2.1 GB of dense workspace, allocated on first use) */
  int32_t pcg_max_iterations;  /* default 500 */
  double pcg_tolerance;        /* relative residual ||r||/||b||, default 1e-10 */
  double initial_trust_region_radius; /* 1e4 */
  double max_trust_region_radius;     /* 1e16 */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3 */
  double min_lm_diagonal;             /* 1e-6 */
  double max_lm_diagonal;             /* 1e32 */
  double function_tolerance;          /* 1e-6 */
  double gradient_tolerance;          /* 1e-10 */
  double parameter_tolerance;         /* 1e-8 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;             /* 1 */
  int32_t device;                     /* CUDA device ordinal, default 0 */
  int32_t profile;                    /* 1 = bracket every kernel with CUDA events, 2 = only the residual/Jacobian kernel */
  /* pba_solve only: GPUs of this node to shard the landmarks over, FROM ONE PROCESS (the reference's caller is a
   * single thread, src/sfm.cpp:1883-1925): devices device .. device + num_gpus - 1, one host thread each, partial
   * reduced camera systems summed with NCCL (ncclCommInitAll, communicators cached per process).  Default 1;
   * 0 = all visible devices. */
  int32_t num_gpus;
  int32_t reserved0_;
} pba_options;

/* One row of Ceres' Solver::Summary::iterations (include/ceres/iteration_callback.h). */
typedef struct pba_iteration {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  int32_t linear_solver_iterations;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double model_cost_change;
  double iteration_time_in_seconds;  /* IterationSummary::iteration_time_in_seconds */
  double cumulative_time_in_seconds; /* since the minimizer started (iteration 0 = the initial evaluation) */
} pba_iteration;

/* Same wall-clock buckets as ceres::Solver::Summary (solver.h:822-851). */
typedef struct pba_summary {
  int32_t termination_type; /* PBA_CONVERGENCE / NO_CONVERGENCE / FAILURE */
  int32_t num_iterations;   /* entries written to `iterations` (incl. iteration 0) */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t num_residual_evaluations;
  int32_t num_jacobian_evaluations;
  int32_t num_linear_solves;
  int32_t rcs_dim;           /* reduced camera system dimension */
  int32_t linear_solver;     /* PBA_SOLVER_* actually used */
  int32_t num_inexact_linear_solves; /* PCG solves that stopped above pcg_tolerance (treated as invalid steps) */
  int64_t rcs_blocks;        /* stored upper-triangular blocks */
  int64_t num_residual_blocks;
  int64_t num_residuals;
  int64_t num_effective_parameters;
  int64_t gpu_kernel_launches;
  double initial_cost;
  double final_cost;
  double setup_time_in_seconds;      /* flatten + sort + H2D */
  double residual_evaluation_time_in_seconds;
  double jacobian_evaluation_time_in_seconds;
  double linear_solver_time_in_seconds;
  double minimizer_time_in_seconds;
  double total_time_in_seconds;
  pba_iteration* iterations;         /* caller-provided, may be NULL */
  int32_t iterations_capacity;
  char message[256];
} pba_summary;

typedef struct pba_kernel_stat {
  char name[48];
  int64_t launches;
  double total_ms; /* CUDA-event time, only when options.profile = 1 */
} pba_kernel_stat;

typedef struct pba_handle pba_handle; /* device-resident problem */

/* ---- library ---- */
int32_t pba_abi_version(void);
const char* pba_status_string(pba_status s);
int32_t pba_device_count(void); /* 0 when no CUDA device is usable */
void pba_options_init(pba_options* o); /* BundleAdjustmentOptions + Ceres defaults */

/* ---- the drop-in call: replaces bundle_adjustment() (map_utils.h:322) ----
 * HOST buffers in, HOST buffers updated in place.  On PBA_FAILURE termination
 * the inputs are left untouched (Ceres restores them, solver.cc:438-447);
 * otherwise the lowest-cost accepted iterate is written back
 * (trust_region_minimizer.cc:316-318). */
pba_status pba_solve(pba_problem* problem, const pba_options* options,
                     pba_summary* summary);

/* Optional: create (and cache for the life of the process) the NCCL communicators that
 * pba_solve(options.num_gpus > 1) uses on devices device .. device + num_gpus - 1 (num_gpus 0 = all from
 * `device` on).  pba_solve does this on demand; calling it up front keeps the seconds ncclCommInitAll
 * takes out of the first solve. */
pba_status pba_multi_gpu_init(int32_t device, int32_t num_gpus);

/* Host-only (needs no GPU): the camera layout pba_create derives — which poses become RCS slots and in what
 * ORDER (the keyframes' own order, or reverse Cuthill-McKee on the covisibility graph when that order is too
 * wide for the banded solvers: loop closures, unordered maps), the half-bandwidth in blocks before / after,
 * and the number of stored upper-triangular RCS blocks.  slot [n_poses] receives the slot or -1; every output
 * may be NULL.  Replaces Ceres' ordering + block-structure detection (trust_region_preprocessor.cc:373,
 * schur_complement_solver.cc:250-297). */
pba_status pba_analyze_structure(const pba_problem* problem, const pba_options* options, int32_t* slot,
                                 int32_t* n_slots, int32_t* bandwidth_natural, int32_t* bandwidth,
                                 int64_t* n_blocks);

/* ---- split entry points (tests / bench): device-resident problem ---- */
/* Replaces Problem construction + Ceres preprocessing (map_utils.h:327-375,
 * trust_region_preprocessor.cc:373): validates, orders observations by
 * (host,target) edge, builds the RCS block structure, uploads everything.
 * (rank, world_size): landmark shard for multi-GPU; (0,1) = whole problem. */
pba_status pba_create(const pba_problem* problem, const pba_options* options,
                      int32_t rank, int32_t world_size, pba_handle** out);
void pba_destroy(pba_handle* h);
/* A destroyed handle's device memory stays in a per-process cache of large chunks and is reused by the
 * next pba_create / pba_solve (allocating and freeing tens of GB costs more than a solve).  This returns
 * the cached chunks to the driver. */
void pba_trim_device_cache(void);

/* Use an existing stream (a cudaStream_t, e.g. torch's current stream) so the
 * caller can bracket work with its own CUDA events.  NULL = library stream. */
pba_status pba_set_stream(pba_handle* h, void* cuda_stream);
pba_status pba_synchronize(pba_handle* h);

/* Replaces ProgramEvaluator::Evaluate (program_evaluator.h:139-286) for this
 * rank's observations: with_jacobian=1 runs the residual+Jacobian kernel (K1)
 * and materialises the robustified local Jacobian; 0 runs the cost-only
 * kernel (K2).  *cost (host, may be NULL) = sum over blocks of rho(s)/2.
 * When cost is NULL the call is asynchronous on the handle's stream. */
pba_status pba_evaluate(pba_handle* h, int32_t with_jacobian, double* cost);

/* Download per-block outputs of the last pba_evaluate(h,1,..) in the
 * CALLER'S observation order (local shard): residuals [n_obs*R], jacobians
 * [n_obs*R*C] row-major per block with columns (host pose 6 | target pose 6 |
 * [affine 2] | rho 1), robustified like residual_block.cc:166-196.
 * These are the cost function's local Jacobians (what the reference's AutoDiff + local
 * parameterisation produce for the block): columns of constant (fixed) poses are NOT zeroed here;
 * constness is applied when the reduced camera system is assembled. */
pba_status pba_get_residuals(pba_handle* h, double* residuals);
pba_status pba_get_jacobians(pba_handle* h, double* jacobians);
/* The same outputs for a SELECTION of blocks only (parity checks on problems whose whole Jacobian
 * is tens of GB): obs_index [n_sel] = caller-order local observation indices (any order, no
 * duplicates); residuals [n_sel*R] and jacobians [n_sel*R*C] (either may be NULL) come back in the
 * order of obs_index. */
pba_status pba_get_blocks(pba_handle* h, int64_t n_sel, const int64_t* obs_index, double* residuals,
                          double* jacobians);

/* Replaces SchurEliminator::Eliminate (schur_eliminator_impl.h:177-306) on the
 * last evaluated Jacobian with Jacobi scaling and LM damping for `radius`
 * (levenberg_marquardt_strategy.cc:76-88); multi-GPU: includes the NCCL
 * all-reduce.  pba_get_rcs copies the dense symmetric RCS [dim*dim] and rhs
 * [dim] to the host (test sizes only). */
pba_status pba_build_rcs(pba_handle* h, double radius);
pba_status pba_get_rcs_dim(pba_handle* h, int32_t* dim);
pba_status pba_get_rcs(pba_handle* h, double* S_dense, double* rhs);

/* Solve the RCS built by pba_build_rcs with the configured solver; copy the
 * (Jacobi-scaled, un-negated) camera solution to `y_cam` [dim] if non-NULL. */
pba_status pba_solve_rcs(pba_handle* h, int32_t solver, double* y_cam,
                         int32_t* iterations);

/* Replaces TrustRegionMinimizer::Minimize (trust_region_minimizer.cc:67-132)
 * on the resident problem.  State stays on the device. */
pba_status pba_minimize(pba_handle* h, pba_summary* summary);

/* One complete LM iteration at the handle's CURRENT state with trust-region
 * radius `radius` — exactly the work of one successful Ceres iteration:
 * residual+Jacobian evaluation (K1), normal equations + Schur elimination
 * (+ all-reduce), RCS solve, back-substitution, model cost, candidate = Plus(x,
 * step) and its cost (K2).  Jacobi scaling / LM diagonal are taken from this
 * Jacobian (iteration-0 semantics).  out->cost = cost at x, out->cost_change =
 * cost(x) - cost(candidate), out->model_cost_change, out->relative_decrease,
 * out->step_norm.  apply != 0 moves the state to the candidate when the step
 * would be accepted (relative_decrease > min_relative_decrease); apply == 0
 * leaves the state untouched, so repeated calls do identical work (bench.py). */
pba_status pba_lm_iterate(pba_handle* h, double radius, int32_t apply, pba_iteration* out);

/* Move optimisation state (poses [n_poses*7], inv_depth [n_landmarks local
 * shard order = caller order], affine [n_poses*2] or NULL). */
pba_status pba_set_state(pba_handle* h, const double* poses,
                         const double* inv_depth, const double* affine);
pba_status pba_get_state(pba_handle* h, double* poses, double* inv_depth,
                         double* affine);

/* Local shard sizes (after landmark partitioning). */
pba_status pba_get_sizes(pba_handle* h, int64_t* n_obs_local,
                         int32_t* n_landmarks_local, int64_t* first_landmark);

/* Kernel accounting: every kernel launch is counted; durations need profile=1. */
pba_status pba_reset_kernel_stats(pba_handle* h);
/* Change pba_options.profile of a live handle (0 / 1 / 2). */
pba_status pba_set_profile(pba_handle* h, int32_t level);
int32_t pba_get_kernel_stats(pba_handle* h, pba_kernel_stat* out, int32_t cap);

/* ---- multi-GPU: one process per GPU, NCCL all-reduce of the partial RCS ----
 * Rank 0 obtains an id (128 bytes), the caller broadcasts it by any means
 * (torch.distributed in bench.py), every rank calls pba_comm_init before
 * pba_build_rcs / pba_minimize.  Communicators are cached per process and id:
 * handles created later with the same id (and world/rank/device) reuse the
 * first communicator, so only the first pba_comm_init pays ncclCommInitRank. */
#define PBA_NCCL_ID_BYTES 128
pba_status pba_nccl_unique_id(uint8_t id[PBA_NCCL_ID_BYTES]);
pba_status pba_comm_init(pba_handle* h, const uint8_t id[PBA_NCCL_ID_BYTES]);
/* Which implementation carries the handle's data-path collectives (the sum of the partial reduced camera
 * systems, the candidate-cost scalars): 0 = none (one rank), 1 = ncclAllReduce, 2 = the library's own
 * peer-memory kernels over NVLink (csrc/peer.cu: every rank's buffer is mapped by its peers — cudaIpc* between
 * processes, peer access inside one; PBA_NO_PEER=1 in the environment keeps NCCL).  pba_comm_init sets the
 * exchange up collectively and falls back to NCCL on every rank if any rank cannot map its peers. */
int32_t pba_collective_kind(pba_handle* h);

/* ---- stand-alone primitives exposed for parity tests ---- */
/* Camera models on the device (camera_models.h project/unproject), n points. */
pba_status pba_camera_project(int32_t model, const double intr[8], int64_t n,
                              const double* xyz, double* uv, double* duv_dxyz);
pba_status pba_camera_unproject(int32_t model, const double intr[8], int64_t n,
                                const double* uv, double* xyz);
/* LocalParameterizationSE3::Plus (local_parameterization_se3.hpp:44-51). */
pba_status pba_se3_plus(int64_t n, const double* poses7, const double* delta6,
                        double* out7);
/* Dense fp64 Cholesky solve A x = b on the device (A symmetric [n*n]). */
pba_status pba_cholesky_solve(int32_t n, const double* A, const double* b,
                              double* x);

/* ---- the callers either side of bundle_adjustment() ----
 * optimize() (src/sfm.cpp:1883-1925) is followed by compute_projections()
 * (src/sfm.cpp:1956-2008) + set_outlier_flags() (src/sfm.cpp:1928-1952) and
 * remove_outlier_landmarks() (src/sfm.cpp:2029-2132); all of them go through
 * Landmark::get_p() (include/visnav/common_types.h:205-217). */

/* OutlierFlags (include/visnav/common_types.h:277-285), same bit values. */
enum {
  PBA_OUTLIER_NONE = 0,
  PBA_OUTLIER_REPROJECTION_ERROR_HUGE = 1 << 0,
  PBA_OUTLIER_REPROJECTION_ERROR_NORMAL = 1 << 1,
  PBA_OUTLIER_CAMERA_DISTANCE = 1 << 2,
  PBA_OUTLIER_Z_COORDINATE = 1 << 3
};

/* The four pangolin::Var thresholds of src/sfm.cpp:254-261 (defaults 40, 3,
 * 0.1, 0.05 through pba_projection_thresholds_init). */
typedef struct pba_projection_thresholds {
  double reprojection_error_huge_pixel;
  double reprojection_error_normal_pixel;
  double camera_center_distance_meter;
  double z_coordinate_meter;
} pba_projection_thresholds;
void pba_projection_thresholds_init(pba_projection_thresholds* t);

/* Landmark::get_p for every landmark: p_w[l] = T_w_host * (normalize(unproject(z_h)) / inv_depth).
 * p_w is HOST memory [n_landmarks*3].  Needs only the pose / intrinsics /
 * landmark-host arrays of `p` (observations and images may be NULL). */
pba_status pba_landmark_positions(const pba_problem* p, int32_t device, double* p_w);

/* compute_projections() + set_outlier_flags() over the inlier observations of
 * every landmark, and the keep/remove decision of remove_outlier_landmarks().
 * Slot layout: landmark l owns slots [lm_obs_ptr[l] + l, lm_obs_ptr[l+1] + l + 1):
 * first its host observation (obs.begin()), then its CSR observations in
 * order; n_slots = n_obs + n_landmarks.  `p->obs_uv` must be set (corner
 * positions), whatever `p->mode` is.  Each observation is projected with the
 * model + intrinsics of the OBSERVING camera (src/sfm.cpp:1974).
 * Outputs are HOST arrays, each optional (NULL = not wanted):
 *   point_reprojected [n_slots*2], point_3d_c [n_slots*3],
 *   reprojection_error [n_slots], outlier_flags [n_slots],
 *   landmark_remove [n_landmarks] (1 = remove_outlier_landmarks would erase it),
 *   any_severe_outliers [1] (src/sfm.cpp:2039-2051). */
pba_status pba_compute_projections(const pba_problem* p, const pba_projection_thresholds* thresholds,
                                   int32_t device, double* point_reprojected, double* point_3d_c,
                                   double* reprojection_error, uint32_t* outlier_flags,
                                   uint8_t* landmark_remove, int32_t* any_severe_outliers);

/* add_new_landmarks_between_cams() (include/visnav/map_utils.h:121-195), the step right BEFORE bundle_adjustment()
 * in the SfM loop: for n feature tracks shared by cameras 0 and 1 (corner pixels uv0 / uv1 [n*2], HOST memory),
 * v = normalize(unproject(z)) in each camera, p = opengv::triangulation::triangulate in camera 0's frame with
 * T_c0_c1 = T_w_c0^-1 T_w_c1 (thirdparty/opengv/src/triangulation/methods.cpp:36-64: null vector of the 4x4 DLT
 * matrix), and the new landmark's inverse distance 1 / |p| (map_utils.h:190, kept as the reference computes it,
 * including its own "TODO check correctness": the distance is measured in camera 0 even when the landmark's host,
 * obs.begin(), is another camera).  Outputs (HOST): p_c0 [n*3] (may be NULL), inv_depth [n]. */
pba_status pba_triangulate_inverse_depth(int32_t model0, const double intr0[8], int32_t model1, const double intr1[8],
                                         const double T_w_c0[7], const double T_w_c1[7], int64_t n,
                                         const double* uv0, const double* uv1, int32_t device, double* p_c0,
                                         double* inv_depth);

/* ---- the feature front-end that produces the BA problem ----
 * The SfM pipeline in front of bundle_adjustment() (src/sfm.cpp:1191-1330): per image
 * detectKeypointsAndDescriptors() (include/visnav/keypoints.h:133-245), then matchDescriptors()
 * (keypoints.h:248-300) for the stereo pairs (match_stereo, src/sfm.cpp:1216-1262, followed by
 * findInliersEssential, include/visnav/matching_utils.h:62-79) and for ALL image pairs (match_all,
 * src/sfm.cpp:1286-1330).  Corner DETECTION stays with the caller (the reference calls
 * cv::goodFeaturesToTrack); everything after it is byte / integer work and is bit-exact here. */

/* computeAngles() + computeDescriptors() (keypoints.h:182-245) for the corners of n_images images of one size.
 * images: HOST, n_images x image_stride bytes, row pitch `pitch`; corner_ptr [n_images + 1] (CSR into corners);
 * corners [n][2] = (x, y) doubles, each at least 19 pixels (EDGE_THRESHOLD, keypoints.h:50) inside its image,
 * as detectKeypoints() guarantees (keypoints.h:146-150) — anything else is PBA_ERR_INVALID_ARGUMENT.
 * rotate_features = 0: all angles 0 (keypoints.h:194).  Outputs (HOST): angles [n] (radians, atan2 of the patch's
 * intensity centroid), descriptors [n][32]: bit d of the reference's std::bitset<256> is bit d % 8 of byte d / 8. */
pba_status pba_corner_descriptors(const uint8_t* images, int32_t n_images, int64_t image_stride, int32_t width,
                                  int32_t height, int32_t pitch, const int32_t* corner_ptr, const double* corners,
                                  int32_t rotate_features, int32_t device, double* angles, uint8_t* descriptors);

/* matchDescriptors() (keypoints.h:248-300) for a LIST of image pairs in one call: brute-force Hamming distance in
 * both directions, best match kept when its distance < threshold and the second best >= best * dist_2_best
 * (ties: the lowest index wins, as the reference's strict comparisons do), and a match (i, j) survives when
 * i -> j and j -> i agree.  set_ptr [n_sets + 1]: CSR into descriptors [.][32] (one set per image); pairs
 * [n_pairs][2] = set indices (first, second).  Outputs (HOST): match_ptr [n_pairs + 1], matches [capacity][2] =
 * (index in the first set, index in the second set) per pair, ascending in the first index (the reference's own
 * order is its unordered_map's, keypoints.h:291-297).  capacity >= sum over pairs of min(|first|, |second|) always
 * suffices; a smaller buffer that overflows returns PBA_ERR_INVALID_ARGUMENT with match_ptr filled in. */
pba_status pba_match_descriptors(int32_t n_sets, const int32_t* set_ptr, const uint8_t* descriptors, int32_t n_pairs,
                                 const int32_t* pairs, int32_t threshold, double dist_2_best, int32_t device,
                                 int64_t* match_ptr, int32_t* matches, int64_t capacity);

/* computeEssential() + findInliersEssential() (matching_utils.h:50-79): E = [t / |t|]x R of T_0_1 (7 doubles, Sophus
 * order) — returned in E_out [9] (row-major, may be NULL) — and inlier[k] = |x0^T E x1| <= threshold for match k,
 * x = unproject(corner) with each camera's model (camera_models.h).  All pointers HOST. */
pba_status pba_epipolar_inliers(int32_t model0, const double intr0[8], int32_t model1, const double intr1[8],
                                const double T_0_1[7], double threshold, int64_t n_matches, const int32_t* matches,
                                const double* corners0, const double* corners1, int32_t device, double* E_out,
                                uint8_t* inlier);

/* TrackBuilder::Build + Filter + Export (include/visnav/tracks.h:53-160; caller build_tracks, src/sfm.cpp:1511-1520):
 * feature tracks = connected components of the graph whose nodes are (image, feature) pairs and whose edges are the
 * inlier matches of the image pairs; a track is dropped when it holds two features of the same image or fewer than
 * min_length features.  feat_ptr [n_images + 1]: CSR of the images' feature counts, node id = feat_ptr[image] + feature;
 * pairs [n_pairs][2] = image indices; match_ptr [n_pairs + 1] / matches [.][2] = (feature in the first image, feature in
 * the second image) per pair — the layout pba_match_descriptors produces.  Outputs (HOST): track_of [n_nodes] = the
 * track's id, which is the SMALLEST node id of the component (the reference's ids are the roots of its union-find forest
 * and depend on the insertion order; only the partition is defined), or -1 for a node in no match or in a dropped track;
 * n_tracks [1] (may be NULL) = tracks kept.  More than 8,192 features in one image is PBA_ERR_UNSUPPORTED (the reference's
 * own limit is 5,000, src/sfm.cpp:197-198). */
pba_status pba_build_tracks(int32_t n_images, const int32_t* feat_ptr, int32_t n_pairs, const int32_t* pairs,
                            const int64_t* match_ptr, const int32_t* matches, int32_t min_length, int32_t device,
                            int32_t* track_of, int32_t* n_tracks);

#ifdef __cplusplus
}
#endif
#endif /* PBA_H_ */
