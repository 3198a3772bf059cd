// pba_internal.h — device-resident problem layout and kernel launchers shared
// by the .cu translation units of libpba_b200.so.  Not part of the ABI.
//
// HBM layout (all fp64 unless noted; n = local observations in EDGE ORDER,
// i.e. sorted by (host pose, target pose, landmark)):
//   J     [R][C+1][ld]      Jacobian planes, column C of each row = residual (SoA: a warp
//                           of 32 consecutive observations writes 256 contiguous bytes per plane)
//   orec  [n][16]           per-observation Schur record: E^T J_h (6) | E^T J_t
//                           (cd) | E^T E | E^T r — AoS, one 128 B line each
//   W     per host group g: [landmarks of g][8*(c_g+1)] dense rows
//                           (F^T E per visible camera | g_l, c_l), zero padded
//   direct partials: per edge chunk 3 cd x cd blocks + 2 cd-vectors
//   schur partials: per host group (c_g+1)(c_g+2)/2 8x8 blocks
//   RCS   [n_blocks][cd*cd] upper-triangular block list + rhs[dim]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include <memory>

#include "pba.h"
#include "pba_math.h"

namespace pba {

// ---------------------------------------------------------------- buffers --
// Host array WITHOUT value-initialisation (std::vector zero-fills serially: ~0.1 s for the
// observation-sized index arrays of pba_create, which are fully overwritten in parallel anyway).
template <class T>
struct RawVec {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  RawVec() {}
  explicit RawVec(size_t count) { resize(count); }
  void resize(size_t count) { p.reset(count ? new T[count] : nullptr); n = count; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  const T* data() const { return p.get(); }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
};

// Device memory of a handle comes from an arena of large chunks (a handle makes ~80 allocations;
// cudaMalloc / cudaFree of tens of GB cost 0.2-0.8 s on a fresh box, more than the 20 LM iterations
// of a solve).  When the handle dies its chunks go to a per-process cache and the next handle
// (the SfM loop calls optimize() again and again) takes them back without touching the driver;
// pba_trim_device_cache() returns them.  Memory from the cache is NOT zeroed.
struct DeviceArena {
  struct Chunk { char* p; size_t bytes; };
  std::vector<Chunk> chunks;
  size_t used = 0;  // bytes used in chunks.back()
  int device = 0;
  void* alloc(size_t bytes, cudaError_t* err);
  void release_to_cache();
  // Stack discipline for set-up temporaries (image staging, per-edge tables): everything allocated
  // after mark() is handed back by rewind(); the caller guarantees the stream is done with it.
  struct Mark { size_t n_chunks, used; };
  Mark mark() const { return Mark{chunks.size(), used}; }
  void rewind(const Mark& m);
  DeviceArena() {}
  DeviceArena(const DeviceArena&) = delete;
  DeviceArena& operator=(const DeviceArena&) = delete;
  ~DeviceArena() { release_to_cache(); }
};
// Set (per thread) while pba_create builds a handle: DevBuf allocations then come from its arena.
DeviceArena*& current_arena();
struct ArenaScope {
  DeviceArena* prev;
  explicit ArenaScope(DeviceArena* a) : prev(current_arena()) { current_arena() = a; }
  ~ArenaScope() { current_arena() = prev; }
};
void trim_device_cache();

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owned = true;  // false: the memory belongs to a DeviceArena
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    n = 0;
    owned = true;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    if (DeviceArena* a = current_arena()) {
      cudaError_t e = cudaSuccess;
      p = static_cast<T*>(a->alloc(count * sizeof(T), &e));
      owned = false;
      return e;
    }
    return cudaMalloc(&p, count * sizeof(T));
  }
  template <class V>
  cudaError_t upload(const V& h, cudaStream_t s = 0) {  // std::vector<T> or RawVec<T>
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

// ---------------------------------------------------------- kernel stats ---
enum KernelId {
  K_INIT_LM = 0,
  K_EDGE_PREP,
  K_RESJAC,      // K1: residual + Jacobian (the roofline kernel)
  K_COST,        // K2: cost-only evaluation
  K_REDUCE_SUM,  // deterministic second-stage reductions
  K_EDGE_GRAM,   // K4a: per-edge F^T F / F^T r
  K_LM_GATHER,   // K3b: per-landmark E^T F rows, c_l, g_l
  K_LM_SCALE,    // per-landmark Jacobi scale / damping / ete^-1
  K_SCHUR_SYRK,  // K4c: per-host-group W^T diag(s) W
  K_RCS_REDUCE,  // K4d: gather partial blocks into the RCS
  K_RCS_SCALE,   // K4e: Jacobi scaling + LM damping of the RCS
  K_CAM_SCALE,   // iteration-0 camera column scales / LM diagonal
  K_DENSE_FILL,
  K_CHOL_PANEL,
  K_CHOL_TRSM,
  K_CHOL_SYRK,   // DMMA trailing update
  K_CHOL_SOLVE,
  K_PCG,
  K_BAND_CHOL,   // single-CTA block-banded Cholesky + solves
  K_BCR,         // block cyclic reduction on super blocks
  K_BACKSUB,     // K6
  K_MODEL_COST,  // K8
  K_RETRACT,     // K7
  K_COPY,
  K_UNPERMUTE,
  K_PRIMITIVE,
  K_NUM
};

extern const char* const kKernelNames[K_NUM];

struct KernelStats {
  int64_t launches[K_NUM] = {0};
  double ms[K_NUM] = {0};
  int profile = 0;  // 0 none, 1 every kernel, 2 only K_RESJAC (the roofline kernel)
  bool timing_now = false;
  // pending event pairs (resolved at synchronisation points)
  struct Pending { int id; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event();
  void begin(int id, cudaStream_t s);
  void end(cudaStream_t s);
  void resolve();  // call after a stream sync
  void reset();
  ~KernelStats();
};

// Stored Jacobian planes: the six target-pose columns of a row are never stored (eval.cu: they are
// the host-pose columns x a per-edge adjoint), so a row of C columns + residual has C + 1 - 6 planes:
// columns 0..5 (host pose) -> planes 0..5, every later column c (affine, inverse distance) and the
// residual (c = C) -> plane c - 6.  Photometric: 10 planes per row, geometric: 8.
// Per-edge record written by k_edge_prep before every evaluation: everything an observation needs
// that is constant along its (host, target) edge sits in ONE 256-byte record, so the dependent load
// chain of the evaluation kernels is obs_edge -> record (it used to be obs_edge -> edge_t -> pose_calib
// -> intrinsics: four dependent global loads before the first useful flop).
//   [0..8] A = R_t^T R_h   [9..11] t = R_t^T (t_h - t_t)   [12] exp(a_t)   [13] b_t
//   [14] target calibration index   [15] host camera model id
//   [16..23] target intrinsics   [24] 1 / fx_host  [25] 1 / fy_host  [26] cx_host  [27] cy_host
constexpr int kEdgeStride = 32;
constexpr int kReduceMid = 64;  // CTAs of the first stage of a long scalar reduction
constexpr int kPhotoPlanes = 10;
constexpr int kGeomPlanes = 8;
__host__ __device__ inline int stored_plane(int c) { return c < 6 ? c : c - 6; }

// ------------------------------------------------------------- the handle --
struct Sizes {
  int mode = 0, R = 2, C = 13, cd = 6;  // residuals/obs, local columns/obs, RCS block dim
  int n_poses = 0, n_calib = 0;
  int n_lm = 0;         // local landmarks
  int64_t n_obs = 0;    // local observations
  int64_t ld = 0;       // plane stride of res / J: n_obs rounded up to 32 (256 B aligned planes)
  int n_edges = 0;
  int n_chunks = 0;     // edge chunks (direct partials)
  int n_groups = 0;     // host groups
  int n_slots = 0, dim = 0;
  int64_t n_blocks = 0; // RCS upper blocks
  int width = 0, height = 0, pitch = 0;
  int64_t image_stride = 0;
};

struct Handle {
  DeviceArena arena;  // first member: destroyed last, after every DevBuf below
  pba_options opt;
  Sizes sz;
  int rank = 0, world = 1;
  int64_t first_landmark = 0;  // caller index of local landmark 0's shard start
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  KernelStats stats;

  // ---- host-side structure kept for state I/O and the LM driver ----
  std::vector<int> lm_order;      // internal landmark -> caller-local landmark (sorted by host)
  RawVec<int64_t> obs_order;      // sorted obs position -> caller-local obs index
  std::vector<int> slot;          // pose -> RCS slot or -1
  std::vector<uint8_t> affine_active;
  std::vector<int> blk_row, blk_col;  // RCS block coordinates (slots), row <= col
  std::vector<int> diag_blk;          // slot -> block index of (slot,slot)
  int64_t n_active_lm = 0;            // landmarks with >= 1 observation (global, for counts)
  int64_t n_obs_global = 0;

  // ---- static device data ----
  DevBuf<double> intr;             // [n_calib*8]
  DevBuf<int> pose_calib, calib_model, d_slot;
  DevBuf<uint8_t> d_affine_active;
  DevBuf<uint32_t> quads;          // photometric: all keyframes as 2x2-footprint words [n_poses][h][w]
  DevBuf<int> edge_h, edge_t;      // [E]
  DevBuf<int64_t> edge_ptr;        // [E+1]
  DevBuf<int> obs_lm, obs_edge;    // [n]
  DevBuf<double> obs_uv;           // geometric [2][n] SoA
  DevBuf<double> lm_uv;            // [n_lm*2] host pixel
  DevBuf<int> lm_host;             // [n_lm]
  DevBuf<double> lm_pat;           // photometric SoA [8][4 = bx,by,bz,I_h][n_lm]; geometric [n_lm][4]
  DevBuf<uint8_t> lm_ok;           // [n_lm]
  DevBuf<int64_t> lm_ptr;          // [n_lm+1] landmark -> positions in sorted obs
  DevBuf<int64_t> lm_pos;          // [n]
  DevBuf<int> lm_group;            // [n_lm]
  DevBuf<int> obs_col;             // [n] column (camera index within the host group) of the target
  DevBuf<int> lm_hostcol;          // [n_lm] column of the host within its group or -1
  // edge chunks
  DevBuf<int> chunk_edge;          // [n_chunks]
  DevBuf<int64_t> chunk_begin, chunk_end;
  // host groups
  DevBuf<int> grp_lm_ptr;          // [G+1] landmark ranges
  DevBuf<int> grp_cam_ptr;         // [G+1] into grp_cams
  DevBuf<int> grp_cams;            // slots, ascending
  DevBuf<int64_t> grp_w_off;       // [G] offset of the group's W rows (doubles)
  DevBuf<int64_t> grp_part_off;    // [G] offset of the group's Schur partial blocks (blocks of 64)
  DevBuf<int> syrk_work;           // [n_syrk_work][2] (group, first tile) of the generic Schur product's passes
  int n_syrk_work = 0;
  DevBuf<int64_t> lm_w_off;        // [n_lm] offset of the landmark's W row
  DevBuf<int> lm_w_stride;         // [n_lm] row length = 8*(c_g+1)
  // RCS structure
  DevBuf<int> d_blk_row, d_blk_col;
  DevBuf<int64_t> blk_dir_ptr, blk_sch_ptr;  // [n_blocks+1]
  DevBuf<int64_t> blk_dir_src, blk_sch_src;  // offsets (doubles) into direct / schur partial buffers
  DevBuf<int64_t> vec_dir_ptr, vec_sch_ptr;  // [n_slots+1]
  DevBuf<int64_t> vec_dir_src, vec_sch_src;
  // symmetric block-row CSR for the PCG SpMV
  DevBuf<int> row_ptr, row_blk, row_col;
  DevBuf<uint8_t> row_trans;

  // ---- state ----
  DevBuf<double> poses, affine, rho;            // current x
  DevBuf<double> poses_c, affine_c, rho_c;      // candidate
  DevBuf<double> poses_best, affine_best, rho_best;

  // ---- per-evaluation data ----
  DevBuf<double> edge_T;       // [E][kEdgeStride] per-edge record (eval.cu: k_edge_prep)
  DevBuf<double> edge_M;       // [E][36] photometric: target-pose Jacobian columns = host-pose columns x M (adjoint)
  DevBuf<double> J;            // interleaved planes [R][C+1][ld]: Jacobian row columns + residual
  DevBuf<double> orec;         // [n][16]
  DevBuf<double> W;            // grouped landmark rows
  DevBuf<double> lm_c, lm_g;   // [n_lm] raw E^T E, E^T r
  DevBuf<double> lm_scale;     // [n_lm] Jacobi scale sigma_l
  DevBuf<double> lm_diag;      // [n_lm] frozen LM diagonal (scaled col norm, clamped)
  DevBuf<double> lm_s2;        // [n_lm] sigma_l^2 / ete_l
  DevBuf<double> lm_iete;      // [n_lm] 1 / ete_l
  DevBuf<double> part_dir;     // direct partials
  DevBuf<double> part_sch;     // schur partials
  DevBuf<double> rcs;          // [n_blocks*cd*cd | rhs dim | diagB dim]  (one all-reduce buffer)
  DevBuf<double> rcs_B;        // this rank's raw direct part B [n_blocks*cd*cd] + its gradient g_c [dim]
                               // (camera part of the model cost change)
  DevBuf<double> cam_scale;    // [dim]
  DevBuf<double> cam_diag;     // [dim] frozen LM diagonal
  DevBuf<double> cam_D2;       // [dim] D^2 of the current solve
  DevBuf<double> y_cam;        // [dim] RCS solution (scaled space)
  DevBuf<double> d_cam;        // [dim] tangent step (unscaled)
  DevBuf<double> d_rho;        // [n_lm]
  DevBuf<double> dense;        // [dim*dim] Cholesky workspace
  DevBuf<double> pcg_ws;       // PCG vectors
  DevBuf<double> blk_inv;      // block-Jacobi inverses [n_slots*cd*cd]
  DevBuf<double> red_ws;       // reduction workspace
  DevBuf<double> red_mid;      // [kReduceMid] first-stage sums of the long reductions
  DevBuf<double> scalars;      // small device scalar bank
  double* h_scalars = nullptr; // pinned mirror

  double t_wait = 0.0;         // host seconds spent waiting for the device (diagnostics, PBA_TIMING)
  bool have_jac = false;
  bool have_rcs = false;
  bool scale_ready = false;
  double rcs_radius = 0;

  DevBuf<int> chol_fail;       // set by the Cholesky when a pivot is not positive
  DevBuf<int> d_diag_blk;      // slot -> diagonal block index
  int schur_tile_l = 32;       // landmark rows staged per pass in k_schur_syrk
  int max_w_stride = 8;        // longest W row (doubles)
  int pcg_grid = 1;            // co-resident CTAs for the cooperative PCG kernel
  int last_solver = 0;
  int rcs_bandwidth = 0;       // max (col - row) over the RCS blocks, in blocks
  DevBuf<int> d_col_blk;       // [n_slots][bw+1] block index of (k, k+i) or -1 (band solver)
  DevBuf<double> band_L;       // band factor [n_slots][bw+1][cd*cd]
  int bcr_m = 0;               // keyframes per BCR super block (0 = BCR not applicable)
  int bcr_levels = 0;
  std::vector<int> bcr_n;      // super blocks per level
  std::vector<size_t> bcr_off; // 7 offsets per level into bcr_ws (A B b L U V y)
  size_t bcr_x_off = 0;
  DevBuf<double> bcr_ws;
  // second-generation BCR (bcr2.cu): padded super blocks Mp = 8 b2_nbk; 0 = not applicable (first generation runs)
  int b2_nbk = 0;
  std::vector<int> b2_n;       // super blocks per level
  std::vector<size_t> b2_off;  // 6 offsets per level into b2_ws (A B b Uh Vh yh)
  size_t b2_x_off = 0;
  DevBuf<double> b2_ws;
  int uniform_model = -1;      // PBA_CAM_* when every calibration uses the same model, else -1

  // NCCL
  void* nccl_comm = nullptr;
  // exchange allocation mapped by every rank (peer.cu): when set, `rcs` points into it and the collectives of the
  // data path run as our own NVLink kernels instead of ncclAllReduce.  Owned by the communicator cache.
  struct PeerExchange* peer = nullptr;
  struct EmuExchange* emu = nullptr;  // ranks emulated on ONE device (host.cu, PBA_EMULATE_RANKS): host barrier + one sum kernel
  // Tail of the all-reduce buffer `rcs` (world > 1): [0] cost, [1] sum of squared landmark gradients,
  // [2 + r] max |landmark gradient| of rank r (every rank writes its own slot, the others stay 0, so the SUM
  // all-reduce gathers them): one collective per Jacobian evaluation carries the RCS and every scalar.
  double* rcs_tail() { return rcs.p + size_t(sz.n_blocks) * sz.cd * sz.cd + 3 * size_t(sz.dim); }

  ~Handle();
};

// scalar bank indices
enum { S_COST = 0, S_COST_C, S_MODEL, S_STEP2, S_XNORM2, S_GMAX, S_GNORM2, S_PCG_ITERS, S_PCG_RES, S_CHOL_FAIL, S_NUM = 16 };

#define PBA_CUDA_OK(x)                          \
  do {                                          \
    cudaError_t _e = (x);                       \
    if (_e != cudaSuccess) return map_cuda(_e); \
  } while (0)

pba_status map_cuda(cudaError_t e);

// -------------------------------------------------------------- launchers --
// eval.cu
pba_status launch_init_landmarks(Handle* h);
pba_status launch_build_quads(Handle* h, cudaStream_t stream, const uint8_t* images_u8, int first, int n_img);
pba_status launch_evaluate(Handle* h, bool with_jacobian, const double* poses, const double* affine,
                           const double* rho, double* cost_out);
pba_status launch_unpermute(Handle* h, int which, double* dst_host_order_dev);
pba_status launch_gather_blocks(Handle* h, int64_t n_sel, const int64_t* sel_pos_dev, double* res_dev, double* jac_dev);
pba_status launch_expand_edges(Handle* h, const int* edge_col_dev);  // fills obs_edge / obs_col from edge_ptr
// schur.cu
pba_status launch_post_jacobian(Handle* h);              // edge Gram + landmark gather (+ scales on first call)
// with_scalars (multi-rank Jacobian evaluations): the all-reduce also carries rcs_tail() — this rank's cost and
// landmark gradient norms — so an evaluation needs ONE collective
pba_status launch_build_rcs(Handle* h, double radius, bool refresh_diag, bool with_scalars = false);
pba_status launch_landmark_gradient_norms(Handle* h);  // world > 1: fills rcs_tail()[1], [2 + rank] before the all-reduce
pba_status launch_backsub(Handle* h);  // also leaves the model cost change in S_MODEL
void launch_reduce_sum(Handle* h, const double* part, int64_t n, double* out);
pba_status launch_retract(Handle* h);
pba_status launch_gradient_norms(Handle* h);
// solve.cu
pba_status launch_cholesky_rcs(Handle* h);
pba_status launch_pcg_rcs(Handle* h);
pba_status launch_band_rcs(Handle* h);
pba_status launch_bcr_rcs(Handle* h);
pba_status bcr_setup(Handle* h);
pba_status bcr2_setup(Handle* h);
pba_status launch_bcr2_rcs(Handle* h);
int band_max_bw(int cd);
pba_status dense_cholesky_solve(Handle* h, double* A, double* b, int ld, int* fail_dev, double* work);
size_t dense_work_size(int ld);
int dense_ld(int n);
int pcg_max_grid(int device);
int schur_tile_l(int max_stride);
void schur_syrk_work(const std::vector<int>& grp_cam_ptr, std::vector<int>* work);  // (group, first tile) per pass
int eval_grid(int64_t n);
// host.cu
// ---- peer-memory collectives (peer.cu) ----
constexpr int kMaxPeers = 8;     // ranks of one NVSwitch domain
constexpr int kPeerSmall = 16;   // doubles per rank in the scalar table
struct PeerExchange {
  int world = 1, rank = 0, device = 0;
  size_t bytes = 0;                          // payload capacity of every rank's allocation
  char* base[kMaxPeers] = {};                // base[rank]: own cudaMalloc; the others: peer mappings
  double* buf[kMaxPeers] = {};               // payload
  unsigned long long* flags[kMaxPeers] = {}; // [3][kMaxPeers] epoch counters: barrier A, barrier B, scalar table
  double* small[kMaxPeers] = {};             // [2][kMaxPeers][kPeerSmall] scalar table, double-buffered by epoch parity
  unsigned long long epoch = 0, epoch_small = 0;
  bool ipc = false;                          // peers opened with cudaIpcOpenMemHandle (one process per GPU)
};
struct PeerArgs {
  double* buf[kMaxPeers];
  unsigned long long* flags[kMaxPeers];
  double* small[kMaxPeers];
  int rank, world;
  size_t count;
  unsigned long long epoch;
  int* fail;
};
inline size_t peer_alloc_bytes(size_t payload) { return ((payload + 255) & ~size_t(255)) + 256 + 2 * kMaxPeers * kPeerSmall * sizeof(double); }
inline void peer_set_layout(PeerExchange* px, int p) {
  char* b = px->base[p];
  px->buf[p] = reinterpret_cast<double*>(b);
  px->flags[p] = reinterpret_cast<unsigned long long*>(b + ((px->bytes + 255) & ~size_t(255)));
  px->small[p] = reinterpret_cast<double*>(b + ((px->bytes + 255) & ~size_t(255)) + 256);
}
pba_status launch_peer_allreduce(Handle* h, size_t count);
pba_status launch_peer_allreduce_small(Handle* h, double* dev, int n);

pba_status allreduce_rcs(Handle* h, bool with_scalars);
pba_status allreduce_scalars(Handle* h, double* dev, int n, bool max_op);

}  // namespace pba
