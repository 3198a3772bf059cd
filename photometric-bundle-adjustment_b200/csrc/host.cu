// host.cu — the C ABI (include/pba.h): problem flattening/ordering, the
// device-resident handle, and the Levenberg-Marquardt driver.
//
// pba_create  replaces ceres::Problem construction + preprocessing
//             (include/visnav/map_utils.h:327-375,
//             internal/ceres/trust_region_preprocessor.cc:373,
//             schur_complement_solver.cc:250-297 for the RCS block pattern).
// pba_minimize replaces TrustRegionMinimizer::Minimize with the
//             LevenbergMarquardtStrategy and the monotonic
//             TrustRegionStepEvaluator (trust_region_minimizer.cc:67-826,
//             levenberg_marquardt_strategy.cc:66-162,
//             trust_region_step_evaluator.cc:52-112), decision logic on the
//             host, every vector operation on the device.
// pba_solve   is the drop-in for visnav::bundle_adjustment (map_utils.h:322).
#include <dlfcn.h>
#include <omp.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <limits>
#include <array>
#include <atomic>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <thread>
#include <string>
#include <unordered_set>

#include "launch.h"
#include "pba_internal.h"

#define PBA_API extern "C" __attribute__((visibility("default")))

namespace pba {

const char* const kKernelNames[K_NUM] = {
    "init_landmarks", "edge_prep",  "residual_jacobian", "cost_only",  "reduce_sum", "edge_gram",
    "landmark_gather", "landmark_scale", "schur_syrk", "rcs_reduce", "rcs_scale",  "cam_scale",
    "dense_fill",     "chol_panel", "chol_trsm",         "chol_syrk_dmma", "chol_solve", "pcg",
    "band_cholesky",  "bcr",        "backsub",           "model_cost", "retract",    "copy",
    "unpermute",      "primitive"};

pba_status map_cuda(cudaError_t e) {
  if (e == cudaSuccess) return PBA_OK;
  if (e == cudaErrorMemoryAllocation) return PBA_ERR_OUT_OF_MEMORY;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return PBA_ERR_NO_DEVICE;
  fprintf(stderr, "[pba_b200] CUDA error: %s\n", cudaGetErrorString(e));
  return PBA_ERR_CUDA;
}

cudaEvent_t KernelStats::get_event() {
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void KernelStats::begin(int id, cudaStream_t s) {
  ++launches[id];
  timing_now = profile == 1 || (profile == 2 && id == K_RESJAC);
  if (!timing_now) return;
  Pending p;
  p.id = id;
  p.a = get_event();
  p.b = get_event();
  cudaEventRecord(p.a, s);
  pending.push_back(p);
}
void KernelStats::end(cudaStream_t s) {
  if (!timing_now) return;
  timing_now = false;
  cudaEventRecord(pending.back().b, s);
}
void KernelStats::resolve() {
  for (auto& p : pending) {
    float ms_ = 0;
    if (cudaEventElapsedTime(&ms_, p.a, p.b) == cudaSuccess) ms[p.id] += ms_;
    pool.push_back(p.a);
    pool.push_back(p.b);
  }
  pending.clear();
}
void KernelStats::reset() {
  resolve();
  for (int i = 0; i < K_NUM; ++i) { launches[i] = 0; ms[i] = 0; }
}
KernelStats::~KernelStats() {
  for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : pool) cudaEventDestroy(e);
}

// ------------------------------------------------------------------ NCCL ---
// Loaded lazily with dlopen so the library has no link-time NCCL dependency
// (single-GPU users never touch it).  The only collective on the path is the
// fp64 sum all-reduce of the partial RCS (plus scalar reductions).
namespace {
struct NcclId { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;  // optional: peer-memory set-up
  int (*CommDestroy)(void*) = nullptr;
  bool load() {
    if (lib) return true;
    const char* env = getenv("PBA_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    CommInitAll = (int (*)(void**, int, const int*))dlsym(lib, "ncclCommInitAll");
    GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
    GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy && CommInitAll && GroupStart && GroupEnd;
  }
};
NcclApi g_nccl;
constexpr int kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2;
}  // namespace

// small pinned blocks (the handle's scalar mirror) are recycled: cudaMallocHost / cudaFreeHost cost a millisecond each
namespace {
std::mutex g_scalar_mu;
std::vector<double*> g_scalar_blocks;  // each kScalarBlock doubles
constexpr size_t kScalarBlock = 64;
double* acquire_scalar_block() {
  {
    std::lock_guard<std::mutex> lock(g_scalar_mu);
    if (!g_scalar_blocks.empty()) { double* p = g_scalar_blocks.back(); g_scalar_blocks.pop_back(); return p; }
  }
  double* p = nullptr;
  if (cudaHostAlloc(&p, sizeof(double) * kScalarBlock, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void release_scalar_block(double* p) {
  std::lock_guard<std::mutex> lock(g_scalar_mu);
  g_scalar_blocks.push_back(p);
}
}  // namespace

Handle::~Handle() {
  // nccl_comm is owned by the process-wide cache (pba_comm_init), not by the handle
  if (stream) cudaStreamSynchronize(stream);  // nothing may still run on memory that goes back to the cache
  if (h_scalars) release_scalar_block(h_scalars);
  if (own_stream && stream) cudaStreamDestroy(stream);
}

// PBA_TIMING: slow driver allocations (they serialise every thread of the process)
std::atomic<int> g_n_cuda_malloc{0}, g_n_host_alloc{0};
std::atomic<int64_t> g_us_cuda_malloc{0}, g_us_host_alloc{0}, g_mb_cuda_malloc{0}, g_mb_host_alloc{0};
static double wall_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- device arena + per-process chunk cache ----
namespace {
constexpr size_t kArenaChunk = size_t(1) << 30;
std::mutex g_cache_mu;
struct CachedChunk { int device; DeviceArena::Chunk c; };
std::vector<CachedChunk> g_chunk_cache;
}  // namespace

DeviceArena*& current_arena() {
  static thread_local DeviceArena* a = nullptr;
  return a;
}

void* DeviceArena::alloc(size_t bytes, cudaError_t* err) {
  *err = cudaSuccess;
  bytes = (bytes + 255) & ~size_t(255);
  if (!chunks.empty() && used + bytes <= chunks.back().bytes) {
    void* r = chunks.back().p + used;
    used += bytes;
    return r;
  }
  Chunk c{nullptr, 0};
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int best = -1;  // smallest cached chunk of this device that fits
    for (int i = 0; i < int(g_chunk_cache.size()); ++i)
      if (g_chunk_cache[i].device == device && g_chunk_cache[i].c.bytes >= bytes &&
          (best < 0 || g_chunk_cache[i].c.bytes < g_chunk_cache[best].c.bytes)) best = i;
    if (best >= 0) {
      c = g_chunk_cache[best].c;
      g_chunk_cache.erase(g_chunk_cache.begin() + best);
    }
  }
  if (!c.p) {
    const size_t want = bytes > kArenaChunk ? bytes : kArenaChunk;
    void* q = nullptr;
    const double t_alloc = wall_s();
    cudaError_t e = cudaMalloc(&q, want);
    ++g_n_cuda_malloc; g_us_cuda_malloc += int64_t(1e6 * (wall_s() - t_alloc)); g_mb_cuda_malloc += int64_t(want >> 20);
    if (e != cudaSuccess) {  // give cached chunks that did not fit back to the driver and retry
      cudaGetLastError();
      trim_device_cache();
      e = cudaMalloc(&q, want);
    }
    if (e != cudaSuccess) { *err = e; return nullptr; }
    c.p = static_cast<char*>(q);
    c.bytes = want;
  }
  chunks.push_back(c);
  used = bytes;
  return c.p;
}

void DeviceArena::release_to_cache() {
  std::lock_guard<std::mutex> lock(g_cache_mu);
  for (const Chunk& c : chunks) g_chunk_cache.push_back({device, c});
  chunks.clear();
  used = 0;
}

void DeviceArena::rewind(const Mark& m) {
  if (chunks.size() > m.n_chunks) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    while (chunks.size() > m.n_chunks) { g_chunk_cache.push_back({device, chunks.back()}); chunks.pop_back(); }
  }
  used = m.used;
}

void trim_device_cache() {
  std::vector<CachedChunk> all;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    all.swap(g_chunk_cache);
  }
  int cur = 0;
  cudaGetDevice(&cur);
  for (const CachedChunk& cc : all) {
    cudaSetDevice(cc.device);
    cudaFree(cc.c.p);
  }
  cudaSetDevice(cur);
}

// (kernel, device) -> largest dynamic shared memory size opted into so far (launch.h)
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes, int device) {
  struct Entry { const void* k; int dev; size_t bytes; };
  static std::mutex mu;
  static std::vector<Entry> table;
  std::lock_guard<std::mutex> lock(mu);
  for (Entry& e : table)
    if (e.k == kernel && e.dev == device) {
      if (e.bytes >= bytes) return cudaSuccess;
      const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
      if (err == cudaSuccess) e.bytes = bytes;
      return err;
    }
  const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (err == cudaSuccess) table.push_back({kernel, device, bytes});
  return err;
}

// ---- ranks emulated on ONE device (PBA_EMULATE_RANKS=1; pba_solve with num_gpus > 1) ----
// The sharded solve with every rank's handle on the same GPU, one host thread per rank as in solve_multi_gpu.  Ranks
// must not wait for one another ON the device there (kernels of different streams need not run concurrently:
// B200_PROFILING.md), so the exchange is a HOST barrier: every rank drains its stream and arrives; the last one
// launches one kernel that sums all ranks' buffers in rank order into all of them, drains it and releases the rest.
// Everything except the transport is the production path: landmark partition, per-rank layout, the packed
// [S | rhs | diag(B) | g_c | scalars] payload, scaling after the reduction, the redundant RCS solve, per-shard
// back-substitution, the scalar exchange of the candidate evaluation.  That is what lets a ONE-GPU box test it.
struct EmuExchange {
  int world = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, generation = 0;
  bool failed = false;
  std::vector<double*> ptr;  // this round's buffer of every rank
  explicit EmuExchange(int w) : world(w), ptr(size_t(w), nullptr) {}
};

namespace {
struct EmuPtrs { double* p[kMaxPeers]; };
__global__ void k_emu_allreduce(EmuPtrs bufs, int world, size_t count, int max_op) {
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) {
    double s = bufs.p[0][i];
    for (int r = 1; r < world; ++r) s = max_op ? fmax(s, bufs.p[r][i]) : s + bufs.p[r][i];
    for (int r = 0; r < world; ++r) bufs.p[r][i] = s;
  }
}
}  // namespace

pba_status emu_allreduce(Handle* h, double* dev, size_t count, bool max_op) {
  EmuExchange* x = h->emu;
  if (x->world > kMaxPeers) return PBA_ERR_INVALID_ARGUMENT;
  const bool mine_ok = cudaStreamSynchronize(h->stream) == cudaSuccess;  // this rank's contribution is complete
  std::unique_lock<std::mutex> lock(x->mu);
  if (!mine_ok) x->failed = true;
  x->ptr[size_t(h->rank)] = dev;
  const int gen = x->generation;
  if (++x->arrived == x->world) {
    if (!x->failed) {
      EmuPtrs b;
      for (int r = 0; r < kMaxPeers; ++r) b.p[r] = r < x->world ? x->ptr[size_t(r)] : nullptr;
      const unsigned grid = unsigned(std::min<size_t>(1024, (count + 255) / 256));
      h->stats.begin(K_COPY, h->stream);
      k_emu_allreduce<<<grid, 256, 0, h->stream>>>(b, x->world, count, max_op ? 1 : 0);
      h->stats.end(h->stream);
      if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) x->failed = true;
    }
    x->arrived = 0;
    ++x->generation;
    x->cv.notify_all();
  } else {
    // a rank that left the solve early (an error on its side) never arrives: give up instead of waiting for ever
    if (!x->cv.wait_for(lock, std::chrono::seconds(120), [&] { return x->generation != gen; })) x->failed = true;
  }
  return x->failed ? PBA_ERR_NCCL : PBA_OK;
}

pba_status allreduce_rcs(Handle* h, bool with_scalars) {
  if (h->world <= 1) return PBA_OK;
  if (h->emu) {
    const Sizes& ze = h->sz;
    return emu_allreduce(h, h->rcs.p, size_t(ze.n_blocks) * ze.cd * ze.cd + 3 * size_t(ze.dim) + (with_scalars ? size_t(2 + h->world) : 0), false);
  }
  if (!h->nccl_comm) return PBA_ERR_NCCL;
  const Sizes& z = h->sz;
  const size_t count = size_t(z.n_blocks) * z.cd * z.cd + 3 * size_t(z.dim) + (with_scalars ? size_t(2 + h->world) : 0);
  if (h->peer) return launch_peer_allreduce(h, count);  // our own NVLink kernel (peer.cu); `rcs` lives in the exchange buffer
  h->stats.begin(K_COPY, h->stream);
  const int rc = g_nccl.AllReduce(h->rcs.p, h->rcs.p, count, kNcclDouble, kNcclSum, h->nccl_comm, h->stream);
  h->stats.end(h->stream);
  return rc == 0 ? PBA_OK : PBA_ERR_NCCL;
}

pba_status allreduce_scalars(Handle* h, double* dev, int n, bool max_op) {
  if (h->world <= 1) return PBA_OK;
  if (h->emu) return emu_allreduce(h, dev, size_t(n), max_op);
  if (!h->nccl_comm) return PBA_ERR_NCCL;
  if (h->peer && !max_op && n <= kPeerSmall) return launch_peer_allreduce_small(h, dev, n);
  const int rc = g_nccl.AllReduce(dev, dev, size_t(n), kNcclDouble, max_op ? kNcclMax : kNcclSum, h->nccl_comm, h->stream);
  return rc == 0 ? PBA_OK : PBA_ERR_NCCL;
}

// ---- exchange allocations of the peer-memory collectives (peer.cu) ----
// One per communicator and rank, created collectively, kept for the life of the process (like the communicator)
// and re-created only when a problem needs a larger payload.  PBA_NO_PEER=1 keeps every collective on NCCL.
namespace {

bool peer_enabled() {
  static const bool off = getenv("PBA_NO_PEER") != nullptr;
  return !off;
}

void peer_release(PeerExchange* px) {
  if (!px) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(px->device);
  for (int p = 0; p < px->world; ++p) {
    if (!px->base[p]) continue;
    if (p == px->rank) cudaFree(px->base[p]);
    else if (px->ipc) cudaIpcCloseMemHandle(px->base[p]);
  }
  cudaGetLastError();
  cudaSetDevice(cur);
  delete px;
}

// one process per GPU: allocate, exchange the IPC handles through the communicator, map the peers.  COLLECTIVE:
// every rank calls it with the same `payload` (the RCS layout is global).  On any failure EVERY rank ends up
// without an exchange (the status is all-reduced), so that nobody takes the peer path alone.
PeerExchange* peer_create_ipc(void* comm, int rank, int world, int device, size_t payload, cudaStream_t stream) {
  if (!peer_enabled() || world > kMaxPeers || !g_nccl.AllGather) return nullptr;
  PeerExchange* px = new PeerExchange;
  px->world = world; px->rank = rank; px->device = device; px->bytes = payload; px->ipc = true;
  bool ok = true;
  const size_t total = peer_alloc_bytes(payload);
  ok = cudaMalloc(reinterpret_cast<void**>(&px->base[rank]), total) == cudaSuccess;
  if (ok) ok = cudaMemset(px->base[rank], 0, total) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaIpcGetMemHandle(&mine, px->base[rank]) == cudaSuccess;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  // [world handles | world status doubles] on the device; the gather runs even after a local failure (collective)
  char* d = nullptr;
  std::vector<cudaIpcMemHandle_t> all(world);
  bool comm_ok = cudaMalloc(reinterpret_cast<void**>(&d), size_t(64) * (world + 1) + 64) == cudaSuccess;
  if (comm_ok) {
    comm_ok = cudaMemcpyAsync(d + size_t(64) * world, &mine, 64, cudaMemcpyHostToDevice, stream) == cudaSuccess &&
              g_nccl.AllGather(d + size_t(64) * world, d, 64, /*ncclInt8*/ 0, comm, stream) == 0 &&
              cudaMemcpyAsync(all.data(), d, size_t(64) * world, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
              cudaStreamSynchronize(stream) == cudaSuccess;
  }
  ok = ok && comm_ok;
  if (ok) {
    for (int p = 0; p < world && ok; ++p) {
      if (p == rank) continue;
      void* q = nullptr;
      ok = cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      px->base[p] = static_cast<char*>(q);
    }
  }
  cudaGetLastError();
  // agree: the number of ranks that failed
  double bad = ok ? 0.0 : 1.0;
  if (comm_ok) {
    double* ds = reinterpret_cast<double*>(d + size_t(64) * (world + 1));
    const bool s_ok = cudaMemcpyAsync(ds, &bad, sizeof(double), cudaMemcpyHostToDevice, stream) == cudaSuccess &&
                      g_nccl.AllReduce(ds, ds, 1, kNcclDouble, kNcclSum, comm, stream) == 0 &&
                      cudaMemcpyAsync(&bad, ds, sizeof(double), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                      cudaStreamSynchronize(stream) == cudaSuccess;
    if (!s_ok) bad = 1.0;
  }
  if (d) cudaFree(d);
  if (bad != 0.0) { peer_release(px); cudaGetLastError(); return nullptr; }
  for (int p = 0; p < world; ++p) peer_set_layout(px, p);
  return px;
}

// one process, n devices starting at device0 (the caller serialises): plain peer access
bool peer_create_local(int device0, int n, size_t payload, std::vector<PeerExchange*>* out) {
  out->assign(n, nullptr);
  if (!peer_enabled() || n > kMaxPeers) return false;
  int cur = 0;
  cudaGetDevice(&cur);
  bool ok = true;
  for (int i = 0; i < n && ok; ++i)
    for (int j = 0; j < n && ok; ++j) {
      if (i == j) continue;
      int can = 0;
      ok = cudaDeviceCanAccessPeer(&can, device0 + i, device0 + j) == cudaSuccess && can;
    }
  std::vector<char*> base(n, nullptr);
  const size_t total = peer_alloc_bytes(payload);
  for (int i = 0; i < n && ok; ++i) {
    ok = cudaSetDevice(device0 + i) == cudaSuccess;
    for (int j = 0; j < n && ok; ++j) {
      if (i == j) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(device0 + j, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else ok = e == cudaSuccess;
    }
    if (ok) ok = cudaMalloc(reinterpret_cast<void**>(&base[i]), total) == cudaSuccess && cudaMemset(base[i], 0, total) == cudaSuccess;
  }
  for (int i = 0; i < n && ok; ++i) { cudaSetDevice(device0 + i); ok = cudaDeviceSynchronize() == cudaSuccess; }
  if (!ok) {
    for (int i = 0; i < n; ++i) if (base[i]) { cudaSetDevice(device0 + i); cudaFree(base[i]); }
    cudaGetLastError();
    cudaSetDevice(cur);
    return false;
  }
  for (int r = 0; r < n; ++r) {
    PeerExchange* px = new PeerExchange;
    px->world = n; px->rank = r; px->device = device0 + r; px->bytes = payload; px->ipc = false;
    for (int p = 0; p < n; ++p) { px->base[p] = base[p]; peer_set_layout(px, p); }
    (*out)[r] = px;
  }
  cudaSetDevice(cur);
  return true;
}

}  // namespace

namespace {

double wall() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Contiguous landmark ranges balanced by observation count.
// Same rule as Python's pba_b200.partition_landmarks.
void partition_landmarks(const int64_t* lm_obs_ptr, int n_lm, int world, std::vector<int>& bounds) {
  bounds.assign(world + 1, 0);
  const int64_t total = lm_obs_ptr[n_lm];
  for (int r = 1; r < world; ++r) {
    const int64_t target = (total * r) / world;
    int b = int(std::lower_bound(lm_obs_ptr, lm_obs_ptr + n_lm + 1, target) - lm_obs_ptr);
    b = std::min(std::max(b, bounds[r - 1]), n_lm);
    bounds[r] = b;
  }
  bounds[world] = n_lm;
}

pba_status validate(const pba_problem* p, const pba_options* o) {
  if (!p || !o) return PBA_ERR_INVALID_ARGUMENT;
  if (o->optimize_intrinsics) return PBA_ERR_UNSUPPORTED;  // map_utils.h:339
  if (p->mode != PBA_MODE_GEOMETRIC && p->mode != PBA_MODE_PHOTOMETRIC) return PBA_ERR_INVALID_ARGUMENT;
  if (p->n_poses < 0 || p->n_calib < 0 || p->n_landmarks < 0 || p->n_obs < 0) return PBA_ERR_INVALID_ARGUMENT;
  if (p->n_poses > 0 && (!p->poses || !p->pose_calib)) return PBA_ERR_INVALID_ARGUMENT;
  if (p->n_calib > 0 && (!p->calib_model || !p->intrinsics)) return PBA_ERR_INVALID_ARGUMENT;
  if (p->n_landmarks > 0 && (!p->inv_depth || !p->lm_host || !p->lm_host_uv)) return PBA_ERR_INVALID_ARGUMENT;
  if (!p->lm_obs_ptr) return PBA_ERR_INVALID_ARGUMENT;
  if (p->n_obs > 0 && !p->obs_target) return PBA_ERR_INVALID_ARGUMENT;
  if (p->lm_obs_ptr[0] != 0 || p->lm_obs_ptr[p->n_landmarks] != p->n_obs) return PBA_ERR_INVALID_ARGUMENT;
  if (o->use_huber && !(o->huber_parameter > 0.0)) return PBA_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < p->n_poses; ++i)
    if (p->pose_calib[i] < 0 || p->pose_calib[i] >= p->n_calib) return PBA_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < p->n_calib; ++i)
    if (p->calib_model[i] < 0 || p->calib_model[i] > PBA_CAM_EUCM) return PBA_ERR_UNSUPPORTED;  // from_data aborts (camera_models.h:469)
  for (int l = 0; l < p->n_landmarks; ++l)  // offsets first: the parallel scan below indexes with them
    if (p->lm_obs_ptr[l + 1] < p->lm_obs_ptr[l] || p->lm_obs_ptr[l + 1] > p->n_obs) return PBA_ERR_INVALID_ARGUMENT;
  // the per-observation index checks (host / target in range, target != host) ride on the observation scan
  // of analyze_cameras: one pass over the 4 B x n_obs target table instead of two
  if (p->mode == PBA_MODE_GEOMETRIC) {
    if (p->n_obs > 0 && !p->obs_uv) return PBA_ERR_INVALID_ARGUMENT;
  } else {
    if (p->n_poses > 0 && !p->images && !p->image_ptrs) return PBA_ERR_INVALID_ARGUMENT;
    if (p->width < 2 || p->height < 2 || p->pitch < p->width) return PBA_ERR_INVALID_ARGUMENT;
    if (p->width > 65535 || p->height > 32767) return PBA_ERR_UNSUPPORTED;  // the kernels pack a pixel cell into 32 bits
  }
  return PBA_OK;
}


// ---- pinned host staging (set-up uploads) ----
// The observation-sized index arrays pba_create builds go to the device from PINNED host memory: the
// copies run at PCIe speed and asynchronously, under the host work that follows.  Pinning hundreds of
// MB costs more than a solve, so released buffers stay in a per-process pool (like the device arena).
struct PinnedPool {
  std::mutex mu;
  struct Buf { void* p; size_t bytes; };
  std::vector<Buf> free_list;
  void* acquire(size_t bytes, size_t* got) {
    // sizes are rounded up to a quarter-octave grid (and 1 MB): the shards of a multi-GPU solve and the
    // successive solves of a growing map ask for slightly different sizes, and a buffer that is a little
    // too small for the next request would otherwise be pinned again (~0.3 ms per MB)
    size_t step = size_t(1) << 20;
    while (step * 8 <= bytes) step <<= 1;
    const size_t want = (bytes + step - 1) / step * step;
    {
      std::lock_guard<std::mutex> lock(mu);
      int best = -1;
      for (int i = 0; i < int(free_list.size()); ++i)
        if (free_list[i].bytes == want) { best = i; break; }  // exact size class: no request steals a larger one's buffer
      if (best >= 0) {
        Buf b = free_list[best];
        free_list.erase(free_list.begin() + best);
        *got = b.bytes;
        return b.p;
      }
    }
    void* q = nullptr;
    const double t_alloc = wall_s();
    const cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocPortable);
    ++g_n_host_alloc; g_us_host_alloc += int64_t(1e6 * (wall_s() - t_alloc)); g_mb_host_alloc += int64_t(want >> 20);
    if (e != cudaSuccess) { cudaGetLastError(); *got = 0; return nullptr; }
    *got = want;
    return q;
  }
  void release(void* p, size_t bytes) {
    std::lock_guard<std::mutex> lock(mu);
    free_list.push_back({p, bytes});
  }
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    for (Buf& b : free_list) cudaFreeHost(b.p);
    free_list.clear();
  }
};
PinnedPool g_pinned;

// Array in pinned staging memory (pageable `new[]` if pinning fails).  Not value-initialised.
template <class T>
struct StageVec {
  T* p = nullptr;
  size_t n = 0, cap_bytes = 0;
  bool pinned = false;
  StageVec() {}
  explicit StageVec(size_t count) { resize(count); }
  StageVec(const StageVec&) = delete;
  StageVec& operator=(const StageVec&) = delete;
  ~StageVec() { reset(); }
  void reset() {
    if (p) { if (pinned) g_pinned.release(p, cap_bytes); else delete[] p; }
    p = nullptr; n = 0;
  }
  void resize(size_t count) {
    reset();
    n = count;
    if (!count) return;
    p = static_cast<T*>(g_pinned.acquire(count * sizeof(T), &cap_bytes));
    pinned = p != nullptr;
    if (!p) p = new T[count];
  }
  void fill(const T& v) { for (size_t i = 0; i < n; ++i) p[i] = v; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  const T* data() const { return p; }
  T* data() { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
};

// ---- camera layout: which poses are optimised, in which ORDER, and the RCS block pattern ----
// Replaces the parameter-block ordering + block-structure detection of Ceres' preprocessor
// (trust_region_preprocessor.cc:373, schur_complement_solver.cc:250-297; Ceres orders the reduced
// camera system with AMD inside its sparse Cholesky).  Here the exact solvers want a small BANDWIDTH:
// when the keyframes' natural order does not give one (loop closures, maps whose keyframes are not in
// temporal order) the free cameras are renumbered by reverse Cuthill-McKee on the covisibility graph.
// The block pattern is the union over host keyframes h of all pairs within {h} + targets(h) — a
// superset of the per-landmark pairs Ceres uses, exactly the blocks the per-host-group Schur products write.
struct CameraLayout {
  std::vector<int> slot;               // pose -> RCS slot or -1 (constant / unused)
  std::vector<uint8_t> affine_active;
  int n_slots = 0;
  int64_t n_active_lm = 0;
  std::vector<std::vector<int>> adj;   // per slot a: ascending slots b >= a with an RCS block (a, b), diagonal included
  int bandwidth = 0;                   // max (b - a) in blocks, final order
  int bandwidth_natural = 0;           // the same in the keyframes' own order
  bool reordered = false;
};

// new position -> node, reverse Cuthill-McKee; nb = symmetric neighbour lists without self loops
std::vector<int> rcm_order(const std::vector<std::vector<int>>& nb) {
  const int n = int(nb.size());
  std::vector<int> order;
  order.reserve(n);
  std::vector<int> level(n, -1);
  std::vector<char> done(n, 0);
  auto bfs_levels = [&](int start, std::vector<int>& visited) {  // fills level[] for the component, returns the depth
    visited.clear();
    visited.push_back(start);
    level[start] = 0;
    for (size_t q = 0; q < visited.size(); ++q)
      for (int v : nb[visited[q]])
        if (level[v] < 0) { level[v] = level[visited[q]] + 1; visited.push_back(v); }
    return level[visited.back()];
  };
  std::vector<int> comp, scratch;
  for (int seed = 0; seed < n; ++seed) {
    if (done[seed]) continue;
    // component of `seed`; start from a pseudo-peripheral node (George-Liu: repeat BFS from a
    // minimum-degree node of the last level while the depth grows)
    int start = seed;
    int depth = bfs_levels(start, comp);
    for (int v : comp) if (nb[v].size() < nb[start].size()) start = v;  // first guess: minimum degree
    for (int v : comp) level[v] = -1;
    depth = bfs_levels(start, comp);
    for (int iter = 0; iter < 8; ++iter) {
      int cand = -1;
      for (int v : comp)
        if (level[v] == depth && (cand < 0 || nb[v].size() < nb[cand].size())) cand = v;
      for (int v : comp) level[v] = -1;
      const int d2 = bfs_levels(cand, scratch);
      if (d2 > depth) { start = cand; depth = d2; comp = scratch; }
      else { for (int v : scratch) level[v] = -1; bfs_levels(start, comp); break; }
    }
    for (int v : comp) level[v] = -1;
    // Cuthill-McKee from `start`: neighbours in order of increasing degree
    const size_t first = order.size();
    order.push_back(start);
    done[start] = 1;
    for (size_t q = first; q < order.size(); ++q) {
      scratch.clear();
      for (int v : nb[order[q]]) if (!done[v]) { done[v] = 1; scratch.push_back(v); }
      std::sort(scratch.begin(), scratch.end(), [&](int a, int b) {
        return nb[a].size() != nb[b].size() ? nb[a].size() < nb[b].size() : a < b;
      });
      order.insert(order.end(), scratch.begin(), scratch.end());
    }
  }
  std::reverse(order.begin(), order.end());
  return order;
}

// Returns PBA_ERR_INVALID_ARGUMENT for a host / target index out of range or a landmark observed by its own
// host keyframe (the part of the input validation that has to look at every observation).
pba_status analyze_cameras(const pba_problem* p, int nthr, int max_band_blocks, CameraLayout* L) {
  const bool timing = getenv("PBA_TIMING") != nullptr;
  double t_mark = wall();
  auto mark = [&](const char* what) {
    if (!timing) return;
    const double now = wall();
    fprintf(stderr, "[analyze_cameras] %-24s %8.1f ms\n", what, 1e3 * (now - t_mark));
    t_mark = now;
  };
  const bool photo = p->mode == PBA_MODE_PHOTOMETRIC;
  const int np = p->n_poses;
  std::vector<uint8_t> used(np, 0), is_target(np, 0);
  std::vector<std::vector<int>> host_targets(np);  // distinct targets of the landmarks hosted by a keyframe
  int64_t n_active = 0;
  int bad = 0;
#pragma omp parallel num_threads(nthr) reduction(+ : n_active) reduction(| : bad)
  {
    std::vector<uint8_t> u(np, 0), t(np, 0);
    std::vector<int> stamp(np, -1);  // stamp[target] = host while consecutive landmarks share the host
    std::vector<std::vector<int>> mine(np);
#pragma omp for schedule(static)
    for (int l = 0; l < p->n_landmarks; ++l) {
      const int64_t k0 = p->lm_obs_ptr[l], k1 = p->lm_obs_ptr[l + 1];
      const int hst = p->lm_host[l];
      if (hst < 0 || hst >= np) { bad |= 1; continue; }
      if (k1 == k0) continue;
      u[hst] = 1;
      ++n_active;
      for (int64_t k = k0; k < k1; ++k) {
        const int tg = p->obs_target[k];
        if (unsigned(tg) >= unsigned(np) || tg == hst) { bad |= 1; continue; }
        if (stamp[tg] != hst) { stamp[tg] = hst; mine[hst].push_back(tg); u[tg] = 1; t[tg] = 1; }
      }
    }
#pragma omp critical
    for (int i = 0; i < np; ++i) {
      used[i] |= u[i]; is_target[i] |= t[i];
      host_targets[i].insert(host_targets[i].end(), mine[i].begin(), mine[i].end());
    }
  }
  mark("observation scan");
  if (bad) return PBA_ERR_INVALID_ARGUMENT;
  for (auto& v : host_targets) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
  L->n_active_lm = n_active;
  L->slot.assign(np, -1);
  L->affine_active.assign(np, 0);
  L->n_slots = 0;
  for (int i = 0; i < np; ++i) {
    const bool fixed = p->pose_fixed && p->pose_fixed[i];
    if (!fixed && used[i]) L->slot[i] = L->n_slots++;
    L->affine_active[i] = photo && !fixed && is_target[i];
  }
  const int ns = L->n_slots;
  // covisibility graph on the provisional (natural) slots
  std::vector<std::vector<int>> nb(ns);
  std::vector<int> cams;
  for (int hst = 0; hst < np; ++hst) {
    if (host_targets[hst].empty()) continue;
    cams.clear();
    if (L->slot[hst] >= 0) cams.push_back(L->slot[hst]);
    for (int tg : host_targets[hst]) if (L->slot[tg] >= 0) cams.push_back(L->slot[tg]);
    for (size_t i = 0; i < cams.size(); ++i)
      for (size_t j = i + 1; j < cams.size(); ++j) { nb[cams[i]].push_back(cams[j]); nb[cams[j]].push_back(cams[i]); }
  }
  int bw_nat = 0;
  for (int a = 0; a < ns; ++a) {
    std::sort(nb[a].begin(), nb[a].end());
    nb[a].erase(std::unique(nb[a].begin(), nb[a].end()), nb[a].end());
    for (int b : nb[a]) bw_nat = std::max(bw_nat, std::abs(a - b));
  }
  mark("covisibility graph");
  L->bandwidth_natural = bw_nat;
  L->bandwidth = bw_nat;
  L->reordered = false;
  std::vector<int> newpos(ns);
  std::iota(newpos.begin(), newpos.end(), 0);
  if (bw_nat > max_band_blocks && ns > 2) {
    // the natural order is too wide for the banded solvers: try reverse Cuthill-McKee
    const std::vector<int> order = rcm_order(nb);
    std::vector<int> pos(ns);
    for (int i = 0; i < ns; ++i) pos[order[i]] = i;
    int bw = 0;
    for (int a = 0; a < ns; ++a)
      for (int b : nb[a]) bw = std::max(bw, std::abs(pos[a] - pos[b]));
    if (bw < bw_nat) {
      newpos = pos;
      L->bandwidth = bw;
      L->reordered = true;
      for (int i = 0; i < np; ++i) if (L->slot[i] >= 0) L->slot[i] = newpos[L->slot[i]];
    }
  }
  L->adj.assign(ns, std::vector<int>());
  for (int a = 0; a < ns; ++a) {
    const int pa = newpos[a];
    L->adj[pa].push_back(pa);
    for (int b : nb[a]) {
      const int pb = newpos[b];
      if (pb > pa) L->adj[pa].push_back(pb);
    }
  }
  for (auto& v : L->adj) std::sort(v.begin(), v.end());
  mark("ordering + pattern");
  return PBA_OK;
}

constexpr int kChunkObs = 1024;

// `shared` (single-process multi-GPU): the global camera layout, computed once by the caller with all host
// cores instead of once per rank thread; the problem has been validated by the caller as well.
pba_status create_impl(const pba_problem* p, const pba_options* o, int rank, int world, Handle** out,
                       const CameraLayout* shared = nullptr) {
  const double t_begin = wall();
  pba_status st = shared ? PBA_OK : validate(p, o);
  if (st != PBA_OK) return st;
  if (world < 1 || rank < 0 || rank >= world) return PBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return PBA_ERR_NO_DEVICE; }
  if (o->device < 0 || o->device >= ndev) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(o->device));

  const bool timing = getenv("PBA_TIMING") != nullptr;
  double t_mark = t_begin;
  auto mark = [&](const char* what) {
    if (!timing) return;
    const double now = wall();
    fprintf(stderr, "[pba_create %d/%d] %-28s %8.1f ms\n", rank, world, what, 1e3 * (now - t_mark));
    t_mark = now;
  };
  std::unique_ptr<Handle> hh(new Handle);
  Handle* h = hh.get();
  h->opt = *o;
  h->rank = rank; h->world = world; h->device = o->device;
  h->arena.device = o->device;
  ArenaScope arena_scope(&h->arena);  // every DevBuf allocated below lives in the handle's arena
  h->stats.profile = o->profile;
  PBA_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  Sizes& z = h->sz;
  const bool photo = p->mode == PBA_MODE_PHOTOMETRIC;
  z.mode = p->mode; z.R = photo ? 8 : 2; z.C = photo ? 15 : 13; z.cd = photo ? 8 : 6;
  z.n_poses = p->n_poses; z.n_calib = p->n_calib;
  z.width = p->width; z.height = p->height; z.pitch = p->pitch;
  const int cd = z.cd;
  h->uniform_model = p->n_calib > 0 ? p->calib_model[0] : -1;
  for (int i = 1; i < p->n_calib; ++i)
    if (p->calib_model[i] != p->calib_model[0]) h->uniform_model = -1;

  mark("validate + device init");
  // One process per GPU shares the host cores: each rank takes its share of the OpenMP threads.
  const int nthr = std::max(1, omp_get_num_procs() / std::max(1, world));
  // ---- this rank's landmark range (contiguous, balanced by observation count) ----
  std::vector<int> bounds;
  partition_landmarks(p->lm_obs_ptr, p->n_landmarks, world, bounds);
  const int lm_lo = bounds[rank], lm_hi = bounds[rank + 1];
  h->first_landmark = lm_lo;
  z.n_lm = lm_hi - lm_lo;
  const int64_t obs_lo = p->lm_obs_ptr[lm_lo];
  z.n_obs = p->lm_obs_ptr[lm_hi] - obs_lo;
  z.ld = (z.n_obs + 31) / 32 * 32;
  const int n_lm = z.n_lm;
  const int64_t n = z.n_obs;

  // ---- keyframe upload, started first and run by a helper thread on its own stream: 8-bit rows go to the
  //      device in batches and are expanded there into the quad layout the evaluation kernels gather from.
  //      It overlaps all the host-side ordering below.  The staging area is the head of the Jacobian buffer
  //      (not needed before the first evaluation). ----
  PBA_CUDA_OK(h->J.alloc(size_t(z.ld) * z.R * (z.C + 1 - 6)));  // stored planes: pba_internal.h
  std::thread image_thread;
  pba_status image_status = PBA_OK;
  DevBuf<uint8_t> own_stage;
  struct JoinGuard { std::thread& t; ~JoinGuard() { if (t.joinable()) t.join(); } } join_guard{image_thread};
  if (photo) {
    z.image_stride = int64_t(p->width) * p->height;  // pixels
    const size_t img_bytes = size_t(p->pitch) * p->height;
    PBA_CUDA_OK(h->quads.alloc(size_t(z.image_stride) * p->n_poses));
    const int batch = std::max(1, std::min(p->n_poses, 256));
    uint8_t* stage = reinterpret_cast<uint8_t*>(h->J.p);
    if (h->J.n * sizeof(double) < img_bytes * batch) {
      PBA_CUDA_OK(own_stage.alloc(img_bytes * batch));
      stage = own_stage.p;
    }
    const int dev = o->device;
    image_thread = std::thread([=, &image_status]() {
      auto fail = [&](cudaError_t e) { image_status = map_cuda(e); };
      cudaError_t e = cudaSetDevice(dev);
      if (e != cudaSuccess) return fail(e);
      cudaStream_t is = nullptr;
      if ((e = cudaStreamCreateWithFlags(&is, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
      for (int f0 = 0; f0 < p->n_poses && image_status == PBA_OK; f0 += batch) {
        const int cnt = std::min(batch, p->n_poses - f0);
        if (!p->image_ptrs && p->image_stride == int64_t(img_bytes)) {
          e = cudaMemcpyAsync(stage, p->images + size_t(f0) * img_bytes, img_bytes * cnt, cudaMemcpyHostToDevice, is);
        } else {
          for (int i = 0; i < cnt && e == cudaSuccess; ++i) {
            const uint8_t* src = p->image_ptrs ? p->image_ptrs[f0 + i] : p->images + size_t(f0 + i) * p->image_stride;
            e = cudaMemcpyAsync(stage + size_t(i) * img_bytes, src, img_bytes, cudaMemcpyHostToDevice, is);
          }
        }
        if (e != cudaSuccess) { fail(e); break; }
        const pba_status ls = launch_build_quads(h, is, stage, f0, cnt);
        if (ls != PBA_OK) image_status = ls;
      }
      e = cudaStreamSynchronize(is);
      if (e != cudaSuccess && image_status == PBA_OK) fail(e);
      cudaStreamDestroy(is);
    });
    h->stats.launches[K_INIT_LM] += (p->n_poses + batch - 1) / batch;
  }
  mark("sizes + start of the keyframe upload");

  // ---- global layout (identical on every rank): which parameter blocks survive Ceres' reduced program,
  //      the order of the free cameras, the RCS block pattern ----
  CameraLayout layout;
  if (shared) layout = *shared;
  else if ((st = analyze_cameras(p, nthr, 128 / cd, &layout)) != PBA_OK) return st;
  h->n_active_lm += layout.n_active_lm;
  h->n_obs_global = p->n_obs;
  h->slot = layout.slot;
  h->affine_active = layout.affine_active;
  z.n_slots = layout.n_slots;
  z.dim = z.n_slots * cd;
  const std::vector<int>& slot = h->slot;
  std::vector<std::vector<int>>& adj = layout.adj;
  mark("camera layout + RCS pattern");
  std::vector<int64_t> adj_ptr(z.n_slots + 1, 0);
  for (int a = 0; a < z.n_slots; ++a) adj_ptr[a + 1] = adj_ptr[a] + int64_t(adj[a].size());  // adj[a][0] == a
  z.n_blocks = adj_ptr[z.n_slots];
  h->blk_row.resize(z.n_blocks); h->blk_col.resize(z.n_blocks); h->diag_blk.assign(z.n_slots, 0);
  for (int a = 0; a < z.n_slots; ++a)
    for (size_t k = 0; k < adj[a].size(); ++k) {
      h->blk_row[adj_ptr[a] + k] = a;
      h->blk_col[adj_ptr[a] + k] = adj[a][k];
      if (adj[a][k] == a) h->diag_blk[a] = int(adj_ptr[a] + k);
    }
  for (int64_t b = 0; b < z.n_blocks; ++b) h->rcs_bandwidth = std::max(h->rcs_bandwidth, h->blk_col[b] - h->blk_row[b]);
  auto block_of = [&](int a, int b) -> int64_t {  // a <= b
    const auto it = std::lower_bound(adj[a].begin(), adj[a].end(), b);
    return adj_ptr[a] + (it - adj[a].begin());
  };

  // ---- local shard: landmarks ordered by host, observations by (host,target) edge ----

  h->lm_order.resize(n_lm);
  bool hosts_sorted = true;
  {
    int unsorted = 0;
#pragma omp parallel for num_threads(nthr) schedule(static) reduction(| : unsorted)
    for (int i = 0; i < n_lm; ++i) {
      h->lm_order[i] = i;
      if (i > 0 && p->lm_host[lm_lo + i - 1] > p->lm_host[lm_lo + i]) unsorted |= 1;
    }
    hosts_sorted = !unsorted;
    if (!hosts_sorted)
      std::stable_sort(h->lm_order.begin(), h->lm_order.end(),
                       [&](int a, int b) { return p->lm_host[lm_lo + a] < p->lm_host[lm_lo + b]; });
  }
  // lm-major observation index space k (internal landmark order); a host group's
  // observations are contiguous in k, and the edge order only permutes inside a group
  // observation-/landmark-sized tables are built in pinned staging memory (StageVec) and uploaded asynchronously;
  // they live until the final synchronisation of this function
  StageVec<int64_t> lm_ptr(size_t(n_lm) + 1);
  std::vector<int> grp_lm_ptr;  // host groups = runs of equal host
  if (hosts_sorted) {
    // the usual case (landmarks arrive grouped by host): the caller's own order, no prefix sum to redo;
    // the run starts are collected per thread and concatenated in thread order
    std::vector<std::vector<int>> starts(nthr);
#pragma omp parallel num_threads(nthr)
    {
      const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
      const int lo = int(int64_t(n_lm) * tid / nt), hi = int(int64_t(n_lm) * (tid + 1) / nt);
      std::vector<int>& mine = starts[tid];
      for (int li = lo; li < hi; ++li) {
        lm_ptr[li] = p->lm_obs_ptr[lm_lo + li] - obs_lo;
        if (li == 0 || p->lm_host[lm_lo + li] != p->lm_host[lm_lo + li - 1]) mine.push_back(li);
      }
    }
    lm_ptr[n_lm] = n;
    for (auto& v : starts) grp_lm_ptr.insert(grp_lm_ptr.end(), v.begin(), v.end());
  } else {
    lm_ptr[0] = 0;
    for (int li = 0; li < n_lm; ++li) {
      const int l = lm_lo + h->lm_order[li];
      lm_ptr[li + 1] = lm_ptr[li] + (p->lm_obs_ptr[l + 1] - p->lm_obs_ptr[l]);
    }
    for (int li = 0; li < n_lm; ++li)
      if (li == 0 || p->lm_host[lm_lo + h->lm_order[li]] != p->lm_host[lm_lo + h->lm_order[li - 1]]) grp_lm_ptr.push_back(li);
  }
  grp_lm_ptr.push_back(n_lm);
  z.n_groups = int(grp_lm_ptr.size()) - 1;
  const int G = z.n_groups;

  // pass A (parallel over groups): distinct targets of every group, ascending, with counts
  std::vector<std::vector<std::pair<int, int64_t>>> grp_tg(G);
#pragma omp parallel num_threads(nthr)
  {
    std::vector<int64_t> cnt(p->n_poses, 0);
    std::vector<int> touched;
#pragma omp for schedule(dynamic, 16)
    for (int g = 0; g < G; ++g) {
      touched.clear();
      for (int li = grp_lm_ptr[g]; li < grp_lm_ptr[g + 1]; ++li) {
        const int l = lm_lo + h->lm_order[li];
        for (int64_t q = p->lm_obs_ptr[l]; q < p->lm_obs_ptr[l + 1]; ++q) {
          const int t = p->obs_target[q];
          if (cnt[t]++ == 0) touched.push_back(t);
        }
      }
      std::sort(touched.begin(), touched.end());
      grp_tg[g].reserve(touched.size());
      for (int t : touched) { grp_tg[g].emplace_back(t, cnt[t]); cnt[t] = 0; }
    }
  }
  // edges (serial prefix): edge e of group g = (host_g, target), observations contiguous
  std::vector<int> edge_h, edge_t, grp_edge0(G + 1, 0);
  std::vector<int64_t> edge_ptr;
  for (int g = 0; g < G; ++g) {
    const int host = p->lm_host[lm_lo + h->lm_order[grp_lm_ptr[g]]];
    int64_t pos = lm_ptr[grp_lm_ptr[g]];
    grp_edge0[g] = int(edge_h.size());
    for (auto& tc : grp_tg[g]) {
      edge_h.push_back(host); edge_t.push_back(tc.first); edge_ptr.push_back(pos);
      pos += tc.second;
    }
  }
  grp_edge0[G] = int(edge_h.size());
  edge_ptr.push_back(n);
  z.n_edges = int(edge_h.size());

  // host-group camera lists (slots, ascending) and buffer offsets (serial prefix, small)
  std::vector<int> grp_cam_ptr(1, 0), grp_cams;
  std::vector<int64_t> grp_w_off(G), grp_part_off(G);
  int64_t w_total = 0, part_total = 0;
  for (int g = 0; g < G; ++g) {
    const int host = p->lm_host[lm_lo + h->lm_order[grp_lm_ptr[g]]];
    const size_t c0 = grp_cams.size();
    if (slot[host] >= 0) grp_cams.push_back(slot[host]);
    for (auto& tc : grp_tg[g]) if (slot[tc.first] >= 0) grp_cams.push_back(slot[tc.first]);
    std::sort(grp_cams.begin() + c0, grp_cams.end());
    grp_cams.erase(std::unique(grp_cams.begin() + c0, grp_cams.end()), grp_cams.end());
    const int c = int(grp_cams.size() - c0);
    grp_cam_ptr.push_back(int(grp_cams.size()));
    const int stride = 8 * (c + 1);
    h->max_w_stride = std::max(h->max_w_stride, stride);
    grp_w_off[g] = w_total;
    grp_part_off[g] = part_total;
    w_total += int64_t(grp_lm_ptr[g + 1] - grp_lm_ptr[g]) * stride;
    part_total += int64_t(c) * (c + 1) / 2 * cd * cd + int64_t(c) * cd;
  }
  h->schur_tile_l = schur_tile_l(h->max_w_stride);
  std::vector<int> syrk_work;
  schur_syrk_work(grp_cam_ptr, &syrk_work);
  h->n_syrk_work = int(syrk_work.size() / 2);

  // pass B (parallel over groups): place every observation at its edge-order position
  StageVec<int> obs_lm(n);  // every entry is written in pass B
  std::vector<int> edge_col(z.n_edges, -1);  // W column of an edge's target; obs_edge / obs_col are expanded on the device
  StageVec<int64_t> lm_pos(n);
  StageVec<int> lm_group(n_lm), lm_hostcol(n_lm), lm_w_stride(n_lm);
  StageVec<int64_t> lm_w_off(n_lm);
  h->obs_order.resize(n);
#pragma omp parallel num_threads(nthr)
  {
    std::vector<int64_t> cursor(p->n_poses, 0);
#pragma omp for schedule(dynamic, 16)
    for (int g = 0; g < G; ++g) {
      const int li0 = grp_lm_ptr[g], li1 = grp_lm_ptr[g + 1];
      const int host = p->lm_host[lm_lo + h->lm_order[li0]];
      const int c0 = grp_cam_ptr[g], c = grp_cam_ptr[g + 1] - c0;
      const int stride = 8 * (c + 1);
      int64_t pos = lm_ptr[li0];
      int e = grp_edge0[g];
      for (auto& tc : grp_tg[g]) {
        cursor[tc.first] = pos;
        pos += tc.second;
        const int sl = slot[tc.first];
        edge_col[e++] = sl >= 0 ? int(std::lower_bound(grp_cams.begin() + c0, grp_cams.begin() + c0 + c, sl) - (grp_cams.begin() + c0)) : -1;
      }
      const int hcol = slot[host] >= 0 ? int(std::lower_bound(grp_cams.begin() + c0, grp_cams.begin() + c0 + c, slot[host]) - (grp_cams.begin() + c0)) : -1;
      for (int li = li0; li < li1; ++li) {
        const int l = lm_lo + h->lm_order[li];
        lm_group[li] = g;
        lm_hostcol[li] = hcol;
        lm_w_off[li] = grp_w_off[g] + int64_t(li - li0) * stride;
        lm_w_stride[li] = stride;
        int64_t k = lm_ptr[li];
        for (int64_t q = p->lm_obs_ptr[l]; q < p->lm_obs_ptr[l + 1]; ++q, ++k) {
          const int t = p->obs_target[q];
          const int64_t at = cursor[t]++;
          obs_lm[at] = li;
          lm_pos[k] = at;
          h->obs_order[at] = q - obs_lo;
        }
      }
    }
  }
  grp_tg.clear();
  mark("edge ordering + host groups");

  // edge chunks (one CTA each in k_edge_gram)
  std::vector<int> chunk_edge;
  std::vector<int64_t> chunk_begin, chunk_end;
  for (int e = 0; e < z.n_edges; ++e)
    for (int64_t b = edge_ptr[e]; b < edge_ptr[e + 1]; b += kChunkObs) {
      chunk_edge.push_back(e); chunk_begin.push_back(b); chunk_end.push_back(std::min<int64_t>(b + kChunkObs, edge_ptr[e + 1]));
    }
  z.n_chunks = int(chunk_edge.size());
  mark("chunks");
  // per-block source lists for the RCS reduction (static)
  const int dir_stride = 3 * cd * cd + 2 * cd;
  std::vector<int64_t> dir_ptr(z.n_blocks + 1, 0), sch_ptr(z.n_blocks + 1, 0), vdir_ptr(z.n_slots + 1, 0), vsch_ptr(z.n_slots + 1, 0);
  std::vector<int64_t> dir_src, sch_src, vdir_src, vsch_src;
  {
    struct Src { int64_t key, off; };
    std::vector<Src> d, s, vd, vs;
    for (int q = 0; q < z.n_chunks; ++q) {
      const int e = chunk_edge[q];
      const int hs = slot[edge_h[e]], ts = slot[edge_t[e]];
      const int64_t base = int64_t(q) * dir_stride;
      if (hs >= 0) { d.push_back({block_of(hs, hs), base}); vd.push_back({hs, base + 3 * cd * cd}); }
      if (ts >= 0) { d.push_back({block_of(ts, ts), base + 2 * cd * cd}); vd.push_back({ts, base + 3 * cd * cd + cd}); }
      if (hs >= 0 && ts >= 0) d.push_back({block_of(std::min(hs, ts), std::max(hs, ts)), base + cd * cd});
    }
    for (int g = 0; g < z.n_groups; ++g) {
      const int c0 = grp_cam_ptr[g], c = grp_cam_ptr[g + 1] - c0;
      int64_t off = grp_part_off[g];
      for (int i = 0; i < c; ++i)
        for (int j = i; j < c; ++j) { s.push_back({block_of(grp_cams[c0 + i], grp_cams[c0 + j]), off}); off += cd * cd; }
      for (int i = 0; i < c; ++i) { vs.push_back({grp_cams[c0 + i], off}); off += cd; }
    }
    auto build = [](std::vector<Src>& v, int64_t nkeys, std::vector<int64_t>& ptr, std::vector<int64_t>& src) {
      // stable counting sort by key (keys are block / slot numbers): sources of a block stay in list order,
      // which fixes the summation order of k_rcs_reduce
      ptr.assign(nkeys + 1, 0);
      src.resize(v.size());
      for (const Src& e : v) ++ptr[e.key + 1];
      for (int64_t i = 0; i < nkeys; ++i) ptr[i + 1] += ptr[i];
      std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
      for (const Src& e : v) src[cur[e.key]++] = e.off;
    };
#pragma omp parallel sections num_threads(4)
    {
#pragma omp section
      build(d, z.n_blocks, dir_ptr, dir_src);
#pragma omp section
      build(s, z.n_blocks, sch_ptr, sch_src);
#pragma omp section
      build(vd, z.n_slots, vdir_ptr, vdir_src);
#pragma omp section
      build(vs, z.n_slots, vsch_ptr, vsch_src);
    }
  }
  // symmetric block-row CSR for the PCG
  std::vector<int> row_ptr(z.n_slots + 1, 0), row_blk, row_col;
  std::vector<uint8_t> row_trans;
  {
    std::vector<std::vector<std::array<int, 3>>> rows(z.n_slots);
    for (int64_t b = 0; b < z.n_blocks; ++b) {
      const int a = h->blk_row[b], c = h->blk_col[b];
      rows[a].push_back({int(b), c, 0});
      if (a != c) rows[c].push_back({int(b), a, 1});
    }
    for (int a = 0; a < z.n_slots; ++a) {
      std::sort(rows[a].begin(), rows[a].end(), [](const std::array<int, 3>& x, const std::array<int, 3>& y) { return x[1] < y[1]; });
      for (auto& e : rows[a]) { row_blk.push_back(e[0]); row_col.push_back(e[1]); row_trans.push_back(uint8_t(e[2])); }
      row_ptr[a + 1] = int(row_blk.size());
    }
  }

  mark("source lists + row CSR");
  // ---- upload static data ----
  cudaStream_t s = h->stream;
  auto up = [&](auto& buf, const auto& vec) { return buf.upload(vec, s); };
  std::vector<double> intr(p->intrinsics, p->intrinsics + size_t(8) * p->n_calib);
  std::vector<int> pose_calib(p->pose_calib, p->pose_calib + p->n_poses), calib_model(p->calib_model, p->calib_model + p->n_calib);
  PBA_CUDA_OK(up(h->intr, intr)); PBA_CUDA_OK(up(h->pose_calib, pose_calib)); PBA_CUDA_OK(up(h->calib_model, calib_model));
  PBA_CUDA_OK(up(h->d_slot, h->slot)); PBA_CUDA_OK(up(h->d_affine_active, h->affine_active));
  PBA_CUDA_OK(up(h->edge_h, edge_h)); PBA_CUDA_OK(up(h->edge_t, edge_t)); PBA_CUDA_OK(up(h->edge_ptr, edge_ptr));
  PBA_CUDA_OK(up(h->obs_lm, obs_lm));
  PBA_CUDA_OK(up(h->lm_ptr, lm_ptr)); PBA_CUDA_OK(up(h->lm_pos, lm_pos));
  PBA_CUDA_OK(up(h->lm_group, lm_group)); PBA_CUDA_OK(up(h->lm_hostcol, lm_hostcol));
  {
    // per-observation edge index and W column: constant along an edge, so expanded from the edge table on the device
    PBA_CUDA_OK(h->obs_edge.alloc(size_t(n))); PBA_CUDA_OK(h->obs_col.alloc(size_t(n)));
    const DeviceArena::Mark tmp_mark = h->arena.mark();  // d_edge_col is a set-up temporary
    DevBuf<int> d_edge_col;
    PBA_CUDA_OK(up(d_edge_col, edge_col));
    if ((st = launch_expand_edges(h, d_edge_col.p)) != PBA_OK) return st;
    d_edge_col.release();  // whatever takes this memory next is ordered behind the kernel on the same stream
    h->arena.rewind(tmp_mark);
  }
  PBA_CUDA_OK(up(h->chunk_edge, chunk_edge)); PBA_CUDA_OK(up(h->chunk_begin, chunk_begin)); PBA_CUDA_OK(up(h->chunk_end, chunk_end));
  PBA_CUDA_OK(up(h->grp_lm_ptr, grp_lm_ptr)); PBA_CUDA_OK(up(h->grp_cam_ptr, grp_cam_ptr)); PBA_CUDA_OK(up(h->grp_cams, grp_cams));
  PBA_CUDA_OK(up(h->grp_w_off, grp_w_off)); PBA_CUDA_OK(up(h->grp_part_off, grp_part_off));
  PBA_CUDA_OK(up(h->syrk_work, syrk_work));
  PBA_CUDA_OK(up(h->lm_w_off, lm_w_off)); PBA_CUDA_OK(up(h->lm_w_stride, lm_w_stride));
  PBA_CUDA_OK(up(h->d_blk_row, h->blk_row)); PBA_CUDA_OK(up(h->d_blk_col, h->blk_col)); PBA_CUDA_OK(up(h->d_diag_blk, h->diag_blk));
  PBA_CUDA_OK(up(h->blk_dir_ptr, dir_ptr)); PBA_CUDA_OK(up(h->blk_dir_src, dir_src));
  PBA_CUDA_OK(up(h->blk_sch_ptr, sch_ptr)); PBA_CUDA_OK(up(h->blk_sch_src, sch_src));
  PBA_CUDA_OK(up(h->vec_dir_ptr, vdir_ptr)); PBA_CUDA_OK(up(h->vec_dir_src, vdir_src));
  PBA_CUDA_OK(up(h->vec_sch_ptr, vsch_ptr)); PBA_CUDA_OK(up(h->vec_sch_src, vsch_src));
  PBA_CUDA_OK(up(h->row_ptr, row_ptr)); PBA_CUDA_OK(up(h->row_blk, row_blk)); PBA_CUDA_OK(up(h->row_col, row_col));
  PBA_CUDA_OK(up(h->row_trans, row_trans));
  if (h->rcs_bandwidth <= band_max_bw(cd)) {
    const int B = h->rcs_bandwidth + 1;
    std::vector<int> col_blk(size_t(z.n_slots) * B, -1);
    for (int64_t b = 0; b < z.n_blocks; ++b) col_blk[size_t(h->blk_row[b]) * B + (h->blk_col[b] - h->blk_row[b])] = int(b);
    PBA_CUDA_OK(up(h->d_col_blk, col_blk));
    PBA_CUDA_OK(h->band_L.alloc(size_t(z.n_slots) * B * cd * cd));
  }
  if ((st = bcr_setup(h)) != PBA_OK) return st;
  StageVec<int> lm_host(n_lm);
  StageVec<double> lm_uv(size_t(2) * n_lm), rho(n_lm), uv;
  {
#pragma omp parallel for num_threads(nthr) schedule(static)
    for (int li = 0; li < n_lm; ++li) {
      const int l = lm_lo + h->lm_order[li];
      lm_host[li] = p->lm_host[l];
      lm_uv[2 * li] = p->lm_host_uv[2 * l]; lm_uv[2 * li + 1] = p->lm_host_uv[2 * l + 1];
      rho[li] = p->inv_depth[l];
    }
    PBA_CUDA_OK(up(h->lm_host, lm_host)); PBA_CUDA_OK(up(h->lm_uv, lm_uv)); PBA_CUDA_OK(up(h->rho, rho));
    if (!photo) {
      uv.resize(size_t(2) * n);
#pragma omp parallel for num_threads(nthr) schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        const int64_t q = obs_lo + h->obs_order[i];
        uv[i] = p->obs_uv[2 * q]; uv[n + i] = p->obs_uv[2 * q + 1];
      }
      PBA_CUDA_OK(up(h->obs_uv, uv));
    }
  }
  {
    std::vector<double> poses(p->poses, p->poses + size_t(7) * p->n_poses);
    PBA_CUDA_OK(up(h->poses, poses));
    std::vector<double> aff(size_t(2) * p->n_poses, 0.0);
    if (photo && p->affine) aff.assign(p->affine, p->affine + size_t(2) * p->n_poses);
    PBA_CUDA_OK(up(h->affine, aff));  // pageable sources are staged by the driver before the call returns
  }
  mark("upload structure");
  // ---- work buffers ----
  const size_t nn = size_t(n);
  PBA_CUDA_OK(h->poses_c.alloc(size_t(7) * p->n_poses)); PBA_CUDA_OK(h->poses_best.alloc(size_t(7) * p->n_poses));
  PBA_CUDA_OK(h->affine_c.alloc(size_t(2) * p->n_poses)); PBA_CUDA_OK(h->affine_best.alloc(size_t(2) * p->n_poses));
  PBA_CUDA_OK(h->rho_c.alloc(n_lm)); PBA_CUDA_OK(h->rho_best.alloc(n_lm));
  PBA_CUDA_OK(h->lm_pat.alloc(size_t(n_lm) * (photo ? 32 : 4))); PBA_CUDA_OK(h->lm_ok.alloc(n_lm));
  PBA_CUDA_OK(h->edge_T.alloc(size_t(kEdgeStride) * z.n_edges));
  PBA_CUDA_OK(h->edge_M.alloc(size_t(36) * z.n_edges));
  PBA_CUDA_OK(h->orec.alloc(nn * 16));
  PBA_CUDA_OK(h->W.alloc(size_t(w_total)));
  if (w_total) PBA_CUDA_OK(cudaMemsetAsync(h->W.p, 0, sizeof(double) * size_t(w_total), s));  // unseen camera slots stay 0
  PBA_CUDA_OK(h->lm_c.alloc(n_lm)); PBA_CUDA_OK(h->lm_g.alloc(n_lm)); PBA_CUDA_OK(h->lm_scale.alloc(n_lm));
  PBA_CUDA_OK(h->lm_diag.alloc(n_lm)); PBA_CUDA_OK(h->lm_s2.alloc(n_lm)); PBA_CUDA_OK(h->lm_iete.alloc(n_lm));
  PBA_CUDA_OK(h->part_dir.alloc(size_t(z.n_chunks) * dir_stride)); PBA_CUDA_OK(h->part_sch.alloc(size_t(part_total)));
  PBA_CUDA_OK(h->rcs.alloc(size_t(z.n_blocks) * cd * cd + 3 * size_t(z.dim) + 2 + size_t(world)));  // + rcs_tail()
  PBA_CUDA_OK(h->rcs_B.alloc(size_t(z.n_blocks) * cd * cd + size_t(z.dim)));
  PBA_CUDA_OK(h->cam_scale.alloc(z.dim)); PBA_CUDA_OK(h->cam_diag.alloc(z.dim)); PBA_CUDA_OK(h->cam_D2.alloc(z.dim));
  PBA_CUDA_OK(h->y_cam.alloc(z.dim)); PBA_CUDA_OK(h->d_cam.alloc(z.dim)); PBA_CUDA_OK(h->d_rho.alloc(n_lm));
  PBA_CUDA_OK(h->blk_inv.alloc(size_t(z.n_slots) * cd * cd));
  h->pcg_grid = pcg_max_grid(h->device);
  PBA_CUDA_OK(h->pcg_ws.alloc(4 * size_t(z.dim) + 3 * size_t(h->pcg_grid) + 8));
  const size_t red = std::max<size_t>({size_t(eval_grid(n)) + 8, 2 * ((size_t(p->n_poses) + n_lm + 255) / 256) + 8,
                                       size_t(n_lm + 127) / 128 + size_t(z.n_blocks + 15) / 16 + 8});
  PBA_CUDA_OK(h->red_ws.alloc(std::max<size_t>(red, (nn + 255) / 256 + 8)));
  PBA_CUDA_OK(h->red_mid.alloc(kReduceMid));
  PBA_CUDA_OK(h->scalars.alloc(S_NUM)); PBA_CUDA_OK(h->chol_fail.alloc(1));
  PBA_CUDA_OK(cudaMemsetAsync(h->scalars.p, 0, sizeof(double) * S_NUM, s));
  PBA_CUDA_OK(cudaMemsetAsync(h->chol_fail.p, 0, sizeof(int), s));
  static_assert(S_NUM + 2 <= int(kScalarBlock), "scalar mirror block too small");
  h->h_scalars = acquire_scalar_block();
  if (!h->h_scalars) return PBA_ERR_OUT_OF_MEMORY;
  mark("allocate work buffers");
  if (image_thread.joinable()) image_thread.join();  // quads are complete (the helper synchronised its stream)
  if (image_status != PBA_OK) return image_status;
  own_stage.release();
  mark("wait for the keyframe upload");
  st = launch_init_landmarks(h);
  if (st != PBA_OK) return st;
  PBA_CUDA_OK(cudaStreamSynchronize(s));
  mark("init landmarks");
  if (timing)
    fprintf(stderr, "[pba_create %d/%d] total %.1f ms   (process so far: %d cudaMalloc %lld MB %.1f ms, %d cudaHostAlloc %lld MB %.1f ms)\n",
            rank, world, 1e3 * (wall() - t_begin), g_n_cuda_malloc.load(), (long long)g_mb_cuda_malloc.load(),
            1e-3 * g_us_cuda_malloc.load(), g_n_host_alloc.load(), (long long)g_mb_host_alloc.load(), 1e-3 * g_us_host_alloc.load());
  *out = hh.release();
  return PBA_OK;
}

// PBA_TIMING: host seconds spent waiting for the device in read_scalars (per handle, reset by minimize_impl);
// the rest of the minimizer's wall time is enqueue work on the host
pba_status read_scalars(Handle* h) {
  const double t_wait0 = wall();
  struct Acc { Handle* h; double t0; ~Acc() { h->t_wait += wall() - t0; } } acc{h, t_wait0};
  PBA_CUDA_OK(cudaMemcpyAsync(h->h_scalars, h->scalars.p, sizeof(double) * S_NUM, cudaMemcpyDeviceToHost, h->stream));
  int* fail = reinterpret_cast<int*>(h->h_scalars + S_NUM);
  PBA_CUDA_OK(cudaMemcpyAsync(fail, h->chol_fail.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->stats.resolve();
  return PBA_OK;
}

int pick_solver(const Handle* h, int requested) {
  int s = requested == PBA_SOLVER_AUTO ? h->opt.solver : requested;
  const bool band_ok = h->rcs_bandwidth <= band_max_bw(h->sz.cd);
  const bool bcr_ok = h->bcr_m > 0;
  const int general = h->sz.dim <= h->opt.cholesky_max_dim ? PBA_SOLVER_CHOLESKY : PBA_SOLVER_PCG;
  // AUTO: exact solvers that exploit windowed covisibility first (parallel cyclic reduction,
  // then the sequential band factorisation), else dense Cholesky while it fits, else PCG.
  // Short chains go to the band factorisation: it costs ~4 us per keyframe on one SM, a BCR
  // level ~90 us (measured: 50 keyframes 0.20 vs 0.30 ms, 200 keyframes 0.83 vs 0.46 ms).
  if (s == PBA_SOLVER_AUTO)
    s = (bcr_ok && !(band_ok && h->sz.n_slots <= 64)) ? PBA_SOLVER_BCR : (band_ok ? PBA_SOLVER_BAND : general);
  if (s == PBA_SOLVER_BCR && !bcr_ok) s = band_ok ? PBA_SOLVER_BAND : general;
  if (s == PBA_SOLVER_BAND && !band_ok) s = general;
  return s;
}

pba_status solve_rcs(Handle* h, int solver) {
  h->last_solver = pick_solver(h, solver);
  if (h->last_solver == PBA_SOLVER_BCR) {
    // second generation (bcr2.cu) whenever its padded super blocks fit; PBA_BCR_V1=1 keeps the first for A/B runs
    static const bool force_v1 = getenv("PBA_BCR_V1") != nullptr;
    return (h->b2_nbk > 0 && !force_v1) ? launch_bcr2_rcs(h) : launch_bcr_rcs(h);
  }
  if (h->last_solver == PBA_SOLVER_BAND) return launch_band_rcs(h);
  return h->last_solver == PBA_SOLVER_CHOLESKY ? launch_cholesky_rcs(h) : launch_pcg_rcs(h);
}

// EvaluateGradientAndJacobian (trust_region_minimizer.cc:228-300) + everything
// that only depends on the new Jacobian: direct partials, landmark rows, the
// RCS for `radius`, gradient norms.
pba_status eval_jacobian_and_build(Handle* h, double radius) {
  // Multi-rank: this rank's cost and landmark gradient norms ride in the tail of the RCS all-reduce
  // (rcs_tail()); launch_gradient_norms then finishes S_COST / S_GMAX / S_GNORM2 from the reduced buffer.
  // One collective per evaluation (it used to be four).
  const bool multi = h->world > 1;
  pba_status st = launch_evaluate(h, true, h->poses.p, h->affine.p, h->rho.p, multi ? h->rcs_tail() : h->scalars.p + S_COST);
  if (st != PBA_OK) return st;
  h->have_jac = true;
  if ((st = launch_post_jacobian(h)) != PBA_OK) return st;
  if (multi && (st = launch_landmark_gradient_norms(h)) != PBA_OK) return st;
  if ((st = launch_build_rcs(h, radius, true, multi)) != PBA_OK) return st;
  return launch_gradient_norms(h);
}

pba_status copy_state(Handle* h, DevBuf<double>& dp, DevBuf<double>& da, DevBuf<double>& dr, const DevBuf<double>& sp,
                      const DevBuf<double>& sa, const DevBuf<double>& sr) {
  PBA_CUDA_OK(cudaMemcpyAsync(dp.p, sp.p, sizeof(double) * sp.n, cudaMemcpyDeviceToDevice, h->stream));
  PBA_CUDA_OK(cudaMemcpyAsync(da.p, sa.p, sizeof(double) * sa.n, cudaMemcpyDeviceToDevice, h->stream));
  if (sr.n) PBA_CUDA_OK(cudaMemcpyAsync(dr.p, sr.p, sizeof(double) * sr.n, cudaMemcpyDeviceToDevice, h->stream));
  return PBA_OK;
}

template <class T>
void swap_buf(DevBuf<T>& a, DevBuf<T>& b) { std::swap(a.p, b.p); std::swap(a.n, b.n); }

pba_status minimize_impl(Handle* h, pba_summary* sum) {
  const double t_start = wall();
  const pba_options& opt = h->opt;
  const Sizes& z = h->sz;
  pba_iteration* its = sum ? sum->iterations : nullptr;
  const int cap = sum ? sum->iterations_capacity : 0;
  if (sum) { memset(sum, 0, sizeof(*sum)); sum->iterations = its; sum->iterations_capacity = cap; }
  int n_it = 0;
  double t_prev_it = t_start;
  auto push = [&](pba_iteration& it) {
    // host clock after the iteration's last device read-back (every iteration ends in read_scalars)
    const double now = wall();
    it.iteration_time_in_seconds = now - t_prev_it;
    it.cumulative_time_in_seconds = now - t_start;
    t_prev_it = now;
    if (its && n_it < cap) its[n_it] = it;
    ++n_it;
  };
  KernelStats& ks = h->stats;
  const int64_t launches0 = std::accumulate(ks.launches, ks.launches + K_NUM, int64_t(0));

  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  double x_cost = 0, x_norm = -1.0, minimum_cost = std::numeric_limits<double>::max();
  int termination = PBA_NO_CONVERGENCE;
  char message[256] = "";
  int num_successful = 0, num_unsuccessful = 0, consecutive_invalid = 0;
  int n_jac = 0, n_res = 0, n_lin = 0, n_inexact = 0;
  pba_status st;
  double* hs = h->h_scalars;
  const int* chol_fail = reinterpret_cast<const int*>(hs + S_NUM);

  h->scale_ready = false;  // Jacobi scaling is computed at iteration 0 of every solve
  h->t_wait = 0.0;
  pba_iteration it;
  memset(&it, 0, sizeof(it));
  if ((st = eval_jacobian_and_build(h, radius)) != PBA_OK) return st;
  if ((st = read_scalars(h)) != PBA_OK) return st;
  ++n_jac; ++n_res;
  x_cost = hs[S_COST];
  bool ok = std::isfinite(x_cost);
  if (!ok) {
    termination = PBA_FAILURE;
    snprintf(message, sizeof(message), "Residual and Jacobian evaluation failed.");
  }
  const double initial_cost = x_cost;
  it.iteration = 0; it.cost = x_cost; it.step_is_valid = 1; it.step_is_successful = 1;
  it.gradient_max_norm = hs[S_GMAX]; it.gradient_norm = sqrt(hs[S_GNORM2]);
  double current_cost_se = x_cost, min_iteration_cost = x_cost;
  bool have_best = false;

  while (ok) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (trust_region_minimizer.cc:311-359)
    if (it.step_is_successful) {
      ++num_successful;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        if ((st = copy_state(h, h->poses_best, h->affine_best, h->rho_best, h->poses, h->affine, h->rho)) != PBA_OK) return st;
        have_best = true;
      }
    } else {
      ++num_unsuccessful;
    }
    it.trust_region_radius = radius;
    min_iteration_cost = std::min(min_iteration_cost, it.cost);
    push(it);
    if (it.iteration >= opt.max_num_iterations) {
      snprintf(message, sizeof(message), "Maximum number of iterations reached. Number of iterations: %d.", it.iteration);
      termination = PBA_NO_CONVERGENCE; break;
    }
    if (it.gradient_max_norm <= opt.gradient_tolerance) {
      snprintf(message, sizeof(message), "Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt.gradient_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    if (radius <= opt.min_trust_region_radius) {
      snprintf(message, sizeof(message), "Minimum trust region radius reached.");
      termination = PBA_CONVERGENCE; break;
    }
    const double prev_gmax = it.gradient_max_norm, prev_gnorm = it.gradient_norm;
    const int next = it.iteration + 1;
    memset(&it, 0, sizeof(it));
    it.iteration = next;

    // ---- ComputeTrustRegionStep; the RCS for `radius` is already on the device ----
    if ((st = solve_rcs(h, PBA_SOLVER_AUTO)) != PBA_OK) return st;
    if ((st = launch_backsub(h)) != PBA_OK) return st;
    // speculative: candidate point and its cost (needed unless the step is invalid)
    if ((st = launch_retract(h)) != PBA_OK) return st;
    if ((st = launch_evaluate(h, false, h->poses_c.p, h->affine_c.p, h->rho_c.p, h->scalars.p + S_COST_C)) != PBA_OK) return st;
    if (h->world > 1) {
      // S_COST_C, S_MODEL, S_STEP2, S_XNORM2 are contiguous
      if ((st = allreduce_scalars(h, h->scalars.p + S_COST_C, 4, false)) != PBA_OK) return st;
    }
    if ((st = read_scalars(h)) != PBA_OK) return st;
    ++n_lin; ++n_res;
    it.linear_solver_iterations = h->last_solver == PBA_SOLVER_PCG ? int(hs[S_PCG_ITERS]) : 1;
    const double model_cost_change = hs[S_MODEL];
    // A direct solver fails on a non-positive pivot; the iterative one when it stops short of its
    // tolerance: the reference's SPARSE_SCHUR solve is exact, so a truncated PCG step is treated as a
    // linear-solver failure (invalid step: the radius shrinks, which also improves the conditioning).
    const bool pcg_short = h->last_solver == PBA_SOLVER_PCG && !(hs[S_PCG_RES] <= opt.pcg_tolerance);
    if (pcg_short) ++n_inexact;
    if (*chol_fail == 3) return PBA_ERR_NCCL;  // a peer-memory collective timed out (peer.cu): a rank is missing
    const bool solved = !(h->last_solver != PBA_SOLVER_PCG && *chol_fail) && !pcg_short && std::isfinite(model_cost_change) &&
                        std::isfinite(hs[S_STEP2]);
    it.model_cost_change = model_cost_change;
    it.step_is_valid = solved && model_cost_change > 0.0;
    if (!it.step_is_valid) {
      // HandleInvalidStep (trust_region_minimizer.cc:450-485)
      if (++consecutive_invalid >= opt.max_num_consecutive_invalid_steps) {
        snprintf(message, sizeof(message), "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt.max_num_consecutive_invalid_steps);
        termination = PBA_FAILURE; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      if ((st = launch_build_rcs(h, radius, false)) != PBA_OK) return st;
      it.cost = x_cost; it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;
      continue;
    }
    consecutive_invalid = 0;
    double cand_cost = hs[S_COST_C];
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached (:706-726) — runs before the accept test
    it.step_norm = sqrt(hs[S_STEP2]);
    if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      snprintf(message, sizeof(message), "Parameter tolerance reached. Relative step_norm: %e <= %e.",
               it.step_norm / (x_norm + opt.parameter_tolerance), opt.parameter_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    // FunctionToleranceReached (:729-748)
    it.cost_change = x_cost - cand_cost;
    if (fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
      snprintf(message, sizeof(message), "Function tolerance reached. |cost_change|/cost: %e <= %e",
               fabs(it.cost_change) / x_cost, opt.function_tolerance);
      termination = PBA_CONVERGENCE; break;
    }
    // IsStepSuccessful (:781-807), monotonic TrustRegionStepEvaluator
    it.relative_decrease = cand_cost >= std::numeric_limits<double>::max()
                               ? std::numeric_limits<double>::lowest()
                               : (current_cost_se - cand_cost) / model_cost_change;
    if (it.relative_decrease > opt.min_relative_decrease) {
      // HandleSuccessfulStep (:812-826): x <- candidate, new Jacobian, StepAccepted
      swap_buf(h->poses, h->poses_c); swap_buf(h->affine, h->affine_c); swap_buf(h->rho, h->rho_c);
      x_norm = sqrt(hs[S_XNORM2]);
      radius = radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      if ((st = eval_jacobian_and_build(h, radius)) != PBA_OK) return st;
      if ((st = read_scalars(h)) != PBA_OK) return st;
      ++n_jac; ++n_res;
      x_cost = hs[S_COST];
      if (!std::isfinite(x_cost)) {
        termination = PBA_FAILURE;
        snprintf(message, sizeof(message), "Residual and Jacobian evaluation failed.");
        break;
      }
      it.cost = x_cost; it.gradient_max_norm = hs[S_GMAX]; it.gradient_norm = sqrt(hs[S_GNORM2]);
      it.step_is_successful = 1;
      current_cost_se = cand_cost;
    } else {
      it.step_is_successful = 0;
      it.cost = cand_cost;
      it.gradient_max_norm = prev_gmax; it.gradient_norm = prev_gnorm;
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      if ((st = launch_build_rcs(h, radius, false)) != PBA_OK) return st;
    }
  }
  // Leave the best accepted iterate as the current state (solver.cc:438-447);
  // on FAILURE the caller keeps its inputs (pba_solve does not write back).
  if (have_best && termination != PBA_FAILURE) {
    if ((st = copy_state(h, h->poses, h->affine, h->rho, h->poses_best, h->affine_best, h->rho_best)) != PBA_OK) return st;
  }
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  ks.resolve();
  h->have_jac = false;  // the stored Jacobian may belong to a later iterate than the state
  h->have_rcs = false;
  if (sum) {
    sum->termination_type = termination;
    sum->num_iterations = cap > 0 ? std::min(n_it, cap) : n_it;
    sum->num_successful_steps = num_successful;
    sum->num_unsuccessful_steps = num_unsuccessful;
    sum->num_residual_evaluations = n_res;
    sum->num_jacobian_evaluations = n_jac;
    sum->num_linear_solves = n_lin;
    sum->rcs_dim = z.dim;
    sum->rcs_blocks = z.n_blocks;
    sum->num_residual_blocks = h->n_obs_global;
    sum->num_residuals = h->n_obs_global * z.R;
    sum->num_effective_parameters = int64_t(z.n_slots) * 6 + h->n_active_lm;
    for (int i = 0; i < z.n_poses; ++i) sum->num_effective_parameters += h->affine_active[i] ? 2 : 0;
    sum->linear_solver = h->last_solver;
    sum->num_inexact_linear_solves = n_inexact;
    sum->gpu_kernel_launches = std::accumulate(ks.launches, ks.launches + K_NUM, int64_t(0)) - launches0;
    sum->initial_cost = initial_cost;
    sum->final_cost = min_iteration_cost;
    sum->jacobian_evaluation_time_in_seconds = 1e-3 * (ks.ms[K_RESJAC] + ks.ms[K_EDGE_PREP]);
    sum->residual_evaluation_time_in_seconds = 1e-3 * ks.ms[K_COST];
    double lin = 0;
    for (int k : {K_EDGE_GRAM, K_LM_GATHER, K_LM_SCALE, K_SCHUR_SYRK, K_RCS_REDUCE, K_RCS_SCALE, K_CAM_SCALE, K_DENSE_FILL,
                  K_CHOL_PANEL, K_CHOL_TRSM, K_CHOL_SYRK, K_CHOL_SOLVE, K_PCG, K_BAND_CHOL, K_BCR, K_BACKSUB})
      lin += ks.ms[k];
    sum->linear_solver_time_in_seconds = 1e-3 * lin;
    sum->minimizer_time_in_seconds = wall() - t_start;
    sum->total_time_in_seconds = sum->minimizer_time_in_seconds;
    if (getenv("PBA_TIMING"))
      fprintf(stderr, "[pba_minimize] rank %d/%d: %d iterations, wall %.1f ms, waiting for the device %.1f ms, host enqueue %.1f ms\n",
              h->rank, h->world, n_it, 1e3 * sum->minimizer_time_in_seconds, 1e3 * h->t_wait,
              1e3 * (sum->minimizer_time_in_seconds - h->t_wait));
    if (n_inexact > 0)
      snprintf(sum->message, sizeof(sum->message), "%.180s [%d PCG solve(s) stopped above pcg_tolerance]", message, n_inexact);
    else
      snprintf(sum->message, sizeof(sum->message), "%s", message);
  }
  return PBA_OK;
}

void print_report(const pba_summary& s, int verbosity) {
  if (verbosity <= 0) return;
  const char* term = s.termination_type == PBA_CONVERGENCE ? "CONVERGENCE" : s.termination_type == PBA_NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE";
  // same one-line shape as ceres::Solver::Summary::BriefReport (map_utils.h:384-388)
  printf("B200 PBA Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s\n", s.num_iterations,
         s.initial_cost, s.final_cost, term);
  if (verbosity >= 2) {
    printf("  residual blocks %lld, residuals %lld, effective parameters %lld, RCS dim %d (%lld blocks)\n",
           (long long)s.num_residual_blocks, (long long)s.num_residuals, (long long)s.num_effective_parameters, s.rcs_dim,
           (long long)s.rcs_blocks);
    printf("  steps: %d successful, %d unsuccessful; evaluations: %d residual, %d jacobian; linear solves %d\n",
           s.num_successful_steps, s.num_unsuccessful_steps, s.num_residual_evaluations, s.num_jacobian_evaluations,
           s.num_linear_solves);
    printf("  time (s): setup %.4f, minimizer %.4f, total %.4f; kernel launches %lld\n  %s\n", s.setup_time_in_seconds,
           s.minimizer_time_in_seconds, s.total_time_in_seconds, (long long)s.gpu_kernel_launches, s.message);
    for (int i = 0; i < s.num_iterations && s.iterations && i < s.iterations_capacity; ++i) {
      const pba_iteration& a = s.iterations[i];
      printf("  %3d cost %.6e change %.2e |grad| %.2e |step| %.2e tr_ratio %.2e radius %.2e ls_iter %d\n", a.iteration,
             a.cost, a.cost_change, a.gradient_max_norm, a.step_norm, a.relative_decrease, a.trust_region_radius,
             a.linear_solver_iterations);
    }
  }
}

pba_status get_state_impl(Handle* h, double* poses, double* inv_depth, double* affine) {
  const Sizes& z = h->sz;
  if (poses) PBA_CUDA_OK(cudaMemcpyAsync(poses, h->poses.p, sizeof(double) * 7 * z.n_poses, cudaMemcpyDeviceToHost, h->stream));
  if (affine && z.mode == PBA_MODE_PHOTOMETRIC)
    PBA_CUDA_OK(cudaMemcpyAsync(affine, h->affine.p, sizeof(double) * 2 * z.n_poses, cudaMemcpyDeviceToHost, h->stream));
  std::vector<double> rho(z.n_lm);
  if (inv_depth && z.n_lm) PBA_CUDA_OK(cudaMemcpyAsync(rho.data(), h->rho.p, sizeof(double) * z.n_lm, cudaMemcpyDeviceToHost, h->stream));
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (inv_depth)
    for (int li = 0; li < z.n_lm; ++li) inv_depth[h->lm_order[li]] = rho[li];
  return PBA_OK;
}

}  // namespace
}  // namespace pba

using namespace pba;

// ============================================================== C ABI =====
PBA_API int32_t pba_abi_version(void) { return PBA_ABI_VERSION; }

PBA_API const char* pba_status_string(pba_status s) {
  switch (s) {
    case PBA_OK: return "PBA_OK";
    case PBA_ERR_INVALID_ARGUMENT: return "PBA_ERR_INVALID_ARGUMENT";
    case PBA_ERR_NO_DEVICE: return "PBA_ERR_NO_DEVICE";
    case PBA_ERR_CUDA: return "PBA_ERR_CUDA";
    case PBA_ERR_UNSUPPORTED: return "PBA_ERR_UNSUPPORTED";
    case PBA_ERR_NUMERICAL_FAILURE: return "PBA_ERR_NUMERICAL_FAILURE";
    case PBA_ERR_NCCL: return "PBA_ERR_NCCL";
    case PBA_ERR_OUT_OF_MEMORY: return "PBA_ERR_OUT_OF_MEMORY";
  }
  return "PBA_ERR_UNKNOWN";
}

PBA_API int32_t pba_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

PBA_API void pba_options_init(pba_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->verbosity_level = 1; o->optimize_intrinsics = 0; o->use_huber = 1; o->huber_parameter = 1.0; o->max_num_iterations = 20;
  o->solver = PBA_SOLVER_AUTO; o->cholesky_max_dim = 16384; o->pcg_max_iterations = 500; o->pcg_tolerance = 1e-10;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1; o->device = 0; o->profile = 0; o->num_gpus = 1;
}

PBA_API pba_status pba_create(const pba_problem* problem, const pba_options* options, int32_t rank, int32_t world_size,
                              pba_handle** out) {
  if (!out) return PBA_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  Handle* h = nullptr;
  pba_status st = create_impl(problem, options, rank, world_size, &h);
  if (st == PBA_OK) *out = reinterpret_cast<pba_handle*>(h);
  return st;
}

PBA_API void pba_destroy(pba_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  delete h;
}

PBA_API void pba_trim_device_cache(void) {
  trim_device_cache();
  g_pinned.trim();
}

// Host-only: the camera layout pba_create would use (no device needed; CPU tests of the ordering logic).
PBA_API pba_status pba_analyze_structure(const pba_problem* problem, const pba_options* options, int32_t* slot,
                                         int32_t* n_slots, int32_t* bandwidth_natural, int32_t* bandwidth,
                                         int64_t* n_blocks) {
  pba_status st = validate(problem, options);
  if (st != PBA_OK) return st;
  const int cd = problem->mode == PBA_MODE_PHOTOMETRIC ? 8 : 6;
  CameraLayout L;
  if ((st = analyze_cameras(problem, std::max(1, omp_get_num_procs()), 128 / cd, &L)) != PBA_OK) return st;
  if (slot) for (int i = 0; i < problem->n_poses; ++i) slot[i] = L.slot[i];
  if (n_slots) *n_slots = L.n_slots;
  if (bandwidth_natural) *bandwidth_natural = L.bandwidth_natural;
  if (bandwidth) *bandwidth = L.bandwidth;
  if (n_blocks) {
    int64_t nb = 0;
    for (const auto& v : L.adj) nb += int64_t(v.size());
    *n_blocks = nb;
  }
  return PBA_OK;
}

PBA_API pba_status pba_set_stream(pba_handle* hh, void* cuda_stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  if (cuda_stream) {
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    PBA_CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return PBA_OK;
}

PBA_API pba_status pba_synchronize(pba_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->stats.resolve();
  return PBA_OK;
}

PBA_API pba_status pba_evaluate(pba_handle* hh, int32_t with_jacobian, double* cost) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  pba_status st = launch_evaluate(h, with_jacobian != 0, h->poses.p, h->affine.p, h->rho.p, h->scalars.p + S_COST);
  if (st != PBA_OK) return st;
  if (with_jacobian) { h->have_jac = true; h->have_rcs = false; }
  if (cost) {
    if (h->world > 1 && (st = allreduce_scalars(h, h->scalars.p + S_COST, 1, false)) != PBA_OK) return st;
    if ((st = read_scalars(h)) != PBA_OK) return st;
    *cost = h->h_scalars[S_COST];
  }
  return PBA_OK;
}

PBA_API pba_status pba_get_residuals(pba_handle* hh, double* residuals) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !residuals || !h->have_jac) return PBA_ERR_INVALID_ARGUMENT;
  const Sizes& z = h->sz;
  if (z.n_obs == 0) return PBA_OK;
  DevBuf<double> tmp;
  PBA_CUDA_OK(tmp.alloc(size_t(z.n_obs) * z.R));
  pba_status st = launch_unpermute(h, 0, tmp.p);
  if (st != PBA_OK) return st;
  PBA_CUDA_OK(cudaMemcpy(residuals, tmp.p, sizeof(double) * z.n_obs * z.R, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_get_jacobians(pba_handle* hh, double* jacobians) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !jacobians || !h->have_jac) return PBA_ERR_INVALID_ARGUMENT;
  const Sizes& z = h->sz;
  if (z.n_obs == 0) return PBA_OK;
  DevBuf<double> tmp;
  PBA_CUDA_OK(tmp.alloc(size_t(z.n_obs) * z.R * z.C));
  pba_status st = launch_unpermute(h, 1, tmp.p);
  if (st != PBA_OK) return st;
  PBA_CUDA_OK(cudaMemcpy(jacobians, tmp.p, sizeof(double) * z.n_obs * z.R * z.C, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_get_blocks(pba_handle* hh, int64_t n_sel, const int64_t* obs_index, double* residuals,
                                  double* jacobians) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || n_sel < 0 || (n_sel > 0 && !obs_index) || !h->have_jac) return PBA_ERR_INVALID_ARGUMENT;
  const Sizes& z = h->sz;
  if (n_sel == 0 || (!residuals && !jacobians)) return PBA_OK;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  // caller index -> edge-order position, for the selection only: one parallel scan of the permutation
  std::vector<int64_t> where(size_t(z.n_obs), -1);
  for (int64_t j = 0; j < n_sel; ++j) {
    if (obs_index[j] < 0 || obs_index[j] >= z.n_obs || where[size_t(obs_index[j])] >= 0) return PBA_ERR_INVALID_ARGUMENT;
    where[size_t(obs_index[j])] = j;
  }
  std::vector<int64_t> pos(size_t(n_sel), -1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < z.n_obs; ++i) {
    const int64_t j = where[size_t(h->obs_order[size_t(i)])];
    if (j >= 0) pos[size_t(j)] = i;
  }
  DevBuf<int64_t> d_pos;
  DevBuf<double> d_res, d_jac;
  PBA_CUDA_OK(d_pos.upload(pos, h->stream));
  if (residuals) PBA_CUDA_OK(d_res.alloc(size_t(n_sel) * z.R));
  if (jacobians) PBA_CUDA_OK(d_jac.alloc(size_t(n_sel) * z.R * z.C));
  pba_status st = launch_gather_blocks(h, n_sel, d_pos.p, d_res.p, d_jac.p);
  if (st != PBA_OK) return st;
  if (residuals) PBA_CUDA_OK(cudaMemcpyAsync(residuals, d_res.p, sizeof(double) * n_sel * z.R, cudaMemcpyDeviceToHost, h->stream));
  if (jacobians) PBA_CUDA_OK(cudaMemcpyAsync(jacobians, d_jac.p, sizeof(double) * n_sel * z.R * z.C, cudaMemcpyDeviceToHost, h->stream));
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  return PBA_OK;
}

PBA_API pba_status pba_build_rcs(pba_handle* hh, double radius) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !(radius > 0.0)) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  pba_status st;
  if (!h->have_jac) {
    if ((st = launch_evaluate(h, true, h->poses.p, h->affine.p, h->rho.p, h->scalars.p + S_COST)) != PBA_OK) return st;
    h->have_jac = true;
  }
  // stand-alone use = iteration 0 semantics: scales and LM diagonal from this Jacobian
  h->scale_ready = false;
  if ((st = launch_post_jacobian(h)) != PBA_OK) return st;
  if ((st = launch_build_rcs(h, radius, true)) != PBA_OK) return st;
  return pba_synchronize(hh);
}

PBA_API pba_status pba_get_rcs_dim(pba_handle* hh, int32_t* dim) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !dim) return PBA_ERR_INVALID_ARGUMENT;
  *dim = h->sz.dim;
  return PBA_OK;
}

PBA_API pba_status pba_get_rcs(pba_handle* hh, double* S_dense, double* rhs) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !S_dense || !rhs || !h->have_rcs) return PBA_ERR_INVALID_ARGUMENT;
  const Sizes& z = h->sz;
  const int cd = z.cd;
  std::vector<double> buf(size_t(z.n_blocks) * cd * cd + z.dim);
  PBA_CUDA_OK(cudaMemcpyAsync(buf.data(), h->rcs.p, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, h->stream));
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  std::fill(S_dense, S_dense + size_t(z.dim) * z.dim, 0.0);
  for (int64_t b = 0; b < z.n_blocks; ++b)
    for (int r = 0; r < cd; ++r)
      for (int c = 0; c < cd; ++c) {
        const double v = buf[b * cd * cd + r * cd + c];
        const int gr = h->blk_row[b] * cd + r, gc = h->blk_col[b] * cd + c;
        S_dense[size_t(gr) * z.dim + gc] = v;
        S_dense[size_t(gc) * z.dim + gr] = v;
      }
  memcpy(rhs, buf.data() + size_t(z.n_blocks) * cd * cd, sizeof(double) * z.dim);
  return PBA_OK;
}

PBA_API pba_status pba_solve_rcs(pba_handle* hh, int32_t solver, double* y_cam, int32_t* iterations) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !h->have_rcs) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  pba_status st = solve_rcs(h, solver);
  if (st != PBA_OK) return st;
  if ((st = read_scalars(h)) != PBA_OK) return st;
  if (h->last_solver != PBA_SOLVER_PCG && *reinterpret_cast<int*>(h->h_scalars + S_NUM)) return PBA_ERR_NUMERICAL_FAILURE;
  if (iterations) *iterations = h->last_solver == PBA_SOLVER_PCG ? int(h->h_scalars[S_PCG_ITERS]) : 1;
  if (y_cam && h->sz.dim) PBA_CUDA_OK(cudaMemcpy(y_cam, h->y_cam.p, sizeof(double) * h->sz.dim, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_minimize(pba_handle* hh, pba_summary* summary) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  return minimize_impl(h, summary);
}

PBA_API pba_status pba_lm_iterate(pba_handle* hh, double radius, int32_t apply, pba_iteration* out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !(radius > 0.0)) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  pba_status st;
  h->scale_ready = false;
  if ((st = eval_jacobian_and_build(h, radius)) != PBA_OK) return st;
  if ((st = solve_rcs(h, PBA_SOLVER_AUTO)) != PBA_OK) return st;
  if ((st = launch_backsub(h)) != PBA_OK) return st;
  if ((st = launch_retract(h)) != PBA_OK) return st;
  if ((st = launch_evaluate(h, false, h->poses_c.p, h->affine_c.p, h->rho_c.p, h->scalars.p + S_COST_C)) != PBA_OK) return st;
  if (h->world > 1 && (st = allreduce_scalars(h, h->scalars.p + S_COST_C, 4, false)) != PBA_OK) return st;
  if ((st = read_scalars(h)) != PBA_OK) return st;
  const double* hs = h->h_scalars;
  const bool chol_failed = h->last_solver != PBA_SOLVER_PCG && *reinterpret_cast<const int*>(hs + S_NUM);
  pba_iteration it;
  memset(&it, 0, sizeof(it));
  it.cost = hs[S_COST];
  it.model_cost_change = hs[S_MODEL];
  it.step_is_valid = !chol_failed && std::isfinite(hs[S_MODEL]) && hs[S_MODEL] > 0.0;
  it.cost_change = hs[S_COST] - hs[S_COST_C];
  it.step_norm = sqrt(hs[S_STEP2]);
  it.gradient_max_norm = hs[S_GMAX];
  it.gradient_norm = sqrt(hs[S_GNORM2]);
  it.trust_region_radius = radius;
  it.linear_solver_iterations = h->last_solver == PBA_SOLVER_PCG ? int(hs[S_PCG_ITERS]) : 1;
  it.relative_decrease = it.step_is_valid ? it.cost_change / it.model_cost_change : 0.0;
  it.step_is_successful = it.step_is_valid && it.relative_decrease > h->opt.min_relative_decrease;
  if (apply && it.step_is_successful) {
    swap_buf(h->poses, h->poses_c); swap_buf(h->affine, h->affine_c); swap_buf(h->rho, h->rho_c);
    h->have_jac = false; h->have_rcs = false;
  }
  if (out) *out = it;
  return PBA_OK;
}

PBA_API pba_status pba_set_state(pba_handle* hh, const double* poses, const double* inv_depth, const double* affine) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  const Sizes& z = h->sz;
  if (poses) PBA_CUDA_OK(cudaMemcpyAsync(h->poses.p, poses, sizeof(double) * 7 * z.n_poses, cudaMemcpyHostToDevice, h->stream));
  if (affine && z.mode == PBA_MODE_PHOTOMETRIC)
    PBA_CUDA_OK(cudaMemcpyAsync(h->affine.p, affine, sizeof(double) * 2 * z.n_poses, cudaMemcpyHostToDevice, h->stream));
  std::vector<double> rho(z.n_lm);
  if (inv_depth && z.n_lm) {
    for (int li = 0; li < z.n_lm; ++li) rho[li] = inv_depth[h->lm_order[li]];
    PBA_CUDA_OK(cudaMemcpyAsync(h->rho.p, rho.data(), sizeof(double) * z.n_lm, cudaMemcpyHostToDevice, h->stream));
  }
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->have_jac = false; h->have_rcs = false;
  return PBA_OK;
}

PBA_API pba_status pba_get_state(pba_handle* hh, double* poses, double* inv_depth, double* affine) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  return get_state_impl(h, poses, inv_depth, affine);
}

PBA_API pba_status pba_get_sizes(pba_handle* hh, int64_t* n_obs_local, int32_t* n_landmarks_local, int64_t* first_landmark) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  if (n_obs_local) *n_obs_local = h->sz.n_obs;
  if (n_landmarks_local) *n_landmarks_local = h->sz.n_lm;
  if (first_landmark) *first_landmark = h->first_landmark;
  return PBA_OK;
}

PBA_API pba_status pba_reset_kernel_stats(pba_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->stats.reset();
  return PBA_OK;
}

PBA_API pba_status pba_set_profile(pba_handle* hh, int32_t level) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || level < 0 || level > 2) return PBA_ERR_INVALID_ARGUMENT;
  PBA_CUDA_OK(cudaStreamSynchronize(h->stream));
  h->stats.resolve();
  h->stats.profile = level;
  return PBA_OK;
}

PBA_API int32_t pba_get_kernel_stats(pba_handle* hh, pba_kernel_stat* out, int32_t cap) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !out) return 0;
  cudaStreamSynchronize(h->stream);
  h->stats.resolve();
  int n = 0;
  for (int i = 0; i < K_NUM && n < cap; ++i) {
    if (!h->stats.launches[i]) continue;
    snprintf(out[n].name, sizeof(out[n].name), "%s", kKernelNames[i]);
    out[n].launches = h->stats.launches[i];
    out[n].total_ms = h->stats.ms[i];
    ++n;
  }
  return n;
}

PBA_API pba_status pba_nccl_unique_id(uint8_t id[PBA_NCCL_ID_BYTES]) {
  if (!id) return PBA_ERR_INVALID_ARGUMENT;
  if (!g_nccl.load()) return PBA_ERR_NCCL;
  NcclId nid;
  if (g_nccl.GetUniqueId(&nid) != 0) return PBA_ERR_NCCL;
  memcpy(id, nid.b, PBA_NCCL_ID_BYTES);
  return PBA_OK;
}

PBA_API pba_status pba_comm_init(pba_handle* hh, const uint8_t id[PBA_NCCL_ID_BYTES]) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !id) return PBA_ERR_INVALID_ARGUMENT;
  if (h->world <= 1) return PBA_OK;
  if (!g_nccl.load()) return PBA_ERR_NCCL;
  PBA_CUDA_OK(cudaSetDevice(h->device));
  NcclId nid;
  memcpy(nid.b, id, PBA_NCCL_ID_BYTES);
  // Communicators are cached per process, keyed by (id, world, rank, device): ncclCommInitRank
  // costs seconds at 8 ranks, a solve milliseconds, so every later handle given the same id
  // (the SfM loop calls optimize() again and again) reuses the first one.
  struct Entry { std::string key; void* comm; PeerExchange* px; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  std::string key(reinterpret_cast<const char*>(id), PBA_NCCL_ID_BYTES);
  key += "/" + std::to_string(h->world) + "/" + std::to_string(h->rank) + "/" + std::to_string(h->device);
  std::lock_guard<std::mutex> lock(mu);
  Entry* ent = nullptr;
  for (auto& e : cache)
    if (e.key == key) ent = &e;
  if (!ent) {
    void* comm = nullptr;
    if (g_nccl.CommInitRank(&comm, h->world, nid, h->rank) != 0) return PBA_ERR_NCCL;
    cache.push_back(Entry{key, comm, nullptr});
    ent = &cache.back();
  }
  h->nccl_comm = ent->comm;
  // the exchange buffer of the peer-memory collectives: every rank sees the same payload size (the RCS layout is
  // global), so every rank takes the same branch here
  const size_t payload = h->rcs.n * sizeof(double);
  if (h->world > 1 && (!ent->px || ent->px->bytes < payload)) {
    // a larger problem than before: the old mappings go first; the gather inside peer_create_ipc is the
    // point after which no rank can still be using them
    PeerExchange* old = ent->px;
    ent->px = nullptr;
    if (old) {
      for (int p = 0; p < old->world; ++p)
        if (p != old->rank && old->base[p]) { cudaIpcCloseMemHandle(old->base[p]); old->base[p] = nullptr; }
    }
    ent->px = peer_create_ipc(ent->comm, h->rank, h->world, h->device, payload + payload / 4, h->stream);
    if (old) peer_release(old);
  }
  if (ent->px) {
    h->peer = ent->px;
    h->rcs.p = ent->px->buf[h->rank];  // arena memory is not freed per buffer: the handle's own block simply stays unused
  }
  return PBA_OK;
}

PBA_API int32_t pba_collective_kind(pba_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || h->world <= 1) return 0;
  return h->peer ? 2 : 1;
}

// ---- single-process multi-GPU (pba_options.num_gpus) ----
// The reference's caller is one thread of one process (src/sfm.cpp:1883-1925), so the drop-in call drives
// all GPUs itself: one host thread per device, each with its landmark shard (create_impl(rank, world)), the
// partial reduced camera systems summed by NCCL.  Communicators come from ncclCommInitAll and are cached per
// process and device list (creating them costs seconds; the SfM loop calls optimize() again and again).
namespace pba {
namespace {

std::mutex g_multi_mu;
struct MultiComms { int device0, n; std::vector<void*> comms; std::vector<PeerExchange*> px; size_t px_bytes = 0; };
std::vector<MultiComms> g_multi;

pba_status multi_comms(int device0, int n, std::vector<void*>* out) {
  if (!g_nccl.load()) return PBA_ERR_NCCL;
  std::lock_guard<std::mutex> lock(g_multi_mu);
  for (const MultiComms& m : g_multi)
    if (m.device0 == device0 && m.n == n) { *out = m.comms; return PBA_OK; }
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = device0 + i;
  std::vector<void*> comms(n, nullptr);
  if (g_nccl.CommInitAll(comms.data(), n, devs.data()) != 0) return PBA_ERR_NCCL;
  // NCCL sets its channels / peer connections up lazily, on the first collective: run one here (a grouped
  // all-reduce of a few doubles on every device) so that it is part of the communicator set-up, not of a solve
  {
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<double*> buf(n, nullptr);
    std::vector<cudaStream_t> st(n, nullptr);
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
      ok = cudaSetDevice(devs[i]) == cudaSuccess && cudaMalloc(&buf[i], 64 * sizeof(double)) == cudaSuccess &&
           cudaMemset(buf[i], 0, 64 * sizeof(double)) == cudaSuccess &&
           cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess;
    }
    if (ok) {
      g_nccl.GroupStart();
      for (int i = 0; i < n; ++i) {
        cudaSetDevice(devs[i]);
        ok = ok && g_nccl.AllReduce(buf[i], buf[i], 64, kNcclDouble, kNcclSum, comms[i], st[i]) == 0;
      }
      ok = g_nccl.GroupEnd() == 0 && ok;
    }
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(devs[i]);
      if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
      if (buf[i]) cudaFree(buf[i]);
    }
    cudaSetDevice(cur);
    if (!ok) return PBA_ERR_NCCL;
  }
  g_multi.push_back(MultiComms{device0, n, comms, {}, 0});
  *out = comms;
  return PBA_OK;
}

// the exchange buffers of the cached communicators, (re)created when a problem needs a larger payload; no solve
// may be running on these devices (pba_solve calls are serial in the caller's thread)
void multi_exchange(int device0, int n, size_t payload, std::vector<PeerExchange*>* out) {
  out->assign(n, nullptr);
  std::lock_guard<std::mutex> lock(g_multi_mu);
  for (MultiComms& m : g_multi) {
    if (m.device0 != device0 || m.n != n) continue;
    if (m.px.empty() || m.px_bytes < payload) {
      if (!m.px.empty()) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (int r = 0; r < n; ++r) {
          if (!m.px[r]) continue;
          cudaSetDevice(device0 + r);
          cudaDeviceSynchronize();
          cudaFree(m.px[r]->base[r]);
          delete m.px[r];
        }
        cudaSetDevice(cur);
        m.px.clear();
      }
      m.px_bytes = payload + payload / 4;
      if (!peer_create_local(device0, n, m.px_bytes, &m.px)) { m.px.assign(n, nullptr); }
    }
    *out = m.px;
    return;
  }
}

// all threads arrive, then everyone learns whether anyone failed (no rank may enter a collective alone)
struct StatusBarrier {
  std::mutex mu;
  std::condition_variable cv;
  int n, arrived = 0, generation = 0;
  pba_status worst = PBA_OK;
  explicit StatusBarrier(int n_) : n(n_) {}
  pba_status wait(pba_status mine) {
    std::unique_lock<std::mutex> lock(mu);
    if (mine != PBA_OK && worst == PBA_OK) worst = mine;
    const int gen = generation;
    if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lock, [&] { return generation != gen; });
    return worst;
  }
};

pba_status solve_multi_gpu(pba_problem* problem, const pba_options* options, int n_gpus, pba_summary* summary, bool emulate) {
  const double t0 = wall();
  std::vector<void*> comms(size_t(n_gpus), nullptr);
  EmuExchange emu(n_gpus);
  pba_status st = emulate ? PBA_OK : multi_comms(options->device, n_gpus, &comms);
  if (st != PBA_OK) return st;
  // the global part of the set-up (camera layout, RCS pattern) once, with every core; the rank threads then
  // order their own shards with their share of the cores
  const bool timing = getenv("PBA_TIMING") != nullptr;
  if (timing) fprintf(stderr, "[pba_solve x%d] communicators %.1f ms\n", n_gpus, 1e3 * (wall() - t0));
  CameraLayout layout;
  st = analyze_cameras(problem, std::max(1, omp_get_num_procs()), 128 / (problem->mode == PBA_MODE_PHOTOMETRIC ? 8 : 6), &layout);
  if (st != PBA_OK) return st;
  if (timing) fprintf(stderr, "[pba_solve x%d] + camera layout %.1f ms\n", n_gpus, 1e3 * (wall() - t0));
  std::vector<PeerExchange*> px(size_t(n_gpus), nullptr);
  if (!emulate) {
    const int cd = problem->mode == PBA_MODE_PHOTOMETRIC ? 8 : 6;
    int64_t n_blocks = 0;
    for (const auto& a : layout.adj) n_blocks += int64_t(a.size());
    const size_t count = size_t(n_blocks) * cd * cd + 3 * size_t(layout.n_slots) * cd + 2 + size_t(n_gpus);
    multi_exchange(options->device, n_gpus, count * sizeof(double), &px);
  }
  StatusBarrier barrier(n_gpus);
  std::vector<pba_status> status(n_gpus, PBA_OK);
  std::vector<double> t_setup(n_gpus, 0.0);
  pba_iteration* its = summary ? summary->iterations : nullptr;
  const int cap = summary ? summary->iterations_capacity : 0;
  auto worker = [&](int r) {
    pba_options o = *options;
    o.device = emulate ? options->device : options->device + r;
    o.num_gpus = 1;
    Handle* h = nullptr;
    pba_status s = create_impl(problem, &o, r, n_gpus, &h, &layout);
    std::unique_ptr<Handle> guard(h);
    t_setup[r] = wall() - t0;
    if (timing) fprintf(stderr, "[pba_solve x%d] rank %d created at %.1f ms\n", n_gpus, r, 1e3 * t_setup[r]);
    if (barrier.wait(s) != PBA_OK) { status[r] = s; return; }
    h->nccl_comm = comms[r];
    if (emulate) h->emu = &emu;
    if (px[r] && px[r]->bytes >= h->rcs.n * sizeof(double)) {
      h->peer = px[r];
      h->rcs.p = px[r]->buf[r];
    }
    pba_summary local;
    memset(&local, 0, sizeof(local));
    if (r == 0) { local.iterations = its; local.iterations_capacity = cap; }
    s = minimize_impl(h, &local);
    // every rank takes the same decisions (all of them see the all-reduced scalars), so termination agrees
    if (s == PBA_OK && local.termination_type != PBA_FAILURE)
      s = get_state_impl(h, r == 0 ? problem->poses : nullptr, problem->inv_depth + h->first_landmark,
                         (r == 0 && problem->mode == PBA_MODE_PHOTOMETRIC) ? problem->affine : nullptr);
    if (r == 0 && summary) *summary = local;
    status[r] = s;
    if (timing) fprintf(stderr, "[pba_solve x%d] rank %d state read at %.1f ms\n", n_gpus, r, 1e3 * (wall() - t0));
    guard.reset();
    if (timing) fprintf(stderr, "[pba_solve x%d] rank %d destroyed at %.1f ms\n", n_gpus, r, 1e3 * (wall() - t0));
  };
  std::vector<std::thread> threads;
  for (int r = 1; r < n_gpus; ++r) threads.emplace_back(worker, r);
  worker(0);
  for (std::thread& t : threads) t.join();
  for (int r = 0; r < n_gpus; ++r)
    if (status[r] != PBA_OK) return status[r];
  if (summary) {
    summary->setup_time_in_seconds = *std::max_element(t_setup.begin(), t_setup.end());
    summary->total_time_in_seconds = wall() - t0;
    print_report(*summary, options->verbosity_level);
  }
  cudaSetDevice(options->device);
  return PBA_OK;
}

}  // namespace
}  // namespace pba

// Creates (and caches) the NCCL communicators pba_solve(num_gpus > 1) uses, so that the first solve does not
// pay for them.  Optional: pba_solve does it on demand.
PBA_API pba_status pba_multi_gpu_init(int32_t device, int32_t num_gpus) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return PBA_ERR_NO_DEVICE; }
  if (num_gpus == 0) num_gpus = ndev - device;
  if (device < 0 || num_gpus < 1 || device + num_gpus > ndev) return PBA_ERR_INVALID_ARGUMENT;
  if (num_gpus == 1) return PBA_OK;
  std::vector<void*> comms;
  return multi_comms(device, num_gpus, &comms);
}

// The drop-in call (map_utils.h:322): host buffers in, updated in place.
PBA_API pba_status pba_solve(pba_problem* problem, const pba_options* options, pba_summary* summary) {
  const double t0 = wall();
  if (!problem || !options) return PBA_ERR_INVALID_ARGUMENT;
  int n_gpus = options->num_gpus;
  bool emulate = false;
  if (n_gpus != 1) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return PBA_ERR_NO_DEVICE; }
    if (n_gpus == 0) n_gpus = ndev - options->device;
    // PBA_EMULATE_RANKS=1: the sharded solve with all ranks on options->device (see EmuExchange) — the multi-rank
    // path on a box with fewer GPUs than ranks
    emulate = n_gpus > 1 && getenv("PBA_EMULATE_RANKS") != nullptr && n_gpus <= kMaxPeers && options->device >= 0 &&
              options->device < ndev;
    if (!emulate && (n_gpus < 1 || options->device < 0 || options->device + n_gpus > ndev)) return PBA_ERR_INVALID_ARGUMENT;
    // a shard needs landmarks to work on
    n_gpus = int(std::max<int64_t>(1, std::min<int64_t>(n_gpus, problem->n_landmarks)));
  }
  if (n_gpus > 1) {
    pba_status vst = validate(problem, options);
    if (vst != PBA_OK) return vst;
    return solve_multi_gpu(problem, options, n_gpus, summary, emulate);
  }
  Handle* h = nullptr;
  pba_status st = create_impl(problem, options, 0, 1, &h);
  if (st != PBA_OK) return st;
  std::unique_ptr<Handle> guard(h);
  const double t1 = wall();
  pba_summary local;
  memset(&local, 0, sizeof(local));
  pba_summary* s = summary ? summary : &local;
  st = minimize_impl(h, s);
  if (st != PBA_OK) return st;
  if (s->termination_type != PBA_FAILURE) {
    st = get_state_impl(h, problem->poses, problem->inv_depth, problem->mode == PBA_MODE_PHOTOMETRIC ? problem->affine : nullptr);
    if (st != PBA_OK) return st;
  }
  s->setup_time_in_seconds = t1 - t0;
  s->total_time_in_seconds = wall() - t0;
  print_report(*s, options->verbosity_level);
  return PBA_OK;
}
