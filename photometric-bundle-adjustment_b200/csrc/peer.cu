// peer.cu — the data-path collectives of the sharded solve as ONE kernel over NVLink peer memory.
//
// A sharded LM iteration has one real exchange: every rank holds a partial reduced camera
// system (its landmarks' Schur contributions, ~12 MB at 2,000 keyframes) and all of them need the sum, plus a
// handful of scalars after the candidate evaluation.  ncclAllReduce does that in 0.12 ms at 8 GPUs
// (profiles/r02b_scale_n8.json: `copy`), most of it protocol latency: the payload is 1.5 MB per rank and
// NVSwitch moves that in a few microseconds.  Here every rank's buffer lives in an exchange allocation its
// peers have mapped (cudaIpc* between processes, cudaDeviceEnablePeerAccess inside one), and one cooperative
// kernel per rank does
//     flag barrier (all partials are in place)
//  -> slice r of the sum: read slice r of EVERY rank's buffer through NVLink, add in rank order (the same
//     bits on every rank, run after run), write the result into every rank's buffer
//  -> flag barrier (all slices have landed).
// The barriers are monotonically increasing epoch counters written into the peers' flag words with system-scope
// release stores and polled locally; a rank that waits longer than ~2 s gives up and raises the handle's failure
// flag instead of hanging the device.
#include <cooperative_groups.h>

#include "launch.h"
#include "pba_internal.h"

namespace cg = cooperative_groups;

namespace pba {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// thread p < world of the calling CTA waits for rank p's flag; then the CTA meets.  Returns false on a time-out.
__device__ __forceinline__ bool wait_flags(const unsigned long long* mine, int world, unsigned long long epoch) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (int(threadIdx.x) < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {
      if (clock64() - t0 > 4000000000LL) { s_ok = 0; break; }  // ~2 s
      __nanosleep(100);
    }
  }
  __syncthreads();
  return s_ok != 0;
}

__global__ void __launch_bounds__(512) k_peer_allreduce(PeerArgs a) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long* my_flags = a.flags[a.rank];
  // ---- barrier A: every rank's partial is complete (the kernels before this one, on every rank) ----
  if (blockIdx.x == 0 && int(threadIdx.x) < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
  }
  bool ok = wait_flags(my_flags, a.world, a.epoch);
  // ---- slice `rank` of the sum ----
  const size_t n2 = (a.count + 1) / 2;                       // double2 elements
  const size_t per = (n2 + a.world - 1) / a.world;
  const size_t lo = per * a.rank, hi = lo + per < n2 ? lo + per : n2;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = lo + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < hi; i += stride) {
    double2 v[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < a.world) v[p] = __ldcv(reinterpret_cast<const double2*>(a.buf[p]) + i);
    double2 s = v[0];
#pragma unroll
    for (int p = 1; p < kMaxPeers; ++p)
      if (p < a.world) { s.x += v[p].x; s.y += v[p].y; }
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < a.world) reinterpret_cast<double2*>(a.buf[p])[i] = s;
  }
  __threadfence_system();
  grid.sync();
  // ---- barrier B: every rank has pushed its slice ----
  if (blockIdx.x == 0 && int(threadIdx.x) < a.world) st_release_sys(a.flags[threadIdx.x] + kMaxPeers + a.rank, a.epoch);
  ok = wait_flags(my_flags + kMaxPeers, a.world, a.epoch) && ok;
  if (!ok && threadIdx.x == 0) *a.fail = 3;
}

// a few scalars: every rank writes its values into its slot of every peer's table, one barrier, local sum in rank order
__global__ void __launch_bounds__(64) k_peer_allreduce_small(PeerArgs a, double* __restrict__ dev, int n) {
  if (int(threadIdx.x) < n) {
    const double v = dev[threadIdx.x];
    for (int p = 0; p < a.world; ++p) a.small[p][(a.epoch & 1) * kMaxPeers * kPeerSmall + a.rank * kPeerSmall + threadIdx.x] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (int(threadIdx.x) < a.world) st_release_sys(a.flags[threadIdx.x] + 2 * kMaxPeers + a.rank, a.epoch);
  const bool ok = wait_flags(a.flags[a.rank] + 2 * kMaxPeers, a.world, a.epoch);
  if (int(threadIdx.x) < n) {
    // the table of the OTHER parity may already be receiving the next call's values from a rank that is ahead
    const double* tab = a.small[a.rank] + (a.epoch & 1) * kMaxPeers * kPeerSmall;
    double s = 0.0;
    for (int p = 0; p < a.world; ++p) s += __ldcv(tab + p * kPeerSmall + threadIdx.x);
    dev[threadIdx.x] = s;
  }
  if (!ok && threadIdx.x == 0) *a.fail = 3;
}

}  // namespace

// The sum of `count` doubles at the start of every rank's exchange buffer, in place.
pba_status launch_peer_allreduce(Handle* h, size_t count) {
  PeerExchange* px = h->peer;
  PeerArgs a;
  for (int p = 0; p < kMaxPeers; ++p) {
    a.buf[p] = p < px->world ? px->buf[p] : nullptr;
    a.flags[p] = p < px->world ? px->flags[p] : nullptr;
    a.small[p] = p < px->world ? px->small[p] : nullptr;
  }
  a.rank = px->rank; a.world = px->world; a.count = count; a.epoch = ++px->epoch; a.fail = h->chol_fail.p;
  // 1.5 MB per rank at 2,000 keyframes: a quarter of the SMs saturate the NVLink ports, and a small grid meets faster
  int n_sm = 0;
  PBA_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
  const size_t per = ((count + 1) / 2 + px->world - 1) / px->world;
  int grid = int(std::min<size_t>(size_t(std::max(1, n_sm / 2)), (per + 511) / 512));
  grid = std::max(grid, 1);
  void* args[] = {(void*)&a};
  h->stats.begin(K_COPY, h->stream);
  const cudaError_t e = cudaLaunchCooperativeKernel((void*)k_peer_allreduce, dim3(grid), dim3(512), args, 0, h->stream);
  h->stats.end(h->stream);
  return e == cudaSuccess ? PBA_OK : PBA_ERR_NCCL;
}

// dev[0..n) <- sum over the ranks, n <= kPeerSmall
pba_status launch_peer_allreduce_small(Handle* h, double* dev, int n) {
  PeerExchange* px = h->peer;
  if (n > kPeerSmall) return PBA_ERR_INVALID_ARGUMENT;
  PeerArgs a;
  for (int p = 0; p < kMaxPeers; ++p) {
    a.buf[p] = p < px->world ? px->buf[p] : nullptr;
    a.flags[p] = p < px->world ? px->flags[p] : nullptr;
    a.small[p] = p < px->world ? px->small[p] : nullptr;
  }
  a.rank = px->rank; a.world = px->world; a.count = size_t(n); a.epoch = ++px->epoch_small; a.fail = h->chol_fail.p;
  k_peer_allreduce_small<<<1, 64, 0, h->stream>>>(a, dev, n);
  return cudaGetLastError() == cudaSuccess ? PBA_OK : PBA_ERR_NCCL;
}

}  // namespace pba
