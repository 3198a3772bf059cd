// corners.cu — the feature front-end of the SfM pipeline that feeds bundle_adjustment(),
// on the device: patch orientation + rotated-BRIEF descriptors per corner (include/visnav/keypoints.h:182-245),
// brute-force mutual descriptor matching for a list of image pairs (keypoints.h:248-300; callers match_stereo /
// match_all, src/sfm.cpp:1216-1262, 1286-1330) and the epipolar inlier test of the stereo pairs
// (include/visnav/matching_utils.h:50-79).  Corner detection itself (cv::goodFeaturesToTrack,
// keypoints.h:133-151) stays with the caller.
//
// This is byte / integer work: results are bit-exact against the reference's own functions
// (oracle/ref/frontend_harness.cpp) except the orientation angle (an fp64 atan2).  No tensor cores: a descriptor
// distance is 8 XOR + POPC on 32-bit words; the matcher is bound by the POPC pipe (16 lanes / clock / SM), the
// target set sits in shared memory and is read with broadcast 16-byte loads.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "launch.h"
#include "pba_internal.h"
#include "pba_math.h"

#define PBA_API extern "C" __attribute__((visibility("default")))

namespace pba {
namespace {

bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

// The 256 point pairs of the descriptor: {x_a, y_a, x_b, y_b} per bit, the values of the reference's four
// pattern_31_* tables (keypoints.h:54-131; OpenCV's ORB learned pattern), interleaved so a test reads one word.
__constant__ signed char c_brief[256][4] = {
    {8, -3, 9, 5}, {4, 2, 7, -12}, {-11, 9, -8, 2}, {7, -12, 12, -13}, {2, -13, 2, 12}, {1, -7, 1, 6},
    {-2, -10, -2, -4}, {-13, -13, -11, -8}, {-13, -3, -12, -9}, {10, 4, 11, 9}, {-13, -8, -8, -9}, {-11, 7, -9, 12},
    {7, 7, 12, 6}, {-4, -5, -3, 0}, {-13, 2, -12, -3}, {-9, 0, -7, 5}, {12, -6, 12, -1}, {-3, 6, -2, 12},
    {-6, -13, -4, -8}, {11, -13, 12, -8}, {4, 7, 5, 1}, {5, -3, 10, -3}, {3, -7, 6, 12}, {-8, -7, -6, -2},
    {-2, 11, -1, -10}, {-13, 12, -8, 10}, {-7, 3, -5, -3}, {-4, 2, -3, 7}, {-10, -12, -6, 11}, {5, -12, 6, -7},
    {5, -6, 7, -1}, {1, 0, 4, -5}, {9, 11, 11, -13}, {4, 7, 4, 12}, {2, -1, 4, 4}, {-4, -12, -2, 7},
    {-8, -5, -7, -10}, {4, 11, 9, 12}, {0, -8, 1, -13}, {-13, -2, -8, 2}, {-3, -2, -2, 3}, {-6, 9, -4, -9},
    {8, 12, 10, 7}, {0, 9, 1, 3}, {7, -5, 11, -10}, {-13, -6, -11, 0}, {10, 7, 12, 1}, {-6, -3, -6, 12},
    {10, -9, 12, -4}, {-13, 8, -8, -12}, {-13, 0, -8, -4}, {3, 3, 7, 8}, {5, 7, 10, -7}, {-1, 7, 1, -12},
    {3, -10, 5, 6}, {2, -4, 3, -10}, {-13, 0, -13, 5}, {-13, -7, -12, 12}, {-13, 3, -11, 8}, {-7, 12, -4, 7},
    {6, -10, 12, 8}, {-9, -1, -7, -6}, {-2, -5, 0, 12}, {-12, 5, -7, 5}, {3, -10, 8, -13}, {-7, -7, -4, 5},
    {-3, -2, -1, -7}, {2, 9, 5, -11}, {-11, -13, -5, -13}, {-1, 6, 0, -1}, {5, -3, 5, 2}, {-4, -13, -4, 12},
    {-9, -6, -9, 6}, {-12, -10, -8, -4}, {10, 2, 12, -3}, {7, 12, 12, 12}, {-7, -13, -6, 5}, {-4, 9, -3, 4},
    {7, -1, 12, 2}, {-7, 6, -5, 1}, {-13, 11, -12, 5}, {-3, 7, -2, -6}, {7, -8, 12, -7}, {-13, -7, -11, -12},
    {1, -3, 12, 12}, {2, -6, 3, 0}, {-4, 3, -2, -13}, {-1, -13, 1, 9}, {7, 1, 8, -6}, {1, -1, 3, 12},
    {9, 1, 12, 6}, {-1, -9, -1, 3}, {-13, -13, -10, 5}, {7, 7, 10, 12}, {12, -5, 12, 9}, {6, 3, 7, 11},
    {5, -13, 6, 10}, {2, -12, 2, 3}, {3, 8, 4, -6}, {2, 6, 12, -13}, {9, -12, 10, 3}, {-8, 4, -7, 9},
    {-11, 12, -4, -6}, {1, 12, 2, -8}, {6, -9, 7, -4}, {2, 3, 3, -2}, {6, 3, 11, 0}, {3, -3, 8, -8},
    {7, 8, 9, 3}, {-11, -5, -6, -4}, {-10, 11, -5, 10}, {-5, -8, -3, 12}, {-10, 5, -9, 0}, {8, -1, 12, -6},
    {4, -6, 6, -11}, {-10, 12, -8, 7}, {4, -2, 6, 7}, {-2, 0, -2, 12}, {-5, -8, -5, 2}, {7, -6, 10, 12},
    {-9, -13, -8, -8}, {-5, -13, -5, -2}, {8, -8, 9, -13}, {-9, -11, -9, 0}, {1, -8, 1, -2}, {7, -4, 9, 1},
    {-2, 1, -1, -4}, {11, -6, 12, -11}, {-12, -9, -6, 4}, {3, 7, 7, 12}, {5, 5, 10, 8}, {0, -4, 2, 8},
    {-9, 12, -5, -13}, {0, 7, 2, 12}, {-1, 2, 1, 7}, {5, 11, 7, -9}, {3, 5, 6, -8}, {-13, -4, -8, 9},
    {-5, 9, -3, -3}, {-4, -7, -3, -12}, {6, 5, 8, 0}, {-7, 6, -6, 12}, {-13, 6, -5, -2}, {1, -10, 3, 10},
    {4, 1, 8, -4}, {-2, -2, 2, -13}, {2, -12, 12, 12}, {-2, -13, 0, -6}, {4, 1, 9, 3}, {-6, -10, -3, -5},
    {-3, -13, -1, 1}, {7, 5, 12, -11}, {4, -2, 5, -7}, {-13, 9, -9, -5}, {7, 1, 8, 6}, {7, -8, 7, 6},
    {-7, -4, -7, 1}, {-8, 11, -7, -8}, {-13, 6, -12, -8}, {2, 4, 3, 9}, {10, -5, 12, 3}, {-6, -5, -6, 7},
    {8, -3, 9, -8}, {2, -12, 2, 8}, {-11, -2, -10, 3}, {-12, -13, -7, -9}, {-11, 0, -10, -5}, {5, -3, 11, 8},
    {-2, -13, -1, 12}, {-1, -8, 0, 9}, {-13, -11, -12, -5}, {-10, -2, -10, 11}, {-3, 9, -2, -13}, {2, -3, 3, 2},
    {-9, -13, -4, 0}, {-4, 6, -3, -10}, {-4, 12, -2, -7}, {-6, -11, -4, 9}, {6, -3, 6, 11}, {-13, 11, -5, 5},
    {11, 11, 12, 6}, {7, -5, 12, -2}, {-1, 12, 0, 7}, {-4, -8, -3, -2}, {-7, 1, -6, 7}, {-13, -12, -8, -13},
    {-7, -2, -6, -8}, {-8, 5, -6, -9}, {-5, -1, -4, 5}, {-13, 7, -8, 10}, {1, 5, 5, -13}, {1, 0, 10, -13},
    {9, 12, 10, -1}, {5, -8, 10, -9}, {-1, 11, 1, -13}, {-9, -3, -6, 2}, {-1, -10, 1, 12}, {-13, 1, -8, -10},
    {8, -11, 10, -6}, {2, -13, 3, -6}, {7, -13, 12, -9}, {-10, -10, -5, -7}, {-10, -8, -8, -13}, {4, -6, 8, 5},
    {3, 12, 8, -13}, {-4, 2, -3, -3}, {5, -13, 10, -12}, {4, -13, 5, -1}, {-9, 9, -4, 3}, {0, 3, 3, -9},
    {-12, 1, -6, 1}, {3, 2, 4, -8}, {-10, -10, -10, 9}, {8, -13, 12, 12}, {-8, -12, -6, -5}, {2, 2, 3, 7},
    {10, 6, 11, -8}, {6, 8, 8, -12}, {-7, 10, -6, 5}, {-3, -9, -3, 9}, {-1, -13, -1, 5}, {-3, -7, -3, 4},
    {-8, -2, -8, 3}, {4, 2, 12, 12}, {2, -5, 3, 11}, {6, -9, 11, -13}, {3, -1, 7, 12}, {11, -1, 12, 4},
    {-3, 0, -3, 6}, {4, -11, 4, 12}, {2, -4, 2, 1}, {-10, -6, -8, 1}, {-13, 7, -11, 1}, {-13, 12, -11, -13},
    {6, 0, 11, -13}, {0, -1, 1, 4}, {-13, 3, -9, -2}, {-9, 8, -6, -3}, {-13, -6, -8, -2}, {5, -9, 8, 10},
    {2, 7, 3, -9}, {-1, -6, -1, -1}, {9, 5, 11, -2}, {11, -3, 12, -8}, {3, 0, 3, 5}, {-1, 4, 0, 10},
    {3, -6, 4, 5}, {-13, 0, -10, 5}, {5, 8, 12, 11}, {8, 9, 9, -6}, {7, -4, 8, -12}, {-10, 4, -10, 9},
    {7, 3, 12, 4}, {9, -7, 10, -2}, {7, 0, 12, -2}, {-1, -6, 0, -11}};

constexpr int kHalfPatch = 15;      // HALF_PATCH_SIZE, keypoints.h:49
constexpr int kEdgeThreshold = 19;  // EDGE_THRESHOLD, keypoints.h:50
constexpr int kDescWarps = 8;

// One warp per corner.  Orientation (computeAngles, keypoints.h:182-212): the reference sums y I and x I over
// {(x, y): |y| <= (int)sqrt(15^2 - x^2)} in doubles; for integers that set is x^2 + y^2 <= 225 and the sums are
// exact integers (< 2^23), so the warp sums them as int32 row by row (31 coalesced bytes per row) and only the
// atan2 is floating point.  Descriptor (computeDescriptors, keypoints.h:214-245): lane L evaluates the bits
// 8 L .. 8 L + 7 and writes byte L.
__global__ void __launch_bounds__(32 * kDescWarps) k_corner_desc(int n, int n_images, const int* __restrict__ corner_ptr,
                                                                  const double* __restrict__ corners,
                                                                  const uint8_t* __restrict__ images, int64_t image_stride,
                                                                  int pitch, int rotate, double* __restrict__ angles,
                                                                  uint8_t* __restrict__ desc) {
  const int c = blockIdx.x * kDescWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= n) return;
  // image of corner c: the last i with corner_ptr[i] <= c
  int lo = 0, hi = n_images - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (corner_ptr[mid] <= c) lo = mid; else hi = mid - 1;
  }
  const uint8_t* img = images + int64_t(lo) * image_stride;
  const int cx = int(corners[2 * c]), cy = int(corners[2 * c + 1]);  // truncation, keypoints.h:188-189
  double angle = 0.0;
  if (rotate) {
    int m01 = 0, m10 = 0;
    const int x = lane - kHalfPatch;  // lanes 0..30
    for (int y = -kHalfPatch; y <= kHalfPatch; ++y) {
      if (lane < 2 * kHalfPatch + 1 && x * x + y * y <= kHalfPatch * kHalfPatch) {
        const int v = img[int64_t(cy + y) * pitch + cx + x];
        m01 += y * v;
        m10 += x * v;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    angle = atan2(double(m01), double(m10));
  }
  if (lane == 0) angles[c] = angle;
  double sn, cs;
  sincos(angle, &sn, &cs);
  unsigned byte = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int d = 8 * lane + b;
    const double xa = c_brief[d][0], ya = c_brief[d][1], xb = c_brief[d][2], yb = c_brief[d][3];
    const int rxa = int(round(cs * xa - sn * ya)), rya = int(round(sn * xa + cs * ya));
    const int rxb = int(round(cs * xb - sn * yb)), ryb = int(round(sn * xb + cs * yb));
    const int pa = img[int64_t(cy + rya) * pitch + cx + rxa], pb = img[int64_t(cy + ryb) * pitch + cx + rxb];
    if (pa < pb) byte |= 1u << b;
  }
  desc[int64_t(c) * 32 + lane] = uint8_t(byte);
}

// ---- matching ----
struct PairInfo {
  int a0, na, b0, nb;    // descriptor rows of the first / second set
  int64_t off12, off21;  // where this pair's best-match arrays start
};

constexpr int kMatchThreads = 128;
constexpr int kMatchChunk = 512;  // target descriptors staged per pass: 16 KB

// matchSets (keypoints.h:248-280) for one direction of one pair: thread = one query descriptor (eight 32-bit words
// in registers), the target set streams through shared memory; every thread reads the same target descriptor at
// the same time (two broadcast 16-byte loads).  Targets are visited in ascending order and the comparisons are
// the reference's strict ones, so ties resolve to the lowest index exactly as there.
__global__ void __launch_bounds__(kMatchThreads) k_match_best(const PairInfo* __restrict__ pairs,
                                                               const uint4* __restrict__ desc, int threshold,
                                                               double dist_2_best, int* __restrict__ best12,
                                                               int* __restrict__ best21) {
  __shared__ uint4 s_t[2 * kMatchChunk];
  const PairInfo p = pairs[blockIdx.z];
  const bool fwd = blockIdx.y == 0;
  const int q0 = fwd ? p.a0 : p.b0, nq = fwd ? p.na : p.nb;
  const int t0 = fwd ? p.b0 : p.a0, nt = fwd ? p.nb : p.na;
  if (int(blockIdx.x) * kMatchThreads >= nq) return;  // block-uniform
  const int q = blockIdx.x * kMatchThreads + threadIdx.x;
  const bool live = q < nq;
  uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
  if (live) { a0 = desc[2 * int64_t(q0 + q)]; a1 = desc[2 * int64_t(q0 + q) + 1]; }
  int smallest = 256, second = 256, best = 0;
  for (int c0 = 0; c0 < nt; c0 += kMatchChunk) {
    const int cn = min(kMatchChunk, nt - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * cn; i += kMatchThreads) s_t[i] = desc[2 * int64_t(t0 + c0) + i];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cn; ++j) {
      const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
      const int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                       __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
      if (dist < smallest) {
        second = smallest; smallest = dist; best = c0 + j;
      } else if (dist < second) {
        second = dist;
      }
    }
  }
  if (!live) return;
  if (smallest >= threshold || nt == 0) best = -1;
  if (double(second) < double(smallest) * dist_2_best) best = -1;
  (fwd ? best12 + p.off12 : best21 + p.off21)[q] = best;
}

// matchDescriptors' consistency test (keypoints.h:291-297), one CTA per pair.  WRITE = false: count; true: the
// matches in ascending first index at match_ptr[pair] (ordered compaction: ballot ranks inside a warp, warp
// totals through shared memory).
template <bool WRITE>
__global__ void __launch_bounds__(kMatchThreads) k_match_mutual(const PairInfo* __restrict__ pairs,
                                                                 const int* __restrict__ best12,
                                                                 const int* __restrict__ best21,
                                                                 int64_t* __restrict__ match_ptr, int* __restrict__ matches,
                                                                 int64_t capacity) {
  __shared__ int s_w[kMatchThreads / 32];
  const PairInfo p = pairs[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t out0 = WRITE ? match_ptr[blockIdx.x] : 0;
  int base = 0;
  for (int i0 = 0; i0 < p.na; i0 += kMatchThreads) {
    const int i = i0 + threadIdx.x;
    int j = -1;
    if (i < p.na) {
      j = best12[p.off12 + i];
      if (j >= 0 && best21[p.off21 + j] != i) j = -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, j >= 0);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kMatchThreads / 32; ++w) {
      if (w < warp) before += s_w[w];
      total += s_w[w];
    }
    if (WRITE && j >= 0) {
      const int64_t o = out0 + base + before + __popc(m & ((1u << lane) - 1u));
      if (o < capacity) { matches[2 * o] = i; matches[2 * o + 1] = j; }
    }
    base += total;
    __syncthreads();
  }
  if (!WRITE && threadIdx.x == 0) match_ptr[blockIdx.x + 1] = base;  // counts; scanned in place by k_scan_counts
}

// match_ptr[0] = 0, match_ptr[k + 1] = counts[0] + ... + counts[k]: a few thousand pairs, one thread
__global__ void k_scan_counts(int n, int64_t* __restrict__ ptr) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t run = 0;
  ptr[0] = 0;
  for (int k = 1; k <= n; ++k) { run += ptr[k]; ptr[k] = run; }
}

// ---- epipolar test ----
struct EpiArgs {
  int64_t n;
  int model0, model1;
  double intr0[8], intr1[8], E[9], threshold;
};

__global__ void k_epipolar(const EpiArgs a, const int* __restrict__ matches, const double* __restrict__ c0,
                           const double* __restrict__ c1, uint8_t* __restrict__ inlier) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= a.n) return;
  const int i = matches[2 * k], j = matches[2 * k + 1];
  double xl[3], xr[3];
  cam_unproject(a.model0, a.intr0, c0[2 * i], c0[2 * i + 1], xl);
  cam_unproject(a.model1, a.intr1, c1[2 * j], c1[2 * j + 1], xr);
  // x_L^T (E x_R), the association Eigen uses for the reference's expression (matching_utils.h:74)
  const double e0 = a.E[0] * xr[0] + a.E[1] * xr[1] + a.E[2] * xr[2];
  const double e1 = a.E[3] * xr[0] + a.E[4] * xr[1] + a.E[5] * xr[2];
  const double e2 = a.E[6] * xr[0] + a.E[7] * xr[1] + a.E[8] * xr[2];
  inlier[k] = fabs(xl[0] * e0 + xl[1] * e1 + xl[2] * e2) <= a.threshold ? 1 : 0;
}


// ---- feature tracks (tracks.h:53-160): connected components of the match graph ----
// Label propagation with the smaller-root-wins hook and pointer jumping: every edge (u, v) hooks the larger of the two
// current roots under the smaller one (atomicMin on the root's parent: integer, so the fixed point — every node
// labelled with the smallest node id of its component — does not depend on the order), a compress pass points every
// node at its root, and the pair repeats until no edge joins two roots any more (a handful of rounds).
constexpr int kTrackMaxFeat = 8192;

__device__ __forceinline__ int track_root(const int* parent, int x) {
  int p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  return x;
}

struct TrackEdges {
  int n_pairs;
  int64_t n_edges;
  const int* feat_ptr;
  const int* pairs;
  const int64_t* match_ptr;
  const int* matches;
};

__device__ __forceinline__ void track_edge(const TrackEdges& g, int64_t e, int& u, int& v) {
  int lo = 0, hi = g.n_pairs - 1;  // the pair of edge e: last k with match_ptr[k] <= e
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (g.match_ptr[mid] <= e) lo = mid; else hi = mid - 1;
  }
  u = g.feat_ptr[g.pairs[2 * lo]] + g.matches[2 * e];
  v = g.feat_ptr[g.pairs[2 * lo + 1]] + g.matches[2 * e + 1];
}

__global__ void k_track_init(int n, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
}

__global__ void k_track_hook(const TrackEdges g, int first, int* parent, uint8_t* __restrict__ touched, int* __restrict__ changed) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= g.n_edges) return;
  int u, v;
  track_edge(g, e, u, v);
  if (first) { touched[u] = 1; touched[v] = 1; }
  const int ru = track_root(parent, u), rv = track_root(parent, v);
  if (ru != rv) {
    atomicMin(parent + max(ru, rv), min(ru, rv));
    *changed = 1;
  }
}

__global__ void k_track_compress(int n, int* parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = track_root(parent, i);
}

__global__ void k_track_count(int n, const int* __restrict__ parent, const uint8_t* __restrict__ touched, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && touched[i]) atomicAdd(count + parent[i], 1);
}

// One CTA per image: two features of the image in one track <=> two equal labels among the image's nodes.  The labels
// are sorted in shared memory (bitonic, padded with distinct sentinels) and neighbours compared.
__global__ void __launch_bounds__(256) k_track_conflict(const int* __restrict__ feat_ptr, const int* __restrict__ parent,
                                                         const uint8_t* __restrict__ touched, uint8_t* __restrict__ conflict) {
  extern __shared__ int s_lab[];
  const int i0 = feat_ptr[blockIdx.x], n = feat_ptr[blockIdx.x + 1] - i0;
  int m = 1;
  while (m < n) m <<= 1;
  for (int x = threadIdx.x; x < m; x += blockDim.x)
    s_lab[x] = (x < n && touched[i0 + x]) ? parent[i0 + x] : -1 - x;  // untouched / padding: all different, all negative
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int x = threadIdx.x; x < m; x += blockDim.x) {
        const int y = x ^ j;
        if (y > x) {
          const int a = s_lab[x], b = s_lab[y];
          if (((x & k) == 0) ? (a > b) : (a < b)) { s_lab[x] = b; s_lab[y] = a; }
        }
      }
      __syncthreads();
    }
  for (int x = threadIdx.x; x + 1 < m; x += blockDim.x)
    if (s_lab[x] >= 0 && s_lab[x] == s_lab[x + 1]) conflict[s_lab[x]] = 1;
}

__global__ void k_track_finish(int n, int min_length, const int* __restrict__ parent, const uint8_t* __restrict__ touched,
                               const int* __restrict__ count, const uint8_t* __restrict__ conflict, int* __restrict__ track_of,
                               int* __restrict__ n_tracks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t = -1;
  if (touched[i]) {
    const int r = parent[i];
    if (count[r] >= min_length && !conflict[r]) {
      t = r;
      if (r == i) atomicAdd(n_tracks, 1);
    }
  }
  track_of[i] = t;
}

}  // namespace
}  // namespace pba

using namespace pba;

PBA_API pba_status pba_corner_descriptors(const uint8_t* images, int32_t n_images, int64_t image_stride, int32_t width,
                                          int32_t height, int32_t pitch, const int32_t* corner_ptr, const double* corners,
                                          int32_t rotate_features, int32_t device, double* angles, uint8_t* descriptors) {
  if (n_images < 0 || !corner_ptr || width <= 0 || height <= 0 || pitch < width || image_stride < int64_t(pitch) * height)
    return PBA_ERR_INVALID_ARGUMENT;
  if (corner_ptr[0] != 0) return PBA_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n_images; ++i)
    if (corner_ptr[i + 1] < corner_ptr[i]) return PBA_ERR_INVALID_ARGUMENT;
  const int n = n_images > 0 ? corner_ptr[n_images] : 0;
  if (n > 0 && (!images || !corners || !angles || !descriptors)) return PBA_ERR_INVALID_ARGUMENT;
  // detectKeypoints() only keeps corners InBounds(x, y, EDGE_THRESHOLD) (keypoints.h:146-150); the patch reads
  // rely on it (orientation disc radius 15, rotated test points within 13 sqrt(2) < 19)
  for (int i = 0; i < n; ++i) {
    const double x = corners[2 * i], y = corners[2 * i + 1];
    if (!(x >= kEdgeThreshold && x < double(width - kEdgeThreshold) && y >= kEdgeThreshold && y < double(height - kEdgeThreshold)))
      return PBA_ERR_INVALID_ARGUMENT;
  }
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (n == 0) return PBA_OK;
  DevBuf<uint8_t> d_img, d_desc;
  DevBuf<int> d_ptr;
  DevBuf<double> d_c, d_ang;
  PBA_CUDA_OK(d_img.alloc(size_t(n_images) * image_stride));
  PBA_CUDA_OK(d_ptr.alloc(size_t(n_images) + 1));
  PBA_CUDA_OK(d_c.alloc(size_t(2) * n)); PBA_CUDA_OK(d_ang.alloc(n)); PBA_CUDA_OK(d_desc.alloc(size_t(32) * n));
  PBA_CUDA_OK(cudaMemcpy(d_img.p, images, size_t(n_images) * image_stride, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(d_ptr.p, corner_ptr, sizeof(int) * (size_t(n_images) + 1), cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(d_c.p, corners, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
  k_corner_desc<<<(n + kDescWarps - 1) / kDescWarps, 32 * kDescWarps>>>(n, n_images, d_ptr.p, d_c.p, d_img.p, image_stride,
                                                                         pitch, rotate_features, d_ang.p, d_desc.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(angles, d_ang.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  PBA_CUDA_OK(cudaMemcpy(descriptors, d_desc.p, size_t(32) * n, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_match_descriptors(int32_t n_sets, const int32_t* set_ptr, const uint8_t* descriptors, int32_t n_pairs,
                                         const int32_t* pairs, int32_t threshold, double dist_2_best, int32_t device,
                                         int64_t* match_ptr, int32_t* matches, int64_t capacity) {
  if (n_sets < 0 || n_pairs < 0 || !set_ptr || !match_ptr || capacity < 0 || (n_pairs > 0 && !pairs)) return PBA_ERR_INVALID_ARGUMENT;
  if (set_ptr[0] != 0) return PBA_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n_sets; ++i)
    if (set_ptr[i + 1] < set_ptr[i]) return PBA_ERR_INVALID_ARGUMENT;
  const int n_desc = n_sets > 0 ? set_ptr[n_sets] : 0;
  if (n_desc > 0 && !descriptors) return PBA_ERR_INVALID_ARGUMENT;
  if (capacity > 0 && !matches) return PBA_ERR_INVALID_ARGUMENT;
  std::vector<PairInfo> info(n_pairs);
  int64_t n12 = 0, n21 = 0;
  int max_q = 0;
  for (int k = 0; k < n_pairs; ++k) {
    const int a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || a >= n_sets || b < 0 || b >= n_sets) return PBA_ERR_INVALID_ARGUMENT;
    PairInfo& p = info[k];
    p.a0 = set_ptr[a]; p.na = set_ptr[a + 1] - set_ptr[a];
    p.b0 = set_ptr[b]; p.nb = set_ptr[b + 1] - set_ptr[b];
    p.off12 = n12; p.off21 = n21;
    n12 += p.na; n21 += p.nb;
    max_q = std::max(max_q, std::max(p.na, p.nb));
  }
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  match_ptr[0] = 0;
  if (n_pairs == 0) return PBA_OK;
  DevBuf<PairInfo> d_info;
  DevBuf<uint8_t> d_desc;
  DevBuf<int> d_b12, d_b21, d_m;
  DevBuf<int64_t> d_ptr;
  PBA_CUDA_OK(d_info.alloc(n_pairs)); PBA_CUDA_OK(d_desc.alloc(size_t(32) * std::max(n_desc, 1)));
  PBA_CUDA_OK(d_b12.alloc(size_t(std::max<int64_t>(n12, 1)))); PBA_CUDA_OK(d_b21.alloc(size_t(std::max<int64_t>(n21, 1))));
  PBA_CUDA_OK(d_ptr.alloc(size_t(n_pairs) + 1)); PBA_CUDA_OK(d_m.alloc(size_t(2) * std::max<int64_t>(capacity, 1)));
  PBA_CUDA_OK(cudaMemcpy(d_info.p, info.data(), sizeof(PairInfo) * n_pairs, cudaMemcpyHostToDevice));
  if (n_desc > 0) PBA_CUDA_OK(cudaMemcpy(d_desc.p, descriptors, size_t(32) * n_desc, cudaMemcpyHostToDevice));
  if (max_q > 0) {
    // pairs go in slices of the grid's z limit
    const int tiles = (max_q + kMatchThreads - 1) / kMatchThreads;
    for (int k0 = 0; k0 < n_pairs; k0 += 65535) {
      const int kn = std::min(65535, n_pairs - k0);
      k_match_best<<<dim3(tiles, 2, kn), kMatchThreads>>>(d_info.p + k0, reinterpret_cast<const uint4*>(d_desc.p), threshold,
                                                           dist_2_best, d_b12.p, d_b21.p);
      PBA_CUDA_OK(cudaGetLastError());
    }
  }
  k_match_mutual<false><<<n_pairs, kMatchThreads>>>(d_info.p, d_b12.p, d_b21.p, d_ptr.p, nullptr, 0);
  PBA_CUDA_OK(cudaGetLastError());
  k_scan_counts<<<1, 32>>>(n_pairs, d_ptr.p);
  PBA_CUDA_OK(cudaGetLastError());
  k_match_mutual<true><<<n_pairs, kMatchThreads>>>(d_info.p, d_b12.p, d_b21.p, d_ptr.p, d_m.p, capacity);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(match_ptr, d_ptr.p, sizeof(int64_t) * (size_t(n_pairs) + 1), cudaMemcpyDeviceToHost));
  const int64_t total = match_ptr[n_pairs];
  if (std::min(total, capacity) > 0)
    PBA_CUDA_OK(cudaMemcpy(matches, d_m.p, sizeof(int) * 2 * size_t(std::min(total, capacity)), cudaMemcpyDeviceToHost));
  return total > capacity ? PBA_ERR_INVALID_ARGUMENT : PBA_OK;
}

PBA_API pba_status pba_epipolar_inliers(int32_t model0, const double intr0[8], int32_t model1, const double intr1[8],
                                        const double T_0_1[7], double threshold, int64_t n_matches, const int32_t* matches,
                                        const double* corners0, const double* corners1, int32_t device, double* E_out,
                                        uint8_t* inlier) {
  if (!intr0 || !intr1 || !T_0_1 || n_matches < 0 || (n_matches > 0 && (!matches || !corners0 || !corners1 || !inlier)))
    return PBA_ERR_INVALID_ARGUMENT;
  if (model0 < 0 || model0 > PBA_CAM_EUCM || model1 < 0 || model1 > PBA_CAM_EUCM) return PBA_ERR_UNSUPPORTED;
  EpiArgs a;
  a.n = n_matches; a.model0 = model0; a.model1 = model1; a.threshold = threshold;
  memcpy(a.intr0, intr0, sizeof(a.intr0)); memcpy(a.intr1, intr1, sizeof(a.intr1));
  {
    // computeEssential (matching_utils.h:50-60): E = [t / |t|]x R
    double R[9];
    quat_to_rot(T_0_1, R);
    const double nt = sqrt(T_0_1[4] * T_0_1[4] + T_0_1[5] * T_0_1[5] + T_0_1[6] * T_0_1[6]);
    const double t[3] = {T_0_1[4] / nt, T_0_1[5] / nt, T_0_1[6] / nt};
    const double S[9] = {0.0, -t[2], t[1], t[2], 0.0, -t[0], -t[1], t[0], 0.0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) a.E[3 * i + j] = S[3 * i] * R[j] + S[3 * i + 1] * R[3 + j] + S[3 * i + 2] * R[6 + j];
    if (E_out) memcpy(E_out, a.E, sizeof(a.E));
  }
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (n_matches == 0) return PBA_OK;
  int max0 = 0, max1 = 0;
  for (int64_t k = 0; k < n_matches; ++k) {
    if (matches[2 * k] < 0 || matches[2 * k + 1] < 0) return PBA_ERR_INVALID_ARGUMENT;
    max0 = std::max(max0, matches[2 * k]); max1 = std::max(max1, matches[2 * k + 1]);
  }
  DevBuf<int> d_m;
  DevBuf<double> d_c0, d_c1;
  DevBuf<uint8_t> d_in;
  PBA_CUDA_OK(d_m.alloc(size_t(2) * n_matches)); PBA_CUDA_OK(d_c0.alloc(size_t(2) * (max0 + 1)));
  PBA_CUDA_OK(d_c1.alloc(size_t(2) * (max1 + 1))); PBA_CUDA_OK(d_in.alloc(size_t(n_matches)));
  PBA_CUDA_OK(cudaMemcpy(d_m.p, matches, sizeof(int) * 2 * n_matches, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(d_c0.p, corners0, sizeof(double) * 2 * (size_t(max0) + 1), cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(d_c1.p, corners1, sizeof(double) * 2 * (size_t(max1) + 1), cudaMemcpyHostToDevice));
  k_epipolar<<<unsigned((n_matches + 127) / 128), 128>>>(a, d_m.p, d_c0.p, d_c1.p, d_in.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(inlier, d_in.p, size_t(n_matches), cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_build_tracks(int32_t n_images, const int32_t* feat_ptr, int32_t n_pairs, const int32_t* pairs,
                                    const int64_t* match_ptr, const int32_t* matches, int32_t min_length, int32_t device,
                                    int32_t* track_of, int32_t* n_tracks) {
  if (n_images < 0 || n_pairs < 0 || !feat_ptr || (n_pairs > 0 && (!pairs || !match_ptr))) return PBA_ERR_INVALID_ARGUMENT;
  if (feat_ptr[0] != 0) return PBA_ERR_INVALID_ARGUMENT;
  int max_feat = 0;
  for (int i = 0; i < n_images; ++i) {
    if (feat_ptr[i + 1] < feat_ptr[i]) return PBA_ERR_INVALID_ARGUMENT;
    max_feat = std::max(max_feat, feat_ptr[i + 1] - feat_ptr[i]);
  }
  const int n = n_images > 0 ? feat_ptr[n_images] : 0;
  if (n > 0 && !track_of) return PBA_ERR_INVALID_ARGUMENT;
  const int64_t n_edges = n_pairs > 0 ? match_ptr[n_pairs] : 0;
  if (n_pairs > 0 && match_ptr[0] != 0) return PBA_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && !matches) return PBA_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < n_pairs; ++k) {
    if (match_ptr[k + 1] < match_ptr[k]) return PBA_ERR_INVALID_ARGUMENT;
    const int a = pairs[2 * k], b = pairs[2 * k + 1];
    if (a < 0 || a >= n_images || b < 0 || b >= n_images) return PBA_ERR_INVALID_ARGUMENT;
    const int na = feat_ptr[a + 1] - feat_ptr[a], nb = feat_ptr[b + 1] - feat_ptr[b];
    for (int64_t e = match_ptr[k]; e < match_ptr[k + 1]; ++e)
      if (matches[2 * e] < 0 || matches[2 * e] >= na || matches[2 * e + 1] < 0 || matches[2 * e + 1] >= nb) return PBA_ERR_INVALID_ARGUMENT;
  }
  if (max_feat > kTrackMaxFeat) return PBA_ERR_UNSUPPORTED;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  PBA_CUDA_OK(cudaSetDevice(device));
  if (n_tracks) *n_tracks = 0;
  if (n == 0) return PBA_OK;
  DevBuf<int> d_fp, d_pairs, d_m, d_parent, d_count, d_track, d_flag;
  DevBuf<int64_t> d_mp;
  DevBuf<uint8_t> d_touched, d_conf;
  PBA_CUDA_OK(d_fp.alloc(size_t(n_images) + 1)); PBA_CUDA_OK(d_pairs.alloc(size_t(2) * std::max(n_pairs, 1)));
  PBA_CUDA_OK(d_mp.alloc(size_t(n_pairs) + 1)); PBA_CUDA_OK(d_m.alloc(size_t(2) * std::max<int64_t>(n_edges, 1)));
  PBA_CUDA_OK(d_parent.alloc(n)); PBA_CUDA_OK(d_count.alloc(n)); PBA_CUDA_OK(d_track.alloc(n)); PBA_CUDA_OK(d_flag.alloc(2));
  PBA_CUDA_OK(d_touched.alloc(n)); PBA_CUDA_OK(d_conf.alloc(n));
  PBA_CUDA_OK(cudaMemcpy(d_fp.p, feat_ptr, sizeof(int) * (size_t(n_images) + 1), cudaMemcpyHostToDevice));
  if (n_pairs > 0) {
    PBA_CUDA_OK(cudaMemcpy(d_pairs.p, pairs, sizeof(int) * 2 * n_pairs, cudaMemcpyHostToDevice));
    PBA_CUDA_OK(cudaMemcpy(d_mp.p, match_ptr, sizeof(int64_t) * (size_t(n_pairs) + 1), cudaMemcpyHostToDevice));
  }
  if (n_edges > 0) PBA_CUDA_OK(cudaMemcpy(d_m.p, matches, sizeof(int) * 2 * n_edges, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemset(d_touched.p, 0, n)); PBA_CUDA_OK(cudaMemset(d_conf.p, 0, n));
  PBA_CUDA_OK(cudaMemset(d_count.p, 0, sizeof(int) * n)); PBA_CUDA_OK(cudaMemset(d_flag.p, 0, sizeof(int) * 2));
  const unsigned gn = unsigned((n + 255) / 256);
  k_track_init<<<gn, 256>>>(n, d_parent.p);
  PBA_CUDA_OK(cudaGetLastError());
  if (n_edges > 0) {
    TrackEdges g;
    g.n_pairs = n_pairs; g.n_edges = n_edges; g.feat_ptr = d_fp.p; g.pairs = d_pairs.p; g.match_ptr = d_mp.p; g.matches = d_m.p;
    const unsigned ge = unsigned((n_edges + 255) / 256);
    for (int round = 0; round < 64; ++round) {  // a component of diameter D settles in O(log D) rounds
      PBA_CUDA_OK(cudaMemset(d_flag.p, 0, sizeof(int)));
      k_track_hook<<<ge, 256>>>(g, round == 0 ? 1 : 0, d_parent.p, d_touched.p, d_flag.p);
      k_track_compress<<<gn, 256>>>(n, d_parent.p);
      PBA_CUDA_OK(cudaGetLastError());
      int changed = 0;
      PBA_CUDA_OK(cudaMemcpy(&changed, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
      if (!changed) break;
      if (round == 63) return PBA_ERR_NUMERICAL_FAILURE;
    }
  }
  k_track_count<<<gn, 256>>>(n, d_parent.p, d_touched.p, d_count.p);
  int m = 1;
  while (m < max_feat) m <<= 1;
  if (n_images > 0 && max_feat > 0) {
    PBA_CUDA_OK(ensure_dynamic_smem((const void*)k_track_conflict, sizeof(int) * size_t(m), device));
    k_track_conflict<<<n_images, 256, sizeof(int) * size_t(m)>>>(d_fp.p, d_parent.p, d_touched.p, d_conf.p);
  }
  k_track_finish<<<gn, 256>>>(n, min_length, d_parent.p, d_touched.p, d_count.p, d_conf.p, d_track.p, d_flag.p + 1);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(track_of, d_track.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (n_tracks) PBA_CUDA_OK(cudaMemcpy(n_tracks, d_flag.p + 1, sizeof(int), cudaMemcpyDeviceToHost));
  return PBA_OK;
}
