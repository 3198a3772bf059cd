// launch.h — kernel launch accounting.  Every launch is counted per kernel id
// (reported as gpu_launches by bench.py); with options.profile = 1 each launch
// is also bracketed by CUDA events on the launching stream.
#pragma once
#include "pba_internal.h"

namespace pba {
// Dynamic shared memory above 48 KB needs cudaFuncAttributeMaxDynamicSharedMemorySize, which is a
// property of (kernel, DEVICE) shared by every live handle.  The opt-in is tracked per (device,
// kernel) as a running maximum under a mutex and raised only when a launch needs more, so a handle
// on a second device, a smaller problem created next to a larger one, and concurrent host threads
// (single-process multi-GPU) all see a sufficient limit.  Defined in host.cu.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes, int device);
}  // namespace pba

#define PBA_LAUNCH(h, id, kernel, grid, block, smem, ...)                                     \
  do {                                                                                        \
    if (size_t(smem) > 48 * 1024) {                                                           \
      cudaError_t _se = ::pba::ensure_dynamic_smem((const void*)(kernel), (smem), (h)->device); \
      if (_se != cudaSuccess) return ::pba::map_cuda(_se);                                    \
    }                                                                                         \
    (h)->stats.begin((id), (h)->stream);                                                      \
    kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);                            \
    (h)->stats.end((h)->stream);                                                              \
    cudaError_t _le = cudaGetLastError();                                                     \
    if (_le != cudaSuccess) return ::pba::map_cuda(_le);                                      \
  } while (0)
