// launch.h — kernel launch accounting.  Every launch is counted per kernel id
// (reported as gpu_launches by bench.py); with options.profile = 1 each launch
// is also bracketed by CUDA events on the launching stream.
#pragma once
#include "pba_internal.h"

#define PBA_LAUNCH(h, id, kernel, grid, block, smem, ...)                 \
  do {                                                                    \
    (h)->stats.begin((id), (h)->stream);                                  \
    kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);        \
    (h)->stats.end((h)->stream);                                          \
    cudaError_t _le = cudaGetLastError();                                 \
    if (_le != cudaSuccess) return ::pba::map_cuda(_le);                  \
  } while (0)
