// synth_gpu.cu — GPU renderer of the synthetic scene (benchmark utility; the
// 2,000-keyframe workload has 722 M pixels to ray-cast).  Same geometry as the
// CPU generator (synth_scene.h); not part of the BA hot path.
#include "pba_internal.h"
#include "pba_synth.h"
#include "synth_scene.h"

namespace {

struct RenderArgs {
  int model, width, height, pitch, first, count;
  double intr[8];
};

__global__ void __launch_bounds__(256) k_synth_render(RenderArgs a, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int i = blockIdx.z;
  if (x >= a.width) return;
  double b[3], T[7], X[3];
  pba::cam_bearing(a.model, a.intr, double(x), double(y), b);
  pba_scene::kf_pose(a.first + i, T);
  pba_scene::ray_wall(T, b, X);
  double f = pba_scene::wall_texture(X[0], X[1]);
  f = f < 0.0 ? 0.0 : (f > 255.0 ? 255.0 : f);
  out[(size_t(i) * a.height + y) * a.pitch + x] = uint8_t(floor(f + 0.5));
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int pba_synth_render_gpu(const pba_synth_params* p, int first,
                                                                            int count, int pitch,
                                                                            uint8_t* images_host) {
  if (!p || !images_host || count < 0 || pitch < p->width) return PBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return PBA_ERR_NO_DEVICE; }
  RenderArgs a;
  a.model = p->model; a.width = p->width; a.height = p->height; a.pitch = pitch;
  for (int k = 0; k < 8; ++k) a.intr[k] = p->intrinsics[k];
  const int batch = 256;  // keyframes per launch (gridDim.z limit 65535; keeps the staging buffer small)
  pba::DevBuf<uint8_t> buf;
  const size_t img = size_t(pitch) * p->height;
  if (buf.alloc(img * size_t(batch < count ? batch : (count > 0 ? count : 1))) != cudaSuccess) return PBA_ERR_OUT_OF_MEMORY;
  for (int f0 = 0; f0 < count; f0 += batch) {
    const int c = count - f0 < batch ? count - f0 : batch;
    a.first = first + f0; a.count = c;
    if (pitch > p->width) cudaMemset(buf.p, 0, img * c);
    k_synth_render<<<dim3((p->width + 255) / 256, p->height, c), 256>>>(a, buf.p);
    if (cudaGetLastError() != cudaSuccess) return PBA_ERR_CUDA;
    if (cudaMemcpy(images_host + img * f0, buf.p, img * c, cudaMemcpyDeviceToHost) != cudaSuccess) return PBA_ERR_CUDA;
  }
  return PBA_OK;
}
