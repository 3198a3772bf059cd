/* pba_synth.h — deterministic synthetic BA scenes.
 * Benchmark / test utility, not part of the drop-in boundary: the reference
 * has no scene generator.  Host generator in libpba_synth.so (csrc/synth.cpp);
 * the GPU renderer (same scene, for 2,000-keyframe benchmarks) in
 * libpba_b200.so (csrc/synth_gpu.cu). */
#ifndef PBA_SYNTH_H_
#define PBA_SYNTH_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pba_synth_params {
  int32_t mode;  /* PBA_MODE_* */
  int32_t n_kf;
  int32_t n_pts;
  int32_t model; /* PBA_CAM_* */
  int32_t width, height;
  int32_t min_len, max_len; /* observation window length incl. host (8..12) */
  uint64_t seed_pix, seed_vis, seed_noise; /* 1234, 99, 42 */
  double pose_sigma;  /* tangent-space sigma of the pose perturbation */
  double rho_sigma;   /* rho <- rho / (1 + N(0, sigma^2)) */
  double pixel_sigma; /* geometric: target pixel noise */
  double affine_a_sigma, affine_b_sigma;
  double intrinsics[8];
} pba_synth_params;

void pba_synth_default_params(pba_synth_params* p, int mode, int n_kf, int n_pts, int model);
int64_t pba_synth_count_obs(const pba_synth_params* p);
int pba_synth_generate(const pba_synth_params* p, double* poses_gt, double* poses, uint8_t* pose_fixed,
                       double* inv_depth_gt, double* inv_depth, int32_t* lm_host, double* lm_host_uv,
                       int64_t* lm_obs_ptr, int32_t* obs_target, double* obs_uv, double* affine);
/* CPU renderer (OpenMP): keyframes [first, first+count) into images[count*pitch*height]. */
int pba_synth_render(const pba_synth_params* p, int first, int count, int pitch, uint8_t* images);
/* GPU renderer: same scene on the current CUDA device, copied to the host buffer.
 * Returns a pba_status. (Pixel values may differ from the CPU renderer by 1 grey
 * level at rounding ties; both engines always consume the same bytes.) */
int pba_synth_render_gpu(const pba_synth_params* p, int first, int count, int pitch, uint8_t* images_host);

#ifdef __cplusplus
}
#endif
#endif
