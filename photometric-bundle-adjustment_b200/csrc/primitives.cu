// primitives.cu — stand-alone device primitives exposed through the C ABI so
// the parity tests can pin them one by one against the reference:
// camera models (include/visnav/camera_models.h), LocalParameterizationSE3::Plus
// (include/visnav/local_parameterization_se3.hpp:44-51) and the dense DMMA Cholesky.
#include <string.h>

#include "launch.h"
#include "pba_internal.h"

#define PBA_API extern "C" __attribute__((visibility("default")))

namespace pba {
namespace {

struct Intr { double v[8]; };

__global__ void k_project(int model, Intr in, int64_t n, const double* __restrict__ xyz, double* __restrict__ uv,
                          double* __restrict__ J) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[2], Jp[6];
  cam_project<true>(model, in.v, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], o, Jp);
  uv[2 * i] = o[0]; uv[2 * i + 1] = o[1];
  if (J) for (int k = 0; k < 6; ++k) J[6 * i + k] = Jp[k];
}

__global__ void k_unproject(int model, Intr in, int64_t n, const double* __restrict__ uv, double* __restrict__ xyz) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[3];
  cam_unproject(model, in.v, uv[2 * i], uv[2 * i + 1], o);
  xyz[3 * i] = o[0]; xyz[3 * i + 1] = o[1]; xyz[3 * i + 2] = o[2];
}

__global__ void k_se3_plus(int64_t n, const double* __restrict__ T, const double* __restrict__ d, double* __restrict__ o) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Ti[7], di[6], oi[7];
  for (int k = 0; k < 7; ++k) Ti[k] = T[7 * i + k];
  for (int k = 0; k < 6; ++k) di[k] = d[6 * i + k];
  se3_plus(Ti, di, oi);
  for (int k = 0; k < 7; ++k) o[7 * i + k] = oi[k];
}

__global__ void k_pad_dense(int n, int ld, const double* __restrict__ A, const double* __restrict__ b,
                            double* __restrict__ Ap, double* __restrict__ bp) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(ld) * ld) return;
  const int r = int(i / ld), c = int(i % ld);
  Ap[i] = (r < n && c < n) ? A[int64_t(r) * n + c] : (r == c ? 1.0 : 0.0);
  if (c == 0) bp[r] = r < n ? b[r] : 0.0;
}

bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

}  // namespace
}  // namespace pba

using namespace pba;

PBA_API pba_status pba_camera_project(int32_t model, const double intr[8], int64_t n, const double* xyz, double* uv,
                                      double* duv_dxyz) {
  if (!intr || !xyz || !uv || n < 0 || model < 0 || model > PBA_CAM_EUCM) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  if (n == 0) return PBA_OK;
  DevBuf<double> dx, du, dj;
  PBA_CUDA_OK(dx.alloc(3 * n)); PBA_CUDA_OK(du.alloc(2 * n));
  if (duv_dxyz) PBA_CUDA_OK(dj.alloc(6 * n));
  PBA_CUDA_OK(cudaMemcpy(dx.p, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  Intr in;
  memcpy(in.v, intr, sizeof(in.v));
  k_project<<<unsigned((n + 127) / 128), 128>>>(model, in, n, dx.p, du.p, dj.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(uv, du.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  if (duv_dxyz) PBA_CUDA_OK(cudaMemcpy(duv_dxyz, dj.p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_camera_unproject(int32_t model, const double intr[8], int64_t n, const double* uv, double* xyz) {
  if (!intr || !xyz || !uv || n < 0 || model < 0 || model > PBA_CAM_EUCM) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  if (n == 0) return PBA_OK;
  DevBuf<double> dx, du;
  PBA_CUDA_OK(dx.alloc(3 * n)); PBA_CUDA_OK(du.alloc(2 * n));
  PBA_CUDA_OK(cudaMemcpy(du.p, uv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
  Intr in;
  memcpy(in.v, intr, sizeof(in.v));
  k_unproject<<<unsigned((n + 127) / 128), 128>>>(model, in, n, du.p, dx.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(xyz, dx.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_se3_plus(int64_t n, const double* poses7, const double* delta6, double* out7) {
  if (!poses7 || !delta6 || !out7 || n < 0) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  if (n == 0) return PBA_OK;
  DevBuf<double> a, b, c;
  PBA_CUDA_OK(a.alloc(7 * n)); PBA_CUDA_OK(b.alloc(6 * n)); PBA_CUDA_OK(c.alloc(7 * n));
  PBA_CUDA_OK(cudaMemcpy(a.p, poses7, sizeof(double) * 7 * n, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(b.p, delta6, sizeof(double) * 6 * n, cudaMemcpyHostToDevice));
  k_se3_plus<<<unsigned((n + 127) / 128), 128>>>(n, a.p, b.p, c.p);
  PBA_CUDA_OK(cudaGetLastError());
  PBA_CUDA_OK(cudaMemcpy(out7, c.p, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost));
  return PBA_OK;
}

PBA_API pba_status pba_cholesky_solve(int32_t n, const double* A, const double* b, double* x) {
  if (!A || !b || !x || n <= 0) return PBA_ERR_INVALID_ARGUMENT;
  if (!have_device()) return PBA_ERR_NO_DEVICE;
  Handle h;  // only its stream (default), device and launch counters are used
  PBA_CUDA_OK(cudaGetDevice(&h.device));
  const int ld = dense_ld(n);
  DevBuf<double> dA, db, dAp;
  DevBuf<int> fail;
  PBA_CUDA_OK(dA.alloc(size_t(n) * n)); PBA_CUDA_OK(db.alloc(n)); PBA_CUDA_OK(dAp.alloc(size_t(ld) * ld + ld + dense_work_size(ld)));
  PBA_CUDA_OK(fail.alloc(1));
  PBA_CUDA_OK(cudaMemset(fail.p, 0, sizeof(int)));
  PBA_CUDA_OK(cudaMemcpy(dA.p, A, sizeof(double) * size_t(n) * n, cudaMemcpyHostToDevice));
  PBA_CUDA_OK(cudaMemcpy(db.p, b, sizeof(double) * n, cudaMemcpyHostToDevice));
  double* bp = dAp.p + size_t(ld) * ld;
  const int64_t tot = int64_t(ld) * ld;
  k_pad_dense<<<unsigned((tot + 255) / 256), 256>>>(n, ld, dA.p, db.p, dAp.p, bp);
  PBA_CUDA_OK(cudaGetLastError());
  pba_status st = dense_cholesky_solve(&h, dAp.p, bp, ld, fail.p, bp + ld);
  if (st != PBA_OK) return st;
  int f = 0;
  PBA_CUDA_OK(cudaMemcpy(&f, fail.p, sizeof(int), cudaMemcpyDeviceToHost));
  PBA_CUDA_OK(cudaMemcpy(x, bp, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return f ? PBA_ERR_NUMERICAL_FAILURE : PBA_OK;
}
