// deposit.cu -- mass assignment (mesh.NGP / CIC / TSC, mesh.py:2240-2595) fused with solver.pm's
// density rescale and rhs_poisson's affine map (solver.py:114-116, 444-449).
//
// One thread per particle, RED.ADD.F32 to the global grid: correct for any N and any particle order.  This is the
// path of meshes the cell-sorted kernels of binned.cu do not take (N < 8 or N % 8 != 0; mesh.can_bin) -- every
// BASELINE configuration goes through psc_deposit_binned.
#include "common.cuh"

namespace psc {

template <int SCHEME>
__global__ void __launch_bounds__(256) deposit_atomic_kernel(const float *__restrict__ pos, int64_t np,
                                                             int N, float *__restrict__ rho) {
  const float Nf = (float)N;
  const size_t N2 = (size_t)N * N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np;
       n += (int64_t)gridDim.x * blockDim.x) {
    float xp = __ldg(&pos[3 * n + 0]) * Nf, yp = __ldg(&pos[3 * n + 1]) * Nf,
          zp = __ldg(&pos[3 * n + 2]) * Nf;
    if (SCHEME == PSC_NGP) {
      int i = (int)xp, j = (int)yp, k = (int)zp;
      atomicAdd(&rho[(size_t)i * N2 + (size_t)j * N + k], 1.0f);
    } else if (SCHEME == PSC_CIC) {
      int i, j, k, i2, j2, k2;
      float wx, wy, wz, dx, dy, dz;
      cic_axis(xp, N, i, i2, wx, dx);
      cic_axis(yp, N, j, j2, wy, dy);
      cic_axis(zp, N, k, k2, wz, dz);
      size_t r0 = (size_t)i * N2, r1 = (size_t)i2 * N2, c0 = (size_t)j * N, c1 = (size_t)j2 * N;
      atomicAdd(&rho[r0 + c0 + k], wx * wy * wz);
      atomicAdd(&rho[r0 + c0 + k2], wx * wy * dz);
      atomicAdd(&rho[r0 + c1 + k], wx * dy * wz);
      atomicAdd(&rho[r0 + c1 + k2], wx * dy * dz);
      atomicAdd(&rho[r1 + c0 + k], dx * wy * wz);
      atomicAdd(&rho[r1 + c0 + k2], dx * wy * dz);
      atomicAdd(&rho[r1 + c1 + k], dx * dy * wz);
      atomicAdd(&rho[r1 + c1 + k2], dx * dy * dz);
    } else {
      int i, j, k;
      float wx[3], wy[3], wz[3];
      tsc_axis(xp, i, wx[0], wx[1], wx[2]);
      tsc_axis(yp, j, wy[0], wy[1], wy[2]);
      tsc_axis(zp, k, wz[0], wz[1], wz[2]);
      int kk[3] = {wrap(k - 1, N), k, wrap(k + 1, N)};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        size_t r = (size_t)wrap(i + a - 1, N) * N2;
#pragma unroll
        for (int b = 0; b < 3; b++) {
          size_t c = r + (size_t)wrap(j + b - 1, N) * N;
          float wxy = wx[a] * wy[b];
#pragma unroll
          for (int g = 0; g < 3; g++) atomicAdd(&rho[c + kk[g]], wxy * wz[g]);
        }
      }
    }
  }
}

// rho = f1 * (scale * rho) + f2  (evaluated in the reference's order: scale first, then affine)
__global__ void rho_affine_kernel(float *rho, int64_t n, float scale, float f1, float f2, int do_scale);
__global__ void __launch_bounds__(256) rho_affine_kernel(float *__restrict__ rho, int64_t n, float scale,
                                                         float f1, float f2, int do_scale) {
  int64_t n4 = n >> 2;
  float4 *r4 = reinterpret_cast<float4 *>(rho);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = r4[i];
    if (do_scale) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
    v.x = f1 * v.x + f2; v.y = f1 * v.y + f2; v.z = f1 * v.z + f2; v.w = f1 * v.w + f2;
    r4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float v = rho[i];
    if (do_scale) v *= scale;
    rho[i] = f1 * v + f2;
  }
}

}  // namespace psc

using namespace psc;

extern "C" int psc_deposit(const float *pos, int64_t np, int N, int scheme, float scale, float f1,
                           float f2, float *rho, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  PSC_CHECK_ARG(N >= 2 && N <= 32767, "N out of range (cell indices are int16 in the reference)");
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  PSC_CHECK_ARG(rho && (pos || np == 0), "null pointer");
  cudaStream_t st = as_stream(stream);
  const int64_t n3 = (int64_t)N * N * N;

  PSC_CUDA(cudaMemsetAsync(rho, 0, sizeof(float) * n3, st));
  if (np > 0) {
    int g = grid_for(np, 256, 16);
    if (scheme == PSC_NGP)
      deposit_atomic_kernel<PSC_NGP><<<g, 256, 0, st>>>(pos, np, N, rho);
    else if (scheme == PSC_CIC)
      deposit_atomic_kernel<PSC_CIC><<<g, 256, 0, st>>>(pos, np, N, rho);
    else
      deposit_atomic_kernel<PSC_TSC><<<g, 256, 0, st>>>(pos, np, N, rho);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  if (scale != 1.0f || f1 != 1.0f || f2 != 0.0f) {
    rho_affine_kernel<<<grid_for((n3 + 3) / 4, 256), 256, 0, st>>>(rho, n3, scale, f1, f2,
                                                                   scale != 1.0f);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return PSC_OK;
}
