// interp.cu -- grid -> particle interpolation (mesh.invNGP/invCIC/invTSC[_vec], mesh.py:2600-3088),
// optionally fused with the second leapfrog half-kick and the max|a|, max|v| reductions
// (integration.py:262, 293-295, 324-326).
//
// Design notes (B200, 512^3, measured -- profiles/r01_bench_kernels_interp_variants.txt and
// profiles/r01_interp4_direct_ncu.txt): the gather is bound by the L1 data pipe (91% of peak), not by
// HBM: a Morton chunk of 32 particles spans 4x4 (i,j) rows, so every warp-wide load touches ~16 lines.
//   direct gather, AoS float3 grid, 81 LDG.32 / particle ........ 9.5 ms
//   direct gather, float4-padded grid, 27 LDG.128 / particle ..... 5.9 ms   <- used by the fused step
//   CTA-shared 12x12x16 smem window (fixed 256-particle slices) .. 8.0 ms  (window reloaded every slice)
//   CTA-shared window, adaptive block-aligned slices ............. 7.0 ms  (3 barriers per slice dominate)
//   per-warp 8^3 smem window, LDG+STS / cp.async fill ............ 10.7 / 7.1 ms
// The shared-memory variants were removed again; the float4 layout is the internal force layout of
// solver.pm (psc_gradient with out_stride = 4).
#include "common.cuh"

namespace psc {

template <int SCHEME, int NCOMP>
__device__ __forceinline__ void interp_one(const float *__restrict__ grid, int N, float x, float y,
                                           float z, float (&acc)[NCOMP]) {
  const float Nf = (float)N;
  const size_t N2 = (size_t)N * N;
  float xp = x * Nf, yp = y * Nf, zp = z * Nf;
  if (SCHEME == PSC_NGP) {
    int i = (int)xp, j = (int)yp, k = (int)zp;
    const float *g = grid + ((size_t)i * N2 + (size_t)j * N + k) * NCOMP;
#pragma unroll
    for (int m = 0; m < NCOMP; m++) acc[m] = __ldg(g + m);
  } else if (SCHEME == PSC_CIC) {
    int i, j, k, i2, j2, k2;
    float wx, wy, wz, dx, dy, dz;
    cic_axis(xp, N, i, i2, wx, dx);
    cic_axis(yp, N, j, j2, wy, dy);
    cic_axis(zp, N, k, k2, wz, dz);
    size_t r[2] = {(size_t)i * N2, (size_t)i2 * N2}, c[2] = {(size_t)j * N, (size_t)j2 * N};
    int kk[2] = {k, k2};
    float ax[2] = {wx, dx}, ay[2] = {wy, dy}, az[2] = {wz, dz};
#pragma unroll
    for (int m = 0; m < NCOMP; m++) acc[m] = 0.0f;
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int g = 0; g < 2; g++) {
          float w = ax[a] * ay[b] * az[g];
          const float *p = grid + (r[a] + c[b] + kk[g]) * NCOMP;
#pragma unroll
          for (int m = 0; m < NCOMP; m++) acc[m] += w * __ldg(p + m);
        }
  } else {
    int i, j, k;
    float wx[3], wy[3], wz[3];
    tsc_axis(xp, i, wx[0], wx[1], wx[2]);
    tsc_axis(yp, j, wy[0], wy[1], wy[2]);
    tsc_axis(zp, k, wz[0], wz[1], wz[2]);
    int kk[3] = {wrap(k - 1, N), k, wrap(k + 1, N)};
#pragma unroll
    for (int m = 0; m < NCOMP; m++) acc[m] = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      size_t r = (size_t)wrap(i + a - 1, N) * N2;
#pragma unroll
      for (int b = 0; b < 3; b++) {
        size_t c = r + (size_t)wrap(j + b - 1, N) * N;
        float wxy = wx[a] * wy[b];
#pragma unroll
        for (int g = 0; g < 3; g++) {
          float w = wxy * wz[g];
          const float *p = grid + (c + kk[g]) * NCOMP;
#pragma unroll
          for (int m = 0; m < NCOMP; m++) acc[m] += w * __ldg(p + m);
        }
      }
    }
  }
}

template <int SCHEME, int NCOMP>
__global__ void __launch_bounds__(256) interp_kernel(const float *__restrict__ grid,
                                                     const float *__restrict__ pos, int64_t np, int N,
                                                     float *__restrict__ out) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np;
       n += (int64_t)gridDim.x * blockDim.x) {
    float acc[NCOMP];
    interp_one<SCHEME, NCOMP>(grid, N, __ldg(&pos[3 * n]), __ldg(&pos[3 * n + 1]), __ldg(&pos[3 * n + 2]), acc);
#pragma unroll
    for (int m = 0; m < NCOMP; m++) out[NCOMP * n + m] = acc[m];
  }
}

template <int SCHEME>
__global__ void __launch_bounds__(256) interp_kick_kernel(const float *__restrict__ force,
                                                          const float *__restrict__ pos,
                                                          float *__restrict__ vel, float *__restrict__ accel,
                                                          int64_t np, int N, float half_dt,
                                                          float *__restrict__ maxout) {
  float ma = 0.0f, mv = 0.0f;
  const float mh = -half_dt;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np;
       n += (int64_t)gridDim.x * blockDim.x) {
    float a[3];
    interp_one<SCHEME, 3>(force, N, __ldg(&pos[3 * n]), __ldg(&pos[3 * n + 1]), __ldg(&pos[3 * n + 2]), a);
    accel[3 * n + 0] = a[0];
    accel[3 * n + 1] = a[1];
    accel[3 * n + 2] = a[2];
    ma = fmaxf(ma, fmaxf(fabsf(a[0]), fmaxf(fabsf(a[1]), fabsf(a[2]))));
    if (vel) {
      float v0 = vel[3 * n + 0] + mh * a[0], v1 = vel[3 * n + 1] + mh * a[1],
            v2 = vel[3 * n + 2] + mh * a[2];
      vel[3 * n + 0] = v0;
      vel[3 * n + 1] = v1;
      vel[3 * n + 2] = v2;
      mv = fmaxf(mv, fmaxf(fabsf(v0), fmaxf(fabsf(v1), fabsf(v2))));
    }
  }
  ma = warp_max(ma);
  mv = warp_max(mv);
  __shared__ float sa[8], sv[8];
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = ma;
    sv[threadIdx.x >> 5] = mv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) {
      ma = fmaxf(ma, sa[w]);
      mv = fmaxf(mv, sv[w]);
    }
    atomic_max_nonneg(&maxout[0], ma);
    atomic_max_nonneg(&maxout[1], mv);
  }
}

// Same fused kernel on a float4-padded force grid [N,N,N] x (fx,fy,fz,pad): one LDG.128 per stencil
// point instead of three LDG.32 (the direct gather is bound by L1 wavefronts, not by HBM).
template <int SCHEME>
__global__ void __launch_bounds__(256) interp_kick4_kernel(const float4 *__restrict__ force4,
                                                           const float *__restrict__ pos, float *__restrict__ vel,
                                                           float *__restrict__ accel, int64_t np, int N,
                                                           float half_dt, float *__restrict__ maxout) {
  float ma = 0.0f, mv = 0.0f;
  const float mh = -half_dt;
  const float Nf = (float)N;
  const size_t N2 = (size_t)N * N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    const float xp = __ldg(&pos[3 * n]) * Nf, yp = __ldg(&pos[3 * n + 1]) * Nf, zp = __ldg(&pos[3 * n + 2]) * Nf;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (SCHEME == PSC_TSC) {
      int i, j, k;
      float wx[3], wy[3], wz[3];
      tsc_axis(xp, i, wx[0], wx[1], wx[2]);
      tsc_axis(yp, j, wy[0], wy[1], wy[2]);
      tsc_axis(zp, k, wz[0], wz[1], wz[2]);
      const int kk[3] = {wrap(k - 1, N), k, wrap(k + 1, N)};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const size_t r = (size_t)wrap(i + a - 1, N) * N2;
#pragma unroll
        for (int b = 0; b < 3; b++) {
          const size_t c = r + (size_t)wrap(j + b - 1, N) * N;
          const float wxy = wx[a] * wy[b];
#pragma unroll
          for (int g = 0; g < 3; g++) {
            const float w = wxy * wz[g];
            const float4 f = __ldg(&force4[c + kk[g]]);
            ax += w * f.x; ay += w * f.y; az += w * f.z;
          }
        }
      }
    } else {
      int i, j, k, i2, j2, k2;
      float wx, wy, wz, dx, dy, dz;
      cic_axis(xp, N, i, i2, wx, dx);
      cic_axis(yp, N, j, j2, wy, dy);
      cic_axis(zp, N, k, k2, wz, dz);
      const size_t r[2] = {(size_t)i * N2, (size_t)i2 * N2}, c[2] = {(size_t)j * N, (size_t)j2 * N};
      const int kk[2] = {k, k2};
      const float fx[2] = {wx, dx}, fy[2] = {wy, dy}, fz[2] = {wz, dz};
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int g = 0; g < 2; g++) {
            const float w = fx[a] * fy[b] * fz[g];
            const float4 f = __ldg(&force4[r[a] + c[b] + kk[g]]);
            ax += w * f.x; ay += w * f.y; az += w * f.z;
          }
    }
    accel[3 * n + 0] = ax; accel[3 * n + 1] = ay; accel[3 * n + 2] = az;
    ma = fmaxf(ma, fmaxf(fabsf(ax), fmaxf(fabsf(ay), fabsf(az))));
    if (vel) {
      const float v0 = vel[3 * n + 0] + mh * ax, v1 = vel[3 * n + 1] + mh * ay, v2 = vel[3 * n + 2] + mh * az;
      vel[3 * n + 0] = v0; vel[3 * n + 1] = v1; vel[3 * n + 2] = v2;
      mv = fmaxf(mv, fmaxf(fabsf(v0), fmaxf(fabsf(v1), fabsf(v2))));
    }
  }
  ma = warp_max(ma);
  mv = warp_max(mv);
  __shared__ float sa[8], sv[8];
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = ma; sv[threadIdx.x >> 5] = mv; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; w++) { ma = fmaxf(ma, sa[w]); mv = fmaxf(mv, sv[w]); }
    atomic_max_nonneg(&maxout[0], ma);
    atomic_max_nonneg(&maxout[1], mv);
  }
}



}  // namespace psc

using namespace psc;

extern "C" {

int psc_interp(const float *grid, const float *pos, int64_t np, int N, int ncomp, int scheme,
               float *out, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  PSC_CHECK_ARG(N >= 2 && N <= 32767, "N out of range");
  PSC_CHECK_ARG(ncomp == 1 || ncomp == 3, "ncomp must be 1 or 3");
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(grid && pos && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  int g = grid_for(np, 256, 16);
#define PSC_LAUNCH(S, C) interp_kernel<S, C><<<g, 256, 0, st>>>(grid, pos, np, N, out)
  if (ncomp == 1) {
    if (scheme == PSC_NGP) PSC_LAUNCH(PSC_NGP, 1);
    else if (scheme == PSC_CIC) PSC_LAUNCH(PSC_CIC, 1);
    else PSC_LAUNCH(PSC_TSC, 1);
  } else {
    if (scheme == PSC_NGP) PSC_LAUNCH(PSC_NGP, 3);
    else if (scheme == PSC_CIC) PSC_LAUNCH(PSC_CIC, 3);
    else PSC_LAUNCH(PSC_TSC, 3);
  }
#undef PSC_LAUNCH
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_interp_kick(const float *force, const float *pos, float *vel, float *acc, int64_t np, int N,
                    int scheme, float half_dt, float *maxout, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  PSC_CHECK_ARG(N >= 2 && N <= 32767, "N out of range");
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(force && pos && acc && maxout, "null pointer");
  cudaStream_t st = as_stream(stream);
  int g = grid_for(np, 256, 16);
  if (scheme == PSC_NGP)
    interp_kick_kernel<PSC_NGP><<<g, 256, 0, st>>>(force, pos, vel, acc, np, N, half_dt, maxout);
  else if (scheme == PSC_CIC)
    interp_kick_kernel<PSC_CIC><<<g, 256, 0, st>>>(force, pos, vel, acc, np, N, half_dt, maxout);
  else
    interp_kick_kernel<PSC_TSC><<<g, 256, 0, st>>>(force, pos, vel, acc, np, N, half_dt, maxout);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_interp_kick4(const float *force4, const float *pos, float *vel, float *acc, int64_t np, int N,
                     int scheme, float half_dt, float *maxout, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  PSC_CHECK_ARG(N >= 2 && N <= 32767, "N out of range");
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(force4 && pos && acc && maxout, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)force4 & 15) == 0, "force4 must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  int g = grid_for(np, 256, 16);
  const float4 *f4 = reinterpret_cast<const float4 *>(force4);
  if (scheme == PSC_CIC)
    interp_kick4_kernel<PSC_CIC><<<g, 256, 0, st>>>(f4, pos, vel, acc, np, N, half_dt, maxout);
  else
    interp_kick4_kernel<PSC_TSC><<<g, 256, 0, st>>>(f4, pos, vel, acc, np, N, half_dt, maxout);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
