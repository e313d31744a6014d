// common.cuh -- shared helpers for libpysco_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pysco_b200.h"

namespace psc {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Streaming kernels are launched as grid-stride loops over a multiple of the SM count of the current device
// (num_sms(): cudaDevAttrMultiProcessorCount, cached; 148 on a B200).
constexpr int kNumSMsB200 = 148;
int num_sms();

#define PSC_CHECK_ARG(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) {                                                \
      psc::set_error("%s: invalid argument: %s", __func__, msg);  \
      return PSC_ERR_INVALID;                                     \
    }                                                             \
  } while (0)

#define PSC_CHECK_LAUNCH()                                                          \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      psc::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__));      \
      return PSC_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define PSC_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      psc::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
      return PSC_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// grid size for a grid-stride loop over n items with `block` threads, capped at `waves` CTAs per SM
static inline int grid_for(int64_t n, int block, int ctas_per_sm = 8) {
  int64_t need = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// periodic index wrap for |offset| < N (the reference relies on numpy negative indexing)
__device__ __forceinline__ int wrap(int i, int N) {
  i = i < 0 ? i + N : i;
  return i >= N ? i - N : i;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// non-negative floats order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
  atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// utils.periodic_wrap (utils.py:1131-1149): tiny negatives snap to 0 (t + 1 would round to 1.0), else shift by one box
__device__ __forceinline__ float wrap01(float t) {
  const double eps = -2.98023223876953125e-08 * (1.0 + 1e-6);  // -(2^-25) * (1 + 1e-6)
  if (t < 0.0f) return ((double)t > eps) ? 0.0f : t + 1.0f;
  if (t >= 1.0f) return t - 1.0f;
  return t;
}

// integration.py:250-258 for one component: v -= half_dt a (mh = -half_dt); x += dt v (in float32, or in float64 when the
// time step was clamped to a snapshot time: F64); periodic wrap.  Explicit fused multiply-adds: every kernel that
// applies -- or predicts -- this update must produce the SAME bits, because bins are counted by one kernel and filled
// by another.
template <bool F64>
__device__ __forceinline__ void kick_drift_wrap1(float &x, float &v, float a, float mh, float dtf, double dt) {
  v = __fmaf_rn(mh, a, v);
  x = F64 ? (float)__fma_rn(dt, (double)v, (double)x) : __fmaf_rn(dtf, v, x);
  x = wrap01(x);
}

// TSC cell + the three 1-D weights of one axis (mesh.py:2502-2522): xp = x*N (float32)
__device__ __forceinline__ void tsc_axis(float xp, int &c, float &wm, float &w0, float &wp) {
  c = (int)xp;  // trunc == floor for xp >= 0 (reference: np.int16(xp))
  float d = xp - 0.5f - (float)c;
  w0 = 0.75f - d * d;
  float m = 0.5f - d, p = 0.5f + d;
  wm = 0.5f * (m * m);
  wp = 0.5f * (p * p);
}
// CIC: second cell offset (sign of d, 0 when d == 0) and weights (mesh.py:2318-2345)
__device__ __forceinline__ void cic_axis(float xp, int N, int &c, int &c2, float &w, float &w2) {
  c = (int)xp;
  float d = xp - 0.5f - (float)c;
  int s = (d > 0.0f) - (d < 0.0f);
  w2 = fabsf(d);
  w = 1.0f - w2;
  c2 = wrap(c + s, N);
}

// the three 1-D weights of one axis for cells c - 1, c, c + 1 (CIC as a 3-point stencil with one zero weight,
// mesh.py:2318-2345: the second cell is c + sign(d); NGP: the cell itself)
// morton.py:42-77: 21-bit magic-mask spread and the key of a position
__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
  x &= 0x1FFFFFull;
  x = (x | x << 32) & 0x1F00000000FFFFull;
  x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full;
  x = (x | x << 4) & 0x10C30C30C30C30C3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__device__ __forceinline__ unsigned long long morton_key(float x, float y, float z) {
  // x * 2^21 is exact in float32; floor then & 0x1FFFFF (two's complement for negatives, as numpy int64 & does)
  const long long xi = (long long)floorf(x * 2097152.0f), yi = (long long)floorf(y * 2097152.0f),
                  zi = (long long)floorf(z * 2097152.0f);
  return spread21((unsigned long long)xi) << 2 | spread21((unsigned long long)yi) << 1 |
         spread21((unsigned long long)zi);
}

template <int SCHEME>
__device__ __forceinline__ void axis_weights(float xp, int N, int &c, float &wm, float &w0, float &wp) {
  if (SCHEME == PSC_TSC) {
    tsc_axis(xp, c, wm, w0, wp);
  } else if (SCHEME == PSC_CIC) {
    c = (int)xp;
    float d = xp - 0.5f - (float)c;
    float ad = fabsf(d);
    w0 = 1.0f - ad;
    wm = d < 0.0f ? ad : 0.0f;
    wp = d > 0.0f ? ad : 0.0f;
  } else {
    c = (int)xp;
    wm = 0.0f; w0 = 1.0f; wp = 0.0f;
  }
}

// The same for the bin kernels, with CIC as a genuine 2-point scheme: the window is shifted so that the two cells with
// weight are always (c - 1, c) -- wp == 0 -- and the kernels skip every offset with a third index: 8 shared-memory
// updates / gathers per particle instead of 27.  The values are those of axis_weights<PSC_CIC>, cell for cell; c may
// be one past the particle's cell (a particle in the upper half of the last cell of a bin reaches tile index 9).
template <int SCHEME>
__device__ __forceinline__ void axis_weights_bin(float xp, int N, int &c, float &wm, float &w0, float &wp) {
  if (SCHEME == PSC_CIC) {
    const int cell = (int)xp;
    const float d = xp - 0.5f - (float)cell;
    const float ad = fabsf(d);
    const bool up = d > 0.0f;            // weight on (cell, cell + 1), else on (cell - 1, cell); d == 0: all on cell
    c = cell + (up ? 1 : 0);
    wm = up ? 1.0f - ad : ad;
    w0 = up ? ad : 1.0f - ad;
    wp = 0.0f;
  } else {
    axis_weights<SCHEME>(xp, N, c, wm, w0, wp);
  }
}
// offsets (a, e, g) in {0, 1, 2}^3 a scheme touches around (c - 1): TSC all, CIC the lower 2^3, NGP the centre
template <int SCHEME>
__device__ __forceinline__ constexpr bool stencil_uses(int a, int e, int g) {
  return SCHEME == PSC_TSC ? true : SCHEME == PSC_CIC ? (a < 2 && e < 2 && g < 2) : (a == 1 && e == 1 && g == 1);
}

}  // namespace psc
