// particles.cu -- particle-array kernels: Morton keys, reorder, leapfrog vector ops.
// HBM-bound streaming kernels: 128-bit loads/stores, grid-stride over a multiple of 148 SMs.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace psc {

// ---------------------------------------------------------------------------- Morton keys
__global__ void __launch_bounds__(256) morton_keys_kernel(const float *__restrict__ pos, int64_t np,
                                                          int64_t *__restrict__ keys) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np;
       n += (int64_t)gridDim.x * blockDim.x) {
    // x * 2^21 is exact in float32; floor then & 0x1FFFFF (two's complement for negatives, as
    // numpy int64 & does)
    keys[n] = (int64_t)morton_key(pos[3 * n + 0], pos[3 * n + 1], pos[3 * n + 2]);
  }
}

__global__ void __launch_bounds__(256) iota_kernel(int64_t *idx, int64_t np) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np;
       n += (int64_t)gridDim.x * blockDim.x)
    idx[n] = n;
}

__global__ void __launch_bounds__(256) gather3_kernel(const int64_t *__restrict__ idx,
                                                      const float *__restrict__ src,
                                                      float *__restrict__ dst, int64_t np) {
  // one thread per output float: consecutive threads write consecutive floats (coalesced stores);
  // the gathered reads are 12-byte rows, consecutive in a Morton-sorted permutation's runs.
  int64_t total = 3 * np;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = t / 3;
    int c = (int)(t - 3 * row);
    dst[t] = __ldg(&src[3 * idx[row] + c]);
  }
}

// ---------------------------------------------------------------------------- vector ops

template <bool F64>
__global__ void __launch_bounds__(256) axpy_kernel(float *__restrict__ y, const float *__restrict__ x,
                                                   double a, int64_t n) {
  const float af = (float)a;
  int64_t n4 = n >> 2;
  float4 *y4 = reinterpret_cast<float4 *>(y);
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 yv = y4[i], xv = x4[i];
    if (F64) {
      yv.x = (float)((double)yv.x + a * (double)xv.x);
      yv.y = (float)((double)yv.y + a * (double)xv.y);
      yv.z = (float)((double)yv.z + a * (double)xv.z);
      yv.w = (float)((double)yv.w + a * (double)xv.w);
    } else {
      yv.x += af * xv.x; yv.y += af * xv.y; yv.z += af * xv.z; yv.w += af * xv.w;
    }
    y4[i] = yv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    y[i] = F64 ? (float)((double)y[i] + a * (double)x[i]) : y[i] + af * x[i];
  }
}

__global__ void __launch_bounds__(256) wrap_kernel(float *__restrict__ x, int64_t n) {
  int64_t n4 = n >> 2;
  float4 *x4 = reinterpret_cast<float4 *>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x = wrap01(v.x); v.y = wrap01(v.y); v.z = wrap01(v.z); v.w = wrap01(v.w);
    x4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    x[i] = wrap01(x[i]);
  }
}

__global__ void __launch_bounds__(256) max_abs_kernel(const float *__restrict__ x, int64_t n,
                                                      float *__restrict__ out) {
  float m = 0.0f;
  int64_t n4 = n >> 2;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(&x4[i]);
    m = fmaxf(fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(x[(n4 << 2) + threadIdx.x]));
  m = warp_max(m);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) atomic_max_nonneg(out, m);
  }
}

// v -= half_dt*a ; x += dt*v ; wrap(x)   -- integration.py:250-258 in one pass (60 B/particle)
template <bool F64>
__global__ void __launch_bounds__(256) kick_drift_wrap_kernel(float *__restrict__ pos,
                                                              float *__restrict__ vel,
                                                              const float *__restrict__ acc, int64_t n,
                                                              float half_dt, double dt) {
  const float dtf = (float)dt;
  const float mh = -half_dt;
  int64_t n4 = n >> 2;
  float4 *p4 = reinterpret_cast<float4 *>(pos);
  float4 *v4 = reinterpret_cast<float4 *>(vel);
  const float4 *a4 = reinterpret_cast<const float4 *>(acc);
#define PSC_KDW(pc, vc, ac) kick_drift_wrap1<F64>(pc, vc, ac, mh, dtf, dt);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = p4[i], v = v4[i], a = __ldg(&a4[i]);
    PSC_KDW(p.x, v.x, a.x) PSC_KDW(p.y, v.y, a.y) PSC_KDW(p.z, v.z, a.z) PSC_KDW(p.w, v.w, a.w)
    p4[i] = p;
    v4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float p = pos[i], v = vel[i], a = acc[i];
    PSC_KDW(p, v, a)
    pos[i] = p;
    vel[i] = v;
  }
#undef PSC_KDW
}

// Same pass on a slab (pysco_b200/slab.py): additionally every particle whose new cell floor(x N) belongs to another
// rank's slab is counted per destination and its row appended to a list, so that the migration that follows touches
// only the leavers instead of scanning all positions twice (psc_slab_count + psc_slab_pack_leavers).
template <bool F64>
__global__ void __launch_bounds__(256) kick_drift_wrap_slab_kernel(float *__restrict__ pos, float *__restrict__ vel,
                                                                   const float *__restrict__ acc, int64_t n,
                                                                   float half_dt, double dt, float Nf, int nxl, int P,
                                                                   int me, unsigned long long *__restrict__ counts,
                                                                   int64_t *__restrict__ rows, int64_t capacity) {
  const float dtf = (float)dt;
  const float mh = -half_dt;
  int64_t n4 = n >> 2;
  float4 *p4 = reinterpret_cast<float4 *>(pos);
  float4 *v4 = reinterpret_cast<float4 *>(vel);
  const float4 *a4 = reinterpret_cast<const float4 *>(acc);
#define PSC_KDW(pc, vc, ac) kick_drift_wrap1<F64>(pc, vc, ac, mh, dtf, dt);
#define PSC_LEAVER(pc, flat)                                                        \
  if ((flat) % 3 == 0) {                                                            \
    const int d = min(max((int)(pc * Nf) / nxl, 0), P - 1);                         \
    if (d != me) {                                                                  \
      atomicAdd(&counts[d], 1ull);                                                  \
      const int64_t slot = (int64_t)atomicAdd(&counts[P], 1ull);                    \
      if (slot < capacity) rows[slot] = (flat) / 3;                                 \
    }                                                                               \
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = p4[i], v = v4[i], a = __ldg(&a4[i]);
    PSC_KDW(p.x, v.x, a.x) PSC_KDW(p.y, v.y, a.y) PSC_KDW(p.z, v.z, a.z) PSC_KDW(p.w, v.w, a.w)
    p4[i] = p;
    v4[i] = v;
    const int64_t f = i << 2;
    PSC_LEAVER(p.x, f) PSC_LEAVER(p.y, f + 1) PSC_LEAVER(p.z, f + 2) PSC_LEAVER(p.w, f + 3)
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float p = pos[i], v = vel[i], a = acc[i];
    PSC_KDW(p, v, a)
    pos[i] = p;
    vel[i] = v;
    PSC_LEAVER(p, i)
  }
#undef PSC_LEAVER
#undef PSC_KDW
}

}  // namespace psc

using namespace psc;

extern "C" {

int psc_morton_keys(const float *pos, int64_t np, int64_t *keys, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos && keys, "null pointer");
  morton_keys_kernel<<<grid_for(np, 256), 256, 0, as_stream(stream)>>>(pos, np, keys);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t psc_argsort_workspace_bytes(int64_t np) {
  if (np <= 0) return 256;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                  (const int64_t *)nullptr, (int64_t *)nullptr, np, 0, 63);
  // keys_out + iota + cub temp
  return align256(sizeof(int64_t) * np) * 2 + align256(cub_bytes) + 256;
}

int psc_argsort_keys(const int64_t *keys, int64_t np, int64_t *idx_out, void *scratch,
                     size_t scratch_bytes, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(keys && idx_out && scratch, "null pointer");
  if (scratch_bytes < psc_argsort_workspace_bytes(np)) {
    set_error("psc_argsort_keys: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  char *base = reinterpret_cast<char *>(scratch);
  uint64_t *keys_sorted = reinterpret_cast<uint64_t *>(base);
  int64_t *iota = reinterpret_cast<int64_t *>(base + align256(sizeof(int64_t) * np));
  void *cub_tmp = base + 2 * align256(sizeof(int64_t) * np);
  size_t cub_bytes = scratch_bytes - 2 * align256(sizeof(int64_t) * np);
  cudaStream_t st = as_stream(stream);
  iota_kernel<<<grid_for(np, 256), 256, 0, st>>>(iota, np);
  count_launch();
  // Morton keys are non-negative 63-bit integers: an unsigned LSD radix sort over bits [0,63) is a
  // stable sort of the signed keys.
  cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes,
                                                  reinterpret_cast<const uint64_t *>(keys), keys_sorted,
                                                  iota, idx_out, np, 0, 63, st);
  count_launch(8);
  if (e != cudaSuccess) {
    set_error("psc_argsort_keys: cub radix sort failed: %s", cudaGetErrorString(e));
    return PSC_ERR_CUDA;
  }
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_gather3(const int64_t *idx, const float *src, float *dst, int64_t np, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(idx && src && dst, "null pointer");
  PSC_CHECK_ARG(src != dst, "gather must be out of place");
  gather3_kernel<<<grid_for(3 * np, 256), 256, 0, as_stream(stream)>>>(idx, src, dst, np);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_axpy(float *y, const float *x, double a, int a_is_f64, int64_t n, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(y && x, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)y & 15) == 0 && ((uintptr_t)x & 15) == 0, "pointers must be 16-byte aligned");
  int g = grid_for((n + 3) / 4, 256);
  if (a_is_f64)
    axpy_kernel<true><<<g, 256, 0, as_stream(stream)>>>(y, x, a, n);
  else
    axpy_kernel<false><<<g, 256, 0, as_stream(stream)>>>(y, x, a, n);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_periodic_wrap(float *x, int64_t n, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(x, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)x & 15) == 0, "pointer must be 16-byte aligned");
  wrap_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(x, n);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_max_abs(const float *x, int64_t n, float *out, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(x && out, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)x & 15) == 0, "pointer must be 16-byte aligned");
  max_abs_kernel<<<grid_for((n + 3) / 4, 256, 4), 256, 0, as_stream(stream)>>>(x, n, out);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_kick_drift_wrap(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                        int dt_is_f64, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos && vel && acc, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)acc) & 15) == 0,
                "pointers must be 16-byte aligned");
  int64_t n = 3 * np;
  int g = grid_for((n + 3) / 4, 256);
  if (dt_is_f64)
    kick_drift_wrap_kernel<true><<<g, 256, 0, as_stream(stream)>>>(pos, vel, acc, n, half_dt, dt);
  else
    kick_drift_wrap_kernel<false><<<g, 256, 0, as_stream(stream)>>>(pos, vel, acc, n, half_dt, dt);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_kick_drift_wrap_slab(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                             int dt_is_f64, int N, int nxl, int P, int me, int64_t *counts, int64_t *leaver_rows,
                             int64_t capacity, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  PSC_CHECK_ARG(N >= 1 && nxl >= 1 && P >= 1 && nxl * P == N && me >= 0 && me < P, "bad slab geometry");
  PSC_CHECK_ARG(counts && (leaver_rows || capacity == 0) && capacity >= 0, "null pointer");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (P + 1), st));
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos && vel && acc, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)acc) & 15) == 0,
                "pointers must be 16-byte aligned");
  int64_t n = 3 * np;
  int g = grid_for((n + 3) / 4, 256);
  unsigned long long *c = reinterpret_cast<unsigned long long *>(counts);
  if (dt_is_f64)
    kick_drift_wrap_slab_kernel<true><<<g, 256, 0, st>>>(pos, vel, acc, n, half_dt, dt, (float)N, nxl, P, me, c,
                                                         leaver_rows, capacity);
  else
    kick_drift_wrap_slab_kernel<false><<<g, 256, 0, st>>>(pos, vel, acc, n, half_dt, dt, (float)N, nxl, P, me, c,
                                                          leaver_rows, capacity);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
