// grid.cu -- grid algebra and finite-difference gradients.
//   utils.linear_operator[_inplace], linear_operator_vectors_inplace (utils.py:644-755)
//   mesh.derivative{2,3,5,7}, derivative*_fR_n{1,2}, add_derivative*_fR_n{1,2} (mesh.py:639-2237)
#include "common.cuh"

namespace psc {

__global__ void __launch_bounds__(256) linear_operator_kernel(const float *__restrict__ x, float f1,
                                                              float f2, float *__restrict__ out,
                                                              int64_t n) {
  int64_t n4 = n >> 2;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x = f1 * v.x + f2; v.y = f1 * v.y + f2; v.z = f1 * v.z + f2; v.w = f1 * v.w + f2;
    o4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = f1 * x[i] + f2;
  }
}

__global__ void __launch_bounds__(256) lincomb_kernel(float *__restrict__ x, float f1,
                                                      const float *__restrict__ y, float f2, int64_t n) {
  int64_t n4 = n >> 2;
  float4 *x4 = reinterpret_cast<float4 *>(x);
  const float4 *y4 = reinterpret_cast<const float4 *>(y);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x4[i], w = __ldg(&y4[i]);
    v.x = f1 * v.x + f2 * w.x; v.y = f1 * v.y + f2 * w.y; v.z = f1 * v.z + f2 * w.z; v.w = f1 * v.w + f2 * w.w;
    x4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    x[i] = f1 * x[i] + f2 * y[i];
  }
}

// --------------------------------------------------------------------------------- gradient
// One CTA = one (i, j-tile) slab of TJ rows x N columns... kept simple and HBM-friendly:
// blockDim.x walks k (coalesced), each thread produces the 3 components of one cell.  The 3-float
// AoS result of a warp (384 contiguous bytes) is staged through shared memory and written as
// coalesced 128-bit stores.
template <int FRN>
__device__ __forceinline__ float fr_val(const float *__restrict__ a, const float *__restrict__ b, float f,
                                        size_t t, bool add) {
  if (FRN == 0) return __ldg(&a[t]);
  float bv = __ldg(&b[t]);
  float pb = FRN == 1 ? bv * bv : bv * bv * bv;
  return add ? f * pb : __ldg(&a[t]) + f * pb;
}

template <int ORDER, int FRN, int STRIDE>
__global__ void __launch_bounds__(128) gradient_kernel(const float *__restrict__ a,
                                                       const float *__restrict__ b, float f, int add,
                                                       int N, float *__restrict__ force) {
  // grid: x = ceil(N/128) k-tiles, y = j, z = i
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int j = blockIdx.y, i = blockIdx.z;
  const size_t N2 = (size_t)N * N;
  __shared__ float stage[128 * 3];
  float g[3] = {0.f, 0.f, 0.f};
  const bool addb = add != 0;
  if (k < N) {
    const float pref = ORDER == 2 ? (float)N : ORDER == 3 ? (float)(0.5 * N)
                     : ORDER == 5 ? (float)(N / 12.0) : (float)(N / 60.0);
#define V(di, dj, dk) \
  fr_val<FRN>(a, b, f, (size_t)wrap(i + (di), N) * N2 + (size_t)wrap(j + (dj), N) * N + wrap(k + (dk), N), addb)
    if (ORDER == 2) {
      float c = V(0, 0, 0);
      g[0] = pref * (-c + V(1, 0, 0));
      g[1] = pref * (-c + V(0, 1, 0));
      g[2] = pref * (-c + V(0, 0, 1));
    } else if (ORDER == 3) {
      g[0] = pref * (-V(-1, 0, 0) + V(1, 0, 0));
      g[1] = pref * (-V(0, -1, 0) + V(0, 1, 0));
      g[2] = pref * (-V(0, 0, -1) + V(0, 0, 1));
    } else if (ORDER == 5) {
      g[0] = pref * (8.0f * (-V(-1, 0, 0) + V(1, 0, 0)) + V(-2, 0, 0) - V(2, 0, 0));
      g[1] = pref * (8.0f * (-V(0, -1, 0) + V(0, 1, 0)) + V(0, -2, 0) - V(0, 2, 0));
      g[2] = pref * (8.0f * (-V(0, 0, -1) + V(0, 0, 1)) + V(0, 0, -2) - V(0, 0, 2));
    } else {
      g[0] = pref * (45.0f * (-V(-1, 0, 0) + V(1, 0, 0)) + 9.0f * (V(-2, 0, 0) - V(2, 0, 0)) - V(-3, 0, 0) + V(3, 0, 0));
      g[1] = pref * (45.0f * (-V(0, -1, 0) + V(0, 1, 0)) + 9.0f * (V(0, -2, 0) - V(0, 2, 0)) - V(0, -3, 0) + V(0, 3, 0));
      g[2] = pref * (45.0f * (-V(0, 0, -1) + V(0, 0, 1)) + 9.0f * (V(0, 0, -2) - V(0, 0, 2)) - V(0, 0, -3) + V(0, 0, 3));
    }
#undef V
  }
  if (STRIDE == 4) {
    // float4-padded layout (fx, fy, fz, 0): one 16-byte store per cell, 512 contiguous bytes per warp
    if (k < N) {
      float4 *dst4 = reinterpret_cast<float4 *>(force) + ((size_t)i * N2 + (size_t)j * N + k);
      if (addb) {
        float4 o = *dst4;
        *dst4 = make_float4(o.x + g[0], o.y + g[1], o.z + g[2], 0.0f);
      } else {
        *dst4 = make_float4(g[0], g[1], g[2], 0.0f);
      }
    }
    return;
  }
  stage[3 * threadIdx.x + 0] = g[0];
  stage[3 * threadIdx.x + 1] = g[1];
  stage[3 * threadIdx.x + 2] = g[2];
  __syncthreads();
  // contiguous output segment of this CTA: 3*min(128, N - k0) floats starting at row base
  const int k0 = blockIdx.x * 128;
  const int nk = min(128, N - k0);
  float *dst = force + ((size_t)i * N2 + (size_t)j * N + k0) * 3;
  for (int t = threadIdx.x; t < 3 * nk; t += 128) {
    if (addb) dst[t] += stage[t];
    else dst[t] = stage[t];
  }
}

}  // namespace psc

using namespace psc;

extern "C" {

int psc_linear_operator(const float *x, float f1, float f2, float *out, int64_t n, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(x && out, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "pointers must be 16-byte aligned");
  linear_operator_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(x, f1, f2, out, n);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_lincomb(float *x, float f1, const float *y, float f2, int64_t n, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(x && y, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "pointers must be 16-byte aligned");
  lincomb_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(x, f1, y, f2, n);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_gradient(const float *a, const float *b, float f, int fr_n, int order, int add, int N,
                 float *force, int out_stride, void *stream) {
  PSC_CHECK_ARG(N >= 4 && N <= 32767, "N out of range");
  PSC_CHECK_ARG(order == 2 || order == 3 || order == 5 || order == 7, "gradient order must be 2, 3, 5 or 7");
  PSC_CHECK_ARG(fr_n >= 0 && fr_n <= 2, "fR_n must be 1 or 2");
  PSC_CHECK_ARG(force && (a || add) && (b || fr_n == 0), "null pointer");
  PSC_CHECK_ARG(!(add && fr_n == 0), "add requires fr_n in {1,2}");
  PSC_CHECK_ARG(N <= 65535, "N too large for grid.y/z");
  PSC_CHECK_ARG(out_stride == 3 || out_stride == 4, "out_stride must be 3 (AoS xyz) or 4 (float4-padded)");
  PSC_CHECK_ARG(out_stride == 3 || ((uintptr_t)force & 15) == 0, "float4 output must be 16-byte aligned");
  dim3 grid((N + 127) / 128, N, N);
  cudaStream_t st = as_stream(stream);
#define PSC_G(O, F)                                                              \
  do {                                                                           \
    if (out_stride == 4) gradient_kernel<O, F, 4><<<grid, 128, 0, st>>>(a, b, f, add, N, force); \
    else gradient_kernel<O, F, 3><<<grid, 128, 0, st>>>(a, b, f, add, N, force); \
  } while (0)
#define PSC_GO(O)            \
  if (fr_n == 0) PSC_G(O, 0); \
  else if (fr_n == 1) PSC_G(O, 1); \
  else PSC_G(O, 2);
  if (order == 2) { PSC_GO(2) }
  else if (order == 3) { PSC_GO(3) }
  else if (order == 5) { PSC_GO(5) }
  else { PSC_GO(7) }
#undef PSC_GO
#undef PSC_G
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
