// fourier.cu -- FFT Poisson solve: cuFFT for the transforms only; hand-written kernels for the
// Green's-function multiply (+ deconvolution + normalisation), the spectral gradient and P(k).
//   fourier.fft_3D_real / ifft_3D_real / ifft_3D_real_grad   fourier.py:104-147, 251-294, 372-410
//   fourier.inverse_laplacian[_compensated|_7pt]             fourier.py:460-595
//   fourier.gradient_inverse_laplacian[_compensated]         fourier.py:606-719
//   fourier.fourier_grid_to_Pk                               fourier.py:22-100
#include <cufft.h>

#include <cmath>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace psc {

struct FftPlan {
  int N;
  cufftHandle r2c, c2r, c2r_vec3;
  bool has_vec3;
  size_t work_bytes;
  // psc_fft_poisson: N batched 2-D transforms over (y, z), one per x plane, and the W_N twiddle table of the x pass
  cufftHandle r2c_yz, c2r_yz;
  bool has_yz;
  float2 *twiddle;
};

static const char *cufft_str(cufftResult r) {
  switch (r) {
    case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
    case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
    case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
    case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
    case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
    case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
    case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
    case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
    default: return "CUFFT_ERROR";
  }
}
#define PSC_CUFFT(call)                                                     \
  do {                                                                      \
    cufftResult r__ = (call);                                               \
    if (r__ != CUFFT_SUCCESS) {                                             \
      psc::set_error("%s: %s failed: %s", __func__, #call, cufft_str(r__)); \
      return PSC_ERR_CUFFT;                                                 \
    }                                                                       \
  } while (0)

__device__ __forceinline__ float sinc_pi(float x) {
  // np.sinc(x) = sin(pi x) / (pi x)
  return x == 0.0f ? 1.0f : sinpif(x) / (3.14159265358979323846f * x);
}
__device__ __forceinline__ float inv_pow_int(float w, int e) {
  // w^(-e), e >= 0 small integer (e = 2p in {0,4,6})
  float r = 1.0f, iw = 1.0f / w;
  for (int n = 0; n < e; n++) r *= iw;
  return r;
}
__device__ __forceinline__ float kfreq(int i, int N) { return (float)(i >= (N >> 1) ? i - N : i); }

// Per-axis factors of the Green's functions, tabulated once per CTA in shared memory so that the
// per-mode work is one LDS, one add and one reciprocal (the kernel then runs at memory speed: 8 B read +
// 8 B write per mode).  tab[n] = {k_n^2 (or sin^2(pi k_n / N)), sinc(k_n / N)^(-2p)} for index n in [0, N).
template <int KIND>
__device__ __forceinline__ float2 green_axis_entry(int n, int N, int p) {
  const float h = 1.0f / (float)N;
  if (KIND == PSC_GREEN_7PT) {
    float s = sinpif(h * kfreq(n, N));
    return make_float2(s * s, 1.0f);
  }
  // plain: |k| folded as N - n above N/2 (fourier.py:475-483); compensated: signed k with n >= N/2 -> n - N
  float k = KIND == PSC_GREEN_PLAIN ? (float)(n > (N >> 1) ? N - n : n) : kfreq(n, N);
  float w = KIND == PSC_GREEN_COMPENSATED ? inv_pow_int(sinc_pi(k * h), 2 * p) : 1.0f;
  return make_float2(k * k, w);
}

// Persistent CTAs walk rows (i, j) of N/2+1 contiguous modes, four rows per iteration (four independent
// 8-byte loads in flight per thread).  The N/2 "even" modes of a row map cleanly onto the threads; the
// Nyquist mode k = N/2 of the four rows is handled by threads 0..3 afterwards.
// The spectrum is [N (kx)][nyl (ky = y0 ..)][N/2+1]: nyl = N, y0 = 0 for a whole grid, a y-block for the
// transposed layout of the slab-decomposed FFT.
template <int KIND>
__global__ void __launch_bounds__(256) green_kernel(float2 *__restrict__ spec, int N, int nyl, int y0, int p,
                                                    float scale) {
  extern __shared__ float2 gtab[];  // [N]
  for (int n = threadIdx.x; n < N; n += blockDim.x) gtab[n] = green_axis_entry<KIND>(n, N, p);
  __syncthreads();
  const int nh = N / 2, nz = nh + 1;
  const int64_t nrows = (int64_t)N * nyl;
  const float h = 1.0f / (float)N;
  // -1/(4 pi^2) for the continuous Green's functions, -(h^2/4) for the 7-point one
  const float c = (KIND == PSC_GREEN_7PT ? -(0.25f * h * h) : -0.0253302959105844f) * scale;
  // row indices fit 32 bits (N * nyl <= 4096^2): 64-bit divisions here made the kernel instruction-bound
  const unsigned nrows32 = (unsigned)nrows, unyl = (unsigned)nyl;
  for (unsigned r0 = blockIdx.x * 4u; r0 < nrows32; r0 += gridDim.x * 4u) {
    float kxy[4], wxy[4];
    float2 *row[4];
    bool dc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned r = min(r0 + u, nrows32 - 1u);
      const unsigned qi = r / unyl;
      const int j = y0 + (int)(r - qi * unyl), i = (int)qi;
      dc[u] = i == 0 && j == 0;
      const float2 tx = gtab[i], ty = gtab[j];
      kxy[u] = tx.x + ty.x;
      wxy[u] = tx.y * ty.y;
      row[u] = spec + (size_t)r * nz;
    }
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
      const float2 tz = gtab[k];
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = row[u][k];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float g = c * (wxy[u] * tz.y) / (kxy[u] + tz.x);
        if (dc[u] && k == 0) g = 0.0f;  // DC mode -> 0 (reference: x[0,0,0] = 0 after the division)
        v[u].x *= g;
        v[u].y *= g;
        if (r0 + u < nrows32) row[u][k] = v[u];
      }
    }
    if (threadIdx.x < 4 && r0 + threadIdx.x < nrows32) {
      const int u = threadIdx.x;
      const float2 tz = gtab[nh];
      const float g = c * (wxy[u] * tz.y) / (kxy[u] + tz.x);
      float2 v = row[u][nh];
      v.x *= g;
      v.y *= g;
      row[u][nh] = v;
    }
  }
}

// out3[t][d] = -i * k_d / (2 pi k^2) * W^-2p * spec[t] * scale       (8 B read, 24 B write)
__global__ void __launch_bounds__(256) grad_green_kernel(const float2 *__restrict__ spec, int N, int p,
                                                         float scale, float2 *__restrict__ out3) {
  const int nz = N / 2 + 1;
  const int64_t total = (int64_t)N * N * nz;
  const float h = 1.0f / (float)N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(t % nz);
    int64_t r = t / nz;
    int j = (int)(r % N), i = (int)(r / N);
    float kx = kfreq(i, N), ky = kfreq(j, N), kz = (float)k;
    float g = 0.159154943091895f / (kx * kx + ky * ky + kz * kz);  // 1/(2 pi)
    if (p) g *= inv_pow_int(sinc_pi(kx * h) * sinc_pi(ky * h) * sinc_pi(kz * h), 2 * p);
    g = (t == 0) ? 0.0f : g * scale;
    float2 v = __ldg(&spec[t]);
    // -i * (re + i im) = im - i re
    float tr = g * v.y, ti = -g * v.x;
    out3[3 * t + 0] = make_float2(tr * kx, ti * kx);
    out3[3 * t + 1] = make_float2(tr * ky, ti * ky);
    out3[3 * t + 2] = make_float2(tr * kz, ti * kz);
  }
}

// P(k): nearest-integer |k| bins, each stored half-spectrum mode counted once (no Hermitian weight).
// Per-CTA shared-memory bins (double), flushed with native global double atomics.
// The spectrum is [N (kx)][nyl (ky = y0 ..)][N/2+1] (nyl = N, y0 = 0: a whole grid; else a y-block of the transposed
// slab layout).
__global__ void __launch_bounds__(256) pk_kernel(float2 *__restrict__ spec, int N, int nyl, int y0, int p,
                                                 double *__restrict__ bins) {
  extern __shared__ double sb[];  // [3][N]
  for (int t = threadIdx.x; t < 3 * N; t += blockDim.x) sb[t] = 0.0;
  __syncthreads();
  const int nz = N / 2 + 1;
  const int64_t total = (int64_t)N * nyl * nz;
  const float h = 1.0f / (float)N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(t % nz);
    int64_t r = t / nz;
    int j = y0 + (int)(r % nyl), i = (int)(r / nyl);
    if (i == 0 && j == 0 && k == 0) {
      spec[t] = make_float2(0.f, 0.f);  // side effect of the reference kept (fourier.py:61)
      continue;
    }
    float kx = kfreq(i, N), ky = kfreq(j, N), kz = (float)k;
    float w = sinc_pi(kx * h) * sinc_pi(ky * h) * sinc_pi(kz * h);
    float iw = inv_pow_int(w, p);
    float2 v = spec[t];
    double re = (double)v.x * (double)iw, im = (double)v.y * (double)iw;
    float knorm = sqrtf(kx * kx + ky * ky + kz * kz);
    int bin = (int)(knorm + 0.5f);
    if (bin < N) {
      atomicAdd(&sb[bin], (double)knorm);
      atomicAdd(&sb[N + bin], re * re + im * im);
      atomicAdd(&sb[2 * N + bin], 1.0);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * N; t += blockDim.x)
    if (sb[t] != 0.0) atomicAdd(&bins[t], sb[t]);
}


// ------------------------------------------------------------------ fused x pass of the FFT Poisson solve
// solver.fft (solver.py:444-500) = rfftn -> Green's function -> irfftn: cuFFT makes three passes over the spectrum each
// way and psc_green a seventh.  psc_fft_poisson lets cuFFT do the two-dimensional (y, z) transforms of every x plane
// and does the rest -- forward transform along x, the Green / deconvolution / 1/N^3 multiply, backward transform along
// x -- in ONE pass: 5 passes over the 0.54 GB spectrum at 512^3 instead of 7.
//
// A CTA owns the x columns of one ky and 16 consecutive kz (128 contiguous bytes per x; 8 for N = 2048): N x 16 complex
// values in shared memory, column pitch N + 1 (the 16 lanes of a half warp work on 16 different columns: 16 different
// 8-byte bank pairs).  N = 8^S x {8, 4, 2}: radix-8 stages and a last radix-8 / 4 / 2 stage on contiguous blocks,
// decimation in frequency on the way forward (natural order in, digit-reversed out),
// decimation in time on the way back (digit-reversed in, natural out), so no reordering is ever needed: the Green
// multiply runs on the digit-reversed modes, in registers, between the last forward and the first backward butterfly
// of a block of 8.  The first stage reads its 8 inputs straight from global memory and the last one stores straight
// back (x = j + (N/8) m for 16 adjacent kz: 128-byte segments).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// X_q = sum_m v_m exp(-+ 2 pi i q m / R), natural order in and out; FWD: minus sign.  R = 8, 4, 2.
template <bool FWD>
__device__ __forceinline__ float2 rot90(float2 a) {   // * (-+ i)
  return FWD ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <bool FWD, int R>
__device__ __forceinline__ void dft_small(float2 (&v)[8]) {
  if (R == 2) {
    const float2 a = cadd(v[0], v[1]), b = csub(v[0], v[1]);
    v[0] = a; v[1] = b;
  } else if (R == 4) {
    const float2 b0 = cadd(v[0], v[2]), b2 = csub(v[0], v[2]), b1 = cadd(v[1], v[3]), b3 = rot90<FWD>(csub(v[1], v[3]));
    v[0] = cadd(b0, b1); v[2] = csub(b0, b1); v[1] = cadd(b2, b3); v[3] = csub(b2, b3);
  } else {
    constexpr float Q = 0.70710678118654752f;
    const float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    const float2 a1 = cadd(v[1], v[5]), t5 = csub(v[1], v[5]);
    const float2 a2 = cadd(v[2], v[6]), a6 = rot90<FWD>(csub(v[2], v[6]));
    const float2 a3 = cadd(v[3], v[7]), t7 = csub(v[3], v[7]);
    // W8^1 = (1 -+ i) / sqrt 2, W8^3 = (-1 -+ i) / sqrt 2
    const float2 a5 = FWD ? make_float2(Q * (t5.x + t5.y), Q * (t5.y - t5.x)) : make_float2(Q * (t5.x - t5.y), Q * (t5.y + t5.x));
    const float2 a7 = FWD ? make_float2(Q * (t7.y - t7.x), -Q * (t7.x + t7.y)) : make_float2(-Q * (t7.x + t7.y), Q * (t7.x - t7.y));
    // even outputs from (a0..a3), odd outputs from (a4, a5, a6, a7)
    const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = rot90<FWD>(csub(a1, a3));
    const float2 c0 = cadd(a4, a6), c2 = csub(a4, a6), c1 = cadd(a5, a7), c3 = rot90<FWD>(csub(a5, a7));
    v[0] = cadd(b0, b1); v[4] = csub(b0, b1); v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
    v[1] = cadd(c0, c1); v[5] = csub(c0, c1); v[3] = cadd(c2, c3); v[7] = csub(c2, c3);
  }
}

// N = 8^S x RM: S radix-8 stages and a last radix RM = 8, 4 or 2 on contiguous blocks
template <int N> struct XfShape {
  static constexpr int RM = (N == 64 || N == 512) ? 8 : (N == 256 || N == 2048) ? 4 : 2;
  static constexpr int S = (N / RM == 8) ? 1 : (N / RM == 64) ? 2 : 3;
  // kz per CTA (64 KB .. 131 KB of columns).  Measured at 512^3: 16 kz x 256 threads x 3 CTAs 0.40 ms, 16 x 512 x 2
  // 0.39 ms, 32 kz x 1024 threads x 1 CTA 0.46 ms; N = 2048 (8 kz): 256 threads 7.3 ms, 1024 threads 5.9 ms per 4.3 GB
  static constexpr int TK = N <= 1024 ? 16 : 8;
  // threads per CTA: the long transforms fit one or two CTAs per SM only, so the CTA itself must fill the SM
  static constexpr int NT = N <= 256 ? 256 : N == 512 ? 512 : 1024;
  static constexpr int MINB = N <= 256 ? 3 : N == 512 ? 2 : 1;   // resident CTAs per SM the register budget must allow
  static_assert(N == 64 || N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048, "unsupported N");
};
__host__ __device__ constexpr int xf_pow8(int e) { return e <= 0 ? 1 : 8 * xf_pow8(e - 1); }

// radix-8 stage s of the forward / backward transform of the TK columns in shared memory: blocks of B = N / 8^(s-1),
// sub-stride L = B / 8, twiddles W_B^(j q) = W_N^(8^(s-1) j q)
template <int N, int TK, int PITCH, int s, bool FWD>
__device__ __forceinline__ void xf_stage(float2 *col, const float2 *tw) {
  constexpr int L = N / xf_pow8(s), B = 8 * L, M = xf_pow8(s - 1), NT = XfShape<N>::NT;
  for (int u = threadIdx.x; u < N * TK / 8; u += NT) {
    const int c = u % TK, t = u / TK, blk = t / L, j = t - blk * L;
    float2 *a = col + c * PITCH + blk * B + j;
    float2 v[8];
    if (FWD) {
#pragma unroll
      for (int m = 0; m < 8; m++) v[m] = a[L * m];
      dft_small<true, 8>(v);
#pragma unroll
      for (int q = 0; q < 8; q++) a[L * q] = q ? cmul(v[q], tw[M * j * q]) : v[q];
    } else {
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = q ? cmulc(a[L * q], tw[M * j * q]) : a[0];
      dft_small<false, 8>(v);
#pragma unroll
      for (int m = 0; m < 8; m++) a[L * m] = v[m];
    }
  }
  __syncthreads();
}

// spec: element (x, row, kz) at x * xstride + row * nz + kz; row = 0 .. nrows-1 is ky = y0 + row.  Whole grid:
// xstride = N * nz, nrows = N, y0 = 0; transposed slab layout [kx][nyl][nz]: xstride = nyl * nz, nrows = nyl.
template <int KIND, int N>
__global__ void __launch_bounds__(XfShape<N>::NT, XfShape<N>::MINB) xfft_green_kernel(float2 *__restrict__ spec,
                                                                    const float2 *__restrict__ twiddle, size_t xstride,
                                                                    int y0, int p, float scale) {
  constexpr int nz = N / 2 + 1, L1 = N / 8, PITCH = N + 1, RM = XfShape<N>::RM, S = XfShape<N>::S, TK = XfShape<N>::TK;
  constexpr int NT = XfShape<N>::NT;
  extern __shared__ float2 xs[];          // [TK][PITCH] columns, then the twiddles [N], then the Green table [N]
  float2 *col = xs, *tw = xs + TK * PITCH, *gtab = tw + N;
  const int ky = y0 + blockIdx.y, kz0 = blockIdx.x * TK;
  for (int n = threadIdx.x; n < N; n += NT) {
    tw[n] = twiddle[n];
    gtab[n] = green_axis_entry<KIND>(n, N, p);
  }
  const float h = 1.0f / (float)N;
  const float cst = (KIND == PSC_GREEN_7PT ? -(0.25f * h * h) : -0.0253302959105844f) * scale;
  float2 *g0 = spec + (size_t)blockIdx.y * nz + kz0;          // + x * xstride + c
  __syncthreads();
  // ---- forward stage 1: x = j + L1 m from global memory, out at j + L1 q, times W_N^(j q)
  for (int u = threadIdx.x; u < N * TK / 8; u += NT) {
    const int c = u % TK, j = u / TK;
    float2 v[8];
    if (kz0 + c < nz) {
#pragma unroll
      for (int m = 0; m < 8; m++) v[m] = g0[(size_t)(j + L1 * m) * xstride + c];
    } else {
#pragma unroll
      for (int m = 0; m < 8; m++) v[m] = make_float2(0.0f, 0.0f);
    }
    dft_small<true, 8>(v);
#pragma unroll
    for (int q = 0; q < 8; q++) col[c * PITCH + j + L1 * q] = q ? cmul(v[q], tw[j * q]) : v[q];
  }
  __syncthreads();
  if constexpr (S >= 2) xf_stage<N, TK, PITCH, 2, true>(col, tw);
  if constexpr (S >= 3) xf_stage<N, TK, PITCH, 3, true>(col, tw);
  // ---- last forward stage, Green's function, first backward stage: blocks of RM in registers.  The block index read
  // backwards in base 8 gives the low digits of kx, the position within the block the highest one.
  for (int u = threadIdx.x; u < N * TK / RM; u += NT) {
    const int c = u % TK, blk = u / TK;
    float2 *a = col + c * PITCH + RM * blk;
    float2 v[8];
#pragma unroll
    for (int m = 0; m < RM; m++) v[m] = a[m];
    dft_small<true, RM>(v);
    const int klow = S == 1 ? blk : S == 2 ? (blk >> 3) + 8 * (blk & 7) : (blk >> 6) + 8 * ((blk >> 3) & 7) + 64 * (blk & 7);
    const int kz = min(kz0 + c, nz - 1);
    const float2 ty = gtab[ky], tz = gtab[kz];
    const float k2 = ty.x + tz.x, w = ty.y * tz.y;
#pragma unroll
    for (int q = 0; q < RM; q++) {
      const int kx = klow + xf_pow8(S) * q;
      const float2 tx = gtab[kx];
      float g = cst * (w * tx.y) / (k2 + tx.x);
      if (kx == 0 && ky == 0 && kz == 0) g = 0.0f;    // DC mode -> 0 (reference: x[0,0,0] = 0 after the division)
      v[q].x *= g;
      v[q].y *= g;
    }
    dft_small<false, RM>(v);
#pragma unroll
    for (int m = 0; m < RM; m++) a[m] = v[m];
  }
  __syncthreads();
  if constexpr (S >= 3) xf_stage<N, TK, PITCH, 3, false>(col, tw);
  if constexpr (S >= 2) xf_stage<N, TK, PITCH, 2, false>(col, tw);
  // ---- backward stage 1, straight to global memory
  for (int u = threadIdx.x; u < N * TK / 8; u += NT) {
    const int c = u % TK, j = u / TK;
    if (kz0 + c >= nz) continue;
    float2 v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const float2 e = col[c * PITCH + j + L1 * q];
      v[q] = q ? cmulc(e, tw[j * q]) : e;
    }
    dft_small<false, 8>(v);
#pragma unroll
    for (int m = 0; m < 8; m++) g0[(size_t)(j + L1 * m) * xstride + c] = v[m];
  }
}

// W_N^k = exp(-2 pi i k / N), k < N, built once per (device, N) from double precision
static float2 *xf_twiddles(int N) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, float2 *> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({dev, N});
  if (it != cache.end()) return it->second;
  std::vector<float2> tw(N);
  for (int k = 0; k < N; k++) {
    const double a = -2.0 * 3.14159265358979323846 * (double)k / (double)N;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  float2 *d = nullptr;
  if (cudaMalloc(&d, sizeof(float2) * N) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  cache[{dev, N}] = d;
  return d;
}

template <int KIND, int N>
static int xfft_green_launch(float2 *spec, size_t xstride, int nrows, int y0, int p, float scale, cudaStream_t st) {
  constexpr int TK = XfShape<N>::TK, nz = N / 2 + 1;
  const size_t smem = sizeof(float2) * (TK * (N + 1) + 2 * N);
  static bool attr = false;
  if (!attr) {
    PSC_CUDA(cudaFuncSetAttribute(xfft_green_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  float2 *tw = xf_twiddles(N);
  if (!tw) {
    set_error("xfft_green: could not allocate the twiddle table");
    return PSC_ERR_CUDA;
  }
  const dim3 grid((nz + TK - 1) / TK, nrows);
  xfft_green_kernel<KIND, N><<<grid, XfShape<N>::NT, smem, st>>>(spec, tw, xstride, y0, p, scale);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

template <int KIND>
static int xfft_green_n(float2 *spec, int N, size_t xstride, int nrows, int y0, int p, float scale, cudaStream_t st) {
  switch (N) {
    case 64: return xfft_green_launch<KIND, 64>(spec, xstride, nrows, y0, p, scale, st);
    case 128: return xfft_green_launch<KIND, 128>(spec, xstride, nrows, y0, p, scale, st);
    case 256: return xfft_green_launch<KIND, 256>(spec, xstride, nrows, y0, p, scale, st);
    case 512: return xfft_green_launch<KIND, 512>(spec, xstride, nrows, y0, p, scale, st);
    case 1024: return xfft_green_launch<KIND, 1024>(spec, xstride, nrows, y0, p, scale, st);
    case 2048: return xfft_green_launch<KIND, 2048>(spec, xstride, nrows, y0, p, scale, st);
  }
  set_error("xfft_green: N must be a power of two in [64, 2048]");
  return PSC_ERR_INVALID;
}

static int xfft_green(float2 *spec, int N, size_t xstride, int nrows, int y0, int kind, int p, float scale,
                      cudaStream_t st) {
  if (kind == PSC_GREEN_PLAIN) return xfft_green_n<PSC_GREEN_PLAIN>(spec, N, xstride, nrows, y0, p, scale, st);
  if (kind == PSC_GREEN_COMPENSATED) return xfft_green_n<PSC_GREEN_COMPENSATED>(spec, N, xstride, nrows, y0, p, scale, st);
  return xfft_green_n<PSC_GREEN_7PT>(spec, N, xstride, nrows, y0, p, scale, st);
}

}  // namespace psc

using namespace psc;

extern "C" {

int psc_fft_plan_create(int N, void **plan_out) {
  PSC_CHECK_ARG(plan_out, "null plan_out");
  PSC_CHECK_ARG(N >= 2 && N <= 4096, "N out of range");
  FftPlan *pl = new FftPlan();
  pl->N = N;
  pl->has_vec3 = false;
  pl->has_yz = false;
  pl->twiddle = nullptr;
  size_t w1 = 0, w2 = 0;
  PSC_CUFFT(cufftCreate(&pl->r2c));
  PSC_CUFFT(cufftMakePlan3d(pl->r2c, N, N, N, CUFFT_R2C, &w1));
  PSC_CUFFT(cufftCreate(&pl->c2r));
  PSC_CUFFT(cufftMakePlan3d(pl->c2r, N, N, N, CUFFT_C2R, &w2));
  pl->work_bytes = w1 > w2 ? w1 : w2;
  *plan_out = pl;
  return PSC_OK;
}

int psc_fft_plan_destroy(void *plan) {
  if (!plan) return PSC_OK;
  FftPlan *pl = reinterpret_cast<FftPlan *>(plan);
  cufftDestroy(pl->r2c);
  cufftDestroy(pl->c2r);
  if (pl->has_vec3) cufftDestroy(pl->c2r_vec3);
  if (pl->has_yz) {
    cufftDestroy(pl->r2c_yz);
    cufftDestroy(pl->c2r_yz);
  }
  delete pl;
  return PSC_OK;
}

size_t psc_fft_plan_workspace_bytes(void *plan) {
  return plan ? reinterpret_cast<FftPlan *>(plan)->work_bytes : 0;
}

int psc_fft_r2c(void *plan, const float *in, float *spec_out, void *stream) {
  PSC_CHECK_ARG(plan && in && spec_out, "null pointer");
  FftPlan *pl = reinterpret_cast<FftPlan *>(plan);
  PSC_CUFFT(cufftSetStream(pl->r2c, as_stream(stream)));
  PSC_CUFFT(cufftExecR2C(pl->r2c, const_cast<float *>(in), reinterpret_cast<cufftComplex *>(spec_out)));
  count_launch(3);
  return PSC_OK;
}

int psc_fft_c2r(void *plan, float *spec_in, float *out, void *stream) {
  PSC_CHECK_ARG(plan && spec_in && out, "null pointer");
  FftPlan *pl = reinterpret_cast<FftPlan *>(plan);
  PSC_CUFFT(cufftSetStream(pl->c2r, as_stream(stream)));
  PSC_CUFFT(cufftExecC2R(pl->c2r, reinterpret_cast<cufftComplex *>(spec_in), out));
  count_launch(3);
  return PSC_OK;
}

int psc_fft_poisson_supported(int N) { return N >= 64 && N <= 2048 && (N & (N - 1)) == 0; }

/* solver.fft (solver.py:444-500) in one call: rhs -> rfftn -> Green's function x W^-2p x scale -> irfftn -> out.
 * cuFFT does the batched 2-D (y, z) transforms of the x planes; the transforms along x and the Green multiply are one
 * kernel (xfft_green_kernel).  spec: [N, N, N/2+1] complex64 scratch; out may alias rhs.  N: power of two, 64..2048. */
int psc_fft_poisson(void *plan, const float *rhs, float *spec, float *out, int kind, int p, float scale, void *stream) {
  PSC_CHECK_ARG(plan && rhs && spec && out, "null pointer");
  FftPlan *pl = reinterpret_cast<FftPlan *>(plan);
  const int N = pl->N;
  PSC_CHECK_ARG(psc_fft_poisson_supported(N), "psc_fft_poisson: N must be a power of two in [64, 2048]");
  PSC_CHECK_ARG(kind >= PSC_GREEN_PLAIN && kind <= PSC_GREEN_7PT, "unknown Green's function");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  cudaStream_t st = as_stream(stream);
  const int nz = N / 2 + 1;
  if (!pl->has_yz) {
    int n[2] = {N, N};
    size_t w = 0;
    PSC_CUFFT(cufftCreate(&pl->r2c_yz));
    PSC_CUFFT(cufftMakePlanMany(pl->r2c_yz, 2, n, nullptr, 1, N * N, nullptr, 1, N * nz, CUFFT_R2C, N, &w));
    PSC_CUFFT(cufftCreate(&pl->c2r_yz));
    PSC_CUFFT(cufftMakePlanMany(pl->c2r_yz, 2, n, nullptr, 1, N * nz, nullptr, 1, N * N, CUFFT_C2R, N, &w));
    pl->has_yz = true;
  }
  PSC_CUFFT(cufftSetStream(pl->r2c_yz, st));
  PSC_CUFFT(cufftExecR2C(pl->r2c_yz, const_cast<float *>(rhs), reinterpret_cast<cufftComplex *>(spec)));
  int rc = xfft_green(reinterpret_cast<float2 *>(spec), N, (size_t)N * nz, N, 0, kind, p, scale, st);
  if (rc != PSC_OK) return rc;
  PSC_CUFFT(cufftSetStream(pl->c2r_yz, st));
  PSC_CUFFT(cufftExecC2R(pl->c2r_yz, reinterpret_cast<cufftComplex *>(spec), out));
  count_launch(4);
  return PSC_OK;
}

/* The x part of the slab-decomposed FFT solve (psc_slab_fft_x forward, psc_green_slab, psc_slab_fft_x backward) in one
 * kernel, in place on the transposed spectrum [N (kx)][nyl (ky = y0 ..)][N/2+1]. */
int psc_xfft_green_slab(float *spec_t, int N, int nyl, int y0, int kind, int p, float scale, void *stream) {
  PSC_CHECK_ARG(spec_t, "null pointer");
  PSC_CHECK_ARG(psc_fft_poisson_supported(N), "psc_xfft_green_slab: N must be a power of two in [64, 2048]");
  PSC_CHECK_ARG(nyl >= 1 && y0 >= 0 && y0 + nyl <= N, "bad y block");
  PSC_CHECK_ARG(kind >= PSC_GREEN_PLAIN && kind <= PSC_GREEN_7PT, "unknown Green's function");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  return xfft_green(reinterpret_cast<float2 *>(spec_t), N, (size_t)nyl * (N / 2 + 1), nyl, y0, kind, p, scale,
                    as_stream(stream));
}

int psc_fft_c2r_vec3(void *plan, float *spec3_in, float *out3, void *stream) {
  PSC_CHECK_ARG(plan && spec3_in && out3, "null pointer");
  FftPlan *pl = reinterpret_cast<FftPlan *>(plan);
  if (!pl->has_vec3) {
    int N = pl->N;
    int n[3] = {N, N, N};
    int inembed[3] = {N, N, N / 2 + 1};
    int onembed[3] = {N, N, N};
    size_t w = 0;
    PSC_CUFFT(cufftCreate(&pl->c2r_vec3));
    PSC_CUFFT(cufftMakePlanMany(pl->c2r_vec3, 3, n, inembed, 3, 1, onembed, 3, 1, CUFFT_C2R, 3, &w));
    pl->has_vec3 = true;
  }
  PSC_CUFFT(cufftSetStream(pl->c2r_vec3, as_stream(stream)));
  PSC_CUFFT(cufftExecC2R(pl->c2r_vec3, reinterpret_cast<cufftComplex *>(spec3_in), out3));
  count_launch(9);
  return PSC_OK;
}

int psc_green(float *spec, int N, int kind, int p, float scale, void *stream) {
  PSC_CHECK_ARG(spec, "null pointer");
  PSC_CHECK_ARG(N >= 2, "N out of range");
  PSC_CHECK_ARG(kind >= PSC_GREEN_PLAIN && kind <= PSC_GREEN_7PT, "unknown Green's function");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  PSC_CHECK_ARG(N <= 4096, "N out of range");
  return psc_green_slab(spec, N, N, 0, kind, p, scale, stream);
}

int psc_green_slab(float *spec_t, int N, int nyl, int y0, int kind, int p, float scale, void *stream) {
  PSC_CHECK_ARG(spec_t, "null pointer");
  PSC_CHECK_ARG(N >= 2 && N <= 4096, "N out of range");
  PSC_CHECK_ARG(nyl >= 1 && y0 >= 0 && y0 + nyl <= N, "bad y block");
  PSC_CHECK_ARG(kind >= PSC_GREEN_PLAIN && kind <= PSC_GREEN_7PT, "unknown Green's function");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  int64_t nrows = (int64_t)N * nyl;
  int g = (int)((nrows + 3) / 4 < (int64_t)num_sms() * 8 ? (nrows + 3) / 4 : (int64_t)num_sms() * 8);
  float2 *s = reinterpret_cast<float2 *>(spec_t);
  cudaStream_t st = as_stream(stream);
  const size_t smem = sizeof(float2) * N;
  if (kind == PSC_GREEN_PLAIN) green_kernel<PSC_GREEN_PLAIN><<<g, 256, smem, st>>>(s, N, nyl, y0, p, scale);
  else if (kind == PSC_GREEN_COMPENSATED) green_kernel<PSC_GREEN_COMPENSATED><<<g, 256, smem, st>>>(s, N, nyl, y0, p, scale);
  else green_kernel<PSC_GREEN_7PT><<<g, 256, smem, st>>>(s, N, nyl, y0, p, scale);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_grad_green(const float *spec, int N, int p, float scale, float *out3, void *stream) {
  PSC_CHECK_ARG(spec && out3, "null pointer");
  PSC_CHECK_ARG(N >= 2, "N out of range");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  int64_t total = (int64_t)N * N * (N / 2 + 1);
  grad_green_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2 *>(spec), N, p, scale, reinterpret_cast<float2 *>(out3));
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_pk(float *spec, int N, int p, double *bins, void *stream) {
  return psc_pk_slab(spec, N, N, 0, p, bins, stream);
}

int psc_pk_slab(float *spec, int N, int nyl, int y0, int p, double *bins, void *stream) {
  PSC_CHECK_ARG(spec && bins, "null pointer");
  PSC_CHECK_ARG(N >= 2 && N <= 2048, "N out of range");
  PSC_CHECK_ARG(nyl >= 1 && y0 >= 0 && y0 + nyl <= N, "bad y block");
  PSC_CHECK_ARG(p >= 0 && p <= 8, "MAS index out of range");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(bins, 0, sizeof(double) * 3 * N, st));
  int64_t total = (int64_t)N * nyl * (N / 2 + 1);
  size_t smem = sizeof(double) * 3 * N;
  if (smem > 48 * 1024)
    PSC_CUDA(cudaFuncSetAttribute(pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pk_kernel<<<grid_for(total, 256, 2), 256, smem, st>>>(reinterpret_cast<float2 *>(spec), N, nyl, y0, p, bins);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
