// multigrid.cu -- 7-point-stencil kernels of the multigrid solvers.
//   laplacian.py: operator :12, residual :63, restrict_residual :125, residual_error :327,
//                 initialise_potential :765, gauss_seidel :844
//   cubic.py / quartic.py: operator, residual_with_rhs, initialise_potential, gauss_seidel[_with_rhs],
//                 residual_error, solution_cubic_equation (cubic.py:162-207, float64),
//                 solution_quartic_equation (quartic.py:157-204, float64)
//   mesh.py: restriction :14, minus_restriction :62, prolongation :180, add_prolongation :334
//   mond.py: rhs_simple/n/beta/gamma/delta :171-932
//
// Thread mapping: threadIdx.x walks k (fastest axis, coalesced 128-byte rows), blockIdx.y/z carry
// (j, i) tiles; neighbour rows come through L1/L2 (each row is re-read by its 4 lateral neighbours
// of the same CTA).  Reductions: warp shuffle -> one double atomic per CTA.
#include <cstdlib>

#include "common.cuh"
#include "fr_roots.cuh"

namespace psc {

constexpr int TJ = 4;  // rows (j) per CTA
constexpr int TKX = 64;  // k-threads per CTA

__device__ __forceinline__ float npow(float v, int kind) {
  return kind == PSC_OP_CUBIC ? v * v : v * v * v;
}

struct Cell {
  int i, j, k;
  size_t t;
  bool ok;
};

__device__ __forceinline__ Cell this_cell(int N) {
  Cell c;
  c.k = blockIdx.x * TKX + threadIdx.x;
  c.j = blockIdx.y * TJ + threadIdx.y;
  c.i = blockIdx.z;
  c.ok = c.k < N && c.j < N;
  c.t = ((size_t)c.i * N + c.j) * N + c.k;
  return c;
}

// sum of the six neighbours (optionally of their squares / cubes)
template <int KIND>
__device__ __forceinline__ float nb6(const float *__restrict__ x, int i, int j, int k, int N) {
  const size_t N2 = (size_t)N * N;
  const size_t ri = (size_t)i * N2, rj = (size_t)j * N;
  float a = x[(size_t)wrap(i - 1, N) * N2 + rj + k];
  float b = x[ri + (size_t)wrap(j - 1, N) * N + k];
  float c = x[ri + rj + wrap(k - 1, N)];
  float d = x[ri + rj + wrap(k + 1, N)];
  float e = x[ri + (size_t)wrap(j + 1, N) * N + k];
  float f = x[(size_t)wrap(i + 1, N) * N2 + rj + k];
  if (KIND == PSC_OP_LAPLACIAN) return a + b + c + d + e + f;
  return npow(a, KIND) + npow(b, KIND) + npow(c, KIND) + npow(d, KIND) + npow(e, KIND) + npow(f, KIND);
}

__global__ void __launch_bounds__(256) diff_sumsq_kernel(const float *__restrict__ a, float fa,
                                                         const float *__restrict__ b, int64_t n,
                                                         double *__restrict__ out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float d = fa * a[i] - b[i];
    s += (double)d * (double)d;
  }
  s = warp_sum(s);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) tot += sm[w];
    atomicAdd(out, tot);
  }
}

// coarse[i,j,k] = 1/8 sum_{2^3 children} (b - Lx): thread per coarse cell, children evaluated in place
__global__ void __launch_bounds__(TKX *TJ) restrict_residual_kernel(const float *__restrict__ x,
                                                                    const float *__restrict__ b, int N,
                                                                    float *__restrict__ coarse) {
  const int Nc = N >> 1;
  Cell c = this_cell(Nc);
  if (!c.ok) return;
  const float invh2 = (float)N * (float)N;
  const size_t N2 = (size_t)N * N;
  float outer = 0.0f, inner = 0.0f, bs = 0.0f;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int g = 0; g < 2; g++) {
        int ii = 2 * c.i + a, jj = 2 * c.j + e, kk = 2 * c.k + g;
        size_t t = (size_t)ii * N2 + (size_t)jj * N + kk;
        outer += x[(size_t)wrap(ii + (a ? 1 : -1), N) * N2 + (size_t)jj * N + kk] +
                 x[(size_t)ii * N2 + (size_t)wrap(jj + (e ? 1 : -1), N) * N + kk] +
                 x[(size_t)ii * N2 + (size_t)jj * N + wrap(kk + (g ? 1 : -1), N)];
        inner += x[t];
        bs += b[t];
      }
  coarse[c.t] = 0.125f * (-(outer - 3.0f * inner) * invh2 + bs);
}

__global__ void __launch_bounds__(TKX *TJ) restriction_kernel(const float *__restrict__ x, int N, float f,
                                                              float *__restrict__ coarse) {
  const int Nc = N >> 1;
  Cell c = this_cell(Nc);
  if (!c.ok) return;
  const size_t N2 = (size_t)N * N;
  const float *p = x + (size_t)(2 * c.i) * N2 + (size_t)(2 * c.j) * N + 2 * c.k;
  float2 r00 = *reinterpret_cast<const float2 *>(p);
  float2 r01 = *reinterpret_cast<const float2 *>(p + N);
  float2 r10 = *reinterpret_cast<const float2 *>(p + N2);
  float2 r11 = *reinterpret_cast<const float2 *>(p + N2 + N);
  coarse[c.t] = f * (r00.x + r00.y + r01.x + r01.y + r10.x + r10.y + r11.x + r11.y);
}

// thread per COARSE cell: reads its 27-neighbourhood once, writes (or adds to) its 8 children
template <bool ADD>
__global__ void __launch_bounds__(TKX *TJ) prolongation_kernel(float *__restrict__ fine,
                                                               const float *__restrict__ coarse, int Nc) {
  Cell c = this_cell(Nc);
  if (!c.ok) return;
  const int N = 2 * Nc;
  const size_t Nc2 = (size_t)Nc * Nc, N2 = (size_t)N * N;
  const float f0 = 27.0f / 64, f1 = 9.0f / 64, f2 = 3.0f / 64, f3 = 1.0f / 64;
  float v[3][3][3];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int e = 0; e < 3; e++)
#pragma unroll
      for (int g = 0; g < 3; g++)
        v[a][e][g] = coarse[(size_t)wrap(c.i + a - 1, Nc) * Nc2 + (size_t)wrap(c.j + e - 1, Nc) * Nc +
                            wrap(c.k + g - 1, Nc)];
  const float t0 = f0 * v[1][1][1];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      float r[2];
#pragma unroll
      for (int g = 0; g < 2; g++) {
        int A = 2 * a, E = 2 * e, G = 2 * g;
        r[g] = t0 + f1 * (v[A][1][1] + v[1][E][1] + v[1][1][G]) +
               f2 * (v[A][E][1] + v[1][E][G] + v[A][1][G]) + f3 * v[A][E][G];
      }
      float2 *dst = reinterpret_cast<float2 *>(fine + (size_t)(2 * c.i + a) * N2 +
                                               (size_t)(2 * c.j + e) * N + 2 * c.k);
      if (ADD) {
        float2 old = *dst;
        *dst = make_float2(old.x + r[0], old.y + r[1]);
      } else {
        *dst = make_float2(r[0], r[1]);
      }
    }
}

// f(R) root solvers: fr_roots.cuh (float32 fast path + the reference's float64 statement)

template <int KIND>
__global__ void __launch_bounds__(256) init_potential_kernel(const float *__restrict__ b, float q_val,
                                                             const float *__restrict__ q_dev, int N,
                                                             float *__restrict__ out, int64_t n) {
  const float q = q_dev ? *q_dev : q_val;  // device-resident q: lets a captured CUDA graph follow q from step to step
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x) {
    if (KIND == PSC_OP_LAPLACIAN) {
      const float h = 1.0f / (float)N;
      const float c = (float)(-(double)h * (double)h / 6.0);
      out[t] = c * b[t];
    } else if (KIND == PSC_OP_CUBIC) {
      // cubic.py:247-258
      const float h2 = 1.0f / ((float)N * (float)N);
      const float threeh2 = 3.0f * h2;
      const double d1 = 27.0 * (double)h2 * (double)q;
      float d0 = -threeh2 * b[t];
      double d03 = (double)d0 * (double)d0 * (double)d0;
      double C = cbrt(0.5 * (d1 + sqrt(d1 * d1 - 4.0 * d03)));
      out[t] = (float)(-(1.0 / 3) * (C + (double)d0 / C));
    } else {
      // quartic.py:241-259
      const double h2 = 1.0 / ((double)N * (double)N);
      const double inv3 = 1.0 / 3;
      const double d0 = 12.0 * h2 * (double)q;
      double p = h2 * (double)b[t];
      double d1 = 27.0 * (p * p);
      double Q = pow(0.5 * (d1 + sqrt(d1 * d1 - 4.0 * (d0 * d0 * d0))), inv3);
      double S = 0.5 * sqrt((Q + d0 / Q) * inv3);
      out[t] = (float)(-S + 0.5 * sqrt(-4.0 * (S * S) + p / S));
    }
  }
}

// One colour of a red-black SOR sweep.  colour = 1: odd i+j+k ("red" in the reference, first),
// colour = 0: even ("black").  Thread per updated cell: threadIdx.x walks k in steps of 2.
template <int KIND>
__global__ void __launch_bounds__(TKX *TJ) gs_colour_kernel(float *__restrict__ x,
                                                            const float *__restrict__ b, float q_val,
                                                            const float *__restrict__ q_dev,
                                                            const float *__restrict__ rhs, int N,
                                                            float f_relax, int colour) {
  const float q = q_dev ? *q_dev : q_val;
  const int kh = blockIdx.x * TKX + threadIdx.x;
  const int j = blockIdx.y * TJ + threadIdx.y;
  const int i = blockIdx.z;
  const int k = 2 * kh + ((i + j + colour) & 1);
  if (k >= N || j >= N) return;
  const size_t t = ((size_t)i * N + j) * N + k;
  const float h2 = 1.0f / ((float)N * (float)N);
  const float invsix = 1.0f / 6.0f;
  float xt = x[t];
  float s = nb6<KIND>(x, i, j, k, N);
  float target;
  if (KIND == PSC_OP_LAPLACIAN) {
    target = (s - h2 * b[t]) * invsix;
  } else {
    float p = h2 * b[t] - invsix * s;
    if (KIND == PSC_OP_CUBIC) {
      float d1 = 27.0f * h2 * q;
      if (rhs) d1 -= 27.0f * rhs[t];
      target = solve_cubic(p, d1);
    } else {
      float qq = q * h2;
      if (rhs) qq -= rhs[t];
      target = solve_quartic(p, qq);
    }
  }
  x[t] = xt + f_relax * (target - xt);
}

// ----------------------------------------------------------------------------------- MOND
template <int FN>
__device__ __forceinline__ float mond_nu(float y, float alpha) {
  if (FN == PSC_MOND_SIMPLE) return 0.5f + sqrtf(0.25f + 1.0f / y);
  if (FN == PSC_MOND_N) {
    int n = (int)alpha;
    return powf(0.5f + sqrtf(0.25f + powf(y, (float)(-n))), 1.0f / (float)n);
  }
  if (FN == PSC_MOND_BETA) {
    float e = expf(-y);
    float nu = alpha * e;
    float om = 1.0f - e;
    if (om > 0.0f) nu += rsqrtf(om);
    return nu;
  }
  if (FN == PSC_MOND_GAMMA) {
    float e = expf(-powf(y, 0.5f * alpha));
    return powf(1.0f - e, -1.0f / alpha) + (1.0f - 1.0f / alpha) * e;
  }
  return powf(1.0f - expf(-powf(y, 0.5f * alpha)), -1.0f / alpha);
}

template <int FN>
__global__ void __launch_bounds__(TKX *TJ) mond_rhs_kernel(const float *__restrict__ phi,
                                                           float *__restrict__ out, int N, float g0,
                                                           float alpha) {
  Cell c = this_cell(N);
  if (!c.ok) return;
  const size_t N2 = (size_t)N * N;
  const float inv_g0 = 1.0f / g0;
  const float invh = (float)N, inv4h = 0.25f * (float)N;
  const int i = c.i, j = c.j, k = c.k;
  const size_t ri[3] = {(size_t)wrap(i - 1, N) * N2, (size_t)i * N2, (size_t)wrap(i + 1, N) * N2};
  const size_t rj[3] = {(size_t)wrap(j - 1, N) * N, (size_t)j * N, (size_t)wrap(j + 1, N) * N};
  const int rk[3] = {wrap(k - 1, N), k, wrap(k + 1, N)};
#define P(a, e, g) phi[ri[(a) + 1] + rj[(e) + 1] + rk[(g) + 1]]
  float p0 = P(0, 0, 0);
  // Point A at -h/2, point B at +h/2 along each axis (mond.py:209-300)
  float Axx = invh * (p0 - P(-1, 0, 0));
  float Axy = inv4h * (P(0, 1, 0) - P(0, -1, 0) + P(-1, 1, 0) - P(-1, -1, 0));
  float Axz = inv4h * (P(0, 0, 1) - P(0, 0, -1) + P(-1, 0, 1) - P(-1, 0, -1));
  float fAx = sqrtf(Axx * Axx + Axy * Axy + Axz * Axz);
  float Bxx = invh * (-p0 + P(1, 0, 0));
  float Bxy = inv4h * (P(1, 1, 0) - P(1, -1, 0) + P(0, 1, 0) - P(0, -1, 0));
  float Bxz = inv4h * (P(1, 0, 1) - P(1, 0, -1) + P(0, 0, 1) - P(0, 0, -1));
  float fBx = sqrtf(Bxx * Bxx + Bxy * Bxy + Bxz * Bxz);
  float Ayy = invh * (p0 - P(0, -1, 0));
  float Ayx = inv4h * (P(1, 0, 0) - P(-1, 0, 0) + P(1, -1, 0) - P(-1, -1, 0));
  float Ayz = inv4h * (P(0, 0, 1) - P(0, 0, -1) + P(0, -1, 1) - P(0, -1, -1));
  float fAy = sqrtf(Ayx * Ayx + Ayy * Ayy + Ayz * Ayz);
  float Byy = invh * (-p0 + P(0, 1, 0));
  float Byx = inv4h * (P(1, 1, 0) - P(-1, 1, 0) + P(1, 0, 0) - P(-1, 0, 0));
  float Byz = inv4h * (P(0, 1, 1) - P(0, 1, -1) + P(0, 0, 1) - P(0, 0, -1));
  float fBy = sqrtf(Byx * Byx + Byy * Byy + Byz * Byz);
  float Azz = invh * (p0 - P(0, 0, -1));
  float Azx = inv4h * (P(1, 0, 0) - P(-1, 0, 0) + P(1, 0, -1) - P(-1, 0, -1));
  float Azy = inv4h * (P(0, 1, 0) - P(0, -1, 0) + P(0, 1, -1) - P(0, -1, -1));
  float fAz = sqrtf(Azx * Azx + Azy * Azy + Azz * Azz);
  float Bzz = invh * (-p0 + P(0, 0, 1));
  float Bzx = inv4h * (P(1, 0, 1) - P(-1, 0, 1) + P(1, 0, 0) - P(-1, 0, 0));
  float Bzy = inv4h * (P(0, 1, 1) - P(0, -1, 1) + P(0, 1, 0) - P(0, -1, 0));
  float fBz = sqrtf(Bzx * Bzx + Bzy * Bzy + Bzz * Bzz);
#undef P
  float r = mond_nu<FN>(fBx * inv_g0, alpha) * Bxx - mond_nu<FN>(fAx * inv_g0, alpha) * Axx +
            mond_nu<FN>(fBy * inv_g0, alpha) * Byy - mond_nu<FN>(fAy * inv_g0, alpha) * Ayy +
            mond_nu<FN>(fBz * inv_g0, alpha) * Bzz - mond_nu<FN>(fAz * inv_g0, alpha) * Azz;
  out[c.t] = invh * r;
}

// The same right-hand side with every face evaluated ONCE: the B face of a cell along an axis is the A face of its
// upper neighbour (same ten potentials, same operations in the same order), so the cell-wise kernel above computes
// every nu(|grad phi| / g0) * d phi twice -- six faces of 10 loads, a square root and an interpolating function each
// (2.1 ms at 512^3, instruction-bound).  Here a CTA owns a 64 (k) x 8 (j) column and marches over MF_CH planes in i:
// a thread computes the three A-face fluxes of its cell, the y / z fluxes of the upper neighbours come from shared
// memory (the threads of the last row / column also compute the face just outside the tile), the x flux of the next
// plane from the next step of the march.
constexpr int MF_K = 64, MF_J = 8, MF_CH = 16;

template <int FN, int AXIS>
__device__ __forceinline__ float mond_face_flux(const float *__restrict__ phi, size_t rim, size_t ri0, size_t rip,
                                                size_t rjm, size_t rj0, size_t rjp, int km, int k0, int kp,
                                                float invh, float inv4h, float inv_g0, float alpha) {
  // A face (at -h/2 along AXIS) of the cell (ri0, rj0, k0): mond.py:209-300
#define PH(a, e, g) phi[(a) + (e) + (g)]
  const float p0 = PH(ri0, rj0, k0);
  float d0, d1, d2;   // derivative along AXIS, then the two transverse ones in the reference's order of summation
  if (AXIS == 0) {
    d0 = invh * (p0 - PH(rim, rj0, k0));
    d1 = inv4h * (PH(ri0, rjp, k0) - PH(ri0, rjm, k0) + PH(rim, rjp, k0) - PH(rim, rjm, k0));
    d2 = inv4h * (PH(ri0, rj0, kp) - PH(ri0, rj0, km) + PH(rim, rj0, kp) - PH(rim, rj0, km));
    return mond_nu<FN>(sqrtf(d0 * d0 + d1 * d1 + d2 * d2) * inv_g0, alpha) * d0;
  } else if (AXIS == 1) {
    d0 = invh * (p0 - PH(ri0, rjm, k0));
    d1 = inv4h * (PH(rip, rj0, k0) - PH(rim, rj0, k0) + PH(rip, rjm, k0) - PH(rim, rjm, k0));
    d2 = inv4h * (PH(ri0, rj0, kp) - PH(ri0, rj0, km) + PH(ri0, rjm, kp) - PH(ri0, rjm, km));
    return mond_nu<FN>(sqrtf(d1 * d1 + d0 * d0 + d2 * d2) * inv_g0, alpha) * d0;
  } else {
    d0 = invh * (p0 - PH(ri0, rj0, km));
    d1 = inv4h * (PH(rip, rj0, k0) - PH(rim, rj0, k0) + PH(rip, rj0, km) - PH(rim, rj0, km));
    d2 = inv4h * (PH(ri0, rjp, k0) - PH(ri0, rjm, k0) + PH(ri0, rjp, km) - PH(ri0, rjm, km));
    return mond_nu<FN>(sqrtf(d1 * d1 + d2 * d2 + d0 * d0) * inv_g0, alpha) * d0;
  }
#undef PH
}

template <int FN>
__global__ void __launch_bounds__(MF_K *MF_J) mond_rhs_march_kernel(const float *__restrict__ phi,
                                                                    float *__restrict__ out, int N, float g0,
                                                                    float alpha) {
  __shared__ float fy[2][MF_J + 1][MF_K], fz[2][MF_J][MF_K + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * MF_K + tx, j = blockIdx.y * MF_J + ty;
  const int i0 = blockIdx.z * MF_CH;
  const bool ok = k < N && j < N;
  const int kc = ok ? k : 0, jc = ok ? j : 0;
  const size_t N2 = (size_t)N * N;
  const float inv_g0 = 1.0f / g0, invh = (float)N, inv4h = 0.25f * (float)N;
  const int km = wrap(kc - 1, N), kp = wrap(kc + 1, N), kpp = wrap(kc + 2, N);
  const size_t rjm = (size_t)wrap(jc - 1, N) * N, rj0 = (size_t)jc * N, rjp = (size_t)wrap(jc + 1, N) * N,
               rjpp = (size_t)wrap(jc + 2, N) * N;
  const bool last_j = ty == MF_J - 1 || j == N - 1, last_k = tx == MF_K - 1 || k == N - 1;
  float fx_prev = 0.0f, s_prev = 0.0f;
  const int nplanes = min(MF_CH, N - i0);
  for (int step = 0; step <= nplanes; step++) {
    const int i = wrap(i0 + step, N);
    const size_t rim = (size_t)wrap(i - 1, N) * N2, ri0 = (size_t)i * N2, rip = (size_t)wrap(i + 1, N) * N2;
    const int buf = step & 1;
    float fx = 0.0f;
    if (ok) {
      fx = mond_face_flux<FN, 0>(phi, rim, ri0, rip, rjm, rj0, rjp, km, kc, kp, invh, inv4h, inv_g0, alpha);
      if (step < nplanes) {
        fy[buf][ty][tx] = mond_face_flux<FN, 1>(phi, rim, ri0, rip, rjm, rj0, rjp, km, kc, kp, invh, inv4h, inv_g0, alpha);
        fz[buf][ty][tx] = mond_face_flux<FN, 2>(phi, rim, ri0, rip, rjm, rj0, rjp, km, kc, kp, invh, inv4h, inv_g0, alpha);
        // the faces just outside the tile: the A face of cell j + 1 (k + 1) is the B face of this one
        if (last_j)
          fy[buf][ty + 1][tx] = mond_face_flux<FN, 1>(phi, rim, ri0, rip, rj0, rjp, rjpp, km, kc, kp, invh, inv4h, inv_g0, alpha);
        if (last_k)
          fz[buf][ty][tx + 1] = mond_face_flux<FN, 2>(phi, rim, ri0, rip, rjm, rj0, rjp, kc, kp, kpp, invh, inv4h, inv_g0, alpha);
      }
    }
    __syncthreads();   // one barrier per plane: the two buffers alternate
    if (ok) {
      if (step > 0) out[(size_t)wrap(i0 + step - 1, N) * N2 + rj0 + kc] = invh * (fx - fx_prev + s_prev);
      if (step < nplanes) s_prev = fy[buf][ty + 1][tx] - fy[buf][ty][tx] + fz[buf][ty][tx + 1] - fz[buf][ty][tx];
      fx_prev = fx;
    }
  }
}

// ---------------------------------------------------------------- plane-marching, shared-memory-staged stencils
// operator / residual / residual norm: a CTA owns a 64 (k) x 8 (j) column and marches over MT_CH planes in i.  The
// plane above / below a cell lives in the thread's registers (each value is loaded from global memory once per
// column instead of three times), the four lateral neighbours come from the plane's tile in shared memory (halo
// rows / columns loaded by the edge threads; two tiles alternate so one barrier per plane suffices).  Compared with
// one thread per cell (7 loads through L1/L2 each, half a million tiny CTAs at 512^3) this is 1 + ~0.3 global loads
// per cell and 16x fewer CTAs.
constexpr int MT_K = 64, MT_J = 8, MT_CH = 16;
enum { MT_OPERATOR = 0, MT_RESIDUAL = 1, MT_SUMSQ = 2 };

template <int KIND, int MODE>
__global__ void __launch_bounds__(MT_K *MT_J) stencil_march_kernel(const float *__restrict__ x,
                                                                   const float *__restrict__ b, float q_val,
                                                                   const float *__restrict__ q_dev,
                                                                   const float *__restrict__ rhs, int N,
                                                                   float *__restrict__ out,
                                                                   double *__restrict__ sumsq) {
  const float q = q_dev ? *q_dev : q_val;
  __shared__ float tile[2][MT_J + 2][MT_K + 2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * MT_K + tx, j = blockIdx.y * MT_J + ty;
  const int i0 = blockIdx.z * MT_CH;
  const bool ok = k < N && j < N;
  const size_t N2 = (size_t)N * N;
  const size_t col = ok ? (size_t)j * N + k : 0;                         // (j, k) offset inside a plane
  const size_t col_km = ok ? (size_t)j * N + wrap(k - 1, N) : 0, col_kp = ok ? (size_t)j * N + wrap(k + 1, N) : 0;
  const size_t col_jm = ok ? (size_t)wrap(j - 1, N) * N + k : 0, col_jp = ok ? (size_t)wrap(j + 1, N) * N + k : 0;
  const bool edge_km = tx == 0, edge_kp = tx == MT_K - 1 || k == N - 1;
  const bool edge_jm = ty == 0, edge_jp = ty == MT_J - 1 || j == N - 1;
  const float h2 = 1.0f / ((float)N * (float)N), invh2 = (float)N * (float)N, invsix = 1.0f / 6.0f;
  float down = ok ? x[(size_t)wrap(i0 - 1, N) * N2 + col] : 0.0f;
  float cur = ok ? x[(size_t)i0 * N2 + col] : 0.0f;
  float up = ok ? x[(size_t)wrap(i0 + 1, N) * N2 + col] : 0.0f;
  double acc = 0.0;
  for (int s = 0; s < MT_CH && i0 + s < N; s++) {
    const int i = i0 + s, buf = s & 1;
    const size_t pl = (size_t)i * N2;
    // two planes of look-ahead: the load issued here is consumed two iterations later
    const float up2 = ok ? x[(size_t)wrap(i + 2, N) * N2 + col] : 0.0f;
    float bt = 0.0f, rt = 0.0f;
    if (ok && (KIND != PSC_OP_LAPLACIAN || MODE != MT_OPERATOR)) bt = b[pl + col];
    if (ok && KIND != PSC_OP_LAPLACIAN && MODE == MT_RESIDUAL) rt = rhs[pl + col];
    if (ok) {
      tile[buf][ty + 1][tx + 1] = cur;
      if (edge_km) tile[buf][ty + 1][tx] = x[pl + col_km];
      if (edge_kp) tile[buf][ty + 1][tx + 2] = x[pl + col_kp];
      if (edge_jm) tile[buf][ty][tx + 1] = x[pl + col_jm];
      if (edge_jp) tile[buf][ty + 2][tx + 1] = x[pl + col_jp];
    }
    __syncthreads();
    if (ok) {
      const float l = tile[buf][ty + 1][tx], r = tile[buf][ty + 1][tx + 2], f = tile[buf][ty][tx + 1],
                  g = tile[buf][ty + 2][tx + 1];
      const size_t t = pl + col;
      float L;
      if (KIND == PSC_OP_LAPLACIAN) {
        // same association as nb6(): ((((a + b) + c) + d) + e) + f with a = i-1, b = j-1, c = k-1, d = k+1, ...
        L = ((((((down + f) + l) + r) + g) + up) - 6.0f * cur) * invh2;
      } else {
        const float s6 = ((((npow(down, KIND) + npow(f, KIND)) + npow(l, KIND)) + npow(r, KIND)) + npow(g, KIND)) +
                         npow(up, KIND);
        const float p = h2 * bt - invsix * s6;
        const float lead = KIND == PSC_OP_CUBIC ? cur * cur * cur : (cur * cur) * (cur * cur);
        L = lead + p * cur + q * h2;
      }
      if (MODE == MT_OPERATOR) out[t] = L;
      else if (MODE == MT_RESIDUAL) out[t] = (KIND == PSC_OP_LAPLACIAN) ? (-L + bt) : (-L + rt);
      else {
        const float res = (KIND == PSC_OP_LAPLACIAN) ? (-L + bt) : L;
        acc += (double)res * (double)res;
      }
    }
    down = cur;
    cur = up;
    up = up2;
  }
  if (MODE == MT_SUMSQ) {
    acc = warp_sum(acc);
    __shared__ double sm[MT_K * MT_J / 32];
    const int tid = ty * MT_K + tx;
    if ((tid & 31) == 0) sm[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < MT_K * MT_J / 32; w++) tot += sm[w];
      atomicAdd(sumsq, tot);
    }
  }
}

static inline dim3 march_grid(int N) {
  return dim3((N + MT_K - 1) / MT_K, (N + MT_J - 1) / MT_J, (N + MT_CH - 1) / MT_CH);
}
static inline dim3 march_block() { return dim3(MT_K, MT_J, 1); }

static inline dim3 cell_grid(int N) { return dim3((N + TKX - 1) / TKX, (N + TJ - 1) / TJ, N); }
static inline dim3 cell_block() { return dim3(TKX, TJ, 1); }

}  // namespace psc

using namespace psc;

#define PSC_KIND_SWITCH(kind, CALL)                     \
  if ((kind) == PSC_OP_LAPLACIAN) { CALL(PSC_OP_LAPLACIAN); } \
  else if ((kind) == PSC_OP_CUBIC) { CALL(PSC_OP_CUBIC); }    \
  else { CALL(PSC_OP_QUARTIC); }

#define PSC_CHECK_GRID(N) PSC_CHECK_ARG((N) >= 2 && (N) <= 32767 && ((N) % 2) == 0, "N must be even, 2..32766")
#define PSC_CHECK_KIND(kind) \
  PSC_CHECK_ARG((kind) >= PSC_OP_LAPLACIAN && (kind) <= PSC_OP_QUARTIC, "unknown operator kind")

static const float *g_q_dev = nullptr;

extern "C" {

/* While set (non-NULL), every f(R) kernel (kinds CUBIC / QUARTIC) reads its q from this device float instead of the
 * by-value argument: a CUDA graph captured with it set keeps working when q changes from step to step (the host
 * rewrites the 4 bytes before each replay).  NULL restores the by-value behaviour. */
int psc_mg_set_q_device(const float *q_dev) {
  g_q_dev = q_dev;
  return PSC_OK;
}
/* (library-internal) the pointer set above, for kernels in other translation units (gs_fused.cu) */
const float *psc_mg_q_device_ptr(void) { return g_q_dev; }

int psc_operator(const float *x, const float *b, float q, int N, int kind, float *out, void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_KIND(kind);
  PSC_CHECK_ARG(x && out && (b || kind == PSC_OP_LAPLACIAN), "null pointer");
#define CALL(K) stencil_march_kernel<K, MT_OPERATOR><<<march_grid(N), march_block(), 0, as_stream(stream)>>>(x, b, q, g_q_dev, nullptr, N, out, nullptr)
  PSC_KIND_SWITCH(kind, CALL)
#undef CALL
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_residual(const float *x, const float *b, float q, const float *rhs, int N, int kind, float *out,
                 void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_KIND(kind);
  PSC_CHECK_ARG(x && b && out && (rhs || kind == PSC_OP_LAPLACIAN), "null pointer");
#define CALL(K) stencil_march_kernel<K, MT_RESIDUAL><<<march_grid(N), march_block(), 0, as_stream(stream)>>>(x, b, q, g_q_dev, rhs, N, out, nullptr)
  PSC_KIND_SWITCH(kind, CALL)
#undef CALL
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_restrict_residual(const float *x, const float *b, int N, float *coarse, void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_ARG(x && b && coarse, "null pointer");
  restrict_residual_kernel<<<cell_grid(N / 2), cell_block(), 0, as_stream(stream)>>>(x, b, N, coarse);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_residual_sumsq(const float *x, const float *b, float q, int N, int kind, double *sumsq_out,
                       void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_KIND(kind);
  PSC_CHECK_ARG(x && b && sumsq_out, "null pointer");
#define CALL(K) stencil_march_kernel<K, MT_SUMSQ><<<march_grid(N), march_block(), 0, as_stream(stream)>>>(x, b, q, g_q_dev, nullptr, N, nullptr, sumsq_out)
  PSC_KIND_SWITCH(kind, CALL)
#undef CALL
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_diff_sumsq(const float *a, float fa, const float *b, int64_t n, double *sumsq_out, void *stream) {
  PSC_CHECK_ARG(n >= 0, "n < 0");
  PSC_CHECK_ARG(a && b && sumsq_out, "null pointer");
  if (n == 0) return PSC_OK;
  diff_sumsq_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(a, fa, b, n, sumsq_out);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_initialise_potential(const float *b, float q, int N, int kind, float *out, void *stream) {
  PSC_CHECK_ARG(N >= 1 && N <= 32767, "N out of range");
  PSC_CHECK_KIND(kind);
  PSC_CHECK_ARG(b && out, "null pointer");
  int64_t n = (int64_t)N * N * N;
#define CALL(K) init_potential_kernel<K><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(b, q, g_q_dev, N, out, n)
  PSC_KIND_SWITCH(kind, CALL)
#undef CALL
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_gauss_seidel(float *x, const float *b, float q, const float *rhs, int N, int kind, float f_relax,
                     void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_KIND(kind);
  PSC_CHECK_ARG(x && b, "null pointer");
  dim3 grid((N / 2 + TKX - 1) / TKX, (N + TJ - 1) / TJ, N);
  for (int colour = 1; colour >= 0; colour--) {
#define CALL(K) gs_colour_kernel<K><<<grid, cell_block(), 0, as_stream(stream)>>>(x, b, q, g_q_dev, rhs, N, f_relax, colour)
    PSC_KIND_SWITCH(kind, CALL)
#undef CALL
    count_launch();
  }
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_restriction(const float *x, int N, float sign, float *coarse, void *stream) {
  PSC_CHECK_GRID(N);
  PSC_CHECK_ARG(x && coarse, "null pointer");
  restriction_kernel<<<cell_grid(N / 2), cell_block(), 0, as_stream(stream)>>>(x, N, sign * 0.125f, coarse);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_prolongation(float *fine, const float *coarse, int Nc, int add, void *stream) {
  PSC_CHECK_ARG(Nc >= 1 && Nc <= 16383, "Nc out of range");
  PSC_CHECK_ARG(fine && coarse, "null pointer");
  if (add)
    prolongation_kernel<true><<<cell_grid(Nc), cell_block(), 0, as_stream(stream)>>>(fine, coarse, Nc);
  else
    prolongation_kernel<false><<<cell_grid(Nc), cell_block(), 0, as_stream(stream)>>>(fine, coarse, Nc);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_mond_rhs(const float *phi, float *out, int N, float g0, int fn, float alpha, void *stream) {
  PSC_CHECK_ARG(N >= 2 && N <= 32767, "N out of range");
  PSC_CHECK_ARG(phi && out && phi != out, "null or aliased pointer");
  PSC_CHECK_ARG(fn >= PSC_MOND_SIMPLE && fn <= PSC_MOND_DELTA, "unknown MOND interpolating function");
  cudaStream_t st = as_stream(stream);
  // every face once (mond_rhs_march_kernel); PSC_MOND_CELLWISE=1 keeps the cell-wise kernel (every face twice)
  static const bool cellwise = getenv("PSC_MOND_CELLWISE") != nullptr;
  const dim3 mgrid((N + MF_K - 1) / MF_K, (N + MF_J - 1) / MF_J, (N + MF_CH - 1) / MF_CH), mblock(MF_K, MF_J);
#define CALL(F)                                                                             \
  do {                                                                                      \
    if (cellwise) mond_rhs_kernel<F><<<cell_grid(N), cell_block(), 0, st>>>(phi, out, N, g0, alpha); \
    else mond_rhs_march_kernel<F><<<mgrid, mblock, 0, st>>>(phi, out, N, g0, alpha);        \
  } while (0)
  switch (fn) {
    case PSC_MOND_SIMPLE: CALL(PSC_MOND_SIMPLE); break;
    case PSC_MOND_N: CALL(PSC_MOND_N); break;
    case PSC_MOND_BETA: CALL(PSC_MOND_BETA); break;
    case PSC_MOND_GAMMA: CALL(PSC_MOND_GAMMA); break;
    default: CALL(PSC_MOND_DELTA); break;
  }
#undef CALL
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
