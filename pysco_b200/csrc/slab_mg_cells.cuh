// slab_mg_cells.cuh -- per-cell bodies of the slab (x-decomposed) multigrid kernels (slab_mg.cu).
//
// A slab level is a box of nxl owned planes of n x n cells.  Potential-like arrays ("xg") carry ONE ghost plane on
// each side along x: array plane p is owned plane p - 1, planes 0 and nxl + 1 mirror the neighbours' boundary planes
// (filled by the host through Comm.exchange_planes).  Right-hand sides ("b") and all outputs hold owned planes only.
// y and z stay periodic inside the box.  The arithmetic (association order included) is the one of the single-domain
// kernels in multigrid.cu, so a one-rank slab reproduces them bit for bit:
//   laplacian.py: operator :12, restrict_residual :125, gauss_seidel :844;  mesh.py: restriction :14,
//   add_prolongation :334;  cubic.py / quartic.py: operator, gauss_seidel[_with_rhs], initialise_potential;
//   mond.py: rhs_simple/n/beta/gamma/delta :171-932.
//
// The bodies are plain functions of (il, j, k) so that the CPU tier can run the very same code through a host
// harness (tests/slab_mg_harness.cpp) -- the __global__ wrappers in slab_mg.cu only map threads to cells.
#pragma once
#include <math.h>
#include <stddef.h>

#include "../../include/pysco_b200.h"
#include "fr_roots.cuh"

#ifdef __CUDACC__
#define PSC_CELL __device__ __forceinline__
#define PSC_UNROLL _Pragma("unroll")
#else
#define PSC_CELL static inline
#define PSC_UNROLL
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif

namespace psc {
namespace box {

PSC_CELL int pwrap(int i, int n) {
  i = i < 0 ? i + n : i;
  return i >= n ? i - n : i;
}

// sum of the six neighbours of owned cell (il, j, k); same order as multigrid.cu nb6(): i-1, j-1, k-1, k+1, j+1, i+1
PSC_CELL float nb6(const float *xg, int il, int j, int k, int n) {
  const size_t n2 = (size_t)n * n;
  const size_t ri = (size_t)(il + 1) * n2, rj = (size_t)j * n;
  float a = xg[ri - n2 + rj + k];
  float b = xg[ri + (size_t)pwrap(j - 1, n) * n + k];
  float c = xg[ri + rj + pwrap(k - 1, n)];
  float d = xg[ri + rj + pwrap(k + 1, n)];
  float e = xg[ri + (size_t)pwrap(j + 1, n) * n + k];
  float f = xg[ri + n2 + rj + k];
  return a + b + c + d + e + f;
}

// one SOR update of owned cell (il, j, k) (laplacian.py:872-1022)
PSC_CELL void gs_cell(float *xg, const float *b, int il, int j, int k, int n, float f_relax) {
  const size_t n2 = (size_t)n * n;
  const size_t tx = (size_t)(il + 1) * n2 + (size_t)j * n + k;
  const size_t tb = (size_t)il * n2 + (size_t)j * n + k;
  const float h2 = 1.0f / ((float)n * (float)n);
  const float invsix = 1.0f / 6.0f;
  float xt = xg[tx];
  float s = nb6(xg, il, j, k, n);
  float target = (s - h2 * b[tb]) * invsix;
  xg[tx] = xt + f_relax * (target - xt);
}

// L x = (sum6 - 6 x) / h^2 at owned cell (il, j, k) (laplacian.py:12-54)
PSC_CELL float operator_cell(const float *xg, int il, int j, int k, int n) {
  const size_t n2 = (size_t)n * n;
  const float invh2 = (float)n * (float)n;
  const float cur = xg[(size_t)(il + 1) * n2 + (size_t)j * n + k];
  return (nb6(xg, il, j, k, n) - 6.0f * cur) * invh2;
}

// coarse owned cell (ci, cj, ck) of the nc = n/2 level: 1/8 sum over its 8 children of (b - Lx)
// (laplacian.py:125-226; 24 outer - 3*8 inner).  Children are owned planes 2ci, 2ci + 1 of the fine slab.
PSC_CELL float restrict_residual_cell(const float *xg, const float *b, int ci, int cj, int ck, int n) {
  const size_t n2 = (size_t)n * n;
  const float invh2 = (float)n * (float)n;
  float outer = 0.0f, inner = 0.0f, bs = 0.0f;
  PSC_UNROLL
  for (int a = 0; a < 2; a++)
    PSC_UNROLL
    for (int e = 0; e < 2; e++)
      PSC_UNROLL
      for (int g = 0; g < 2; g++) {
        const int il = 2 * ci + a, jj = 2 * cj + e, kk = 2 * ck + g;
        const size_t pl = (size_t)(il + 1) * n2;  // plane of the child inside xg
        const size_t t = pl + (size_t)jj * n + kk;
        outer += xg[(a ? pl + n2 : pl - n2) + (size_t)jj * n + kk] +
                 xg[pl + (size_t)pwrap(jj + (e ? 1 : -1), n) * n + kk] +
                 xg[pl + (size_t)jj * n + pwrap(kk + (g ? 1 : -1), n)];
        inner += xg[t];
        bs += b[(size_t)il * n2 + (size_t)jj * n + kk];
      }
  return 0.125f * (-(outer - 3.0f * inner) * invh2 + bs);
}

// coarse owned cell = f * sum of its 8 children; fine holds owned planes only (mesh.py:14-108, f = +-1/8)
PSC_CELL float restriction_cell(const float *fine, int ci, int cj, int ck, int n, float f) {
  const size_t n2 = (size_t)n * n;
  const float *p = fine + (size_t)(2 * ci) * n2 + (size_t)(2 * cj) * n + 2 * ck;
  return f * (p[0] + p[1] + p[n] + p[n + 1] + p[n2] + p[n2 + 1] + p[n2 + n] + p[n2 + n + 1]);
}

// fine_g (+)= trilinear prolongation of coarse_g (mesh.py:334-453), both with ghost planes: the 8 children of coarse
// owned cell (ci, cj, ck) are written from its 27-neighbourhood (planes ci .. ci + 2 of coarse_g)
PSC_CELL void prolong_add_cell(float *fine_g, const float *coarse_g, int ci, int cj, int ck, int nc) {
  const int n = 2 * nc;
  const size_t nc2 = (size_t)nc * nc, n2 = (size_t)n * n;
  const float f0 = 27.0f / 64, f1 = 9.0f / 64, f2 = 3.0f / 64, f3 = 1.0f / 64;
  float v[3][3][3];
  PSC_UNROLL
  for (int a = 0; a < 3; a++)
    PSC_UNROLL
    for (int e = 0; e < 3; e++)
      PSC_UNROLL
      for (int g = 0; g < 3; g++)
        v[a][e][g] = coarse_g[(size_t)(ci + a) * nc2 + (size_t)pwrap(cj + e - 1, nc) * nc + pwrap(ck + g - 1, nc)];
  const float t0 = f0 * v[1][1][1];
  PSC_UNROLL
  for (int a = 0; a < 2; a++)
    PSC_UNROLL
    for (int e = 0; e < 2; e++)
      PSC_UNROLL
      for (int g = 0; g < 2; g++) {
        const int A = 2 * a, E = 2 * e, G = 2 * g;
        const float r = t0 + f1 * (v[A][1][1] + v[1][E][1] + v[1][1][G]) +
                        f2 * (v[A][E][1] + v[1][E][G] + v[A][1][G]) + f3 * v[A][E][G];
        fine_g[(size_t)(2 * ci + a + 1) * n2 + (size_t)(2 * cj + e) * n + 2 * ck + g] += r;
      }
}

// ----------------------------------------------------------------------------------- f(R) scalaron (FAS)
// closed-form roots: fr_roots.cuh (psc::solve_cubic / psc::solve_quartic), shared with multigrid.cu

// KIND = PSC_OP_CUBIC: squares of the neighbours, PSC_OP_QUARTIC: cubes
template <int KIND>
PSC_CELL float npow(float v) {
  return KIND == PSC_OP_CUBIC ? v * v : v * v * v;
}
template <int KIND>
PSC_CELL float nb6_pow(const float *xg, int il, int j, int k, int n) {
  const size_t n2 = (size_t)n * n;
  const size_t ri = (size_t)(il + 1) * n2, rj = (size_t)j * n;
  float a = xg[ri - n2 + rj + k];
  float b = xg[ri + (size_t)pwrap(j - 1, n) * n + k];
  float c = xg[ri + rj + pwrap(k - 1, n)];
  float d = xg[ri + rj + pwrap(k + 1, n)];
  float e = xg[ri + (size_t)pwrap(j + 1, n) * n + k];
  float f = xg[ri + n2 + rj + k];
  return npow<KIND>(a) + npow<KIND>(b) + npow<KIND>(c) + npow<KIND>(d) + npow<KIND>(e) + npow<KIND>(f);
}

// one nonlinear SOR update (cubic.py:269-627, quartic.py:270-628); rhs may be NULL (finest level)
template <int KIND>
PSC_CELL void gs_fr_cell(float *xg, const float *b, const float *rhs, float q, int il, int j, int k, int n,
                         float f_relax) {
  const size_t n2 = (size_t)n * n;
  const size_t tx = (size_t)(il + 1) * n2 + (size_t)j * n + k;
  const size_t tb = (size_t)il * n2 + (size_t)j * n + k;
  const float h2 = 1.0f / ((float)n * (float)n);
  const float invsix = 1.0f / 6.0f;
  float xt = xg[tx];
  float s = nb6_pow<KIND>(xg, il, j, k, n);
  float p = h2 * b[tb] - invsix * s;
  float target;
  if (KIND == PSC_OP_CUBIC) {
    float d1 = 27.0f * h2 * q;
    if (rhs) d1 -= 27.0f * rhs[tb];
    target = solve_cubic(p, d1);
  } else {
    float qq = q * h2;
    if (rhs) qq -= rhs[tb];
    target = solve_quartic(p, qq);
  }
  xg[tx] = xt + f_relax * (target - xt);
}

// L(u) = u^3 + p u + q h^2 (cubic.py:23-81) or u^4 + p u + q h^2 (quartic.py), p = h^2 b - sum6(u^2 | u^3) / 6
template <int KIND>
PSC_CELL float operator_fr_cell(const float *xg, const float *b, float q, int il, int j, int k, int n) {
  const size_t n2 = (size_t)n * n;
  const float h2 = 1.0f / ((float)n * (float)n), invsix = 1.0f / 6.0f;
  const float cur = xg[(size_t)(il + 1) * n2 + (size_t)j * n + k];
  const float s6 = nb6_pow<KIND>(xg, il, j, k, n);
  const float p = h2 * b[(size_t)il * n2 + (size_t)j * n + k] - invsix * s6;
  const float lead = KIND == PSC_OP_CUBIC ? cur * cur * cur : (cur * cur) * (cur * cur);
  return lead + p * cur + q * h2;
}

// first guess from the density term alone (cubic.py:217-259, quartic.py:214-260)
template <int KIND>
PSC_CELL float init_fr_cell(float bt, float q, int n) {
  if (KIND == PSC_OP_CUBIC) {
    const float h2 = 1.0f / ((float)n * (float)n);
    const float threeh2 = 3.0f * h2;
    const double d1 = 27.0 * (double)h2 * (double)q;
    float d0 = -threeh2 * bt;
    double d03 = (double)d0 * (double)d0 * (double)d0;
    double C = cbrt(0.5 * (d1 + sqrt(d1 * d1 - 4.0 * d03)));
    return (float)(-(1.0 / 3) * (C + (double)d0 / C));
  }
  const double h2 = 1.0 / ((double)n * (double)n);
  const double inv3 = 1.0 / 3;
  const double d0 = 12.0 * h2 * (double)q;
  double p = h2 * (double)bt;
  double d1 = 27.0 * (p * p);
  double Q = pow(0.5 * (d1 + sqrt(d1 * d1 - 4.0 * (d0 * d0 * d0))), inv3);
  double S = 0.5 * sqrt((Q + d0 / Q) * inv3);
  return (float)(-S + 0.5 * sqrt(-4.0 * (S * S) + p / S));
}

// ----------------------------------------------------------------------------------- MOND
// interpolating functions nu(y) (mond.py:16-162), as in multigrid.cu mond_nu()
template <int FN>
PSC_CELL float mond_nu(float y, float alpha) {
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

// QUMOND source div(nu(|grad phi_N| / g0) grad phi_N) at owned cell (il, j, k) (mond.py:171-932): phig is the
// Newtonian potential with one ghost plane per side; the six half-cell points A (-h/2) and B (+h/2) of each axis
template <int FN>
PSC_CELL float mond_rhs_cell(const float *phig, int il, int j, int k, int n, float g0, float alpha) {
  const size_t n2 = (size_t)n * n;
  const float inv_g0 = 1.0f / g0;
  const float invh = (float)n, inv4h = 0.25f * (float)n;
  const size_t ri[3] = {(size_t)il * n2, (size_t)(il + 1) * n2, (size_t)(il + 2) * n2};
  const size_t rj[3] = {(size_t)pwrap(j - 1, n) * n, (size_t)j * n, (size_t)pwrap(j + 1, n) * n};
  const int rk[3] = {pwrap(k - 1, n), k, pwrap(k + 1, n)};
#define PSC_P(a, e, g) phig[ri[(a) + 1] + rj[(e) + 1] + rk[(g) + 1]]
  float p0 = PSC_P(0, 0, 0);
  float Axx = invh * (p0 - PSC_P(-1, 0, 0));
  float Axy = inv4h * (PSC_P(0, 1, 0) - PSC_P(0, -1, 0) + PSC_P(-1, 1, 0) - PSC_P(-1, -1, 0));
  float Axz = inv4h * (PSC_P(0, 0, 1) - PSC_P(0, 0, -1) + PSC_P(-1, 0, 1) - PSC_P(-1, 0, -1));
  float fAx = sqrtf(Axx * Axx + Axy * Axy + Axz * Axz);
  float Bxx = invh * (-p0 + PSC_P(1, 0, 0));
  float Bxy = inv4h * (PSC_P(1, 1, 0) - PSC_P(1, -1, 0) + PSC_P(0, 1, 0) - PSC_P(0, -1, 0));
  float Bxz = inv4h * (PSC_P(1, 0, 1) - PSC_P(1, 0, -1) + PSC_P(0, 0, 1) - PSC_P(0, 0, -1));
  float fBx = sqrtf(Bxx * Bxx + Bxy * Bxy + Bxz * Bxz);
  float Ayy = invh * (p0 - PSC_P(0, -1, 0));
  float Ayx = inv4h * (PSC_P(1, 0, 0) - PSC_P(-1, 0, 0) + PSC_P(1, -1, 0) - PSC_P(-1, -1, 0));
  float Ayz = inv4h * (PSC_P(0, 0, 1) - PSC_P(0, 0, -1) + PSC_P(0, -1, 1) - PSC_P(0, -1, -1));
  float fAy = sqrtf(Ayx * Ayx + Ayy * Ayy + Ayz * Ayz);
  float Byy = invh * (-p0 + PSC_P(0, 1, 0));
  float Byx = inv4h * (PSC_P(1, 1, 0) - PSC_P(-1, 1, 0) + PSC_P(1, 0, 0) - PSC_P(-1, 0, 0));
  float Byz = inv4h * (PSC_P(0, 1, 1) - PSC_P(0, 1, -1) + PSC_P(0, 0, 1) - PSC_P(0, 0, -1));
  float fBy = sqrtf(Byx * Byx + Byy * Byy + Byz * Byz);
  float Azz = invh * (p0 - PSC_P(0, 0, -1));
  float Azx = inv4h * (PSC_P(1, 0, 0) - PSC_P(-1, 0, 0) + PSC_P(1, 0, -1) - PSC_P(-1, 0, -1));
  float Azy = inv4h * (PSC_P(0, 1, 0) - PSC_P(0, -1, 0) + PSC_P(0, 1, -1) - PSC_P(0, -1, -1));
  float fAz = sqrtf(Azx * Azx + Azy * Azy + Azz * Azz);
  float Bzz = invh * (-p0 + PSC_P(0, 0, 1));
  float Bzx = inv4h * (PSC_P(1, 0, 1) - PSC_P(-1, 0, 1) + PSC_P(1, 0, 0) - PSC_P(-1, 0, 0));
  float Bzy = inv4h * (PSC_P(0, 1, 1) - PSC_P(0, -1, 1) + PSC_P(0, 1, 0) - PSC_P(0, -1, 0));
  float fBz = sqrtf(Bzx * Bzx + Bzy * Bzy + Bzz * Bzz);
#undef PSC_P
  float r = mond_nu<FN>(fBx * inv_g0, alpha) * Bxx - mond_nu<FN>(fAx * inv_g0, alpha) * Axx +
            mond_nu<FN>(fBy * inv_g0, alpha) * Byy - mond_nu<FN>(fAy * inv_g0, alpha) * Ayy +
            mond_nu<FN>(fBz * inv_g0, alpha) * Bzz - mond_nu<FN>(fAz * inv_g0, alpha) * Azz;
  return invh * r;
}

}  // namespace box
}  // namespace psc
