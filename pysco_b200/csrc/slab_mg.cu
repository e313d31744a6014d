// slab_mg.cu -- multigrid kernels on an x-slab with ghost planes (stencils / GS / prolongation with
// ghost-plane copies from the neighbours; the host side is pysco_b200/slab.py: Slab.multigrid_poisson).
//
// Layout and arithmetic: slab_mg_cells.cuh.  Thread mapping as in multigrid.cu: threadIdx.x walks k (coalesced rows),
// blockIdx.y / threadIdx.y carry j, blockIdx.z the owned plane.  The red-black parity of a cell is the parity of its
// GLOBAL indices, x0 + il + j + k, so that P slabs together perform exactly the single-domain sweep.
#include "common.cuh"
#include "slab_mg_cells.cuh"

namespace psc {

constexpr int BX_K = 64;  // k-threads per CTA
constexpr int BX_J = 4;   // rows (j) per CTA

// colour = 1: odd x0 + il + j + k ("red", first in the reference), colour = 0: even
__global__ void __launch_bounds__(BX_K *BX_J) box_gs_colour_kernel(float *__restrict__ xg,
                                                                   const float *__restrict__ b, int nxl, int n,
                                                                   int x0, int colour, float f_relax) {
  const int kh = blockIdx.x * BX_K + threadIdx.x;
  const int j = blockIdx.y * BX_J + threadIdx.y;
  const int il = blockIdx.z;
  const int k = 2 * kh + ((x0 + il + j + colour) & 1);
  if (k >= n || j >= n || il >= nxl) return;
  box::gs_cell(xg, b, il, j, k, n, f_relax);
}

__global__ void __launch_bounds__(BX_K *BX_J) box_operator_kernel(const float *__restrict__ xg, int nxl, int n,
                                                                  float *__restrict__ out) {
  const int k = blockIdx.x * BX_K + threadIdx.x;
  const int j = blockIdx.y * BX_J + threadIdx.y;
  const int il = blockIdx.z;
  if (k >= n || j >= n || il >= nxl) return;
  out[((size_t)il * n + j) * n + k] = box::operator_cell(xg, il, j, k, n);
}

// thread per COARSE owned cell; nxlc = nxl / 2 coarse planes
__global__ void __launch_bounds__(BX_K *BX_J) box_restrict_residual_kernel(const float *__restrict__ xg,
                                                                           const float *__restrict__ b, int nxlc,
                                                                           int n, float *__restrict__ coarse) {
  const int nc = n >> 1;
  const int ck = blockIdx.x * BX_K + threadIdx.x;
  const int cj = blockIdx.y * BX_J + threadIdx.y;
  const int ci = blockIdx.z;
  if (ck >= nc || cj >= nc || ci >= nxlc) return;
  coarse[((size_t)ci * nc + cj) * nc + ck] = box::restrict_residual_cell(xg, b, ci, cj, ck, n);
}

__global__ void __launch_bounds__(BX_K *BX_J) box_restriction_kernel(const float *__restrict__ fine, int nxlc, int n,
                                                                     float f, float *__restrict__ coarse) {
  const int nc = n >> 1;
  const int ck = blockIdx.x * BX_K + threadIdx.x;
  const int cj = blockIdx.y * BX_J + threadIdx.y;
  const int ci = blockIdx.z;
  if (ck >= nc || cj >= nc || ci >= nxlc) return;
  coarse[((size_t)ci * nc + cj) * nc + ck] = box::restriction_cell(fine, ci, cj, ck, n, f);
}

// thread per COARSE owned cell: writes its 8 children in the fine slab
__global__ void __launch_bounds__(BX_K *BX_J) box_add_prolongation_kernel(float *__restrict__ fine_g,
                                                                          const float *__restrict__ coarse_g,
                                                                          int nxlc, int nc) {
  const int ck = blockIdx.x * BX_K + threadIdx.x;
  const int cj = blockIdx.y * BX_J + threadIdx.y;
  const int ci = blockIdx.z;
  if (ck >= nc || cj >= nc || ci >= nxlc) return;
  box::prolong_add_cell(fine_g, coarse_g, ci, cj, ck, nc);
}

// f(R): KIND = PSC_OP_CUBIC (n = 1) or PSC_OP_QUARTIC (n = 2); rhs (FAS coarse levels) may be NULL
template <int KIND>
__global__ void __launch_bounds__(BX_K *BX_J) box_gs_colour_fr_kernel(float *__restrict__ xg,
                                                                      const float *__restrict__ b,
                                                                      const float *__restrict__ rhs, float q,
                                                                      int nxl, int n, int x0, int colour,
                                                                      float f_relax) {
  const int kh = blockIdx.x * BX_K + threadIdx.x;
  const int j = blockIdx.y * BX_J + threadIdx.y;
  const int il = blockIdx.z;
  const int k = 2 * kh + ((x0 + il + j + colour) & 1);
  if (k >= n || j >= n || il >= nxl) return;
  box::gs_fr_cell<KIND>(xg, b, rhs, q, il, j, k, n, f_relax);
}

template <int KIND>
__global__ void __launch_bounds__(BX_K *BX_J) box_operator_fr_kernel(const float *__restrict__ xg,
                                                                     const float *__restrict__ b, float q, int nxl,
                                                                     int n, float *__restrict__ out) {
  const int k = blockIdx.x * BX_K + threadIdx.x;
  const int j = blockIdx.y * BX_J + threadIdx.y;
  const int il = blockIdx.z;
  if (k >= n || j >= n || il >= nxl) return;
  out[((size_t)il * n + j) * n + k] = box::operator_fr_cell<KIND>(xg, b, q, il, j, k, n);
}

template <int KIND>
__global__ void __launch_bounds__(256) box_init_fr_kernel(const float *__restrict__ b, float q, int n, int64_t count,
                                                          float *__restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
    out[t] = box::init_fr_cell<KIND>(b[t], q, n);
}

template <int FN>
__global__ void __launch_bounds__(BX_K *BX_J) box_mond_rhs_kernel(const float *__restrict__ phig,
                                                                  float *__restrict__ out, int nxl, int n, float g0,
                                                                  float alpha) {
  const int k = blockIdx.x * BX_K + threadIdx.x;
  const int j = blockIdx.y * BX_J + threadIdx.y;
  const int il = blockIdx.z;
  if (k >= n || j >= n || il >= nxl) return;
  out[((size_t)il * n + j) * n + k] = box::mond_rhs_cell<FN>(phig, il, j, k, n, g0, alpha);
}

static inline dim3 box_grid(int nk, int nj, int planes) {
  return dim3((nk + BX_K - 1) / BX_K, (nj + BX_J - 1) / BX_J, planes);
}
static inline dim3 box_block() { return dim3(BX_K, BX_J, 1); }

}  // namespace psc

using namespace psc;

#define PSC_CHECK_BOX(nxl, n)                                                                  \
  PSC_CHECK_ARG((n) >= 2 && (n) <= 32766 && ((n) % 2) == 0, "n must be even, 2..32766");       \
  PSC_CHECK_ARG((nxl) >= 1 && (nxl) <= 32766, "nxl must be 1..32766")

extern "C" {

int psc_box_gauss_seidel_colour(float *xg, const float *b, int nxl, int n, int x0, int colour, float f_relax,
                                void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_ARG(xg && b, "null pointer");
  PSC_CHECK_ARG(x0 >= 0 && (colour == 0 || colour == 1), "x0 must be >= 0 and colour 0 or 1");
  box_gs_colour_kernel<<<box_grid(n / 2, n, nxl), box_block(), 0, as_stream(stream)>>>(xg, b, nxl, n, x0, colour,
                                                                                      f_relax);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_operator(const float *xg, int nxl, int n, float *out, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_ARG(xg && out, "null pointer");
  box_operator_kernel<<<box_grid(n, n, nxl), box_block(), 0, as_stream(stream)>>>(xg, nxl, n, out);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_restrict_residual(const float *xg, const float *b, int nxl, int n, float *coarse, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_ARG((nxl % 2) == 0, "nxl must be even (2:1 coarsening of the owned planes)");
  PSC_CHECK_ARG(xg && b && coarse, "null pointer");
  box_restrict_residual_kernel<<<box_grid(n / 2, n / 2, nxl / 2), box_block(), 0, as_stream(stream)>>>(
      xg, b, nxl / 2, n, coarse);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_restriction(const float *fine, int nxl, int n, float sign, float *coarse, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_ARG((nxl % 2) == 0, "nxl must be even (2:1 coarsening of the owned planes)");
  PSC_CHECK_ARG(fine && coarse, "null pointer");
  box_restriction_kernel<<<box_grid(n / 2, n / 2, nxl / 2), box_block(), 0, as_stream(stream)>>>(
      fine, nxl / 2, n, sign * 0.125f, coarse);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_add_prolongation(float *fine_g, const float *coarse_g, int nxlc, int nc, void *stream) {
  PSC_CHECK_ARG(nc >= 1 && nc <= 16383 && nxlc >= 1 && nxlc <= 16383, "nc / nxlc out of range");
  PSC_CHECK_ARG(fine_g && coarse_g, "null pointer");
  box_add_prolongation_kernel<<<box_grid(nc, nc, nxlc), box_block(), 0, as_stream(stream)>>>(fine_g, coarse_g, nxlc,
                                                                                            nc);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

#define PSC_CHECK_FR_KIND(kind) \
  PSC_CHECK_ARG((kind) == PSC_OP_CUBIC || (kind) == PSC_OP_QUARTIC, "kind must be PSC_OP_CUBIC or PSC_OP_QUARTIC")

int psc_box_gauss_seidel_colour_fr(float *xg, const float *b, const float *rhs, float q, int nxl, int n, int x0,
                                   int colour, float f_relax, int kind, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_FR_KIND(kind);
  PSC_CHECK_ARG(xg && b, "null pointer");
  PSC_CHECK_ARG(x0 >= 0 && (colour == 0 || colour == 1), "x0 must be >= 0 and colour 0 or 1");
  if (kind == PSC_OP_CUBIC)
    box_gs_colour_fr_kernel<PSC_OP_CUBIC><<<box_grid(n / 2, n, nxl), box_block(), 0, as_stream(stream)>>>(
        xg, b, rhs, q, nxl, n, x0, colour, f_relax);
  else
    box_gs_colour_fr_kernel<PSC_OP_QUARTIC><<<box_grid(n / 2, n, nxl), box_block(), 0, as_stream(stream)>>>(
        xg, b, rhs, q, nxl, n, x0, colour, f_relax);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_operator_fr(const float *xg, const float *b, float q, int nxl, int n, int kind, float *out,
                        void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_FR_KIND(kind);
  PSC_CHECK_ARG(xg && b && out, "null pointer");
  if (kind == PSC_OP_CUBIC)
    box_operator_fr_kernel<PSC_OP_CUBIC><<<box_grid(n, n, nxl), box_block(), 0, as_stream(stream)>>>(xg, b, q, nxl,
                                                                                                    n, out);
  else
    box_operator_fr_kernel<PSC_OP_QUARTIC><<<box_grid(n, n, nxl), box_block(), 0, as_stream(stream)>>>(xg, b, q, nxl,
                                                                                                      n, out);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_initialise_potential_fr(const float *b, float q, int nxl, int n, int kind, float *out, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_FR_KIND(kind);
  PSC_CHECK_ARG(b && out, "null pointer");
  const int64_t count = (int64_t)nxl * n * n;
  if (kind == PSC_OP_CUBIC)
    box_init_fr_kernel<PSC_OP_CUBIC><<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(b, q, n, count, out);
  else
    box_init_fr_kernel<PSC_OP_QUARTIC><<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(b, q, n, count, out);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_box_mond_rhs(const float *phig, float *out, int nxl, int n, float g0, int fn, float alpha, void *stream) {
  PSC_CHECK_BOX(nxl, n);
  PSC_CHECK_ARG(phig && out && phig != out, "null or aliased pointer");
  PSC_CHECK_ARG(fn >= PSC_MOND_SIMPLE && fn <= PSC_MOND_DELTA, "unknown MOND interpolating function");
  cudaStream_t st = as_stream(stream);
#define CALL(F) box_mond_rhs_kernel<F><<<box_grid(n, n, nxl), box_block(), 0, st>>>(phig, out, nxl, n, g0, alpha)
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
