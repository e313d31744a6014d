// Host harness for the per-cell bodies of the slab multigrid kernels (pysco_b200/csrc/slab_mg_cells.cuh).
// TEST INFRASTRUCTURE ONLY: compiled with g++ by tests/slab_oracle_ops.py so that the CPU tier runs the very code the
// CUDA kernels of csrc/slab_mg.cu execute per thread.  The loops below enumerate cells exactly as the kernels map
// threads to cells (one "thread" per updated / coarse cell).
#include "../pysco_b200/csrc/slab_mg_cells.cuh"

using namespace psc::box;

extern "C" {

void hx_gs_colour(float *xg, const float *b, int nxl, int n, int x0, int colour, float f_relax) {
  for (int il = 0; il < nxl; il++)
    for (int j = 0; j < n; j++)
      for (int kh = 0; kh < n / 2; kh++) {
        const int k = 2 * kh + ((x0 + il + j + colour) & 1);
        gs_cell(xg, b, il, j, k, n, f_relax);
      }
}

void hx_operator(const float *xg, int nxl, int n, float *out) {
  for (int il = 0; il < nxl; il++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++) out[((size_t)il * n + j) * n + k] = operator_cell(xg, il, j, k, n);
}

void hx_restrict_residual(const float *xg, const float *b, int nxl, int n, float *coarse) {
  const int nc = n / 2;
  for (int ci = 0; ci < nxl / 2; ci++)
    for (int cj = 0; cj < nc; cj++)
      for (int ck = 0; ck < nc; ck++)
        coarse[((size_t)ci * nc + cj) * nc + ck] = restrict_residual_cell(xg, b, ci, cj, ck, n);
}

void hx_restriction(const float *fine, int nxl, int n, float sign, float *coarse) {
  const int nc = n / 2;
  for (int ci = 0; ci < nxl / 2; ci++)
    for (int cj = 0; cj < nc; cj++)
      for (int ck = 0; ck < nc; ck++)
        coarse[((size_t)ci * nc + cj) * nc + ck] = restriction_cell(fine, ci, cj, ck, n, sign * 0.125f);
}

void hx_add_prolongation(float *fine_g, const float *coarse_g, int nxlc, int nc) {
  for (int ci = 0; ci < nxlc; ci++)
    for (int cj = 0; cj < nc; cj++)
      for (int ck = 0; ck < nc; ck++) prolong_add_cell(fine_g, coarse_g, ci, cj, ck, nc);
}

void hx_gs_colour_fr(float *xg, const float *b, const float *rhs, float q, int nxl, int n, int x0, int colour,
                     float f_relax, int kind) {
  for (int il = 0; il < nxl; il++)
    for (int j = 0; j < n; j++)
      for (int kh = 0; kh < n / 2; kh++) {
        const int k = 2 * kh + ((x0 + il + j + colour) & 1);
        if (kind == PSC_OP_CUBIC) gs_fr_cell<PSC_OP_CUBIC>(xg, b, rhs, q, il, j, k, n, f_relax);
        else gs_fr_cell<PSC_OP_QUARTIC>(xg, b, rhs, q, il, j, k, n, f_relax);
      }
}

void hx_operator_fr(const float *xg, const float *b, float q, int nxl, int n, int kind, float *out) {
  for (int il = 0; il < nxl; il++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++)
        out[((size_t)il * n + j) * n + k] = kind == PSC_OP_CUBIC ? operator_fr_cell<PSC_OP_CUBIC>(xg, b, q, il, j, k, n)
                                                                 : operator_fr_cell<PSC_OP_QUARTIC>(xg, b, q, il, j, k, n);
}

void hx_init_fr(const float *b, float q, int nxl, int n, int kind, float *out) {
  const size_t count = (size_t)nxl * n * n;
  for (size_t t = 0; t < count; t++)
    out[t] = kind == PSC_OP_CUBIC ? init_fr_cell<PSC_OP_CUBIC>(b[t], q, n) : init_fr_cell<PSC_OP_QUARTIC>(b[t], q, n);
}

void hx_mond_rhs(const float *phig, float *out, int nxl, int n, float g0, int fn, float alpha) {
  for (int il = 0; il < nxl; il++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++) {
        float r;
        switch (fn) {
          case PSC_MOND_SIMPLE: r = mond_rhs_cell<PSC_MOND_SIMPLE>(phig, il, j, k, n, g0, alpha); break;
          case PSC_MOND_N: r = mond_rhs_cell<PSC_MOND_N>(phig, il, j, k, n, g0, alpha); break;
          case PSC_MOND_BETA: r = mond_rhs_cell<PSC_MOND_BETA>(phig, il, j, k, n, g0, alpha); break;
          case PSC_MOND_GAMMA: r = mond_rhs_cell<PSC_MOND_GAMMA>(phig, il, j, k, n, g0, alpha); break;
          default: r = mond_rhs_cell<PSC_MOND_DELTA>(phig, il, j, k, n, g0, alpha); break;
        }
        out[((size_t)il * n + j) * n + k] = r;
      }
}

}  // extern "C"
