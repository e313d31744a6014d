// deposit_tiled.cu -- host launcher of the warp-private walking-window deposit (deposit_window.cuh).
#include <cstdlib>
#include <cstring>

#include "deposit_window.cuh"

namespace psc {

// rho = f1 * (scale * rho) + f2, defined in deposit.cu
__global__ void rho_affine_kernel(float *rho, int64_t n, float scale, float f1, float f2, int do_scale);

static int deposit_mode() {
  // PSC_DEPOSIT_MODE=atomic forces the one-RED-per-stencil-point kernel (A/B measurements)
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("PSC_DEPOSIT_MODE");
    mode = (e && strcmp(e, "atomic") == 0) ? 1 : 0;
  }
  return mode;
}

template <int SCHEME, int DBG = 0>
static int launch_window(const float *pos, int64_t np, int N, float *rho, cudaStream_t st, DepositStats *stats) {
  const size_t smem = sizeof(float) * DW_WARPS * DW_CAP;
  static bool configured = false;
  if (!configured) {
    PSC_CUDA(cudaFuncSetAttribute(deposit_window_kernel<SCHEME, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int64_t nchunks = (np + 31) >> 5;
  int64_t ctas = (int64_t)kNumSMs * DW_CTAS_PER_SM;
  int64_t warps = ctas * DW_WARPS;
  // at least 8 chunks per warp so that a window is reused; fewer CTAs for small inputs
  int64_t cpw = (nchunks + warps - 1) / warps;
  if (cpw < 8) cpw = 8;
  ctas = (nchunks + cpw * DW_WARPS - 1) / (cpw * DW_WARPS);
  if (ctas < 1) ctas = 1;
  deposit_window_kernel<SCHEME, DBG><<<(int)ctas, DW_WARPS * 32, smem, st>>>(pos, np, N, rho, cpw, stats);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int deposit_tiled(const float *pos, int64_t np, int N, int scheme, float scale, float f1, float f2, float *rho,
                  cudaStream_t st) {
  if (deposit_mode() == 1 || (N & 3) != 0 || N < 32 || np == 0) return 0;
  const int64_t n3 = (int64_t)N * N * N;
  PSC_CUDA(cudaMemsetAsync(rho, 0, sizeof(float) * n3, st));
  int rc;
  if (scheme == PSC_TSC) rc = launch_window<PSC_TSC>(pos, np, N, rho, st, nullptr);
  else if (scheme == PSC_CIC) rc = launch_window<PSC_CIC>(pos, np, N, rho, st, nullptr);
  else rc = launch_window<PSC_NGP>(pos, np, N, rho, st, nullptr);
  if (rc != PSC_OK) return rc;
  if (scale != 1.0f || f1 != 1.0f || f2 != 0.0f) {
    rho_affine_kernel<<<grid_for((n3 + 3) / 4, 256), 256, 0, st>>>(rho, n3, scale, f1, f2, scale != 1.0f);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return 1;
}

}  // namespace psc

// debug / measurement entry (not part of the reference-facing ABI): deposit with statistics
extern "C" int psc_deposit_window_dbg(const float *pos, int64_t np, int N, int dbg, float *rho, void *stream) {
  // timing experiments only (results are WRONG for dbg != 0): bit0 no __syncwarp, bit1 no duplicate merge, bit2 no flush
  using namespace psc;
  cudaStream_t st = as_stream(stream);
  switch (dbg) {
    case 1: return launch_window<PSC_TSC, 1>(pos, np, N, rho, st, nullptr);
    case 2: return launch_window<PSC_TSC, 2>(pos, np, N, rho, st, nullptr);
    case 3: return launch_window<PSC_TSC, 3>(pos, np, N, rho, st, nullptr);
    case 4: return launch_window<PSC_TSC, 4>(pos, np, N, rho, st, nullptr);
    case 7: return launch_window<PSC_TSC, 7>(pos, np, N, rho, st, nullptr);
    default: return launch_window<PSC_TSC, 0>(pos, np, N, rho, st, nullptr);
  }
}

extern "C" int psc_deposit_window_stats(const float *pos, int64_t np, int N, int scheme, float *rho,
                                        unsigned long long *stats3, void *stream) {
  using namespace psc;
  PSC_CHECK_ARG((N & 3) == 0 && N >= 32, "window deposit needs N % 4 == 0 and N >= 32");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(rho, 0, sizeof(float) * (size_t)N * N * N, st));
  PSC_CUDA(cudaMemsetAsync(stats3, 0, sizeof(DepositStats), st));
  DepositStats *s = reinterpret_cast<DepositStats *>(stats3);
  if (scheme == PSC_TSC) return launch_window<PSC_TSC>(pos, np, N, rho, st, s);
  if (scheme == PSC_CIC) return launch_window<PSC_CIC>(pos, np, N, rho, st, s);
  return launch_window<PSC_NGP>(pos, np, N, rho, st, s);
}
