// gs_fused.cu -- one red-black SOR sweep (laplacian.gauss_seidel laplacian.py:844-1022, cubic.gauss_seidel[_with_rhs]
// cubic.py:269-627, quartic.gauss_seidel[_with_rhs]) as ONE plane-marching, shared-memory-staged kernel.
//
// The two-launch sweep of multigrid.cu (gs_colour_kernel: odd-parity cells, then even-parity cells) moves 24 B per cell
// (x is read and written by both colour passes, b by both).  Here a CTA owns a (32 x 64) tile of (j, k) columns and
// marches over a chunk of planes i with a ring of four planes of x in shared memory:
//     step s:  load plane s + 2                         (TMA bulk copies: one or two cp.async.bulk per 288-byte row,
//                                                        completion on an mbarrier; or plain LDG.128 / STS.128)
//              red cells of plane s + 1  <- old black neighbours in planes s, s + 1, s + 2   (written into the ring)
//              black cells of plane s    <- updated red neighbours in planes s - 1, s, s + 1 (written to x_out with the
//                                                                                             plane's updated red cells)
// Red cells are updated on the tile plus a ring of one cell (the black update of the tile's edge needs them), so a plane
// is staged with a halo of two: x is read 1.2 times, b / rhs once, x_out written once -- 13 B per cell instead of 24.
// The sweep is OUT OF PLACE: a neighbouring CTA needs the OLD values of this CTA's edge cells whenever it gets to them.
// Arithmetic and association of the neighbour sum are those of gs_colour_kernel, so the result is bit-identical to the
// two-launch sweep (tests/test_gs_fused_gpu.py).
//
// Measured at 512^3 (profiles/r02_multigrid_kernels.txt, r02_grid_kernels_ncu.txt): DRAM traffic 1.71 GB per sweep
// against 3.14 GB for the two colour launches, 0.687 ms against 0.717 ms -- the kernel is bound by instruction issue and
// barrier / copy latency (three barriers per plane, 47 % issue-active), not by HBM.  A second version (lanes of a warp on
// two rows so that the stride-2 colour accesses hit distinct banks, fully unrolled task loops, five-plane ring with the
// next plane's TMA in flight under the arithmetic on a second mbarrier) removed the 2-way bank conflicts but issued more
// instructions and ran at 0.79 ms; this first version is the one kept.
#include <cuda/barrier>
#include <cuda/ptx>

#include "common.cuh"
#include "fr_roots.cuh"

namespace psc {

constexpr int GF_TJ = 32, GF_TK = 64;      // tile of updated cells
constexpr int GF_RJ = GF_TJ + 4;           // staged rows: halo of two
constexpr int GF_RK = GF_TK + 8;           // staged columns: halo of four, so that a row starts 16-byte aligned
constexpr int GF_PLANE = GF_RJ * GF_RK;    // 2592 floats = 10.1 KB
constexpr int GF_THREADS = 256;
constexpr int GF_CHUNK = 32;               // planes per CTA

template <int KIND>
__device__ __forceinline__ float gsf_pw(float v) {
  return KIND == PSC_OP_LAPLACIAN ? v : (KIND == PSC_OP_CUBIC ? v * v : v * v * v);
}

// x + f_relax (target - x) for one cell: s6 = sum over the six neighbours (of v, v^2 or v^3), in gs_colour_kernel's order
template <int KIND>
__device__ __forceinline__ float gsf_update(float xt, float s6, float bt, float q, bool has_rhs, float rt, float h2,
                                            float f_relax) {
  const float invsix = 1.0f / 6.0f;
  float target;
  if (KIND == PSC_OP_LAPLACIAN) {
    target = (s6 - h2 * bt) * invsix;
  } else {
    const float p = h2 * bt - invsix * s6;
    if (KIND == PSC_OP_CUBIC) {
      float d1 = 27.0f * h2 * q;
      if (has_rhs) d1 -= 27.0f * rt;
      target = solve_cubic(p, d1);
    } else {
      float qq = q * h2;
      if (has_rhs) qq -= rt;
      target = solve_quartic(p, qq);
    }
  }
  return xt + f_relax * (target - xt);
}

using gsf_barrier = cuda::barrier<cuda::thread_scope_block>;

// stage plane gi of x (rows j0 - 2 .. j0 + 33, columns k0 - 4 .. k0 + 67, periodic) into `dst`
template <bool TMA>
__device__ __forceinline__ void gsf_stage_plane(const float *__restrict__ x, int N, int gi, int j0, int k0,
                                                float *dst, gsf_barrier *bar) {
  const size_t plane = (size_t)gi * N * N;
  if (TMA) {
    // one thread per row: the row is contiguous in memory except where it crosses the periodic boundary in k
    const int r = threadIdx.x;
    if (r < GF_RJ) {
      const float *row = x + plane + (size_t)wrap(j0 - 2 + r, N) * N;
      float *d = dst + r * GF_RK;
      if (k0 == 0) {
        cuda::memcpy_async(d, row + (N - 4), cuda::aligned_size_t<16>(16), *bar);
        cuda::memcpy_async(d + 4, row, cuda::aligned_size_t<16>(sizeof(float) * (GF_RK - 4)), *bar);
      } else if (k0 + GF_TK == N) {
        cuda::memcpy_async(d, row + (k0 - 4), cuda::aligned_size_t<16>(sizeof(float) * (GF_RK - 4)), *bar);
        cuda::memcpy_async(d + (GF_RK - 4), row, cuda::aligned_size_t<16>(16), *bar);
      } else {
        cuda::memcpy_async(d, row + (k0 - 4), cuda::aligned_size_t<16>(sizeof(float) * GF_RK), *bar);
      }
    }
  } else {
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    const int n4 = N >> 2;
    for (int t = threadIdx.x; t < GF_RJ * (GF_RK / 4); t += GF_THREADS) {
      const int r = t / (GF_RK / 4), q = t - r * (GF_RK / 4);
      int g4 = (k0 >> 2) - 1 + q;
      g4 += g4 < 0 ? n4 : 0;
      g4 -= g4 >= n4 ? n4 : 0;
      d4[t] = __ldg(&x4[(plane + (size_t)wrap(j0 - 2 + r, N) * N) / 4 + g4]);
    }
  }
}

template <int KIND, bool TMA>
__global__ void __launch_bounds__(GF_THREADS) gs_fused_kernel(const float *__restrict__ x, const float *__restrict__ b,
                                                              float q_val, const float *__restrict__ q_dev,
                                                              const float *__restrict__ rhs, int N, float f_relax,
                                                              float *__restrict__ out) {
  __shared__ __align__(128) float ring[4][GF_PLANE];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ gsf_barrier bar;
  const float q = q_dev ? *q_dev : q_val;
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * GF_TK, j0 = blockIdx.y * GF_TJ, i0 = blockIdx.z * GF_CHUNK;
  const int i1 = min(N, i0 + GF_CHUNK);
  const float h2 = 1.0f / ((float)N * (float)N);
  const bool has_rhs = rhs != nullptr;
  const size_t N2 = (size_t)N * N;
  if (TMA) {
    if (tid == 0) {
      init(&bar, GF_THREADS);
      cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    }
    __syncthreads();
  }

  // all threads: wait until the planes staged since the last call have landed
  auto landed = [&]() {
    if (TMA) {
      bar.arrive_and_wait();
    } else {
      __syncthreads();
    }
  };
  // before the async proxy overwrites a ring slot that this CTA has read / written through the generic proxy
  auto release_slot = [&]() {
    if (TMA) cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    __syncthreads();
  };

  // red cells (odd i + j + k) of plane p on the tile + one ring; neighbours: old black cells
  auto red_stage = [&](int p) {
    const float *lo = ring[(p - 1) & 3], *hi = ring[(p + 1) & 3];
    float *mid = ring[p & 3];
    const int gp = wrap(p, N);
    for (int idx = tid; idx < (GF_TJ + 2) * (GF_TK / 2 + 1); idx += GF_THREADS) {
      const int rr = idx / (GF_TK / 2 + 1), m = idx - rr * (GF_TK / 2 + 1);
      const int r = 1 + rr;                                  // staged row (global j = j0 - 2 + r)
      const int c = 3 + 2 * m + ((p + r) & 1);               // staged column (global k = k0 - 4 + c), odd parity
      const int o = r * GF_RK + c;
      const float s6 = ((((gsf_pw<KIND>(lo[o]) + gsf_pw<KIND>(mid[o - GF_RK])) + gsf_pw<KIND>(mid[o - 1])) +
                         gsf_pw<KIND>(mid[o + 1])) + gsf_pw<KIND>(mid[o + GF_RK])) + gsf_pw<KIND>(hi[o]);
      const size_t t = (size_t)gp * N2 + (size_t)wrap(j0 - 2 + r, N) * N + wrap(k0 - 4 + c, N);
      mid[o] = gsf_update<KIND>(mid[o], s6, __ldg(&b[t]), q, has_rhs, has_rhs ? __ldg(&rhs[t]) : 0.0f, h2, f_relax);
    }
  };
  // black cells (even parity) of plane p on the tile, from the updated red cells; writes the finished plane
  auto black_stage = [&](int p) {
    const float *lo = ring[(p - 1) & 3], *mid = ring[p & 3], *hi = ring[(p + 1) & 3];
    const size_t base = (size_t)p * N2;
    for (int idx = tid; idx < GF_TJ * (GF_TK / 2); idx += GF_THREADS) {
      const int rr = idx / (GF_TK / 2), m = idx - rr * (GF_TK / 2);
      const int r = 2 + rr;
      const int ob = (p + r) & 1;                            // which of the pair (4 + 2m, 5 + 2m) is black
      const int o = r * GF_RK + 4 + 2 * m + ob;
      const float s6 = ((((gsf_pw<KIND>(lo[o]) + gsf_pw<KIND>(mid[o - GF_RK])) + gsf_pw<KIND>(mid[o - 1])) +
                         gsf_pw<KIND>(mid[o + 1])) + gsf_pw<KIND>(mid[o + GF_RK])) + gsf_pw<KIND>(hi[o]);
      const size_t t = base + (size_t)(j0 + rr) * N + (k0 + 2 * m);
      const float nb = gsf_update<KIND>(mid[o], s6, __ldg(&b[t + ob]), q, has_rhs, has_rhs ? __ldg(&rhs[t + ob]) : 0.0f,
                                        h2, f_relax);
      const float red = mid[o + 1 - 2 * ob];
      *reinterpret_cast<float2 *>(out + t) = ob ? make_float2(red, nb) : make_float2(nb, red);
    }
  };

  // prologue: planes i0 - 2 .. i0 + 1, red cells of planes i0 - 1 and i0
  for (int p = i0 - 2; p <= i0 + 1; p++) gsf_stage_plane<TMA>(x, N, wrap(p, N), j0, k0, ring[p & 3], &bar);
  landed();
  red_stage(i0 - 1);
  __syncthreads();
  red_stage(i0);
  for (int s = i0; s < i1; s++) {
    release_slot();                                   // plane s - 2 is dead: its slot takes plane s + 2
    gsf_stage_plane<TMA>(x, N, wrap(s + 2, N), j0, k0, ring[(s + 2) & 3], &bar);
    landed();
    red_stage(s + 1);
    __syncthreads();
    black_stage(s);
  }
}

}  // namespace psc

using namespace psc;

extern "C" {

// defined in multigrid.cu: the device-resident q of a captured f(R) graph
const float *psc_mg_q_device_ptr(void);

int psc_gauss_seidel_fused_supported(int N) { return N >= 128 && (N % GF_TK) == 0 ? 1 : 0; }

int psc_gauss_seidel_fused(const float *x, const float *b, float q, const float *rhs, int N, int kind, float f_relax,
                           float *x_out, int use_tma, void *stream) {
  PSC_CHECK_ARG(psc_gauss_seidel_fused_supported(N), "the fused sweep needs N >= 128, N % 64 == 0");
  PSC_CHECK_ARG(kind >= PSC_OP_LAPLACIAN && kind <= PSC_OP_QUARTIC, "unknown operator kind");
  PSC_CHECK_ARG(x && b && x_out && x != x_out, "null or aliased pointer (the fused sweep is out of place)");
  PSC_CHECK_ARG((((uintptr_t)x | (uintptr_t)x_out) & 15) == 0, "x and x_out must be 16-byte aligned");
  dim3 grid(N / GF_TK, N / GF_TJ, (N + GF_CHUNK - 1) / GF_CHUNK);
  cudaStream_t st = as_stream(stream);
  const float *qd = psc_mg_q_device_ptr();
#define GSF(K)                                                                                       \
  do {                                                                                               \
    if (use_tma) gs_fused_kernel<K, true><<<grid, GF_THREADS, 0, st>>>(x, b, q, qd, rhs, N, f_relax, x_out); \
    else gs_fused_kernel<K, false><<<grid, GF_THREADS, 0, st>>>(x, b, q, qd, rhs, N, f_relax, x_out);        \
  } while (0)
  if (kind == PSC_OP_LAPLACIAN) GSF(PSC_OP_LAPLACIAN);
  else if (kind == PSC_OP_CUBIC) GSF(PSC_OP_CUBIC);
  else GSF(PSC_OP_QUARTIC);
#undef GSF
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
