// tools/ubench_deposit2.cu -- second deposit micro-benchmark (see DESIGN.md "deposit").
//   var0  global RED.F32, one per stencil point (27 / particle)
//   var4  global RED vectorised along k (RED.F32x2 / x4 chosen by alignment)
//   var5  per-WARP private shared-memory tile, conflict-free plain LDS/FADD/STS in 27 phases
//         (duplicate cells merged by match_any + shuffles), flushed with RED.F32x4
// Particles are generated on the device in Morton order with Gaussian jitter ("sorted a few steps ago").
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <math.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int wrapN(int i, int N) { i = i < 0 ? i + N : i; return i >= N ? i - N : i; }
__device__ __forceinline__ void tsc_axis(float xp, int &c, float &wm, float &w0, float &wp) {
  c = (int)xp; float d = xp - 0.5f - (float)c; w0 = 0.75f - d * d; float m = 0.5f - d, p = 0.5f + d;
  wm = 0.5f * (m * m); wp = 0.5f * (p * p);
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t compact3(uint64_t k) { k &= 0x1249249249249249ull; k = (k ^ (k >> 2)) & 0x10C30C30C30C30C3ull;
  k = (k ^ (k >> 4)) & 0x100F00F00F00F00Full; k = (k ^ (k >> 8)) & 0x1F0000FF0000FFull; k = (k ^ (k >> 16)) & 0x1F00000000FFFFull;
  k = (k ^ (k >> 32)) & 0x1FFFFF; return (uint32_t)k; }

__global__ void gen_particles(float *pos, int64_t np, int N, float sigma, int lexicographic) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    int i, j, k;
    if (lexicographic) { k = n % N; j = (n / N) % N; i = n / ((int64_t)N * N); }
    else { i = compact3((uint64_t)n >> 2); j = compact3((uint64_t)n >> 1); k = compact3((uint64_t)n); }
    float c[3] = {(float)i, (float)j, (float)k};
    for (int d = 0; d < 3; d++) {
      uint32_t h1 = hash32((uint32_t)(n * 6 + 2 * d + 1)), h2 = hash32((uint32_t)(n * 6 + 2 * d + 2) ^ 0x9e3779b9U);
      float u = (h1 + 1.0f) * (1.0f / 4294967808.0f), v = h2 * (1.0f / 4294967296.0f);
      float g = sqrtf(-2.0f * logf(u)) * cospif(2.0f * v);
      float x = (c[d] + 0.5f + sigma * g) / (float)N; x -= floorf(x); if (x >= 1.0f) x = 0.0f;
      pos[3 * n + d] = x;
    }
  }
}

__global__ void k_global(const float *__restrict__ pos, int64_t np, int N, float *__restrict__ rho) {
  const float Nf = (float)N; const size_t N2 = (size_t)N * N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    int i, j, k; float wx[3], wy[3], wz[3];
    tsc_axis(pos[3*n] * Nf, i, wx[0], wx[1], wx[2]); tsc_axis(pos[3*n+1] * Nf, j, wy[0], wy[1], wy[2]); tsc_axis(pos[3*n+2] * Nf, k, wz[0], wz[1], wz[2]);
    int kk[3] = {wrapN(k-1,N), k, wrapN(k+1,N)};
    for (int a = 0; a < 3; a++) { size_t r = (size_t)wrapN(i+a-1,N) * N2;
      for (int b = 0; b < 3; b++) { size_t c = r + (size_t)wrapN(j+b-1,N) * N; float wxy = wx[a]*wy[b];
        for (int g = 0; g < 3; g++) atomicAdd(&rho[c + kk[g]], wxy * wz[g]); } }
  }
}

// var4: the three k-cells of a row as one RED.F32x4 (when k-1 is 0 or 1 mod 4 and no wrap) or two RED.F32x2
__global__ void k_global_vec(const float *__restrict__ pos, int64_t np, int N, float *__restrict__ rho) {
  const float Nf = (float)N; const size_t N2 = (size_t)N * N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    int i, j, k; float wx[3], wy[3], wz[3];
    tsc_axis(pos[3*n] * Nf, i, wx[0], wx[1], wx[2]); tsc_axis(pos[3*n+1] * Nf, j, wy[0], wy[1], wy[2]); tsc_axis(pos[3*n+2] * Nf, k, wz[0], wz[1], wz[2]);
    const int km = k - 1;
    const int m4 = km & 3;
    const bool interior = km >= 0 && k + 1 < N;
    for (int a = 0; a < 3; a++) { size_t r = (size_t)wrapN(i+a-1,N) * N2;
      for (int b = 0; b < 3; b++) { float *row = rho + r + (size_t)wrapN(j+b-1,N) * N; float wxy = wx[a]*wy[b];
        float v0 = wxy * wz[0], v1 = wxy * wz[1], v2 = wxy * wz[2];
        if (interior && m4 <= 1) {
          float4 v = m4 == 0 ? make_float4(v0, v1, v2, 0.f) : make_float4(0.f, v0, v1, v2);
          atomicAdd(reinterpret_cast<float4 *>(row + (km & ~3)), v);
        } else if (interior) {
          // m4 == 2: cells (km,km+1) | (km+2, pad) ; m4 == 3: (pad,km) | (km+1,km+2)
          if (m4 == 2) { atomicAdd(reinterpret_cast<float2 *>(row + km), make_float2(v0, v1)); atomicAdd(reinterpret_cast<float2 *>(row + km + 2), make_float2(v2, 0.f)); }
          else { atomicAdd(reinterpret_cast<float2 *>(row + km - 1), make_float2(0.f, v0)); atomicAdd(reinterpret_cast<float2 *>(row + km + 1), make_float2(v1, v2)); }
        } else {
          atomicAdd(row + wrapN(km, N), v0); atomicAdd(row + k, v1); atomicAdd(row + wrapN(k + 1, N), v2);
        }
      } }
  }
}

// var5: per-warp private tile.  WT = tile capacity in cells per warp.
template <int WT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_warp_tile(const float *__restrict__ pos, int64_t np, int N, float *__restrict__ rho,
                                                           unsigned long long *__restrict__ stats) {
  __shared__ __align__(16) float tiles[WARPS][WT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *tile = tiles[warp];
  const float Nf = (float)N; const size_t N2 = (size_t)N * N;
  const int64_t nwarps_total = (int64_t)gridDim.x * WARPS;
  const int64_t nchunks = (np + 31) / 32;
  unsigned long long nfall = 0, nflush = 0;
  for (int64_t chunk = (int64_t)blockIdx.x * WARPS + warp; chunk < nchunks; chunk += nwarps_total) {
    const int64_t n = chunk * 32 + lane;
    const bool valid = n < np;
    int i = 0, j = 0, k = 0; float wx[3], wy[3], wz[3];
    if (valid) { tsc_axis(pos[3*n] * Nf, i, wx[0], wx[1], wx[2]); tsc_axis(pos[3*n+1] * Nf, j, wy[0], wy[1], wy[2]); tsc_axis(pos[3*n+2] * Nf, k, wz[0], wz[1], wz[2]); }
    // frame anchored at lane 0's cell (relative coordinates in (-N/2, N/2])
    const int a0 = __shfl_sync(0xffffffffu, i, 0), a1 = __shfl_sync(0xffffffffu, j, 0), a2 = __shfl_sync(0xffffffffu, k, 0);
    int r0 = i - a0, r1 = j - a1, r2 = k - a2;
    r0 += (r0 < -N / 2) ? N : 0; r0 -= (r0 > N / 2) ? N : 0;
    r1 += (r1 < -N / 2) ? N : 0; r1 -= (r1 > N / 2) ? N : 0;
    r2 += (r2 < -N / 2) ? N : 0; r2 -= (r2 > N / 2) ? N : 0;
    const bool nearby = valid && abs(r0) < 32 && abs(r1) < 32 && abs(r2) < 64;
    int mn0 = nearby ? r0 : 0, mn1 = nearby ? r1 : 0, mn2 = nearby ? r2 : 0, mx0 = mn0, mx1 = mn1, mx2 = mn2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o)); mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
      mn2 = min(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx2 = max(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    }
    // tile origin (relative frame) incl. the -1 halo; k origin aligned to 4 in absolute coordinates
    const int o0 = mn0 - 1, o1 = mn1 - 1;
    int o2 = mn2 - 1; { int kabs = o2 + a2; o2 -= (kabs & 3); }
    int d0 = mx0 + 2 - o0, d1 = mx1 + 2 - o1, d2 = ((mx2 + 2 - o2) + 3) & ~3;
    // clip to capacity (uniform across the warp)
    while (d0 * d1 * d2 > WT) { if (d0 >= d1 && 4 * d0 >= d2) d0--; else if (4 * d1 >= d2) d1--; else d2 -= 4; }
    const int ncell = d0 * d1 * d2;
    for (int t = lane; t < ncell; t += 32) tile[t] = 0.0f;
    const int t0 = r0 - o0, t1 = r1 - o1, t2 = r2 - o2;
    const bool inside = nearby && t0 >= 1 && t0 + 1 < d0 && t1 >= 1 && t1 + 1 < d1 && t2 >= 1 && t2 + 1 < d2;
    // 27 weights
    float w[27];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) { float wxy = wx[a] * wy[b];
#pragma unroll
        for (int g = 0; g < 3; g++) w[(a * 3 + b) * 3 + g] = wxy * wz[g]; }
    // merge lanes that sit in the same cell: the lowest lane of each group accumulates its peers
    const int cellkey = inside ? (t0 * d1 + t1) * d2 + t2 : -1 - lane;
    unsigned peers = __match_any_sync(0xffffffffu, cellkey);
    const int leader = __ffs(peers) - 1;
    const bool is_leader = inside && leader == lane;
    unsigned rest = is_leader ? (peers & ~(1u << lane)) : 0u;
    while (__any_sync(0xffffffffu, rest != 0u)) {
      const int src = rest ? (__ffs(rest) - 1) : lane;
#pragma unroll
      for (int q = 0; q < 27; q++) { float v = __shfl_sync(0xffffffffu, w[q], src); if (rest) w[q] += v; }
      rest &= rest - 1;
    }
    __syncwarp();
    // 27 conflict-free phases: in one phase every leader adds to (its cell + the same offset)
    float *c0 = tile + ((t0 - 1) * d1 + (t1 - 1)) * d2 + (t2 - 1);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++)
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (is_leader) { float *p = c0 + (a * d1 + b) * d2 + g; *p += w[(a * 3 + b) * 3 + g]; }
          __syncwarp();
        }
    // stragglers outside the tile: direct global REDs
    if (valid && !inside) {
      nfall++;
      int kk[3] = {wrapN(k-1,N), k, wrapN(k+1,N)};
      for (int a = 0; a < 3; a++) { size_t r = (size_t)wrapN(i+a-1,N) * N2;
        for (int b = 0; b < 3; b++) { size_t c = r + (size_t)wrapN(j+b-1,N) * N;
          for (int g = 0; g < 3; g++) atomicAdd(&rho[c + kk[g]], w[(a * 3 + b) * 3 + g]); } }
    }
    // flush non-zero quads
    const int dq = d2 >> 2, nq = ncell >> 2;
    for (int t = lane; t < nq; t += 32) {
      float4 v = *reinterpret_cast<const float4 *>(tile + 4 * t);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        int kq = t % dq, r = t / dq; int b = r % d1, a = r / d1;
        int gi = a0 + o0 + a, gj = a1 + o1 + b, gk = a2 + o2 + 4 * kq;
        gi %= N; gi += gi < 0 ? N : 0; gj %= N; gj += gj < 0 ? N : 0; gk %= N; gk += gk < 0 ? N : 0;
        atomicAdd(reinterpret_cast<float4 *>(rho + (size_t)gi * N2 + (size_t)gj * N + gk), v);
        nflush++;
      }
    }
    __syncwarp();
  }
  if (stats) { if (nfall) atomicAdd(&stats[0], nfall); if (nflush) atomicAdd(&stats[1], nflush); }
}

int main(int argc, char **argv) {
  int N = argc > 1 ? atoi(argv[1]) : 256;
  float sigma = argc > 2 ? atof(argv[2]) : 0.3f;
  int lexi = argc > 3 ? atoi(argv[3]) : 0;
  int64_t np = (int64_t)N * N * N;
  printf("N=%d np=%lld sigma=%.2f cells, %s order\n", N, (long long)np, sigma, lexi ? "lexicographic" : "Morton");
  float *pos, *rho, *rho_ref; unsigned long long *stats;
  CK(cudaMalloc(&pos, sizeof(float) * 3 * np)); CK(cudaMalloc(&rho, sizeof(float) * np)); CK(cudaMalloc(&rho_ref, sizeof(float) * np));
  CK(cudaMalloc(&stats, 32));
  gen_particles<<<148 * 8, 256>>>(pos, np, N, sigma, lexi); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> href(np), hout(np);
  auto timeit = [&](auto launch, float *out, const char *name, bool ref) {
    float best = 1e30f;
    for (int r = 0; r < 4; r++) { CK(cudaMemset(out, 0, sizeof(float) * np)); CK(cudaMemset(stats, 0, 32)); cudaEventRecord(e0); launch(out); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
    CK(cudaGetLastError());
    unsigned long long st[4]; CK(cudaMemcpy(st, stats, 32, cudaMemcpyDeviceToHost));
    if (ref) { CK(cudaMemcpy(href.data(), out, sizeof(float) * np, cudaMemcpyDeviceToHost)); printf("%-28s: %8.3f ms  %6.3f ns/particle\n", name, best, best * 1e6 / np); }
    else { CK(cudaMemcpy(hout.data(), out, sizeof(float) * np, cudaMemcpyDeviceToHost));
      double maxd = 0, sum = 0; for (int64_t n = 0; n < np; n++) { maxd = std::max(maxd, (double)fabsf(hout[n] - href[n])); sum += hout[n]; }
      printf("%-28s: %8.3f ms  %6.3f ns/particle  maxdiff %.2e sum/np %.8f fallback %.3f%% flushx4/particle %.3f\n", name, best, best * 1e6 / np, maxd, sum / np,
             100.0 * st[0] / np, (double)st[1] / np); }
  };
  timeit([&](float *o) { k_global<<<148 * 16, 256>>>(pos, np, N, o); }, rho_ref, "var0 global RED x1", true);
  timeit([&](float *o) { k_global_vec<<<148 * 16, 256>>>(pos, np, N, o); }, rho, "var4 global RED x2/x4", false);
  timeit([&](float *o) { k_warp_tile<512, 8><<<148 * 8, 256>>>(pos, np, N, o, stats); }, rho, "var5 warp tile 512c 8w", false);
  timeit([&](float *o) { k_warp_tile<512, 16><<<148 * 4, 512>>>(pos, np, N, o, stats); }, rho, "var5 warp tile 512c 16w", false);
  timeit([&](float *o) { k_warp_tile<1024, 8><<<148 * 6, 256>>>(pos, np, N, o, stats); }, rho, "var5 warp tile 1024c 8w", false);
  timeit([&](float *o) { k_warp_tile<256, 16><<<148 * 4, 512>>>(pos, np, N, o, stats); }, rho, "var5 warp tile 256c 16w", false);
  return 0;
}
