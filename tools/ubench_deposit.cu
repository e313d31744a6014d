// tools/ubench_deposit.cu -- micro-benchmark that decided the design of the tiled TSC deposit.
// Variants of per-CTA shared-memory tile accumulation over chunks of spatially sorted particles:
//   0  global RED.F32 per particle (baseline, any order)
//   1  smem float atomicAdd (CAS loop)
//   2  smem u64 fixed-point atomicAdd (Q24.40, order independent)
//   3  smem u32 fixed-point atomicAdd (Q8.24, speed reference only)
// All tile variants flush non-zero tile cells with global float REDs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_deposit tools/ubench_deposit.cu
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

// ---- tile kernel --------------------------------------------------------------------------------
// CTA = P consecutive particles.  Frame: anchor A = cell(first particle) - N/2 per axis; relative cell
// r = (c - A) mod N.  Block-min of r gives the tile origin O = rmin - 1 (TSC halo), k-origin rounded
// down to a multiple of 4.  Tile dims (TX,TY,TZ) are chosen per CTA from the block max, clipped to the
// smem budget; particles whose 3^3 footprint leaves the tile fall back to global REDs.
template <int VAR> struct Acc;
template <> struct Acc<1> { typedef float T; static __device__ __forceinline__ void add(T *p, float w) { atomicAdd(p, w); }
  static __device__ __forceinline__ float get(T v) { return v; } };
template <> struct Acc<2> { typedef unsigned long long T; static __device__ __forceinline__ void add(T *p, float w) {
    atomicAdd(p, (unsigned long long)__float2ull_rn(w * 1099511627776.0f)); }
  static __device__ __forceinline__ float get(T v) { return (float)((double)v * (1.0 / 1099511627776.0)); } };
template <> struct Acc<3> { typedef unsigned int T; static __device__ __forceinline__ void add(T *p, float w) {
    atomicAdd(p, __float2uint_rn(w * 16777216.0f)); }
  static __device__ __forceinline__ float get(T v) { return (float)v * (1.0f / 16777216.0f); } };

template <int VAR, int P, int NT>
__global__ void __launch_bounds__(NT) k_tile(const float *__restrict__ pos, int64_t np, int N, float *__restrict__ rho,
                                             int max_cells, unsigned long long *__restrict__ stats) {
  typedef typename Acc<VAR>::T T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *tile = reinterpret_cast<T *>(smem_raw);
  __shared__ int s_min[3], s_max[3], s_dim[3], s_org[3];
  const float Nf = (float)N; const size_t N2 = (size_t)N * N;
  const int64_t base = (int64_t)blockIdx.x * P;
  const int cnt = (int)min((int64_t)P, np - base);
  constexpr int PPT = P / NT;
  // anchor from the first particle of the chunk
  int A[3];
  { float x0 = pos[3*base] * Nf, y0 = pos[3*base+1] * Nf, z0 = pos[3*base+2] * Nf;
    A[0] = (int)x0 - N/2; A[1] = (int)y0 - N/2; A[2] = (int)z0 - N/2; }
  if (threadIdx.x < 3) { s_min[threadIdx.x] = 1 << 30; s_max[threadIdx.x] = -1; }
  __syncthreads();
  float px[PPT], py[PPT], pz[PPT];
  int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-1, -1, -1};
#pragma unroll
  for (int q = 0; q < PPT; q++) {
    int idx = q * NT + threadIdx.x;
    if (idx < cnt) {
      px[q] = pos[3*(base+idx)] * Nf; py[q] = pos[3*(base+idx)+1] * Nf; pz[q] = pos[3*(base+idx)+2] * Nf;
      int r0 = wrapN((int)px[q] - A[0], N), r1 = wrapN((int)py[q] - A[1], N), r2 = wrapN((int)pz[q] - A[2], N);
      // ignore far outliers when sizing the tile (they take the fallback path)
      if (abs(r0 - N/2) <= 48 && abs(r1 - N/2) <= 48 && abs(r2 - N/2) <= 48) {
        mn[0] = min(mn[0], r0); mn[1] = min(mn[1], r1); mn[2] = min(mn[2], r2);
        mx[0] = max(mx[0], r0); mx[1] = max(mx[1], r1); mx[2] = max(mx[2], r2);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    for (int o = 16; o > 0; o >>= 1) { mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_min[d], mn[d]); atomicMax(&s_max[d], mx[d]); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o0 = s_min[0] - 1, o1 = s_min[1] - 1, o2 = (s_min[2] - 1);
    // align k origin to a multiple of 4 in ABSOLUTE coordinates
    int k_abs = o2 + A[2]; int k_al = k_abs & ~3; o2 -= (k_abs - k_al);
    int d0 = s_max[0] + 2 - o0, d1 = s_max[1] + 2 - o1, d2 = ((s_max[2] + 2 - o2) + 3) & ~3;
    // clip to the smem budget: shrink the largest dimension until it fits
    while ((long long)d0 * d1 * d2 > max_cells) { if (d0 >= d1 && d0 * 4 >= d2) d0--; else if (d1 * 4 >= d2) d1--; else d2 -= 4; }
    s_org[0] = o0; s_org[1] = o1; s_org[2] = o2; s_dim[0] = d0; s_dim[1] = d1; s_dim[2] = d2;
  }
  __syncthreads();
  const int O0 = s_org[0], O1 = s_org[1], O2 = s_org[2], D0 = s_dim[0], D1 = s_dim[1], D2 = s_dim[2];
  const int ncell = D0 * D1 * D2;
  for (int t = threadIdx.x; t < ncell; t += NT) tile[t] = 0;
  __syncthreads();
  int nfall = 0;
#pragma unroll
  for (int q = 0; q < PPT; q++) {
    int idx = q * NT + threadIdx.x;
    if (idx < cnt) {
      int i, j, k; float wx[3], wy[3], wz[3];
      tsc_axis(px[q], i, wx[0], wx[1], wx[2]); tsc_axis(py[q], j, wy[0], wy[1], wy[2]); tsc_axis(pz[q], k, wz[0], wz[1], wz[2]);
      int t0 = wrapN(i - A[0], N) - O0, t1 = wrapN(j - A[1], N) - O1, t2 = wrapN(k - A[2], N) - O2;
      if (t0 >= 1 && t0 + 1 < D0 && t1 >= 1 && t1 + 1 < D1 && t2 >= 1 && t2 + 1 < D2) {
        T *c0 = tile + ((t0 - 1) * D1 + (t1 - 1)) * D2 + (t2 - 1);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) { float wxy = wx[a] * wy[b]; T *r = c0 + (a * D1 + b) * D2;
            Acc<VAR>::add(r, wxy * wz[0]); Acc<VAR>::add(r + 1, wxy * wz[1]); Acc<VAR>::add(r + 2, wxy * wz[2]); }
      } else {
        nfall++;
        int kk[3] = {wrapN(k-1,N), k, wrapN(k+1,N)};
        for (int a = 0; a < 3; a++) { size_t r = (size_t)wrapN(i+a-1,N) * N2;
          for (int b = 0; b < 3; b++) { size_t c = r + (size_t)wrapN(j+b-1,N) * N; float wxy = wx[a]*wy[b];
            for (int g = 0; g < 3; g++) atomicAdd(&rho[c + kk[g]], wxy * wz[g]); } }
      }
    }
  }
  __syncthreads();
  // flush: one thread per 4 consecutive k cells (16-byte aligned in global memory)
  int nflush = 0;
  const int nq = ncell >> 2, DQ = D2 >> 2;
  for (int t = threadIdx.x; t < nq; t += NT) {
    int kq = t % DQ; int r = t / DQ; int b = r % D1, a = r / D1;
    const T *s = tile + (size_t)t * 4;
    float4 v = make_float4(Acc<VAR>::get(s[0]), Acc<VAR>::get(s[1]), Acc<VAR>::get(s[2]), Acc<VAR>::get(s[3]));
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
      int gi = wrapN(wrapN(O0 + a + A[0], N), N), gj = wrapN(wrapN(O1 + b + A[1], N), N);
      int gk = O2 + 4 * kq + A[2]; gk %= N; if (gk < 0) gk += N;
      gi %= N; if (gi < 0) gi += N; gj %= N; if (gj < 0) gj += N;
      atomicAdd(reinterpret_cast<float4 *>(rho + (size_t)gi * N2 + (size_t)gj * N + gk), v);
      nflush++;
    }
  }
  if (stats) { if (nfall) atomicAdd(&stats[0], (unsigned long long)nfall); if (nflush) atomicAdd(&stats[1], (unsigned long long)nflush);
    if (threadIdx.x == 0) atomicAdd(&stats[2], (unsigned long long)ncell); }
}

// ---- host ---------------------------------------------------------------------------------------
static uint64_t spread(uint64_t x) { x &= 0x1FFFFF; x = (x | x << 32) & 0x1F00000000FFFFull; x = (x | x << 16) & 0x1F0000FF0000FFull;
  x = (x | x << 8) & 0x100F00F00F00F00Full; x = (x | x << 4) & 0x10C30C30C30C30C3ull; x = (x | x << 2) & 0x1249249249249249ull; return x; }

template <int VAR, int P, int NT>
float run_tile(const float *pos, int64_t np, int N, float *rho, int max_cells, unsigned long long *stats, int reps) {
  size_t smem = (size_t)max_cells * sizeof(typename Acc<VAR>::T);
  CK(cudaFuncSetAttribute(k_tile<VAR, P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((np + P - 1) / P);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaMemset(rho, 0, sizeof(float) * (size_t)N * N * N));
    CK(cudaMemset(stats, 0, 32));
    cudaEventRecord(e0);
    k_tile<VAR, P, NT><<<grid, NT, smem>>>(pos, np, N, rho, max_cells, stats);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char **argv) {
  int N = argc > 1 ? atoi(argv[1]) : 256;
  float sigma = argc > 2 ? atof(argv[2]) : 0.3f;
  int64_t np = (int64_t)N * N * N;
  printf("N=%d np=%lld sigma=%.2f cells, Morton-sorted\n", N, (long long)np, sigma);
  std::vector<float> h(3 * np);
  std::vector<std::pair<uint64_t, uint32_t>> keys(np);
  srand(42);
  auto gauss = [&]() { float u = (rand() + 1.0f) / (RAND_MAX + 2.0f), v = (rand() + 1.0f) / (RAND_MAX + 2.0f); return sqrtf(-2 * logf(u)) * cosf(6.2831853f * v); };
  for (int64_t n = 0; n < np; n++) {
    int k = n % N, j = (n / N) % N, i = n / ((int64_t)N * N);
    float x = (i + 0.5f + sigma * gauss()) / N, y = (j + 0.5f + sigma * gauss()) / N, z = (k + 0.5f + sigma * gauss()) / N;
    x -= floorf(x); y -= floorf(y); z -= floorf(z); if (x >= 1) x = 0; if (y >= 1) y = 0; if (z >= 1) z = 0;
    h[3*n] = x; h[3*n+1] = y; h[3*n+2] = z;
    keys[n] = { spread((uint64_t)(x * 2097152.0f)) << 2 | spread((uint64_t)(y * 2097152.0f)) << 1 | spread((uint64_t)(z * 2097152.0f)), (uint32_t)n };
  }
  std::sort(keys.begin(), keys.end());
  std::vector<float> hs(3 * np);
  for (int64_t n = 0; n < np; n++) { uint32_t s = keys[n].second; hs[3*n] = h[3*s]; hs[3*n+1] = h[3*s+1]; hs[3*n+2] = h[3*s+2]; }
  float *pos, *rho, *rho_ref; unsigned long long *stats;
  CK(cudaMalloc(&pos, sizeof(float) * 3 * np)); CK(cudaMalloc(&rho, sizeof(float) * np)); CK(cudaMalloc(&rho_ref, sizeof(float) * np));
  CK(cudaMalloc(&stats, 32));
  CK(cudaMemcpy(pos, hs.data(), sizeof(float) * 3 * np, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { CK(cudaMemset(rho_ref, 0, sizeof(float) * np)); cudaEventRecord(e0);
    k_global<<<148 * 16, 256>>>(pos, np, N, rho_ref); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
  printf("var0 global RED          : %8.3f ms  %6.2f ns/particle\n", best, best * 1e6 / np);
  std::vector<float> href(np), hout(np);
  CK(cudaMemcpy(href.data(), rho_ref, sizeof(float) * np, cudaMemcpyDeviceToHost));
  auto report = [&](const char *name, float ms) {
    CK(cudaMemcpy(hout.data(), rho, sizeof(float) * np, cudaMemcpyDeviceToHost));
    unsigned long long st[4]; CK(cudaMemcpy(st, stats, 32, cudaMemcpyDeviceToHost));
    double maxd = 0, sum = 0; for (int64_t n = 0; n < np; n++) { maxd = std::max(maxd, (double)fabsf(hout[n] - href[n])); sum += hout[n]; }
    printf("%-25s: %8.3f ms  %6.2f ns/particle  maxdiff %.2e sum/np %.8f  fallback %.3f%% flushx4/particle %.3f tilecells/particle %.2f\n",
           name, ms, ms * 1e6 / np, maxd, sum / np, 100.0 * st[0] / np, (double)st[1] / np, (double)st[2] / np);
  };
  report("var1 f32 CAS P2048 T512 28K", run_tile<1, 2048, 512>(pos, np, N, rho, 28 * 1024, stats, 3));
  report("var1 f32 CAS P2048 T512 14K", run_tile<1, 2048, 512>(pos, np, N, rho, 14 * 1024, stats, 3));
  report("var2 u64 fix P2048 T512 14K", run_tile<2, 2048, 512>(pos, np, N, rho, 14 * 1024, stats, 3));
  report("var2 u64 fix P2048 T512 24K", run_tile<2, 2048, 512>(pos, np, N, rho, 24 * 1024, stats, 3));
  report("var3 u32 fix P2048 T512 28K", run_tile<3, 2048, 512>(pos, np, N, rho, 28 * 1024, stats, 3));
  report("var3 u32 fix P2048 T512 14K", run_tile<3, 2048, 512>(pos, np, N, rho, 14 * 1024, stats, 3));
  report("var3 u32 fix P1024 T256 14K", run_tile<3, 1024, 256>(pos, np, N, rho, 14 * 1024, stats, 3));
  report("var3 u32 fix P4096 T1024 28K", run_tile<3, 4096, 1024>(pos, np, N, rho, 28 * 1024, stats, 3));
  report("var1 f32 CAS P4096 T1024 28K", run_tile<1, 4096, 1024>(pos, np, N, rho, 28 * 1024, stats, 3));
  report("var2 u64 fix P4096 T1024 24K", run_tile<2, 4096, 1024>(pos, np, N, rho, 24 * 1024, stats, 3));
  return 0;
}
