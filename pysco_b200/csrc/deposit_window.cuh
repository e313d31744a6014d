// deposit_window.cuh -- warp-private "walking window" mass deposition (TSC / CIC / NGP).
//
// Why this shape (measured on B200, tools/ubench_deposit*.cu, profiles/r01_ubench_deposit.txt):
//   * shared-memory atomics (float CAS, u32/u64 add alike) run at ~2 cycles/lane  -> 3.3 ms @256^3,
//     slower than L2 float REDs (~380 G float-adds/s -> 1.2 ms @256^3, 9.5 ms @512^3);
//   * RED.F32x2/x4 do not raise the L2 rate: the L2 atomic throughput is per float;
//   => the only lever is to ADD FEWER FLOATS to global memory, and to accumulate on chip without
//      atomics.
//
// Each warp walks a contiguous range of the (spatially coherent, e.g. Morton-ordered) particle array
// in chunks of 32.  It owns a private window of the density grid in shared memory.  Per chunk:
//   1. the 3x3x3 weights of the lane's particle are formed in registers;
//   2. lanes sitting in the same cell are merged (match_any + shuffles) so that leaders have distinct cells;
//   3. 27 phases: in phase (a,b,c) every leader does a plain LDS/FADD/STS on window[cell + (a,b,c)].
//      Distinct cells + identical offset => distinct addresses, so no atomics are needed; phases are
//      ordered by __syncwarp();
//   4. when a chunk leaves the window, the window's non-zero quads are flushed with RED.F32x4 and
//      the window is re-anchored around the chunk.  Particles that do not fit (incoherent input)
//      fall back to direct global REDs, so any particle order is handled correctly.
// Net effect on sorted input: ~2-3 floats RED'ed per particle instead of 27.
#pragma once
#include "common.cuh"

namespace psc {

constexpr int DW_WARPS = 7;                        // warps per CTA
constexpr int DW_D0 = 12, DW_D1 = 12, DW_D2 = 16;  // window dims per warp: an 8^3 Morton block + 2 cells each side
// Shared-memory pitches chosen so that the 32 leaders of a 4x4x2 Morton chunk hit 32 distinct banks:
// row pitch 18 (t1*18 mod 32 = 0,18,4,22), plane pitch 216 (t0*216 mod 32 = 0,24,16,8), t2 in {0,1}.
// (ncu, first version with pitches 16/192: 52% of the shared wavefronts were bank conflicts and the
// shared pipe sat at 90% of peak -- profiles/r01_deposit_window_v1_ncu.txt.)
constexpr int DW_P1 = 18, DW_P0 = DW_D1 * DW_P1;   // 216
constexpr int DW_CAP = DW_D0 * DW_P0;              // 2592 floats = 10.1 KB per warp, 70.9 KB per CTA
constexpr int DW_CTAS_PER_SM = 3;

template <int SCHEME>
__device__ __forceinline__ void axis_weights(float xp, int N, int &c, float &wm, float &w0, float &wp) {
  if (SCHEME == PSC_TSC) {
    tsc_axis(xp, c, wm, w0, wp);
  } else if (SCHEME == PSC_CIC) {
    // CIC as a 3-point stencil with one zero weight (mesh.py:2318-2345): second cell is c + sign(d)
    c = (int)xp;
    float d = xp - 0.5f - (float)c;
    float ad = fabsf(d);
    w0 = 1.0f - ad;
    wm = d < 0.0f ? ad : 0.0f;
    wp = d > 0.0f ? ad : 0.0f;
  } else {
    c = (int)xp;
    wm = 0.0f; w0 = 1.0f; wp = 0.0f;
  }
}

__device__ __forceinline__ int rel_coord(int c, int anchor, int N) {
  int r = c - anchor;
  r += (r < -(N >> 1)) ? N : 0;
  r -= (r > (N >> 1)) ? N : 0;
  return r;
}
__device__ __forceinline__ int mod_pos(int v, int N) {
  v %= N;
  return v < 0 ? v + N : v;
}

struct DepositStats {
  unsigned long long fallback, flushed_quads, reanchors;
};

// window state (uniform across the warp): absolute cell coordinates of window cell (0,0,0)
struct Window {
  int o0, o1, o2;
  bool live;
};

__device__ __forceinline__ void window_flush(float *win, const Window &w, int N, float *__restrict__ rho, int lane,
                                             unsigned long long &nflush) {
  const size_t N2 = (size_t)N * N;
  constexpr int DQ = DW_D2 / 4, NQ = DW_D0 * DW_D1 * DQ;
#pragma unroll 2
  for (int t = lane; t < NQ; t += 32) {
    const int kq = t % DQ, r = t / DQ;
    const int b = r % DW_D1, a = r / DW_D1;
    float *q = win + a * DW_P0 + b * DW_P1 + 4 * kq;
    const float2 lo = *reinterpret_cast<const float2 *>(q), hi = *reinterpret_cast<const float2 *>(q + 2);
    if (lo.x != 0.f || lo.y != 0.f || hi.x != 0.f || hi.y != 0.f) {
      const int gi = mod_pos(w.o0 + a, N), gj = mod_pos(w.o1 + b, N), gk = mod_pos(w.o2 + 4 * kq, N);
      atomicAdd(reinterpret_cast<float4 *>(rho + (size_t)gi * N2 + (size_t)gj * N + gk),
                make_float4(lo.x, lo.y, hi.x, hi.y));
      *reinterpret_cast<float2 *>(q) = make_float2(0.f, 0.f);
      *reinterpret_cast<float2 *>(q + 2) = make_float2(0.f, 0.f);
      nflush++;
    }
  }
  __syncwarp();
}

template <int SCHEME, int DBG = 0>
__global__ void __launch_bounds__(DW_WARPS * 32, DW_CTAS_PER_SM) deposit_window_kernel(const float *__restrict__ pos, int64_t np,
                                                                      int N, float *__restrict__ rho,
                                                                      int64_t chunks_per_warp,
                                                                      DepositStats *__restrict__ stats) {
  extern __shared__ __align__(16) float dw_windows[];  // [DW_WARPS][DW_CAP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *win = dw_windows + warp * DW_CAP;
  const float Nf = (float)N;
  const size_t N2 = (size_t)N * N;
  const int64_t gw = (int64_t)blockIdx.x * DW_WARPS + warp;
  const int64_t nchunks = (np + 31) >> 5;
  const int64_t c_begin = gw * chunks_per_warp;
  const int64_t c_end = min(nchunks, c_begin + chunks_per_warp);
  unsigned long long nfall = 0, nflush = 0, nre = 0;

  for (int t = lane; t < DW_CAP; t += 32) win[t] = 0.0f;
  __syncwarp();
  Window w;
  w.o0 = w.o1 = w.o2 = 0;
  w.live = false;

  // software prefetch of the next chunk's positions
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if (c_begin < c_end) {
    int64_t n = c_begin * 32 + lane;
    if (n < np) { nx = __ldg(&pos[3 * n]); ny = __ldg(&pos[3 * n + 1]); nz = __ldg(&pos[3 * n + 2]); }
  }
  for (int64_t chunk = c_begin; chunk < c_end; chunk++) {
    const int64_t n = chunk * 32 + lane;
    const bool valid = n < np;
    const float px = nx, py = ny, pz = nz;
    if (chunk + 1 < c_end) {
      int64_t m = n + 32;
      if (m < np) { nx = __ldg(&pos[3 * m]); ny = __ldg(&pos[3 * m + 1]); nz = __ldg(&pos[3 * m + 2]); }
    }
    int i = 0, j = 0, k = 0;
    float wx[3], wy[3], wz[3];
    axis_weights<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);

    // chunk bounding box in a frame anchored at lane 0's cell (relative coords in (-N/2, N/2])
    const int a0 = __shfl_sync(0xffffffffu, i, 0), a1 = __shfl_sync(0xffffffffu, j, 0),
              a2 = __shfl_sync(0xffffffffu, k, 0);
    const int r0 = rel_coord(i, a0, N), r1 = rel_coord(j, a1, N), r2 = rel_coord(k, a2, N);
    const bool nearby = valid && abs(r0) < DW_D0 && abs(r1) < DW_D1 && abs(r2) < DW_D2;
    // warp-wide bounding box with REDUX (one instruction per bound instead of 5 shuffles)
    const int mn0 = __reduce_min_sync(0xffffffffu, nearby ? r0 : 0), mx0 = __reduce_max_sync(0xffffffffu, nearby ? r0 : 0);
    const int mn1 = __reduce_min_sync(0xffffffffu, nearby ? r1 : 0), mx1 = __reduce_max_sync(0xffffffffu, nearby ? r1 : 0);
    const int mn2 = __reduce_min_sync(0xffffffffu, nearby ? r2 : 0), mx2 = __reduce_max_sync(0xffffffffu, nearby ? r2 : 0);
    // does the chunk's footprint (bbox +- 1) fit the current window?
    bool fits = false;
    if (w.live) {
      int lo0 = rel_coord(a0 + mn0 - 1, w.o0, N), lo1 = rel_coord(a1 + mn1 - 1, w.o1, N),
          lo2 = rel_coord(a2 + mn2 - 1, w.o2, N);
      fits = lo0 >= 0 && lo0 + (mx0 - mn0) + 2 < DW_D0 && lo1 >= 0 && lo1 + (mx1 - mn1) + 2 < DW_D1 &&
             lo2 >= 0 && lo2 + (mx2 - mn2) + 2 < DW_D2;
    }
    if (!fits) {
      if (w.live && !(DBG & 4)) window_flush(win, w, N, rho, lane, nflush);
      // re-anchor.  Preferred: the 8^3-aligned block holding the chunk's low corner, minus 2 cells
      // (Morton order visits such a block as 16 consecutive chunks; the margin absorbs the TSC halo
      // and one cell of drift), k origin lowered to a multiple of 4 for 16-byte REDs.  If the chunk does
      // not fit that way, centre it instead.
      const int c0 = mod_pos(a0 + mn0, N), c1 = mod_pos(a1 + mn1, N), c2 = mod_pos(a2 + mn2, N);
      int q0 = (c0 & ~7) - 2, q1 = (c1 & ~7) - 2, q2 = (c2 & ~7) - 4;
      const int e0 = mx0 - mn0 + 3, e1 = mx1 - mn1 + 3, e2 = mx2 - mn2 + 3;  // footprint extents
      const bool ok = (c0 - 1 - q0) + e0 <= DW_D0 && (c1 - 1 - q1) + e1 <= DW_D1 && (c2 - 1 - q2) + e2 <= DW_D2;
      if (!ok) {
        q0 = c0 - 1 - max(0, (DW_D0 - e0) >> 1);
        q1 = c1 - 1 - max(0, (DW_D1 - e1) >> 1);
        q2 = c2 - 1 - max(0, (DW_D2 - e2 - 3) >> 1);
        q2 -= mod_pos(q2, N) & 3;
      }
      w.o0 = mod_pos(q0, N);
      w.o1 = mod_pos(q1, N);
      w.o2 = mod_pos(q2, N);
      w.live = true;
      nre++;
    }
    // window-relative cell of this lane's particle
    const int t0 = rel_coord(i, w.o0, N), t1 = rel_coord(j, w.o1, N), t2 = rel_coord(k, w.o2, N);
    const bool inside = valid && t0 >= 1 && t0 + 1 < DW_D0 && t1 >= 1 && t1 + 1 < DW_D1 && t2 >= 1 && t2 + 1 < DW_D2;

    float wgt[27];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) {
        const float wxy = wx[a] * wy[b];
#pragma unroll
        for (int g = 0; g < 3; g++) wgt[(a * 3 + b) * 3 + g] = wxy * wz[g];
      }

    // merge lanes of equal cell into the lowest lane of the group
    const int cellkey = inside ? t0 * DW_P0 + t1 * DW_P1 + t2 : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, cellkey);
    const bool is_leader = inside && (__ffs(peers) - 1) == lane;
    unsigned rest = is_leader ? (peers & ~(1u << lane)) : 0u;
    while (!(DBG & 2) && __any_sync(0xffffffffu, rest != 0u)) {
      const int src = rest ? (__ffs(rest) - 1) : lane;
#pragma unroll
      for (int q = 0; q < 27; q++) {
        float v = __shfl_sync(0xffffffffu, wgt[q], src);
        if (rest) wgt[q] += v;
      }
      rest &= rest - 1;
    }

    // 27 conflict-free phases (plain read-modify-write, ordered by __syncwarp)
    float *cell0 = win + (t0 - 1) * DW_P0 + (t1 - 1) * DW_P1 + (t2 - 1);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++)
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (SCHEME == PSC_NGP && !(a == 1 && b == 1 && g == 1)) continue;
          if (is_leader) {
            float *p = cell0 + a * DW_P0 + b * DW_P1 + g;
            *p += wgt[(a * 3 + b) * 3 + g];
          }
          if (!(DBG & 1)) __syncwarp();
        }

    // particles that do not fit the window: direct global REDs (keeps any input order correct)
    if (valid && !inside) {
      nfall++;
      const int kk[3] = {wrap(k - 1, N), k, wrap(k + 1, N)};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const size_t r = (size_t)wrap(i + a - 1, N) * N2;
#pragma unroll
        for (int b = 0; b < 3; b++) {
          const size_t c = r + (size_t)wrap(j + b - 1, N) * N;
#pragma unroll
          for (int g = 0; g < 3; g++) {
            const float v = wgt[(a * 3 + b) * 3 + g];
            if (SCHEME == PSC_TSC || v != 0.0f) atomicAdd(&rho[c + kk[g]], v);
          }
        }
      }
    }
  }
  if (w.live) window_flush(win, w, N, rho, lane, nflush);
  if (stats) {
    if (nfall) atomicAdd(&stats->fallback, nfall);
    if (nflush) atomicAdd(&stats->flushed_quads, nflush);
    if (lane == 0 && nre) atomicAdd(&stats->reanchors, nre);
  }
}

}  // namespace psc
