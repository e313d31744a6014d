// binned.cu -- order-independent particle <-> mesh kernels on a per-step "shadow" counting sort of the particles.
//
// The particle arrays keep the reference's order (bit-exact ordering parity), so every step the positions are
// COPIED into CELL order: bins of 8^3 cells, and inside a bin the cells in (i, j, k) order, k fastest.
//
//   psc_bin_particles   count per cell (warp-aggregated int atomics; fused into the kick-drift-wrap in the step)
//                       -> in-place exclusive scan (CUB) -> scatter of (x, y, z, source row).  After the scatter
//                       cell[c] / cell[c + 1] are the first / one-past-last record of cell c, and cell[512 b] the
//                       first record of bin b.
//   psc_deposit_binned  one CTA per bin, LANES OWN CELLS (round 2): warp w owns plane w of the bin, a lane one
//                       (j, k) column of it.  A lane sums the 27 weights of the particles of its cell in
//                       registers, then the warp adds them to its private 3-plane tile in 27 phases of plain
//                       LDS/FADD/STS: the 32 lanes are 4 rows x 8 consecutive cells and the row pitch is 24, so
//                       every phase touches 32 distinct banks -- no merging of equal cells, no conflicts, one
//                       read-modify-write per CELL instead of per particle.  The eight tiles are summed; the 6^3
//                       cells no other bin can reach are stored plainly, the shell goes to L2 as float REDs.
//                       Bins whose fill is too uneven for a lane-per-cell loop (one cell far above the mean) and
//                       bins of more than BIN_PART particles take the round-1 path (lanes own particles, equal
//                       cells merged by match_any + shuffles).
//   psc_interp_kick*_binned  one CTA per bin: the 10^3 float4 force tile is staged once in shared memory
//                       (16 KB), every particle gathers its 27 points with LDS.128, the result is written to
//                       the particle's ORIGINAL row (acceleration, velocity kick) through the source index.  The
//                       records arrive cell-sorted, so the lanes of a quarter-warp read (mostly) consecutive
//                       float4 of one tile row.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace psc {

constexpr int BB = 8;           // bin edge in cells
constexpr int BT = BB + 2;      // tile edge (one halo cell per side)
constexpr int BD_WARPS = 4;     // warps per CTA in the binned deposit
constexpr int BD_P1 = 12;       // row pitch of the per-warp tile; 12 / 144 spreads a Morton chunk (2x4x4 cells) over the banks
constexpr int BD_P0 = 144;
constexpr int BD_TILE = BT * BD_P0;  // 1100 floats per warp

// A bin is processed by one CTA up to BIN_PART particles; the rest of a heavier bin (a halo core can hold 10^6
// particles of a 512^3 run at z = 0) is cut into parts of BIN_PART particles listed in BinLayout::heavy and processed
// by a second, persistent launch -- otherwise one CTA would serialise the whole bin.
constexpr int BIN_PART = 4096;

constexpr int CPB = BB * BB * BB;   // cells per bin

struct BinLayout {
  int NB;            // bins along y and z (N / 8)
  int NBX;           // bins along x (owned planes / 8; == NB for the periodic single-domain case)
  int x0;            // first owned x cell (0 unless the mesh is slab-decomposed)
  int64_t nbins;     // NBX * NB^2
  int64_t ncells;    // 512 nbins
  // [ncells + 2]  cell[0] = 0.  cell[1 + c] holds the count of cell c after the count pass, its first record after
  // the scan, and its one-past-last record after the scatter (every particle claims its slot by an atomic +1).
  // So, once the binning is complete: records of cell c = [cell[c], cell[c + 1]), of bin b = [cell[512 b], cell[512 (b + 1)]).
  int *cell;
  float4 *rec;       // [np] sorted particles: (x, y, z, source row as int bits) -- one 16-byte access per particle
  int *dense_count;  // [1] number of entries of `dense`
  int *dense;        // [nbins] bins the lane-per-cell deposit handed to the lane-per-particle path
  int *heavy_count;  // [1] number of entries of `heavy`
  int2 *heavy;       // [heavy_cap] (bin, part >= 1): the parts beyond the first BIN_PART particles of a bin
  int heavy_cap;
  void *cub_tmp;
  size_t cub_bytes;
};

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t scan_tmp_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int *)nullptr, (int *)nullptr, (int)n);
  return b;
}

static bool bin_layout(void *scratch, size_t bytes, int64_t np, int N, int x0, int nxl, BinLayout &L) {
  L.NB = N / BB;
  L.NBX = nxl / BB;
  L.x0 = x0;
  L.nbins = (int64_t)L.NBX * L.NB * L.NB;
  L.ncells = L.nbins * CPB;
  char *p = reinterpret_cast<char *>(scratch);
  size_t off = 0;
  L.cell = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (size_t)(L.ncells + 2));
  L.rec = reinterpret_cast<float4 *>(p + off); off += a256(sizeof(float4) * (size_t)np);
  L.dense_count = reinterpret_cast<int *>(p + off); off += 128;
  L.heavy_count = reinterpret_cast<int *>(p + off); off += 128;
  L.dense = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (size_t)L.nbins);
  L.heavy_cap = (int)(np / BIN_PART) + 1;
  L.heavy = reinterpret_cast<int2 *>(p + off); off += a256(sizeof(int2) * (size_t)L.heavy_cap);
  L.cub_tmp = p + off;
  L.cub_bytes = scan_tmp_bytes(L.ncells + 1);
  off += a256(L.cub_bytes);
  return off <= bytes;
}

// Sort key of a position: 512 * bin + 64 (i & 7) + 8 (j & 7) + (k & 7), bin = (bi * NB + bj) * NB + bk.
// x0 / NBX: the slab of owned planes [x0, x0 + 8 NBX); a particle outside it (the host migrates particles before
// binning) is clamped into the edge plane, and y / z outside [0, 1) (an exact 1.0 from an external snapshot, an inf)
// into the edge cells, so that nothing is ever written out of bounds.
__device__ __forceinline__ int cell_of(float x, float y, float z, float Nf, int N, int NB, int x0, int NBX) {
  int i = (int)(x * Nf) - x0, j = (int)(y * Nf), k = (int)(z * Nf);
  i = min(max(i, 0), BB * NBX - 1);
  j = min(max(j, 0), N - 1);
  k = min(max(k, 0), N - 1);
  const int b = ((i >> 3) * NB + (j >> 3)) * NB + (k >> 3);
  return (b << 9) | ((i & 7) << 6) | ((j & 7) << 3) | (k & 7);
}

// pass 1: cnt[cell] += 1 (cnt = BinLayout::cell + 1), one atomic per distinct cell per warp
__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ pos, int64_t np, int N, int NB,
                                                        int x0, int NBX, int *__restrict__ cnt) {
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    int c = -1 - lane;
    if (n < np) c = cell_of(__ldg(&pos[3 * n]), __ldg(&pos[3 * n + 1]), __ldg(&pos[3 * n + 2]), Nf, N, NB, x0, NBX);
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&cnt[c], __popc(peers));
  }
}

// pass 1 fused into the first half of the leapfrog step (integration.py:250-258): v -= half_dt a; x += dt v; wrap(x);
// cnt[cell(x)] += 1.  Four particles (three float4 per array) per thread.  Saves the separate read of the positions
// that bin_count_kernel does.  The particles of one cell sit next to each other in a Morton-ordered array, so equal
// keys of a thread's four particles are merged first and the rest across the warp: a dense cell (a halo core) costs
// one atomic per warp, not one per particle.
template <bool F64>
__global__ void __launch_bounds__(256) kick_drift_wrap_count_kernel(float *__restrict__ pos, float *__restrict__ vel,
                                                                    const float *__restrict__ acc, int64_t np,
                                                                    float half_dt, double dt, int N, int NB,
                                                                    int *__restrict__ cnt) {
  const float dtf = (float)dt, mh = -half_dt, Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nq = np >> 2;
  float4 *p4 = reinterpret_cast<float4 *>(pos);
  float4 *v4 = reinterpret_cast<float4 *>(vel);
  const float4 *a4 = reinterpret_cast<const float4 *>(acc);
  const int64_t wstride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < nq; base += wstride) {
    const int64_t q = base + lane;
    const bool valid = q < nq;
    float f[12];
    int c[4] = {-1 - lane, -1 - lane, -1 - lane, -1 - lane};
    if (valid) {
      float v[12], a[12];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const float4 P = p4[3 * q + d], V = v4[3 * q + d], A = __ldg(&a4[3 * q + d]);
        f[4 * d] = P.x; f[4 * d + 1] = P.y; f[4 * d + 2] = P.z; f[4 * d + 3] = P.w;
        v[4 * d] = V.x; v[4 * d + 1] = V.y; v[4 * d + 2] = V.z; v[4 * d + 3] = V.w;
        a[4 * d] = A.x; a[4 * d + 1] = A.y; a[4 * d + 2] = A.z; a[4 * d + 3] = A.w;
      }
#pragma unroll
      for (int d = 0; d < 12; d++) {
        v[d] += mh * a[d];
        f[d] = F64 ? (float)((double)f[d] + dt * (double)v[d]) : f[d] + dtf * v[d];
        f[d] = wrap01(f[d]);
      }
#pragma unroll
      for (int d = 0; d < 3; d++) {
        p4[3 * q + d] = make_float4(f[4 * d], f[4 * d + 1], f[4 * d + 2], f[4 * d + 3]);
        v4[3 * q + d] = make_float4(v[4 * d], v[4 * d + 1], v[4 * d + 2], v[4 * d + 3]);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) c[r] = cell_of(f[3 * r], f[3 * r + 1], f[3 * r + 2], Nf, N, NB, 0, NB);
    }
    // runs of equal keys inside the thread: the run's count rides on its first particle
    int m[4] = {1, 1, 1, 1};
#pragma unroll
    for (int r = 3; r > 0; r--)
      if (c[r] == c[r - 1]) { m[r - 1] += m[r]; m[r] = 0; }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int key = m[r] ? c[r] : -1 - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      int tot = m[r];
      if (__popc(peers) > 1) {
        // (rare at one particle per cell) sum the run lengths of the lanes that share the key
        tot = 0;
        for (unsigned rest = peers; rest; rest &= rest - 1) tot += __shfl_sync(peers, m[r], __ffs(rest) - 1);
      }
      if (valid && key >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&cnt[key], tot);
    }
  }
  // the last np % 4 particles
  if (blockIdx.x == 0 && threadIdx.x < (np & 3)) {
    const int64_t n = (nq << 2) + threadIdx.x;
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      float v = vel[3 * n + d] + mh * acc[3 * n + d];
      float p = pos[3 * n + d];
      p = F64 ? (float)((double)p + dt * (double)v) : p + dtf * v;
      p = wrap01(p);
      vel[3 * n + d] = v;
      pos[3 * n + d] = p;
      x[d] = p;
    }
    atomicAdd(&cnt[cell_of(x[0], x[1], x[2], Nf, N, NB, 0, NB)], 1);
  }
}

// after the scan (start[c] = cell[1 + c]): list the extra parts of the bins that hold more than BIN_PART particles
__global__ void __launch_bounds__(256) bin_heavy_list_kernel(const int *__restrict__ start, int nbins,
                                                             int *__restrict__ heavy_count, int2 *__restrict__ heavy,
                                                             int heavy_cap) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int extra = (start[(size_t)(b + 1) * CPB] - start[(size_t)b * CPB] - 1) / BIN_PART;
  if (extra <= 0) return;
  const int base = atomicAdd(heavy_count, extra);
  for (int e = 0; e < extra; e++)
    if (base + e < heavy_cap) heavy[base + e] = make_int2(b, e + 1);
}

// pass 2: every particle claims the next free record of its cell (cur = BinLayout::cell + 1 holds the first record of
// each cell after the scan and is counted up to the cell's end)
__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ pos, int64_t np, int N, int NB,
                                                          int x0, int NBX, int *__restrict__ cur,
                                                          float4 *__restrict__ brec) {
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    int c = -1 - lane;
    if (n < np) {
      x = __ldg(&pos[3 * n]); y = __ldg(&pos[3 * n + 1]); z = __ldg(&pos[3 * n + 2]);
      c = cell_of(x, y, z, Nf, N, NB, x0, NBX);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (c >= 0 && leader == lane) base = atomicAdd(&cur[c], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (c >= 0) brec[base + __popc(peers & ((1u << lane) - 1u))] = make_float4(x, y, z, __int_as_float((int)n));
  }
}

// ------------------------------------------------------------------------------------- deposit
// particles [beg, end) of bin b; shared = other CTAs deposit into the same bin (its interior cells need atomics too)
template <int SCHEME>
__device__ __forceinline__ void deposit_bin_range(float (*tiles)[BD_TILE], const float4 *__restrict__ brec, int b,
                                                  int beg, int end, bool shared, int N, int NB, int x0, int xoff,
                                                  int nxa, float *__restrict__ rho) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bk = b % NB, bj = (b / NB) % NB, bi = b / (NB * NB);
  const int oi = x0 + bi * BB - 1, oj = bj * BB - 1, ok = bk * BB - 1;  // global cell of tile cell (0,0,0)
  const int pl0 = bi * BB - 1 + xoff;  // its plane in rho (nxa planes; periodic only when nxa == N)
  const float Nf = (float)N;
  float *tile = tiles[warp];
  for (int t = lane; t < BD_TILE; t += 32) tile[t] = 0.0f;
  __syncwarp();
  for (int c = beg + warp * 32; c < end; c += BD_WARPS * 32) {
    const int n = c + lane;
    const bool valid = n < end;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (valid) { const float4 r = __ldg(&brec[n]); px = r.x; py = r.y; pz = r.z; }
    int i, j, k;
    float wx[3], wy[3], wz[3];
    axis_weights<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
    const int t0 = min(max(i - oi, 1), BB), t1 = min(max(j - oj, 1), BB), t2 = min(max(k - ok, 1), BB);  // in [1, 8] for every particle of this bin (clamped like the sort key, cell_of)
    float wgt[27];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const float wxy = wx[a] * wy[e];
#pragma unroll
        for (int g = 0; g < 3; g++) wgt[(a * 3 + e) * 3 + g] = wxy * wz[g];
      }
    // merge lanes of equal cell into the lowest lane of the group
    const int cellkey = valid ? t0 * BD_P0 + t1 * BD_P1 + t2 : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, cellkey);
    const bool is_leader = valid && (__ffs(peers) - 1) == lane;
    unsigned rest = is_leader ? (peers & ~(1u << lane)) : 0u;
    // Dense cells (halo cores): Morton order puts runs of particles of ONE cell into a chunk; a chunk whose 32 lanes
    // all share a cell is summed by a 5-step butterfly instead of 31 merge rounds.
    if (__all_sync(0xffffffffu, peers == 0xffffffffu)) {
#pragma unroll
      for (int q = 0; q < 27; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wgt[q] += __shfl_xor_sync(0xffffffffu, wgt[q], o);
      }
      rest = 0u;
    }
    // (the peer's nine 1-D weights are shuffled and its 27 products re-formed here: 9 SHFL instead of 27 -- the
    // shuffles share the saturated shared-memory pipe with the phases below, the multiplies are free)
    while (__any_sync(0xffffffffu, rest != 0u)) {
      const int src = rest ? (__ffs(rest) - 1) : lane;
      float qx[3], qy[3], qz[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        qx[a] = __shfl_sync(0xffffffffu, wx[a], src);
        qy[a] = __shfl_sync(0xffffffffu, wy[a], src);
        qz[a] = __shfl_sync(0xffffffffu, wz[a], src);
      }
      if (rest) {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int e = 0; e < 3; e++) {
            const float qxy = qx[a] * qy[e];
#pragma unroll
            for (int g = 0; g < 3; g++) wgt[(a * 3 + e) * 3 + g] += qxy * qz[g];
          }
      }
      rest &= rest - 1;
    }
    // 27 conflict-free phases: distinct cells + identical offset => distinct addresses
    float *cell0 = tile + (t0 - 1) * BD_P0 + (t1 - 1) * BD_P1 + (t2 - 1);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++)
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (SCHEME == PSC_NGP && !(a == 1 && e == 1 && g == 1)) continue;
          if (is_leader) {
            float *p = cell0 + a * BD_P0 + e * BD_P1 + g;
            *p += wgt[(a * 3 + e) * 3 + g];
          }
          __syncwarp();
        }
  }
  __syncthreads();
  // sum the per-warp tiles; cells no other bin can reach (2 <= t <= 7 in every dimension) are stored,
  // the shell is added to L2
  // one half-warp per (i, j) row of the tile: wraps and the 64-bit row address once per row, not per cell (the
  // cell-wise loop was 36% of the kernel's instructions in ncu)
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  if (hl < BT) {
    const int gk = wrap(ok + hl, N);
    const bool kin = hl >= 2 && hl <= BT - 3;
    for (int row = hw; row < BT * BT; row += BD_WARPS * 2) {
      const int a = row / BT, e = row - a * BT;
      const int s = a * BD_P0 + e * BD_P1 + hl;
      float v = tiles[0][s];
#pragma unroll
      for (int w = 1; w < BD_WARPS; w++) v += tiles[w][s];
      const int gi = wrap(pl0 + a, nxa), gj = wrap(oj + e, N);
      float *dst = rho + ((size_t)gi * N + gj) * N + gk;
      const bool mine = !shared && kin && a >= 2 && a <= BT - 3 && e >= 2 && e <= BT - 3;
      if (mine) *dst = v;
      else if (v != 0.0f) atomicAdd(dst, v);
    }
  }
}

// ---- lanes own cells (the path of a bin whose cells are about equally filled)
constexpr int DC_WARPS = BB;                // warp w owns plane w of the bin
constexpr int DC_P1 = 24;                   // row pitch of a warp's tile: lanes = 4 rows x 8 cells, 24 r mod 32 = 0, 24, 16, 8
constexpr int DC_PA = BT * DC_P1;           // 240 floats per plane
constexpr int DC_TILE = 3 * DC_PA;          // planes w - 1, w, w + 1

// A bin takes this path when no cell holds more than 8 + 4 x (mean fill) particles: a lane loops over the particles
// of its cell, so a warp is as slow as its fullest cell.
__device__ __forceinline__ int dc_fill_limit(int nbin) { return 8 + (nbin >> 7); }

template <int SCHEME>
__global__ void __launch_bounds__(DC_WARPS * 32) deposit_cells_kernel(const float4 *__restrict__ brec,
                                                                      const int *__restrict__ cell, int N, int NB,
                                                                      int x0, int xoff, int nxa, float *__restrict__ rho,
                                                                      int *__restrict__ dense_count,
                                                                      int *__restrict__ dense, int force_dense) {
  __shared__ int s_cell[CPB + 1];
  __shared__ __align__(16) float s_tile[DC_WARPS][DC_TILE];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int *cb = cell + (size_t)b * CPB;
  for (int t = tid; t <= CPB; t += DC_WARPS * 32) s_cell[t] = __ldg(&cb[t]);
  {
    float4 *t4 = reinterpret_cast<float4 *>(s_tile[w]);
    for (int t = lane; t < DC_TILE / 4; t += 32) t4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int nbin = s_cell[CPB] - s_cell[0];
  if (nbin == 0) return;  // rho was zeroed by the caller
  {
    const int c0 = s_cell[2 * tid + 1] - s_cell[2 * tid], c1 = s_cell[2 * tid + 2] - s_cell[2 * tid + 1];
    const int uneven = __syncthreads_or(max(c0, c1) > dc_fill_limit(nbin));
    if (uneven || nbin > BIN_PART || force_dense) {
      if (tid == 0) dense[atomicAdd(dense_count, 1)] = b;
      return;
    }
  }
  const int bk = b % NB, bj = (b / NB) % NB, bi = b / (NB * NB);
  const int oj = bj * BB - 1, ok = bk * BB - 1;  // global cell of tile cell (., 0, 0)
  const int pl0 = bi * BB - 1 + xoff;            // plane of tile plane 0 in rho (nxa planes; periodic only when nxa == N)
  const float Nf = (float)N;
  float *tile = s_tile[w];
  const int jl = lane >> 3, k = lane & 7;
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    const int lj = 4 * h + jl;
    const int c = (w << 6) | (lj << 3) | k;
    int p = s_cell[c];
    const int pe = s_cell[c + 1];
    const int nmax = __reduce_max_sync(0xffffffffu, pe - p);
    if (nmax == 0) continue;
    float acc[27];
#pragma unroll
    for (int q = 0; q < 27; q++) acc[q] = 0.0f;
    for (int it = 0; it < nmax; it++, p++) {
      if (p < pe) {
        const float4 r = __ldg(&brec[p]);
        int ci, cj, ck;
        float wx[3], wy[3], wz[3];
        axis_weights<SCHEME>(r.x * Nf, N, ci, wx[0], wx[1], wx[2]);
        axis_weights<SCHEME>(r.y * Nf, N, cj, wy[0], wy[1], wy[2]);
        axis_weights<SCHEME>(r.z * Nf, N, ck, wz[0], wz[1], wz[2]);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int e = 0; e < 3; e++) {
            const float wxy = wx[a] * wy[e];
#pragma unroll
            for (int g = 0; g < 3; g++) acc[(a * 3 + e) * 3 + g] += wxy * wz[g];
          }
      }
    }
    // 27 phases: the lanes are 4 rows x 8 consecutive cells of one plane, the offset is the same for all of them
    // => 32 distinct banks, no two lanes on one address.  Tile row lj + e is bin row lj - 1 + e, column k + g is k - 1 + g.
    float *base = tile + lj * DC_P1 + k;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++)
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (SCHEME == PSC_NGP && !(a == 1 && e == 1 && g == 1)) continue;
          base[a * DC_PA + e * DC_P1 + g] += acc[(a * 3 + e) * 3 + g];
          __syncwarp();
        }
  }
  __syncthreads();
  // tile plane A of the bin (bin plane A - 1) = plane a of the tile of warp A - a, a = 0..2.  One half-warp per (A, E)
  // row; cells no other bin can reach (2 <= . <= 7 in every dimension) are stored, the shell is added to L2.
  const int hl = tid & 15, hw = tid >> 4;
  if (hl < BT) {
    const int gk = wrap(ok + hl, N);
    const bool kin = hl >= 2 && hl <= BT - 3;
    for (int row = hw; row < BT * BT; row += DC_WARPS * 2) {
      const int A = row / BT, E = row - A * BT;
      float v = 0.0f;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int ww = A - a;
        if (ww >= 0 && ww < DC_WARPS) v += s_tile[ww][a * DC_PA + E * DC_P1 + hl];
      }
      const int gi = wrap(pl0 + A, nxa), gj = wrap(oj + E, N);
      float *dst = rho + ((size_t)gi * N + gj) * N + gk;
      const bool mine = kin && A >= 2 && A <= BT - 3 && E >= 2 && E <= BT - 3;
      if (mine) *dst = v;
      else if (v != 0.0f) atomicAdd(dst, v);
    }
  }
}

// the bins deposit_cells_kernel listed in `dense` (first BIN_PART particles): lanes own particles (round-1 path)
template <int SCHEME>
__global__ void __launch_bounds__(BD_WARPS * 32) deposit_dense_kernel(const float4 *__restrict__ brec,
                                                                      const int *__restrict__ cell,
                                                                      const int *__restrict__ dense_count,
                                                                      const int *__restrict__ dense, int N, int NB,
                                                                      int x0, int xoff, int nxa, float *__restrict__ rho) {
  __shared__ float tiles[BD_WARPS][BD_TILE];
  const int nitems = *dense_count;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int b = dense[it];
    const int beg = cell[(size_t)b * CPB], end = cell[(size_t)(b + 1) * CPB];
    deposit_bin_range<SCHEME>(tiles, brec, b, beg, min(end, beg + BIN_PART), end - beg > BIN_PART, N, NB, x0, xoff, nxa,
                              rho);
    __syncthreads();
  }
}

// the parts beyond BIN_PART particles of the heavy bins (persistent CTAs over BinLayout::heavy)
template <int SCHEME>
__global__ void __launch_bounds__(BD_WARPS * 32) deposit_heavy_kernel(const float4 *__restrict__ brec,
                                                                      const int *__restrict__ cell,
                                                                      const int *__restrict__ heavy_count,
                                                                      const int2 *__restrict__ heavy, int heavy_cap,
                                                                      int N, int NB, int x0, int xoff, int nxa,
                                                                      float *__restrict__ rho) {
  __shared__ float tiles[BD_WARPS][BD_TILE];
  const int nitems = min(*heavy_count, heavy_cap);
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int2 w = heavy[it];
    const int beg = cell[(size_t)w.x * CPB] + w.y * BIN_PART, end = min(cell[(size_t)(w.x + 1) * CPB], beg + BIN_PART);
    deposit_bin_range<SCHEME>(tiles, brec, w.x, beg, end, true, N, NB, x0, xoff, nxa, rho);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------- interpolation
constexpr int BI_THREADS = 128;

template <int SCHEME>
__global__ void __launch_bounds__(BI_THREADS) interp_kick4_binned_kernel(
    const float4 *__restrict__ force4, const float4 *__restrict__ brec,
    const int *__restrict__ cell, float *__restrict__ vel, float *__restrict__ accel, int N, int NB,
    int x0, int xoff, int nxa, float half_dt, float *__restrict__ maxout, int nbins,
    const int *__restrict__ heavy_count, const int2 *__restrict__ heavy) {
  __shared__ float4 tile[BT * BT * BT];  // 16,000 B
  __shared__ unsigned s_max[BI_THREADS / 32][2];
  int b = blockIdx.x, part = 0;
  if (b >= nbins) {
    const int it = b - nbins;
    if (it >= *heavy_count) return;
    const int2 w = heavy[it];
    b = w.x;
    part = w.y;
  }
  const int beg = cell[(size_t)b * CPB] + part * BIN_PART, end = min(cell[(size_t)(b + 1) * CPB], beg + BIN_PART);
  if (beg >= end) return;
  const int bk = b % NB, bj = (b / NB) % NB, bi = b / (NB * NB);
  const int oi = x0 + bi * BB - 1, oj = bj * BB - 1, ok = bk * BB - 1;
  const int pl0 = bi * BB - 1 + xoff;
  const size_t N2 = (size_t)N * N;
  for (int t = threadIdx.x; t < BT * BT * BT; t += BI_THREADS) {
    const int g = t % BT, r = t / BT;
    const int e = r % BT, a = r / BT;
    tile[t] = __ldg(&force4[(size_t)wrap(pl0 + a, nxa) * N2 + (size_t)wrap(oj + e, N) * N + wrap(ok + g, N)]);
  }
  __syncthreads();
  const float Nf = (float)N;
  const float mh = -half_dt;
  unsigned ma = 0u, mv = 0u;   // maxima of |.| as bit patterns: orders like the floats and lets a NaN win
  for (int n = beg + threadIdx.x; n < end; n += BI_THREADS) {
    const float4 rec = __ldg(&brec[n]);
    const float px = rec.x, py = rec.y, pz = rec.z;
    const int row = __float_as_int(rec.w);
    int i, j, k;
    float wx[3], wy[3], wz[3];
    axis_weights<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
    const float4 *c0 = tile + ((min(max(i - oi, 1), BB) - 1) * BT + (min(max(j - oj, 1), BB) - 1)) * BT + (min(max(k - ok, 1), BB) - 1);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const float wxy = wx[a] * wy[e];
#pragma unroll
        for (int g = 0; g < 3; g++) {
          const float w = wxy * wz[g];
          const float4 f = c0[(a * BT + e) * BT + g];
          ax += w * f.x; ay += w * f.y; az += w * f.z;
        }
      }
    float *ap = accel + 3 * (size_t)row;
    ap[0] = ax; ap[1] = ay; ap[2] = az;
    ma = max(ma, max(__float_as_uint(fabsf(ax)), max(__float_as_uint(fabsf(ay)), __float_as_uint(fabsf(az)))));
    if (vel) {
      float *vp = vel + 3 * (size_t)row;
      const float v0 = vp[0] + mh * ax, v1 = vp[1] + mh * ay, v2 = vp[2] + mh * az;
      vp[0] = v0; vp[1] = v1; vp[2] = v2;
      mv = max(mv, max(__float_as_uint(fabsf(v0)), max(__float_as_uint(fabsf(v1)), __float_as_uint(fabsf(v2)))));
    }
  }
  ma = __reduce_max_sync(0xffffffffu, ma);
  mv = __reduce_max_sync(0xffffffffu, mv);
  if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5][0] = ma; s_max[threadIdx.x >> 5][1] = mv; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < BI_THREADS / 32; w++) { ma = max(ma, s_max[w][0]); mv = max(mv, s_max[w][1]); }
    atomicMax(reinterpret_cast<unsigned *>(&maxout[0]), ma);
    atomicMax(reinterpret_cast<unsigned *>(&maxout[1]), mv);
  }
}

// ------------------------------------------------------- gradient fused into the interpolation
// Same as interp_kick4_binned_kernel, but the CTA derives its 10^3 force tile itself from the potential:
// it stages the (10 + 2H)^3 potential tile (H = stencil reach: 1, 1, 2, 3 for orders 2, 3, 5, 7), applies
// the finite-difference stencil of mesh.derivative{2,3,5,7} (mesh.py:639-850) in shared memory and gathers
// from the result.  For f(R) the tile holds phi + f * u^(n+1) (mesh.derivative*_fR_n{1,2}).  This removes
// the gradient kernel and the force grid (16 B/cell written + re-read) from the step.
template <int ORDER> struct Reach { static constexpr int H = ORDER == 7 ? 3 : ORDER == 5 ? 2 : 1; };

constexpr int BP_THREADS = 256;  // gradient + interpolation kernel

// Finite-difference gradient of one half (5 cells, HH = 0 / 1) of a (i, j) row of the force tile, from the potential
// tile in shared memory.  The row is read as 16-byte groups and the k derivative is formed in registers; the rows at
// +-s in i and j contribute their 5 cells as one LDS.128 + one LDS.32.  With the potential rows 20 floats apart, eight
// consecutive rows fall on the eight distinct 16-byte bank groups, so lanes that walk consecutive rows do not collide.
// Window position p of a row is tile cell g = p - 3 (the window is the aligned 16 floats [8 bk - 4, 8 bk + 12)).
constexpr int PKR = 20;

template <int HH>
__device__ __forceinline__ void load5(const float *r, float v[5]) {
  if (HH == 0) {
    v[0] = r[3];
    const float4 t = *reinterpret_cast<const float4 *>(r + 4);
    v[1] = t.x; v[2] = t.y; v[3] = t.z; v[4] = t.w;
  } else {
    const float4 t = *reinterpret_cast<const float4 *>(r + 8);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    v[4] = r[12];
  }
}

template <int ORDER, int HH>
__device__ __forceinline__ void diff5(const float *c0, int s, const float cen[5], float pref, float out[5]) {
  float m1[5], p1[5];
  load5<HH>(c0 + s, p1);
  if (ORDER == 2) {
#pragma unroll
    for (int q = 0; q < 5; q++) out[q] = pref * (-cen[q] + p1[q]);
    return;
  }
  load5<HH>(c0 - s, m1);
  if (ORDER == 3) {
#pragma unroll
    for (int q = 0; q < 5; q++) out[q] = pref * (-m1[q] + p1[q]);
    return;
  }
  float m2[5], p2[5];
  load5<HH>(c0 - 2 * s, m2);
  load5<HH>(c0 + 2 * s, p2);
  if (ORDER == 5) {
#pragma unroll
    for (int q = 0; q < 5; q++) out[q] = pref * (8.0f * (-m1[q] + p1[q]) + m2[q] - p2[q]);
    return;
  }
  float m3[5], p3[5];
  load5<HH>(c0 - 3 * s, m3);
  load5<HH>(c0 + 3 * s, p3);
#pragma unroll
  for (int q = 0; q < 5; q++) out[q] = pref * (45.0f * (-m1[q] + p1[q]) + 9.0f * (m2[q] - p2[q]) - m3[q] + p3[q]);
}

template <int ORDER, int HH>
__device__ __forceinline__ void force_half_row(const float *ptile, int a, int e, float pref, float4 *dst) {
  constexpr int H = Reach<ORDER>::H;
  constexpr int PT = BT + 2 * H;
  const float *c0 = ptile + ((a + H) * PT + (e + H)) * PKR;
  // own row, window positions [4 HH, 4 HH + 12): cell q of this half is window position 5 HH + 3 + q = c[HH + 3 + q]
  float c[12];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const float4 v = *reinterpret_cast<const float4 *>(c0 + 4 * HH + 4 * t);
    c[4 * t] = v.x; c[4 * t + 1] = v.y; c[4 * t + 2] = v.z; c[4 * t + 3] = v.w;
  }
  constexpr int CI = HH + 3;
  float cen[5], gx[5], gy[5], gz[5];
#pragma unroll
  for (int q = 0; q < 5; q++) {
    cen[q] = c[CI + q];
    if (ORDER == 2) gz[q] = pref * (-c[CI + q] + c[CI + q + 1]);
    else if (ORDER == 3) gz[q] = pref * (-c[CI + q - 1] + c[CI + q + 1]);
    else if (ORDER == 5) gz[q] = pref * (8.0f * (-c[CI + q - 1] + c[CI + q + 1]) + c[CI + q - 2] - c[CI + q + 2]);
    else gz[q] = pref * (45.0f * (-c[CI + q - 1] + c[CI + q + 1]) + 9.0f * (c[CI + q - 2] - c[CI + q + 2]) - c[CI + q - 3] + c[CI + q + 3]);
  }
  diff5<ORDER, HH>(c0, PT * PKR, cen, pref, gx);
  diff5<ORDER, HH>(c0, PKR, cen, pref, gy);
#pragma unroll
  for (int q = 0; q < 5; q++) dst[5 * HH + q] = make_float4(gx[q], gy[q], gz[q], 0.0f);
}

// TP1 / TP0: row / plane pitch of the float4 force tile.  Measured at 512^3 (Morton order, round 1): 10/100 4.41 ms,
// 11/110 4.62, 12/120 4.59, 12/144 4.75, 14/140 4.70, 11/112 5.19, 10/104 5.39 -- the dense tile is the best of those.
// GRAD = 1: row-wise gradient stage (force_half_row); GRAD = 0: the round-1 cell-wise stage (one thread per tile cell,
// 12 scalar LDS each with two-way bank conflicts: a third of this kernel's shared-memory wavefronts in ncu) -- kept
// for A/B measurements (PSC_INTERP_MODE=1).
template <int SCHEME, int ORDER, int GRAD, int TP1 = BT, int TP0 = BT * BT>
__global__ void __launch_bounds__(BP_THREADS) interp_kick_phi_binned_kernel(
    const float *__restrict__ phi, const float *__restrict__ u, float f, int fr_n,
    const float4 *__restrict__ brec, const int *__restrict__ cell,
    float *__restrict__ vel, float *__restrict__ accel, int N, int NB, int x0, int xoff, int nxa, float half_dt,
    float *__restrict__ maxout, int nbins, const int *__restrict__ heavy_count, const int2 *__restrict__ heavy) {
  constexpr int H = Reach<ORDER>::H;
  constexpr int PT = BT + 2 * H;          // potential tile edge
  constexpr int PK = GRAD ? PKR : 16;     // k pitch of the potential tile; a row holds the aligned 16-float window
  __shared__ __align__(16) float ptile[PT * PT * PK];
  __shared__ float4 tile[BT * TP0];
  __shared__ unsigned s_max[BP_THREADS / 32][2];
  // CTAs [0, nbins): the first BIN_PART particles of bin blockIdx.x; CTAs beyond: one listed part of a heavy bin
  int b = blockIdx.x, part = 0;
  if (b >= nbins) {
    const int it = b - nbins;
    if (it >= *heavy_count) return;
    const int2 w = heavy[it];
    b = w.x;
    part = w.y;
  }
  const int beg = cell[(size_t)b * CPB] + part * BIN_PART, end = min(cell[(size_t)(b + 1) * CPB], beg + BIN_PART);
  if (beg >= end) return;
  const int bk = b % NB, bj = (b / NB) % NB, bi = b / (NB * NB);
  const int oi = x0 + bi * BB - 1, oj = bj * BB - 1, ok = bk * BB - 1;
  const int pl0 = bi * BB - 1 + xoff;
  {
    // Every (i, j) row of the potential tile is fetched as the aligned 16-float window [8 bk - 4, 8 bk + 12) that
    // contains the PT cells the stencils need: four LDG.128 per row (784 per tile instead of 2744 scalar loads,
    // whose index arithmetic was 59% of this kernel's instructions in ncu), all issued before the first store.  A
    // window group never straddles the periodic boundary because N % 4 == 0.
    const float4 *phi4 = reinterpret_cast<const float4 *>(phi);
    const float4 *u4 = reinterpret_cast<const float4 *>(u);
    float4 *pt4 = reinterpret_cast<float4 *>(ptile);
    const int n4 = N >> 2;
    constexpr int NITEM = PT * PT * 4;
    constexpr int NIT = (NITEM + BP_THREADS - 1) / BP_THREADS;
    float4 v[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int item = threadIdx.x + it * BP_THREADS;
      const int row = item >> 2, q = item & 3;
      const int a = row / PT, e = row - a * PT;
      int gi = pl0 - H + a, gj = oj - H + e, g4 = 2 * bk - 1 + q;
      gi += gi < 0 ? nxa : 0; gi -= gi >= nxa ? nxa : 0;
      gj += gj < 0 ? N : 0; gj -= gj >= N ? N : 0;
      g4 += g4 < 0 ? n4 : 0; g4 -= g4 >= n4 ? n4 : 0;
      const size_t c = ((size_t)gi * N + gj) * n4 + g4;
      v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (item < NITEM) {
        v[it] = __ldg(&phi4[c]);
        if (fr_n) {
          const float4 w = __ldg(&u4[c]);
          v[it].x += f * (fr_n == 1 ? w.x * w.x : w.x * w.x * w.x);
          v[it].y += f * (fr_n == 1 ? w.y * w.y : w.y * w.y * w.y);
          v[it].z += f * (fr_n == 1 ? w.z * w.z : w.z * w.z * w.z);
          v[it].w += f * (fr_n == 1 ? w.w * w.w : w.w * w.w * w.w);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int item = threadIdx.x + it * BP_THREADS;
      if (item < NITEM) pt4[(item >> 2) * (PK / 4) + (item & 3)] = v[it];
    }
  }
  __syncthreads();
  const float pref = ORDER == 2 ? (float)N : ORDER == 3 ? (float)(0.5 * N) : ORDER == 5 ? (float)(N / 12.0) : (float)(N / 60.0);
  if (GRAD) {
    // warps 0-3: cells 0..4 of rows 0..99, warps 4-7: cells 5..9 (the half is warp-uniform: no divergence)
    const int row = threadIdx.x & 127;
    if (row < BT * BT) {
      const int a = row / BT, e = row - a * BT;
      float4 *dst = tile + a * TP0 + e * TP1;
      if (threadIdx.x < 128) force_half_row<ORDER, 0>(ptile, a, e, pref, dst);
      else force_half_row<ORDER, 1>(ptile, a, e, pref, dst);
    }
  } else {
    for (int t = threadIdx.x; t < BT * BT * BT; t += BP_THREADS) {
      const int g = t % BT, r = t / BT;
      const int e = r % BT, a = r / BT;
      // tile cell g is global k = 8 bk - 1 + g = window position g + 3
      const float *c = ptile + ((a + H) * PT + (e + H)) * PK + (g + 3);
      float gr[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const int s = d == 0 ? PT * PK : d == 1 ? PK : 1;
        if (ORDER == 2) gr[d] = pref * (-c[0] + c[s]);
        else if (ORDER == 3) gr[d] = pref * (-c[-s] + c[s]);
        else if (ORDER == 5) gr[d] = pref * (8.0f * (-c[-s] + c[s]) + c[-2 * s] - c[2 * s]);
        else gr[d] = pref * (45.0f * (-c[-s] + c[s]) + 9.0f * (c[-2 * s] - c[2 * s]) - c[-3 * s] + c[3 * s]);
      }
      tile[a * TP0 + e * TP1 + g] = make_float4(gr[0], gr[1], gr[2], 0.0f);
    }
  }
  __syncthreads();
  const float Nf = (float)N;
  const float mh = -half_dt;
  unsigned ma = 0u, mv = 0u;   // maxima of |.| as bit patterns: orders like the floats and lets a NaN win
  for (int n = beg + threadIdx.x; n < end; n += BP_THREADS) {
    const float4 rec = __ldg(&brec[n]);
    const float px = rec.x, py = rec.y, pz = rec.z;
    const int row = __float_as_int(rec.w);
    int i, j, k;
    float wx[3], wy[3], wz[3];
    axis_weights<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
    const float4 *c0 = tile + (min(max(i - oi, 1), BB) - 1) * TP0 + (min(max(j - oj, 1), BB) - 1) * TP1 + (min(max(k - ok, 1), BB) - 1);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const float wxy = wx[a] * wy[e];
#pragma unroll
        for (int g = 0; g < 3; g++) {
          const float w = wxy * wz[g];
          const float4 ff = c0[a * TP0 + e * TP1 + g];
          ax += w * ff.x; ay += w * ff.y; az += w * ff.z;
        }
      }
    float *ap = accel + 3 * (size_t)row;
    ap[0] = ax; ap[1] = ay; ap[2] = az;
    ma = max(ma, max(__float_as_uint(fabsf(ax)), max(__float_as_uint(fabsf(ay)), __float_as_uint(fabsf(az)))));
    if (vel) {
      float *vp = vel + 3 * (size_t)row;
      const float v0 = vp[0] + mh * ax, v1 = vp[1] + mh * ay, v2 = vp[2] + mh * az;
      vp[0] = v0; vp[1] = v1; vp[2] = v2;
      mv = max(mv, max(__float_as_uint(fabsf(v0)), max(__float_as_uint(fabsf(v1)), __float_as_uint(fabsf(v2)))));
    }
  }
  ma = __reduce_max_sync(0xffffffffu, ma);
  mv = __reduce_max_sync(0xffffffffu, mv);
  if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5][0] = ma; s_max[threadIdx.x >> 5][1] = mv; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < BP_THREADS / 32; w++) { ma = max(ma, s_max[w][0]); mv = max(mv, s_max[w][1]); }
    atomicMax(reinterpret_cast<unsigned *>(&maxout[0]), ma);
    atomicMax(reinterpret_cast<unsigned *>(&maxout[1]), mv);
  }
}

// rho = f1 * (scale * rho) + f2, defined in deposit.cu
__global__ void rho_affine_kernel(float *rho, int64_t n, float scale, float f1, float f2, int do_scale);

}  // namespace psc

using namespace psc;

extern "C" {

static bool slab_ok(int N, int x0, int nxl) {
  return N >= BB && (N % BB) == 0 && N <= 32767 && nxl >= BB && (nxl % BB) == 0 && x0 >= 0 && x0 + nxl <= N;
}

size_t psc_bin_workspace_bytes_slab(int64_t np, int N, int nxl) {
  if (np < 0 || !slab_ok(N, 0, nxl)) return 0;
  const int64_t nbins = (int64_t)(nxl / BB) * (N / BB) * (N / BB);
  const int64_t ncells = nbins * CPB;
  if (ncells + 2 >= ((int64_t)1 << 31)) return 0;
  return a256(sizeof(int) * (size_t)(ncells + 2)) + a256(sizeof(float4) * (size_t)np) + 256 +
         a256(sizeof(int) * (size_t)nbins) + a256(sizeof(int2) * (size_t)(np / BIN_PART + 1)) +
         a256(scan_tmp_bytes(ncells + 1)) + 256;
}
size_t psc_bin_workspace_bytes(int64_t np, int N) { return psc_bin_workspace_bytes_slab(np, N, N); }

// scan + heavy list + scatter from the per-cell counts in L.cell[1 ..]
static int finish_binning(const float *pos, int64_t np, int N, int x0, const BinLayout &L, cudaStream_t st) {
  cudaError_t e = cub::DeviceScan::ExclusiveSum(L.cub_tmp, const_cast<size_t &>(L.cub_bytes), L.cell + 1, L.cell + 1,
                                                (int)(L.ncells + 1), st);
  count_launch(2);
  if (e != cudaSuccess) {
    set_error("psc_bin_particles: cub scan failed: %s", cudaGetErrorString(e));
    return PSC_ERR_CUDA;
  }
  PSC_CUDA(cudaMemsetAsync(L.dense_count, 0, 256, st));   // dense_count and heavy_count
  bin_heavy_list_kernel<<<(int)((L.nbins + 255) / 256), 256, 0, st>>>(L.cell + 1, (int)L.nbins, L.heavy_count, L.heavy,
                                                                     L.heavy_cap);
  count_launch();
  if (np > 0) {
    bin_scatter_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, x0, L.NBX, L.cell + 1, L.rec);
    count_launch();
  }
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_bin_particles_slab(const float *pos, int64_t np, int N, int x0, int nxl, void *scratch, size_t scratch_bytes,
                           void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl), "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(scratch && (pos || np == 0), "null pointer");
  PSC_CHECK_ARG(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, x0, nxl, L)) {
    set_error("psc_bin_particles: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(L.cell, 0, sizeof(int) * (size_t)(L.ncells + 2), st));
  if (np > 0) {
    bin_count_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, x0, L.NBX, L.cell + 1);
    count_launch();
  }
  return finish_binning(pos, np, N, x0, L, st);
}
int psc_bin_particles(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, void *stream) {
  return psc_bin_particles_slab(pos, np, N, 0, N, scratch, scratch_bytes, stream);
}

int psc_kick_drift_wrap_count(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                              int dt_is_f64, int N, int64_t np_total, void *scratch, size_t scratch_bytes,
                              int zero_counts, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np <= np_total && np_total < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && ((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np_total, N, 0, N, L)) {
    set_error("psc_kick_drift_wrap_count: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (zero_counts) PSC_CUDA(cudaMemsetAsync(L.cell, 0, sizeof(int) * (size_t)(L.ncells + 2), st));
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos && vel && acc, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)acc) & 15) == 0, "pointers must be 16-byte aligned");
  const int g = grid_for((np + 3) / 4, 256, 8);
  if (dt_is_f64)
    kick_drift_wrap_count_kernel<true><<<g, 256, 0, st>>>(pos, vel, acc, np, half_dt, dt, N, L.NB, L.cell + 1);
  else
    kick_drift_wrap_count_kernel<false><<<g, 256, 0, st>>>(pos, vel, acc, np, half_dt, dt, N, L.NB, L.cell + 1);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_bin_particles_counted(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && (pos || np == 0), "null pointer");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, 0, N, L)) {
    set_error("psc_bin_particles_counted: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  return finish_binning(pos, np, N, 0, L, as_stream(stream));
}

// ghost = 0: periodic N^3 grid (x0 = 0, nxl = N); ghost = 1: rho has nxl + 2 planes, plane 0 / nxl + 1 collect the
// mass that belongs to the neighbouring slabs
static int deposit_binned_impl(const void *scratch, size_t scratch_bytes, int64_t np, int N, int x0, int nxl,
                               int ghost, int scheme, float scale, float f1, float f2, float *rho, void *stream) {
  BinLayout L;
  if (!bin_layout(const_cast<void *>(scratch), scratch_bytes, np, N, x0, nxl, L)) {
    set_error("psc_deposit_binned: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int nxa = nxl + 2 * ghost;
  const int64_t n3 = (int64_t)nxa * N * N;
  PSC_CUDA(cudaMemsetAsync(rho, 0, sizeof(float) * n3, st));
  if (np > 0) {
    const int grid = (int)L.nbins;
    const int hgrid = num_sms() * 4;   // persistent CTAs over the listed bins / parts (exit at once when there are none)
    const int force_dense = deposit_mode() == 1;
    PSC_CUDA(cudaMemsetAsync(L.dense_count, 0, sizeof(int), st));
#define PSC_DEP(S)                                                                                                   \
  deposit_cells_kernel<S><<<grid, DC_WARPS * 32, 0, st>>>(L.rec, L.cell, N, L.NB, x0, ghost, nxa, rho, L.dense_count, \
                                                          L.dense, force_dense);                                      \
  deposit_dense_kernel<S><<<2 * hgrid, BD_WARPS * 32, 0, st>>>(L.rec, L.cell, L.dense_count, L.dense, N, L.NB, x0,    \
                                                              ghost, nxa, rho);                                       \
  deposit_heavy_kernel<S><<<hgrid, BD_WARPS * 32, 0, st>>>(L.rec, L.cell, L.heavy_count, L.heavy, L.heavy_cap, N,     \
                                                          L.NB, x0, ghost, nxa, rho)
    if (scheme == PSC_TSC) { PSC_DEP(PSC_TSC); }
    else if (scheme == PSC_CIC) { PSC_DEP(PSC_CIC); }
    else { PSC_DEP(PSC_NGP); }
#undef PSC_DEP
    count_launch(3);
    PSC_CHECK_LAUNCH();
  }
  if (scale != 1.0f || f1 != 1.0f || f2 != 0.0f) {
    rho_affine_kernel<<<grid_for((n3 + 3) / 4, 256), 256, 0, st>>>(rho, n3, scale, f1, f2, scale != 1.0f);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return PSC_OK;
}

int psc_deposit_binned(const void *scratch, size_t scratch_bytes, int64_t np, int N, int scheme, float scale, float f1,
                       float f2, float *rho, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && rho, "null pointer");
  return deposit_binned_impl(scratch, scratch_bytes, np, N, 0, N, 0, scheme, scale, f1, f2, rho, stream);
}

int psc_deposit_binned_slab(const void *scratch, size_t scratch_bytes, int64_t np, int N, int x0, int nxl, int scheme,
                            float *rho_ghost, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl), "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(scratch && rho_ghost, "null pointer");
  return deposit_binned_impl(scratch, scratch_bytes, np, N, x0, nxl, 1, scheme, 1.0f, 1.0f, 0.0f, rho_ghost, stream);
}

int psc_interp_kick4_binned(const float *force4, const void *scratch, size_t scratch_bytes, float *vel, float *acc,
                            int64_t np, int N, int scheme, float half_dt, float *maxout, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(force4 && scratch && acc && maxout, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)force4 & 15) == 0, "force4 must be 16-byte aligned");
  if (np == 0) return PSC_OK;
  BinLayout L;
  if (!bin_layout(const_cast<void *>(scratch), scratch_bytes, np, N, 0, N, L)) {
    set_error("psc_interp_kick4_binned: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const float4 *f4 = reinterpret_cast<const float4 *>(force4);
  const int grid = (int)L.nbins + L.heavy_cap;
  if (scheme == PSC_TSC)
    interp_kick4_binned_kernel<PSC_TSC><<<grid, BI_THREADS, 0, st>>>(f4, L.rec, L.cell, vel, acc, N, L.NB, 0, 0, N, half_dt, maxout, (int)L.nbins, L.heavy_count, L.heavy);
  else
    interp_kick4_binned_kernel<PSC_CIC><<<grid, BI_THREADS, 0, st>>>(f4, L.rec, L.cell, vel, acc, N, L.NB, 0, 0, N, half_dt, maxout, (int)L.nbins, L.heavy_count, L.heavy);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

// ghost = 0: periodic N^3 grids; ghost = G >= 1 + reach(order): phi (and u) have nxl + 2G planes, the owned ones
// start at plane G
static int interp_kick_phi_impl(const float *phi, const float *u, float f, int fr_n, int order, int x0, int nxl,
                                int ghost, const void *scratch, size_t scratch_bytes, float *vel, float *acc,
                                int64_t np, int N, int scheme, float half_dt, float *maxout, void *stream) {
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG((((uintptr_t)phi | (uintptr_t)u) & 15) == 0, "phi and u must be 16-byte aligned");
  BinLayout L;
  if (!bin_layout(const_cast<void *>(scratch), scratch_bytes, np, N, x0, nxl, L)) {
    set_error("psc_interp_kick_phi_binned: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int grid = (int)L.nbins + L.heavy_cap;   // the CTAs of unused heavy-part slots exit at once
  const int nxa = nxl + 2 * ghost;
  const bool cellwise = interp_mode() == 1;
#define PSC_IKP(S, O)                                                                                               \
  do {                                                                                                              \
    if (cellwise)                                                                                                   \
      interp_kick_phi_binned_kernel<S, O, 0><<<grid, BP_THREADS, 0, st>>>(phi, u, f, fr_n, L.rec, L.cell, vel, acc, \
                                                                          N, L.NB, x0, ghost, nxa, half_dt, maxout, \
                                                                          (int)L.nbins, L.heavy_count, L.heavy);    \
    else                                                                                                            \
      interp_kick_phi_binned_kernel<S, O, 1><<<grid, BP_THREADS, 0, st>>>(phi, u, f, fr_n, L.rec, L.cell, vel, acc, \
                                                                          N, L.NB, x0, ghost, nxa, half_dt, maxout, \
                                                                          (int)L.nbins, L.heavy_count, L.heavy);    \
  } while (0)
#define PSC_IKP_O(S)               \
  if (order == 2) PSC_IKP(S, 2);    \
  else if (order == 3) PSC_IKP(S, 3); \
  else if (order == 5) PSC_IKP(S, 5); \
  else PSC_IKP(S, 7);
  if (scheme == PSC_TSC) { PSC_IKP_O(PSC_TSC) } else { PSC_IKP_O(PSC_CIC) }
#undef PSC_IKP_O
#undef PSC_IKP
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_interp_kick_phi_binned(const float *phi, const float *u, float f, int fr_n, int order, const void *scratch,
                               size_t scratch_bytes, float *vel, float *acc, int64_t np, int N, int scheme,
                               float half_dt, float *maxout, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  PSC_CHECK_ARG(order == 2 || order == 3 || order == 5 || order == 7, "gradient order must be 2, 3, 5 or 7");
  PSC_CHECK_ARG(fr_n >= 0 && fr_n <= 2, "fR_n must be 1 or 2");
  PSC_CHECK_ARG(N >= 2 * BB && (N % BB) == 0, "N must be a multiple of 8 and >= 16");
  PSC_CHECK_ARG(phi && scratch && acc && maxout && (u || fr_n == 0), "null pointer");
  return interp_kick_phi_impl(phi, u, f, fr_n, order, 0, N, 0, scratch, scratch_bytes, vel, acc, np, N, scheme,
                              half_dt, maxout, stream);
}

int psc_interp_kick_phi_binned_slab(const float *phi_ghost, const float *u_ghost, float f, int fr_n, int order,
                                    int x0, int nxl, int ghost, const void *scratch, size_t scratch_bytes, float *vel,
                                    float *acc, int64_t np, int N, int scheme, float half_dt, float *maxout,
                                    void *stream) {
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  PSC_CHECK_ARG(order == 2 || order == 3 || order == 5 || order == 7, "gradient order must be 2, 3, 5 or 7");
  PSC_CHECK_ARG(fr_n >= 0 && fr_n <= 2, "fR_n must be 1 or 2");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl) && N >= 2 * BB, "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(ghost >= 1 + (order == 7 ? 3 : order == 5 ? 2 : 1), "not enough ghost planes for this stencil");
  PSC_CHECK_ARG(phi_ghost && scratch && acc && maxout && (u_ghost || fr_n == 0), "null pointer");
  return interp_kick_phi_impl(phi_ghost, u_ghost, f, fr_n, order, x0, nxl, ghost, scratch, scratch_bytes, vel, acc, np,
                              N, scheme, half_dt, maxout, stream);
}

}  // extern "C"
