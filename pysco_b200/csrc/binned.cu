// binned.cu -- order-independent particle <-> mesh kernels on a per-step "shadow" binning of the particles.
//
// Motivation (measured): the walking-window deposit and the direct gather depend on how
// well the reference-ordered particle array is still sorted (deposit 5.0 ms right after a Morton reorder,
// 2x that before the next one, 69 ms for a random order).  The reference order itself must not change
// (bit-exact ordering parity), so every step the positions are COPIED into bins of 8^3 cells:
//
//   psc_bin_particles   count (warp-aggregated int atomics) -> exclusive scan (CUB) -> scatter of
//                       (x,y,z) and the source row into bin order.  ~32 B / particle of traffic.
//   psc_deposit_binned  one CTA per bin: every particle of the bin lies inside the CTA's 10^3-cell tile
//                       (8^3 + one halo cell), so there is no bounding-box logic, no re-anchoring and no
//                       fallback.  Each warp accumulates chunks of 32 particles into its private tile with
//                       the conflict-free phase scheme of deposit_window.cuh (merge equal cells, 27 plain
//                       LDS/FADD/STS phases); the four tiles are summed, the 6^3 cells no other bin can
//                       touch are stored plainly and only the shell (784 cells per bin, ~1.5 per particle)
//                       goes to L2 as float REDs.
//   psc_interp_kick4_binned  one CTA per bin: the 10^3 float4 force tile is staged once in shared memory
//                       (16 KB), every particle gathers its 27 points with LDS.128, the result is written to
//                       the particle's ORIGINAL row (acceleration, velocity kick) through the source index.
//
// Round 2 measured a CELL-granular counting sort (records in cell order, lanes of the deposit owning cells instead of
// particles, 27 conflict-free read-modify-writes per CELL) against this scheme and rejected it: the per-cell counters
// (0.54 GB at 512^3) make the count / scatter passes atomic-latency bound (kick+drift+count 1.47 -> 1.87 ms, scan +
// scatter 0.95 -> 1.84 ms, 5.5 ms once the source order has drifted by a cell), the lane-per-cell loop runs at the
// speed of the fullest of a warp's 32 cells (11 of 32 lanes busy on a jittered lattice: 3.5 ms, against 2.9 ms here),
// and the gather of the interpolation only drops from 7.9 to 6.5 wavefronts per LDS.128 on cell-sorted records.
// Whole step 17.5 ms against 13.7 ms.  Evidence: profiles/r02_cellsort_experiment/.
#include <cstdlib>
#include <type_traits>
#include <cstring>

#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda/barrier>
#include <cuda/ptx>

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

struct BinLayout {
  int NB;            // bins along y and z (N / 8)
  int NBX;           // bins along x (owned planes / 8; == NB for the periodic single-domain case)
  int x0;            // first owned x cell (0 unless the mesh is slab-decomposed)
  int64_t nbins;     // NBX * NB^2
  int64_t nrec;      // records `rec` can hold: np + np / 8 + 40 nbins (room for the slack of the direct scatter)
  int *counts;       // [nbins + 1]  result of a count pass (input of the scan of the exact binning)
  int *fill;         // [nbins + 1]  records written to each bin: the particles of bin b are rec[base[b] .. base[b] + fill[b])
  int *base;         // [nbins + 1]  first record of each bin.  Exact binning: the exclusive scan of counts (no gaps).
                     //              Direct scatter (psc_kick_drift_wrap_count, mode 1): the scan of capacities derived
                     //              from the PREVIOUS step's fill, so every bin has some slack and the kick-drift-wrap
                     //              pass can drop its records straight into place without a count pass.
  int *tmp;          // [nbins + 1]  capacities / fall-back offsets
  int *overflow;     // [1] set by the direct scatter when a bin ran out of slack: the exact binning is redone
  // sorted layout only (with_rec = false): a SECOND bin table.  The arrays a step reads are described by one table,
  // the arrays it writes by the other (psc_step_sort: src_table / 1 - src_table).
  int *fill2, *base2;
  float4 *rec;       // [nrec] binned particles: (x, y, z, source row as int bits) -- one 16-byte access per particle
  int *heavy_count;  // [1] number of entries of `heavy`
  int2 *heavy;       // [heavy_cap] (bin, part >= 1): the parts beyond the first BIN_PART particles of a bin
  int heavy_cap;
  void *cub_tmp;
  size_t cub_bytes;
};

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

static bool env_is(const char *name, const char *value) {
  const char *e = getenv(name);
  return e && strcmp(e, value) == 0;
}

static size_t scan_tmp_bytes(int64_t n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int *)nullptr, (int *)nullptr, (int)n);
  return b;
}

static int64_t rec_capacity(int64_t np, int64_t nbins) { return np + np / 8 + 40 * nbins + 64; }
constexpr int MB_PER_BIN = 64;   // micro-blocks (2^3 cells) per bin: the in-bin order of the sorted layout

static bool bin_layout(void *scratch, size_t bytes, int64_t np, int N, int x0, int nxl, BinLayout &L,
                       bool with_rec = true) {
  L.NB = N / BB;
  L.NBX = nxl / BB;
  L.x0 = x0;
  L.nbins = (int64_t)L.NBX * L.NB * L.NB;
  char *p = reinterpret_cast<char *>(scratch);
  size_t off = 0;
  L.nrec = with_rec ? rec_capacity(np, L.nbins) : 0;   // sorted particle arrays need no binned copy
  L.counts = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (L.nbins + 1));
  L.fill = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (L.nbins + 1));
  L.base = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (L.nbins + 1));
  L.tmp = reinterpret_cast<int *>(p + off); off += a256(sizeof(int) * (L.nbins + 1));
  L.rec = reinterpret_cast<float4 *>(p + off); off += a256(sizeof(float4) * (size_t)L.nrec);
  L.overflow = reinterpret_cast<int *>(p + off); off += 128;
  L.heavy_count = reinterpret_cast<int *>(p + off); off += 128;
  L.fill2 = reinterpret_cast<int *>(p + off); off += with_rec ? 0 : a256(sizeof(int) * (L.nbins + 1));
  L.base2 = reinterpret_cast<int *>(p + off); off += with_rec ? 0 : a256(sizeof(int) * (L.nbins + 1));
  L.heavy_cap = (int)(np / BIN_PART) + 1;
  L.heavy = reinterpret_cast<int2 *>(p + off); off += a256(sizeof(int2) * (size_t)L.heavy_cap);
  L.cub_tmp = p + off;
  L.cub_bytes = scan_tmp_bytes(L.nbins + 1);
  off += a256(L.cub_bytes);
  return off <= bytes;
}

// x0 / NBX: the slab of owned planes [x0, x0 + 8 NBX); a particle outside it (the host migrates particles
// before binning) is clamped into the edge bin so that nothing is ever written out of bounds.
// y / z outside [0, 1) (an exact 1.0 from an external snapshot, an inf) are clamped into the edge bins too.
__device__ __forceinline__ int bin_of(float x, float y, float z, float Nf, int NB, int x0, int NBX) {
  const int i = (int)(x * Nf) - x0, j = (int)(y * Nf), k = (int)(z * Nf);
  const int bi = min(max(i >> 3, 0), NBX - 1), bj = min(max(j >> 3, 0), NB - 1), bk = min(max(k >> 3, 0), NB - 1);
  return (bi * NB + bj) * NB + bk;
}

// pass 1: counts[bin] += 1, one atomic per distinct bin per warp
__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ pos, int64_t np, int N, int NB,
                                                        int x0, int NBX, int *__restrict__ counts,
                                                        const int *__restrict__ only_if) {
  if (only_if && *only_if == 0) return;
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    int b = -1 - lane;
    if (n < np) b = bin_of(__ldg(&pos[3 * n]), __ldg(&pos[3 * n + 1]), __ldg(&pos[3 * n + 2]), Nf, NB, x0, NBX);
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (b >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&counts[b], __popc(peers));
  }
}

// pass 1 fused into the first half of the leapfrog step (integration.py:250-258): v -= half_dt a; x += dt v; wrap(x);
// counts[bin(x)] += 1.  Four particles (three float4 per array) per thread; the common case "all four in one bin" costs
// one warp-aggregated atomic.  Saves the separate read of the positions that bin_count_kernel does.
// DIRECT: instead of counting, every particle claims the next free record of its bin (fill[] is the cursor, base[] the
// first record of each bin, with slack from the previous step's fill) and its (x, y, z, row) record is written at once:
// the binning costs no second pass over the positions.  A bin whose slack is exhausted sets *overflow (the caller
// then redoes the exact binning); row0 = global row of pos[0] (chunked uploads).
template <bool F64, bool DIRECT>
__global__ void __launch_bounds__(256) kick_drift_wrap_count_kernel(float *__restrict__ pos, float *__restrict__ vel,
                                                                    const float *__restrict__ acc, int64_t np,
                                                                    float half_dt, double dt, int N, int NB,
                                                                    int *__restrict__ counts,
                                                                    const int *__restrict__ bbase, int64_t nrec,
                                                                    float4 *__restrict__ brec, int row0,
                                                                    int *__restrict__ overflow) {
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
    int b[4] = {-1 - lane, -1 - lane, -1 - lane, -1 - lane};
    if (valid) {
      float v[12], a[12];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 P = p4[3 * q + c], V = v4[3 * q + c], A = __ldg(&a4[3 * q + c]);
        f[4 * c] = P.x; f[4 * c + 1] = P.y; f[4 * c + 2] = P.z; f[4 * c + 3] = P.w;
        v[4 * c] = V.x; v[4 * c + 1] = V.y; v[4 * c + 2] = V.z; v[4 * c + 3] = V.w;
        a[4 * c] = A.x; a[4 * c + 1] = A.y; a[4 * c + 2] = A.z; a[4 * c + 3] = A.w;
      }
#pragma unroll
      for (int c = 0; c < 12; c++) {
        kick_drift_wrap1<F64>(f[c], v[c], a[c], mh, dtf, dt);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        p4[3 * q + c] = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
        v4[3 * q + c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) b[r] = bin_of(f[3 * r], f[3 * r + 1], f[3 * r + 2], Nf, NB, 0, NB);
    }
    const bool same = b[0] == b[1] && b[1] == b[2] && b[2] == b[3];
    if (!DIRECT) {
      if (__all_sync(0xffffffffu, same)) {
        const unsigned peers = __match_any_sync(0xffffffffu, b[0]);
        if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&counts[b[0]], 4 * __popc(peers));
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const unsigned peers = __match_any_sync(0xffffffffu, b[r]);
          if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&counts[b[r]], __popc(peers));
        }
      }
    } else {
      const int row = row0 + (int)(4 * q);
      if (__all_sync(0xffffffffu, same)) {
        // one atomic per distinct bin of the warp claims 4 records per lane of the group
        const unsigned peers = __match_any_sync(0xffffffffu, b[0]);
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (valid && leader == lane) first = atomicAdd(&counts[b[0]], 4 * __popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (valid) {
          const int64_t slot = (int64_t)bbase[b[0]] + first + 4 * __popc(peers & ((1u << lane) - 1u));
          if (slot + 4 <= (int64_t)bbase[b[0] + 1] && slot + 4 <= nrec) {
#pragma unroll
            for (int r = 0; r < 4; r++)
              brec[slot + r] = make_float4(f[3 * r], f[3 * r + 1], f[3 * r + 2], __int_as_float(row + r));
          } else {
            *overflow = 1;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const unsigned peers = __match_any_sync(0xffffffffu, b[r]);
          const int leader = __ffs(peers) - 1;
          int first = 0;
          if (valid && leader == lane) first = atomicAdd(&counts[b[r]], __popc(peers));
          first = __shfl_sync(0xffffffffu, first, leader);
          if (valid) {
            const int64_t slot = (int64_t)bbase[b[r]] + first + __popc(peers & ((1u << lane) - 1u));
            if (slot < (int64_t)bbase[b[r] + 1] && slot < nrec)
              brec[slot] = make_float4(f[3 * r], f[3 * r + 1], f[3 * r + 2], __int_as_float(row + r));
            else
              *overflow = 1;
          }
        }
      }
    }
  }
  // the last np % 4 particles
  if (blockIdx.x == 0 && threadIdx.x < (np & 3)) {
    const int64_t n = (nq << 2) + threadIdx.x;
    float x[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float v = vel[3 * n + c], p = pos[3 * n + c];
      kick_drift_wrap1<F64>(p, v, acc[3 * n + c], mh, dtf, dt);
      vel[3 * n + c] = v;
      pos[3 * n + c] = p;
      x[c] = p;
    }
    const int bb = bin_of(x[0], x[1], x[2], Nf, NB, 0, NB);
    const int first = atomicAdd(&counts[bb], 1);
    if (DIRECT) {
      const int64_t slot = (int64_t)bbase[bb] + first;
      if (slot < (int64_t)bbase[bb + 1] && slot < nrec)
        brec[slot] = make_float4(x[0], x[1], x[2], __int_as_float(row0 + (int)n));
      else
        *overflow = 1;
    }
  }
}

// capacities for the next direct scatter from this step's fill: fill + 1/8 + 32 records per bin (sum <= rec capacity)
__global__ void __launch_bounds__(256) bin_caps_kernel(const int *__restrict__ fill, int nbins, int64_t np,
                                                       int *__restrict__ caps) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbins) return;
  int f = 0;
  if (b < nbins) f = (int)min((int64_t)max(fill[b], 0), np);
  caps[b] = b < nbins ? f + (f >> 3) + 32 : 0;
}

// fall-back of the direct scatter: adopt the exact offsets and restart the cursors
__global__ void __launch_bounds__(256) bin_adopt_kernel(const int *__restrict__ overflow, const int *__restrict__ exact,
                                                        int nbins, int *__restrict__ base, int *__restrict__ fill) {
  if (*overflow == 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbins) return;
  base[b] = exact[b];
  fill[b] = 0;
}

// after the scan: list the extra parts of the bins that hold more than BIN_PART particles
__global__ void __launch_bounds__(256) bin_heavy_list_kernel(const int *__restrict__ fill, int nbins,
                                                             int *__restrict__ heavy_count, int2 *__restrict__ heavy,
                                                             int heavy_cap) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int extra = (fill[b] - 1) / BIN_PART;
  if (extra <= 0) return;
  const int base = atomicAdd(heavy_count, extra);
  for (int e = 0; e < extra; e++)
    if (base + e < heavy_cap) heavy[base + e] = make_int2(b, e + 1);
}

// pass 2: every particle claims the next record of its bin: slot = base[bin] + fill[bin]++ (fill ends as the count)
__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ pos, int64_t np, int N, int NB,
                                                          int x0, int NBX, int *__restrict__ fill,
                                                          const int *__restrict__ base, float4 *__restrict__ brec,
                                                          const int *__restrict__ only_if) {
  if (only_if && *only_if == 0) return;
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    int b = -1 - lane;
    if (n < np) {
      x = __ldg(&pos[3 * n]); y = __ldg(&pos[3 * n + 1]); z = __ldg(&pos[3 * n + 2]);
      b = bin_of(x, y, z, Nf, NB, x0, NBX);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    int first = 0;
    if (b >= 0 && leader == lane) first = atomicAdd(&fill[b], __popc(peers));
    first = __shfl_sync(0xffffffffu, first, leader);
    if (b >= 0) {
      const int slot = base[b] + first + __popc(peers & ((1u << lane) - 1u));
      brec[slot] = make_float4(x, y, z, __int_as_float((int)n));
    }
  }
}

// ------------------------------------------------------------------- particle arrays kept in bin order
// Round 2, after measuring what the shadow binning costs once the source order has decayed (per-particle random
// velocities: the rows of a bin's particles are spread over the array, the interpolation's v / a accesses by source row
// become one 32-byte sector per lane -- 5.0 -> 7.3 ms at 512^3): the time loop keeps position / velocity /
// acceleration THEMSELVES in bin order and re-sorts them every step, fused with the first half of the leapfrog:
//   pass 1 (PASS = 1)  read x, v, a; form the new position in registers; count it in its bin.  Nothing is written.
//   scan               base[] = first row of every bin (tight: no gaps, the arrays stay [np, 3]).
//   pass 2 (PASS = 2)  read x, v, a (+ id) again, redo the same arithmetic (bit-identical), claim the next row of the
//                      bin and write x', v' (and the particle's id = its row in the reference's order) there.
// 36 + 40 B read, 28 B written per particle, all streaming; the deposit then reads 12 B and the interpolation 24 + 24 B
// per particle in bin order -- no binned copy, no source-row indirection anywhere.  The reference's particle order is
// restored from the ids when somebody asks for it (utils.reference_order: snapshots, NumPy callers, reorder_particles).
template <bool F64, int PASS>
__global__ void __launch_bounds__(256) step_sort_kernel(const float *__restrict__ pos, const float *__restrict__ vel,
                                                        const float *__restrict__ acc, const int *__restrict__ ids,
                                                        int64_t np, float half_dt, double dt, int N, int NB,
                                                        int *__restrict__ cnt, const int *__restrict__ bbase,
                                                        float *__restrict__ pos_out, float *__restrict__ vel_out,
                                                        int *__restrict__ ids_out) {
  const float dtf = (float)dt, mh = -half_dt, Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nq = np >> 2;
  const float4 *p4 = reinterpret_cast<const float4 *>(pos);
  const float4 *v4 = reinterpret_cast<const float4 *>(vel);
  const float4 *a4 = reinterpret_cast<const float4 *>(acc);
  const int64_t wstride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
  for (int64_t wb = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; wb < nq; wb += wstride) {
    const int64_t q = wb + lane;
    const bool valid = q < nq;
    float f[12], v[12];
    int b[4] = {-1 - lane, -1 - lane, -1 - lane, -1 - lane};
    int id[4] = {0, 0, 0, 0};
    if (valid) {
      float a[12];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 P = __ldg(&p4[3 * q + c]), V = __ldg(&v4[3 * q + c]), A = __ldg(&a4[3 * q + c]);
        f[4 * c] = P.x; f[4 * c + 1] = P.y; f[4 * c + 2] = P.z; f[4 * c + 3] = P.w;
        v[4 * c] = V.x; v[4 * c + 1] = V.y; v[4 * c + 2] = V.z; v[4 * c + 3] = V.w;
        a[4 * c] = A.x; a[4 * c + 1] = A.y; a[4 * c + 2] = A.z; a[4 * c + 3] = A.w;
      }
#pragma unroll
      for (int c = 0; c < 12; c++) {
        kick_drift_wrap1<F64>(f[c], v[c], a[c], mh, dtf, dt);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) b[r] = bin_of(f[3 * r], f[3 * r + 1], f[3 * r + 2], Nf, NB, 0, NB);
      if (PASS == 2) {
        if (ids) {
          const int4 I = __ldg(reinterpret_cast<const int4 *>(ids) + q);
          id[0] = I.x; id[1] = I.y; id[2] = I.z; id[3] = I.w;
        } else {
#pragma unroll
          for (int r = 0; r < 4; r++) id[r] = (int)(4 * q) + r;
        }
      }
    }
    const bool same = b[0] == b[1] && b[1] == b[2] && b[2] == b[3];
    if (__all_sync(0xffffffffu, same)) {
      // the common case in a bin-ordered array: one atomic per distinct bin of the warp
      const unsigned peers = __match_any_sync(0xffffffffu, b[0]);
      const int leader = __ffs(peers) - 1;
      int first = 0;
      if (valid && leader == lane) first = atomicAdd(&cnt[b[0]], 4 * __popc(peers));
      if (PASS == 2) {
        first = __shfl_sync(0xffffffffu, first, leader);
        if (valid) {
          const size_t slot = (size_t)bbase[b[0]] + first + 4 * __popc(peers & ((1u << lane) - 1u));
          // four consecutive rows: 48 bytes each of position and velocity, 16-byte aligned (slot is a multiple of 4
          // only by chance, so plain float stores)
#pragma unroll
          for (int c = 0; c < 12; c++) {
            pos_out[3 * slot + c] = f[c];
            vel_out[3 * slot + c] = v[c];
          }
#pragma unroll
          for (int r = 0; r < 4; r++) ids_out[slot + r] = id[r];
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const unsigned peers = __match_any_sync(0xffffffffu, b[r]);
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (valid && leader == lane) first = atomicAdd(&cnt[b[r]], __popc(peers));
        if (PASS == 2) {
          first = __shfl_sync(0xffffffffu, first, leader);
          if (valid) {
            const size_t slot = (size_t)bbase[b[r]] + first + __popc(peers & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < 3; c++) {
              pos_out[3 * slot + c] = f[3 * r + c];
              vel_out[3 * slot + c] = v[3 * r + c];
            }
            ids_out[slot] = id[r];
          }
        }
      }
    }
  }
  // the last np % 4 particles
  if (blockIdx.x == 0 && threadIdx.x < (np & 3)) {
    const int64_t n = (nq << 2) + threadIdx.x;
    float x[3], w[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float vv = vel[3 * n + c], p = pos[3 * n + c];
      kick_drift_wrap1<F64>(p, vv, acc[3 * n + c], mh, dtf, dt);
      x[c] = p;
      w[c] = vv;
    }
    const int bb = bin_of(x[0], x[1], x[2], Nf, NB, 0, NB);
    const int first = atomicAdd(&cnt[bb], 1);
    if (PASS == 2) {
      const size_t slot = (size_t)bbase[bb] + first;
#pragma unroll
      for (int c = 0; c < 3; c++) { pos_out[3 * slot + c] = x[c]; vel_out[3 * slot + c] = w[c]; }
      ids_out[slot] = ids ? ids[n] : (int)n;
    }
  }
}

// The same two passes when the INPUT arrays are already in bin order (every step but the first after a reorder): one
// CTA per source bin.  A particle moves at most one cell per step, so it stays in its bin or goes to one of the 26
// neighbours: the CTA counts its particles per destination bin in shared memory (pass 1: <= 27 global atomics per CTA
// instead of one per particle) and, in pass 2, sorts its particles in shared memory by (destination bin, 2^3-cell
// micro-block in Morton order), claims ONE contiguous block per destination bin and writes the sorted runs with
// coalesced stores.  The rows of a bin are then ~90 % one run in micro-block order (plus short runs from the
// neighbours): consecutive particles sit in neighbouring cells -- what the bank spreading of the deposit / force tiles
// is made for (deposit 3.7 -> 3.1 ms, interpolation 5.7 -> 4.2 ms at 512^3 against the arrival order of the
// one-atomic-per-particle scatter) -- at none of the cost of a global sort on the finer key (measured: 6.5 ms, the
// 67 MB of counters are atomic-latency bound).  A particle that jumped further (never under the Courant condition)
// takes a slow path through global atomics.
constexpr int SL_CHUNK = 640;                 // particles per round of a CTA: a whole bin at the mean density + 5 sigma
constexpr int SL_R = 3;                       // = ceil(SL_CHUNK / 256) particles per thread
// sort keys of a round: the nself (64 micro-blocks; 512 cells in an experiment) keys of the particles that stay in the
// bin, then the 26 neighbouring bins
constexpr int SL_F3 = 3 * SL_CHUNK + 4;       // floats of one staged [n][3] array (+ 16-byte alignment slack)
constexpr int SL_I = SL_CHUNK + 4;            // ints of the staged ids
constexpr int SL_STAGE = 3 * SL_F3 + SL_I;    // one stage: position, velocity, acceleration, ids (25664 bytes)
constexpr size_t SL_SMEM = 2 * SL_STAGE * sizeof(float);
static_assert((SL_F3 * 4) % 16 == 0 && (SL_I * 4) % 16 == 0, "stage arrays must stay 16-byte aligned");

using sl_barrier = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ int axis_delta(int nb, int src, int NB) {
  // 0 / 1 / 2 for destination bin src - 1 / src / src + 1 (periodic with period NB), -1 otherwise
  int d = nb - src;
  d += d > 1 ? -NB : (d < -1 ? NB : 0);
  return (unsigned)(d + 1) <= 2u ? d + 1 : -1;
}

// One bulk copy (TMA, completion on the stage's mbarrier) of the 16-byte aligned superset of src[first, first + n);
// what the aligned superset would read beyond the end of the array (< 16 bytes, last round of the last bin only) is
// left out.  Called by one thread.
template <class T>
__device__ __forceinline__ void sl_fetch(T *dst, const T *src, int64_t first, int64_t n, int64_t total,
                                         sl_barrier &bar) {
  constexpr int64_t q = 16 / sizeof(T);   // elements per 16 bytes
  const int64_t a0 = first & ~(q - 1);
  int64_t a1 = (first + n + q - 1) & ~(q - 1);
  if (a1 > total) {
    a1 = total & ~(q - 1);
    for (int64_t t = a1 > a0 ? a1 : a0; t < total; t++) dst[t - a0] = src[t];
  }
  if (a1 > a0) cuda::memcpy_async(dst, src + a0, cuda::aligned_size_t<16>(sizeof(T) * (size_t)(a1 - a0)), bar);
}

// Pass 2 of the local sort (pass 1 = the count pass of step_sort_kernel).  Persistent CTAs, each walking over bins
// blockIdx.x, blockIdx.x + gridDim.x, ...; the rows of the NEXT bin (position, velocity, acceleration, ids: four bulk
// copies, 25 KB) are in flight on the other stage's mbarrier while the current bin is sorted, so the DRAM latency of a
// bin is hidden behind the sort of the one before.  A stage that has been read into registers becomes the staging area
// of the sorted output, which leaves the CTA as contiguous float runs (one run per destination bin).
//
// SLAB = true is the same sort for the particles of an x-slab (psc_sort_by_bin_slab): the kick + drift + wrap and the
// migration have already been applied in place (so nothing is recomputed and the acceleration is not read), ids are
// 64-bit, bins are those of the slab [x0, x0 + 8 NBX) (periodic in x only when the slab is the whole box), and only the
// first `rows` rows are described by the source table (the migration may have shortened or extended the arrays; rows
// that hold arrivals or moved tail particles are wherever the migration put them: they take the slow path).
template <bool F64, int nself, bool SLAB>
__global__ void __launch_bounds__(256) step_sort_local_kernel(
    const float *__restrict__ pos, const float *__restrict__ vel, const float *__restrict__ acc,
    const typename std::conditional<SLAB, long long, int>::type *__restrict__ ids, int64_t np, int64_t rows,
    const int *__restrict__ base_src, const int *__restrict__ fill_src, float half_dt, double dt, int N, int NB, int x0,
    int NBX, int nbins, int *__restrict__ cnt, const int *__restrict__ base_dst, float *__restrict__ pos_out,
    float *__restrict__ vel_out, typename std::conditional<SLAB, long long, int>::type *__restrict__ ids_out) {
  using IdT = typename std::conditional<SLAB, long long, int>::type;
  extern __shared__ __align__(128) float sl_smem[];
  constexpr int nkeys = nself + 26, per = (nkeys + 31) / 32;
  __shared__ int hist[32 * per];
  __shared__ int s_dst[27], s_wsum[8], s_near;
  __shared__ sl_barrier bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float dtf = (float)dt, mh = -half_dt, Nf = (float)N;
  if (tid == 0) {
    init(&bar[0], 256);
    init(&bar[1], 256);
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
  }
  for (int t = tid; t < 32 * per; t += 256) hist[t] = 0;
  // the rounds of this CTA: (bin, first particle of the round within the bin), empty bins skipped
  // particles of source bin sb that are still there: its rows below `rows`
  auto live = [&](int sb) {
    if (!SLAB) return __ldg(&fill_src[sb]);   // the single-domain arrays never change length
    const int64_t left = rows - (int64_t)__ldg(&base_src[sb]);
    return (int)max((int64_t)0, min((int64_t)__ldg(&fill_src[sb]), left));
  };
  const int NBXp = NBX == NB ? NB : 0x40000000;   // period of the bin index along x (none inside a slab)
  // The rounds of this CTA are described by thread 0, one round ahead (it needs them to issue the copies): the source
  // bin, its particles of the round, their first row and the bin's coordinates.  Everybody else reads the descriptor
  // from shared memory after the stage's mbarrier -- no per-thread divisions, no per-thread loads of the bin table.
  struct Round { int b, n, sbi, sbj, sbk, c0, nb; long long first; };
  __shared__ Round s_round[2];
  auto describe = [&](int rb, int rc0, int rnb, Round &R) {   // thread 0; rb >= nbins: no such round
    R.b = rb; R.c0 = rc0; R.nb = rnb;
    if (rb < nbins) {
      R.n = min(rnb - rc0, SL_CHUNK);
      R.first = (long long)__ldg(&base_src[rb]) + rc0;
      R.sbk = rb % NB; R.sbj = (rb / NB) % NB; R.sbi = rb / (NB * NB);
    }
  };
  auto fetch = [&](const Round &R, int stage) {   // thread 0
    float *S = sl_smem + stage * SL_STAGE;
    const int64_t first = R.first, n = R.n;
    sl_fetch(S, pos, 3 * first, 3 * n, 3 * np, bar[stage]);
    sl_fetch(S + SL_F3, vel, 3 * first, 3 * n, 3 * np, bar[stage]);
    if (SLAB) {   // 64-bit ids in the (unused) acceleration area
      if (ids) sl_fetch(reinterpret_cast<IdT *>(S + 2 * SL_F3), ids, first, n, np, bar[stage]);
    } else {
      sl_fetch(S + 2 * SL_F3, acc, 3 * first, 3 * n, 3 * np, bar[stage]);
      if (ids) sl_fetch(reinterpret_cast<IdT *>(S + 3 * SL_F3), ids, first, n, np, bar[stage]);
    }
  };
  if (tid == 0) {
    int b = blockIdx.x, nb = 0;
    while (b < nbins && (nb = live(b)) == 0) b += gridDim.x;
    describe(b, 0, nb, s_round[0]);
    if (b < nbins) fetch(s_round[0], 0);
  }
  __syncthreads();
  for (int it = 0;; it++) {
    const int stage = it & 1;
    if (s_round[stage].b >= nbins) break;
    if (tid == 0) {
      // the round after this one
      const Round &R = s_round[stage];
      int b2 = R.b, c2 = R.c0 + SL_CHUNK, nb2 = R.nb;
      if (c2 >= nb2) {
        c2 = 0;
        b2 += gridDim.x;
        while (b2 < nbins && (nb2 = live(b2)) == 0) b2 += gridDim.x;
      }
      describe(b2, c2, nb2, s_round[stage ^ 1]);
      if (b2 < nbins) {
        cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);   // the other stage was last written as an output area
        fetch(s_round[stage ^ 1], stage ^ 1);
      }
    }
    bar[stage].arrive_and_wait();
    float *S = sl_smem + stage * SL_STAGE;
    const int64_t first = s_round[stage].first;
    const int n = s_round[stage].n;
    const int off3 = (int)((3 * first) & 3), off1 = (int)(first & (16 / sizeof(IdT) - 1));
    const int sbk = s_round[stage].sbk, sbj = s_round[stage].sbj, sbi = s_round[stage].sbi;
    float f[SL_R][3], v[SL_R][3];
    IdT id[SL_R];
    int key[SL_R], dbin[SL_R], rank[SL_R], dst[SL_R];
#pragma unroll
    for (int r = 0; r < SL_R; r++) {
      const int m = tid + 256 * r;
      key[r] = -2;   // no particle
      if (m < n) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int o = off3 + 3 * m + c;
          if (SLAB) {
            f[r][c] = S[o];
            v[r][c] = S[SL_F3 + o];
          } else {
            float vv = S[SL_F3 + o], p = S[o];
            kick_drift_wrap1<F64>(p, vv, S[2 * SL_F3 + o], mh, dtf, dt);
            f[r][c] = p;
            v[r][c] = vv;
          }
        }
        id[r] = ids ? reinterpret_cast<const IdT *>(S + (SLAB ? 2 : 3) * SL_F3)[off1 + m] : (IdT)(first + m);
        // the cell, clamped into the slab like bin_of does
        const int i = min(max((int)(f[r][0] * Nf) - x0, 0), 8 * NBX - 1), j = min(max((int)(f[r][1] * Nf), 0), N - 1),
                  k = min(max((int)(f[r][2] * Nf), 0), N - 1);
        const int di = axis_delta(i >> 3, sbi, NBXp), dj = axis_delta(j >> 3, sbj, NB), dk = axis_delta(k >> 3, sbk, NB);
        dbin[r] = ((i >> 3) * NB + (j >> 3)) * NB + (k >> 3);
        if ((di | dj | dk) >= 0) {
          const int ii = (i >> 1) & 3, jj = (j >> 1) & 3, kk = (k >> 1) & 3;
          const int mb = ((ii >> 1) << 5) | ((jj >> 1) << 4) | ((kk >> 1) << 3) | ((ii & 1) << 2) | ((jj & 1) << 1) | (kk & 1);
          dst[r] = (di * 3 + dj) * 3 + dk;
          const int self = nself == MB_PER_BIN ? mb : ((i & 7) << 6) | ((j & 7) << 3) | (k & 7);
          key[r] = dst[r] == 13 ? self : nself + dst[r] - (dst[r] > 13);
          rank[r] = atomicAdd(&hist[key[r]], 1);
        } else {
          key[r] = -1;   // further than a neighbouring bin (never under the Courant condition): slow path
        }
      }
    }
    __syncthreads();   // counts complete; the stage has been read by everybody
    if (warp == 0) {
      // exclusive scan of the nself + 26 counts: a run of `per` consecutive keys per lane
      int h[per], sum = 0;
#pragma unroll
      for (int q = 0; q < per; q++) {
        h[q] = hist[lane * per + q];     // the keys beyond nkeys are never counted: zero
        sum += h[q];
      }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int run = incl - sum;
#pragma unroll
      for (int q = 0; q < per; q++) {
        hist[lane * per + q] = run;
        run += h[q];
      }
      if (lane == 31) s_near = incl;
    }
    __syncthreads();
    // the sorted round: nstay stayers, then the leavers by bin (read here: hist is cleared again by whoever finishes
    // the write-out first)
    const int nnear = s_near, nstay = hist[nself];
    float *opos = S, *ovel = S + SL_F3;
    IdT *oid = reinterpret_cast<IdT *>(S + 2 * SL_F3);
    unsigned char *od = reinterpret_cast<unsigned char *>(S + 3 * SL_F3);
    if (tid < 27) {
      // one contiguous block of destination rows per destination bin: s_dst[d] + (index in the sorted round) = row.
      // The round trip of the global atomic runs under the staging of the sorted records below.
      const int k0 = tid == 13 ? 0 : nself + tid - (tid > 13);              // first key of destination tid
      const int k1 = tid == 13 ? nself : k0 + 1;
      const int head = hist[k0], tot = (k1 < nself + 26 ? hist[k1] : s_near) - head;
      int row0 = 0;
      if (tot > 0) {
        const int di = tid / 9, dj = (tid / 3) % 3, dk = tid % 3;
        const int bi = (sbi + di - 1 + NBX) % NBX, bj = (sbj + dj - 1 + NB) % NB, bk = (sbk + dk - 1 + NB) % NB;
        const int db = (bi * NB + bj) * NB + bk;
        row0 = __ldg(&base_dst[db]) + atomicAdd(&cnt[db], tot) - head;
      }
      s_dst[tid] = row0;
    }
#pragma unroll
    for (int r = 0; r < SL_R; r++) {
      if (key[r] >= 0) {
        const int si = hist[key[r]] + rank[r];
#pragma unroll
        for (int c = 0; c < 3; c++) { opos[3 * si + c] = f[r][c]; ovel[3 * si + c] = v[r][c]; }
        oid[si] = id[r];
        od[si] = (unsigned char)dst[r];
      } else if (key[r] == -1) {
        const size_t slot = (size_t)__ldg(&base_dst[dbin[r]]) + atomicAdd(&cnt[dbin[r]], 1);
#pragma unroll
        for (int c = 0; c < 3; c++) { pos_out[3 * slot + c] = f[r][c]; vel_out[3 * slot + c] = v[r][c]; }
        ids_out[slot] = id[r];
      }
    }
    __syncthreads();
    {
      const int64_t g0 = (int64_t)3 * s_dst[13];
      for (int t = tid; t < 3 * nstay; t += 256) {           // the stayers: one contiguous run of floats
        pos_out[g0 + t] = opos[t];
        vel_out[g0 + t] = ovel[t];
      }
      for (int t = 3 * nstay + tid; t < 3 * nnear; t += 256) {
        const int64_t g = (int64_t)3 * s_dst[od[t / 3]] + t;    // consecutive t: consecutive floats of a run
        pos_out[g] = opos[t];
        vel_out[g] = ovel[t];
      }
      IdT *io = ids_out + s_dst[13];
      for (int t = tid; t < nstay; t += 256) io[t] = oid[t];
      for (int t = nstay + tid; t < nnear; t += 256) ids_out[s_dst[od[t]] + t] = oid[t];
    }
    for (int t = tid; t < 32 * per; t += 256) hist[t] = 0;   // all of them: the scan wrote the unused tail too
    // this stage is the destination of the next bulk copy: every thread orders its own generic-proxy accesses to the
    // output area before the async proxy (the documented pattern: fence in every thread, barrier, then the copy)
    cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);
    __syncthreads();   // output area, hist and s_dst are free again
  }
}

// ---- utils.reorder_particles (utils.py:1019-1075) for bin-ordered arrays: the Morton order without moving a particle.
// The arrays of the time loop are re-sorted into bins every step, so the reference's periodic Morton reorder changes
// nothing for the kernels; what it does change is the ROW every particle has in the reference from then on.  That is
// the ids' business: ids[n] = rank of particle n in the order of its Morton key.  An 8^3-cell bin of a power-of-two
// mesh is one contiguous range of Morton keys, so the rank is (particles of the bins before it in Z order) + (rank of
// the key within the bin): an exclusive scan of the bin counts in Z order and one block-wide radix sort of ~512 keys
// per bin in shared memory -- ~1 ms at 512^3 instead of a 134 M-key device radix sort and three gathers (34 ms).
// Equal keys (identical 21-bit coordinates) keep their current row order; the reference's np.argsort is unstable there.
__device__ __forceinline__ unsigned spread10(unsigned x) {
  x &= 0x3FFu;
  x = (x | x << 16) & 0x30000FFu;
  x = (x | x << 8) & 0x300F00Fu;
  x = (x | x << 4) & 0x30C30C3u;
  x = (x | x << 2) & 0x9249249u;
  return x;
}

__device__ __forceinline__ int zorder_of_bin(int b, int NB) {
  const int bk = b % NB, bj = (b / NB) % NB, bi = b / (NB * NB);
  return (int)(spread10(bi) << 2 | spread10(bj) << 1 | spread10(bk));
}

__global__ void __launch_bounds__(256) zorder_fill_kernel(const int *__restrict__ fill, int nbins, int NB,
                                                          int *__restrict__ zfill) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbins) zfill[zorder_of_bin(b, NB)] = fill[b];
  if (b == nbins) zfill[nbins] = 0;
}

template <int ITEMS>
__global__ void __launch_bounds__(256) morton_rank_kernel(const float *__restrict__ pos, const int *__restrict__ base,
                                                          const int *__restrict__ fill, const int *__restrict__ zscan,
                                                          int NB, int lo, int key_bits, int *__restrict__ ids_out,
                                                          int *__restrict__ too_big) {
  using Sort = cub::BlockRadixSort<unsigned long long, 256, ITEMS, int>;   // 4-bit digits: 8.3 ms at 512^3, 6-bit: 10.2
  __shared__ typename Sort::TempStorage tmp;
  const int b = blockIdx.x, n = fill[b];
  if (n <= lo) return;
  if (n > 256 * ITEMS) {
    if (threadIdx.x == 0) *too_big = 1;
    return;
  }
  const int beg = base[b];
  unsigned long long key[ITEMS];
  int val[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int m = threadIdx.x * ITEMS + i;
    val[i] = m;
    key[i] = ~0ull;
    if (m < n) {
      const size_t g = (size_t)beg + m;
      key[i] = morton_key(__ldg(&pos[3 * g]), __ldg(&pos[3 * g + 1]), __ldg(&pos[3 * g + 2]));
    }
  }
  Sort(tmp).Sort(key, val, 0, key_bits);   // the bits above key_bits are the bin's: equal for all its particles
  const int first = zscan[zorder_of_bin(b, NB)];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    if (val[i] < n) ids_out[beg + val[i]] = first + threadIdx.x * ITEMS + i;
}

// slab flavour of the sort (the kick + drift + wrap is a separate pass there: migration sits in between): pass 2 of a
// counting sort of (position, velocity, 64-bit id) into bin order; pass 1 is bin_count_kernel
__global__ void __launch_bounds__(256) sort_scatter_kernel(const float *__restrict__ pos, const float *__restrict__ vel,
                                                           const int64_t *__restrict__ ids, int64_t np, int N, int NB,
                                                           int x0, int NBX, int *__restrict__ fill,
                                                           const int *__restrict__ bbase, float *__restrict__ pos_out,
                                                           float *__restrict__ vel_out, int64_t *__restrict__ ids_out) {
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    int b = -1 - lane;
    if (n < np) {
      x = __ldg(&pos[3 * n]); y = __ldg(&pos[3 * n + 1]); z = __ldg(&pos[3 * n + 2]);
      b = bin_of(x, y, z, Nf, NB, x0, NBX);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    int first = 0;
    if (b >= 0 && leader == lane) first = atomicAdd(&fill[b], __popc(peers));
    first = __shfl_sync(0xffffffffu, first, leader);
    if (b >= 0) {
      const size_t slot = (size_t)bbase[b] + first + __popc(peers & ((1u << lane) - 1u));
      pos_out[3 * slot] = x; pos_out[3 * slot + 1] = y; pos_out[3 * slot + 2] = z;
      vel_out[3 * slot] = __ldg(&vel[3 * n]); vel_out[3 * slot + 1] = __ldg(&vel[3 * n + 1]);
      vel_out[3 * slot + 2] = __ldg(&vel[3 * n + 2]);
      ids_out[slot] = ids[n];
    }
  }
}

// out[ids[n]] = in[n] for three-column arrays: back to the reference's particle order
__global__ void __launch_bounds__(256) scatter3_by_id_kernel(const int *__restrict__ ids, const float *__restrict__ in,
                                                             float *__restrict__ out, int64_t np) {
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    const size_t d = 3 * (size_t)ids[n];
    out[d] = in[3 * n]; out[d + 1] = in[3 * n + 1]; out[d + 2] = in[3 * n + 2];
  }
}

// ------------------------------------------------------------------------------------- deposit
// particles [beg, end) of bin b; shared = other CTAs deposit into the same bin (its interior cells need atomics too)
// Where a kernel finds particle n of a bin.  SHADOW (SORTED = false): `src` is the binned copy float4[.] = (x, y, z, source
// row), the particle arrays themselves are in another order and results go to `row`.  SORTED = true: `src` is the
// position array [np, 3] itself, stored in bin order (csrc/sorted.cu): the row IS n.
template <bool SORTED>
__device__ __forceinline__ void load_particle(const void *__restrict__ src, int n, float &px, float &py, float &pz,
                                              int &row) {
  if (SORTED) {
    const float *p = reinterpret_cast<const float *>(src) + 3 * (size_t)n;
    px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
    row = n;
  } else {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(src) + n);
    px = r.x; py = r.y; pz = r.z;
    row = __float_as_int(r.w);
  }
}

// scale / f1: the density rescale and the slope of rhs_poisson's affine map (solver.py:114-116, 444-449) applied to the
// bin's sums before they leave the CTA; the caller has initialised rho with the map's constant f2, so that
// rho = f1 * (scale * sum) + f2 needs no further pass over the grid (a cell fed by several bins gets f2 once and the
// bins' scaled sums through REDs).
template <int SCHEME, bool SORTED>
__device__ __forceinline__ void deposit_bin_range(float (*tiles)[BD_TILE], const void *__restrict__ brec, int b,
                                                  int beg, int end, bool shared, int N, int NB, int x0, int xoff,
                                                  int nxa, float scale, float f1, float f2,
                                                  float *__restrict__ rho) {
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
    int row_unused;
    if (valid) load_particle<SORTED>(brec, n, px, py, pz, row_unused);
    int i, j, k;
    float wx[3], wy[3], wz[3];
    axis_weights_bin<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights_bin<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights_bin<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
    // in [1, 8] for every particle of this bin (clamped like bin_of); the shifted CIC window reaches 9
    constexpr int TMAX = SCHEME == PSC_CIC ? BB + 1 : BB;
    const int t0 = min(max(i - oi, 1), TMAX), t1 = min(max(j - oj, 1), TMAX), t2 = min(max(k - ok, 1), TMAX);
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
          if (!stencil_uses<SCHEME>(a, e, g)) continue;
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
      v = f1 * (scale * v);
      const int gi = wrap(pl0 + a, nxa), gj = wrap(oj + e, N);
      float *dst = rho + ((size_t)gi * N + gj) * N + gk;
      const bool mine = !shared && kin && a >= 2 && a <= BT - 3 && e >= 2 && e <= BT - 3;
      if (mine) *dst = v + f2;
      else if (v != 0.0f) atomicAdd(dst, v);
    }
  }
}

template <int SCHEME, bool SORTED>
__global__ void __launch_bounds__(BD_WARPS * 32) deposit_binned_kernel(const void *__restrict__ brec,
                                                                       const int *__restrict__ base,
                                                                       const int *__restrict__ fill, int N, int NB,
                                                                       int x0, int xoff, int nxa, float scale, float f1,
                                                                       float f2, float *__restrict__ rho) {
  __shared__ float tiles[BD_WARPS][BD_TILE];
  const int b = blockIdx.x;
  const int beg = base[b], end = beg + fill[b];
  if (beg == end) return;  // rho was initialised (f2) by the caller
  deposit_bin_range<SCHEME, SORTED>(tiles, brec, b, beg, min(end, beg + BIN_PART), end - beg > BIN_PART, N, NB, x0, xoff,
                                    nxa, scale, f1, f2, rho);
}

// the parts beyond BIN_PART particles of the heavy bins (persistent CTAs over BinLayout::heavy)
template <int SCHEME, bool SORTED>
__global__ void __launch_bounds__(BD_WARPS * 32) deposit_heavy_kernel(const void *__restrict__ brec,
                                                                      const int *__restrict__ base,
                                                                      const int *__restrict__ fill,
                                                                      const int *__restrict__ heavy_count,
                                                                      const int2 *__restrict__ heavy, int heavy_cap,
                                                                      int N, int NB, int x0, int xoff, int nxa,
                                                                      float scale, float f1, float f2,
                                                                      float *__restrict__ rho) {
  __shared__ float tiles[BD_WARPS][BD_TILE];
  const int nitems = min(*heavy_count, heavy_cap);
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int2 w = heavy[it];
    const int beg = base[w.x] + w.y * BIN_PART, end = min(base[w.x] + fill[w.x], beg + BIN_PART);
    deposit_bin_range<SCHEME, SORTED>(tiles, brec, w.x, beg, end, true, N, NB, x0, xoff, nxa, scale, f1, f2, rho);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------- interpolation
constexpr int BI_THREADS = 128;

template <int SCHEME>
__global__ void __launch_bounds__(BI_THREADS) interp_kick4_binned_kernel(
    const float4 *__restrict__ force4, const float4 *__restrict__ brec,
    const int *__restrict__ base, const int *__restrict__ fill, float *__restrict__ vel, float *__restrict__ accel,
    int N, int NB,
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
  const int beg = base[b] + part * BIN_PART, end = min(base[b] + fill[b], beg + BIN_PART);
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
    axis_weights_bin<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
    axis_weights_bin<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
    axis_weights_bin<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
    constexpr int TMAX = SCHEME == PSC_CIC ? BB + 1 : BB;   // the shifted CIC window reaches tile index 9
    const float4 *c0 = tile + ((min(max(i - oi, 1), TMAX) - 1) * BT + (min(max(j - oj, 1), TMAX) - 1)) * BT + (min(max(k - ok, 1), TMAX) - 1);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const float wxy = wx[a] * wy[e];
#pragma unroll
        for (int g = 0; g < 3; g++) {
          if (!stencil_uses<SCHEME == PSC_NGP ? PSC_TSC : SCHEME>(a, e, g)) continue;   // CIC: 8 of the 27 points
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

// Speculative bin count of the NEXT step's sort (psc_step_sort pass 1) inside the interpolation kernel: the particle's
// position, its velocity after the second half-kick and its new acceleration are all in registers here, so with the
// next time step in hand -- it is known in advance whenever the scale-factor criterion binds (integration.py:329-358
// depends on a(t) only) -- the kernel applies the next kick + drift + wrap (same bits as every other kernel:
// kick_drift_wrap1) and counts the destination bins.  The host uses the counts only if the time step it then computes
// is the predicted one; otherwise the sort counts again.  counts == nullptr: off.
struct BinPredict {
  int *counts;
  float mh, dtf;
  double dt;
  int f64;
};

constexpr int BP_THREADS = 256;  // gradient + interpolation kernel.  Resident CTAs per SM (register bound), 512^3, bin-ordered
                                 // arrays: 5 (48 registers) 4.44 ms, 6 (40 registers) 4.31 ms, 7 (32 registers) 4.52 ms

// TP1 / TP0: row / plane pitch of the float4 force tile.  Measured at 512^3 (Morton order): 10/100 4.41 ms, 11/110 4.62,
// 12/120 4.59, 12/144 4.75, 14/140 4.70, 11/112 5.19, 10/104 5.39 -- the dense tile is the best of those.  Round 2,
// bin-ordered arrays (profiles/r02_exp_inbin_order_x_pitch.txt): micro-block order 10/100 4.44 ms, 16/160 6.21 ms;
// particles in row-major CELL order within the bin (PSC_SORT_KEY=cell), where any 8 consecutive cells of a 16/160 tile
// fall into 8 different 16-byte bank groups: 10/100 4.33 ms, 16/160 4.26 ms -- removing the LDS.128 conflicts buys 4 %,
// less than the finer sort key costs in the sort (+0.17 ms) and the deposit (+0.08 ms).
template <int SCHEME, int ORDER, bool SORTED, int TP1 = BT, int TP0 = BT * BT>
__global__ void __launch_bounds__(BP_THREADS, 6) interp_kick_phi_binned_kernel(
    const float *__restrict__ phi, const float *__restrict__ u, float f, int fr_n,
    const void *__restrict__ brec, const int *__restrict__ base, const int *__restrict__ fill,
    float *__restrict__ vel, float *__restrict__ accel, int N, int NB, int x0, int xoff, int nxa, float half_dt,
    float *__restrict__ maxout, int nbins, const int *__restrict__ heavy_count, const int2 *__restrict__ heavy,
    BinPredict pred) {
  constexpr int H = Reach<ORDER>::H;
  constexpr int PT = BT + 2 * H;  // potential tile edge
  constexpr int PK = 16;          // k pitch of the potential tile: the aligned 16-float window
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
  const int beg = base[b] + part * BIN_PART, end = min(base[b] + fill[b], beg + BIN_PART);
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
      if (item < NITEM) pt4[item] = v[it];
    }
  }
  __syncthreads();
  const float pref = ORDER == 2 ? (float)N : ORDER == 3 ? (float)(0.5 * N) : ORDER == 5 ? (float)(N / 12.0) : (float)(N / 60.0);
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
  __syncthreads();
  const float Nf = (float)N;
  const float mh = -half_dt;
  unsigned ma = 0u, mv = 0u;   // maxima of |.| as bit patterns: orders like the floats and lets a NaN win
  const int lane = threadIdx.x & 31;
  for (int n0 = beg + (threadIdx.x & ~31); n0 < end; n0 += BP_THREADS) {   // warp-uniform trip count
    const int n = n0 + lane;
    int db = -1 - lane;      // destination bin of the particle in the next step (prediction), none
    if (n < end) {
      float px, py, pz;
      int row;
      load_particle<SORTED>(brec, n, px, py, pz, row);
      int i, j, k;
      float wx[3], wy[3], wz[3];
      axis_weights_bin<SCHEME>(px * Nf, N, i, wx[0], wx[1], wx[2]);
      axis_weights_bin<SCHEME>(py * Nf, N, j, wy[0], wy[1], wy[2]);
      axis_weights_bin<SCHEME>(pz * Nf, N, k, wz[0], wz[1], wz[2]);
      constexpr int TMAX = SCHEME == PSC_CIC ? BB + 1 : BB;   // the shifted CIC window reaches tile index 9
      const float4 *c0 = tile + (min(max(i - oi, 1), TMAX) - 1) * TP0 + (min(max(j - oj, 1), TMAX) - 1) * TP1 + (min(max(k - ok, 1), TMAX) - 1);
      float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const float wxy = wx[a] * wy[e];
#pragma unroll
          for (int g = 0; g < 3; g++) {
            if (!stencil_uses<SCHEME == PSC_NGP ? PSC_TSC : SCHEME>(a, e, g)) continue;   // CIC: 8 of the 27 points
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
        float v0 = vp[0] + mh * ax, v1 = vp[1] + mh * ay, v2 = vp[2] + mh * az;
        vp[0] = v0; vp[1] = v1; vp[2] = v2;
        mv = max(mv, max(__float_as_uint(fabsf(v0)), max(__float_as_uint(fabsf(v1)), __float_as_uint(fabsf(v2)))));
        if (SORTED && pred.counts) {
          if (pred.f64) {
            kick_drift_wrap1<true>(px, v0, ax, pred.mh, pred.dtf, pred.dt);
            kick_drift_wrap1<true>(py, v1, ay, pred.mh, pred.dtf, pred.dt);
            kick_drift_wrap1<true>(pz, v2, az, pred.mh, pred.dtf, pred.dt);
          } else {
            kick_drift_wrap1<false>(px, v0, ax, pred.mh, pred.dtf, pred.dt);
            kick_drift_wrap1<false>(py, v1, ay, pred.mh, pred.dtf, pred.dt);
            kick_drift_wrap1<false>(pz, v2, az, pred.mh, pred.dtf, pred.dt);
          }
          db = bin_of(px, py, pz, Nf, NB, 0, NB);
        }
      }
    }
    if (SORTED && pred.counts) {
      const unsigned peers = __match_any_sync(0xffffffffu, db);
      if (db >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&pred.counts[db], __popc(peers));
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

__global__ void __launch_bounds__(256) fill_kernel(float *__restrict__ x, int64_t n, float v) {
  const int64_t n4 = n >> 2;
  float4 *x4 = reinterpret_cast<float4 *>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    x4[i] = make_float4(v, v, v, v);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) x[(n4 << 2) + threadIdx.x] = v;
}

}  // namespace psc

using namespace psc;

extern "C" {

static bool slab_ok(int N, int x0, int nxl) {
  return N >= BB && (N % BB) == 0 && N <= 32767 && nxl >= BB && (nxl % BB) == 0 && x0 >= 0 && x0 + nxl <= N;
}

size_t psc_bin_workspace_bytes_slab(int64_t np, int N, int nxl) {
  if (np < 0 || !slab_ok(N, 0, nxl)) return 0;
  const int64_t nbins = (int64_t)(nxl / BB) * (N / BB) * (N / BB);
  return 4 * a256(sizeof(int) * (nbins + 1)) + a256(sizeof(float4) * (size_t)rec_capacity(np, nbins)) + 256 +
         a256(sizeof(int2) * (size_t)(np / BIN_PART + 1)) + a256(scan_tmp_bytes(nbins + 1)) + 256;
}
size_t psc_bin_workspace_bytes(int64_t np, int N) { return psc_bin_workspace_bytes_slab(np, N, N); }

static int scan_bins(const BinLayout &L, const int *in, int *out, cudaStream_t st) {
  size_t bytes = L.cub_bytes;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(L.cub_tmp, bytes, in, out, (int)(L.nbins + 1), st);
  count_launch(2);
  if (e != cudaSuccess) {
    set_error("psc_bin_particles: cub scan failed: %s", cudaGetErrorString(e));
    return PSC_ERR_CUDA;
  }
  return PSC_OK;
}

// exact binning from the per-bin counts in L.counts: base = scan(counts), scatter with fill as the cursor, heavy list
static int finish_exact(const float *pos, int64_t np, int N, int x0, const BinLayout &L, cudaStream_t st) {
  int rc = scan_bins(L, L.counts, L.base, st);
  if (rc != PSC_OK) return rc;
  PSC_CUDA(cudaMemsetAsync(L.fill, 0, sizeof(int) * (L.nbins + 1), st));
  if (np > 0) {
    bin_scatter_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, x0, L.NBX, L.fill, L.base, L.rec, nullptr);
    count_launch();
  }
  PSC_CUDA(cudaMemsetAsync(L.heavy_count, 0, sizeof(int), st));
  bin_heavy_list_kernel<<<(int)((L.nbins + 255) / 256), 256, 0, st>>>(L.fill, (int)L.nbins, L.heavy_count, L.heavy,
                                                                     L.heavy_cap);
  count_launch();
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
  PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
  if (np > 0) {
    bin_count_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, x0, L.NBX, L.counts, nullptr);
    count_launch();
  }
  return finish_exact(pos, np, N, x0, L, st);
}
int psc_bin_particles(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, void *stream) {
  return psc_bin_particles_slab(pos, np, N, 0, N, scratch, scratch_bytes, stream);
}

/* mode 0: kick + drift + wrap + per-bin COUNT of the new positions (psc_bin_particles_counted(mode 0) then scans and
 * scatters).  mode 1: DIRECT scatter -- `scratch` still holds the fill of the previous step's binning of (about) the
 * same particles: the bins get that fill + 1/8 + 32 records of room and every particle drops its record straight into
 * its bin (no count pass, no second read of the positions); psc_bin_particles_counted(mode 1) finishes, redoing the
 * exact binning on the device when a bin ran out of room.  zero_counts = 0 for the later chunks of a chunked upload
 * (row0 = global row of pos[0]). */
int psc_kick_drift_wrap_count(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                              int dt_is_f64, int N, int64_t np_total, void *scratch, size_t scratch_bytes,
                              int zero_counts, int mode, int64_t row0, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np <= np_total && np_total < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(row0 >= 0 && row0 + np <= np_total, "row0 out of range");
  PSC_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (count) or 1 (direct scatter)");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && ((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np_total, N, 0, N, L)) {
    set_error("psc_kick_drift_wrap_count: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (zero_counts) {
    if (mode == 1) {
      // room for every bin from the previous step's fill, then fresh cursors
      bin_caps_kernel<<<(int)((L.nbins + 256) / 256), 256, 0, st>>>(L.fill, (int)L.nbins, np_total, L.tmp);
      count_launch();
      int rc = scan_bins(L, L.tmp, L.base, st);
      if (rc != PSC_OK) return rc;
      PSC_CUDA(cudaMemsetAsync(L.fill, 0, sizeof(int) * (L.nbins + 1), st));
      PSC_CUDA(cudaMemsetAsync(L.overflow, 0, sizeof(int), st));
    } else {
      PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
    }
  }
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos && vel && acc, "null pointer");
  PSC_CHECK_ARG((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)acc) & 15) == 0, "pointers must be 16-byte aligned");
  const int g = grid_for((np + 3) / 4, 256, 8);
#define PSC_KDW(F64, DIRECT, CNT)                                                                                 \
  kick_drift_wrap_count_kernel<F64, DIRECT><<<g, 256, 0, st>>>(pos, vel, acc, np, half_dt, dt, N, L.NB, CNT, L.base, \
                                                               L.nrec, L.rec, (int)row0, L.overflow)
  if (mode == 1) {
    if (dt_is_f64) PSC_KDW(true, true, L.fill);
    else PSC_KDW(false, true, L.fill);
  } else {
    if (dt_is_f64) PSC_KDW(true, false, L.counts);
    else PSC_KDW(false, false, L.counts);
  }
#undef PSC_KDW
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_bin_particles_counted(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, int mode,
                              void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (count) or 1 (direct scatter)");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && (pos || np == 0), "null pointer");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, 0, N, L)) {
    set_error("psc_bin_particles_counted: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (mode == 0) return finish_exact(pos, np, N, 0, L, st);
  // direct scatter: the records are in place unless a bin overflowed; in that case (flag on the device, no host
  // round trip) the same stream redoes the exact binning: count -> scan -> adopt offsets -> scatter.  Every kernel
  // of the fall-back returns at once when the flag is clear; the 1 MB scan runs either way.
  PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
  if (np > 0) {
    bin_count_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, 0, L.NBX, L.counts, L.overflow);
    count_launch();
  }
  int rc = scan_bins(L, L.counts, L.tmp, st);
  if (rc != PSC_OK) return rc;
  bin_adopt_kernel<<<(int)((L.nbins + 256) / 256), 256, 0, st>>>(L.overflow, L.tmp, (int)L.nbins, L.base, L.fill);
  count_launch();
  if (np > 0) {
    bin_scatter_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, 0, L.NBX, L.fill, L.base, L.rec,
                                                             L.overflow);
    count_launch();
  }
  PSC_CUDA(cudaMemsetAsync(L.heavy_count, 0, sizeof(int), st));
  bin_heavy_list_kernel<<<(int)((L.nbins + 255) / 256), 256, 0, st>>>(L.fill, (int)L.nbins, L.heavy_count, L.heavy,
                                                                     L.heavy_cap);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

// ghost = 0: periodic N^3 grid (x0 = 0, nxl = N); ghost = 1: rho has nxl + 2 planes, plane 0 / nxl + 1 collect the
// mass that belongs to the neighbouring slabs
// table 0 / 1 of the sorted layout -> (base, fill) views of the layout
static void use_table(BinLayout &L, int table) {
  if (table == 1) {
    int *b = L.base2, *f = L.fill2;
    L.base2 = L.base; L.fill2 = L.fill;
    L.base = b; L.fill = f;
  }
}

static int deposit_binned_impl(const void *scratch, size_t scratch_bytes, int64_t np, int N, int x0, int nxl,
                               int ghost, int scheme, float scale, float f1, float f2, float *rho, void *stream,
                               const float *sorted_pos = nullptr, int table = 0) {
  BinLayout L;
  if (!bin_layout(const_cast<void *>(scratch), scratch_bytes, np, N, x0, nxl, L, sorted_pos == nullptr)) {
    set_error("psc_deposit_binned: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  use_table(L, table);
  cudaStream_t st = as_stream(stream);
  const int nxa = nxl + 2 * ghost;
  const int64_t n3 = (int64_t)nxa * N * N;
  // rho starts as the constant of the affine map (0 for a raw deposit); the kernels add f1 * (scale * sums)
  if (f2 == 0.0f) {
    PSC_CUDA(cudaMemsetAsync(rho, 0, sizeof(float) * n3, st));
  } else {
    fill_kernel<<<grid_for((n3 + 3) / 4, 256), 256, 0, st>>>(rho, n3, f2);
    count_launch();
  }
  if (np > 0) {
    const int grid = (int)L.nbins;
    const int hgrid = num_sms() * 4;   // persistent CTAs over the heavy-bin parts (exit at once when there are none)
#define PSC_DEP(S)                                                                                                   \
  do {                                                                                                               \
    if (sorted_pos) {                                                                                                \
      deposit_binned_kernel<S, true><<<grid, BD_WARPS * 32, 0, st>>>(sorted_pos, L.base, L.fill, N, L.NB, x0, ghost,  \
                                                                     nxa, scale, f1, f2, rho);                       \
      deposit_heavy_kernel<S, true><<<hgrid, BD_WARPS * 32, 0, st>>>(sorted_pos, L.base, L.fill, L.heavy_count,       \
                                                                    L.heavy, L.heavy_cap, N, L.NB, x0, ghost, nxa,   \
                                                                    scale, f1, f2, rho);                             \
    } else {                                                                                                         \
      deposit_binned_kernel<S, false><<<grid, BD_WARPS * 32, 0, st>>>(L.rec, L.base, L.fill, N, L.NB, x0, ghost, nxa, \
                                                                      scale, f1, f2, rho);                           \
      deposit_heavy_kernel<S, false><<<hgrid, BD_WARPS * 32, 0, st>>>(L.rec, L.base, L.fill, L.heavy_count, L.heavy,  \
                                                                     L.heavy_cap, N, L.NB, x0, ghost, nxa, scale,    \
                                                                     f1, f2, rho);                                   \
    }                                                                                                                \
  } while (0)
    if (scheme == PSC_TSC) { PSC_DEP(PSC_TSC); }
    else if (scheme == PSC_CIC) { PSC_DEP(PSC_CIC); }
    else { PSC_DEP(PSC_NGP); }
#undef PSC_DEP
    count_launch(2);
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
    interp_kick4_binned_kernel<PSC_TSC><<<grid, BI_THREADS, 0, st>>>(f4, L.rec, L.base, L.fill, vel, acc, N, L.NB, 0, 0, N, half_dt, maxout, (int)L.nbins, L.heavy_count, L.heavy);
  else
    interp_kick4_binned_kernel<PSC_CIC><<<grid, BI_THREADS, 0, st>>>(f4, L.rec, L.base, L.fill, vel, acc, N, L.NB, 0, 0, N, half_dt, maxout, (int)L.nbins, L.heavy_count, L.heavy);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

// ghost = 0: periodic N^3 grids; ghost = G >= 1 + reach(order): phi (and u) have nxl + 2G planes, the owned ones
// start at plane G
static int interp_kick_phi_impl(const float *phi, const float *u, float f, int fr_n, int order, int x0, int nxl,
                                int ghost, const void *scratch, size_t scratch_bytes, float *vel, float *acc,
                                int64_t np, int N, int scheme, float half_dt, float *maxout, void *stream,
                                const float *sorted_pos = nullptr, int table = 0, const BinPredict *predict = nullptr) {
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG((((uintptr_t)phi | (uintptr_t)u) & 15) == 0, "phi and u must be 16-byte aligned");
  BinLayout L;
  if (!bin_layout(const_cast<void *>(scratch), scratch_bytes, np, N, x0, nxl, L, sorted_pos == nullptr)) {
    set_error("psc_interp_kick_phi_binned: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  use_table(L, table);
  cudaStream_t st = as_stream(stream);
  const int grid = (int)L.nbins + L.heavy_cap;   // the CTAs of unused heavy-part slots exit at once
  const int nxa = nxl + 2 * ghost;
  BinPredict pr = {nullptr, 0.0f, 0.0f, 0.0, 0};
  if (predict && sorted_pos && vel) {
    pr = *predict;
    pr.counts = L.counts;
    PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
  }
#define PSC_IKP(S, O)                                                                                               \
  do {                                                                                                              \
    if (sorted_pos)                                                                                                 \
      interp_kick_phi_binned_kernel<S, O, true><<<grid, BP_THREADS, 0, st>>>(phi, u, f, fr_n, sorted_pos, L.base,   \
                                                                             L.fill, vel, acc, N, L.NB, x0, ghost,  \
                                                                             nxa, half_dt, maxout, (int)L.nbins,    \
                                                                             L.heavy_count, L.heavy, pr);           \
    else                                                                                                            \
      interp_kick_phi_binned_kernel<S, O, false><<<grid, BP_THREADS, 0, st>>>(phi, u, f, fr_n, L.rec, L.base,       \
                                                                              L.fill, vel, acc, N, L.NB, x0, ghost, \
                                                                              nxa, half_dt, maxout, (int)L.nbins,   \
                                                                              L.heavy_count, L.heavy, pr);          \
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

/* ----------------------------------------------------------- particle arrays in bin order (the time loop) */
static size_t sorted_bytes(int64_t np, int64_t nbins) {
  return 6 * a256(sizeof(int) * (nbins + 1)) + 256 + a256(sizeof(int2) * (size_t)(np / BIN_PART + 1)) +
         a256(scan_tmp_bytes(nbins + 1)) + 256;
}
size_t psc_sorted_workspace_bytes(int64_t np, int N) {
  if (np < 0 || !slab_ok(N, 0, N)) return 0;
  return sorted_bytes(np, (int64_t)(N / BB) * (N / BB) * (N / BB));
}

/* src_table: -1 when the input arrays are in no particular order (first step, after utils.reorder_particles): both
 * passes go through one global atomic per particle and the result is table 0.  0 / 1 when the input arrays are the
 * bin-ordered output of a previous call described by that table: one CTA per source bin sorts its particles in shared
 * memory (step_sort_local_kernel) and the result is table 1 - src_table. */
int psc_step_sort(const float *pos, const float *vel, const float *acc, const int *ids, int64_t np, float half_dt,
                  double dt, int dt_is_f64, int N, int src_table, int counts_ready, void *scratch,
                  size_t scratch_bytes, float *pos_out, float *vel_out, int *ids_out, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(src_table >= -1 && src_table <= 1, "src_table must be -1, 0 or 1");
  PSC_CHECK_ARG(scratch && ((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, 0, N, L, false)) {
    set_error("psc_step_sort: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  const int *base_src = src_table == 1 ? L.base2 : L.base, *fill_src = src_table == 1 ? L.fill2 : L.fill;
  use_table(L, src_table < 0 ? 0 : 1 - src_table);     // L.base / L.fill: the table being written
  cudaStream_t st = as_stream(stream);
  // counts_ready: the previous psc_interp_kick_phi_sorted(predict = 1) already counted the destination bins of exactly
  // this (half_dt, dt, dt_is_f64) on exactly these arrays
  if (!counts_ready) PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
  PSC_CUDA(cudaMemsetAsync(L.fill, 0, sizeof(int) * (L.nbins + 1), st));
  PSC_CUDA(cudaMemsetAsync(L.heavy_count, 0, sizeof(int), st));
  if (np > 0) {
    PSC_CHECK_ARG(pos && vel && acc && pos_out && vel_out && ids_out, "null pointer");
    PSC_CHECK_ARG(pos != pos_out && vel != vel_out && ids != ids_out, "the sort is out of place");
    PSC_CHECK_ARG((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)acc | (uintptr_t)ids) & 15) == 0,
                  "pointers must be 16-byte aligned");
  }
  const int g = grid_for((np + 3) / 4, 256, 8);
  if (np > 0 && !counts_ready) {
    // pass 1, either input order: destination-bin counts, one atomic per distinct bin of a warp
    if (dt_is_f64)
      step_sort_kernel<true, 1><<<g, 256, 0, st>>>(pos, vel, acc, ids, np, half_dt, dt, N, L.NB, L.counts, nullptr,
                                                   pos_out, vel_out, ids_out);
    else
      step_sort_kernel<false, 1><<<g, 256, 0, st>>>(pos, vel, acc, ids, np, half_dt, dt, N, L.NB, L.counts, nullptr,
                                                    pos_out, vel_out, ids_out);
    count_launch();
  }
  int rc = scan_bins(L, L.counts, L.base, st);
  if (rc != PSC_OK) return rc;
  if (np > 0 && src_table < 0) {
    if (dt_is_f64)
      step_sort_kernel<true, 2><<<g, 256, 0, st>>>(pos, vel, acc, ids, np, half_dt, dt, N, L.NB, L.fill, L.base,
                                                   pos_out, vel_out, ids_out);
    else
      step_sort_kernel<false, 2><<<g, 256, 0, st>>>(pos, vel, acc, ids, np, half_dt, dt, N, L.NB, L.fill, L.base,
                                                    pos_out, vel_out, ids_out);
    count_launch();
  } else if (np > 0) {
    static const bool cell_key = env_is("PSC_SORT_KEY", "cell");   // experiment: row-major cells instead of micro-blocks
    const int gl = (int)std::min<int64_t>(L.nbins, (int64_t)num_sms() * 4);
#define PSC_SORT_LOCAL(F64, NSELF)                                                                                    \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      PSC_CUDA(cudaFuncSetAttribute(step_sort_local_kernel<F64, NSELF, false>,                                        \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL_SMEM));                      \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    step_sort_local_kernel<F64, NSELF, false><<<gl, 256, SL_SMEM, st>>>(                                              \
        pos, vel, acc, ids, np, np, base_src, fill_src, half_dt, dt, N, L.NB, 0, L.NB, (int)L.nbins, L.fill, L.base,  \
        pos_out, vel_out, ids_out);                                                                                   \
  } while (0)
    if (cell_key) {
      if (dt_is_f64) PSC_SORT_LOCAL(true, 512);
      else PSC_SORT_LOCAL(false, 512);
    } else {
      if (dt_is_f64) PSC_SORT_LOCAL(true, MB_PER_BIN);
      else PSC_SORT_LOCAL(false, MB_PER_BIN);
    }
#undef PSC_SORT_LOCAL
    count_launch();
  }
  bin_heavy_list_kernel<<<(int)((L.nbins + 255) / 256), 256, 0, st>>>(L.fill, (int)L.nbins, L.heavy_count, L.heavy,
                                                                     L.heavy_cap);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_deposit_sorted(const float *pos_sorted, const void *scratch, size_t scratch_bytes, int table, int64_t np, int N,
                       int scheme, float scale, float f1, float f2, float *rho, void *stream) {
  PSC_CHECK_ARG(table == 0 || table == 1, "table must be 0 or 1");
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  PSC_CHECK_ARG(slab_ok(N, 0, N), "N must be a multiple of 8");
  PSC_CHECK_ARG(scratch && rho && (pos_sorted || np == 0), "null pointer");
  static const float dummy = 0.0f;
  return deposit_binned_impl(scratch, scratch_bytes, np, N, 0, N, 0, scheme, scale, f1, f2, rho, stream,
                             pos_sorted ? pos_sorted : &dummy, table);
}

int psc_interp_kick_phi_sorted(const float *phi, const float *u, float f, int fr_n, int order, const float *pos_sorted,
                               const void *scratch, size_t scratch_bytes, int table, float *vel_sorted,
                               float *acc_sorted, int64_t np, int N, int scheme, float half_dt, float *maxout,
                               int predict, float next_half_dt, double next_dt, int next_dt_is_f64, void *stream) {
  PSC_CHECK_ARG(table == 0 || table == 1, "table must be 0 or 1");
  BinPredict pr = {nullptr, -next_half_dt, (float)next_dt, next_dt, next_dt_is_f64 != 0};
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  PSC_CHECK_ARG(order == 2 || order == 3 || order == 5 || order == 7, "gradient order must be 2, 3, 5 or 7");
  PSC_CHECK_ARG(fr_n >= 0 && fr_n <= 2, "fR_n must be 1 or 2");
  PSC_CHECK_ARG(N >= 2 * BB && (N % BB) == 0, "N must be a multiple of 8 and >= 16");
  PSC_CHECK_ARG(phi && scratch && acc_sorted && maxout && (u || fr_n == 0) && (pos_sorted || np == 0), "null pointer");
  return interp_kick_phi_impl(phi, u, f, fr_n, order, 0, N, 0, scratch, scratch_bytes, vel_sorted, acc_sorted, np, N,
                              scheme, half_dt, maxout, stream, pos_sorted, table, predict ? &pr : nullptr);
}

/* ids_out[n] = rank of particle n (row n of the bin-ordered pos_sorted) in Morton-key order: after the call
 * utils.reference_order returns the arrays the reference has after utils.reorder_particles (utils.py:1019-1075).
 * N must be a power of two.  *too_big (device int, zeroed here) becomes 1 if a bin holds more than 2048 particles:
 * the caller then falls back to the global sort.  Uses the scratch's count / tmp tables. */
int psc_morton_ids_sorted(const float *pos_sorted, void *scratch, size_t scratch_bytes, int table, int64_t np, int N,
                          int *ids_out, int *too_big, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, 0, N) && (N & (N - 1)) == 0 && N <= 8192, "N must be a power of two in [8, 8192]");
  PSC_CHECK_ARG(table == 0 || table == 1, "table must be 0 or 1");
  PSC_CHECK_ARG(scratch && too_big, "null pointer");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, 0, N, L, false)) {
    set_error("psc_morton_ids_sorted: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  use_table(L, table);
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(too_big, 0, sizeof(int), st));
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(pos_sorted && ids_out, "null pointer");
  const int nbins = (int)L.nbins;
  zorder_fill_kernel<<<nbins / 256 + 1, 256, 0, st>>>(L.fill, nbins, L.NB, L.counts);
  count_launch();
  int rc = scan_bins(L, L.counts, L.tmp, st);
  if (rc != PSC_OK) return rc;
  int nb_bits = 0;
  while ((1 << nb_bits) < L.NB) nb_bits++;
  const int key_bits = 3 * (21 - nb_bits);
  morton_rank_kernel<3><<<nbins, 256, 0, st>>>(pos_sorted, L.base, L.fill, L.tmp, L.NB, 0, key_bits, ids_out, too_big);
  // bins of 769 .. 2048 particles; the launch above only flagged them (and wrote nothing for them)
  PSC_CUDA(cudaMemsetAsync(too_big, 0, sizeof(int), st));
  morton_rank_kernel<8><<<nbins, 256, 0, st>>>(pos_sorted, L.base, L.fill, L.tmp, L.NB, 768, key_bits, ids_out,
                                               too_big);
  count_launch(2);
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_scatter3_by_id(const int *ids, const float *in, float *out, int64_t np, void *stream) {
  PSC_CHECK_ARG(np >= 0, "np < 0");
  if (np == 0) return PSC_OK;
  PSC_CHECK_ARG(ids && in && out && in != out, "null or aliased pointer");
  scatter3_by_id_kernel<<<grid_for(np, 256, 8), 256, 0, as_stream(stream)>>>(ids, in, out, np);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

/* ---- the same on a slab: the particle arrays of the rank are sorted into bin order after the migration */
size_t psc_sorted_workspace_bytes_slab(int64_t np, int N, int nxl) {
  if (np < 0 || !slab_ok(N, 0, nxl)) return 0;
  return sorted_bytes(np, (int64_t)(nxl / BB) * (N / BB) * (N / BB));
}

/* src_table = -1: the input arrays are in no particular order (after set_particles / the Morton reorder): one global
 * atomic per particle, result in table 0.  src_table = 0 / 1: rows [0, src_rows) are the bin-ordered arrays that table
 * describes, as the in-place kick + drift and the migration left them (arrivals in the holes of the leavers, tail rows
 * moved down or new rows [src_rows, np) appended): one CTA per source bin sorts in shared memory
 * (step_sort_local_kernel<.., SLAB>), the appended rows go through the per-particle path; result in table
 * 1 - src_table. */
int psc_sort_by_bin_slab(const float *pos, const float *vel, const int64_t *ids, int64_t np, int N, int x0, int nxl,
                         int src_table, int64_t src_rows, void *scratch, size_t scratch_bytes, float *pos_out,
                         float *vel_out, int64_t *ids_out, void *stream) {
  PSC_CHECK_ARG(np >= 0 && np < ((int64_t)1 << 31), "np out of range");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl), "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(src_table >= -1 && src_table <= 1 && src_rows >= 0, "src_table must be -1, 0 or 1");
  PSC_CHECK_ARG(scratch && ((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  BinLayout L;
  if (!bin_layout(scratch, scratch_bytes, np, N, x0, nxl, L, false)) {
    set_error("psc_sort_by_bin_slab: scratch too small");
    return PSC_ERR_WORKSPACE;
  }
  const int *base_src = src_table == 1 ? L.base2 : L.base, *fill_src = src_table == 1 ? L.fill2 : L.fill;
  use_table(L, src_table < 0 ? 0 : 1 - src_table);
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(L.counts, 0, sizeof(int) * (L.nbins + 1), st));
  PSC_CUDA(cudaMemsetAsync(L.fill, 0, sizeof(int) * (L.nbins + 1), st));
  PSC_CUDA(cudaMemsetAsync(L.heavy_count, 0, sizeof(int), st));
  const bool local = src_table >= 0 && np > 0 && ((((uintptr_t)pos | (uintptr_t)vel | (uintptr_t)ids) & 15) == 0);
  if (np > 0) {
    PSC_CHECK_ARG(pos && vel && ids && pos_out && vel_out && ids_out, "null pointer");
    PSC_CHECK_ARG(pos != pos_out && vel != vel_out && ids != ids_out, "the sort is out of place");
    bin_count_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, L.NB, x0, L.NBX, L.counts, nullptr);
    count_launch();
  }
  int rc = scan_bins(L, L.counts, L.base, st);
  if (rc != PSC_OK) return rc;
  // rows the per-particle scatter takes: all of them, or the ones appended behind the rows the source table describes
  const int64_t tail0 = local ? std::min(src_rows, np) : 0;
  if (local) {
    static bool attr_set = false;
    if (!attr_set) {
      PSC_CUDA(cudaFuncSetAttribute(step_sort_local_kernel<false, MB_PER_BIN, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL_SMEM));
      attr_set = true;
    }
    const int gl = (int)std::min<int64_t>(L.nbins, (int64_t)num_sms() * 4);
    step_sort_local_kernel<false, MB_PER_BIN, true><<<gl, 256, SL_SMEM, st>>>(
        pos, vel, nullptr, reinterpret_cast<const long long *>(ids), np, tail0, base_src, fill_src, 0.0f, 0.0, N, L.NB, x0,
        L.NBX, (int)L.nbins, L.fill, L.base, pos_out, vel_out, reinterpret_cast<long long *>(ids_out));
    count_launch();
  }
  if (np > tail0) {
    sort_scatter_kernel<<<grid_for(np - tail0, 256, 8), 256, 0, st>>>(pos + 3 * tail0, vel + 3 * tail0, ids + tail0,
                                                                     np - tail0, N, L.NB, x0, L.NBX, L.fill, L.base,
                                                                     pos_out, vel_out, ids_out);
    count_launch();
  }
  bin_heavy_list_kernel<<<(int)((L.nbins + 255) / 256), 256, 0, st>>>(L.fill, (int)L.nbins, L.heavy_count, L.heavy,
                                                                     L.heavy_cap);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_deposit_sorted_slab(const float *pos_sorted, const void *scratch, size_t scratch_bytes, int table, int64_t np,
                            int N, int x0, int nxl, int scheme, float *rho_ghost, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_NGP || scheme == PSC_CIC || scheme == PSC_TSC, "unknown mass scheme");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl), "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(scratch && rho_ghost && (pos_sorted || np == 0), "null pointer");
  PSC_CHECK_ARG(table == 0 || table == 1, "table must be 0 or 1");
  static const float dummy = 0.0f;
  return deposit_binned_impl(scratch, scratch_bytes, np, N, x0, nxl, 1, scheme, 1.0f, 1.0f, 0.0f, rho_ghost, stream,
                             pos_sorted ? pos_sorted : &dummy, table);
}

int psc_interp_kick_phi_sorted_slab(const float *phi_ghost, const float *u_ghost, float f, int fr_n, int order, int x0,
                                    int nxl, int ghost, const float *pos_sorted, const void *scratch,
                                    size_t scratch_bytes, int table, float *vel_sorted, float *acc_sorted, int64_t np,
                                    int N, int scheme, float half_dt, float *maxout, void *stream) {
  PSC_CHECK_ARG(scheme == PSC_CIC || scheme == PSC_TSC, "mass scheme must be CIC or TSC");
  PSC_CHECK_ARG(order == 2 || order == 3 || order == 5 || order == 7, "gradient order must be 2, 3, 5 or 7");
  PSC_CHECK_ARG(fr_n >= 0 && fr_n <= 2, "fR_n must be 1 or 2");
  PSC_CHECK_ARG(slab_ok(N, x0, nxl) && N >= 2 * BB, "N and the slab thickness must be multiples of 8");
  PSC_CHECK_ARG(ghost >= 1 + (order == 7 ? 3 : order == 5 ? 2 : 1), "not enough ghost planes for this stencil");
  PSC_CHECK_ARG(phi_ghost && scratch && acc_sorted && maxout && (u_ghost || fr_n == 0) && (pos_sorted || np == 0),
                "null pointer");
  PSC_CHECK_ARG(table == 0 || table == 1, "table must be 0 or 1");
  return interp_kick_phi_impl(phi_ghost, u_ghost, f, fr_n, order, x0, nxl, ghost, scratch, scratch_bytes, vel_sorted,
                              acc_sorted, np, N, scheme, half_dt, maxout, stream, pos_sorted, table);
}

}  // extern "C"
