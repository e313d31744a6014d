// slab.cu -- kernels of the x-slab decomposed particle-mesh step (one slab of N/P planes per GPU).
//
// The reference has no distributed path (README.md:49); these kernels implement the decomposition SURVEY 8(e)
// lays out for its hot path:
//   * particle migration after the drift (integration.py:252-258 moves every particle by < 1 cell per step):
//     psc_slab_count / psc_slab_pack_leavers / psc_slab_unpack_rows / psc_slab_move_rows.  Particles that left
//     the slab are packed into 32-byte records (x, v, 64-bit id) ordered by destination rank, the holes they
//     leave are filled by the arrivals (and, if more left than arrived, by particles from the tail), so a
//     migration touches O(migrants) rows, not O(Np).
//   * the transposed FFT of fourier.fft_3D_real / ifft_3D_real (fourier.py:104-147, 251-294): batched 2-D
//     R2C over the owned planes -> pack by y-block -> all-to-all (host, NCCL) -> strided 1-D C2C along x; the
//     Green's function (psc_green_slab, fourier.cu) is applied in the transposed [x][y_local][kz] layout;
//     the inverse mirrors it.
// Ghost-plane exchanges are plain contiguous plane copies and are done by the host (pysco_b200/slab.py).
#include <cufft.h>

#include "common.cuh"

namespace psc {

constexpr int REC = 8;  // floats per migration record: x y z vx vy vz id_lo id_hi

__device__ __forceinline__ int owner_of(float x, float Nf, int nxl, int P) {
  const int i = (int)(x * Nf);
  return min(max(i / nxl, 0), P - 1);
}

// counts[d] = number of particles owned by rank d != me (stayers are the vast majority and are not counted with
// atomics: the host gets counts[me] as np - sum of the others); one aggregated atomic per distinct owner per warp
__global__ void __launch_bounds__(256) slab_count_kernel(const float *__restrict__ pos, int64_t np, int N, int nxl,
                                                         int P, int me, unsigned long long *__restrict__ counts) {
  const float Nf = (float)N;
  const int lane = threadIdx.x & 31;
  const int64_t nwarp_iters = (np + 31) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwarp_iters; w += wstride) {
    const int64_t n = w * 32 + lane;
    int d = me;
    if (n < np) d = owner_of(__ldg(&pos[3 * n]), Nf, nxl, P);
    if (__any_sync(0xffffffffu, d != me)) {
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (d != me && (__ffs(peers) - 1) == lane) atomicAdd(&counts[d], (unsigned long long)__popc(peers));
    }
  }
}

// every particle whose owner is not `me` is written to sendbuf[offsets[owner] + slot] and its row recorded
__global__ void __launch_bounds__(256) slab_pack_kernel(const float *__restrict__ pos, const float *__restrict__ vel,
                                                        const int64_t *__restrict__ ids, int64_t np, int N, int nxl,
                                                        int P, int me, const int64_t *__restrict__ offsets,
                                                        unsigned long long *__restrict__ cursor,
                                                        float *__restrict__ sendbuf, int64_t *__restrict__ holes) {
  const float Nf = (float)N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < np; n += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(&pos[3 * n]);
    const int d = owner_of(x, Nf, nxl, P);
    if (d == me) continue;
    const int64_t slot = offsets[d] + (int64_t)atomicAdd(&cursor[d], 1ull);
    float4 *rec = reinterpret_cast<float4 *>(sendbuf + REC * slot);
    const int64_t id = ids[n];
    rec[0] = make_float4(x, pos[3 * n + 1], pos[3 * n + 2], vel[3 * n]);
    rec[1] = make_float4(vel[3 * n + 1], vel[3 * n + 2], __int_as_float((int)(id & 0xffffffffll)),
                         __int_as_float((int)(id >> 32)));
    holes[slot] = n;
  }
}

// same, over a list of candidate rows (written by psc_kick_drift_wrap_slab) instead of all particles
__global__ void __launch_bounds__(256) slab_pack_rows_kernel(const float *__restrict__ pos, const float *__restrict__ vel,
                                                             const int64_t *__restrict__ ids,
                                                             const int64_t *__restrict__ rows, int64_t nrows, int N,
                                                             int nxl, int P, int me,
                                                             const int64_t *__restrict__ offsets,
                                                             unsigned long long *__restrict__ cursor,
                                                             float *__restrict__ sendbuf, int64_t *__restrict__ holes) {
  const float Nf = (float)N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = rows[t];
    const float x = pos[3 * n];
    const int d = owner_of(x, Nf, nxl, P);
    if (d == me) continue;
    const int64_t slot = offsets[d] + (int64_t)atomicAdd(&cursor[d], 1ull);
    float4 *rec = reinterpret_cast<float4 *>(sendbuf + REC * slot);
    const int64_t id = ids[n];
    rec[0] = make_float4(x, pos[3 * n + 1], pos[3 * n + 2], vel[3 * n]);
    rec[1] = make_float4(vel[3 * n + 1], vel[3 * n + 2], __int_as_float((int)(id & 0xffffffffll)),
                         __int_as_float((int)(id >> 32)));
    holes[slot] = n;
  }
}

// Fixed-capacity packing for the single-round neighbour migration: leavers towards the left / right slab go to
// sendL / sendR (capacity `cap` records after a one-record header), their rows to holesL / holesR.  Everything is
// decided on the device -- the candidate list written by psc_kick_drift_wrap_slab is used if it is complete
// (counts[P] <= list_capacity), otherwise all np particles are scanned -- so the host needs no count before the
// exchange.  status[0] / [1] = leavers towards left / right (may exceed cap: those beyond cap are NOT packed),
// status[2] = leavers whose owner is not a neighbour (an error: Courant condition violated).
__global__ void __launch_bounds__(256) slab_pack_fixed_kernel(const float *__restrict__ pos, const float *__restrict__ vel,
                                                              const int64_t *__restrict__ ids, int64_t np,
                                                              const int64_t *__restrict__ rows,
                                                              const int64_t *__restrict__ counts, int64_t list_capacity,
                                                              int N, int nxl, int P, int me, int left, int right,
                                                              int64_t cap, float *__restrict__ sendL,
                                                              float *__restrict__ sendR, int64_t *__restrict__ holesL,
                                                              int64_t *__restrict__ holesR,
                                                              unsigned long long *__restrict__ status) {
  const float Nf = (float)N;
  const int64_t nlist = counts[P];
  const bool use_list = rows != nullptr && nlist <= list_capacity;
  const int64_t total = use_list ? nlist : np;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = use_list ? rows[t] : t;
    const float x = pos[3 * n];
    const int d = owner_of(x, Nf, nxl, P);
    if (d == me) continue;
    const bool to_left = d == left;
    if (!to_left && d != right) { atomicAdd(&status[2], 1ull); continue; }
    const int64_t slot = (int64_t)atomicAdd(&status[to_left ? 0 : 1], 1ull);
    if (slot >= cap) continue;
    float4 *rec = reinterpret_cast<float4 *>((to_left ? sendL : sendR) + REC * (slot + 1));
    const int64_t id = ids[n];
    rec[0] = make_float4(x, pos[3 * n + 1], pos[3 * n + 2], vel[3 * n]);
    rec[1] = make_float4(vel[3 * n + 1], vel[3 * n + 2], __int_as_float((int)(id & 0xffffffffll)),
                         __int_as_float((int)(id >> 32)));
    (to_left ? holesL : holesR)[slot] = n;
  }
}

// header record of both send buffers = the true leaver counts (int64 in the first two floats)
__global__ void slab_pack_header_kernel(const unsigned long long *__restrict__ status, float *__restrict__ sendL,
                                        float *__restrict__ sendR) {
  if (threadIdx.x == 0) {
    reinterpret_cast<unsigned long long *>(sendL)[0] = status[0];
    reinterpret_cast<unsigned long long *>(sendR)[0] = status[1];
  }
}

__global__ void __launch_bounds__(256) slab_unpack_kernel(const float *__restrict__ recvbuf,
                                                          const int64_t *__restrict__ rows, int64_t n,
                                                          float *__restrict__ pos, float *__restrict__ vel,
                                                          int64_t *__restrict__ ids) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float4 *rec = reinterpret_cast<const float4 *>(recvbuf + REC * t);
    const float4 a = rec[0], b = rec[1];
    const int64_t r = rows[t];
    pos[3 * r] = a.x; pos[3 * r + 1] = a.y; pos[3 * r + 2] = a.z;
    vel[3 * r] = a.w; vel[3 * r + 1] = b.x; vel[3 * r + 2] = b.y;
    ids[r] = (int64_t)(unsigned int)__float_as_int(b.z) | ((int64_t)__float_as_int(b.w) << 32);
  }
}

__global__ void __launch_bounds__(256) slab_move_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                        int64_t n, float *__restrict__ pos, float *__restrict__ vel,
                                                        int64_t *__restrict__ ids) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = src[t], d = dst[t];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      pos[3 * d + c] = pos[3 * s + c];
      vel[3 * d + c] = vel[3 * s + c];
    }
    ids[d] = ids[s];
  }
}

// [nxl][N][nz] -> [P][nxl][nyl][nz]  (to_blocks) and back; rows of nz complex numbers stay contiguous
template <bool TO_BLOCKS>
__global__ void __launch_bounds__(256) yblock_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int N,
                                                     int nxl, int nyl, int nz) {
  const int64_t nrows = (int64_t)nxl * N;
  const int rows_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_cta + wid; r < nrows; r += (int64_t)gridDim.x * rows_per_cta) {
    const int y = (int)(r % N), x = (int)(r / N);
    const int d = y / nyl, yl = y - d * nyl;
    const int64_t a = r * nz;                                         // [x][y][.]
    const int64_t b = (((int64_t)d * nxl + x) * nyl + yl) * nz;       // [d][x][yl][.]
    const float2 *s = in + (TO_BLOCKS ? a : b);
    float2 *t = out + (TO_BLOCKS ? b : a);
    for (int k = lane; k < nz; k += 32) t[k] = s[k];
  }
}

// Transposes of the slab FFT as ONE kernel over NVLink peer memory: every rank stores the rows of its spectrum
// straight into the receive buffers of their destination ranks (peers[d] = rank d's symmetric buffer, mapped into
// this process), already in the layout the next transform wants.  No pack pass, no staging buffer, no NCCL
// all-to-all; the self block is an ordinary local copy running at HBM speed alongside the remote stores.
//   FWD: in [nxl][N][nz] (2-D spectra of my planes) -> peer d = y / nyl gets row (me nxl + x, y - d nyl) of its
//        [N][nyl][nz] buffer;
//   !FWD: in [N][nyl][nz] -> peer d = x / nxl gets row (x - d nxl, me nyl + yl) of its [nxl][N][nz] buffer.
template <bool FWD>
__global__ void __launch_bounds__(256) slab_put_kernel(const float2 *__restrict__ in, float2 *const *__restrict__ peers,
                                                       int N, int nxl, int nyl, int nz, int me) {
  const int64_t nrows = FWD ? (int64_t)nxl * N : (int64_t)N * nyl;
  const int rows_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_cta + wid; r < nrows; r += (int64_t)gridDim.x * rows_per_cta) {
    int d;
    int64_t drow;
    if (FWD) {
      const int y = (int)(r % N), x = (int)(r / N);
      d = y / nyl;
      drow = ((int64_t)me * nxl + x) * nyl + (y - d * nyl);
    } else {
      const int yl = (int)(r % nyl), x = (int)(r / nyl);
      d = x / nxl;
      drow = (int64_t)(x - d * nxl) * N + (int64_t)me * nyl + yl;
    }
    const float2 *s = in + r * nz;
    float2 *t = peers[d] + drow * nz;
    for (int k = lane; k < nz; k += 32) t[k] = s[k];
  }
}

struct SlabFftPlan {
  int N, nxl, nyl;
  cufftHandle r2c, c2r, c2c;
  size_t work_bytes;
};

static const char *cufft_str2(cufftResult r) {
  switch (r) {
    case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
    case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
    case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
    case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
    case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
    case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
    case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
    case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
    default: return "CUFFT_ERROR";
  }
}
#define PSC_CUFFT2(call)                                                     \
  do {                                                                       \
    cufftResult r__ = (call);                                                \
    if (r__ != CUFFT_SUCCESS) {                                              \
      psc::set_error("%s: %s failed: %s", __func__, #call, cufft_str2(r__)); \
      return PSC_ERR_CUFFT;                                                  \
    }                                                                        \
  } while (0)

}  // namespace psc

using namespace psc;

extern "C" {

int psc_slab_count(const float *pos, int64_t np, int N, int nxl, int P, int me, int64_t *counts, void *stream) {
  PSC_CHECK_ARG(np >= 0 && N >= 1 && nxl >= 1 && P >= 1 && nxl * P == N && me >= 0 && me < P, "bad slab geometry");
  PSC_CHECK_ARG(counts && (pos || np == 0), "null pointer");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * P, st));
  if (np > 0) {
    slab_count_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, np, N, nxl, P, me,
                                                            reinterpret_cast<unsigned long long *>(counts));
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return PSC_OK;
}

int psc_slab_pack_leavers(const float *pos, const float *vel, const int64_t *ids, int64_t np, int N, int nxl, int P,
                          int me, const int64_t *offsets, int64_t *cursor, float *sendbuf, int64_t *holes,
                          void *stream) {
  PSC_CHECK_ARG(np >= 0 && N >= 1 && nxl >= 1 && P >= 1 && nxl * P == N && me >= 0 && me < P, "bad slab geometry");
  PSC_CHECK_ARG(offsets && cursor && (np == 0 || (pos && vel && ids)), "null pointer");
  PSC_CHECK_ARG(((uintptr_t)sendbuf & 15) == 0, "sendbuf must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int64_t) * P, st));
  if (np > 0) {
    slab_pack_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(pos, vel, ids, np, N, nxl, P, me, offsets,
                                                           reinterpret_cast<unsigned long long *>(cursor), sendbuf,
                                                           holes);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return PSC_OK;
}

int psc_slab_pack_rows(const float *pos, const float *vel, const int64_t *ids, const int64_t *rows, int64_t nrows,
                       int N, int nxl, int P, int me, const int64_t *offsets, int64_t *cursor, float *sendbuf,
                       int64_t *holes, void *stream) {
  PSC_CHECK_ARG(nrows >= 0 && N >= 1 && nxl >= 1 && P >= 1 && nxl * P == N && me >= 0 && me < P, "bad slab geometry");
  PSC_CHECK_ARG(offsets && cursor && (nrows == 0 || (pos && vel && ids && rows)), "null pointer");
  PSC_CHECK_ARG(((uintptr_t)sendbuf & 15) == 0, "sendbuf must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int64_t) * P, st));
  if (nrows > 0) {
    slab_pack_rows_kernel<<<grid_for(nrows, 256, 8), 256, 0, st>>>(pos, vel, ids, rows, nrows, N, nxl, P, me, offsets,
                                                                   reinterpret_cast<unsigned long long *>(cursor),
                                                                   sendbuf, holes);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  return PSC_OK;
}

int psc_slab_pack_fixed(const float *pos, const float *vel, const int64_t *ids, int64_t np, const int64_t *rows,
                        const int64_t *counts, int64_t list_capacity, int N, int nxl, int P, int me, int64_t cap,
                        float *sendL, float *sendR, int64_t *holesL, int64_t *holesR, int64_t *status, void *stream) {
  PSC_CHECK_ARG(np >= 0 && N >= 1 && nxl >= 1 && P >= 2 && nxl * P == N && me >= 0 && me < P && cap >= 0,
                "bad slab geometry");
  PSC_CHECK_ARG(counts && sendL && sendR && holesL && holesR && status && (np == 0 || (pos && vel && ids)),
                "null pointer");
  PSC_CHECK_ARG((((uintptr_t)sendL | (uintptr_t)sendR) & 15) == 0, "send buffers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  PSC_CUDA(cudaMemsetAsync(status, 0, sizeof(int64_t) * 3, st));
  const int left = (me + P - 1) % P, right = (me + 1) % P;
  unsigned long long *stt = reinterpret_cast<unsigned long long *>(status);
  if (np > 0) {
    // grid-stride loop: the same grid serves the short candidate list and the (rare) full scan
    slab_pack_fixed_kernel<<<grid_for(np, 256, 8), 256, 0, st>>>(
        pos, vel, ids, np, rows, counts, list_capacity, N, nxl, P, me, left, right, cap, sendL, sendR, holesL, holesR,
        stt);
    count_launch();
    PSC_CHECK_LAUNCH();
  }
  slab_pack_header_kernel<<<1, 32, 0, st>>>(stt, sendL, sendR);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_slab_unpack_rows(const float *recvbuf, const int64_t *rows, int64_t n, float *pos, float *vel, int64_t *ids,
                         void *stream) {
  PSC_CHECK_ARG(n >= 0, "negative count");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(recvbuf && rows && pos && vel && ids, "null pointer");
  PSC_CHECK_ARG(((uintptr_t)recvbuf & 15) == 0, "recvbuf must be 16-byte aligned");
  slab_unpack_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(recvbuf, rows, n, pos, vel, ids);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_slab_move_rows(const int64_t *src, const int64_t *dst, int64_t n, float *pos, float *vel, int64_t *ids,
                       void *stream) {
  PSC_CHECK_ARG(n >= 0, "negative count");
  if (n == 0) return PSC_OK;
  PSC_CHECK_ARG(src && dst && pos && vel && ids, "null pointer");
  slab_move_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(src, dst, n, pos, vel, ids);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_slab_fft_plan_create(int N, int nxl, int nyl, void **plan_out) {
  PSC_CHECK_ARG(plan_out, "null plan_out");
  PSC_CHECK_ARG(N >= 2 && N <= 4096 && nxl >= 1 && nyl >= 1 && N % nxl == 0 && N % nyl == 0, "bad slab geometry");
  SlabFftPlan *pl = new SlabFftPlan();
  pl->N = N; pl->nxl = nxl; pl->nyl = nyl;
  const int nz = N / 2 + 1;
  size_t w[3] = {0, 0, 0};
  int n2[2] = {N, N};
  PSC_CUFFT2(cufftCreate(&pl->r2c));
  PSC_CUFFT2(cufftSetAutoAllocation(pl->r2c, 0));
  PSC_CUFFT2(cufftMakePlanMany(pl->r2c, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, nxl, &w[0]));
  PSC_CUFFT2(cufftCreate(&pl->c2r));
  PSC_CUFFT2(cufftSetAutoAllocation(pl->c2r, 0));
  PSC_CUFFT2(cufftMakePlanMany(pl->c2r, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, nxl, &w[1]));
  int n1[1] = {N};
  int emb[1] = {N};
  const int stride = nyl * nz;
  PSC_CUFFT2(cufftCreate(&pl->c2c));
  PSC_CUFFT2(cufftSetAutoAllocation(pl->c2c, 0));
  PSC_CUFFT2(cufftMakePlanMany(pl->c2c, 1, n1, emb, stride, 1, emb, stride, 1, CUFFT_C2C, stride, &w[2]));
  pl->work_bytes = w[0] > w[1] ? w[0] : w[1];
  if (w[2] > pl->work_bytes) pl->work_bytes = w[2];
  *plan_out = pl;
  return PSC_OK;
}

int psc_slab_fft_plan_destroy(void *plan) {
  if (!plan) return PSC_OK;
  SlabFftPlan *pl = reinterpret_cast<SlabFftPlan *>(plan);
  cufftDestroy(pl->r2c);
  cufftDestroy(pl->c2r);
  cufftDestroy(pl->c2c);
  delete pl;
  return PSC_OK;
}

size_t psc_slab_fft_workspace_bytes(void *plan) {
  return plan ? reinterpret_cast<SlabFftPlan *>(plan)->work_bytes : 0;
}

/* one work area (>= psc_slab_fft_workspace_bytes) shared by the three cuFFT plans */
int psc_slab_fft_set_workspace(void *plan, void *work) {
  PSC_CHECK_ARG(plan && work, "null pointer");
  SlabFftPlan *pl = reinterpret_cast<SlabFftPlan *>(plan);
  PSC_CUFFT2(cufftSetWorkArea(pl->r2c, work));
  PSC_CUFFT2(cufftSetWorkArea(pl->c2r, work));
  PSC_CUFFT2(cufftSetWorkArea(pl->c2c, work));
  return PSC_OK;
}

int psc_slab_fft_r2c_planes(void *plan, const float *planes, float *spec2d, void *stream) {
  PSC_CHECK_ARG(plan && planes && spec2d, "null pointer");
  SlabFftPlan *pl = reinterpret_cast<SlabFftPlan *>(plan);
  PSC_CUFFT2(cufftSetStream(pl->r2c, as_stream(stream)));
  PSC_CUFFT2(cufftExecR2C(pl->r2c, const_cast<float *>(planes), reinterpret_cast<cufftComplex *>(spec2d)));
  count_launch(2);
  return PSC_OK;
}

int psc_slab_fft_c2r_planes(void *plan, float *spec2d, float *planes, void *stream) {
  PSC_CHECK_ARG(plan && planes && spec2d, "null pointer");
  SlabFftPlan *pl = reinterpret_cast<SlabFftPlan *>(plan);
  PSC_CUFFT2(cufftSetStream(pl->c2r, as_stream(stream)));
  PSC_CUFFT2(cufftExecC2R(pl->c2r, reinterpret_cast<cufftComplex *>(spec2d), planes));
  count_launch(2);
  return PSC_OK;
}

int psc_slab_fft_x(void *plan, float *spec_t, int inverse, void *stream) {
  PSC_CHECK_ARG(plan && spec_t, "null pointer");
  SlabFftPlan *pl = reinterpret_cast<SlabFftPlan *>(plan);
  PSC_CUFFT2(cufftSetStream(pl->c2c, as_stream(stream)));
  cufftComplex *p = reinterpret_cast<cufftComplex *>(spec_t);
  PSC_CUFFT2(cufftExecC2C(pl->c2c, p, p, inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
  count_launch(1);
  return PSC_OK;
}

int psc_slab_transpose_put(const float *in, const void *peer_ptrs_dev, int N, int nxl, int nyl, int P, int me,
                           int forward, void *stream) {
  PSC_CHECK_ARG(in && peer_ptrs_dev, "null pointer");
  PSC_CHECK_ARG(N >= 2 && nxl >= 1 && nyl >= 1 && P >= 1 && nxl * P == N && nyl * P == N && me >= 0 && me < P,
                "bad slab geometry");
  const int nz = N / 2 + 1;
  const int64_t nrows = forward ? (int64_t)nxl * N : (int64_t)N * nyl;
  const int grid = grid_for(nrows * 32, 256, 8);
  const float2 *i2 = reinterpret_cast<const float2 *>(in);
  float2 *const *peers = reinterpret_cast<float2 *const *>(peer_ptrs_dev);
  if (forward) slab_put_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(i2, peers, N, nxl, nyl, nz, me);
  else slab_put_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(i2, peers, N, nxl, nyl, nz, me);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

int psc_slab_yblocks(const float *in, float *out, int N, int nxl, int nyl, int to_blocks, void *stream) {
  PSC_CHECK_ARG(in && out && in != out, "null or aliased pointer");
  PSC_CHECK_ARG(N >= 2 && nxl >= 1 && nyl >= 1 && N % nyl == 0, "bad slab geometry");
  const int nz = N / 2 + 1;
  const int64_t nrows = (int64_t)nxl * N;
  const int grid = grid_for(nrows * 32, 256, 8);
  const float2 *i2 = reinterpret_cast<const float2 *>(in);
  float2 *o2 = reinterpret_cast<float2 *>(out);
  if (to_blocks) yblock_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(i2, o2, N, nxl, nyl, nz);
  else yblock_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(i2, o2, N, nxl, nyl, nz);
  count_launch();
  PSC_CHECK_LAUNCH();
  return PSC_OK;
}

}  // extern "C"
