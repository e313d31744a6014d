/*
 * pysco_b200.h -- C ABI of libpysco_b200.so: the B200 (sm_100a) particle-mesh gravity step for PySCo.
 *
 * The reference (mianbreton/pysco 1.0.9) has no FFI layer: its hot path is a set of Numba-jitted
 * module-level Python functions.  Each entry point below replaces one (or a fused group) of those
 * functions; the reference file:line it stands in for is cited next to it.  INTEGRATION.md shows
 * the ctypes binding a PySCo maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; no ownership is transferred
 *     and nothing is allocated behind the caller's back except cuFFT plans (psc_fft_plan_*).
 *   - scalar grids [N,N,N] float32 C-order (k fastest); vector grids AoS [N,N,N,3]; particles AoS
 *     [Np,3]; rfft half-spectrum [N,N,N/2+1] complex64 interleaved (same as numpy.fft.rfftn).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - return value: 0 on success, negative psc_status otherwise; psc_last_error() gives the text.
 *   - no entry point synchronises the stream unless documented (results that are scalars are
 *     written to device memory so that callers choose when to read them).
 */
#ifndef PYSCO_B200_H
#define PYSCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PSC_OK = 0,
  PSC_ERR_INVALID = -1, /* bad argument (maps to ValueError / NotImplementedError in the Python shim) */
  PSC_ERR_CUDA = -2,    /* CUDA runtime error */
  PSC_ERR_CUFFT = -3,   /* cuFFT error */
  PSC_ERR_WORKSPACE = -4 /* caller-provided scratch too small */
} psc_status;

/* mass-assignment schemes (param["mass_scheme"], MAS_index = scheme + 1 for CIC/TSC) */
enum { PSC_NGP = 0, PSC_CIC = 1, PSC_TSC = 2 };
/* Green's functions: fourier.inverse_laplacian / _compensated / _7pt */
enum { PSC_GREEN_PLAIN = 0, PSC_GREEN_COMPENSATED = 1, PSC_GREEN_7PT = 2 };
/* smoother / operator families: laplacian.py, cubic.py (f(R) n=1), quartic.py (f(R) n=2) */
enum { PSC_OP_LAPLACIAN = 0, PSC_OP_CUBIC = 1, PSC_OP_QUARTIC = 2 };
/* QUMOND interpolating functions: mond.py:16-162 */
enum { PSC_MOND_SIMPLE = 0, PSC_MOND_N = 1, PSC_MOND_BETA = 2, PSC_MOND_GAMMA = 3, PSC_MOND_DELTA = 4 };

const char *psc_last_error(void);
int psc_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t psc_launch_count(void);

/* ---------------------------------------------------------------- particles ---------------- */
/* morton.positions_to_keys (morton.py:42-137): 21 bits per axis, key = X<<2 | Y<<1 | Z */
int psc_morton_keys(const float *pos, int64_t np, int64_t *keys, void *stream);
/* stable argsort of int64 keys, as utils.reorder_particles with nthreads == 1 (utils.py:1019-1075,
 * np.argsort).  idx_out[i] = source row of sorted row i.  scratch from psc_argsort_workspace_bytes. */
size_t psc_argsort_workspace_bytes(int64_t np);
int psc_argsort_keys(const int64_t *keys, int64_t np, int64_t *idx_out, void *scratch,
                     size_t scratch_bytes, void *stream);
/* utils.injection_with_indices (utils.py:894-924) on [Np,3] rows: dst[i] = src[idx[i]] */
int psc_gather3(const int64_t *idx, const float *src, float *dst, int64_t np, void *stream);
/* utils.add_vector_scalar_inplace (utils.py:264-297): y += a*x.  a_is_f64 selects the float64-scalar
 * typing of the reference (drift with a snapshot-clamped dt, integration.py:84-85, 252) */
int psc_axpy(float *y, const float *x, double a, int a_is_f64, int64_t n, void *stream);
/* utils.periodic_wrap (utils.py:1120-1149) */
int psc_periodic_wrap(float *x, int64_t n, void *stream);
/* utils.max_abs (utils.py:220-240): *out = max(*out, max|x|); caller zeroes *out (device float) */
int psc_max_abs(const float *x, int64_t n, float *out, void *stream);
/* fused first half of integration.leapfrog (integration.py:250-258): v -= half_dt*a; x += dt*v;
 * periodic_wrap(x) */
int psc_kick_drift_wrap(float *pos, float *vel, const float *acc, int64_t np, float half_dt,
                        double dt, int dt_is_f64, void *stream);

/* ---------------------------------------------------------------- mesh <-> particles ------- */
/* mesh.NGP/CIC/TSC/TSC_seq (mesh.py:2240-2595) followed by solver.pm's density rescale and
 * rhs_poisson's affine map (solver.py:114-116, 444-449): rho = f1*(scale*deposit) + f2.
 * Pass scale = 1, f1 = 1, f2 = 0 for the bare mass assignment.  rho is overwritten. */
int psc_deposit(const float *pos, int64_t np, int N, int scheme, float scale, float f1, float f2,
                float *rho, void *stream);
/* mesh.invNGP/invCIC/invTSC and their _vec forms (mesh.py:2600-3088); ncomp = 1 or 3 */
int psc_interp(const float *grid, const float *pos, int64_t np, int N, int ncomp, int scheme,
               float *out, void *stream);
/* mesh.inv{CIC,TSC}_vec fused with the second half-kick and the two reductions the next
 * integration.integrate needs (solver.py:205-213; integration.py:262, 293-295, 324-326):
 * acc = interp(force, pos); vel -= half_dt*acc; maxout[0] = max|acc|, maxout[1] = max|vel|
 * (maxout is max-combined: caller zeroes it).  half_dt = 0 and vel = NULL gives plain interp+max. */
int psc_interp_kick(const float *force, const float *pos, float *vel, float *acc, int64_t np, int N,
                    int scheme, float half_dt, float *maxout, void *stream);

/* same on a float4-padded force grid [N,N,N,4] produced by psc_gradient(out_stride = 4); CIC or TSC */
int psc_interp_kick4(const float *force4, const float *pos, float *vel, float *acc, int64_t np, int N,
                     int scheme, float half_dt, float *maxout, void *stream);

/* Order-independent variants on a per-step binning of the particles into 8^3-cell bins (the particle
 * arrays themselves keep the reference's order).  psc_bin_particles fills `scratch` (size from
 * psc_bin_workspace_bytes, 256-byte aligned) with bin offsets, the binned copy of the positions and the
 * source row of every binned particle; the two kernels below consume it.  N must be a multiple of 8.
 * psc_deposit_binned == psc_deposit, psc_interp_kick4_binned == psc_interp_kick4 (results are written to the
 * particles' original rows). */
size_t psc_bin_workspace_bytes(int64_t np, int N);
int psc_bin_particles(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, void *stream);
int psc_deposit_binned(const void *scratch, size_t scratch_bytes, int64_t np, int N, int scheme, float scale,
                       float f1, float f2, float *rho, void *stream);
int psc_interp_kick4_binned(const float *force4, const void *scratch, size_t scratch_bytes, float *vel, float *acc,
                            int64_t np, int N, int scheme, float half_dt, float *maxout, void *stream);

/* psc_kick_drift_wrap with the binning of the NEW positions folded in, on the bin scratch of np_total particles.
 * mode 0: per-bin COUNT (the binning's first pass then costs no extra read of the positions);
 * psc_bin_particles_counted(mode 0) finishes with scan + scatter.
 * mode 1: DIRECT scatter -- the scratch still holds the previous step's binning of (about) the same particles: every
 * bin gets its previous fill + 1/8 + 32 records of room and each particle drops its (x, y, z, row) record straight
 * into its bin, so there is no count pass and no second read of the positions; psc_bin_particles_counted(mode 1)
 * finishes (heavy-bin list) and, when a bin ran out of room, redoes the exact binning on the device.
 * zero_counts: 1 for the first (or only) chunk of a step, 0 for the later chunks of a chunked upload; row0: global row
 * of pos[0]. */
int psc_kick_drift_wrap_count(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                              int dt_is_f64, int N, int64_t np_total, void *scratch, size_t scratch_bytes,
                              int zero_counts, int mode, int64_t row0, void *stream);
int psc_bin_particles_counted(const float *pos, int64_t np, int N, void *scratch, size_t scratch_bytes, int mode,
                              void *stream);

/* ---- the time loop on particle arrays kept in BIN order (no binned copy, no source-row indirection).
 * psc_step_sort: integration.leapfrog's first half (integration.py:250-258: v -= half_dt a; x += dt v; periodic_wrap)
 * fused with a counting sort of the particles into 8^3-cell bins: pass 1 counts the NEW positions (nothing written),
 * scan, pass 2 redoes the arithmetic and writes x', v' and the particle's id (ids == NULL: the input row) to its row of
 * the bin-ordered output arrays (out of place).  `scratch` (psc_sorted_workspace_bytes) then holds the bin table that
 * psc_deposit_sorted (mesh.TSC/CIC/NGP + rescale + rhs_poisson's affine map, as psc_deposit_binned) and
 * psc_interp_kick_phi_sorted (as psc_interp_kick_phi_binned; vel / acc rows are the rows of pos_sorted) consume.
 * psc_scatter3_by_id: out[ids[n]] = in[n], the way back to the reference's particle order. */
size_t psc_sorted_workspace_bytes(int64_t np, int N);
/* The scratch holds TWO bin tables.  src_table = -1: the input arrays are in no particular order (first step, after a
 * Morton reorder): one global atomic per particle and pass, result in table 0.  src_table = 0 / 1: the input arrays
 * are the bin-ordered output of a previous call, described by that table: one CTA per source bin counts and sorts its
 * particles in shared memory by (destination bin, 2^3-cell micro-block in Morton order) and writes contiguous runs;
 * result in table 1 - src_table.  `table` of the two consumers = the table the sort wrote. */
int psc_step_sort(const float *pos, const float *vel, const float *acc, const int *ids, int64_t np, float half_dt,
                  double dt, int dt_is_f64, int N, int src_table, int counts_ready, void *scratch,
                  size_t scratch_bytes, float *pos_out, float *vel_out, int *ids_out, void *stream);
int psc_deposit_sorted(const float *pos_sorted, const void *scratch, size_t scratch_bytes, int table, int64_t np, int N,
                       int scheme, float scale, float f1, float f2, float *rho, void *stream);
/* predict = 1: the kernel also applies the NEXT step's first half (next_half_dt, next_dt, next_dt_is_f64: the values
 * the next psc_step_sort will be called with -- known in advance when the scale-factor criterion of
 * integration.py:329-358 binds) to every particle in registers and counts its destination bin, into the scratch's count
 * table; that psc_step_sort is then called with counts_ready = 1 and skips its count pass.  If the host ends up with
 * another time step it passes counts_ready = 0 and the sort counts again. */
int psc_interp_kick_phi_sorted(const float *phi, const float *u, float f, int fr_n, int order, const float *pos_sorted,
                               const void *scratch, size_t scratch_bytes, int table, float *vel_sorted,
                               float *acc_sorted, int64_t np, int N, int scheme, float half_dt, float *maxout,
                               int predict, float next_half_dt, double next_dt, int next_dt_is_f64, void *stream);
/* utils.reorder_particles (utils.py:1019-1075) for bin-ordered arrays: no particle moves; ids_out[n] = rank of the
 * particle in row n in Morton-key order, i.e. its row in the reference after the reorder.  Per-bin block radix sort in
 * shared memory + a scan of the bin counts in Z order.  N: power of two.  *too_big (device int) = 1 when a bin holds
 * more than 2048 particles (the caller then sorts globally: psc_morton_keys + psc_argsort_keys). */
int psc_morton_ids_sorted(const float *pos_sorted, void *scratch, size_t scratch_bytes, int table, int64_t np, int N,
                          int *ids_out, int *too_big, void *stream);
int psc_scatter3_by_id(const int *ids, const float *in, float *out, int64_t np, void *stream);
/* the bin-ordered layout on a slab: after the migration the rank's (position, velocity, 64-bit id) arrays are sorted
 * into bin order (out of place); the deposit / interpolation then read them in place, as the two entries above.
 * src_table / src_rows: -1 / 0 for arrays in no particular order (result: table 0); else the table that describes rows
 * [0, src_rows) of the input as the previous sort left them -- kicked, drifted and migrated in place since (result:
 * table 1 - src_table; per-source-bin sort in shared memory).  `table` of the consumers = the table the sort wrote. */
size_t psc_sorted_workspace_bytes_slab(int64_t np, int N, int nxl);
int psc_sort_by_bin_slab(const float *pos, const float *vel, const int64_t *ids, int64_t np, int N, int x0, int nxl,
                         int src_table, int64_t src_rows, void *scratch, size_t scratch_bytes, float *pos_out,
                         float *vel_out, int64_t *ids_out, void *stream);
int psc_deposit_sorted_slab(const float *pos_sorted, const void *scratch, size_t scratch_bytes, int table, int64_t np,
                            int N, int x0, int nxl, int scheme, float *rho_ghost, void *stream);
int psc_interp_kick_phi_sorted_slab(const float *phi_ghost, const float *u_ghost, float f, int fr_n, int order, int x0,
                                    int nxl, int ghost, const float *pos_sorted, const void *scratch,
                                    size_t scratch_bytes, int table, float *vel_sorted, float *acc_sorted, int64_t np,
                                    int N, int scheme, float half_dt, float *maxout, void *stream);

/* mesh.derivative / derivative_fR (mesh.py:639-2174) fused into the binned interpolation + kick: every bin's CTA
 * derives its force tile from the potential phi (and, for f(R), the scalaron u: phi + f * u^(fr_n+1)) in shared
 * memory, so neither the gradient kernel nor the force grid exist in the step.  fr_n = 0: plain. */
int psc_interp_kick_phi_binned(const float *phi, const float *u, float f, int fr_n, int order, const void *scratch,
                               size_t scratch_bytes, float *vel, float *acc, int64_t np, int N, int scheme,
                               float half_dt, float *maxout, void *stream);

/* ---------------------------------------------------------------- x-slab decomposition ------ */
/* The reference is single-process (README.md:49).  These entry points are the per-GPU pieces of the slab
 * decomposition of its hot path: rank r of P owns the planes [x0, x0 + nxl) = [r N/P, (r+1) N/P)
 * of every grid and the particles whose cell floor(x N) lies in them.  Collectives (NCCL) are issued by the
 * host (pysco_b200/slab.py); nothing here communicates.
 *
 * Binned particle <-> mesh kernels on a slab: bins cover the owned planes only (nxl % 8 == 0).
 * psc_deposit_binned_slab writes raw TSC/CIC/NGP sums into rho_ghost[nxl + 2][N][N]; plane 0 and plane
 * nxl + 1 hold what belongs to the left / right neighbour's last / first owned plane (mesh.py:2468-2595 with
 * the periodic wrap in x replaced by ghost planes).  psc_interp_kick_phi_binned_slab reads phi_ghost (and
 * u_ghost) [nxl + 2*ghost][N][N] whose owned planes start at plane `ghost` >= 1 + stencil reach. */
size_t psc_bin_workspace_bytes_slab(int64_t np, int N, int nxl);
int psc_bin_particles_slab(const float *pos, int64_t np, int N, int x0, int nxl, void *scratch, size_t scratch_bytes,
                           void *stream);
int psc_deposit_binned_slab(const void *scratch, size_t scratch_bytes, int64_t np, int N, int x0, int nxl, int scheme,
                            float *rho_ghost, void *stream);
int psc_interp_kick_phi_binned_slab(const float *phi_ghost, const float *u_ghost, float f, int fr_n, int order,
                                    int x0, int nxl, int ghost, const void *scratch, size_t scratch_bytes, float *vel,
                                    float *acc, int64_t np, int N, int scheme, float half_dt, float *maxout,
                                    void *stream);
/* Particle migration after the drift (integration.py:252-258; utils.periodic_wrap utils.py:1120).
 * psc_slab_count: counts[d] (device int64[P], overwritten) = particles whose owner floor(x N) / nxl is rank
 *   d != me; counts[me] is left 0 (the stayers are np minus the others).
 * psc_slab_pack_leavers: every particle not owned by `me` is packed as a 32-byte record (x y z vx vy vz id)
 *   at sendbuf[8 * (offsets[owner] + slot)], its row stored in holes[same index]; offsets = exclusive prefix
 *   of the send counts by destination (device int64[P]); cursor = device int64[P] scratch.
 * psc_slab_unpack_rows: record t of recvbuf -> row rows[t] of pos / vel / ids.
 * psc_slab_move_rows: row src[t] -> row dst[t] (fills the holes that the arrivals did not fill). */
/* psc_kick_drift_wrap on a slab with the leavers detected in the same pass: counts[0..P) (device int64[P + 1],
 * overwritten) = leavers per destination, counts[P] = number of leavers; their rows are appended to
 * leaver_rows[capacity] (unordered).  If counts[P] > capacity the list is incomplete: fall back to
 * psc_slab_pack_leavers.  psc_slab_pack_rows = psc_slab_pack_leavers over such a list. */
int psc_kick_drift_wrap_slab(float *pos, float *vel, const float *acc, int64_t np, float half_dt, double dt,
                             int dt_is_f64, int N, int nxl, int P, int me, int64_t *counts, int64_t *leaver_rows,
                             int64_t capacity, void *stream);
int psc_slab_pack_rows(const float *pos, const float *vel, const int64_t *ids, const int64_t *rows, int64_t nrows,
                       int N, int nxl, int P, int me, const int64_t *offsets, int64_t *cursor, float *sendbuf,
                       int64_t *holes, void *stream);
/* Single-round neighbour migration: leavers towards the left / right slab are packed into fixed-capacity buffers
 * sendL / sendR [(cap + 1) records]; record 0 is a header holding the true count (int64), records 1.. the leavers,
 * holesL / holesR [cap] their rows.  rows / counts / list_capacity = the candidate list of
 * psc_kick_drift_wrap_slab (rows may be NULL: all np particles are scanned; an incomplete list is detected on the
 * device and also falls back to the scan).  status (device int64[3]) = leavers to the left, to the right (either may
 * exceed cap: the excess is not packed and the exchange must be redone with a larger cap), and leavers to a
 * non-neighbour (error).  P == 2: both neighbours are the same rank and every leaver counts as "left". */
int psc_slab_pack_fixed(const float *pos, const float *vel, const int64_t *ids, int64_t np, const int64_t *rows,
                        const int64_t *counts, int64_t list_capacity, int N, int nxl, int P, int me, int64_t cap,
                        float *sendL, float *sendR, int64_t *holesL, int64_t *holesR, int64_t *status, void *stream);
int psc_slab_count(const float *pos, int64_t np, int N, int nxl, int P, int me, int64_t *counts, void *stream);
int psc_slab_pack_leavers(const float *pos, const float *vel, const int64_t *ids, int64_t np, int N, int nxl, int P,
                          int me, const int64_t *offsets, int64_t *cursor, float *sendbuf, int64_t *holes,
                          void *stream);
int psc_slab_unpack_rows(const float *recvbuf, const int64_t *rows, int64_t n, float *pos, float *vel, int64_t *ids,
                         void *stream);
int psc_slab_move_rows(const int64_t *src, const int64_t *dst, int64_t n, float *pos, float *vel, int64_t *ids,
                       void *stream);
/* Transposed FFT of fourier.fft_3D_real / ifft_3D_real (fourier.py:104-147, 251-294) for one slab:
 *   forward: psc_slab_fft_r2c_planes ([nxl][N][N] -> [nxl][N][N/2+1]) -> psc_slab_yblocks(to_blocks = 1)
 *            ([P][nxl][nyl][N/2+1]) -> all-to-all (host) = [N][nyl][N/2+1] -> psc_slab_fft_x(inverse = 0)
 *   psc_green_slab on the transposed layout, then the mirror image back.  All transforms are unnormalised.
 * The three cuFFT plans share ONE caller-provided work area (psc_slab_fft_workspace_bytes /
 * psc_slab_fft_set_workspace). */
int psc_slab_fft_plan_create(int N, int nxl, int nyl, void **plan_out);
int psc_slab_fft_plan_destroy(void *plan);
size_t psc_slab_fft_workspace_bytes(void *plan);
int psc_slab_fft_set_workspace(void *plan, void *work);
int psc_slab_fft_r2c_planes(void *plan, const float *planes, float *spec2d, void *stream);
int psc_slab_fft_c2r_planes(void *plan, float *spec2d, float *planes, void *stream);
int psc_slab_fft_x(void *plan, float *spec_t, int inverse, void *stream);
int psc_slab_yblocks(const float *in, float *out, int N, int nxl, int nyl, int to_blocks, void *stream);
/* The same transposes as ONE kernel over NVLink peer memory (no pack pass, no NCCL all-to-all): peer_ptrs_dev is a
 * DEVICE array of P pointers, entry d = rank d's receive buffer mapped into this process (symmetric memory).
 * forward != 0: in = [nxl][N][N/2+1] -> the destinations' [N][nyl][N/2+1] buffers; forward == 0: in =
 * [N][nyl][N/2+1] -> the destinations' [nxl][N][N/2+1] buffers.  The caller orders the kernel against the peers'
 * use of the buffers with inter-GPU barriers (pysco_b200/slab.py). */
int psc_slab_transpose_put(const float *in, const void *peer_ptrs_dev, int N, int nxl, int nyl, int P, int me,
                           int forward, void *stream);
/* psc_green / psc_pk on a y-block of the transposed spectrum [N (kx)][nyl (ky = y0 + .)][N/2+1]; the P(k) bins of the
 * ranks are summed by the host (all-reduce) */
int psc_pk_slab(float *spec_t, int N, int nyl, int y0, int p, double *bins, void *stream);
int psc_green_slab(float *spec_t, int N, int nyl, int y0, int kind, int p, float scale, void *stream);

/* ---------------------------------------------------------------- grid algebra ------------- */
/* utils.linear_operator[_inplace] (utils.py:644-717): out = f1*x + f2 (out may alias x) */
int psc_linear_operator(const float *x, float f1, float f2, float *out, int64_t n, void *stream);
/* utils.linear_operator_vectors_inplace (utils.py:721-755): x = f1*x + f2*y */
int psc_lincomb(float *x, float f1, const float *y, float f2, int64_t n, void *stream);
/* mesh.derivative / derivative_fR / add_derivative_fR (mesh.py:639-2237).  order in {2,3,5,7};
 * fr_n = 0 (plain), 1 (a + f*b^2), 2 (a + f*b^3); add != 0: force += f * grad(b^(fr_n+1)).
 * out_stride = 3: the reference's AoS [N,N,N,3]; out_stride = 4: float4-padded [N,N,N,4] (fx,fy,fz,0),
 * the internal layout consumed by psc_interp_kick4 (one 16-byte load per stencil point) */
int psc_gradient(const float *a, const float *b, float f, int fr_n, int order, int add, int N,
                 float *force, int out_stride, void *stream);

/* ---------------------------------------------------------------- Fourier ------------------ */
/* fourier.fft_3D_real / ifft_3D_real (fourier.py:104-147, 251-294): cuFFT R2C / C2R plans for an
 * N^3 grid.  C2R overwrites its input (cuFFT semantics) and is UNNORMALISED: fold 1/N^3 into the
 * `scale` of psc_green / psc_grad_green, or call psc_linear_operator afterwards. */
int psc_fft_plan_create(int N, void **plan_out);
int psc_fft_plan_destroy(void *plan);
size_t psc_fft_plan_workspace_bytes(void *plan);
int psc_fft_r2c(void *plan, const float *in, float *spec_out, void *stream);
int psc_fft_c2r(void *plan, float *spec_in, float *out, void *stream);
/* fourier.ifft_3D_real_grad (fourier.py:372-410): three C2R transforms of an interleaved
 * [N,N,N/2+1,3] spectrum into an AoS [N,N,N,3] grid */
/* solver.fft (solver.py:444-500) in one call: out = irfftn(G * rfftn(rhs)) with G = Green's function (kind) x
 * W^-2p x scale.  cuFFT runs the batched 2-D (y, z) transforms of the x planes; the forward and backward transforms
 * along x and the Green multiply are ONE kernel (shared-memory radix-8 FFT, 5 passes over the spectrum instead of 7).
 * N = a power of two in [64, 2048] (psc_fft_poisson_supported); spec = [N, N, N/2+1] complex64 scratch; out may alias
 * rhs.  psc_xfft_green_slab: the same kernel on the transposed spectrum [N (kx)][nyl (ky = y0 ..)][N/2+1] of the
 * slab-decomposed solve, replacing psc_slab_fft_x (forward) + psc_green_slab + psc_slab_fft_x (backward). */
int psc_fft_poisson_supported(int N);
int psc_fft_poisson(void *plan, const float *rhs, float *spec, float *out, int kind, int p, float scale, void *stream);
int psc_xfft_green_slab(float *spec_t, int N, int nyl, int y0, int kind, int p, float scale, void *stream);
int psc_fft_c2r_vec3(void *plan, float *spec3_in, float *out3, void *stream);
/* fourier.inverse_laplacian / _compensated / _7pt (fourier.py:460-595), in place, DC zeroed;
 * every mode is additionally multiplied by `scale` */
int psc_green(float *spec, int N, int kind, int p, float scale, void *stream);
/* fourier.gradient_inverse_laplacian[_compensated] (fourier.py:606-719): out3[...,d] =
 * -i k_d/(2 pi k^2) W^-2p spec * scale; p = 0 means uncompensated */
int psc_grad_green(const float *spec, int N, int p, float scale, float *out3, void *stream);
/* fourier.fourier_grid_to_Pk (fourier.py:22-100): bins[3][N] (device, double) receive per nearest-
 * integer |k| bin: sum |k|, sum |delta_k W^-p|^2, mode count; also zeroes spec[0,0,0] like the
 * reference.  The host shim divides and slices bins 1..int(2*(N/2)/3)-1. bins is overwritten. */
int psc_pk(float *spec, int N, int p, double *bins, void *stream);

/* ---------------------------------------------------------------- multigrid ---------------- */
/* laplacian.operator (laplacian.py:12-54) / cubic.operator (cubic.py:23-81) / quartic.operator
 * (quartic.py:20-77).  kind = PSC_OP_*; b and q ignored for the Laplacian. */
int psc_operator(const float *x, const float *b, float q, int N, int kind, float *out, void *stream);
/* laplacian.residual (laplacian.py:63-117): b - Lx;  cubic/quartic.residual_with_rhs
 * (cubic.py:90-154, quartic.py:86-149): rhs - L(x; b, q) */
int psc_residual(const float *x, const float *b, float q, const float *rhs, int N, int kind,
                 float *out, void *stream);
/* laplacian.restrict_residual (laplacian.py:125-226): coarse = R(b - Lx), fused */
int psc_restrict_residual(const float *x, const float *b, int N, float *coarse, void *stream);
/* laplacian.residual_error (laplacian.py:327-381), cubic/quartic.residual_error (cubic.py:844-901,
 * quartic.py:844-901): *sumsq_out (device double) += sum of squared residuals; caller zeroes it and
 * takes the square root. */
int psc_residual_sumsq(const float *x, const float *b, float q, int N, int kind, double *sumsq_out,
                       void *stream);
/* sum (fa*a - b)^2 into *sumsq_out: the reductions of laplacian.truncation_error
 * (laplacian.py:502-533, fa = 1) and cubic/quartic.truncation_error (cubic.py:1021-1061, fa = 4) */
int psc_diff_sumsq(const float *a, float fa, const float *b, int64_t n, double *sumsq_out, void *stream);
/* laplacian.initialise_potential (laplacian.py:765-796) kind 0; cubic/quartic.initialise_potential
 * (cubic.py:217-259, quartic.py:214-260) kinds 1, 2 */
int psc_initialise_potential(const float *b, float q, int N, int kind, float *out, void *stream);
/* one red-black SOR sweep: laplacian.gauss_seidel (laplacian.py:844-1022), cubic.gauss_seidel[_with_rhs]
 * (cubic.py:269-627), quartic.gauss_seidel[_with_rhs] (quartic.py:270-628).  rhs may be NULL. */
int psc_gauss_seidel(float *x, const float *b, float q, const float *rhs, int N, int kind,
                     float f_relax, void *stream);
/* The same sweep as ONE plane-marching kernel (csrc/gs_fused.cu): a CTA stages four planes of x in shared memory (TMA
 * bulk copies when use_tma != 0, LDG/STS otherwise), updates the red cells of plane s + 1 and then the black cells of
 * plane s, and writes the finished plane to x_out -- 13 B per cell instead of the 24 B of the two colour passes,
 * bit-identical result.  OUT OF PLACE (x_out != x).  Needs N >= 128 and N % 64 == 0
 * (psc_gauss_seidel_fused_supported). */
int psc_gauss_seidel_fused_supported(int N);
int psc_gauss_seidel_fused(const float *x, const float *b, float q, const float *rhs, int N, int kind, float f_relax,
                           float *x_out, int use_tma, void *stream);
const float *psc_mg_q_device_ptr(void);
/* While set (non-NULL) the f(R) kernels above read q from this device float instead of their by-value argument, so
 * that a CUDA graph captured over a FAS cycle follows q from step to step; NULL restores the by-value behaviour. */
int psc_mg_set_q_device(const float *q_dev);
/* mesh.restriction / minus_restriction (mesh.py:14-108): coarse = sign/8 * sum of 8 children */
int psc_restriction(const float *x, int N, float sign, float *coarse, void *stream);
/* mesh.prolongation / add_prolongation (mesh.py:180-453): fine (2Nc) (+)= P(coarse Nc) */
int psc_prolongation(float *fine, const float *coarse, int Nc, int add, void *stream);
/* mond.rhs_simple/n/beta/gamma/delta (mond.py:171-932) */
int psc_mond_rhs(const float *phi, float *out, int N, float g0, int fn, float alpha, void *stream);

/* ---- multigrid on an x-slab (ghost-plane stencils).  A slab level holds nxl owned planes of n x n cells;
 * "xg" arrays are [nxl + 2][n][n] with one ghost plane on each side along x (the host fills them from the neighbours,
 * pysco_b200/slab.py), "b" / outputs hold the owned planes only; y and z are periodic.  Same arithmetic as the
 * single-domain entries above. */
/* one colour of laplacian.gauss_seidel (laplacian.py:844-1022): colour 1 = cells with odd x0 + il + j + k (the
 * reference's first half-sweep), colour 0 = even; x0 = global index of the first owned plane.  The ghost planes must
 * be refreshed between the two colours. */
int psc_box_gauss_seidel_colour(float *xg, const float *b, int nxl, int n, int x0, int colour, float f_relax,
                                void *stream);
/* laplacian.operator (laplacian.py:12-54): out[nxl][n][n] = L xg */
int psc_box_operator(const float *xg, int nxl, int n, float *out, void *stream);
/* laplacian.restrict_residual (laplacian.py:125-226): coarse[nxl/2][n/2][n/2] = R(b - L xg); nxl even */
int psc_box_restrict_residual(const float *xg, const float *b, int nxl, int n, float *coarse, void *stream);
/* mesh.restriction / minus_restriction (mesh.py:14-108) of owned planes: coarse[nxl/2][n/2][n/2]; nxl even */
int psc_box_restriction(const float *fine, int nxl, int n, float sign, float *coarse, void *stream);
/* mesh.add_prolongation (mesh.py:334-453): fine_g[2 nxlc + 2][2 nc][2 nc] += P(coarse_g[nxlc + 2][nc][nc]); the coarse
 * ghost planes must be current, the fine ghost planes are not touched */
int psc_box_add_prolongation(float *fine_g, const float *coarse_g, int nxlc, int nc, void *stream);
/* f(R) scalaron on the slab, kind = PSC_OP_CUBIC (fR_n = 1) or PSC_OP_QUARTIC (fR_n = 2): one colour of
 * cubic/quartic.gauss_seidel[_with_rhs] (cubic.py:269-627, quartic.py:270-628; rhs may be NULL), cubic/quartic.operator
 * (cubic.py:23-81) and cubic/quartic.initialise_potential (cubic.py:217-259, quartic.py:214-260; b, out = owned planes) */
int psc_box_gauss_seidel_colour_fr(float *xg, const float *b, const float *rhs, float q, int nxl, int n, int x0,
                                   int colour, float f_relax, int kind, void *stream);
int psc_box_operator_fr(const float *xg, const float *b, float q, int nxl, int n, int kind, float *out, void *stream);
int psc_box_initialise_potential_fr(const float *b, float q, int nxl, int n, int kind, float *out, void *stream);
/* mond.rhs_simple/n/beta/gamma/delta (mond.py:171-932) on the slab: phig[nxl + 2][n][n] is the Newtonian potential with
 * its ghost planes, out[nxl][n][n] the QUMOND source */
int psc_box_mond_rhs(const float *phig, float *out, int nxl, int n, float g0, int fn, float alpha, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSCO_B200_H */
