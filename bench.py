#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full Newtonian FFT particle-mesh step (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --steps 3 --warmup 1     # CPU arm: the oracle port on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU (x-slab decomposition)

Workload: N^3 particles on an N^3 mesh (default N = 512), synthetic ICs = cell-centre lattice + Gaussian
displacement 0.3 cells, velocities N(0, 1e-3) per particle and component (counter-based generator `hash_normal`,
the SAME function in the NumPy (CPU arm) and the torch (GPU arms, per x-slab) flavour), Morton-ordered as after a
reorder, TSC deposit, compensated-Green FFT solve, 5-point gradient, TSC interpolation, leapfrog KDK through the
public API `pysco_b200.integration.integrate` (N = 1) / `pysco_b200.slab.Slab.integrate` (N > 1) with analytic
background tables; `utils.reorder_particles` every n_reorder = 50 steps.  The K timed steps sit in the middle of a
reorder cycle ((50 - K) / 2 untimed steps after a reorder come first) and the measured reorder time is amortised as
K / 50 reorders; K = 50 times one whole cycle including its reorder.

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same call with HOST (pinned)
buffers; `extra` (N = 1) = the same step on the coherent-flow and clustered particle sets, through the slab path on
one rank, and BASELINE configs 2-4; `config5` (N = 8) = BASELINE config 5 (2048^3 on 8 GPUs).
STRONG scaling: the same N^3 problem on every GPU count.
"""
import argparse
import gc
import json
import os

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")   # no lazy kernel loads inside the timed steps
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

METRIC = "particle_updates_per_sec_full_pm_step"
UNIT = "particle-updates/s"
N_REORDER = 50

# algorithmic bytes per particle (= per cell, Np = N^3) of each C-ABI call of the Newtonian FFT step
# (SURVEY 8d): the figure `roofline.achieved` is computed from.
ALGO_BYTES = {
    "psc_kick_drift_wrap": 60.0,   # read x,v,a (36) + write x,v (24)
    "psc_kick_drift_wrap_count": 60.0,   # same pass + the per-bin counts of the new positions
    "psc_bin_particles_counted": 0.0,    # scan + scatter of the per-step binning: overhead, not in the 176 B budget
    "psc_deposit": 16.0,           # read x (12) + write rho (4); rescale + RHS affine fused (0)
    "psc_fft_r2c": 8.0,            # read 4 + write 4 (half-spectrum ~ 4 B per real cell)
    "psc_green": 8.0,
    "psc_fft_c2r": 8.0,
    "psc_fft_poisson": 24.0,       # = r2c + Green + c2r in one call (the x passes and the Green multiply fused)
    "psc_gradient": 16.0,          # read phi (4) + write force (12)
    "psc_interp_kick4": 60.0,      # read x,v (24) + force once per cell (12) + write v,a (24)
    "psc_interp_kick4_binned": 60.0,
    "psc_interp_kick_phi_binned": 76.0,   # gradient (16) + interpolation/kick (60) in one kernel
    "psc_deposit_binned": 16.0,
    "psc_bin_particles": 0.0,
    # the bin-ordered time loop: kick + drift + wrap fused with the re-sort of the particle arrays (60 B algorithmic;
    # the second read of x, v, a and the ids are the price of the order, not in the 176 B budget)
    "psc_step_sort": 60.0,
    "psc_deposit_sorted": 16.0,
    "psc_interp_kick_phi_sorted": 76.0,
    # x-slab path (every kernel works on N^3 / P particles or cells)
    "psc_kick_drift_wrap_slab": 60.0,      # kick + drift + wrap + leaver detection
    "psc_bin_particles_slab": 0.0,
    "psc_deposit_binned_slab": 16.0,
    "psc_linear_operator": 8.0,            # RHS affine map (fused into the deposit on the single-domain path)
    "psc_slab_fft_r2c_planes": 8.0,        # 2-D R2C of the owned planes
    "psc_slab_yblocks": 0.0,               # pack / unpack around the all-to-all: overhead
    "psc_slab_fft_x": 8.0,                 # 1-D C2C along x on the transposed spectrum
    "psc_green_slab": 8.0,
    "psc_slab_fft_c2r_planes": 8.0,
    "psc_interp_kick_phi_binned_slab": 76.0,
    "psc_slab_count": 0.0, "psc_slab_pack_leavers": 0.0, "psc_slab_unpack_rows": 0.0, "psc_slab_move_rows": 0.0,
    # multigrid / f(R) / MOND kernels (configs 2-4), per cell of the grid the call works on
    "psc_gauss_seidel": 12.0,              # one red + black sweep: read x, b + write x
    "psc_restrict_residual": 8.5,          # read x, b (8) + write the coarse grid (1/8 * 4)
    "psc_prolongation": 8.5,               # read the coarse grid (0.5) + read / write the fine grid (8)
    "psc_residual_sumsq": 8.0,
    "psc_mond_rhs": 8.0,
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this workload
# (512^3, one GPU), filled from profiles/ by load_ncu_traffic()
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic_512.json")

# whole Newtonian FFT step: kick/drift/wrap 60 + deposit 16 + FFT 8 + Green 8 + inverse FFT 8 + gradient 16 +
# interpolation/kick 60 = 176 B per particle-update
STEP_ALGO_BYTES = 176.0


def make_tables():
    """Analytic background (EdS-like supercomoving time, D1 = a); same layout as cosmotable.generate."""
    from scipy.interpolate import interp1d
    lna = np.linspace(np.log(1.0 / 201), 0.0, 20001)
    a = np.exp(lna)
    t = -2.0 * (a ** -0.5 - 1.0)
    tabs = [interp1d(t, lna, fill_value="extrapolate"), interp1d(lna, t, fill_value="extrapolate"),
            interp1d(lna, 72.0 * a ** -1.5, fill_value="extrapolate"), interp1d(lna, a, fill_value="extrapolate")]
    return tabs + [interp1d(lna, np.ones_like(a), fill_value="extrapolate")] * 9


def make_param(ncoarse, nthreads, **over):
    import pandas as pd
    N = 2 ** ncoarse
    p = {
        "nthreads": nthreads, "theory": "newton", "H0": 72, "Om_m": 0.25733, "boxlen": 100, "ncoarse": ncoarse,
        "npart": N ** 3, "integrator": "leapfrog", "mass_scheme": "TSC", "n_reorder": N_REORDER,
        "Courant_factor": 1.0, "max_aexp_stepping": 10, "linear_newton_solver": "fft",
        "gradient_stencil_order": 5, "save_power_spectrum": "no", "aexp": 0.02, "aexp_old": 0.02, "nsteps": 0,
        "write_snapshot": False, "w0": -1.0, "wa": 0.0, "Om_r": 8.0763e-05, "Om_lambda": 1 - 0.25733 - 8.0763e-05,
        "parametrized_mu0": 0.0, "epsrel": 1e-2, "Npre": 2, "Npost": 1,
    }
    p.update(over)
    return pd.Series(p)


# ------------------------------------------------------------------------------- synthetic ICs
_M32 = 0xFFFFFFFF


def _mix32(x):
    """murmur3 finaliser on uint32 values held in int64 (NumPy arrays or torch tensors alike)"""
    x = ((x ^ (x >> 16)) * 0x85EBCA6B) & _M32
    x = ((x ^ (x >> 13)) * 0xC2B2AE35) & _M32
    return x ^ (x >> 16)


def _hash_uniform(idx, stream, seed):
    """uniform in (0, 1] from the particle's global index (int64), a stream number and the seed"""
    salt = (stream * 0x9E3779B1 + seed * 0x85EBCA77 + 0x27D4EB2F) & _M32
    h = _mix32((idx & _M32) ^ _mix32(((idx >> 32) + salt) & _M32))
    return h


def hash_normal(idx, stream, seed, like):
    """N(0, 1) per element of `idx` (global particle index) -- counter-based (Box-Muller on two hashed uniforms), so
    any slab of the particle set can be generated on its own, on the host (NumPy) or on the device (torch), and both
    arms of the bench see the same particles.  like: np or torch."""
    two32 = 1.0 / 4294967296.0
    h1, h2 = _hash_uniform(idx, 2 * stream, seed), _hash_uniform(idx, 2 * stream + 1, seed)
    if like is np:
        u1 = (h1.astype(np.float64) + 0.5) * two32
        u2 = (h2.astype(np.float64) + 0.5) * two32
        return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)
    u1 = (h1.double() + 0.5) * two32
    u2 = (h2.double() + 0.5) * two32
    return (like.sqrt(-2.0 * like.log(u1)) * like.cos(2.0 * np.pi * u2)).float()


def _planes(nxl, N):
    step = max(1, min(nxl, (1 << 24) // (N * N) or 1))
    return [(a, min(nxl, a + step)) for a in range(0, nxl, step)]


def lattice_ics(like, N, x0, nxl, seed=42, vel_rms=1e-3, sigma_cells=0.3, device=None):
    """SURVEY 8(d) synthetic ICs of the planes [x0, x0 + nxl): cell-centre lattice ((i + 1/2) / N) in lexicographic
    (i, j, k) order + N(0, sigma_cells / N) displacement per axis, wrapped; velocities N(0, vel_rms); ids = the
    lexicographic lattice index (the reference's particle order).  Returns (pos, vel, ids)."""
    n = nxl * N * N
    if like is np:
        pos, vel = np.empty((n, 3), dtype=np.float32), np.empty((n, 3), dtype=np.float32)
        ids = np.arange(x0 * N * N, (x0 + nxl) * N * N, dtype=np.int64)
    else:
        pos = like.empty((n, 3), dtype=like.float32, device=device)
        vel = like.empty((n, 3), dtype=like.float32, device=device)
        ids = like.arange(x0 * N * N, (x0 + nxl) * N * N, dtype=like.int64, device=device)
    for a, b in _planes(nxl, N):
        sl = slice(a * N * N, b * N * N)
        idx = ids[sl]
        cell = (idx // (N * N), (idx // N) % N, idx % N)
        for d in range(3):
            if like is np:
                c = (cell[d].astype(np.float32) + np.float32(0.5)) / np.float32(N)
                p = c + hash_normal(idx, d, seed, like) * np.float32(sigma_cells / N)
                p = p - np.floor(p)
                p[p >= 1.0] = 0.0
            else:
                c = (cell[d].float() + 0.5) / N
                p = c + hash_normal(idx, d, seed, like) * (sigma_cells / N)
                p = p - like.floor(p)
                p[p >= 1.0] = 0.0
            pos[sl, d] = p
            vel[sl, d] = hash_normal(idx, 3 + d, seed, like) * (np.float32(vel_rms) if like is np else vel_rms)
    return pos, vel, ids


def analytic_velocity(pos, seed, rms=1e-3, nmodes=8, kmax=4):
    """Coherent large-scale flow evaluated at the particle positions: sum of `nmodes` long-wavelength plane waves
    per component (integer wave vectors, |k_i| <= kmax box modes), rms `rms` -- the SECOND labelled particle set of
    the bench (`extra.coherent_flows`): like the flows of real initial conditions it keeps the Morton order of the
    particle arrays coherent over a reorder cycle, which uncorrelated velocities do not."""
    import torch
    rng = np.random.default_rng(seed)
    vel = torch.zeros_like(pos)
    for d in range(3):
        for _ in range(nmodes):
            k = rng.integers(-kmax, kmax + 1, size=3)
            if not k.any():
                k[d] = 1
            ph = float(rng.uniform(0, 2 * np.pi))
            arg = (pos[:, 0] * float(k[0]) + pos[:, 1] * float(k[1]) + pos[:, 2] * float(k[2])) * (2 * np.pi) + ph
            vel[:, d] += torch.sin(arg)
            del arg
    vel *= rms * np.sqrt(2.0 / nmodes)
    return vel


def slab_ics(N, x0, nxl, seed=42, vel_rms=1e-3, velocities="random"):
    """Device ICs of the planes [x0, x0 + nxl) (lattice_ics on the GPU); velocities = "random"  or
    "coherent" (analytic_velocity)."""
    import torch
    pos, vel, ids = lattice_ics(torch, N, x0, nxl, seed=seed, vel_rms=vel_rms, device="cuda")
    if velocities == "coherent":
        vel = analytic_velocity(pos, seed + 1, vel_rms)
    return pos, vel, ids


def clustered_positions(pos, N, seed=7, nblobs=4096, sigma_cells=0.7):
    """The clustered particle set: half of the particles moved into `nblobs` Gaussian blobs of sigma_cells cells
    (bins with ~10^4 particles next to almost empty ones): deposit contention and load imbalance of an evolved field."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = pos.shape[0]
    centres = torch.rand((nblobs, 3), generator=g, device="cuda")
    which = torch.randint(0, nblobs, (n // 2,), generator=g, device="cuda")
    pc = centres[which] + torch.randn((n // 2, 3), generator=g, device="cuda") * (sigma_cells / N)
    out = pos.clone()
    out[n - n // 2:] = pc
    out -= torch.floor(out)
    out[out >= 1.0] = 0.0
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  nvidia-smi needs a
    fraction of a second to start, which is as long as the timed region of a short run: the sampler is started ahead
    of it (before the untimed pre-roll steps, same kernels, same load) and every sample carries its arrival time, so
    that only the samples taken between mark_begin() and mark_end() are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is not None and not any(self.t0 <= t <= self.t1 + 0.06 for t, _ in self.rows):
            time.sleep(0.15)          # let a sample that was in flight arrive
        self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 7]
        inside = [r for t, r in self.rows if len(r) >= 7 and self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        window = "timed region"
        if not inside:                # a timed region shorter than the sampling period: the samples next to it
            inside, window = rows[-3:], "last samples before the end of the timed region (same load: pre-roll steps)"
        sm = [float(r[0]) for r in inside if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu_traffic():
    if os.path.exists(NCU_TRAFFIC_FILE):
        return json.load(open(NCU_TRAFFIC_FILE))
    return {}


# ------------------------------------------------------------------------------ CPU (oracle) arm
def cpu_step_rate(ncoarse, steps, warmup, threads):
    """Times the oracle port (C/OpenMP + numpy FFT, the reference's algorithm) for full leapfrog steps
    at 2^ncoarse cells per side on the host cores, on the ICs of lattice_ics.  Returns (particle-updates/s, ms/step)."""
    import oracle
    from oracle import host
    oracle.set_num_threads(threads)
    N = 2 ** ncoarse
    tables = make_tables()
    param = make_param(ncoarse, threads)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    host.set_units(param)
    pos, vel, _ = lattice_ics(np, N, 0, N)
    pos, vel = oracle.utils.reorder_particles(pos, vel)
    acc, pot, add = host.pm(pos, param)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        param["nsteps"] += 1
        pos, vel, acc, pot, add = host.integrate(pos, vel, acc, pot, add, tables, param, 1e30)
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times[warmup:]))
    return N ** 3 / dt, dt * 1e3


def cpu_baseline_record(nc_sample):
    threads = os.cpu_count() or 1
    v, ms = cpu_step_rate(nc_sample, 1, 1, threads)
    Nc = 2 ** nc_sample
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{Nc}^3 particles / {Nc}^3 mesh full leapfrog step of the same workload shape "
                      f"(1 warm-up + 1 timed, {ms:.0f} ms); numpy-pocketfft FFT (single thread), as the reference "
                      "without pyfftw"}


def run_reference_arm(args):
    """The reference's CPU implementation of the step = the oracle port (the reference itself is Numba + astropy and
    cannot be installed on the GPU box: its build backend pdm-backend is not in the wheelhouse).  The
    full-size step is timed when the K + W steps fit the time budget, else a smaller sample -- `config.workload`
    names the size that actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nc = args.ncoarse
    probe_nc = min(args.cpu_ncoarse, nc)
    value, ms = cpu_step_rate(probe_nc, 1, 1, threads)
    ran_nc, steps_run = probe_nc, 1
    if nc > probe_nc and not args.cpu_sample_only:
        est_step_s = ms * 1e-3 * 8.5 ** (nc - probe_nc)
        est_setup_s = 60.0 * 8 ** (nc - 9)          # ICs (hash generator) + Morton sort on the host at 512^3
        if est_step_s * (args.steps + args.warmup + 1) + est_setup_s <= args.cpu_budget_s:
            value, ms = cpu_step_rate(nc, args.steps, args.warmup, threads)
            ran_nc, steps_run = nc, args.steps
    elif nc == probe_nc:
        value, ms = cpu_step_rate(nc, args.steps, args.warmup, threads)
        steps_run = args.steps
    N, Nw = 2 ** ran_nc, 2 ** nc
    sample = (f"{N}^3 particles / {N}^3 mesh full leapfrog step, {steps_run} timed step(s)"
              + ("" if ran_nc == nc else f" (1/{(Nw // N) ** 3} of the {Nw}^3 workload: the full size does not fit "
                                         f"the {args.cpu_budget_s:.0f} s budget of this arm)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(ran_nc),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "fft": "numpy-pocketfft (single thread), as the reference without pyfftw"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(ncoarse):
    N = 2 ** ncoarse
    return {"workload": f"Newtonian FFT-PM leapfrog step, {N}^3 particles on {N}^3 mesh, TSC, compensated Green, "
                        f"5-pt gradient, n_reorder={N_REORDER} (BASELINE configs[0] shape at {N}^3)",
            "ncells_1d": N, "npart": N ** 3,
            "ics": "SURVEY 8(d): lattice + N(0, 0.3 cell) displacement, velocities N(0, 1e-3) per particle, "
                   "counter-based generator seed 42 (bench.hash_normal: same particles in every arm / per x-slab), "
                   "Morton-ordered",
            "l2_policy": "inputs larger than L2 (particle arrays 3 x %.1f GB, grids %.2f GB vs 126 MB L2)" % (
                12 * N ** 3 / 1e9, 4 * N ** 3 / 1e9)}


# ---------------------------------------------------------------------------------- helpers of the GPU arms
def kernel_table(records, steps, N, world, peak, particle_share=None):
    per = {}
    for name, a, b in records:
        per.setdefault(name, []).append(a.elapsed_time(b))
    kern = {k: {"calls_per_step": len(v) / steps, "ms_per_call": float(np.mean(v)),
                "ms_per_step": float(np.sum(v)) / steps,
                "ms_first_last": [float(np.mean(v[:3])), float(np.mean(v[-3:]))]} for k, v in per.items()}
    for k, d in kern.items():
        if ALGO_BYTES.get(k, 0) > 0 and k not in ("psc_gauss_seidel", "psc_restrict_residual", "psc_prolongation",
                                                  "psc_residual_sumsq"):
            d["algo_bytes"] = ALGO_BYTES[k] * N ** 3 / world
            d["achieved_gbs"] = d["algo_bytes"] / (d["ms_per_call"] * 1e-3) / 1e9
            d["frac_of_peak"] = d["achieved_gbs"] / peak
    return kern


def time_steps(step, nsteps):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nsteps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nsteps


def single_gpu_state(N, param, tables, velocities="random", clustered=False, vel_rms=1e-3, sigma_cells=0.3):
    """Morton-ordered ICs + first force: [pos, vel, acc, pot, add]"""
    import torch
    from pysco_b200 import solver, utils
    pos, vel, _ = lattice_ics(torch, N, 0, N, vel_rms=vel_rms, sigma_cells=sigma_cells, device="cuda")
    if clustered:
        pos = clustered_positions(pos, N)
    if velocities == "coherent":
        vel = analytic_velocity(pos, 43, vel_rms)
    pos, vel = utils.reorder_particles(pos, vel)
    torch.cuda.empty_cache()
    return [pos, vel] + list(solver.pm(pos, param, tables=tables))


def short_run(N, nc, steps, label, velocities="random", clustered=False, **over):
    """`steps` timed steps (after 3 warm-up steps) of a variant of the workload on one GPU; returns a record"""
    import torch
    from pysco_b200 import _lib, integration, utils
    tables = make_tables()
    param = make_param(nc, 1, **over)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    sig = 0.02 if param["theory"] == "fr" else 0.3
    vr = 1e-5 if param["theory"] == "fr" else 1e-3
    state = single_gpu_state(N, param, tables, velocities, clustered, vel_rms=vr, sigma_cells=sig)

    def step():
        param["nsteps"] += 1
        state[:] = integration.integrate(*state, tables, param, 1e30)

    for _ in range(3):
        step()
    _lib.enable_timing(True)
    ms = time_steps(step, steps)
    records = _lib.timing_records()
    _lib.enable_timing(False)
    peak, _ = measured_peak_gbs()
    kern = kernel_table(records, steps, N, 1, peak)
    del state[:]
    torch.cuda.empty_cache()
    return {"what": label, "ncells_1d": N, "steps": steps, "ms_per_step": ms, "value": N ** 3 / (ms * 1e-3),
            "unit": UNIT, "kernels": {k: {"ms_per_step": v["ms_per_step"], "calls_per_step": v["calls_per_step"],
                                          **({"frac_of_peak": v["frac_of_peak"]} if "frac_of_peak" in v else {})}
                                      for k, v in kern.items()}}


def grid_kernel_rooflines(N):
    """Fine-level multigrid / f(R) / MOND / spectral kernels timed alone on an N^3 grid (they run inside a CUDA graph
    in the step, where per-call events cannot be recorded): achieved GB/s from SURVEY 8(d)'s bytes per cell."""
    import torch
    from pysco_b200 import cubic, laplacian, mesh, mond
    peak, _ = measured_peak_gbs()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((N, N, N), generator=g, device="cuda") * 1e-3
    b = torch.randn((N, N, N), generator=g, device="cuda")
    xc = torch.randn((N // 2,) * 3, generator=g, device="cuda") * 1e-3
    u = torch.rand((N, N, N), generator=g, device="cuda") * 0.1 + 0.05
    out = torch.empty_like(x)

    def t(fn, reps=5):
        fn()
        return min(time_steps(fn, 1) for _ in range(reps))
    rows = {
        "gauss_seidel sweep (laplacian.py:844)": (12.0, t(lambda: laplacian.gauss_seidel(x, b, np.float32(1.25)))),
        "residual_error (laplacian.py:327)": (8.0, t(lambda: laplacian.residual_error(x, b))),
        "restrict_residual (laplacian.py:125)": (8.5, t(lambda: laplacian.restrict_residual(x, b))),
        "add_prolongation (mesh.py:334)": (8.5, t(lambda: mesh.add_prolongation(x, xc))),
        "cubic gauss_seidel sweep (cubic.py:269)": (12.0, t(lambda: cubic.gauss_seidel(u, b, np.float32(-0.01), np.float32(1.25)))),
        "mond rhs_simple (mond.py:171)": (8.0, t(lambda: mond.rhs_simple(x, out, np.float32(0.05)))),
    }
    return {k: {"algo_bytes_per_cell": by, "ms": ms, "achieved_gbs": by * N ** 3 / (ms * 1e-3) / 1e9,
                "frac_of_peak": by * N ** 3 / (ms * 1e-3) / 1e9 / peak} for k, (by, ms) in rows.items()}


# ---------------------------------------------------------------------------------- B200 arm, one GPU
def run_gpu_arm(args):
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from pysco_b200 import _lib, integration, utils
    _lib.load()

    nc = args.ncoarse
    N = 2 ** nc
    K = args.steps
    tables = make_tables()
    param = make_param(nc, 1)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    state = single_gpu_state(N, param, tables)

    def step():
        param["nsteps"] += 1
        state[:] = integration.integrate(*state, tables, param, 1e30)
        if param["nsteps"] % N_REORDER == 0:
            state[0], state[1], state[2] = utils.reorder_particles(state[0], state[1], state[2])
            return True
        return False

    for _ in range(args.warmup):
        step()
    # reorder cost (amortised below for the part of a reorder that does not land in the timed steps); the first
    # call warms the allocator
    state[0], state[1], state[2] = utils.reorder_particles(state[0], state[1], state[2])
    torch.cuda.synchronize()
    _lib.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rp, rv, ra = utils.reorder_particles(state[0], state[1], state[2])
    e1.record()
    torch.cuda.synchronize()
    t_reorder_ms = e0.elapsed_time(e1)
    reorder_calls = {}
    for name, a, b in _lib.timing_records():
        reorder_calls[name] = reorder_calls.get(name, 0.0) + a.elapsed_time(b)
    _lib.enable_timing(False)
    state[0], state[1], state[2] = rp, rv, ra
    del rp, rv, ra
    # the timed window sits in the middle of the 50-step reorder cycle
    preroll = max(0, (N_REORDER - K) // 2) if K < N_REORDER else 0
    param["nsteps"] = 0
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    for _ in range(preroll):
        step()

    from pysco_b200 import mesh as _mesh

    def count_hits():   # steps whose sort used the bin counts predicted by the previous interpolation kernel
        return sum(getattr(sb, "counts_skipped", 0) for sb in _mesh._step_sorted.values())
    hits0 = count_hits()
    _lib.enable_timing(True)
    launches0 = _lib.launch_count()
    gc.collect()
    gc.freeze()       # see slab_measure: no full pass of the cyclic collector over the import-time heap in a step
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()   # `ncu --profile-from-start off` lists the timed steps only
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_reorders = 0
    for _ in range(K):
        n_reorders += bool(step())
    ev1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    hits = int(count_hits() - hits0)
    sampler.mark_end()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    records = _lib.timing_records()
    _lib.enable_timing(False)
    t_ms = ev0.elapsed_time(ev1)
    t_ms_total = t_ms + (K / N_REORDER - n_reorders) * t_reorder_ms   # K/50 reorders belong to K steps
    ms_per_step = t_ms_total / K
    value = N ** 3 / (ms_per_step * 1e-3)

    peak, peak_src = measured_peak_gbs()
    kern = kernel_table(records, K, N, 1, peak)
    # what the device spends outside the library's calls: the wait for the host after the 8-byte read of max|a|, max|v|
    # that the next step's dt needs (integration.py:40-60), event overhead
    outside_ms = t_ms / K - sum(v["ms_per_step"] for v in kern.values())
    step_algo_bytes = STEP_ALGO_BYTES * N ** 3
    dom = max((k for k in kern if "algo_bytes" in kern[k]), key=lambda k: kern[k]["ms_per_step"])
    traffic = load_ncu_traffic() if N == 512 else {}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kern[dom]["achieved_gbs"] / peak,
                "traffic": traffic.get(dom), "traffic_source": traffic.get("source"),
                "peak_source": peak_src, "algo_bytes_per_launch": kern[dom]["algo_bytes"],
                "ms_per_launch": kern[dom]["ms_per_call"],
                "whole_step": {"algo_bytes": step_algo_bytes,
                               "achieved": step_algo_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": step_algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak}}

    # ---- e2e: same public call with HOST (pinned) buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in state[:4]]
        host.append(state[4])
        del state[:]
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.e2e_steps, K))
        npart = host[0].shape[0]
        for i in range(1 + e2e_steps):
            if i == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            param["nsteps"] += 1
            host[:] = integration.integrate(*host, tables, param, 1e30)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        # bytes integration._leapfrog_pinned moves per step: x, v, a up; x, v, a and the potential down
        e2e = {"value": N ** 3 / dt, "unit": UNIT, "h2d_bytes_per_step": 36 * npart,
               "d2h_bytes_per_step": 36 * npart + 4 * N ** 3, "steps": e2e_steps, "ms_per_step": dt * 1e3,
               "api": "pysco_b200.integration.integrate(pinned host tensors)"}
        del host
    else:
        del state[:]
    torch.cuda.empty_cache()

    extra = None
    if not args.no_extras:
        ks = args.extra_steps
        extra = {}
        for name, fn in (
            ("coherent_flows", lambda: short_run(N, nc, ks, "same step, velocities = 8 long-wavelength plane waves per "
                                                 "component (rms 1e-3): the particle arrays stay Morton-coherent",
                                                 velocities="coherent")),
            ("clustered", lambda: short_run(N, nc, ks, "same step, half of the particles in 4096 Gaussian blobs of "
                                            "0.7 cell (bins of ~10^4 particles): the lanes-own-particles deposit path",
                                            clustered=True)),
            ("slab_p1", lambda: slab_short_run(N, nc, ks)),
            ("config2_newton_multigrid_256", lambda: short_run(256, 8, ks, "BASELINE config 2: Newtonian 256^3, "
                                                              "multigrid V-cycles", linear_newton_solver="multigrid")),
            ("config3_fr_256", lambda: short_run(256, 8, ks, "BASELINE config 3: f(R) n=1 |fR0|=1e-5 256^3, FAS "
                                                 "multigrid, screened regime a = 0.02 .. 0.07 (cubic.py:196 has "
                                                 "no real branch from a = 0.13 on for these particles and stops "
                                                 "the reference too)", theory="fr",
                                                 fR_logfR0=5, fR_n=1, linear_newton_solver="multigrid")),
            ("config4_mond_512", lambda: short_run(N, nc, ks, "BASELINE config 4: QUMOND 512^3, fft_7pt + MOND source "
                                                   "+ second solve", theory="mond", mond_function="simple", mond_g0=1.2,
                                                   mond_scale_factor_exponent=0, mond_alpha=1,
                                                   linear_newton_solver="fft_7pt")),
            ("grid_kernels_512", lambda: grid_kernel_rooflines(N)),
            ("grid_kernels_256", lambda: grid_kernel_rooflines(256)),
        ):
            try:
                extra[name] = fn()
            except Exception as exc:   # an extra must never cost the headline line
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()

    cpu = None if args.no_cpu_baseline else cpu_baseline_record(args.cpu_ncoarse)
    cfg = workload_config(nc)
    cfg["timed_window"] = (f"steps {preroll + 1}..{preroll + K} of the {N_REORDER}-step reorder cycle"
                           if K < N_REORDER else "whole reorder cycle(s)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "kernels": kern,
        "ms_per_step_outside_calls": outside_ms,
        "predicted_bin_counts_used": {"steps": hits, "of": K,
                                      "what": "steps whose sort skipped its count pass: the previous interpolation "
                                              "kernel had counted the bins under the time step that was then taken"},
        "reorder": {"ms": t_reorder_ms, "in_timed_steps": n_reorders, "amortised_over": N_REORDER,
                    "calls_ms": reorder_calls},
        "extra": extra,
    }
    emit(line)


def slab_short_run(N, nc, steps):
    """The x-slab code path on ONE rank (the P = 1 point of the scaling curve: same kernels, slab bookkeeping)"""
    import torch
    from pysco_b200 import slab, utils
    tables = make_tables()
    param = make_param(nc, 1)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    S = slab.Slab(N, comm=slab.SelfComm(), capacity_factor=1.05)
    pos, vel, ids = slab_ics(N, S.x0, S.nxl)
    S.set_particles(pos, vel, ids)
    del pos, vel, ids
    torch.cuda.empty_cache()
    S.reorder()
    S.pm(param)

    def step():
        param["nsteps"] += 1
        S.integrate(tables, param, 1e30)

    from pysco_b200 import _lib
    for _ in range(3):
        step()
    _lib.enable_timing(True)
    ms = time_steps(step, steps)
    records = _lib.timing_records()
    _lib.enable_timing(False)
    peak, _ = measured_peak_gbs()
    kern = kernel_table(records, steps, N, 1, peak)
    del S
    torch.cuda.empty_cache()
    return {"what": "pysco_b200.slab.Slab.integrate on one rank (P = 1 point of the slab scaling curve)",
            "ncells_1d": N, "steps": steps, "ms_per_step": ms, "value": N ** 3 / (ms * 1e-3), "unit": UNIT,
            "kernels": {k: {"ms_per_step": v["ms_per_step"], "calls_per_step": v["calls_per_step"]}
                        for k, v in kern.items()}}


# ------------------------------------------------------------------- B200 arm, x-slab decomposition
def slab_measure(S, param, tables, K, warmup, world, rank, local_rank, with_phases=True):
    """W warm-up steps, reorder timing, (50 - K) / 2 pre-roll steps, K timed steps of S.integrate.  Returns a dict."""
    import torch
    import torch.distributed as dist
    from pysco_b200 import _lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        param["nsteps"] += 1
        S.integrate(tables, param, 1e30)
        if param["nsteps"] % N_REORDER == 0:
            S.reorder()
            return True
        return False

    N = S.N
    for _ in range(warmup):
        step()
    S.reorder()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.reorder()
    e1.record()
    torch.cuda.synchronize()
    t_reorder_ms = e0.elapsed_time(e1)
    preroll = max(0, (N_REORDER - K) // 2) if K < N_REORDER else 0
    param["nsteps"] = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(preroll):
        step()

    _lib.enable_timing(True)
    S.phase_marks = [] if with_phases else None
    launches0 = _lib.launch_count()
    # Python's cyclic collector runs a full (generation 2) pass about every 24 steps of this loop and takes ~3 ms on
    # a heap with torch, pandas and scipy loaded -- and a rank that pauses stalls all the others at the next exchange
    # (measured: one 7.6 ms step among 4.6 ms ones, tools/trace_slab_steps.py).  Collect now and freeze what exists,
    # so that later passes only look at objects created since (microseconds); main.run / slab.run do the same.
    gc.collect()
    gc.freeze()
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_reorders = 0
    migrated = 0
    for _ in range(K):
        n_reorders += bool(step())
        migrated += S.migrated_last[0]
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    records = _lib.timing_records()
    _lib.enable_timing(False)
    marks, S.phase_marks = (S.phase_marks or []), None
    phase_samples = {}
    for (n0, e0_), (n1, e1_) in zip(marks[:-1], marks[1:]):
        if n1 != "start":   # device time between consecutive phase boundaries (kernels + collectives + bubbles)
            phase_samples.setdefault(n1, []).append(e0_.elapsed_time(e1_))
    phases = {k: {"mean": float(np.mean(v)), "median": float(np.median(v)), "max": float(np.max(v)),
                  "argmax_step": int(np.argmax(v))} for k, v in phase_samples.items()}
    t_ms_total = ev0.elapsed_time(ev1) + (K / N_REORDER - n_reorders) * t_reorder_ms
    tt = torch.tensor([t_ms_total, float(S.np), float(migrated)], device="cuda", dtype=torch.float64)
    mx = tt.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    ms_per_step = mx[0].item() / K
    assert int(round(tt[1].item())) == N ** 3 or world == 1, "particles lost in migration"
    peak, peak_src = measured_peak_gbs()
    kern = kernel_table(records, K, N, world, peak)
    kernel_ms = sum(d["ms_per_step"] for d in kern.values())
    step_algo_bytes = STEP_ALGO_BYTES * N ** 3
    dom = max((k for k in kern if "algo_bytes" in kern[k]), key=lambda k: kern[k]["ms_per_step"])
    agg = step_algo_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "algo_bytes_per_launch": kern[dom]["algo_bytes"], "ms_per_launch": kern[dom]["ms_per_call"],
                "whole_step": {"algo_bytes": step_algo_bytes, "achieved": agg, "aggregate_peak": peak * world,
                               "frac": agg / (peak * world)}}
    return {"ms_per_step": ms_per_step, "value": N ** 3 / (ms_per_step * 1e-3), "roofline": roofline, "kernels": kern,
            "kernel_ms_per_step": kernel_ms, "comm_and_host_ms_per_step": ms_per_step - kernel_ms,
            "phases_ms_per_step_rank0": phases, "gpu_launches": int(launches), "clocks": clocks,
            "reorder": {"ms": t_reorder_ms, "in_timed_steps": n_reorders, "amortised_over": N_REORDER},
            "particles_migrated_per_step": tt[2].item() / K, "preroll": preroll}


def run_slab_arm(args):
    """N GPUs, one x-slab of the mesh (and its particles) per rank: pysco_b200/slab.py.  STRONG scaling: the
    same N^3 problem on every GPU count; with 8 ranks BASELINE config 5 (2048^3) is timed as well (`config5`)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    from pysco_b200 import _lib, distributed, slab, utils
    distributed.init_from_env("nccl")
    _lib.load()
    nc = args.ncoarse
    N = 2 ** nc
    K = args.steps
    tables = make_tables()
    comm = slab.default_comm()

    def build(ncc):
        Nn = 2 ** ncc
        p = make_param(ncc, 1)
        p["t"] = float(tables[1](np.log(p["aexp"])))
        utils.set_units(p)
        # spare rows for arrivals: migration moves O(N^2 / P) particles per step, a few per cent is plenty
        S_ = slab.Slab(Nn, comm=comm, capacity_factor=1.3 if Nn ** 3 / world < 2e8 else 1.1)
        pos, vel, ids = slab_ics(Nn, S_.x0, S_.nxl, seed=42)
        S_.set_particles(pos, vel, ids)
        del pos, vel, ids
        torch.cuda.empty_cache()
        S_.reorder()
        S_.pm(p)
        return S_, p

    S, param = build(nc)
    m = slab_measure(S, param, tables, K, args.warmup, world, rank, local_rank)

    # ---- e2e: every step uploads the rank's x, v, a from pinned host memory and downloads them afterwards
    e2e = None
    if not args.no_e2e and 36 * S.np * 1.2 > 8e9:
        e2e = {"skipped": "pinned host staging of %.1f GB per rank exceeds the 8 GB guard" % (36 * S.np * 1.2 / 1e9)}
    elif not args.no_e2e:
        n = S.np
        hp, hv, ha = (torch.empty((int(n * 1.2) + 1024, 3), dtype=torch.float32, pin_memory=True) for _ in range(3))
        hp[:n].copy_(S.position); hv[:n].copy_(S.velocity); ha[:n].copy_(S.acceleration)
        e2e_steps = max(1, min(args.e2e_steps, K))
        moved_bytes = 0
        for i in range(1 + e2e_steps):
            if i == 1:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                moved_bytes = 0
            n = S.np
            S.pos[:n].copy_(hp[:n], non_blocking=True)
            S.vel[:n].copy_(hv[:n], non_blocking=True)
            S.acc[:n].copy_(ha[:n], non_blocking=True)
            param["nsteps"] += 1
            S.integrate(tables, param, 1e30)
            n2 = S.np
            hp[:n2].copy_(S.position, non_blocking=True)
            hv[:n2].copy_(S.velocity, non_blocking=True)
            ha[:n2].copy_(S.acceleration, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            moved_bytes += 36 * (n + n2)
        dt = (time.perf_counter() - t0) / e2e_steps
        t2 = torch.tensor([dt, moved_bytes / e2e_steps / 2], device="cuda", dtype=torch.float64)
        t2s = t2.clone()
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            dist.all_reduce(t2s, op=dist.ReduceOp.SUM)
        e2e = {"value": N ** 3 / t2[0].item(), "unit": UNIT, "h2d_bytes_per_step": int(t2s[1].item()),
               "d2h_bytes_per_step": int(t2s[1].item()), "steps": e2e_steps, "ms_per_step": t2[0].item() * 1e3,
               "api": "pysco_b200.slab.Slab.integrate with the rank's x, v, a uploaded from / downloaded to pinned "
                      "host memory every step (particle ids stay on the device)"}
        del hp, hv, ha
    peer = bool(S._peer)
    nxl = S.nxl

    line = None
    if rank == 0:
        cfg = workload_config(nc)   # identical to the reference arm's config (same N^3 problem on every GPU count)
        cfg["timed_window"] = (f"steps {m['preroll'] + 1}..{m['preroll'] + K} of the {N_REORDER}-step reorder cycle"
                               if K < N_REORDER else "whole reorder cycle(s)")
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": m["roofline"], "cpu_baseline": None, "e2e": e2e, "gpu_launches": m["gpu_launches"],
            "clocks": m["clocks"], "kernels": m["kernels"],
            "kernel_ms_per_step": m["kernel_ms_per_step"],
            "comm_and_host_ms_per_step": m["comm_and_host_ms_per_step"],
            "phases_ms_per_step_rank0": m["phases_ms_per_step_rank0"], "reorder": m["reorder"],
            "multi_gpu": {"decomposition": f"x-slabs of {nxl} planes per GPU",
                          "fft_transposes": "one kernel each, remote stores over NVLink peer memory + inter-GPU barrier"
                          if peer else "NCCL all-to-all",
                          "collectives_per_step": "1 neighbour exchange of migrants, 2 ghost-plane exchanges (density "
                                                  "add, potential copy), 2 transposes of the half-spectrum, "
                                                  "max of 3 floats over the ranks",
                          "particles_migrated_per_step": m["particles_migrated_per_step"]},
            "config5": None,
        }

    # ---- BASELINE config 5 (the north-star target): 2048^3 on 8 GPUs, a few steps, its own roofline.  It runs AFTER
    # the 512^3 measurement and must never cost that line: a watchdog prints the line without it and ends every rank if
    # the big run does not come back (a rank that failed inside a collective leaves the others waiting).
    config5 = None
    if world == 8 and nc < 11 and not args.no_config5:
        del S
        torch.cuda.empty_cache()
        # 2048^3 on 8 GPUs peaks at ~150 GB per GPU (the Morton reorder's sort buffers on top of 1.07 G particles):
        # attempted only when EVERY rank has the room, so that no rank can run out of memory inside a collective
        free = torch.tensor([torch.cuda.mem_get_info()[0] / 2 ** 30], device="cuda", dtype=torch.float64)
        dist.all_reduce(free, op=dist.ReduceOp.MIN)
        if free.item() < args.config5_min_free_gib:
            config5 = {"skipped": f"least free device memory over the ranks {free.item():.0f} GiB < "
                                  f"{args.config5_min_free_gib:.0f} GiB needed for 2048^3 on 8 GPUs"}
    if config5 is None and world == 8 and nc < 11 and not args.no_config5:
        import threading

        def give_up():
            if line is not None:
                line["config5"] = {"error": f"2048^3 run did not finish within {args.config5_timeout_s:.0f} s"}
                emit(line)
                sys.stdout.flush()
            os._exit(0)

        watchdog = threading.Timer(args.config5_timeout_s, give_up)
        watchdog.daemon = True
        watchdog.start()
        try:
            S5, p5 = build(11)
            m5 = slab_measure(S5, p5, tables, args.config5_steps, 3, world, rank, local_rank)
            config5 = {"config": workload_config(11), "steps": args.config5_steps, "warmup": 3, "scaling": "strong",
                       "target": "north_star: >= 0.60 of the aggregate HBM roofline",
                       **{k: m5[k] for k in ("ms_per_step", "value", "roofline", "kernels", "kernel_ms_per_step",
                                             "comm_and_host_ms_per_step", "phases_ms_per_step_rank0", "reorder")}}
            del S5
        except Exception as exc:
            config5 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if line is not None:      # the other ranks may be waiting in a collective this rank has left
                line["config5"] = config5
                emit(line)
                sys.stdout.flush()
                os._exit(0)
            os._exit(0)
        watchdog.cancel()
        torch.cuda.empty_cache()

    if rank == 0:
        line["config5"] = config5
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at
    stderr for the run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncoarse", type=int, default=9, help="log2 cells per side of the workload (9 -> 512^3)")
    ap.add_argument("--cpu-ncoarse", type=int, default=8, help="log2 cells per side of the CPU sample (8 -> 256^3)")
    ap.add_argument("--cpu-budget-s", type=float, default=270.0,
                    help="reference arm: run the full-size step only if the K + W steps are expected to fit")
    ap.add_argument("--cpu-sample-only", action="store_true", help="reference arm: time the --cpu-ncoarse sample only")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--extra-steps", type=int, default=8)
    ap.add_argument("--config5-steps", type=int, default=6)
    ap.add_argument("--config5-min-free-gib", type=float, default=160.0)
    ap.add_argument("--config5-timeout-s", type=float, default=420.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--decomposition", default="auto", choices=["auto", "slab", "single"],
                    help="x-slabs (default for N > 1) or the single-domain path (N = 1)")
    args = ap.parse_args()
    quiet_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.decomposition == "slab" or (args.decomposition == "auto" and world > 1):
        run_slab_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
