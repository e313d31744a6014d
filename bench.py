#!/usr/bin/env python
"""bench.py -- particle-updates/s of one full Newtonian FFT particle-mesh step (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --steps 3 --warmup 1     # CPU arm: the oracle port on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU

Workload: N^3 particles on an N^3 mesh (default N = 512), synthetic ICs = cell-centre
lattice + Gaussian displacement 0.3 cells (seed 42), Morton-ordered as after a reorder, TSC deposit,
compensated-Green FFT solve, 5-point gradient, TSC interpolation, leapfrog KDK through the public API
`pysco_b200.integration.integrate` with analytic background tables; `utils.reorder_particles` every
n_reorder = 50 steps (its cost is measured in the same run and amortised as t_reorder/50 when no
reorder falls inside the K timed steps).

One JSON line on stdout (rank 0).  `value` = device-resident throughput; `e2e` = the same call with
HOST (pinned) buffers, i.e. every step uploads position/velocity/acceleration and downloads the results.
"""
import argparse
import json
import os
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
    "psc_kick_drift_wrap_count": 60.0,   # same pass + the bin counts of the new positions
    "psc_bin_particles_counted": 0.0,
    "psc_deposit": 16.0,           # read x (12) + write rho (4); rescale + RHS affine fused (0)
    "psc_fft_r2c": 8.0,            # read 4 + write 4 (half-spectrum ~ 4 B per real cell)
    "psc_green": 8.0,
    "psc_fft_c2r": 8.0,
    "psc_gradient": 16.0,          # read phi (4) + write force (12)
    "psc_interp_kick4": 60.0,      # read x,v (24) + force once per cell (12) + write v,a (24)
    "psc_interp_kick4_binned": 60.0,
    "psc_interp_kick_phi_binned": 76.0,   # gradient (16) + interpolation/kick (60) in one kernel
    "psc_deposit_binned": 16.0,
    "psc_bin_particles": 0.0,      # pure overhead of the order-independent scheme (not in the 176 B budget)
    # x-slab path (every kernel works on N^3 / P particles or cells)
    "psc_kick_drift_wrap_slab": 60.0,      # kick + drift + wrap + leaver detection
    "psc_bin_particles_slab": 0.0,
    "psc_deposit_binned_slab": 16.0,
    "psc_linear_operator": 8.0,            # RHS affine map (fused into the deposit on the single-domain path)
    "psc_slab_fft_r2c_planes": 8.0,        # 2-D R2C of the owned planes
    "psc_slab_yblocks": 0.0,               # pack / unpack around the all-to-all: overhead
    "psc_slab_fft_x": 8.0,                 # strided 1-D C2C along x on the transposed spectrum
    "psc_green_slab": 8.0,
    "psc_slab_fft_c2r_planes": 8.0,
    "psc_interp_kick_phi_binned_slab": 76.0,
    "psc_slab_count": 0.0, "psc_slab_pack_leavers": 0.0, "psc_slab_unpack_rows": 0.0, "psc_slab_move_rows": 0.0,
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this same workload
# (512^3, one GPU): profiles/r01_final_kernels_ncu.txt
NCU_TRAFFIC_512 = {"psc_interp_kick_phi_binned": 4.900826e9 + 3.384417e9, "psc_deposit_binned": 2.688113e9 + 0.523306e9,
                   "psc_kick_drift_wrap_count": 4.835572e9 + 3.177392e9}

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


def make_param(ncoarse, nthreads):
    import pandas as pd
    N = 2 ** ncoarse
    return pd.Series({
        "nthreads": nthreads, "theory": "newton", "H0": 72, "Om_m": 0.25733, "boxlen": 100, "ncoarse": ncoarse,
        "npart": N ** 3, "integrator": "leapfrog", "mass_scheme": "TSC", "n_reorder": N_REORDER,
        "Courant_factor": 1.0, "max_aexp_stepping": 10, "linear_newton_solver": "fft",
        "gradient_stencil_order": 5, "save_power_spectrum": "no", "aexp": 0.02, "aexp_old": 0.02, "nsteps": 0,
        "write_snapshot": False, "w0": -1.0, "wa": 0.0, "Om_r": 8.0763e-05, "Om_lambda": 1 - 0.25733 - 8.0763e-05,
        "parametrized_mu0": 0.0, "epsrel": 1e-2, "Npre": 2, "Npost": 1,
    })


def synthetic_ics_numpy(N, seed=42):
    import cases
    pos = cases.lattice_particles(N, 0.3, seed=seed)
    vel = cases.velocities(N ** 3, seed=seed + 1, scale=1e-3)
    return pos, vel


def smooth_velocity_field(N, seed, rms=1e-3, corr_cells=16.0):
    """Gaussian random velocity field with N(0, rms) marginals and a correlation length of corr_cells
    cells, sampled at the lattice sites: like the large-scale flows of real N-body initial conditions
    (uncorrelated per-particle velocities would scramble the Morton order within ~10 steps, which no
    cosmological run does)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    k1 = torch.fft.fftfreq(N, device="cuda") * N
    kz = torch.fft.rfftfreq(N, device="cuda") * N
    k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
    filt = torch.exp(-0.5 * k2 * (2 * np.pi * corr_cells / N) ** 2 / (2 * np.pi) ** 2 * 4.0)
    comps = []
    for _ in range(3):
        w = torch.randn((N, N, N), generator=g, device="cuda", dtype=torch.float32)
        f = torch.fft.irfftn(torch.fft.rfftn(w) * filt, s=(N, N, N))
        comps.append((f * (rms / f.std())).reshape(-1))
        del w, f
    return torch.stack(comps, dim=1).contiguous()


def synthetic_ics_device(N, seed=42):
    """Positions as synthetic_ics_numpy (lattice + N(0, 0.3 cell) jitter), generated on the device;
    velocities: smooth Gaussian field with N(0, 1e-3) marginals (see smooth_velocity_field)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    ax = (torch.arange(N, device="cuda", dtype=torch.float32) + 0.5) / N
    pos = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    pos = pos + torch.randn(pos.shape, generator=g, device="cuda", dtype=torch.float32) * (0.3 / N)
    pos = pos - torch.floor(pos)
    pos[pos >= 1.0] = 0.0
    vel = smooth_velocity_field(N, seed + 1)
    return pos.contiguous(), vel


def analytic_velocity(pos, seed, rms=1e-3, nmodes=8, kmax=4):
    """Coherent large-scale flow evaluated at the particle positions: sum of `nmodes` long-wavelength plane waves
    per component (integer wave vectors, |k_i| <= kmax box modes), rms `rms`.  Rank-local (needs no mesh), so the
    slab-decomposed runs generate it per slab; same role as smooth_velocity_field."""
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


def slab_ics(N, x0, nxl, seed=42, vel_rms=1e-3):
    """Synthetic ICs of the planes [x0, x0 + nxl): cell-centre lattice + N(0, 0.3 cell) jitter,
    analytic coherent velocities, ids = lexicographic lattice index (the reference's particle order)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed * 4099 + x0)
    ax = (torch.arange(N, device="cuda", dtype=torch.float32) + 0.5) / N
    pos = torch.stack(torch.meshgrid(ax[x0:x0 + nxl], ax, ax, indexing="ij"), dim=-1).reshape(-1, 3).contiguous()
    pos += torch.randn(pos.shape, generator=g, device="cuda", dtype=torch.float32) * (0.3 / N)
    pos -= torch.floor(pos)
    pos[pos >= 1.0] = 0.0
    ids = torch.arange(x0 * N * N, (x0 + nxl) * N * N, device="cuda", dtype=torch.int64)
    vel = analytic_velocity(pos, seed + 1, vel_rms)
    return pos, vel, ids


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ CPU (oracle) arm
def cpu_step_rate(ncoarse, steps, warmup, threads):
    """Times the oracle port (C/OpenMP + numpy FFT, the reference's algorithm) for full leapfrog steps
    at 2^ncoarse cells per side on the host cores.  Returns (particle-updates/s, ms/step)."""
    import oracle
    from oracle import host
    oracle.set_num_threads(threads)
    N = 2 ** ncoarse
    tables = make_tables()
    param = make_param(ncoarse, threads)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    host.set_units(param)
    pos, vel = synthetic_ics_numpy(N)
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nc = args.cpu_ncoarse
    value, ms = cpu_step_rate(nc, args.steps, args.warmup, threads)
    N = 2 ** nc
    sample = f"{N}^3 particles / {N}^3 mesh full leapfrog step (1/{(2 ** args.ncoarse // N) ** 3} of the {2 ** args.ncoarse}^3 workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.ncoarse),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "fft": "numpy-pocketfft (single thread), as the reference without pyfftw"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(ncoarse):
    N = 2 ** ncoarse
    return {"workload": f"Newtonian FFT-PM leapfrog step, {N}^3 particles on {N}^3 mesh, TSC, compensated Green, "
                        f"5-pt gradient, n_reorder={N_REORDER} (BASELINE configs[0] shape at the metric's {N}^3 size)",
            "ncells_1d": N, "npart": N ** 3, "ics": "lattice + N(0, 0.3 cell) displacement, seed 42 (generated per x-slab), Morton-ordered; velocities: 8 long-wavelength plane waves per component, rms 1e-3 (coherent flows)",
            "l2_policy": "inputs larger than L2 (particle arrays 3 x %.1f GB, grids %.2f GB vs 126 MB L2)" % (
                12 * N ** 3 / 1e9, 4 * N ** 3 / 1e9)}


# ---------------------------------------------------------------------------------- B200 arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)

    import pysco_b200
    from pysco_b200 import _lib, distributed, integration, solver, utils
    distributed.init_from_env("nccl")
    _lib.load()

    nc = args.ncoarse
    N = 2 ** nc
    tables = make_tables()
    param = make_param(nc, 1)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)

    # Multi-GPU: particle-parallel / mesh-replicated (pysco_b200/distributed.py).  The SAME N^3 problem is
    # split by particle index over the ranks (strong scaling); per step one all-reduce(sum) of the density
    # grid and one all-reduce(max) of two floats.
    # same ICs as the slab arm (lattice + jitter, analytic coherent velocities) so that the per-N values compare
    pos, vel, _ = slab_ics(N, 0, N, seed=42)
    pos, vel = utils.reorder_particles(pos, vel)
    if world > 1:
        lo, hi = distributed.local_range(pos.shape[0])
        pos, vel = pos[lo:hi].clone(), vel[lo:hi].clone()
        torch.cuda.empty_cache()
    acc, pot, add = solver.pm(pos, param)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = [pos, vel, acc, pot, add]

    def step():
        param["nsteps"] += 1
        state[:] = integration.integrate(*state, tables, param, 1e30)
        if param["nsteps"] % N_REORDER == 0:
            state[0], state[1], state[2] = utils.reorder_particles(state[0], state[1], state[2])
            return True
        return False

    for _ in range(args.warmup):
        step()
    # reorder cost (amortised below if no reorder lands in the timed steps); first call warms the allocator
    state[0], state[1], state[2] = utils.reorder_particles(state[0], state[1], state[2])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rp, rv, ra = utils.reorder_particles(state[0], state[1], state[2])
    e1.record()
    torch.cuda.synchronize()
    t_reorder_ms = e0.elapsed_time(e1)
    state[0], state[1], state[2] = rp, rv, ra
    del rp, rv, ra

    sampler = ClockSampler(local_rank)
    _lib.enable_timing(True)
    launches0 = _lib.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_reorders = 0
    for _ in range(args.steps):
        n_reorders += bool(step())
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    records = _lib.timing_records()
    _lib.enable_timing(False)
    t_ms = ev0.elapsed_time(ev1)
    # amortised reorder: K/50 reorders belong to K steps
    t_ms_total = t_ms + (args.steps / N_REORDER - n_reorders) * t_reorder_ms
    tt = torch.tensor([t_ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms_total = tt.item()
    ms_per_step = t_ms_total / args.steps
    value = N ** 3 / (ms_per_step * 1e-3)   # whole job: the N^3 particles are shared by all ranks

    # per-kernel device times from the CUDA-event pairs recorded around every C-ABI call
    per = {}
    for name, a, b in records:
        per.setdefault(name, []).append(a.elapsed_time(b))
    kern = {k: {"calls_per_step": len(v) / args.steps, "ms_per_call": float(np.mean(v)),
                "ms_per_step": float(np.sum(v)) / args.steps,
                "ms_first_last": [float(np.mean(v[:3])), float(np.mean(v[-3:]))]} for k, v in per.items()}
    peak, peak_src = measured_peak_gbs()
    for k, d in kern.items():
        if ALGO_BYTES.get(k, 0) > 0:
            # particle kernels see N^3 / world particles per rank, grid kernels the full (replicated) mesh
            units = N ** 3 / world if k in ("psc_kick_drift_wrap", "psc_kick_drift_wrap_count", "psc_interp_kick4", "psc_interp_kick4_binned") else N ** 3
            if k == "psc_interp_kick_phi_binned":
                units = N ** 3 * (60.0 / world + 16.0) / 76.0
            if k in ("psc_deposit", "psc_deposit_binned"):
                units = N ** 3 * (12.0 / world + 4.0) / 16.0
            d["algo_bytes"] = ALGO_BYTES[k] * units
            d["achieved_gbs"] = d["algo_bytes"] / (d["ms_per_call"] * 1e-3) / 1e9
            d["frac_of_peak"] = d["achieved_gbs"] / peak
    step_algo_bytes = STEP_ALGO_BYTES * N ** 3
    dom = max((k for k in kern if ALGO_BYTES.get(k, 0) > 0), key=lambda k: kern[k]["ms_per_step"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kern[dom]["achieved_gbs"] / peak,
                "traffic": NCU_TRAFFIC_512.get(dom) if (N == 512 and world == 1) else None,
                "traffic_source": "profiles/r01_final_kernels_ncu.txt (ncu --set full, bytes per launch)",
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
        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        h2d = sum(t.numel() * 4 for t in host[:4])
        d2h = h2d
        for i in range(1 + e2e_steps):
            if i == 1:
                barrier()
                t0 = time.perf_counter()
            param["nsteps"] += 1
            host[:] = integration.integrate(*host, tables, param, 1e30)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": N ** 3 / tt.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": tt.item() * 1e3,
               "api": "pysco_b200.integration.integrate(pinned host tensors)"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, ms = cpu_step_rate(args.cpu_ncoarse, 1, 1, threads)
            Nc = 2 ** args.cpu_ncoarse
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{Nc}^3 particles / {Nc}^3 mesh full leapfrog step of the same workload shape "
                             f"(1 warm-up + 1 timed, {ms:.0f} ms); numpy-pocketfft FFT"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(nc),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "kernels": kern,
            "reorder": {"ms": t_reorder_ms, "in_timed_steps": n_reorders, "amortised_over": N_REORDER},
            "multi_gpu": ("particle-parallel, mesh-replicated: all-reduce(sum) of the %d^3 density grid + all-reduce(max) of "
                          "2 floats per step (NCCL)" % N) if world > 1 else None,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------- B200 arm, x-slab decomposition
def run_slab_arm(args):
    """N GPUs, one x-slab of the mesh (and its particles) per rank: pysco_b200/slab.py.  STRONG scaling: the
    same N^3 problem on every GPU count (BASELINE quotes the metric at 512^3 on 1/2/4/8 GPUs and 2048^3 on 8:
    --ncoarse 11)."""
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
    tables = make_tables()
    param = make_param(nc, 1)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    comm = slab.default_comm()
    # spare rows for arrivals: migration moves O(N^2 / P) particles per step, a few per cent is plenty
    S = slab.Slab(N, comm=comm, capacity_factor=1.3 if N ** 3 / world < 2e8 else 1.1)
    pos, vel, ids = slab_ics(N, S.x0, S.nxl, seed=42)
    S.set_particles(pos, vel, ids)
    del pos, vel, ids
    torch.cuda.empty_cache()
    S.reorder()
    S.pm(param)

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

    for _ in range(args.warmup):
        step()
    S.reorder()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.reorder()
    e1.record()
    torch.cuda.synchronize()
    t_reorder_ms = e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank)
    _lib.enable_timing(True)
    S.phase_marks = []
    launches0 = _lib.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_reorders = 0
    migrated = 0
    for _ in range(args.steps):
        n_reorders += bool(step())
        migrated += S.migrated_last[0]
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    records = _lib.timing_records()
    _lib.enable_timing(False)
    marks, S.phase_marks = S.phase_marks, None
    phase_samples = {}
    for (n0, e0_), (n1, e1_) in zip(marks[:-1], marks[1:]):
        if n1 != "start":   # device time between consecutive phase boundaries (kernels + collectives + bubbles)
            phase_samples.setdefault(n1, []).append(e0_.elapsed_time(e1_))
    phases = {k: {"mean": float(np.mean(v)), "median": float(np.median(v)), "max": float(np.max(v)),
                  "argmax_step": int(np.argmax(v))} for k, v in phase_samples.items()}
    t_ms_total = ev0.elapsed_time(ev1) + (args.steps / N_REORDER - n_reorders) * t_reorder_ms
    tt = torch.tensor([t_ms_total, float(S.np), float(migrated)], device="cuda", dtype=torch.float64)
    mx = tt.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    ms_per_step = mx[0].item() / args.steps
    assert int(round(tt[1].item())) == N ** 3 or world == 1, "particles lost in migration"
    value = N ** 3 / (ms_per_step * 1e-3)

    per = {}
    for name, a, b in records:
        per.setdefault(name, []).append(a.elapsed_time(b))
    kern = {k: {"calls_per_step": len(v) / args.steps, "ms_per_call": float(np.mean(v)),
                "ms_per_step": float(np.sum(v)) / args.steps} for k, v in per.items()}
    peak, peak_src = measured_peak_gbs()
    for k, d in kern.items():
        if ALGO_BYTES.get(k, 0) > 0:
            d["algo_bytes"] = ALGO_BYTES[k] * N ** 3 / world
            d["achieved_gbs"] = d["algo_bytes"] / (d["ms_per_call"] * 1e-3) / 1e9
            d["frac_of_peak"] = d["achieved_gbs"] / peak
    kernel_ms = sum(d["ms_per_step"] for d in kern.values())
    step_algo_bytes = STEP_ALGO_BYTES * N ** 3
    dom = max((k for k in kern if ALGO_BYTES.get(k, 0) > 0), key=lambda k: kern[k]["ms_per_step"])
    agg = step_algo_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "algo_bytes_per_launch": kern[dom]["algo_bytes"], "ms_per_launch": kern[dom]["ms_per_call"],
                "whole_step": {"algo_bytes": step_algo_bytes, "achieved": agg, "aggregate_peak": peak * world,
                               "frac": agg / (peak * world)}}

    # ---- e2e: every step uploads the rank's x, v, a from pinned host memory and downloads them afterwards
    e2e = None
    if not args.no_e2e and 36 * S.np * 1.2 > 8e9:
        e2e = {"skipped": "pinned host staging of %.1f GB per rank exceeds the 8 GB guard" % (36 * S.np * 1.2 / 1e9)}
    elif not args.no_e2e:
        n = S.np
        hp, hv, ha = (torch.empty((int(n * 1.2) + 1024, 3), dtype=torch.float32, pin_memory=True) for _ in range(3))
        hp[:n].copy_(S.position); hv[:n].copy_(S.velocity); ha[:n].copy_(S.acceleration)
        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        moved_bytes = 0
        for i in range(1 + e2e_steps):
            if i == 1:
                barrier()
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

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, ms = cpu_step_rate(args.cpu_ncoarse, 1, 1, threads)
            Nc = 2 ** args.cpu_ncoarse
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{Nc}^3 particles / {Nc}^3 mesh full leapfrog step of the same workload shape "
                             f"(1 warm-up + 1 timed, {ms:.0f} ms); numpy-pocketfft FFT"}
        cfg = workload_config(nc)   # identical to the reference arm's config (same N^3 problem on every GPU count)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "kernels": kern,
            "kernel_ms_per_step": kernel_ms, "comm_and_host_ms_per_step": ms_per_step - kernel_ms,
            "phases_ms_per_step_rank0": phases,
            "reorder": {"ms": t_reorder_ms, "in_timed_steps": n_reorders, "amortised_over": N_REORDER},
            "multi_gpu": {"decomposition": f"x-slabs of {S.nxl} planes per GPU",
                          "fft_transposes": "one kernel each, remote stores over NVLink peer memory + inter-GPU barrier"
                          if S._peer else "NCCL all-to-all",
                          "collectives_per_step": "1 neighbour exchange of migrants, 2 ghost-plane exchanges (density "
                                                  "add, potential copy), 2 transposes of the half-spectrum, "
                                                  "all-reduce(max) of 3 floats",
                          "particles_migrated_per_step": tt[2].item() / args.steps},
        }
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
    ap.add_argument("--steps", type=int, default=50, help="default = one full n_reorder cycle")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncoarse", type=int, default=9, help="log2 cells per side of the workload (9 -> 512^3)")
    ap.add_argument("--cpu-ncoarse", type=int, default=8, help="log2 cells per side of the CPU sample (8 -> 256^3)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decomposition", default="auto", choices=["auto", "slab", "replicated"],
                    help="multi-GPU layout: x-slabs (default for N > 1) or particle-parallel / mesh-replicated")
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
