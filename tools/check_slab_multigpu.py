#!/usr/bin/env python
"""Multi-GPU consistency check of the x-slab path over NCCL: three leapfrog steps on WORLD_SIZE ranks must
reproduce the single-domain CUDA path (computed by rank 0 before the process group exists).
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_slab_multigpu.py [ncoarse=7] [solver=fft] [theory=newton]
solver: fft | fft_7pt | multigrid; theory: newton | mond | fr (the slab multigrid / QUMOND / FAS paths over NCCL).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import distributed, integration, slab, solver, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 7
SOLVER = sys.argv[2] if len(sys.argv) > 2 else "fft"
THEORY = sys.argv[3] if len(sys.argv) > 3 else "newton"
N = 2 ** nc
NSTEPS = 3
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
tables = bench.make_tables()


def fresh_param():
    param = bench.make_param(nc, 1)
    param["linear_newton_solver"], param["theory"] = SOLVER, THEORY
    for key, value in dict(mond_function="simple", mond_g0=1.2, mond_scale_factor_exponent=0, mond_alpha=1,
                           fR_logfR0=5, fR_n=1).items():
        param[key] = value
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    return param


ref = None
if rank == 0:
    param = fresh_param()
    pos, vel, ids = bench.slab_ics(N, 0, N, seed=7, vel_rms=0.05)
    acc, pot, add = solver.pm(pos, param, tables=tables)
    state = [pos, vel, acc, pot, add]
    for _ in range(NSTEPS):
        param["nsteps"] += 1
        state = list(integration.integrate(*state, tables, param, 1e30))
    order = torch.argsort(ids)
    state[:3] = utils.reference_order(*state[:3])      # the device-resident loop keeps its arrays in bin order
    ref = [state[0][order].cpu(), state[1][order].cpu(), state[2][order].cpu(), float(param["t"])]
    del state, pos, vel, acc, pot, add
    torch.cuda.empty_cache()

distributed.init_from_env("nccl")
comm = slab.default_comm()
param = fresh_param()
s = slab.Slab(N, comm=comm)
# every rank generates the same global ICs and adopts a strided 1/P of them: set_particles routes them home
pos, vel, ids = bench.slab_ics(N, 0, N, seed=7, vel_rms=0.05)
s.set_particles(pos[rank::world].contiguous(), vel[rank::world].contiguous(), ids[rank::world].contiguous())
del pos, vel, ids
s.pm(param, tables=tables)
moved = 0
for step in range(NSTEPS):
    param["nsteps"] += 1
    if step == 1:
        s.reorder()
    s.integrate(tables, param, 1e30)
    moved += s.migrated_last[0]
print(f"rank {rank}: np = {s.np}, migrated out over {NSTEPS} steps = {moved}", flush=True)
res = s.gather_to_root(N ** 3)
if rank == 0:
    ok = True
    for name, mine, full in zip(("pos", "vel", "acc"), res, ref[:3]):
        d = (mine - full).abs()
        if name == "pos":
            d = torch.minimum(d, 1 - d)
        err = (d.max() / full.abs().max().clamp_min(1e-30)).item()
        print(f"{name}: max rel diff vs single-domain path {err:.2e}")
        ok &= err < 1e-4
    assert abs(param["t"] - ref[3]) < 1e-9 * abs(ref[3]), "time steps differ"
    assert ok
    print(f"SLAB MULTI-GPU OK (P = {world}, N = {N}, solver = {SOLVER}, theory = {THEORY})")
torch.distributed.barrier()
torch.distributed.destroy_process_group()
