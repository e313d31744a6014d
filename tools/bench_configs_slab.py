#!/usr/bin/env python
"""Step time of the other BASELINE configs on x-slabs (not the contract bench): strong scaling over the ranks of one node.
  2: Newtonian 256^3, multigrid     3: f(R) n=1 |fR0|=1e-5 256^3, FAS multigrid     4: QUMOND 512^3 (fft_7pt)
usage: torchrun --nproc-per-node P --master-addr 127.0.0.1 tools/bench_configs_slab.py [2|3|4] [steps=10]
       (python tools/bench_configs_slab.py ... runs one rank)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import distributed, slab, utils  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    distributed.init_from_env("nccl")
nc = 9 if cfg == 4 else 8
N = 2 ** nc
tables = bench.make_tables()
param = bench.make_param(nc, 1)
if cfg == 2:
    param["linear_newton_solver"] = "multigrid"
elif cfg == 3:
    param["theory"], param["fR_logfR0"], param["fR_n"] = "fr", 5, 1
    param["linear_newton_solver"] = "multigrid"
    param["aexp"] = param["aexp_old"] = 0.05
else:
    param["theory"], param["mond_function"], param["mond_g0"] = "mond", "simple", 1.2
    param["mond_scale_factor_exponent"], param["mond_alpha"] = 0, 1
    param["linear_newton_solver"] = "fft_7pt"
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)

S = slab.Slab(N, comm=slab.default_comm())
nxl = N // world
if cfg == 3:
    # nearly linear density field (see tools/bench_configs.py): every rank builds the global lattice and keeps its slab
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    pos = torch.from_numpy(cases.lattice_particles(N, 0.02, seed=5))
    keep = ((pos[:, 0] * N).long() // nxl) == rank
    ids = torch.nonzero(keep).squeeze(1)
    pos = pos[keep].cuda()
    vel = torch.zeros_like(pos)
    ids = ids.cuda()
else:
    pos, vel, ids = bench.slab_ics(N, rank * nxl, nxl, vel_rms=1e-3)
S.set_particles(pos, vel, ids)
del pos, vel, ids
S.pm(param, tables=tables)
for _ in range(3):
    param["nsteps"] += 1
    S.integrate(tables, param, 1e30)
if world > 1:
    torch.distributed.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    param["nsteps"] += 1
    S.integrate(tables, param, 1e30)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
    ms = ms.item()
    print(f"config {cfg} on {world} slab(s): {N}^3, {ms:.3f} ms/step (max over ranks), "
          f"{N ** 3 / ms / 1e6:.2f} G particle-updates/s")
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
