#!/usr/bin/env python
"""Multi-GPU consistency check: two leapfrog steps of the particle-parallel / mesh-replicated path on
WORLD_SIZE ranks must reproduce the single-GPU result on every rank's particle range.
  python tools/check_multigpu.py ref  /tmp/ref.pt         # single process: writes the reference
  torchrun --nproc-per-node 2 tools/check_multigpu.py dist /tmp/ref.pt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import distributed, integration, solver, utils  # noqa: E402

mode, path = sys.argv[1], sys.argv[2]
nc = 7
N = 2 ** nc
if mode == "dist":
    distributed.init_from_env("nccl")
else:
    torch.cuda.set_device(0)
tables = bench.make_tables()
param = bench.make_param(nc, 1)
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)
pos, vel = bench.synthetic_ics_device(N, seed=7)
pos, vel = utils.reorder_particles(pos, vel)
lo, hi = distributed.local_range(pos.shape[0])
pos, vel = pos[lo:hi].clone(), vel[lo:hi].clone()
acc, pot, add = solver.pm(pos, param)
state = [pos, vel, acc, pot, add]
for _ in range(2):
    param["nsteps"] += 1
    state = list(integration.integrate(*state, tables, param, 1e30))
if mode == "ref":
    torch.save({"pos": state[0].cpu(), "vel": state[1].cpu(), "acc": state[2].cpu(), "pot": state[3].cpu(),
                "t": float(param["t"])}, path)
    print("reference written", path)
else:
    ref = torch.load(path, weights_only=False)
    ok = True
    for name, mine, full in (("pos", state[0], ref["pos"][lo:hi]), ("vel", state[1], ref["vel"][lo:hi]),
                             ("acc", state[2], ref["acc"][lo:hi]), ("pot", state[3], ref["pot"])):
        full = full.cuda()
        err = ((mine - full).abs().max() / full.abs().max().clamp_min(1e-30)).item()
        print(f"rank {distributed.rank()} {name}: max rel diff vs single GPU {err:.2e}")
        ok &= err < 1e-4
    assert abs(param["t"] - ref["t"]) < 1e-9 * abs(ref["t"]), "time steps differ"
    assert ok
    print(f"rank {distributed.rank()} OK")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
