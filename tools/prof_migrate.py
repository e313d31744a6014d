#!/usr/bin/env python
"""Host-side timing of the pieces of Slab.migrate_neighbours (torchrun, >= 2 GPUs)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pysco_b200 import distributed, slab, utils
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
distributed.init_from_env("nccl")
nc = 9; N = 2 ** nc
tables = bench.make_tables(); param = bench.make_param(nc, 1)
param["t"] = float(tables[1](np.log(param["aexp"]))); utils.set_units(param)
comm = slab.default_comm(); S = slab.Slab(N, comm=comm)
pos, vel, ids = bench.slab_ics(N, S.x0, S.nxl); S.set_particles(pos, vel, ids); del pos, vel, ids
S.reorder(); S.pm(param)
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
    return w
S.ops.pack_fixed = timed("pack_fixed", S.ops.pack_fixed)
comm.neighbor_exchange = timed("neighbor_exchange", comm.neighbor_exchange)
S._apply_migration = timed("apply", S._apply_migration)
S.migrate_neighbours = timed("migrate_total", S.migrate_neighbours)
for _ in range(8):
    param["nsteps"] += 1; S.integrate(tables, param, 1e30)
if comm.rank == 0:
    for k, v in T.items(): print(k, [round(x, 3) for x in v[2:]])
    print("cap", S._mig_cap, "migrated", S.migrated_last)
torch.distributed.barrier(); torch.distributed.destroy_process_group()
