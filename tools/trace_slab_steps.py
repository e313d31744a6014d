#!/usr/bin/env python
"""Per-step wall time of the slab step with the quantities that can trigger a reallocation (particle count, migration
capacity, scratch sizes), every rank.  usage: torchrun --nproc-per-node P tools/trace_slab_steps.py [nc=9] [steps=40]"""
import os
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", os.environ.get("TRACE_LOADING", "EAGER"))

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import distributed, slab, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = 2 ** nc
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
distributed.init_from_env("nccl")
tables = bench.make_tables()
param = bench.make_param(nc, 1)
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)
S = slab.Slab(N, comm=slab.default_comm(), capacity_factor=1.15)
pos, vel, ids = bench.slab_ics(N, S.x0, S.nxl)
S.set_particles(pos, vel, ids)
del pos, vel, ids
S.reorder()
S.pm(param)
torch.cuda.synchronize()
for s in range(steps):
    param["nsteps"] += 1
    stats0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    S.integrate(tables, param, 1e30)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    st = torch.cuda.memory_stats()
    print(f"rank {rank} step {s:3d} {dt:8.3f} ms np {S.np} mig_cap {S._mig_cap} want {S._mig_want} "
          f"sorted {S.ops._sorted.numel() if S.ops._sorted is not None else 0} table {S.ops._table} "
          f"cudaMalloc {st['num_device_alloc'] - stats0['num_device_alloc']} "
          f"free {st['num_device_free'] - stats0['num_device_free']} mbox {S.comm._mbox[2] if S.comm._mbox else None}",
          flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
