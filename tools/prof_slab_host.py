#!/usr/bin/env python
"""Host-side profile of the slab step under torchrun: cProfile of rank 0 over a few dozen steps at 2^nc cells per side.
usage: torchrun --nproc-per-node P --master-addr 127.0.0.1 tools/prof_slab_host.py [nc=9] [steps=40]"""
import cProfile
import os
import pstats
import sys
import time

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


def run(k):
    for _ in range(k):
        param["nsteps"] += 1
        S.integrate(tables, param, 1e30)
    torch.cuda.synchronize()


run(10)
torch.distributed.barrier()
t0 = time.perf_counter()
run(steps)
dt = (time.perf_counter() - t0) / steps * 1e3
pr = cProfile.Profile()
pr.enable()
run(steps)
pr.disable()
if rank == 0:
    print(f"N={N} P={S.P}: {dt:.3f} ms/step wall")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(60)
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
