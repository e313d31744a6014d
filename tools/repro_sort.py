#!/usr/bin/env python
"""Repro driver: a few single-domain steps with fast particles (many bin changes per step).
usage: python tools/repro_sort.py [nc=7] [vel_rms=0.05] [steps=3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import integration, solver, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 7
vr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = 2 ** nc
tables = bench.make_tables()
param = bench.make_param(nc, 1)
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)
pos, vel, ids = bench.slab_ics(N, 0, N, seed=7, vel_rms=vr)
acc, pot, add = solver.pm(pos, param, tables=tables)
state = [pos, vel, acc, pot, add]
for s in range(steps):
    param["nsteps"] += 1
    state = list(integration.integrate(*state, tables, param, 1e30))
    torch.cuda.synchronize()
    p = state[0]
    print("step", s, "ok; pos range", float(p.min()), float(p.max()), "ids ok",
          bool((torch.sort(utils.particle_ids(p)).values == torch.arange(N ** 3, device="cuda", dtype=torch.int32)).all()),
          flush=True)
print("done")
