#!/usr/bin/env python
"""Tiny driver for ncu captures: a few full PM steps (or only the deposit) at 2^nc cells per side.
usage: python tools/prof_step.py [nc=9] [what=step|step+reorder|deposit|interp]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import _lib, integration, mesh, solver, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
what = sys.argv[2] if len(sys.argv) > 2 else "step"
N = 2 ** nc
pos, vel, _ = bench.slab_ics(N, 0, N)
pos, vel = utils.reorder_particles(pos, vel)
if what == "deposit":
    for _ in range(3):
        rho = mesh.TSC(pos, N)
elif what == "interp":
    force = torch.randn((N, N, N, 3), device="cuda")
    for _ in range(3):
        mesh.interp_kick(force, pos, vel, 2, 0.01)
else:
    tables = bench.make_tables()
    param = bench.make_param(nc, 1)
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    state = list(solver.pm(pos, param))
    state = [pos, vel] + state
    for it in range(3):
        param["nsteps"] += 1
        state = list(integration.integrate(*state, tables, param, 1e30))
        if it == 1 and what == "step+reorder":    # utils.reorder_particles on the bin-ordered arrays: the id relabel
            state[:3] = utils.reorder_particles(*state[:3])
torch.cuda.synchronize()
print("done", what, N)
