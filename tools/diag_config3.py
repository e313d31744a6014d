#!/usr/bin/env python
"""Diagnostic: BASELINE config 3 (f(R), 256^3 by default) step by step, printing max|a|, NaN counts of every array and
dt; variants through the environment (PSC_NO_FUSED_GS=1, PSC_ORDER=reference).  usage: diag_config3.py [nc=8] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import integration, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
N = 2 ** nc
tables = bench.make_tables()
param = bench.make_param(nc, 1, theory="fr", fR_logfR0=5, fR_n=1, linear_newton_solver="multigrid", aexp=0.05,
                         aexp_old=0.05)
if os.environ.get("PSC_ORDER"):
    param["particle_order"] = os.environ["PSC_ORDER"]
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)
state = bench.single_gpu_state(N, param, tables, vel_rms=1e-5, sigma_cells=0.02)


def nans(t):
    return int(torch.isnan(t).sum()) if isinstance(t, torch.Tensor) and t.numel() else 0


print("variant", {k: v for k, v in os.environ.items() if k.startswith("PSC_")}, flush=True)
print("init max|a|", float(state[2].abs().max()), "nan", [nans(t) for t in state], flush=True)
for s in range(steps):
    param["nsteps"] += 1
    try:
        state = list(integration.integrate(*state, tables, param, 1e30))
    except Exception as exc:
        print("step", s, "raised", type(exc).__name__, exc, flush=True)
        break
    add = state[4]
    print("step", s, "aexp %.5f" % param["aexp"], "max|a| %.4e" % float(state[2].abs().max()),
          "nan", [nans(t) for t in state],
          "scalaron min/max %.4e %.4e" % (float(add.min()), float(add.max())) if isinstance(add, torch.Tensor) and add.numel() else "",
          flush=True)
