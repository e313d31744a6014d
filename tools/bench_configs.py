#!/usr/bin/env python
"""Step time and per-kernel breakdown of the other BASELINE configs on one GPU (not the contract bench):
  2: Newtonian 256^3, multigrid     3: f(R) n=1 |fR0|=1e-5 256^3, FAS multigrid     4: QUMOND 512^3 (fft_7pt)
usage: python tools/bench_configs.py [2|3|4] [steps=10]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import _lib, integration, solver, utils  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
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
pos, vel, _ = bench.slab_ics(N, 0, N, vel_rms=1e-3 if cfg != 3 else 1e-5)
if cfg == 3:
    # nearly linear density field: with the 0.3-cell jitter of the PM bench the reference's cubic root formula has no
    # real branch after a few steps (cubic.py:196-197) and the run stops with a math domain error, as in the reference
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    pos = torch.from_numpy(cases.lattice_particles(N, 0.02, seed=5)).cuda()
pos, vel = utils.reorder_particles(pos, vel)
state = [pos, vel] + list(solver.pm(pos, param, tables=tables))
for _ in range(3):
    param["nsteps"] += 1
    state = list(integration.integrate(*state, tables, param, 1e30))
_lib.enable_timing(True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    param["nsteps"] += 1
    state = list(integration.integrate(*state, tables, param, 1e30))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
per = {}
for name, a, b in _lib.timing_records():
    per.setdefault(name, []).append(a.elapsed_time(b))
print(f"config {cfg}: {N}^3, {ms:.3f} ms/step, {N ** 3 / ms / 1e6:.2f} G particle-updates/s")
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"   {k:32s} {len(v) / steps:6.1f} calls/step {sum(v) / steps:8.3f} ms/step")
