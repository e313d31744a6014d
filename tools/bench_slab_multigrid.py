"""Time the slab-decomposed multigrid step (one rank, SelfComm) against the single-domain multigrid step on one GPU.

    python tools/bench_slab_multigrid.py [N] [steps]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import cases  # noqa: E402
from pysco_b200 import _lib, integration, slab, solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
tables = cases.toy_tables()
pos = cases.lattice_particles(N, 0.3, seed=11)
vel = cases.velocities(N ** 3, seed=12, scale=0.05)


def param0():
    param = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver="multigrid")
    param["aexp"] = 0.2
    param["t"] = float(tables[1](np.log(param["aexp"])))
    from pysco_b200 import utils
    utils.set_units(param)
    return param


def timed(fn, n):
    torch.cuda.synchronize()
    l0, t0 = _lib.launch_count(), time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, (_lib.launch_count() - l0) / n


# slab, one rank
param = param0()
s = slab.Slab(N, comm=slab.SelfComm())
s.set_particles(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(),
                torch.arange(N ** 3, dtype=torch.int64, device="cuda"))
s.pm(param, tables=tables)


def slab_step():
    param["nsteps"] += 1
    s.integrate(tables, param, 1e30)


slab_step()
ms, launches = timed(slab_step, STEPS)
print(f"slab multigrid  N={N}: {ms:.2f} ms/step, {launches:.0f} kernel launches/step")

# single domain (CUDA-graph V-cycle)
param = param0()
p, v = torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda()
state = [p, v, *solver.pm(p, param)]


def dom_step():
    global state
    param["nsteps"] += 1
    state = list(integration.integrate(*state, tables, param, 1e30))


dom_step()
ms, launches = timed(dom_step, STEPS)
print(f"single domain   N={N}: {ms:.2f} ms/step, {launches:.0f} kernel launches/step")
