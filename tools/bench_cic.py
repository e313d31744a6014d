#!/usr/bin/env python
"""deposit / gradient + interpolation of the bin kernels with mass_scheme = CIC against TSC at 2^nc cells per side
(bin-ordered arrays).  usage: python tools/bench_cic.py [nc=9]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import mesh, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
N = 2 ** nc
pos, vel, _ = bench.slab_ics(N, 0, N)
pos, vel = utils.reorder_particles(pos, vel)
sb = mesh.step_sorted(pos.shape[0], N)
z = torch.zeros_like(pos)
sp, sv, sid = mesh.step_sort(pos, vel, z, None, np.float32(0), np.float32(0), 0, sb)
sp, sv, sid = mesh.step_sort(sp, sv, z, sid, np.float32(0), np.float32(0), 0, sb)     # micro-block order
phi = torch.randn((N, N, N), device="cuda")


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for name, sid_ in (("TSC", 2), ("CIC", 1)):
    td = timeit(lambda: mesh.deposit_rhs(sp, N, sid_, 1.0, 1.0, 0.0, sb))
    ti = timeit(lambda: mesh.interp_kick_phi(phi, None, 0.0, 0, 5, sp, sv, sid_, 0.0, sb))
    print(f"N={N} {name}: deposit {td:.3f} ms | gradient + interpolation + kick {ti:.3f} ms", flush=True)
