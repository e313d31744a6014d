#!/usr/bin/env python
"""Host-side cost of one step: the PM step at a mesh so small that the GPU time is negligible, under cProfile.
usage: python tools/prof_host.py [nc=6] [steps=300]"""
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
from pysco_b200 import integration, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
N = 2 ** nc
tables = bench.make_tables()
param = bench.make_param(nc, 1)
param["t"] = float(tables[1](np.log(param["aexp"])))
utils.set_units(param)
state = bench.single_gpu_state(N, param, tables)


def run(k):
    global state
    for _ in range(k):
        param["nsteps"] += 1
        state = list(integration.integrate(*state, tables, param, 1e30))
    torch.cuda.synchronize()


run(20)
t0 = time.perf_counter()
run(steps)
print(f"N={N}: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step wall (host-bound)")
pr = cProfile.Profile()
pr.enable()
run(steps)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
