#!/usr/bin/env python
"""Wall time of initial_conditions.generate (host white noise + device LPT) at 2^nc particles per side.
usage: python tools/bench_ics.py [nc=8] [order=2LPT]"""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
from pysco_b200 import cosmotable, initial_conditions as ic, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
order = sys.argv[2] if len(sys.argv) > 2 else "2LPT"
N = 2 ** nc
base = tempfile.mkdtemp()
param = pd.Series(cases.ic_param(base, npart=N ** 3, initial_conditions=order))
param["aexp"] = 1.0 / (1 + param["z_start"])
utils.set_units(param)
tables = cosmotable.generate(param)
t0 = time.perf_counter()
d = ic.generate_density_fourier(param)
torch.cuda.synchronize()
t1 = time.perf_counter()
pos, vel = ic.generate(param, tables, write_snapshot=False)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N} {order}: host white noise + transfer grid + upload alone {t1 - t0:.2f} s (first call, cold); whole "
      f"generate() {t2 - t1:.2f} s; particles {tuple(pos.shape)}")
