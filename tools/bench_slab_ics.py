#!/usr/bin/env python
"""Host cost of one rank's white-noise block (initial_conditions.white_noise_fourier_block) against the full draw
every rank of the replicated generator makes.  usage: bench_slab_ics.py N P"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysco_b200 import initial_conditions as ic  # noqa: E402

N, P = int(sys.argv[1]), int(sys.argv[2])
nyl = N // P
t = time.perf_counter()
b = ic.white_noise_fourier_block(N, 42, nyl, nyl)
t1 = time.perf_counter() - t
t = time.perf_counter()
full = ic.white_noise_fourier(N, np.random.default_rng(42))
t2 = time.perf_counter() - t
same = np.array_equal(b.view(np.float32), np.ascontiguousarray(full[:, nyl:2 * nyl]).view(np.float32))
print(f"N {N} P {P}: block of rank 1 {t1:.2f} s ({b.nbytes / 1e6:.0f} MB), full draw {t2:.2f} s "
      f"({full.nbytes / 1e6:.0f} MB), identical bits: {same}")
