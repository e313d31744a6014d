#!/usr/bin/env python
"""Stress of the sort after a speculative count that the caller voids (tests/test_zz_sort_gpu.py::
test_predicted_bin_count_equals_the_count_pass, last part), repeated.  usage: stress_sort2.py [reps=40]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases  # noqa: E402
import pysco_b200 as psc  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N, n = 128, 500009
nb = N // 8


def key_of(p):
    c = np.minimum((p * N).astype(np.int64), N - 1) >> 3
    return (c[:, 0] * nb + c[:, 1]) * nb + c[:, 2]


phi = torch.from_numpy(cases.scalar_grid(N, seed=43, smooth=True) * np.float32(2e-3)).cuda()
bad = 0
for rep in range(reps):
    pos, vel = cases.particles(N, n, seed=41 + rep), cases.velocities(n, seed=42 + rep, scale=5e-3)
    sb = psc.mesh.step_sorted(n, N)
    zero = torch.zeros((n, 3), device="cuda")
    p, v, i = psc.mesh.step_sort(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), zero, None,
                                 np.float32(0), np.float32(0), 0, sb)
    dt2 = np.float32(3.0)
    half2 = np.float32(0.5 * dt2)
    for s in range(4):
        sb.predict_next = (half2, dt2, 0)
        acc, _ = psc.mesh.interp_kick_phi(phi, None, 0.0, 0, 5, p, v, 2, np.float32(0.01), sb)
        if s % 2:
            v += 1e-4            # voids the guess
        p2, v2, i2 = psc.mesh.step_sort(p, v, acc, i, half2, dt2, 0, sb)
        torch.cuda.synchronize()
        k = key_of(p2.cpu().numpy())
        ids = i2.cpu().numpy()
        ok = bool(np.all(np.diff(k) >= 0)) and bool(np.array_equal(np.sort(ids), np.arange(n)))
        if not ok:
            bad += 1
            dup = n - len(np.unique(ids))
            print(f"rep {rep} sort {s} ({'voided' if s % 2 else 'predicted'}): rows out of order "
                  f"{int(np.sum(np.diff(k) < 0))}, duplicate ids {dup}, skipped so far {sb.counts_skipped}", flush=True)
        p, v, i = p2, v2, i2
print("failures:", bad, "of", reps * 4)
