#!/usr/bin/env python
"""Stress of psc_step_sort's local path: repeated sorts with large drifts (many bin changes, some beyond the
neighbouring bin), checking sortedness, the id set and the bin table every time.  usage: stress_sort.py [reps=40]"""
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
nbins = (N // 8) ** 3
lib, L = psc._lib, psc._lib.load()


def key_of(p):
    c = np.minimum((p * N).astype(np.int64), N - 1) >> 3
    return (c[:, 0] * (N // 8) + c[:, 1]) * (N // 8) + c[:, 2]


bad = 0
for rep in range(reps):
    rng = np.random.default_rng(rep)
    pos = cases.particles(N, n, seed=100 + rep)
    vel = cases.velocities(n, seed=200 + rep, scale=5e-3)
    sb = psc.mesh.step_sorted(n, N)
    zero = torch.zeros((n, 3), device="cuda")
    p, v, i = psc.mesh.step_sort(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), zero, None,
                                 np.float32(0), np.float32(0), 0, sb)
    for s, dt in enumerate((np.float32(0.5), np.float32(3.0), np.float32(3.0), np.float32(6.0))):
        acc = torch.randn((n, 3), device="cuda") * 1e-4
        before = key_of(p.cpu().numpy())
        rp, rv = p.clone(), v.clone()
        lib.check(L.psc_kick_drift_wrap(lib.ptr(rp), lib.ptr(rv), lib.ptr(acc), n, float(np.float32(0.5 * dt)), float(dt), 0, lib.stream()))
        p2, v2, i2 = psc.mesh.step_sort(p, v, acc, i, np.float32(0.5 * dt), dt, 0, sb)
        torch.cuda.synchronize()
        k = key_of(p2.cpu().numpy())
        ok_sorted = bool(np.all(np.diff(k) >= 0))
        ids = i2.cpu().numpy()
        ok_ids = bool(np.array_equal(np.sort(ids), np.arange(n)))
        order = np.argsort(i.cpu().numpy())[ids]
        ok_bits = bool(np.array_equal(p2.cpu().numpy(), rp.cpu().numpy()[order]))
        if not (ok_sorted and ok_ids and ok_bits):
            bad += 1
            after = key_of(rp.cpu().numpy())
            nb = N // 8
            d = np.abs((after // (nb * nb)) - (before // (nb * nb)))
            far = int(np.sum(np.minimum(d, nb - d) > 1))
            print(f"rep {rep} sort {s} dt {dt}: sorted {ok_sorted} ids {ok_ids} bits {ok_bits}; rows out of order "
                  f"{int(np.sum(np.diff(k) < 0))}; far movers along x {far}; table {sb.table}", flush=True)
        p, v, i = p2, v2, i2
print("failures:", bad, "of", reps * 4)
