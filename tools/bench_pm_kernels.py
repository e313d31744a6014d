#!/usr/bin/env python
"""Timing of the binned particle<->mesh kernels alone (binning, deposit, gradient+interpolation+kick) on one GPU, for
a freshly Morton-sorted order, an order that has drifted by one cell rms, a Poisson (uniform random) set and a
clustered set; plus the two halves of the binning inside a step (count fused into kick+drift+wrap, scan + scatter).
usage: python tools/bench_pm_kernels.py [ncoarse=9]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import _lib, mesh, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
N = 2 ** nc
_lib.load()


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


pos, vel, _ = bench.slab_ics(N, 0, N)
pos_mor, vel = utils.reorder_particles(pos, vel)
del pos
g = torch.Generator(device="cuda").manual_seed(1)
p = pos_mor + torch.randn(pos_mor.shape, generator=g, device="cuda") * (1.0 / N)
p = p - torch.floor(p)
p[p >= 1.0] = 0.0
cases = {"morton": pos_mor, "morton+drift1.0": p.contiguous()}
# Poisson: uniform random positions, Morton-ordered (an evolved but unclustered field)
pr = torch.rand(pos_mor.shape, generator=g, device="cuda")
pr[pr >= 1.0] = 0.0
cases["poisson (Morton)"] = utils.reorder_particles(pr.contiguous())
del pr
# clustered: half of the particles in 4096 Gaussian blobs of sigma = 0.7 cell (bins with ~10^4 particles)
n = pos_mor.shape[0]
centres = torch.rand((4096, 3), generator=g, device="cuda")
which = torch.randint(0, 4096, (n // 2,), generator=g, device="cuda")
pc = torch.cat([pos_mor[: n - n // 2], centres[which] + torch.randn((n // 2, 3), generator=g, device="cuda") * (0.7 / N)])
pc = pc - torch.floor(pc)
pc[pc >= 1.0] = 0.0
cases["clustered (Morton)"] = utils.reorder_particles(pc.contiguous())
del pc, which
phi = torch.randn((N, N, N), device="cuda")
acc = torch.randn(pos_mor.shape, device="cuda") * 1e-3
for name, p in cases.items():
    t_bin = timeit(lambda: mesh.bin_particles(p, N))
    bn = mesh.bin_particles(p, N)
    out = [f"N={N} {name:18s} bin {t_bin:6.3f} ms"]
    t_dep = timeit(lambda: mesh.deposit_rhs(p, N, 2, 1.0, 1.0, 0.0, bn))
    out.append(f"deposit {t_dep:6.3f} ms")
    v2 = vel.clone()
    t_int = timeit(lambda: mesh.interp_kick_phi(phi, None, 0.0, 0, 5, p, v2, 2, 0.0, bn))
    out.append(f"grad+interp+kick {t_int:6.3f} ms")
    print(" | ".join(out), flush=True)
    # the same particle set with the arrays THEMSELVES in bin order (the time loop's layout)
    z = torch.zeros_like(p)
    sb = mesh.step_sorted(p.shape[0], N)
    t_sort = timeit(lambda: mesh.step_sort(p, vel, z, None, np.float32(0), np.float32(0), 0, sb))
    sp, sv, sid = mesh.step_sort(p, vel, z, None, np.float32(0), np.float32(0), 0, sb)
    t_dep = timeit(lambda: mesh.deposit_rhs(sp, N, 2, 1.0, 1.0, 0.0, sb))
    t_int = timeit(lambda: mesh.interp_kick_phi(phi, None, 0.0, 0, 5, sp, sv, 2, 0.0, sb))
    print(f"N={N} {name:18s} [bin-ordered arrays, global-atomic sort] kick+drift+wrap+sort {t_sort:6.3f} ms | "
          f"deposit {t_dep:6.3f} ms | grad+interp+kick {t_int:6.3f} ms", flush=True)
    # the steady state of the time loop: the input of the sort is the bin-ordered output of the previous one (one CTA
    # per source bin, shared-memory sort by destination bin and micro-block); a quarter-cell rms drift per step
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sp, sv, sid = mesh.step_sort(sp, sv, z, sid, np.float32(0), np.float32(0.25 / (N * 1e-3)), 0, sb)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t_dep = timeit(lambda: mesh.deposit_rhs(sp, N, 2, 1.0, 1.0, 0.0, sb))
    t_int = timeit(lambda: mesh.interp_kick_phi(phi, None, 0.0, 0, 5, sp, sv, 2, 0.0, sb))
    print(f"N={N} {name:18s} [bin-ordered arrays, local sort, 6th step] kick+drift+wrap+sort {min(ts[1:]):6.3f} ms | "
          f"deposit {t_dep:6.3f} ms | grad+interp+kick {t_int:6.3f} ms", flush=True)
    del sp, sv, sid, z
# kick + drift + wrap + binning inside a step: count -> scan -> scatter (mode 0) against the direct scatter (mode 1)
p = pos_mor.clone()
v = vel.clone()
acc0 = torch.zeros_like(acc)
bn = mesh.alloc_binned(n, N)


def pair(events):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    mesh.kick_drift_wrap_count(p, v, acc0, np.float32(0), np.float32(1e-2), 0, bn)
    e[1].record()
    mesh.finish_binning(p, bn)
    e[2].record()
    torch.cuda.synchronize()
    events.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))


for mode in (0, 1):
    ev = []
    for _ in range(6):
        bn.ready = bool(mode)
        pair(ev)
    k, f = min(x[0] for x in ev[1:]), min(x[1] for x in ev[1:])
    print(f"N={N} mode {mode} ({'direct scatter' if mode else 'count -> scan -> scatter'}): kick+drift+wrap(+bin) {k:6.3f} ms"
          f" | finish {f:6.3f} ms | sum {k + f:6.3f} ms", flush=True)
