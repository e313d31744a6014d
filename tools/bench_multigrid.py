#!/usr/bin/env python
"""Timings of the multigrid kernels and of a full multigrid.linear solve (config 2 shape).
usage: python tools/bench_multigrid.py [ncoarse=8]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import bench  # noqa: E402
from pysco_b200 import _lib, laplacian, mesh, multigrid, cubic, fourier  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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


g = torch.Generator(device="cuda").manual_seed(3)
b = torch.randn((N, N, N), generator=g, device="cuda")
bk = torch.fft.rfftn(b)
k1 = torch.fft.fftfreq(N, device="cuda") * N
kz = torch.fft.rfftfreq(N, device="cuda") * N
k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
b = torch.fft.irfftn(bk * torch.exp(-k2 / (2 * (N / 16.0) ** 2)), s=(N, N, N)).contiguous()
b -= b.mean()
b /= b.std()
x = laplacian.initialise_potential(b)
peak = bench.measured_peak_gbs()[0]
cells = N ** 3


def report(name, ms, bytes_per_cell):
    gbs = bytes_per_cell * cells / (ms * 1e-3) / 1e9
    print(f"N={N} {name:34s} {ms:8.3f} ms  {gbs:7.0f} GB/s algorithmic ({bytes_per_cell} B/cell)  frac {gbs / peak:.3f}", flush=True)


report("gauss_seidel (1 sweep, two launches)", timeit(lambda: laplacian.gauss_seidel(x, b, np.float32(1.25))), 12)
report("smoothing x2 (fused pair), per sweep", timeit(lambda: laplacian.smoothing(x, b, 2)) / 2, 12)
report("restrict_residual", timeit(lambda: laplacian.restrict_residual(x, b)), 8.5)
xc = mesh.restriction(x)
report("add_prolongation", timeit(lambda: mesh.add_prolongation(x, xc)), 8.5)
report("residual_error", timeit(lambda: laplacian.residual_error(x, b)), 8)
report("operator", timeit(lambda: laplacian.operator(x)), 8)
u = torch.ones_like(b) + 0.05 * b
bb = 2.0 * (1 + 0.1 * b)
report("cubic gauss_seidel (two launches)", timeit(lambda: cubic.gauss_seidel(u, bb, np.float32(-2.0), np.float32(1.25))), 12)
report("cubic smoothing x2 (fused), per sweep", timeit(lambda: cubic.smoothing(u, bb, np.float32(-2.0), 2)) / 2, 12)
param = bench.make_param(nc, 1)
param["linear_newton_solver"] = "multigrid"
param["compute_additional_field"] = False


def solve():
    p = param.copy()
    y = laplacian.initialise_potential(b)
    multigrid.linear(y, b, p)


report("multigrid.linear (cold start)", timeit(solve, reps=3), 80)
p = param.copy()
y = laplacian.initialise_potential(b)
report("V_cycle", timeit(lambda: multigrid.V_cycle(y, b, p), reps=3), 80)
spec = fourier.fft_3D_real(b)
report("inverse_laplacian_compensated", timeit(lambda: fourier.inverse_laplacian_compensated(spec, 3, 1.0)), 8)
report("inverse_laplacian_7pt", timeit(lambda: fourier.inverse_laplacian_7pt(spec, 1.0)), 8)
report("derivative5 (float4 out)", timeit(lambda: mesh.derivative(b, 5, padded=True)), 16)
report("derivative5 (AoS out)", timeit(lambda: mesh.derivative(b, 5)), 16)
