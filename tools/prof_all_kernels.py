#!/usr/bin/env python
"""One launch (after a warm-up launch) of every grid / spectral / multigrid / f(R) / MOND / reorder kernel of the
library at 2^nc cells per side, for an `ncu --set full` capture (evidence for every kernel, not only the
particle kernels).  usage: ncu --set full -k regex:psc:: ... python tools/prof_all_kernels.py [nc=9]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import _lib, cubic, fourier, laplacian, mesh, mond, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
N = 2 ** nc
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn((N, N, N), generator=g, device="cuda") * 1e-3
b = torch.randn((N, N, N), generator=g, device="cuda")
xc = torch.randn((N // 2,) * 3, generator=g, device="cuda") * 1e-3
u = torch.rand((N, N, N), generator=g, device="cuda") * 0.1 + 1.0
bu = 2.0 + 0.3 * b
out = torch.empty_like(x)
for rep in range(2):   # launch 1 warms up (lazy module load, plans), launch 2 is the one to read
    spec = fourier.fft_3D_real(b.clone(), 1)
    fourier.inverse_laplacian(spec, 1.0)
    fourier.inverse_laplacian_compensated(spec, 3, 1.0)
    fourier.inverse_laplacian_7pt(spec, 1.0)
    fourier.fourier_grid_to_Pk(spec, 3)
    fourier.ifft_3D_real(spec, 1)
    del spec
    mesh.derivative(x, 5)
    utils.linear_operator_inplace(out, np.float32(0.3), np.float32(-0.3))
    laplacian.gauss_seidel(x, b, np.float32(1.25))
    laplacian.smoothing(x, b, 2)                                 # the fused sweep (TMA) on grids >= 128
    laplacian.residual_error(x, b)
    laplacian.operator(x)
    laplacian.restrict_residual(x, b)
    laplacian.initialise_potential(b)
    mesh.restriction(x)
    mesh.add_prolongation(x, xc)
    cubic.gauss_seidel(u, bu, np.float32(-2.0), np.float32(1.25))
    cubic.smoothing(u, bu, np.float32(-2.0), 2)
    cubic.residual_error(u, bu, np.float32(-2.0))
    mond.rhs_simple(x, out, np.float32(0.05))
torch.cuda.synchronize()
# reorder: keys + radix sort + gathers
pos, vel, _ = bench.slab_ics(min(N, 256), 0, min(N, 256))
for rep in range(2):
    utils.reorder_particles(pos, vel)
torch.cuda.synchronize()
print("done", N)
