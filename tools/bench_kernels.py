#!/usr/bin/env python
"""Per-kernel timings of the particle<->mesh kernels on one GPU (not the contract bench; see bench.py).
usage: python tools/bench_kernels.py [ncoarse=9]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pysco_b200 import _lib, mesh, utils  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 9
N = 2 ** nc
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.psc_deposit_window_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]


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


pos_lex, vel = bench.synthetic_ics_device(N)
pos_mor = utils.reorder_particles(pos_lex)
g = torch.Generator(device="cuda").manual_seed(1)
cases = {"morton": pos_mor, "lexicographic": pos_lex}
for sig in (1.0, 3.0):
    p = pos_mor + torch.randn(pos_mor.shape, generator=g, device="cuda") * (sig / N)
    p = p - torch.floor(p)
    p[p >= 1.0] = 0.0
    cases[f"morton+drift{sig}"] = p.contiguous()
cases["random"] = pos_mor[torch.randperm(pos_mor.shape[0], device="cuda")].contiguous()
np_ = pos_mor.shape[0]
rho = torch.empty((N, N, N), device="cuda")
stats = torch.zeros(3, dtype=torch.int64, device="cuda")
force = torch.randn((N, N, N, 3), device="cuda")
force4 = torch.cat([force, torch.zeros((N, N, N, 1), device="cuda")], dim=3).contiguous()
raw.psc_interp_kick4.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
mxo = torch.zeros(2, device="cuda")
acc4 = torch.empty((pos_mor.shape[0], 3), device="cuda")
raw.psc_deposit_window_dbg.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
for dbg in (0, 1, 2, 3, 4, 7):
    t = timeit(lambda: raw.psc_deposit_window_dbg(pos_mor.data_ptr(), np_, N, dbg, rho.data_ptr(), None))
    print(f"window kernel only, dbg={dbg} (1: no syncwarp, 2: no merge, 4: no flush): {t:7.3f} ms", flush=True)
for name, p in cases.items():
    os.environ.pop("PSC_DEPOSIT_MODE", None)
    t_win = timeit(lambda: _lib.check(lib.psc_deposit(p.data_ptr(), np_, N, 2, 1.0, 1.0, 0.0, rho.data_ptr(), 0)))
    ref = rho.clone()
    raw.psc_deposit_window_stats(p.data_ptr(), np_, N, 2, rho.data_ptr(), stats.data_ptr(), None)
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    err = (rho - ref).abs().max().item()
    mass = rho.sum(dtype=torch.float64).item() / np_
    vel2 = vel.clone()
    t_int = timeit(lambda: mesh.interp_kick(force, p, vel2, 2, 0.0))
    os.environ["PSC_INTERP_MODE"] = "direct"
    t_int4 = timeit(lambda: raw.psc_interp_kick4(force4.data_ptr(), p.data_ptr(), vel2.data_ptr(), acc4.data_ptr(), np_, N, 2, 0.0, mxo.data_ptr(), None))
    a_ref, _ = mesh.interp_kick(force, p, None, 2, 0.0)
    err4 = (a_ref - acc4).abs().max().item()
    os.environ.pop("PSC_INTERP_MODE", None)
    t_ww = timeit(lambda: raw.psc_interp_kick4(force4.data_ptr(), p.data_ptr(), vel2.data_ptr(), acc4.data_ptr(), np_, N, 2, 0.0, mxo.data_ptr(), None))
    errw = (a_ref - acc4).abs().max().item()
    from pysco_b200 import mesh as _m
    t_bin = timeit(lambda: _m.bin_particles(p, N))
    bn = _m.bin_particles(p, N)
    t_dep_b = timeit(lambda: _m.deposit_rhs(p, N, 2, 1.0, 1.0, 0.0, bn))
    rho_b = _m.deposit_rhs(p, N, 2, 1.0, 1.0, 0.0, bn)
    err_b = (rho_b - ref).abs().max().item()
    vel3 = vel.clone()
    t_int_b = timeit(lambda: _m.interp_kick(force4, p, vel3, 2, 0.0, bn))
    a_b, _ = _m.interp_kick(force4, p, None, 2, 0.0, bn)
    err_ib = (a_b - a_ref).abs().max().item()
    print(f"N={N} {name:18s} BINNED: bin {t_bin:6.3f} ms  deposit {t_dep_b:6.3f} ms (maxdiff {err_b:.1e})  "
          f"interp_kick {t_int_b:6.3f} ms (maxdiff {err_ib:.1e})", flush=True)
    print(f"N={N} {name:18s} deposit(total) {t_win:7.3f} ms  fallback {100.0 * st[0] / np_:6.2f}%  "
          f"floats RED'ed/particle {4.0 * st[1] / np_:5.2f}  reanchors/1k {1000.0 * st[2] / np_:6.2f}  "
          f"rerun maxdiff {err:.1e} mass {mass:.7f} | interp_kick {t_int:7.3f} ms | direct float4 {t_int4:7.3f} ms (maxdiff {err4:.1e}) | warp-window float4 {t_ww:7.3f} ms (maxdiff {errw:.1e})", flush=True)
