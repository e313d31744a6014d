#!/usr/bin/env python
"""psc_xfft_green_slab alone: the per-rank block of BASELINE config 5 (N = 2048, 256 ky per rank) and a 512^3 spectrum.
usage: python tools/bench_xfft.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysco_b200 import _lib  # noqa: E402

L = _lib.load()
for N, nyl in ((512, 512), (1024, 256), (2048, 256)):
    nz = N // 2 + 1
    a = torch.view_as_complex(torch.randn((N, nyl, nz, 2), device="cuda"))
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.psc_xfft_green_slab(_lib.ptr(a), N, nyl, 0, 1, 3, 1.0, _lib.stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    gb = 2 * a.numel() * 8 / 1e9
    print(f"N={N} nyl={nyl}: {min(ts[1:]):.3f} ms, {gb / min(ts[1:]) * 1e3 / 1e3:.2f} TB/s (read + write of the block)", flush=True)
    del a
