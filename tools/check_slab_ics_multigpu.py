#!/usr/bin/env python
"""Multi-GPU check of the per-slab initial conditions over NCCL: slab.run from a parameter file (2LPT, z = 49 -> 0)
with `slab_ics = slab` (every rank generates its own lattice planes: the transposes of the LPT chain go through
all_to_all_single over NVLink) against `slab_ics = replicated` (every rank generates everything), same kernels.
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_slab_ics_multigpu.py [ncoarse=7]"""
import faulthandler
import logging
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import cases  # noqa: E402
from pysco_b200 import distributed, initial_conditions as ic, slab  # noqa: E402

logging.disable(logging.WARNING)
faulthandler.enable()
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 7
rank = int(os.environ.get("RANK", "0"))
distributed.init_from_env()
comm = slab.default_comm()
base = os.path.join(tempfile.gettempdir(), "psc_slab_ics_check")
if rank == 0:
    pk = cases.ic_pk_file(base)
comm.barrier()
pk = os.path.join(base, "pk_test.dat")
res, secs = {}, {}
for how in ("slab", "replicated"):
    param = cases.run_param(os.path.join(base, how) + "/", "fft", ncoarse=nc)
    param.update(power_spectrum_file=pk, z_out="[30, 0]", save_power_spectrum="no", slab_ics=how)
    torch.cuda.synchronize()
    t = time.perf_counter()
    res[how] = slab.run(param, comm=comm)
    torch.cuda.synchronize()
    secs[how] = time.perf_counter() - t
    if rank == 0:
        print(f"slab.run with slab_ics = {how}: {secs[how]:.1f} s", flush=True)
    sys.stdout.flush()
if rank == 0:
    pa, va = (x.numpy() for x in res["slab"])
    pb, vb = (x.numpy() for x in res["replicated"])
    d = np.abs(pa - pb)
    d = np.minimum(d, 1 - d)
    print(f"N = {2 ** nc}, {comm.size} ranks over NCCL: final particles, per-slab ICs vs replicated ICs: max |dx| "
          f"{d.max():.2e} box units ({d.max() * 2 ** nc:.2e} cells), max |dv| / max |v| "
          f"{np.abs(va - vb).max() / np.abs(vb).max():.2e}; whole run {secs['slab']:.1f} s vs {secs['replicated']:.1f} s")
    print("OK" if d.max() < 1e-4 else "MISMATCH", flush=True)
try:
    # the generator alone, timed
    param = cases.run_param(os.path.join(base, "gen") + "/", "fft", ncoarse=nc)
    param.update(power_spectrum_file=pk)
    import pandas as pd  # noqa: E402
    from pysco_b200 import cosmotable, utils  # noqa: E402
    param = pd.Series(param)
    param["base"] = ""
    tables = cosmotable.generate(param)
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    torch.cuda.synchronize()
    t = time.perf_counter()
    p, v, i = ic.generate_slab(param, tables, comm)
    torch.cuda.synchronize()
    t_slab = time.perf_counter() - t
    if rank == 0:
        print(f"generate_slab alone: {t_slab:.2f} s for {p.shape[0]} particles on rank 0")
except Exception as e:  # noqa: BLE001
    print("generator timing failed:", repr(e)[:300])
comm.barrier()
torch.distributed.barrier()
torch.distributed.destroy_process_group()
