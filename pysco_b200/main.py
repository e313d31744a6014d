"""Mirror of pysco/main.py: ``run(param)`` -- the N-body driver loop (main.py:30-156) over the B200
PM step, and the ``-c param.ini`` command line (main.py:159-169).

Scope: the per-step path (integration.integrate -> solver.pm, Morton reorder every
n_reorder steps, snapshots) runs on the GPU.  Initial conditions: 1LPT / 2LPT / 3LPT generation on the
device (pysco_b200/initial_conditions.py), ``initial_state=(position, velocity)``, a snapshot number
(``initial_conditions = <int>``, parquet) or an ``.npz`` file holding ``position``/``velocity``; the
RayGal ``.h5`` and Gadget readers (h5py / readgadget are not in this image) raise NotImplementedError.
"""
import logging
import os
from typing import Dict

import numpy as np
import pandas as pd
import torch

from . import _lib, cosmotable, integration, iostream, solver, utils


def _initial_state(param, initial_state, tables):
    if initial_state is not None:
        return initial_state
    ic = param["initial_conditions"]
    if isinstance(ic, str) and ic.casefold() in ("1lpt", "2lpt", "3lpt"):
        from . import initial_conditions
        return initial_conditions.generate(param, tables)
    if isinstance(ic, (int, np.integer)):
        # restart from snapshot i (initial_conditions.py:79-107): every parameter saved with the snapshot comes back
        # (aexp, t, nsteps, i_snap, but also the multigrid tolerances `tolerance`, `tolerance_FAS`, so that a restarted
        # multigrid / f(R) run refreshes them at the same step phase as the uninterrupted one).  The reference means to
        # keep the caller's nthreads (its `is not` test on strings never fires, so it overwrites it as well); host
        # threads do not exist on this path, the caller's value is kept.
        fmt = str(param["output_snapshot_format"]).casefold()
        if fmt == "hdf5":
            raise NotImplementedError("restart from an HDF5 snapshot needs h5py, which this image does not have; "
                                      "write snapshots with output_snapshot_format = parquet")
        if fmt != "parquet":
            raise ValueError(f"{param['output_snapshot_format']=}, should be 'parquet' or 'hdf5'")
        d = f"{param['base']}/output_{int(ic):05d}"
        pos, vel = iostream.read_snapshot_particles_parquet(f"{d}/particles_{param['extra']}.parquet")
        saved = iostream.read_param_file(f"{d}/param_{param['extra']}_{int(ic):05d}.txt")
        for key in saved.index:
            if key.casefold() != "nthreads":
                param[key] = saved[key]
        param["nsteps"], param["i_snap"] = int(param["nsteps"]), int(param["i_snap"])
        return pos.astype(np.float32), vel.astype(np.float32)
    if isinstance(ic, str) and ic.endswith(".npz"):
        z = np.load(ic)
        return z["position"].astype(np.float32), z["velocity"].astype(np.float32)
    raise NotImplementedError(
        f"initial_conditions={ic!r}: should be 1LPT, 2LPT, 3LPT, a snapshot number or an .npz file "
        "(the reference's RayGal .h5 and Gadget readers need h5py / readgadget, absent here); "
        "or pass initial_state=(position, velocity)")


def run(param, initial_state=None):
    """main.py:30-156.  Returns (position, velocity) device tensors of the final state in addition to
    writing the reference's snapshots / P(k) files."""
    if param["verbose"] == 0:
        level = logging.ERROR
    elif param["verbose"] == 1:
        level = logging.WARNING
    elif param["verbose"] == 2:
        level = logging.INFO
    else:
        raise ValueError(f"{param['verbose']=}, should be 0, 1 or 2")
    logging.basicConfig(level=level, format="%(message)s", force=True)
    if isinstance(param, Dict):
        param = pd.Series(param)
    elif not isinstance(param, pd.Series):
        raise ValueError(f"{type(param)=}, should be a dictionnary or a Pandas Series")
    param["write_snapshot"] = False
    if param["nthreads"] <= 0:
        param["nthreads"] = os.cpu_count()
    logging.warning(f"{param['nthreads']=} (host threads are not used by the B200 path)")
    logging.warning(f"FFT module: cuFFT; device: {torch.cuda.get_device_name(_lib.device())}")

    extra = param["theory"].casefold()
    if extra == "fr":
        extra += f"{param['fR_logfR0']}_n{param['fR_n']}"
    elif extra == "mond":
        mond_function = param["mond_function"].casefold()
        extra += f"_g0_{param['mond_g0']}_exponent_{param['mond_scale_factor_exponent']}_{mond_function}"
        if "simple" != mond_function:
            extra += f"_{param['mond_alpha']}"
    elif extra == "parametrized":
        extra += f"_mu0_{param['parametrized_mu0']}"
    extra += f"_{param['linear_newton_solver']}_ncoarse{param['ncoarse']}"
    param["extra"] = extra
    z_out = iostream.parse_z_out(param)
    os.makedirs(f"{param['base']}/power", exist_ok=True)
    for i in range(len(z_out) + 1):
        os.makedirs(f"{param['base']}/output_{i:05d}", exist_ok=True)

    tables = cosmotable.generate(param)
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    if "nsteps" not in param.index:
        param["nsteps"] = 0
    position, velocity = _initial_state(param, initial_state, tables)
    utils.set_units(param)
    param["t"] = tables[1](np.log(param["aexp"]))
    logging.warning(f"{param['aexp']=} {param['t']=}")

    dev = _lib.device()
    if not isinstance(position, torch.Tensor):
        position = torch.as_tensor(np.ascontiguousarray(position), dtype=torch.float32)
        velocity = torch.as_tensor(np.ascontiguousarray(velocity), dtype=torch.float32)
    position = position.to(dev, torch.float32).contiguous()
    velocity = velocity.to(dev, torch.float32).contiguous()
    acceleration, potential, additional_field = solver.pm(position, param)
    aexp_out = 1.0 / (np.array(z_out) + 1)
    aexp_out.sort()
    t_out = tables[1](np.log(aexp_out))
    if "i_snap" not in param.index:
        param["i_snap"] = 1
    else:
        param["i_snap"] += 1

    # the time loop allocates a few thousand small Python objects per step: without this the cyclic collector makes a
    # full pass over the whole import-time heap (torch, pandas, scipy: ~3 ms) every couple of dozen steps
    import gc
    gc.collect()
    gc.freeze()
    while param["aexp"] < aexp_out[-1]:
        param["nsteps"] += 1
        position, velocity, acceleration, potential, additional_field = integration.integrate(
            position, velocity, acceleration, potential, additional_field, tables, param,
            t_out[param["i_snap"] - 1])
        if (param["nsteps"] % param["n_reorder"]) == 0:
            logging.info("Reordering particles")
            position, velocity, acceleration = utils.reorder_particles(position, velocity, acceleration)
        if param["write_snapshot"]:
            # the device-resident arrays are in bin order between reorders: snapshots keep the reference's rows
            iostream.write_snapshot_particles(*utils.reference_order(position, velocity), param)
            param["i_snap"] += 1
        logging.warning(f"{param['nsteps']=} {param['aexp']=} z = {1.0 / param['aexp'] - 1}")
    return utils.reference_order(position, velocity)


def main():
    import argparse
    from time import perf_counter

    parser = argparse.ArgumentParser()
    parser.add_argument("-c", "--config_file", help="Configuration file", required=True)
    args = parser.parse_args()
    param = iostream.read_param_file(args.config_file)
    print(param)
    t_start = perf_counter()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # launched by torchrun, one process per GPU: the same run on x-slabs (pysco_b200/slab.py)
        import torch.distributed as dist
        from . import distributed, slab
        distributed.init_from_env()
        slab.run(param)
        dist.barrier()
        dist.destroy_process_group()
    else:
        run(param)
    print(f"Simulation run time: {perf_counter() - t_start} seconds.")


if __name__ == "__main__":
    main()
