"""Host-side I/O of the run loop (reference pysco/iostream.py): parameter file
reader, parquet/HDF5 particle snapshots and the P(k) ascii writer, with the reference's file names
and on-disk layout.  Device tensors are copied to the host here."""
import ast
import logging
import os

import numpy as np
import pandas as pd


def _host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def read_param_file(name: str) -> pd.Series:
    """iostream.py:13-69: ``key = value  # comment`` lines; values evaluated when they are Python
    literals / arithmetic (e.g. ``128**3``), lists kept as strings, true/false case-folded."""
    out = {}
    with open(name) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            key, val = (s.strip() for s in line.split("=", 1))
            if val == "":
                val = "False"
            if val.casefold() in ("true", "false"):
                val = val.capitalize()
            try:
                v = eval(val, {"__builtins__": {}}, {})  # same contract as the reference (eval per value)
                out[key] = val if isinstance(v, list) else v
            except Exception:
                out[key] = val
    return pd.Series(out)


def write_power_spectrum_to_ascii_file(k, Pk, Nmodes, param) -> None:
    """iostream.py:268-304"""
    output_pk = f"{param['base']}/power/pk_{param['extra']}_{param['nsteps']:05d}.dat"
    os.makedirs(os.path.dirname(output_pk), exist_ok=True)
    logging.warning(f"Write P(k) in {output_pk}")
    np.savetxt(
        output_pk, np.c_[k, Pk, Nmodes],
        header=f"aexp = {param['aexp']}\nboxlen = {param['boxlen']} Mpc/h \nnpart = {param['npart']} \n"
               f"k [h/Mpc] P(k) [Mpc/h]^3 Nmodes")


def write_snapshot_particles(position, velocity, param) -> None:
    """iostream.py:136-182"""
    fmt = param["output_snapshot_format"].casefold()
    position, velocity = _host(position), _host(velocity)
    d = f"{param['base']}/output_{param['i_snap']:05d}"
    os.makedirs(d, exist_ok=True)
    if fmt == "parquet":
        import pyarrow as pa
        import pyarrow.parquet as pq
        filename = f"{d}/particles_{param['extra']}.parquet"
        pq.write_table(pa.table({"x": position[:, 0], "y": position[:, 1], "z": position[:, 2],
                                 "vx": velocity[:, 0], "vy": velocity[:, 1], "vz": velocity[:, 2]}), filename)
        param.to_csv(f"{d}/param_{param['extra']}_{param['i_snap']:05d}.txt", sep="=", header=False)
    elif fmt == "hdf5":
        import h5py  # optional dependency, as in the reference
        filename = f"{d}/particles_{param['extra']}.h5"
        with h5py.File(filename, "w") as h5f:
            h5f.create_dataset("position", data=position)
            h5f.create_dataset("velocity", data=velocity)
            for key, item in param.items():
                h5f.attrs[key] = item
    else:
        raise NotImplementedError(f"{param['output_snapshot_format']=}, should be 'parquet' or 'hdf5'")
    logging.warning(f"Snapshot written at ...{filename=} {param['aexp']=}")


def write_snapshot_particles_parquet(filename: str, position, velocity) -> None:
    """iostream.py:185-214"""
    import pyarrow as pa
    import pyarrow.parquet as pq
    position, velocity = _host(position), _host(velocity)
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    pq.write_table(pa.table({"x": position[:, 0], "y": position[:, 1], "z": position[:, 2],
                             "vx": velocity[:, 0], "vy": velocity[:, 1], "vz": velocity[:, 2]}), filename)


def read_snapshot_particles_parquet(filename: str):
    """iostream.py:111-133"""
    import pyarrow.parquet as pq
    def cols(names):
        t = pq.read_table(filename, columns=names)
        return np.ascontiguousarray(np.stack([np.asarray(t.column(n)) for n in names], axis=1))   # [Npart, 3]

    return cols(["x", "y", "z"]), cols(["vx", "vy", "vz"])


# ---------------------------------------------------------------------------- per-slab snapshots (multi-GPU runs)
# A slab-decomposed run writes a snapshot as a DIRECTORY with the name of the reference's file,
#     output_0000i/particles_<extra>.parquet/part-<rank>.parquet      columns x, y, z, vx, vy, vz, id
# one part per rank, written by that rank from its own slab: no rank ever holds the global arrays (103 GB of x, v at
# 2048^3).  pyarrow reads the directory as one table, so `iostream.read_snapshot_particles_parquet` of the reference
# still works on it (rows in slab order; `id` is the row of the particle in the reference's order).
def slab_snapshot_dir(param):
    return f"{param['base']}/output_{param['i_snap']:05d}/particles_{param['extra']}.parquet"


def write_snapshot_slab_part(position, velocity, ids, param, rank, write_param) -> str:
    """iostream.py:136-214 for one slab: this rank's particles as part-<rank>.parquet (+ the parameter file, once)"""
    import pyarrow as pa
    import pyarrow.parquet as pq
    if str(param["output_snapshot_format"]).casefold() != "parquet":
        raise NotImplementedError("per-slab snapshots are written as parquet (HDF5 needs h5py: absent)")
    position, velocity, ids = _host(position), _host(velocity), _host(ids)
    d = slab_snapshot_dir(param)
    os.makedirs(d, exist_ok=True)
    filename = f"{d}/part-{rank:05d}.parquet"
    pq.write_table(pa.table({"x": position[:, 0], "y": position[:, 1], "z": position[:, 2],
                             "vx": velocity[:, 0], "vy": velocity[:, 1], "vz": velocity[:, 2],
                             "id": ids.astype(np.int64)}), filename)
    if write_param:
        param.to_csv(f"{os.path.dirname(d)}/param_{param['extra']}_{param['i_snap']:05d}.txt", sep="=", header=False)
    return filename


def read_snapshot_slab_parts(path, rank, size):
    """The parts rank `rank` of `size` adopts when restarting from a per-slab snapshot (parts rank, rank + size, ...:
    the number of ranks may differ from the run that wrote it), or its strided share of a single-file snapshot.
    Returns (position [n,3], velocity [n,3], ids [n]) float32 / int64."""
    import pyarrow.parquet as pq

    def load(f, with_id):
        cols = ["x", "y", "z", "vx", "vy", "vz"] + (["id"] if with_id else [])
        t = pq.read_table(f, columns=cols)
        a = [np.asarray(t.column(c)) for c in cols]
        return (np.stack(a[0:3], axis=1).astype(np.float32), np.stack(a[3:6], axis=1).astype(np.float32),
                a[6].astype(np.int64) if with_id else None)

    if os.path.isdir(path):
        parts = sorted(f for f in os.listdir(path) if f.startswith("part-") and f.endswith(".parquet"))
        mine = [load(os.path.join(path, f), True) for f in parts[rank::size]]
        if not mine:
            return (np.empty((0, 3), np.float32), np.empty((0, 3), np.float32), np.empty((0,), np.int64))
        return tuple(np.concatenate([m[i] for m in mine]) for i in range(3))
    pos, vel, _ = load(path, False)      # a gathered snapshot: the row is the id
    ids = np.arange(pos.shape[0], dtype=np.int64)
    return (np.ascontiguousarray(pos[rank::size]), np.ascontiguousarray(vel[rank::size]),
            np.ascontiguousarray(ids[rank::size]))


def parse_z_out(param):
    return ast.literal_eval(param["z_out"]) if isinstance(param["z_out"], str) else list(param["z_out"])
