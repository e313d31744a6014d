"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path (pysco_b200/distributed.py).

The decomposition itself (particle-parallel, mesh-replicated: local deposit -> all-reduce(sum) -> replicated
solve -> local interpolation, all-reduce(max) for the time step) is exercised with the CPU oracle standing in
for the CUDA kernels, and compared with the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases
    import oracle
    from oracle import host
    from pysco_b200 import distributed
    distributed.init_from_env("gloo")
    assert distributed.is_active() and distributed.world_size() == world and distributed.rank() == rank

    # balanced contiguous ranges that tile [0, n)
    n = 32 ** 3 + 5
    lo, hi = distributed.local_range(n)
    ranges = [distributed.local_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1

    # decomposition: local deposit + all-reduce == global deposit; replicated solve; local interpolation
    N = 32
    pos = cases.lattice_particles(N, 0.3, seed=3)
    vel = cases.velocities(N ** 3, seed=4, scale=2e-3)
    lo, hi = distributed.local_range(pos.shape[0])
    param = cases.base_param(5, N ** 3, linear_newton_solver="fft")
    host.set_units(param)
    rho_local = torch.from_numpy(oracle.mesh.TSC_seq(np.ascontiguousarray(pos[lo:hi]), N))
    distributed.allreduce_sum_(rho_local)
    rho = rho_local.numpy()
    f1 = np.float32(1.5 * param["aexp"] * param["Om_m"])
    oracle.utils.linear_operator_inplace(rho, f1, -f1)
    param["MAS_index"] = 3
    param["compute_additional_field"] = False
    param["save_pk"] = False
    phi = host.fft(rho, param)
    acc_local = oracle.mesh.invTSC_vec(oracle.mesh.derivative(phi, 5), np.ascontiguousarray(pos[lo:hi]))
    mx = torch.tensor([float(np.abs(acc_local).max()), float(np.abs(vel[lo:hi]).max())])
    distributed.allreduce_max_(mx)
    if rank == 0:
        acc_ref, phi_ref, _ = host.pm(pos, cases.base_param(5, N ** 3, linear_newton_solver="fft").pipe(
            lambda p: (host.set_units(p), p)[1]))
        out["phi_err"] = float(np.abs(phi - phi_ref).max() / np.abs(phi_ref).max())
        out["acc_err"] = float(np.abs(acc_local - acc_ref[lo:hi]).max() / np.abs(acc_ref).max())
        out["max_ok"] = bool(abs(mx[0].item() - np.abs(acc_ref).max()) < 1e-5 * np.abs(acc_ref).max()
                             and mx[1].item() == np.abs(vel).max())
    dist.barrier()
    dist.destroy_process_group()


def test_particle_parallel_decomposition_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out["phi_err"] < 1e-5, out["phi_err"]
        assert out["acc_err"] < 1e-5, out["acc_err"]
        assert out["max_ok"]


def test_single_process_is_inactive():
    from pysco_b200 import distributed
    assert not distributed.is_active()
    assert distributed.world_size() == 1 and distributed.rank() == 0
    assert distributed.local_range(10) == (0, 10)
    t = torch.ones(3)
    assert distributed.allreduce_sum_(t) is t
