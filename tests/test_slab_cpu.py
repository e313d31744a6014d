"""CPU tests of the x-slab decomposition's host logic (pysco_b200/slab.py): migration bookkeeping, ghost-plane
exchanges, the transposed-FFT all-to-alls and the time-step reduction, with the oracle standing in for the CUDA
kernels (tests/slab_oracle_ops.py).  Two transports are covered: ThreadComm (P = 1, 2, 4 virtual ranks in one
process) and TorchComm over gloo with world_size 2 (spawned processes).  The yardstick is the oracle's
single-process leapfrog (oracle/host.py)."""
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
import oracle  # noqa: E402
from oracle import host  # noqa: E402

NSTEPS = 3


def _setup(N, solver="fft", **overrides):
    tables = cases.toy_tables()
    pos = cases.lattice_particles(N, 0.4, seed=11)
    vel = cases.velocities(N ** 3, seed=12, scale=0.3)  # large enough that particles cross slab boundaries
    aexp = overrides.pop("aexp", 0.2) if overrides else 0.2
    if overrides and overrides.pop("boost_slab0_right", False):
        # a stream of particles about to leave slab 0 (of 4) to the right: one neighbour pair migrates ~7x more
        # particles than every other pair
        m = (pos[:, 0] > 0.21) & (pos[:, 0] < 0.25)
        vel[m, 0] = np.float32(0.9)
    param = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver=solver, **overrides)
    param["aexp"] = param["aexp_old"] = aexp
    param["t"] = float(tables[1](np.log(param["aexp"])))
    host.set_units(param)
    return tables, pos, vel, param


def _reference(N, solver="fft", **overrides):
    tables, pos, vel, param = _setup(N, solver, **overrides)
    pos, vel = pos.copy(), vel.copy()
    acc, phi, add = host.pm(pos, param, tables=tables)
    state = [pos, vel, acc, phi, add]
    for _ in range(NSTEPS):
        param["nsteps"] += 1
        state = list(host.integrate(*state, tables, param, 1e30))
    return state, float(param["t"])


def _run_rank(N, comm, out, reorder_at=None, solver="fft", mig_cap=None, **overrides):
    from pysco_b200 import slab
    from slab_oracle_ops import OracleOps
    tables, pos, vel, param = _setup(N, solver, **overrides)
    P, r = comm.size, comm.rank
    # every rank adopts an arbitrary 1/P of the particles: set_particles must route them to their owners
    ids = np.arange(N ** 3, dtype=np.int64)
    mine = slice(r, None, P)
    s = slab.Slab(N, comm=comm, ops=OracleOps(N, P, r))
    s.set_particles(torch.from_numpy(pos[mine].copy()), torch.from_numpy(vel[mine].copy()),
                    torch.from_numpy(ids[mine].copy()))
    own = (s.position[:, 0].numpy() * np.float32(N)).astype(np.int64) // (N // P)
    assert (own == r).all()
    if P == 4:
        s._mig_cap = 8     # far too small: the first steps must take the overflow (repeat) path
    if mig_cap is not None:
        s._mig_cap = mig_cap
    s.pm(param, tables=tables)
    moved = 0
    for step in range(NSTEPS):
        param["nsteps"] += 1
        if reorder_at is not None and step == reorder_at:
            s.reorder()
        s.integrate(tables, param, 1e30)
        moved += s.migrated_last[0]
    phi_planes = s.potential.clone()
    res = s.gather_to_root(N ** 3)
    tot = torch.tensor([float(moved)])
    comm.allreduce_sum_(tot)
    # the potential: gather the owned planes on rank 0
    counts = [0] * P
    counts[0] = phi_planes.shape[0]
    phi = comm.all_to_all_v(phi_planes.reshape(phi_planes.shape[0], -1), counts, comm.exchange_counts(counts))
    add = None
    if s.additional_field is not None:      # MOND: the Newtonian potential, f(R): the scalaron
        ap = s.additional_field.clone()
        add = comm.all_to_all_v(ap.reshape(ap.shape[0], -1), counts, comm.exchange_counts(counts))
    out.setdefault("redo", {})[r] = s.redo_count
    if r == 0:
        out["state"] = [t.numpy() for t in res] + [phi.numpy().reshape(N, N, N)]
        out["additional_field"] = None if add is None else add.numpy().reshape(N, N, N)
        out["t"] = float(param["t"])
        out["moved"] = float(tot[0])


def _check(out, ref, ref_t, P):
    pos, vel, acc, phi = out["state"]
    rpos, rvel, racc, rphi = ref[0], ref[1], ref[2], ref[3]
    assert abs(out["t"] - ref_t) <= 1e-6 * abs(ref_t)
    if P > 1:
        assert out["moved"] > 0, "the test must exercise migration"

    def rel(a, b):
        return float(np.max(np.abs(a.astype(np.float64) - b)) / max(np.sqrt(np.mean(b.astype(np.float64) ** 2)), 1e-30))
    d = np.abs(pos - rpos)
    d = np.minimum(d, 1 - d)  # periodic distance
    assert d.max() < 2e-6
    assert rel(vel, rvel) < 5e-5
    assert rel(acc, racc) < 2e-4
    assert rel(phi, rphi) < 2e-4
    if out.get("additional_field") is not None:
        assert len(ref[4]) and rel(out["additional_field"], np.asarray(ref[4]).reshape(phi.shape)) < 2e-4


@pytest.mark.parametrize("P,solver,N", [(1, "fft", 32), (2, "fft", 32), (4, "fft", 32), (1, "multigrid", 32),
                                        (2, "multigrid", 32), (4, "multigrid", 32), (8, "multigrid", 64)])
def test_slab_threads_vs_oracle(P, solver, N):
    """multigrid: P = 1, 2 keep every level distributed (two planes per rank on the coarsest 4^3 grid at P = 2);
    P = 4 gathers the 4^3 level and smooths it redundantly; P = 8 at 64^3 gathers the 8^3 level and runs the rest of
    the V-cycle on it."""
    from pysco_b200 import slab
    ref, ref_t = _reference(N, solver)
    comms = slab.ThreadComm.world(P) if P > 1 else [slab.SelfComm()]
    out, errs = {}, []

    def work(c):
        try:
            _run_rank(N, c, out, reorder_at=1, solver=solver)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            if P > 1:
                c.w.barrier.abort()

    ts = [threading.Thread(target=work, args=(c,)) for c in comms]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    _check(out, ref, ref_t, P)


def test_migration_overflow_of_one_pair_only():
    """the overflow decision is per message.  Only the pair (0, 1) of four ranks migrates more particles
    than the fixed-capacity buffers hold; ranks 0 and 1 repeat that one message point to point while ranks 2 and 3,
    whose messages all fitted, must neither wait for them nor be disturbed -- and the result is the oracle's."""
    N, P = 32, 4
    ref, ref_t = _reference(N, "fft", boost_slab0_right=True)
    out = _run_threads(P, N, "fft", dict(boost_slab0_right=True, mig_cap=300))
    _check(out, ref, ref_t, P)
    redo = out["redo"]
    assert redo[0] > 0 and redo[1] > 0, redo          # the two ends of the overflowing message
    assert redo[2] == 0 and redo[3] == 0, redo        # bystanders


MOND_CASES = [(1, "fft_7pt", dict(theory="mond")),
              (2, "fft_7pt", dict(theory="mond")),
              (4, "fft_7pt", dict(theory="mond", mond_function="beta", mond_alpha=1.5)),
              (2, "multigrid", dict(theory="mond", mond_function="n", mond_alpha=2)),
              (4, "multigrid", dict(theory="mond", mond_function="gamma", mond_alpha=1.5, mass_scheme="CIC"))]


def _run_threads(P, N, solver, overrides, runner=None):
    from pysco_b200 import slab
    comms = slab.ThreadComm.world(P) if P > 1 else [slab.SelfComm()]
    out, errs = {}, []

    def work(c):
        try:
            (runner or _run_rank)(N, c, out, reorder_at=1, solver=solver, **overrides)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            if P > 1:
                c.w.barrier.abort()

    ts = [threading.Thread(target=work, args=(c,)) for c in comms]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("P,solver,overrides", MOND_CASES)
def test_slab_mond_threads_vs_oracle(P, solver, overrides):
    """theory = mond on slabs (QUMOND: Newtonian solve -> psc_box_mond_rhs on the ghosted Newtonian potential ->
    second solve, with the two warm starts and the two tolerances of the multigrid variant): three steps against the
    oracle's single-process step."""
    N = 32
    ref, ref_t = _reference(N, solver, **dict(overrides))
    out = _run_threads(P, N, solver, dict(overrides))
    _check(out, ref, ref_t, P)


# f(R) in the screened (early) regime, where the reference's cubic root stays on its defined branches
FR_CASES = [(1, "multigrid", dict(theory="fr", fR_n=1, aexp=0.05)),
            (2, "fft", dict(theory="fr", fR_n=1, aexp=0.05)),
            (4, "multigrid", dict(theory="fr", fR_n=1, aexp=0.05)),
            (2, "fft", dict(theory="fr", fR_n=2, fR_logfR0=6, aexp=0.05)),
            (4, "fft_7pt", dict(theory="fr", fR_n=2, fR_logfR0=6, aexp=0.05, mass_scheme="CIC"))]


@pytest.mark.parametrize("P,solver,overrides", FR_CASES)
def test_slab_fr_threads_vs_oracle(P, solver, overrides):
    """theory = fr on slabs: scalaron by the FAS cycle on ghosted slabs (nonlinear red-black sweeps, coarse problem
    4 R(res) + L(R x), gathered 4^3 level at P = 4), warm start from the previous scalaron, then the Newtonian solve and
    the fifth force through the gradient of phi + f u^(n+1) -- three steps against the oracle's single-process step."""
    N = 32
    ref, ref_t = _reference(N, solver, **dict(overrides))
    out = _run_threads(P, N, solver, dict(overrides))
    _check(out, ref, ref_t, P)


def test_slab_fr_8_ranks_gathered_fas_levels():
    """8 ranks at 64^3: the 8^3 FAS level is gathered (x_c, b_c and the residual) and the rest of the FAS V-cycle runs
    redundantly on every rank."""
    over = dict(theory="fr", fR_n=1, aexp=0.05)
    ref, ref_t = _reference(64, "fft", **dict(over))
    out = _run_threads(8, 64, "fft", dict(over))
    _check(out, ref, ref_t, 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out, solver="fft"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from pysco_b200 import distributed, slab
    distributed.init_from_env("gloo")
    comm = slab.default_comm()
    assert isinstance(comm, slab.TorchComm) and comm.size == world
    # ghost-plane exchange semantics: what I get "from the left" is what my left neighbour sent "to the right"
    a = torch.full((1, 4), float(10 * rank + 1))
    b = torch.full((1, 4), float(10 * rank + 2))
    fl, fr = comm.exchange_planes(a, b)
    left, right = (rank - 1) % world, (rank + 1) % world
    assert fl[0, 0].item() == 10 * left + 2 and fr[0, 0].item() == 10 * right + 1
    local = {}
    _run_rank(32, comm, local, solver=solver)
    if rank == 0:
        out.update(local)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("solver", ["fft", "multigrid"])
def test_slab_gloo_world2_vs_oracle(solver):
    import torch.multiprocessing as mp
    ref, ref_t = _reference(32, solver)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, out, solver), nprocs=2, join=True)
        out = dict(out)
    _check(out, ref, ref_t, 2)


@pytest.mark.parametrize("solver", ["fft", "multigrid"])
def test_slab_run_loop_threads_vs_reference_snapshot(tmp_path, solver):
    """slab.run (main.run on slabs: snapshot schedule, per-slab reorder, gather in reference order, P(k) files) on 2
    virtual ranks with the oracle standing in for the kernels, z = 49 -> 0 at 32^3, against the final snapshot of the
    unmodified reference (tests/golden/run.npz), for the FFT and the multigrid solver (warm starts, truncation error
    refreshed every third step, the reference's stopping rule)."""
    import glob
    from pysco_b200 import slab
    from slab_oracle_ops import OracleOps
    with np.load(os.path.join(ROOT, "tests", "golden", "run.npz")) as z:
        # materialised here: NpzFile reads lazily through one zip handle, which the rank threads must not share
        g = {k: z[k] for k in ("ic_pos", "ic_vel", f"{solver}_pos", f"{solver}_vel", f"{solver}_pk_last")}
    base = str(tmp_path) + "/"
    out, errs = {}, []
    comms = slab.ThreadComm.world(2)

    def work(c):
        try:
            param = cases.run_param(base, solver)     # save_power_spectrum = z_out: P(k) at every snapshot
            res = slab.run(param, comm=c, initial_state=(g["ic_pos"].copy(), g["ic_vel"].copy()),
                           ops_factory=OracleOps)
            if c.rank == 0:
                out["pos"], out["vel"] = res[0].numpy(), res[1].numpy()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            c.w.barrier.abort()

    ts = [threading.Thread(target=work, args=(c,)) for c in comms]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    d = np.abs(out["pos"] - g[f"{solver}_pos"])
    d = np.minimum(d, 1 - d)
    assert d.max() < 1e-5, d.max()
    assert np.abs(out["vel"] - g[f"{solver}_vel"]).max() < 1e-4 * np.abs(g[f"{solver}_vel"]).max() + 1e-7
    assert len(glob.glob(os.path.join(base, "output_0000[1-6]", "particles_*.parquet"))) == 6
    # the last P(k) file (fft: spectrum of the RHS inside the solve; multigrid: forward transform of the density)
    pks = sorted(glob.glob(os.path.join(base, "power", "*.dat")))
    mine, ref = np.loadtxt(pks[-1]), g[f"{solver}_pk_last"]
    assert np.array_equal(mine[:, 2], ref[:, 2])
    np.testing.assert_allclose(mine[:, 0], ref[:, 0], rtol=1e-6)
    np.testing.assert_allclose(mine[:, 1], ref[:, 1], rtol=1e-4)


def _gloo3_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from pysco_b200 import distributed, slab
    distributed.init_from_env("gloo")
    comm = slab.default_comm()
    left, right = (rank - 1) % world, (rank + 1) % world
    ok = True
    # ghost planes: from_left is the left neighbour's to_right, from_right the right neighbour's to_left
    fl, fr = comm.exchange_planes(torch.full((2, 3), 10.0 * rank + 1), torch.full((2, 3), 10.0 * rank + 2))
    ok &= bool((fl == 10.0 * left + 2).all() and (fr == 10.0 * right + 1).all())
    # variable-size neighbour messages (migration records): sizes differ per rank and direction
    n_l, n_r = rank + 1, 2 * rank + 1
    from_l, from_r = comm.neighbor_counts(n_l, n_r)
    ok &= from_l == 2 * left + 1 and from_r == right + 1
    got_l, got_r = comm.neighbor_exchange(torch.full((n_l, 8), float(rank)), torch.full((n_r, 8), float(rank) + 0.5),
                                          from_l, from_r)
    ok &= got_l.shape == (from_l, 8) and got_r.shape == (from_r, 8)
    ok &= bool((got_l == left + 0.5).all() and (got_r == float(right)).all())
    # equal-split all-to-all (the NCCL fallback of the FFT transposes): block s of rank r goes to rank s
    send = torch.arange(world * 4, dtype=torch.float32).view(world, 4) + 100.0 * rank
    recv = torch.empty_like(send)
    comm.all_to_all_equal(send, recv)
    want = torch.stack([torch.arange(4, dtype=torch.float32) + 4 * rank + 100.0 * s for s in range(world)])
    ok &= bool(torch.equal(recv, want))
    t = torch.tensor([float(rank), 5.0 - rank])
    comm.allreduce_max_(t)
    ok &= t.tolist() == [float(world - 1), 5.0]
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_torchcomm_semantics_gloo_world3():
    """TorchComm with three ranks (left and right neighbours are different processes): the point-to-point pairing of
    ghost planes and migration messages, the equal all-to-all and the reductions."""
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo3_worker, args=(3, port, out), nprocs=3, join=True)
        assert dict(out) == {0: True, 1: True, 2: True}


def _run_slab_threads(P, base, overrides, g):
    """slab.run on P virtual ranks (oracle kernels); returns {rank: result}"""
    from pysco_b200 import slab
    from slab_oracle_ops import OracleOps
    out, errs = {}, []
    comms = slab.ThreadComm.world(P) if P > 1 else [slab.SelfComm()]

    def work(c):
        try:
            param = cases.run_param(base, "fft")
            param.update(overrides)
            init = None if isinstance(param["initial_conditions"], int) else (g["ic_pos"].copy(), g["ic_vel"].copy())
            out[c.rank] = slab.run(param, comm=c, initial_state=init, ops_factory=OracleOps)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            if P > 1:
                c.w.barrier.abort()

    ts = [threading.Thread(target=work, args=(c,)) for c in comms]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    return out


def test_slab_snapshots_in_parts_and_restart(tmp_path):
    """SURVEY 8(f) rank 2 on slabs: every rank writes its own slab of a snapshot (no gather), the directory reads back
    as one table with the reference's reader, the run restarts from those parts on a DIFFERENT number of ranks
    (initial_conditions = snapshot number, initial_conditions.py:79-107) and ends where the uninterrupted run ends --
    which is the reference's final snapshot (tests/golden/run.npz)."""
    import glob
    from pysco_b200 import iostream
    with np.load(os.path.join(ROOT, "tests", "golden", "run.npz")) as z:
        g = {k: z[k] for k in ("ic_pos", "ic_vel", "fft_pos", "fft_vel")}
    base = str(tmp_path) + "/"
    over = dict(slab_snapshots="parts", save_power_spectrum="no")
    full = _run_slab_threads(2, base, over, g)

    def assemble(res):
        pos = torch.cat([r[0].cpu() for r in res.values()]).numpy()
        vel = torch.cat([r[1].cpu() for r in res.values()]).numpy()
        ids = torch.cat([r[2].cpu() for r in res.values()]).numpy()
        assert np.array_equal(np.sort(ids), np.arange(len(ids)))
        o = np.argsort(ids)
        return pos[o], vel[o]

    pos, vel = assemble(full)
    d = np.abs(pos - g["fft_pos"])
    assert np.minimum(d, 1 - d).max() < 1e-5
    # on disk: one directory per snapshot, one part per rank, readable as one table by the reference's reader
    snap = glob.glob(os.path.join(base, "output_00003", "particles_*.parquet"))
    assert len(snap) == 1 and os.path.isdir(snap[0])
    assert sorted(os.listdir(snap[0])) == ["part-00000.parquet", "part-00001.parquet"]
    p3, v3 = iostream.read_snapshot_particles_parquet(snap[0])
    assert p3.shape == (32 ** 3, 3) and v3.shape == (32 ** 3, 3)
    # restart from snapshot 3 on ONE rank, then from the same snapshot on FOUR ranks
    for P in (1, 4):
        again = _run_slab_threads(P, base, dict(over, initial_conditions=3), g)
        pos2, vel2 = assemble(again)
        d = np.abs(pos2 - pos)
        assert np.minimum(d, 1 - d).max() < 2e-6, f"restart on {P} rank(s)"
        assert np.abs(vel2 - vel).max() < 2e-5 * np.abs(vel).max()
