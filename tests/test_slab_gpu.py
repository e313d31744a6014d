"""GPU parity of the x-slab decomposed step (pysco_b200/slab.py + csrc/slab.cu + the slab forms of the binned
kernels): P = 1, 2, 4 virtual ranks (threads sharing cuda:0, ThreadComm) run the whole path -- migration, ghost
planes, transposed FFT -- through the C ABI and are compared with the oracle's single-process leapfrog and with
the single-domain CUDA path."""
import os
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
import test_slab_cpu as cpu  # noqa: E402

pytestmark = pytest.mark.gpu


def _run_rank_cuda(N, comm, out, order, reorder_at=None, solver="fft", **overrides):
    from pysco_b200 import slab
    tables, pos, vel, param = cpu._setup(N, solver, **overrides)
    param["gradient_stencil_order"] = order
    P, r = comm.size, comm.rank
    ids = np.arange(N ** 3, dtype=np.int64)
    mine = slice(r, None, P)
    s = slab.Slab(N, comm=comm)
    assert isinstance(s.ops, slab.CudaOps)
    s.set_particles(torch.from_numpy(pos[mine].copy()).cuda(), torch.from_numpy(vel[mine].copy()).cuda(),
                    torch.from_numpy(ids[mine].copy()).cuda())
    s.pm(param, tables=tables)
    moved = 0
    for step in range(cpu.NSTEPS):
        param["nsteps"] += 1
        if reorder_at is not None and step == reorder_at:
            s.reorder()
        s.integrate(tables, param, 1e30)
        moved += s.migrated_last[0]
    phi_planes = s.potential.clone()
    res = s.gather_to_root(N ** 3)
    tot = torch.tensor([float(moved)], device="cuda")
    comm.allreduce_sum_(tot)
    counts = [0] * P
    counts[0] = phi_planes.shape[0]
    phi = comm.all_to_all_v(phi_planes.reshape(phi_planes.shape[0], -1), counts, comm.exchange_counts(counts))
    add = None
    if s.additional_field is not None:      # MOND: the Newtonian potential, f(R): the scalaron
        ap = s.additional_field.clone()
        add = comm.all_to_all_v(ap.reshape(ap.shape[0], -1), counts, comm.exchange_counts(counts))
    if r == 0:
        out["additional_field"] = None if add is None else add.cpu().numpy().reshape(N, N, N)
        out["state"] = [t.numpy() for t in res] + [phi.cpu().numpy().reshape(N, N, N)]
        out["t"] = float(param["t"])
        out["moved"] = float(tot[0])
    s.ops.close()


def _threads(P, fn):
    from pysco_b200 import slab
    comms = slab.ThreadComm.world(P) if P > 1 else [slab.SelfComm()]
    out, errs = {}, []

    def work(c):
        try:
            torch.cuda.set_device(0)
            fn(c, out)
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


@pytest.mark.parametrize("P,N,order", [(1, 32, 5), (2, 32, 5), (4, 32, 5), (2, 32, 7), (4, 64, 3), (2, 32, 2)])
def test_slab_cuda_vs_oracle(P, N, order):
    from oracle import host
    tables, pos, vel, param = cpu._setup(N)
    param["gradient_stencil_order"] = order
    pos, vel = pos.copy(), vel.copy()
    acc, phi, add = host.pm(pos, param)
    state = [pos, vel, acc, phi, add]
    for _ in range(cpu.NSTEPS):
        param["nsteps"] += 1
        state = list(host.integrate(*state, tables, param, 1e30))
    out = _threads(P, lambda c, o: _run_rank_cuda(N, c, o, order, reorder_at=1))
    cpu._check(out, state, float(param["t"]), P)


@pytest.mark.parametrize("P,N", [(1, 32), (2, 32), (4, 32), (8, 64)])
def test_slab_cuda_multigrid_vs_oracle(P, N):
    """linear_newton_solver = multigrid on slabs (csrc/slab_mg.cu + slab_multigrid.py): ghost-plane red-black sweeps,
    local restriction / prolongation, gathered coarse levels (P = 4: the 4^3 level; P = 8 at 64^3: the 8^3 level, which
    then runs the single-domain V-cycle), warm start from the previous potential -- three steps against the oracle's
    single-process multigrid."""
    ref, ref_t = cpu._reference(N, "multigrid")
    out = _threads(P, lambda c, o: _run_rank_cuda(N, c, o, 5, reorder_at=1, solver="multigrid"))
    cpu._check(out, ref, ref_t, P)


def test_slab_cuda_multigrid_matches_single_domain_kernels():
    """One rank, one V-cycle: the ghost-plane kernels reproduce the single-domain multigrid kernels (same per-cell
    arithmetic, csrc/slab_mg_cells.cuh vs csrc/multigrid.cu; bit for bit when the compiler contracts both alike)."""
    from pysco_b200 import multigrid, slab
    from pysco_b200.slab_multigrid import SlabMultigrid
    N = 32
    g = torch.Generator().manual_seed(5)
    b = (torch.randn((N, N, N), generator=g) * 3).cuda()
    b -= b.mean()
    param = cases.base_param(5, N ** 3, linear_newton_solver="multigrid")
    x_ref = torch.zeros((N, N, N), device="cuda")
    multigrid._cycle("V", x_ref, b, param, 0)
    s = slab.Slab(N, comm=slab.SelfComm())
    mg = SlabMultigrid(s.comm, s.ops, N)
    xg = torch.zeros((N + 2, N, N), device="cuda")
    mg.v_cycle(xg, b, N, N, param, 0)
    diff = (xg[1:N + 1] - x_ref).abs().max().item()
    print(f"slab V-cycle vs single-domain V-cycle: max |diff| = {diff:.3e} (0 = bit-identical)")
    assert diff <= 2e-6 * x_ref.abs().max().item()
    r_ref = np.float32(np.sqrt(((b - multigrid.laplacian.operator(x_ref)).double() ** 2).sum().item()))
    assert abs(mg.residual_error(xg, b, N, N) - r_ref) <= 1e-5 * r_ref
    s.ops.close()


@pytest.mark.parametrize("N,P", [(64, 4), (128, 2)])
def test_slab_cuda_matches_single_domain_path(N, P):
    """Same kernels, same inputs: the slab path on P virtual ranks against integration.integrate on one domain.
    128^3 on 2 ranks: 2048 bins per rank for 592 persistent CTAs -- every CTA of the per-source-bin sort walks over
    several bins (TMA pipeline, stage reuse), with migration in between."""
    from pysco_b200 import integration, solver
    tables, pos, vel, param = cpu._setup(N)
    p, v = torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda()
    acc, phi, add = solver.pm(p, param)
    state = [p, v, acc, phi, add]
    for _ in range(cpu.NSTEPS):
        param["nsteps"] += 1
        state = list(integration.integrate(*state, tables, param, 1e30))
    from pysco_b200 import utils
    state[:3] = utils.reference_order(*state[:3])      # the device-resident loop keeps its arrays in bin order
    ref = [state[0].cpu().numpy(), state[1].cpu().numpy(), state[2].cpu().numpy(), state[3].cpu().numpy()]
    out = _threads(P, lambda c, o: _run_rank_cuda(N, c, o, 5))
    cpu._check(out, ref, float(param["t"]), P)


@pytest.mark.parametrize("solver_name", ["fft", "multigrid"])
def test_pinned_host_pipeline_matches_device_path(solver_name):
    """integration.integrate with PINNED host tensors (chunked, overlapped PCIe transfers) against the same call
    with device tensors: same kernels, same arithmetic (only the order of the float atomics differs)."""
    from pysco_b200 import integration, solver
    N = 64
    tables, pos, vel, param = cpu._setup(N)
    param["linear_newton_solver"] = solver_name
    p2 = param.copy()

    def run(to, prm):
        state = [to(torch.from_numpy(pos.copy())), to(torch.from_numpy(vel.copy()))]
        a, phi, add = solver.pm(state[0], prm)
        state += [a, phi, add]
        for _ in range(3):
            prm["nsteps"] += 1
            state = list(integration.integrate(*state, tables, prm, 1e30))
        return state

    from pysco_b200 import utils
    dev = run(lambda t: t.cuda(), param)
    dev[:3] = utils.reference_order(*dev[:3])          # bin order -> the rows the caller passed in
    host = run(lambda t: t.pin_memory(), p2)
    assert all(not t.is_cuda and t.is_pinned() for t in host[:4])
    assert abs(param["t"] - p2["t"]) <= 1e-7 * abs(param["t"])
    for name, a, b in zip(("pos", "vel", "acc", "phi"), host[:4], dev[:4]):
        b = b.cpu()
        err = ((a - b).abs().max() / b.abs().max()).item()
        assert err < 1e-5, (name, err)


@pytest.mark.parametrize("solver", ["fft", "multigrid"])
def test_slab_whole_run_matches_reference_snapshot(tmp_path, solver):
    """BASELINE config 1 shape at 32^3, z = 49 -> 0, through slab.run on 2 virtual ranks: the final state, put back
    in the reference's particle order, against the unmodified reference's final snapshot (tests/golden/run.npz).
    The slab path restores the reference's particle order through the particle ids, so the comparison is row by
    row."""
    import cases
    from pysco_b200 import slab
    with np.load(os.path.join(ROOT, "tests", "golden", "run.npz")) as z:
        # materialised here: NpzFile reads lazily through one zip handle, which the rank threads must not share
        g = {k: z[k] for k in ("ic_pos", "ic_vel", f"{solver}_pos", f"{solver}_vel", f"{solver}_nsteps",
                               f"{solver}_pk_last")}
    base = str(tmp_path) + "/"
    out = {}

    def work(c, o):
        param = cases.run_param(base, solver)    # save_power_spectrum = z_out: P(k) at every snapshot
        res = slab.run(param, comm=c, initial_state=(g["ic_pos"].copy(), g["ic_vel"].copy()))
        if c.rank == 0:
            o["pos"], o["vel"] = res[0].numpy(), res[1].numpy()

    out = _threads(2, work)
    # fewer than n_reorder = 50 steps: the reference never reorders, its final rows are the initial (lattice = id) rows
    assert int(g[f"{solver}_nsteps"][0]) < 50
    d = np.abs(out["pos"] - g[f"{solver}_pos"])
    d = np.minimum(d, 1 - d)
    assert d.max() < 1e-5, d.max()
    assert np.abs(out["vel"] - g[f"{solver}_vel"]).max() < 1e-4 * np.abs(g[f"{solver}_vel"]).max() + 1e-7
    glob = __import__("glob")
    snaps = glob.glob(os.path.join(base, "output_0000[1-6]", "particles_*.parquet"))
    assert len(snaps) == 6
    # the last P(k) file against the reference's (north-star bar 1e-4), summed over the two slabs' bins
    pks = sorted(glob.glob(os.path.join(base, "power", "*.dat")))
    mine, ref = np.loadtxt(pks[-1]), g[f"{solver}_pk_last"]
    assert np.array_equal(mine[:, 2], ref[:, 2])
    np.testing.assert_allclose(mine[:, 0], ref[:, 0], rtol=1e-6)
    np.testing.assert_allclose(mine[:, 1], ref[:, 1], rtol=1e-4)


@pytest.mark.parametrize("P", [2, 4])
def test_slab_pm_edge_positions_and_ragged_counts(P):
    """Slab deposit / interpolation with particles exactly on slab boundaries (x = r/P), at 0 and just below 1, with
    npart != N^3 and very uneven counts per slab: acceleration of every particle against the oracle's solver.pm."""
    from oracle import host
    from pysco_b200 import slab
    N, npart = 32, 20011
    pos = cases.particles(N, npart, seed=41)
    k = npart // 2
    pos[16:16 + P, 0] = (np.arange(P, dtype=np.float32) / P)                     # exactly on the slab boundaries
    pos[32:k, 0] = pos[32:k, 0] * np.float32(0.5 / P)                            # half of the particles in slab 0
    pos = np.ascontiguousarray(pos.astype(np.float32))
    param_ref = cases.base_param(5, npart, linear_newton_solver="fft")
    host.set_units(param_ref)
    acc_ref, _, _ = host.pm(pos.copy(), param_ref)

    def work(c, o):
        param = cases.base_param(5, npart, linear_newton_solver="fft")
        host.set_units(param)
        s = slab.Slab(N, comm=c)
        ids = torch.arange(npart, dtype=torch.int64, device="cuda")
        mine = slice(c.rank, None, c.size)
        s.set_particles(torch.from_numpy(pos).cuda()[mine].contiguous(), torch.zeros((npart, 3), device="cuda")[mine],
                        ids[mine].contiguous())
        s.pm(param)
        res = s.gather_to_root(npart)
        if c.rank == 0:
            o["acc"] = res[2].numpy()
            o["pos"] = res[0].numpy()
        s.ops.close()

    out = _threads(P, work)
    assert np.array_equal(out["pos"], pos)
    err = np.max(np.abs(out["acc"] - acc_ref)) / np.sqrt(np.mean(acc_ref.astype(np.float64) ** 2))
    assert err < 5e-5, err
