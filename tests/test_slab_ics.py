"""Initial conditions generated per x-slab (pysco_b200/initial_conditions.py: white_noise_fourier_block, SlabLayout,
generate_slab) against the unmodified reference's particles (tests/golden/ics.npz): every rank draws only its block
of the white noise -- bit-identical to the slice of the full draw -- and runs the LPT chain on its planes / its
transposed spectrum block; the union of the ranks' particles, put back in lattice order by their ids, is the
reference's particle set.  P = 1, 2, 4 virtual ranks (ThreadComm) on CPU tensors, world_size-2 gloo, and through
slab.run from a parameter file."""
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

with np.load(os.path.join(ROOT, "tests", "golden", "ics.npz")) as _z:
    G = {k: _z[k] for k in _z.files}     # materialised: NpzFile reads lazily through one zip handle, not thread-safe


def _ranks(P, fn):
    """fn(comm) on P virtual ranks (threads); the first real exception of a rank aborts the others' barriers and is
    re-raised here (no rank is left waiting)"""
    from pysco_b200.slab import ThreadComm
    errs = []

    def work(comm):
        try:
            fn(comm)
        except threading.BrokenBarrierError:
            pass
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            comm.w.barrier.abort()
    ts = [threading.Thread(target=work, args=(c,)) for c in ThreadComm.world(P)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]


@pytest.mark.parametrize("N,seed", [(8, 5), (12, 6), (32, 1)])
def test_noise_block_is_the_slice_of_the_full_draw(N, seed):
    """random access into the PCG64 stream: same bits as drawing everything (pinned on the reference by
    tests/test_ics.py::test_white_noise_matches_reference)"""
    from pysco_b200 import initial_conditions as ic
    full = ic.white_noise_fourier(N, np.random.default_rng(seed))
    fixed = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), True)
    unpaired = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), False)
    for P in (1, 2, 4):
        nyl = N // P
        for r in range(P):
            sl = slice(r * nyl, (r + 1) * nyl)
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl).view(np.float32),
                                  np.ascontiguousarray(full[:, sl]).view(np.float32))
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl, True, True).view(np.float32),
                                  np.ascontiguousarray(fixed[:, sl]).view(np.float32))
            assert np.array_equal(ic.white_noise_fourier_block(N, seed, r * nyl, nyl, True, False).view(np.float32),
                                  np.ascontiguousarray(unpaired[:, sl]).view(np.float32))


def test_transfer_block_is_the_slice(tmp_path):
    import pandas as pd
    from pysco_b200 import initial_conditions as ic
    param = pd.Series(cases.ic_param(str(tmp_path), npart=16 ** 3, seed=9))
    full = ic.get_transfer_grid(param)
    for y0, nyl in ((0, 16), (4, 4), (8, 8)):
        assert np.array_equal(ic.get_transfer_grid_block(param, y0, nyl), full[:, y0:y0 + nyl])


@pytest.mark.parametrize("P", [1, 2, 4])
def test_distributed_transforms_match_rfftn(P):
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    N = 16
    x = torch.from_numpy(cases.scalar_grid(N, seed=3))
    spec = torch.fft.rfftn(x, dim=(0, 1, 2))
    out = {}

    def rank(comm):
        L = ic.SlabLayout(comm, N)
        mine = x[L.x0:L.x0 + L.nxl].clone()
        s = L.fft(mine)
        out[comm.rank] = (s, L.ifft(s.clone()), L.y0, L.nyl, L.x0, L.nxl)
    _ranks(P, rank)
    assert len(out) == P
    scale = float(spec.abs().max())
    for r, (s, back, y0, nyl, x0, nxl) in out.items():
        assert float((s - spec[:, y0:y0 + nyl]).abs().max()) < 1e-5 * scale
        assert float((back - x[x0:x0 + nxl]).abs().max()) < 1e-5 * float(x.abs().max())


def _tables_of(name):
    t = G[f"{name}_tables"]   # [H(lna), D1(0), D1(lna), f1, D2, f2, D3a, f3a, D3b, f3b, D3c, f3c]
    return [None, None, lambda x: t[0], lambda x: t[1] if x == 0 else t[2]] + \
           [(lambda v: (lambda x: v))(v) for v in t[3:]]


def _param_of(name, tmp_path):
    import pandas as pd
    from pysco_b200 import utils
    param = pd.Series(cases.ic_param(str(tmp_path), **cases.IC_CASES[name]))
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    return param


def _check_against_reference(name, parts):
    rpos, rvel = G[f"{name}_pos"], G[f"{name}_vel"]
    ids = np.concatenate([p[2] for p in parts])
    assert np.array_equal(np.sort(ids), np.arange(len(rpos)))
    pos, vel = np.empty_like(rpos), np.empty_like(rvel)
    pos[ids] = np.concatenate([p[0] for p in parts])
    vel[ids] = np.concatenate([p[1] for p in parts])
    d = np.abs(pos - rpos)
    d = np.minimum(d, 1 - d)
    assert d.max() < 2e-6, d.max()                                   # box units (cell = 1/16): test_ics.py's bar
    assert np.max(np.abs(vel - rvel)) < 2e-5 * np.sqrt(np.mean(rvel.astype(np.float64) ** 2)) + 1e-9
    assert pos.min() >= 0 and pos.max() < 1


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("name", ["lpt1_edge", "lpt2", "lpt3", "lpt2_fixed_paired", "lpt3_dealiased"])
def test_generate_slab_matches_reference(name, P, tmp_path):
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    out = {}
    param0 = _param_of(name, tmp_path)      # written once (the power-spectrum file), copied per rank
    tables = _tables_of(name)

    def rank(comm):
        pos, vel, ids = ic.generate_slab(param0.copy(), tables, comm, device="cpu")
        n3 = 16 ** 3 // P
        assert pos.shape == (n3, 3) and ids[0] == comm.rank * n3 and ids[-1] == (comm.rank + 1) * n3 - 1
        out[comm.rank] = (pos.numpy(), vel.numpy(), ids.numpy())
    _ranks(P, rank)
    assert len(out) == P
    _check_against_reference(name, [out[r] for r in range(P)])
    # the layout is per thread and gone afterwards: the single-domain generator is untouched
    assert ic._layout() is ic._WHOLE


@pytest.mark.parametrize("P", [1, 2, 4])
def test_slab_pad_and_trim_match_the_whole_grid(P):
    """the 3/2-rule regridding of transposed blocks (ky rows change owner) against initial_conditions.pad / trim"""
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    N = 16
    rng = np.random.default_rng(4)
    x = torch.from_numpy((rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1)))
                         .astype(np.complex64))
    big = ic.pad(x)
    Ne = big.shape[0]
    y = torch.from_numpy((rng.standard_normal(tuple(big.shape)) + 1j * rng.standard_normal(tuple(big.shape)))
                         .astype(np.complex64))
    small = ic.trim(y)
    out = {}

    def rank(comm):
        L = ic.SlabLayout(comm, N)
        r, nyl, nye = comm.rank, N // P, Ne // P
        out[r] = (L.regrid(x[:, r * nyl:(r + 1) * nyl].clone(), Ne), L.regrid(y[:, r * nye:(r + 1) * nye].clone(), N))
    _ranks(P, rank)
    for r in range(P):
        assert torch.equal(out[r][0], big[:, r * (Ne // P):(r + 1) * (Ne // P)])
        assert torch.equal(out[r][1], small[:, r * (N // P):(r + 1) * (N // P)])


def test_generate_slab_dealiased_needs_a_divisible_grid(tmp_path):
    """3 N / 2 = 24 planes do not split over 16 ranks: NotImplementedError (slab_ics = replicated is the way then)"""
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    param = _param_of("lpt3_dealiased", tmp_path)
    ic._tls.layout = ic.SlabLayout(ThreadComm.world(16)[3], 16)
    try:
        with pytest.raises(NotImplementedError):
            ic._dealias_in(param, torch.zeros((16, 1, 9), dtype=torch.complex64))
        ic._tls.layout = ic.SlabLayout(ThreadComm.world(8)[3], 16)       # 24 planes over 8 ranks: fine
        param["dealiased_ICS"] = False
        (x,) = ic._dealias_in(param, torch.zeros((16, 2, 9), dtype=torch.complex64))
        assert x.shape == (16, 2, 9)
    finally:
        del ic._tls.layout
    assert ic._layout() is ic._WHOLE


def _run_slab(base, how, P, cuda=False, ncoarse=4, **over):
    from pysco_b200 import slab
    OracleOps = None
    if not cuda:
        from slab_oracle_ops import OracleOps
    os.makedirs(base, exist_ok=True)
    pk = cases.ic_pk_file(base)
    out, errs = {}, []

    def work(c):
        try:
            param = cases.run_param(base + "/", "fft", ncoarse=ncoarse)
            param.update(power_spectrum_file=pk, z_out="[30, 0]", save_power_spectrum="no", slab_ics=how)
            param.update(over)
            res = slab.run(param, comm=c, ops_factory=OracleOps)
            out[c.rank] = res
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            c.w.barrier.abort()
    ts = [threading.Thread(target=work, args=(c,)) for c in slab.ThreadComm.world(P)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]
    return out


def test_slab_run_generates_its_initial_conditions_per_slab(tmp_path):
    """slab.run from a parameter file on 2 virtual ranks (oracle kernels): per-slab generation (the default on more
    than one rank) against the replicated generator -- same initial snapshot, same final particles"""
    import glob
    import pyarrow.parquet as pq
    a = _run_slab(str(tmp_path / "slab"), "auto", 2)
    b = _run_slab(str(tmp_path / "repl"), "replicated", 2)
    ics = []
    for d in ("slab", "repl"):
        f = glob.glob(str(tmp_path / d / "output_00000" / "particles_*.parquet"))
        assert len(f) == 1 and glob.glob(str(tmp_path / d / "output_00000" / "param_*.txt"))
        t = pq.read_table(f[0])
        ics.append(np.stack([np.asarray(t.column(c)) for c in ("x", "y", "z", "vx", "vy", "vz")], axis=1))
    dx = np.abs(ics[0][:, :3] - ics[1][:, :3])
    assert np.minimum(dx, 1 - dx).max() < 2e-6
    assert np.abs(ics[0][:, 3:] - ics[1][:, 3:]).max() < 2e-5 * np.abs(ics[1][:, 3:]).max()
    pa, va = (t.numpy() for t in a[0])
    pb, vb = (t.numpy() for t in b[0])
    d = np.abs(pa - pb)
    assert np.minimum(d, 1 - d).max() < 1e-4           # a cell is 6e-2; the ICs differ by transform rounding only
    assert np.abs(va - vb).max() < 1e-3 * np.abs(vb).max()
    assert len(glob.glob(str(tmp_path / "slab" / "output_0000[12]" / "particles_*.parquet"))) == 2


def test_slab_run_per_slab_ics_in_parts(tmp_path):
    """slab_snapshots = parts: the initial snapshot is written slab by slab as well; together the parts hold every
    lattice particle exactly once"""
    import glob
    from pysco_b200 import iostream
    out = _run_slab(str(tmp_path / "parts"), "slab", 2, slab_snapshots="parts")
    assert sorted(out) == [0, 1] and all(len(o) == 3 for o in out.values())
    d = glob.glob(str(tmp_path / "parts" / "output_00000" / "particles_*.parquet"))
    assert len(d) == 1 and len(glob.glob(d[0] + "/part-*.parquet")) == 2
    ids = np.concatenate([iostream.read_snapshot_slab_parts(d[0], r, 2)[2] for r in range(2)])
    assert np.array_equal(np.sort(ids), np.arange(16 ** 3))


def test_slab_ics_switch_is_validated(tmp_path):
    from pysco_b200 import slab
    import pandas as pd
    p = pd.Series({"slab_ics": "sometimes", "dealiased_ICS": False})
    with pytest.raises(NotImplementedError):
        slab._ics_per_slab(p, slab.SelfComm())
    assert not slab._ics_per_slab(pd.Series({"dealiased_ICS": False}), slab.SelfComm())
    comms = slab.ThreadComm.world(2)
    assert slab._ics_per_slab(pd.Series({"dealiased_ICS": False}), comms[0])
    assert slab._ics_per_slab(pd.Series({"dealiased_ICS": True, "npart": 16 ** 3}), comms[0])
    assert not slab._ics_per_slab(pd.Series({"dealiased_ICS": True, "npart": 16 ** 3}), slab.ThreadComm.world(16)[0])
    assert not slab._ics_per_slab(pd.Series({"dealiased_ICS": True, "npart": 16 ** 3, "slab_ics": "replicated"}), comms[0])


def _gloo_ics_worker(rank, world, port, base, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pathlib
    import torch.distributed as dist
    from pysco_b200 import distributed, initial_conditions as ic, slab
    distributed.init_from_env("gloo")
    comm = slab.default_comm()
    assert isinstance(comm, slab.TorchComm) and comm.size == world
    sub = pathlib.Path(base) / f"rank{rank}"          # every process writes its own copy of the P(k) table
    sub.mkdir(parents=True, exist_ok=True)
    pos, vel, ids = ic.generate_slab(_param_of(name, sub), _tables_of(name), comm, device="cpu")
    out[rank] = (pos.numpy(), vel.numpy(), ids.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["lpt2", "lpt3_dealiased"])
def test_generate_slab_gloo_world2(name, tmp_path):
    """the same over torch.distributed (gloo, two processes): the transposes go through all_to_all_single, the ky rows
    of the dealiasing grid through all_to_all_single with split sizes"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_ics_worker, args=(2, port, str(tmp_path), name, out), nprocs=2, join=True)
        out = dict(out)
    _check_against_reference(name, [out[0], out[1]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lpt2", "lpt3"])
def test_generate_slab_matches_reference_gpu(name, tmp_path):
    """the per-slab generator on CUDA tensors (cuFFT through torch.fft), 2 virtual ranks sharing cuda:0"""
    from pysco_b200 import initial_conditions as ic
    from pysco_b200.slab import ThreadComm
    out = {}
    param0 = _param_of(name, tmp_path)
    tables = _tables_of(name)

    def rank(comm):
        pos, vel, ids = ic.generate_slab(param0.copy(), tables, comm)
        assert pos.is_cuda
        out[comm.rank] = (pos.cpu().numpy(), vel.cpu().numpy(), ids.cpu().numpy())
    _ranks(2, rank)
    _check_against_reference(name, [out[0], out[1]])


@pytest.mark.gpu
def test_slab_run_per_slab_ics_gpu(tmp_path):
    """slab.run from a parameter file with the CUDA kernels on 2 virtual ranks at 32^3, z = 49 -> 0: initial
    conditions generated per slab against the replicated generator"""
    a = _run_slab(str(tmp_path / "slab"), "auto", 2, cuda=True, ncoarse=5)
    b = _run_slab(str(tmp_path / "repl"), "replicated", 2, cuda=True, ncoarse=5)
    pa, va = (t.numpy() for t in a[0])
    pb, vb = (t.numpy() for t in b[0])
    d = np.abs(pa - pb)
    # two runs whose initial conditions differ by float32 rounding of the transforms (1e-7), ~45 steps of growth and
    # float atomics in the deposit: measured 1e-5 box units / 2e-4 of max |v| (128^3 over NCCL,
    # profiles/r02_slab_ics_nccl_n2.txt); wrong initial conditions are off by a cell (3e-2) or more
    assert np.minimum(d, 1 - d).max() < 2e-4, np.minimum(d, 1 - d).max()
    assert np.abs(va - vb).max() < 2e-3 * np.abs(vb).max()


def test_command_line_runs_on_slabs_under_torchrun(tmp_path, monkeypatch):
    """`torchrun -m pysco_b200.main -c param.ini`: with WORLD_SIZE > 1 the command line hands the run to slab.run"""
    import torch.distributed as dist
    from pysco_b200 import distributed, main, slab
    cfg = tmp_path / "param.ini"
    cfg.write_text("theory = newton\nnpart = 4096\nncoarse = 4\n")
    called = {}
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(sys, "argv", ["pysco_b200", "-c", str(cfg)])
    monkeypatch.setattr(distributed, "init_from_env", lambda backend=None: called.setdefault("init", True))
    monkeypatch.setattr(slab, "run", lambda param, **kw: called.setdefault("param", param))
    monkeypatch.setattr(main, "run", lambda param: called.setdefault("single", True))
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    main.main()
    assert called.get("init") and "single" not in called and int(called["param"]["ncoarse"]) == 4
    called.clear()
    monkeypatch.setenv("WORLD_SIZE", "1")
    main.main()
    assert called == {"single": True}
