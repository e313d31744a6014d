"""GPU tests of the cell-sorted particle <-> mesh kernels (csrc/binned.cu) through the C ABI: the counting sort's
invariants, and both paths of the deposit (lanes own cells / lanes own particles) and of the gradient stage of the
interpolation (row-wise / cell-wise) against the oracle, on uniform, clustered and mixed particle sets.

Tolerance: max|diff| <= 5e-6 rms(reference) for float32 fields; the sort (keys, permutation) is exact."""
import numpy as np
import pytest

import cases
from conftest import assert_close, rel_err

pytestmark = pytest.mark.gpu

TOL = 5e-6


@pytest.fixture(scope="module")
def psc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pysco_b200
    pysco_b200._lib.load()
    yield pysco_b200
    pysco_b200._lib.load().psc_set_kernel_modes(0, 0)


@pytest.fixture(scope="module")
def orc():
    import oracle
    oracle.build()
    return oracle


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _a256(x):
    return (x + 255) & ~255


def _cell_key(pos, N):
    """512 * bin + 64 (i & 7) + 8 (j & 7) + (k & 7): the sort key of binned.cu (cell_of)."""
    c = np.minimum((pos * np.float32(N)).astype(np.int64), N - 1)
    NB = N // 8
    b = ((c[:, 0] >> 3) * NB + (c[:, 1] >> 3)) * NB + (c[:, 2] >> 3)
    return (b << 9) | ((c[:, 0] & 7) << 6) | ((c[:, 1] & 7) << 3) | (c[:, 2] & 7)


def _mixed_particles(N, seed=3):
    """uniform background + one very dense blob + a moderately dense region: bins of all three kinds"""
    rng = np.random.default_rng(seed)
    bg = cases.lattice_particles(N, 0.3, seed=seed)
    blob = (0.31 + 0.004 * rng.standard_normal((60000, 3))).astype(np.float32)
    mid = (np.array([0.7, 0.2, 0.55], dtype=np.float32) + 0.05 * rng.standard_normal((150000, 3))).astype(np.float32)
    pos = np.concatenate([bg, blob % 1.0, mid % 1.0]).astype(np.float32)
    pos[pos >= 1.0] = 0.0
    return np.ascontiguousarray(pos[rng.permutation(len(pos))])


@pytest.mark.parametrize("kind", ["lattice", "mixed"])
def test_counting_sort_invariants(psc, kind):
    """cell[] is the exclusive scan of the per-cell counts, the records are a permutation of the rows, every record
    sits in the range of its own cell and carries the position of its source row (bit-exact)."""
    N = 64
    pos = cases.lattice_particles(N, 0.3, seed=5) if kind == "lattice" else _mixed_particles(N)
    n = len(pos)
    tp = _cuda(pos)
    bn = psc.mesh.bin_particles(tp, N)
    ncells = N ** 3
    raw = bn.scratch.cpu().numpy()
    cell = raw[: 4 * (ncells + 2)].view(np.int32)
    off = _a256(4 * (ncells + 2))
    rec = raw[off: off + 16 * n].view(np.float32).reshape(n, 4)
    rows = rec[:, 3].copy().view(np.int32)
    key = _cell_key(pos, N)
    counts = np.bincount(key, minlength=ncells)
    assert cell[0] == 0 and cell[ncells] == n and cell[ncells + 1] == n
    assert np.array_equal(np.diff(cell[: ncells + 1].astype(np.int64)), counts)
    assert np.array_equal(np.sort(rows), np.arange(n))
    assert np.array_equal(rec[:, :3], pos[rows])
    assert np.all(np.diff(key[rows]) >= 0)


def test_kick_drift_count_matches_separate_binning(psc, orc):
    """the count pass folded into kick+drift+wrap produces the same sort as psc_bin_particles of the new positions"""
    import torch
    N = 32
    n = 50003   # not a multiple of 4: the scalar tail of the fused kernel
    pos, vel = cases.particles(N, n, seed=21), cases.velocities(n, seed=22, scale=5e-3)
    acc = cases.velocities(n, seed=23, scale=1.0)
    tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
    bn = psc.mesh.alloc_binned(n, N)
    dt = np.float32(0.21)
    psc.mesh.kick_drift_wrap_count(tp, tv, ta, np.float32(0.5 * dt), dt, 0, bn)
    psc.mesh.finish_binning(tp, bn)
    p, v = pos.copy(), vel.copy()
    orc.utils.add_vector_scalar_inplace(v, acc, -np.float32(0.5 * dt))
    orc.utils.add_vector_scalar_inplace(p, v, dt)
    orc.utils.periodic_wrap(p)
    assert np.max(np.abs(tp.cpu().numpy() - p)) <= 6e-8
    ref = psc.mesh.bin_particles(tp, N)
    ncells = N ** 3
    a = bn.scratch[: 4 * (ncells + 2)].cpu().numpy().view(np.int32)
    b = ref.scratch[: 4 * (ncells + 2)].cpu().numpy().view(np.int32)
    assert np.array_equal(a, b)
    rho_a = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, bn)
    rho_b = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, ref)
    assert_close(rho_a.cpu().numpy(), rho_b.cpu().numpy(), 1e-6, "deposit from the fused count")
    torch.cuda.synchronize()


@pytest.mark.parametrize("scheme", ["TSC", "CIC", "NGP"])
@pytest.mark.parametrize("mode", [0, 1])
def test_deposit_paths_vs_oracle(psc, orc, scheme, mode):
    """mode 0: lanes own cells (uneven bins fall back per bin); mode 1: every bin on the lanes-own-particles path"""
    N = 64
    psc._lib.load().psc_set_kernel_modes(mode, -1)
    try:
        for name, pos in (("lattice", cases.lattice_particles(N, 0.3, seed=5)), ("mixed", _mixed_particles(N))):
            sid = {"NGP": 0, "CIC": 1, "TSC": 2}[scheme]
            exact = orc.mesh.deposit_f64(pos, N, sid)
            ref = getattr(orc.mesh, "TSC_seq" if scheme == "TSC" else scheme)(pos, N)
            bound = max(TOL, 3.0 * rel_err(ref, exact))
            rho = getattr(psc.mesh, scheme)(pos, N)
            assert_close(rho, exact, bound, f"{scheme} {name} mode {mode}")
            assert abs(float(rho.sum(dtype=np.float64)) - len(pos)) < 2e-6 * len(pos)
    finally:
        psc._lib.load().psc_set_kernel_modes(0, -1)


@pytest.mark.parametrize("order", [2, 3, 5, 7])
@pytest.mark.parametrize("mode", [0, 1])
def test_interp_gradient_stage_vs_oracle(psc, orc, order, mode):
    """gradient fused into the interpolation (row-wise stage, mode 0; cell-wise stage, mode 1) against
    mesh.derivative + mesh.invTSC_vec / invCIC_vec of the oracle, plain and f(R)"""
    N = 32
    psc._lib.load().psc_set_kernel_modes(-1, mode)
    try:
        pos = _mixed_particles(N, seed=9)[:120001]
        phi = cases.scalar_grid(N, seed=31, smooth=True)
        u = cases.scalaron_field(N)
        tp = _cuda(pos)
        bn = psc.mesh.bin_particles(tp, N)
        for scheme, sid in (("TSC", 2), ("CIC", 1)):
            f_ref = orc.mesh.derivative(phi, order)
            a_ref = getattr(orc.mesh, f"inv{scheme}_vec")(f_ref, pos)
            vel = cases.velocities(len(pos), seed=4, scale=1e-2)
            tv = _cuda(vel)
            a, mx = psc.mesh.interp_kick_phi(_cuda(phi), None, 0.0, 0, order, tp, tv, sid, np.float32(0.013), bn)
            assert_close(a.cpu().numpy(), a_ref, 2 * TOL, f"acc {scheme} order {order}")
            v_ref = vel.copy()
            orc.utils.add_vector_scalar_inplace(v_ref, a_ref, -np.float32(0.013))
            assert_close(tv.cpu().numpy(), v_ref, 2 * TOL, f"vel {scheme} order {order}")
            np.testing.assert_allclose(mx.cpu().numpy()[0], orc.utils.max_abs(a_ref), rtol=2e-5)
        f_ref = orc.mesh.derivative_fR(phi, u, np.float32(0.37), 1, order)
        a_ref = orc.mesh.invTSC_vec(f_ref, pos)
        a, _ = psc.mesh.interp_kick_phi(_cuda(phi), _cuda(u), np.float32(0.37), 1, order, tp, None, 2, 0.0, bn)
        assert_close(a.cpu().numpy(), a_ref, 2 * TOL, f"f(R) acc order {order}")
    finally:
        psc._lib.load().psc_set_kernel_modes(-1, 0)


def test_positions_on_the_box_edge_do_not_leave_the_grid(psc, orc):
    """y or z exactly 1.0 (an external snapshot) must not index past the cell table"""
    N = 16
    pos = cases.particles(N, 4096, seed=2)
    pos[0] = (0.5, 1.0, 0.25)
    pos[1] = (0.25, 0.5, 1.0)
    tp = _cuda(pos)
    bn = psc.mesh.bin_particles(tp, N)
    ncells = N ** 3
    cell = bn.scratch[: 4 * (ncells + 2)].cpu().numpy().view(np.int32)
    assert cell[ncells] == len(pos) and np.all(np.diff(cell[: ncells + 1]) >= 0)


def test_small_mesh_path_without_binning(psc, orc):
    """N % 8 != 0: mesh.can_bin is false, the step runs on psc_deposit (global REDs) and psc_interp_kick4 (direct
    gather) -- the documented path of meshes the cell-sorted kernels do not take"""
    N = 12
    pos = cases.particles(N, 3000, seed=17)
    assert not psc.mesh.can_bin(N, len(pos))
    for scheme in ("TSC", "CIC"):
        ref = getattr(orc.mesh, "TSC_seq" if scheme == "TSC" else scheme)(pos, N)
        assert_close(getattr(psc.mesh, scheme)(pos, N), ref, TOL, f"{scheme} N=12")
    phi = cases.scalar_grid(N, seed=3, smooth=True)
    force4 = psc.mesh.derivative(_cuda(phi), 5, padded=True)
    a_ref = orc.mesh.invTSC_vec(orc.mesh.derivative(phi, 5), pos)
    vel = cases.velocities(len(pos), seed=4, scale=1e-2)
    tv = _cuda(vel)
    a, mx = psc.mesh.interp_kick(force4, _cuda(pos), tv, 2, np.float32(0.02))
    assert_close(a.cpu().numpy(), a_ref, TOL, "interp_kick4 N=12")
    v_ref = vel.copy()
    orc.utils.add_vector_scalar_inplace(v_ref, a_ref, -np.float32(0.02))
    assert_close(tv.cpu().numpy(), v_ref, TOL, "kick N=12")
