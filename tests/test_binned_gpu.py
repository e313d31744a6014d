"""GPU tests of the binned particle <-> mesh kernels (csrc/binned.cu) through the C ABI: the invariants of the per-step
binning (8^3-cell bins), the deposit and the gradient + interpolation + kick kernel against the oracle on uniform,
clustered and mixed particle sets (bins of all fills, including bins split into parts), positions on the box edge,
and the path of meshes that cannot be binned (N % 8 != 0).

Tolerance: max|diff| <= 5e-6 rms(reference) for float32 fields; the binning (bin of every record, permutation) is exact."""
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
    return pysco_b200


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


def _bin_key(pos, N):
    """(bi * NB + bj) * NB + bk: the bin of binned.cu (bin_of)"""
    c = np.minimum((pos * np.float32(N)).astype(np.int64), N - 1)
    NB = N // 8
    return ((c[:, 0] >> 3) * NB + (c[:, 1] >> 3)) * NB + (c[:, 2] >> 3)


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
def test_binning_invariants(psc, kind):
    """offsets[] is the exclusive scan of the per-bin counts, the records are a permutation of the rows, every record
    sits in the range of its own bin and carries the position of its source row (bit-exact)."""
    N = 64
    pos = cases.lattice_particles(N, 0.3, seed=5) if kind == "lattice" else _mixed_particles(N)
    n = len(pos)
    tp = _cuda(pos)
    bn = psc.mesh.bin_particles(tp, N)
    nbins = (N // 8) ** 3
    raw = bn.scratch.cpu().numpy()
    o = _a256(4 * (nbins + 1))
    offsets = raw[o: o + 4 * (nbins + 1)].view(np.int32)
    rec = raw[2 * o: 2 * o + 16 * n].view(np.float32).reshape(n, 4)
    rows = rec[:, 3].copy().view(np.int32)
    key = _bin_key(pos, N)
    counts = np.bincount(key, minlength=nbins)
    assert offsets[0] == 0 and offsets[nbins] == n
    assert np.array_equal(np.diff(offsets.astype(np.int64)), counts)
    assert np.array_equal(np.sort(rows), np.arange(n))
    assert np.array_equal(rec[:, :3], pos[rows])
    assert np.all(np.diff(key[rows]) >= 0)


def test_kick_drift_count_matches_separate_binning(psc, orc):
    """the count pass folded into kick+drift+wrap produces the same binning as psc_bin_particles of the new positions"""
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
    assert np.max(np.abs(tp.cpu().numpy() - p)) <= 1.2e-7      # one ulp below 1.0 (fused multiply-add in the drift)
    ref = psc.mesh.bin_particles(tp, N)
    nbins = (N // 8) ** 3
    o = _a256(4 * (nbins + 1))
    a = bn.scratch[o: o + 4 * (nbins + 1)].cpu().numpy().view(np.int32)
    b = ref.scratch[o: o + 4 * (nbins + 1)].cpu().numpy().view(np.int32)
    assert np.array_equal(a, b)
    rho_a = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, bn)
    rho_b = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, ref)
    assert_close(rho_a.cpu().numpy(), rho_b.cpu().numpy(), 1e-6, "deposit from the fused count")
    torch.cuda.synchronize()


@pytest.mark.parametrize("scheme", ["TSC", "CIC", "NGP"])
def test_deposit_vs_float64_sum(psc, orc, scheme):
    """yardstick = the float64-accumulated deposit; bound = the error of the reference-ordered float32 sum itself"""
    N = 64
    for name, pos in (("lattice", cases.lattice_particles(N, 0.3, seed=5)), ("mixed", _mixed_particles(N))):
        sid = {"NGP": 0, "CIC": 1, "TSC": 2}[scheme]
        exact = orc.mesh.deposit_f64(pos, N, sid)
        ref = getattr(orc.mesh, "TSC_seq" if scheme == "TSC" else scheme)(pos, N)
        bound = max(TOL, 3.0 * rel_err(ref, exact))
        rho = getattr(psc.mesh, scheme)(pos, N)
        assert_close(rho, exact, bound, f"{scheme} {name}")
        assert abs(float(rho.sum(dtype=np.float64)) - len(pos)) < 2e-6 * len(pos)


@pytest.mark.parametrize("order", [2, 3, 5, 7])
def test_interp_gradient_stage_vs_oracle(psc, orc, order):
    """gradient fused into the interpolation against mesh.derivative + mesh.invTSC_vec / invCIC_vec of the oracle,
    plain and f(R), on a particle set with empty, ordinary and split (> 4096 particles) bins"""
    N = 32
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


def test_positions_on_the_box_edge_do_not_leave_the_grid(psc, orc):
    """y or z exactly 1.0 (an external snapshot) must not index past the cell table"""
    N = 16
    pos = cases.particles(N, 4096, seed=2)
    pos[0] = (0.5, 1.0, 0.25)
    pos[1] = (0.25, 0.5, 1.0)
    tp = _cuda(pos)
    bn = psc.mesh.bin_particles(tp, N)
    nbins = (N // 8) ** 3
    o = _a256(4 * (nbins + 1))
    offsets = bn.scratch[o: o + 4 * (nbins + 1)].cpu().numpy().view(np.int32)
    assert offsets[nbins] == len(pos) and np.all(np.diff(offsets) >= 0)
    rho = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, bn)
    assert abs(float(rho.sum(dtype=__import__("torch").float64)) - len(pos)) < 1e-5 * len(pos)


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
