"""GPU tests of the binned particle <-> mesh kernels (csrc/binned.cu) through the C ABI: the invariants of the per-step
binning (8^3-cell bins), the deposit and the gradient + interpolation + kick kernel against the oracle on uniform,
clustered and mixed particle sets (bins of all fills, including bins split into parts), positions on the box edge,
and the path of meshes that cannot be binned (N % 8 != 0).

Tolerance: max|diff| <= 5e-6 rms(reference) for float32 fields; the binning (bin of every record, permutation) is exact."""
import os

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


def _layout(bn, n, N):
    """(fill, base, records) of a Binned scratch: counts | fill | base | tmp (nbins + 1 ints each, 256-byte aligned),
    then the records (csrc/binned.cu: bin_layout)"""
    nbins = (N // 8) ** 3
    raw = bn.scratch.cpu().numpy()
    o = _a256(4 * (nbins + 1))
    fill = raw[o: o + 4 * (nbins + 1)].view(np.int32)
    base = raw[2 * o: 2 * o + 4 * (nbins + 1)].view(np.int32)
    nrec = n + n // 8 + 40 * nbins + 64
    rec = raw[4 * o: 4 * o + 16 * nrec].view(np.float32).reshape(nrec, 4)
    return fill, base, rec


def _check_binning(bn, pos, N, exact):
    """every bin's records are exactly the particles of that bin (positions bit-exact, rows a permutation)"""
    n = len(pos)
    nbins = (N // 8) ** 3
    fill, base, rec = _layout(bn, n, N)
    key = _bin_key(pos, N)
    counts = np.bincount(key, minlength=nbins)
    assert np.array_equal(fill[:nbins], counts)
    assert np.all(np.diff(base.astype(np.int64)) >= counts if not exact else np.diff(base.astype(np.int64)) == counts)
    assert base[0] == 0
    sel = np.concatenate([np.arange(base[b], base[b] + fill[b]) for b in np.nonzero(counts)[0]]) if n else np.zeros(0, int)
    rows = rec[sel, 3].copy().view(np.int32)
    assert np.array_equal(np.sort(rows), np.arange(n))
    assert np.array_equal(rec[sel, :3], pos[rows])
    assert np.array_equal(key[rows], np.repeat(np.nonzero(counts)[0], counts[np.nonzero(counts)[0]]))


def _bin_key(pos, N):
    """(bi * NB + bj) * NB + bk: the bin of binned.cu (bin_of)"""
    c = np.minimum((pos * np.float32(N)).astype(np.int64), N - 1)
    NB = N // 8
    return ((c[:, 0] >> 3) * NB + (c[:, 1] >> 3)) * NB + (c[:, 2] >> 3)


def _sort_report(psc, N, p_in, v_in, a_in, id_in, half, dt, f64, p_out, v_out, id_out, sb, what):
    """None when (p_out, v_out, id_out) is psc_kick_drift_wrap of the inputs in bin order with the ids carried along
    and the bin table of `sb` describes it; otherwise a description of what is wrong (also appended to
    gpurun_out/sort_failures.txt: a rare failure of this check is worth its details)"""
    import torch
    lib, L = psc._lib, psc._lib.load()
    n, nbins = p_in.shape[0], (N // 8) ** 3
    rp, rv = p_in.clone(), v_in.clone()
    lib.check(L.psc_kick_drift_wrap(lib.ptr(rp), lib.ptr(rv), lib.ptr(a_in), n, float(half), float(dt), f64, lib.stream()))
    torch.cuda.synchronize()
    rp, rv, po, vo, io = (t.cpu().numpy() for t in (rp, rv, p_out, v_out, id_out))
    idi = np.arange(n) if id_in is None else id_in.cpu().numpy()
    k = _bin_key(po, N)
    msg = []
    uniq, cnt = np.unique(io, return_counts=True)
    in_range = (io >= 0) & (io < n)
    if len(uniq) != n or not in_range.all():
        msg.append(f"ids: {n - len(uniq)} rows repeat an id, {int((~in_range).sum())} ids out of range")
    row_of_id = np.empty(n, np.int64)
    row_of_id[idi] = np.arange(n)
    src = row_of_id[np.clip(io, 0, n - 1)]                       # input row of every output row
    badp = np.nonzero(np.any(po != rp[src], axis=1) | np.any(vo != rv[src], axis=1))[0]
    if len(badp):
        pin = p_in.cpu().numpy()
        stale = int(np.sum(np.all(po[badp] == pin[src[badp]], axis=1)))
        zero = int(np.sum(np.all(po[badp] == 0, axis=1)))
        msg.append(f"{len(badp)} rows differ from kick_drift_wrap of their id (rows {badp[0]}..{badp[-1]}; {stale} hold "
                   f"the input position, {zero} are zero); first: " +
                   "; ".join(f"row {r} id {io[r]} bin {k[r]} got {po[r]} want {rp[src[r]]}" for r in badp[:4]))
    ooo = np.nonzero(np.diff(k) < 0)[0]
    if len(ooo):
        msg.append(f"{len(ooo)} descents of the bin key (rows {ooo[0]}..{ooo[-1]}); first: " +
                   "; ".join(f"row {r}: keys {k[max(r - 2, 0): r + 4].tolist()}" for r in ooo[:4]))
    raw = sb.scratch.cpu().numpy()
    o = _a256(4 * (nbins + 1))
    of, ob = (o, 2 * o) if sb.table == 0 else (4 * o + 256, 5 * o + 256)     # csrc/binned.cu bin_layout
    fill = raw[of: of + 4 * nbins].view(np.int32)
    base = raw[ob: ob + 4 * (nbins + 1)].view(np.int32)
    want = np.bincount(_bin_key(rp, N), minlength=nbins)
    if not (np.array_equal(fill, want) and np.array_equal(np.diff(base), want) and base[0] == 0):
        bf, bb = np.nonzero(fill != want)[0], np.nonzero(np.diff(base) != want)[0]
        msg.append(f"table {sb.table}: fill differs in {len(bf)} bins {bf[:6].tolist()}, base in {len(bb)} bins "
                   f"{bb[:6].tolist()} (fill sum {int(fill.sum())}, base[-1] {int(base[-1])}, n {n})")
    if not msg:
        return None
    text = f"{what} (N {N}, n {n}, dt {dt!r}, table {sb.table}): " + " | ".join(msg)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "sort_failures.txt"), "a") as fh:
            fh.write(text + "\n")
    except OSError:
        pass
    return text


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
    """base[] is the exclusive scan of the per-bin counts, fill[] the counts, the records of a bin are exactly its
    particles (a permutation of the rows, positions bit-exact)"""
    N = 64
    pos = cases.lattice_particles(N, 0.3, seed=5) if kind == "lattice" else _mixed_particles(N)
    bn = psc.mesh.bin_particles(_cuda(pos), N)
    _check_binning(bn, pos, N, exact=True)


def _kdw_reference(orc, pos, vel, acc, dt):
    p, v = pos.copy(), vel.copy()
    orc.utils.add_vector_scalar_inplace(v, acc, -np.float32(0.5 * dt))
    orc.utils.add_vector_scalar_inplace(p, v, dt)
    orc.utils.periodic_wrap(p)
    return p, v


def test_kick_drift_count_then_direct_scatter(psc, orc):
    """step 1: the count pass folded into kick+drift+wrap (mode 0) gives the binning psc_bin_particles gives; step 2 on
    the same scratch: direct scatter (mode 1) -- no count pass, the records land in bins sized from step 1's fill;
    step 3: direct scatter in two chunks (row0), as the pinned-host pipeline calls it"""
    import torch
    N = 32
    n = 50003   # not a multiple of 4: the scalar tail of the fused kernel
    pos, vel = cases.particles(N, n, seed=21), cases.velocities(n, seed=22, scale=5e-3)
    acc = cases.velocities(n, seed=23, scale=1.0)
    tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
    bn = psc.mesh.alloc_binned(n, N)
    dt = np.float32(0.21)
    for step in range(3):
        assert bn.ready == (step > 0)
        if step < 2:
            psc.mesh.kick_drift_wrap_count(tp, tv, ta, np.float32(0.5 * dt), dt, 0, bn)
        else:
            h = 20000   # a multiple of 4 keeps the second chunk 16-byte aligned
            psc.mesh.kick_drift_wrap_count(tp[:h], tv[:h], ta[:h], np.float32(0.5 * dt), dt, 0, bn, zero_counts=True)
            psc.mesh.kick_drift_wrap_count(tp[h:], tv[h:], ta[h:], np.float32(0.5 * dt), dt, 0, bn, zero_counts=False,
                                           row0=h)
        assert bn.mode == (1 if step > 0 else 0)
        psc.mesh.finish_binning(tp, bn)
        pos, vel = _kdw_reference(orc, pos, vel, acc, dt)
        got = tp.cpu().numpy()
        assert np.max(np.abs(got - pos)) <= 1.2e-7 * (step + 1)     # one ulp below 1.0 per step (fused multiply-add)
        pos, vel = got, tv.cpu().numpy()                            # follow the device's bits from here on
        _check_binning(bn, pos, N, exact=(step == 0))
        rho = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, bn)
        assert_close(rho.cpu().numpy(), orc.mesh.TSC_seq(pos, N), TOL, f"deposit, step {step}")
    torch.cuda.synchronize()


def test_direct_scatter_overflow_falls_back_to_exact_binning(psc, orc):
    """the particles move far more than the slack of the bins allows (here: everything collapses into a corner
    between two steps): the direct scatter overflows, the device redoes the exact binning, nothing is lost"""
    N = 32
    n = 40000
    pos = cases.particles(N, n, seed=31)
    vel = np.zeros((n, 3), dtype=np.float32)
    acc = np.zeros((n, 3), dtype=np.float32)
    tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
    bn = psc.mesh.alloc_binned(n, N)
    psc.mesh.kick_drift_wrap_count(tp, tv, ta, np.float32(0), np.float32(0), 0, bn)
    psc.mesh.finish_binning(tp, bn)
    _check_binning(bn, pos, N, exact=True)
    tp.mul_(0.2)                                            # all particles into 1/125 of the box
    pos2 = tp.cpu().numpy()
    psc.mesh.kick_drift_wrap_count(tp, tv, ta, np.float32(0), np.float32(0), 0, bn)   # mode 1, must overflow
    assert bn.mode == 1
    psc.mesh.finish_binning(tp, bn)
    nbins = (N // 8) ** 3
    o = _a256(4 * (nbins + 1))
    flag_off = 4 * o + _a256(16 * (n + n // 8 + 40 * nbins + 64))
    assert bn.scratch[flag_off: flag_off + 4].cpu().numpy().view(np.int32)[0] == 1, "the test must overflow"
    _check_binning(bn, pos2, N, exact=True)
    rho = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, bn)
    exact = orc.mesh.deposit_f64(pos2, N, 2)
    assert_close(rho.cpu().numpy(), exact, max(TOL, 3 * rel_err(orc.mesh.TSC_seq(pos2, N), exact)), "deposit after overflow")
    # and the next step is a direct scatter again, now with room for the collapsed distribution
    psc.mesh.kick_drift_wrap_count(tp, tv, ta, np.float32(0), np.float32(0), 0, bn)
    psc.mesh.finish_binning(tp, bn)
    assert bn.scratch[flag_off: flag_off + 4].cpu().numpy().view(np.int32)[0] == 0
    _check_binning(bn, pos2, N, exact=False)


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
    fill, base, _ = _layout(bn, len(pos), N)
    assert base[nbins] == len(pos) and np.all(np.diff(base) >= 0) and fill[:nbins].sum() == len(pos)
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


@pytest.mark.parametrize("N", [32, 128])
def test_bin_ordered_loop_matches_row_preserving_loop(psc, N):
    """integration.integrate on device tensors keeps the arrays in bin order (default) or leaves every particle in
    its row (param['particle_order'] = 'reference', the shadow binning): same physics -- after utils.reference_order
    the two agree to float32 summation order; NumPy callers always get their rows back"""
    import torch
    tables = cases.toy_tables()
    pos = cases.lattice_particles(N, 0.3, seed=60)
    vel = cases.velocities(N ** 3, seed=61, scale=2e-3)
    out = {}
    for mode in ("bins", "reference", "numpy"):
        param = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver="fft")
        param["aexp"] = 0.2
        param["t"] = float(tables[1](np.log(param["aexp"])))
        param["particle_order"] = "reference" if mode == "reference" else "bins"
        psc.utils.set_units(param)
        p, v = (pos.copy(), vel.copy()) if mode == "numpy" else (_cuda(pos), _cuda(vel))
        a, pot, add = psc.solver.pm(p, param)
        for step in range(4):
            param["nsteps"] += 1
            p, v, a, pot, add = psc.integration.integrate(p, v, a, pot, add, tables, param, 1e30)
            if mode == "bins":
                assert psc.utils.particle_ids(p) is not None
                k = _bin_key(p.cpu().numpy(), N)
                assert np.all(np.diff(k) >= 0), "arrays of the bin-ordered loop are sorted by bin"
            elif mode == "reference":
                assert psc.utils.particle_ids(p) is None
            if step == 1:
                # utils.py:1019-1075, as main.run does every n_reorder steps: physically Morton-sorted arrays for the
                # row-preserving loops; the bin-ordered arrays only get new ids (the same tensors come back)
                q = psc.utils.reorder_particles(p, v, a)
                if mode == "bins":
                    assert q[0] is p and q[1] is v and q[2] is a
                    ids = psc.utils.particle_ids(p).cpu().numpy()
                    assert np.array_equal(np.sort(ids), np.arange(N ** 3))
                    keys = psc.morton.positions_to_keys(p).cpu().numpy()
                    assert np.all(np.diff(keys[np.argsort(ids)]) >= 0), "ids = rank in Morton-key order"
                p, v, a = q
        if mode != "numpy":
            p, v, a = psc.utils.reference_order(p, v, a)
            p, v, a = (t.cpu().numpy() for t in (p, v, a))
        out[mode] = (p, v, a)
    for other in ("reference", "numpy"):
        d = np.abs(out["bins"][0] - out[other][0])
        assert np.minimum(d, 1 - d).max() < 2e-7, other
        assert_close(out["bins"][1], out[other][1], 1e-5, f"velocity vs {other}")
        # float32 summation order of the deposit differs between the layouts; 2e-4 is the bar after steps
        assert_close(out["bins"][2], out[other][2], 5e-5 if N == 32 else 2e-4, f"acceleration vs {other}")


def test_morton_relabel_of_bin_ordered_arrays(psc):
    """utils.reorder_particles on the bin-ordered arrays of the time loop: ids = rank of the Morton key (per-bin sort in
    shared memory: bins of <= 768 and of 769..2048 particles), against numpy's argsort of the same keys; a bin beyond
    2048 particles falls back to the global sort (new arrays in Morton order, no ids)"""
    import torch
    N = 32
    rng = np.random.default_rng(77)
    base = cases.particles(N, 60000, seed=78)
    mid = (rng.random((1500, 3), dtype=np.float32) * np.float32(8 / N) + np.float32(0.25)).astype(np.float32)
    for extra in (mid, np.concatenate([mid, mid[:900] * np.float32(0.999)])):      # a bin of ~1600, then of ~2500
        pos = np.ascontiguousarray(np.concatenate([base, extra]))
        n = len(pos)
        vel = cases.velocities(n, seed=79, scale=1e-3)
        tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(np.zeros_like(vel))
        sb = psc.mesh.step_sorted(n, N)
        sp, sv, sid = psc.mesh.step_sort(tp, tv, ta, None, np.float32(0), np.float32(0), 0, sb)
        psc.utils.set_particle_ids((sp, sv), sid)
        fill_max = np.bincount(_bin_key(sp.cpu().numpy(), N)).max()
        q = psc.utils.reorder_particles(sp, sv)
        keys = psc.morton.positions_to_keys(sp).cpu().numpy()
        want = pos[np.argsort(psc.morton.positions_to_keys(tp).cpu().numpy(), kind="stable")]
        if fill_max <= 2048:
            assert 768 < fill_max and q[0] is sp
            ids = psc.utils.particle_ids(sp).cpu().numpy()
            assert np.array_equal(np.sort(ids), np.arange(n))
            assert np.all(np.diff(keys[np.argsort(ids)]) >= 0)
        else:
            assert q[0] is not sp and psc.utils.particle_ids(q[0]) is None
        got = psc.utils.reference_order(*q)[0].cpu().numpy()
        assert np.array_equal(np.sort(keys), psc.morton.positions_to_keys(_cuda(got)).cpu().numpy())
        # distinct keys: the Morton-sorted arrays are unique
        if len(np.unique(keys)) == n:
            assert np.array_equal(got, want)
