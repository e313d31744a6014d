"""GPU tests of psc_step_sort (csrc/binned.cu: the kick + drift + wrap fused with the re-sort of the particle arrays
into bin order, and the speculative count pass of the interpolation kernel), through the C ABI.

They live at the END of the suite on purpose (open item): twice in about 45 runs one sort of the
N = 128 cases left rows out of bin order, cause not proven removed; the driver runs the suite with -x, and a rare
failure here must not hide the rest.  The checks are strict (no retry); a failing sort is described in detail
(_sort_report, also appended to gpurun_out/sort_failures.txt)."""
import os

import numpy as np
import pytest

import cases
from conftest import assert_close
from test_binned_gpu import TOL, _a256, _bin_key, _cuda, _sort_report, orc, psc  # noqa: F401  (psc, orc: fixtures)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,n", [(32, 50003), (128, 600011)])
def test_step_sort_is_kick_drift_wrap_plus_a_permutation(psc, orc, N, n):
    """psc_step_sort (first half of the leapfrog fused with the re-sort into bin order): the output arrays hold exactly
    the particles psc_kick_drift_wrap produces, in bin order, with ids = their input rows; the following calls start
    from bin-ordered arrays (one CTA per source bin, shared-memory sort, tables 0 -> 1 -> 0 ...) and carry the ids
    through: float64 drift (snapshot-clamped dt), drifts that move many particles into neighbouring bins and a drift of
    a third of the box (the slow path of the local sort); PSC_NO_LOCAL_SORT=1 keeps the global-atomic sort"""
    # N = 128: 4096 bins for 592 persistent CTAs, every CTA walks over several bins (the TMA pipeline of the local sort)
    import torch
    pos, vel = cases.particles(N, n, seed=21), cases.velocities(n, seed=22, scale=5e-3)
    acc = cases.velocities(n, seed=23, scale=1e-4)
    lib, L = psc._lib, psc._lib.load()
    tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
    ids, prev_table = None, 0
    nbins = (N // 8) ** 3
    for step, dt in enumerate((np.float32(0.021), 0.0193456789012, np.float32(4.0), np.float32(30.0), 2.5,
                               np.float32(0.3))):
        half = np.float32(0.5 * dt)
        f64 = 0 if isinstance(dt, np.float32) else 1
        rp, rv = tp.clone(), tv.clone()
        lib.check(L.psc_kick_drift_wrap(lib.ptr(rp), lib.ptr(rv), lib.ptr(ta), n, float(half), float(dt), f64, lib.stream()))
        sb = psc.mesh.step_sorted(n, N)
        if step == 5:
            os.environ["PSC_NO_LOCAL_SORT"] = "1"
        try:
            sp, sv, sid = psc.mesh.step_sort(tp, tv, ta, ids, half, dt, f64, sb)
        finally:
            os.environ.pop("PSC_NO_LOCAL_SORT", None)
        torch.cuda.synchronize()
        assert sb.table == (0 if step in (0, 5) else 1 - prev_table) and sb.describes(sp)
        report = _sort_report(psc, N, tp, tv, ta, ids, half, dt, f64, sp, sv, sid, sb, f"step {step}")
        assert report is None, report
        prev_table = sb.table
        h_id = sid.cpu().numpy()
        assert np.array_equal(np.sort(h_id), np.arange(n))
        src = h_id if ids is None else np.argsort(ids.cpu().numpy())[h_id]     # input row of every output row
        assert np.array_equal(sp.cpu().numpy(), rp.cpu().numpy()[src])         # same arithmetic: bit-identical
        assert np.array_equal(sv.cpu().numpy(), rv.cpu().numpy()[src])
        key = _bin_key(sp.cpu().numpy(), N)
        assert np.all(np.diff(key) >= 0)
        raw = sb.scratch.cpu().numpy()
        o = _a256(4 * (nbins + 1))
        of, ob = (o, 2 * o) if sb.table == 0 else (4 * o + 256, 5 * o + 256)     # csrc/binned.cu bin_layout
        fill = raw[of: of + 4 * nbins].view(np.int32)
        base = raw[ob: ob + 4 * (nbins + 1)].view(np.int32)
        counts = np.bincount(key, minlength=nbins)
        assert np.array_equal(fill, counts) and np.array_equal(np.diff(base), counts) and base[0] == 0
        # deposit / interpolation on the sorted arrays against the oracle
        rho = psc.mesh.deposit_rhs(sp, N, psc._lib.TSC, 1.0, 1.0, 0.0, sb)
        assert_close(rho.cpu().numpy(), orc.mesh.TSC_seq(sp.cpu().numpy(), N), TOL, "deposit on sorted arrays")
        phi = cases.scalar_grid(N, seed=31, smooth=True)
        a_ref = orc.mesh.invTSC_vec(orc.mesh.derivative(phi, 5), sp.cpu().numpy())
        v0 = sv.clone()
        a, mx = psc.mesh.interp_kick_phi(_cuda(phi), None, 0.0, 0, 5, sp, sv, 2, np.float32(0.013), sb)
        assert_close(a.cpu().numpy(), a_ref, 2 * TOL, "interpolation on sorted arrays")
        v_ref = v0.cpu().numpy().copy()
        orc.utils.add_vector_scalar_inplace(v_ref, a_ref, -np.float32(0.013))
        assert_close(sv.cpu().numpy(), v_ref, 2 * TOL, "kick on sorted arrays")
        # next step starts from the sorted state: positions / velocities / ids of this step, the same acceleration rows
        ta = ta[torch.from_numpy(src).cuda()].contiguous()
        tp, tv, ids = sp, v0, sid


@pytest.mark.parametrize("dt2", [np.float32(3.0), 2.7182818284])
def test_predicted_bin_count_equals_the_count_pass(psc, orc, dt2):
    """psc_interp_kick_phi_sorted(predict = 1) counts, in the interpolation kernel, the bins the NEXT psc_step_sort will
    fill (speculative count pass): the count table must equal the bins of psc_kick_drift_wrap's new positions, and the
    sort that skips its count pass must produce exactly what the counting sort does (float32 and float64 time steps)"""
    import torch
    N, n = 128, 500009
    pos, vel = cases.particles(N, n, seed=41), cases.velocities(n, seed=42, scale=5e-3)
    lib, L = psc._lib, psc._lib.load()
    sb = psc.mesh.step_sorted(n, N)
    zero = torch.zeros((n, 3), device="cuda")
    sp, sv, sid = psc.mesh.step_sort(_cuda(pos), _cuda(vel), zero, None, np.float32(0), np.float32(0), 0, sb)
    phi = _cuda(cases.scalar_grid(N, seed=43, smooth=True) * np.float32(2e-3))
    half2 = np.float32(0.5 * dt2)
    f64 = 0 if isinstance(dt2, np.float32) else 1
    sb.predict_next = (half2, dt2, f64)
    acc, _ = psc.mesh.interp_kick_phi(phi, None, 0.0, 0, 5, sp, sv, 2, np.float32(0.01), sb)
    assert sb.predicted[:3] == (float(half2), float(dt2), f64)
    # reference: the stand-alone kick + drift + wrap on copies
    rp, rv = sp.clone(), sv.clone()
    lib.check(L.psc_kick_drift_wrap(lib.ptr(rp), lib.ptr(rv), lib.ptr(acc), n, float(half2), float(dt2), f64, lib.stream()))
    torch.cuda.synchronize()
    nbins = (N // 8) ** 3
    want = np.bincount(_bin_key(rp.cpu().numpy(), N), minlength=nbins)
    got = sb.scratch.cpu().numpy()[: 4 * nbins].view(np.int32)          # the count table opens the scratch
    assert np.array_equal(got, want)
    assert np.any(want != np.bincount(_bin_key(sp.cpu().numpy(), N), minlength=nbins)), "particles must change bins"
    skipped = sb.counts_skipped
    p2, v2, i2 = psc.mesh.step_sort(sp, sv, acc, sid, half2, dt2, f64, sb)
    assert sb.counts_skipped == skipped + 1
    torch.cuda.synchronize()
    order = np.argsort(sid.cpu().numpy())[i2.cpu().numpy()]            # input row of every output row
    assert np.array_equal(p2.cpu().numpy(), rp.cpu().numpy()[order])
    assert np.array_equal(v2.cpu().numpy(), rv.cpu().numpy()[order])
    key = _bin_key(p2.cpu().numpy(), N)
    assert np.all(np.diff(key) >= 0)
    # another time step than the predicted one: the sort counts for itself
    sb.predict_next = (half2, dt2, f64)
    acc2, _ = psc.mesh.interp_kick_phi(phi, None, 0.0, 0, 5, p2, v2, 2, np.float32(0.01), sb)
    p3, v3, i3 = psc.mesh.step_sort(p2, v2, acc2, i2, np.float32(0.25), np.float32(0.5), 0, sb)
    assert sb.counts_skipped == skipped + 1
    report = _sort_report(psc, N, p2, v2, acc2, i2, np.float32(0.25), np.float32(0.5), 0, p3, v3, i3, sb, "other step")
    assert report is None, report
    # the predicted step, but the caller touched the velocities in between: the guess is void
    sb.predict_next = (half2, dt2, f64)
    acc3, _ = psc.mesh.interp_kick_phi(phi, None, 0.0, 0, 5, p3, v3, 2, np.float32(0.01), sb)
    v3 += 1e-4
    p4, v4, i4 = psc.mesh.step_sort(p3, v3, acc3, i3, half2, dt2, f64, sb)
    assert sb.counts_skipped == skipped + 1
    report = _sort_report(psc, N, p3, v3, acc3, i3, half2, dt2, f64, p4, v4, i4, sb, "voided guess")
    assert report is None, report
