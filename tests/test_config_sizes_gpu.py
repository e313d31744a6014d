"""BASELINE.json configs 1-4 at their REAL sizes, CUDA path against the oracle (oracle.host.pm / integrate) on the same
seeded inputs -- direct comparisons, not properties ("Parity gaps").

  config 1: Newtonian 128^3, FFT, TSC: solver.pm + three leapfrog steps (one clamped to a snapshot, one reorder)
  config 2: Newtonian 256^3, multigrid (V-cycles, red-black Gauss-Seidel)
  config 3: f(R) Hu-Sawicki n = 1, |fR0| = 1e-5, 256^3, FAS multigrid with the cubic smoother.  Screened regime
            (a = 0.05): at low redshift the reference's cubic root leaves its real branch and the reference itself
            stops (cubic.py:196-197), so that is where parity is defined.
  config 4: QUMOND 512^3, fft_7pt Newtonian solve + nu-weighted source + second solve
  + integration.euler against the reference's golden vectors (tests/golden/euler.npz).

Tolerances: max|diff| <= tol * rms(reference); potential 3e-5 (FFT) / 1e-4 (multigrid: both
sides stop at the same cycle count, the iterates differ by float32 summation order), acceleration 1e-4 (2e-4 after
steps), positions 1e-6 box units."""
import numpy as np
import pytest

import cases
from conftest import assert_close
from test_oracle_golden import _euler_steps, check_euler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def psc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pysco_b200
    pysco_b200._lib.load()
    return pysco_b200


@pytest.fixture(scope="module")
def host():
    import oracle
    oracle.build()
    oracle.set_num_threads(__import__("os").cpu_count() or 1)
    from oracle import host
    return host


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def _pair(ncoarse, npart, psc, host, **over):
    p1 = cases.base_param(ncoarse, npart, **over)
    p2 = p1.copy()
    psc.utils.set_units(p1)
    host.set_units(p2)
    return p1, p2


def test_euler_steps_vs_golden(psc, golden):
    pos, vel, acc, pot, dts, param = _euler_steps(psc.solver.pm, psc.integration.integrate, psc.utils.set_units, _cuda)
    check_euler(golden("euler"), _np(pos), _np(vel), _np(acc), _np(pot), dts, param, 2e-4)


def test_config1_newton_fft_128_pm_and_steps(psc, host):
    import oracle
    N = 128
    tables = cases.toy_tables()
    pos = cases.lattice_particles(N, 0.3, seed=71)
    vel = cases.velocities(N ** 3, seed=72, scale=2e-3)
    p1, p2 = _pair(7, N ** 3, psc, host, linear_newton_solver="fft")
    for p in (p1, p2):
        p["aexp"] = 0.2
        p["t"] = float(tables[1](np.log(p["aexp"])))
    psc.utils.set_units(p1)
    host.set_units(p2)
    s1 = [_cuda(pos), _cuda(vel)] + list(psc.solver.pm(_cuda(pos), p1))
    s2 = [pos.copy(), vel.copy()] + list(host.pm(pos.copy(), p2))
    assert_close(_np(s1[3]), s2[3], 3e-5, "config 1 potential")
    assert_close(_np(s1[2]), s2[2], 1e-4, "config 1 acceleration")
    for step in range(3):
        t_snap = 1e30 if step < 2 else p2["t"] + 0.4 * dt
        t0 = p2["t"]
        for p in (p1, p2):
            p["nsteps"] += 1
        s1 = list(psc.integration.integrate(*s1, tables, p1, t_snap))
        s2 = list(host.integrate(*s2, tables, p2, t_snap))
        dt = p2["t"] - t0
        np.testing.assert_allclose(p1["t"], p2["t"], rtol=1e-6)
        if step == 1:
            s1[0], s1[1], s1[2] = psc.utils.reorder_particles(s1[0], s1[1], s1[2])
            s2[0], s2[1], s2[2] = oracle.utils.reorder_particles(s2[0], s2[1], s2[2])
    assert bool(p1["write_snapshot"]) and bool(p2["write_snapshot"])
    assert np.max(np.abs(_np(s1[0]) - s2[0])) < 1e-6, "positions (row by row: same particle order)"
    assert_close(_np(s1[1]), s2[1], 2e-4, "velocity after 3 steps")
    assert_close(_np(s1[2]), s2[2], 2e-4, "acceleration after 3 steps")
    assert_close(_np(s1[3]), s2[3], 2e-4, "potential after 3 steps")


def test_config2_newton_multigrid_256_pm(psc, host):
    N = 256
    pos = cases.lattice_particles(N, 0.3, seed=81)
    p1, p2 = _pair(8, N ** 3, psc, host, linear_newton_solver="multigrid")
    acc, pot, _ = psc.solver.pm(_cuda(pos), p1)
    acc_ref, pot_ref, _ = host.pm(pos, p2)
    np.testing.assert_allclose(p1["tolerance"], p2["tolerance"], rtol=1e-3)
    assert_close(_np(pot), pot_ref, 1e-4, "config 2 potential")
    assert_close(_np(acc), acc_ref, 2e-4, "config 2 acceleration")


def test_config3_fr_256_pm(psc, host):
    N = 256
    pos = cases.lattice_particles(N, 0.3, seed=82)
    over = dict(theory="fr", fR_n=1, fR_logfR0=5, linear_newton_solver="multigrid", aexp=0.05, aexp_old=0.05)
    p1, p2 = _pair(8, N ** 3, psc, host, **over)
    tables = cases.toy_tables()
    acc, pot, u = psc.solver.pm(_cuda(pos), p1, tables=tables)
    acc_ref, pot_ref, u_ref = host.pm(pos, p2, tables=tables)
    assert_close(_np(u), u_ref, 1e-4, "config 3 scalaron")
    assert_close(_np(pot), pot_ref, 1e-4, "config 3 potential")
    assert_close(_np(acc), acc_ref, 2e-4, "config 3 acceleration")


def test_config4_mond_512_pm(psc, host):
    N = 512
    pos = cases.lattice_particles(N, 0.3, seed=83)
    over = dict(theory="mond", linear_newton_solver="fft_7pt", mond_function="simple", mond_g0=1.2)
    p1, p2 = _pair(9, N ** 3, psc, host, **over)
    tp = _cuda(pos)
    acc, pot, add = psc.solver.pm(tp, p1)
    acc, pot, add = _np(acc), _np(pot), _np(add)
    del tp
    acc_ref, pot_ref, add_ref = host.pm(pos, p2)
    assert_close(add, add_ref, 3e-5, "config 4 Newtonian potential")
    assert_close(pot, pot_ref, 1e-4, "config 4 MOND potential")
    assert_close(acc, acc_ref, 2e-4, "config 4 acceleration")
