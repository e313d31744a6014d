"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test drives the CUDA path through the reference-shaped Python API (pysco_b200.*, which calls
the C ABI of libpysco_b200.so) and compares with
  (1) tests/golden/*.npz -- outputs of the unmodified reference on the same seeded inputs, and
  (2) the CPU oracle (oracle/) at sizes it finishes in seconds,
and checks size-independent properties at larger sizes (mass conservation, zero-mean force, ...).

Tolerances: Morton keys / orderings bit-exact; float32 fields max|diff| <= tol * rms(reference).
"""
import os

import numpy as np
import pytest

import cases
from conftest import assert_close

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


# ------------------------------------------------------------------------------ particles
@pytest.mark.parametrize("N,npart", [(16, 5000), (32, 3000)])
def test_particles_mesh_vs_golden(psc, golden, N, npart):
    g = golden("particles")
    pos = cases.particles(N, npart)
    t = f"N{N}"
    keys = psc.morton.positions_to_keys(pos)
    assert keys.dtype == np.int64 and np.array_equal(keys, g[f"keys_{t}"])
    idx = psc.utils.argsort_keys(_cuda(keys)).cpu().numpy()
    assert np.array_equal(idx, g[f"argsort_{t}"])
    assert_close(psc.mesh.TSC(pos, N), g[f"tsc_{t}"], TOL, "TSC")
    assert_close(psc.mesh.TSC_seq(pos, N), g[f"tsc_{t}"], TOL, "TSC_seq")
    assert_close(psc.mesh.CIC(pos, N), g[f"cic_{t}"], TOL, "CIC")
    assert np.array_equal(psc.mesh.NGP(pos, N), g[f"ngp_{t}"])
    f3, f1 = cases.vector_grid(N), cases.scalar_grid(N)
    assert_close(psc.mesh.invTSC_vec(f3, pos), g[f"invtsc_vec_{t}"], TOL, "invTSC_vec")
    assert_close(psc.mesh.invCIC_vec(f3, pos), g[f"invcic_vec_{t}"], TOL, "invCIC_vec")
    assert np.array_equal(psc.mesh.invNGP_vec(f3, pos), g[f"invngp_vec_{t}"])
    assert_close(psc.mesh.invTSC(f1, pos), g[f"invtsc_{t}"], TOL, "invTSC")
    assert_close(psc.mesh.invCIC(f1, pos), g[f"invcic_{t}"], TOL, "invCIC")
    assert np.array_equal(psc.mesh.invNGP(f1, pos), g[f"invngp_{t}"])


def test_reorder_and_vector_ops_vs_golden(psc, golden):
    g = golden("particles")
    pos, vel = cases.particles(16, 5000), cases.velocities(5000)
    acc = cases.velocities(5000, seed=44, scale=1.0)
    p2, v2, a2 = psc.utils.reorder_particles(pos.copy(), vel.copy(), acc.copy())
    assert np.array_equal(p2, g["reorder_pos"])
    k2 = psc.morton.positions_to_keys(p2)
    assert np.all(np.diff(k2) >= 0)
    tie = np.zeros(len(k2), dtype=bool)
    tie[1:] |= k2[1:] == k2[:-1]
    tie[:-1] |= k2[1:] == k2[:-1]
    for mine, ref in ((v2, g["reorder_vel"]), (a2, g["reorder_acc"])):
        assert np.array_equal(mine[~tie], ref[~tie])
        assert np.array_equal(np.sort(mine[tie], axis=0), np.sort(ref[tie], axis=0))
    # stable sort: among equal keys the original order is kept (pos[6] before pos[7])
    assert np.array_equal(v2[tie], vel[[6, 7]])
    y = vel.copy()
    psc.utils.add_vector_scalar_inplace(y, acc, np.float32(-0.0123))
    assert_close(y, g["axpy_f32"], 1e-6, "axpy f32")
    y = pos.copy()
    psc.utils.add_vector_scalar_inplace(y, vel, 0.731)
    assert np.array_equal(y, g["axpy_f64"])
    w = g["wrap_in"].copy()
    psc.utils.periodic_wrap(w)
    assert np.array_equal(w, g["wrap_out"])
    assert psc.utils.max_abs(acc) == g["max_abs"][0]


@pytest.mark.parametrize("scheme", ["TSC", "CIC", "NGP"])
@pytest.mark.parametrize("order", ["lattice", "morton", "random"])
def test_deposit_interp_vs_oracle_1M(psc, orc, scheme, order):
    """128^3 mesh, 128^3 particles in the three orderings the deposit / interpolation kernels see."""
    N = 128
    pos = cases.lattice_particles(N, 0.3, seed=5)
    if order == "morton":
        pos = orc.utils.reorder_particles(pos)
    elif order == "random":
        pos = np.ascontiguousarray(pos[np.random.default_rng(1).permutation(len(pos))])
    rho = getattr(psc.mesh, scheme)(pos, N)
    ref = getattr(orc.mesh, "TSC_seq" if scheme == "TSC" else scheme)(pos, N)
    assert_close(rho, ref, TOL, f"{scheme} deposit ({order})")
    assert abs(float(rho.sum(dtype=np.float64)) - N ** 3) < 1e-5 * N ** 3
    f3 = cases.vector_grid(N, seed=3)
    out = getattr(psc.mesh, f"inv{scheme}_vec")(f3, pos)
    assert_close(out, getattr(orc.mesh, f"inv{scheme}_vec")(f3, pos), TOL, f"inv{scheme}_vec ({order})")


def test_deposit_clustered_and_ragged(psc, orc):
    """Heavy clustering (many particles per cell), Np != N^3, tiny and empty inputs."""
    N = 64
    rng = np.random.default_rng(7)
    blob = (0.5 + 0.01 * rng.standard_normal((200000, 3))).astype(np.float32)
    pos = np.ascontiguousarray(np.concatenate([blob % 1.0, cases.particles(N, 50001, seed=8)]).astype(np.float32))
    pos[pos >= 1.0] = 0.0
    from conftest import rel_err
    for scheme, sid in (("TSC", 2), ("CIC", 1)):
        # thousands of particles per cell: float32 accumulation noise dominates (the reference's own
        # TSC vs TSC_seq differ at this level), so the yardstick is the float64-accumulated sum and
        # the bound is the error of the reference-ordered float32 sum itself
        exact = orc.mesh.deposit_f64(pos, N, sid)
        ref = getattr(orc.mesh, "TSC_seq" if scheme == "TSC" else scheme)(pos, N)
        bound = max(2e-5, 3.0 * rel_err(ref, exact))
        assert_close(getattr(psc.mesh, scheme)(pos, N), exact, bound, f"{scheme} clustered")
        pos_sorted = orc.utils.reorder_particles(pos)
        assert_close(getattr(psc.mesh, scheme)(pos_sorted, N), exact, bound, f"{scheme} clustered+sorted")
    one = np.array([[0.999, 0.001, 0.5]], dtype=np.float32)
    assert_close(psc.mesh.TSC(one, 8), orc.mesh.TSC_seq(one, 8), 1e-6, "single particle")
    empty = np.zeros((0, 3), dtype=np.float32)
    assert float(np.abs(psc.mesh.TSC(empty, 8)).max()) == 0.0
    assert psc.mesh.invTSC_vec(cases.vector_grid(8), empty).shape == (0, 3)


def test_fused_particle_kernels_vs_oracle(psc, orc):
    """kick+drift+wrap and interp+kick+max: fused CUDA kernels vs the reference's separate calls."""
    import torch
    N, npart = 32, 40000
    pos, vel = cases.particles(N, npart, seed=11), cases.velocities(npart, seed=12, scale=5e-3)
    acc = cases.velocities(npart, seed=13, scale=1.0)
    for dt in (np.float32(0.0371), 0.0371234567891):
        p, v = pos.copy(), vel.copy()
        half = np.float32(0.5 * dt)
        orc.utils.add_vector_scalar_inplace(v, acc, -half)
        orc.utils.add_vector_scalar_inplace(p, v, dt)
        orc.utils.periodic_wrap(p)
        tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
        lib = psc._lib
        lib.check(lib.load().psc_kick_drift_wrap(tp.data_ptr(), tv.data_ptr(), ta.data_ptr(), npart, float(half),
                                                 float(dt), 0 if isinstance(dt, np.float32) else 1, lib.stream()))
        assert np.max(np.abs(tp.cpu().numpy() - p)) <= 6e-8
        assert_close(tv.cpu().numpy(), v, 1e-6, "kick")
        assert tp.min().item() >= 0.0 and tp.max().item() < 1.0
    force = cases.vector_grid(N, seed=14)
    for scheme, sid in (("TSC", lib.TSC), ("CIC", lib.CIC)):
        a_ref = getattr(orc.mesh, f"inv{scheme}_vec")(force, pos)
        v_ref = vel.copy()
        orc.utils.add_vector_scalar_inplace(v_ref, a_ref, -np.float32(0.0123))
        tv = _cuda(vel)
        a, mx = psc.mesh.interp_kick(_cuda(force), _cuda(pos), tv, sid, np.float32(0.0123))
        assert_close(a.cpu().numpy(), a_ref, TOL, f"interp_kick acc {scheme}")
        assert_close(tv.cpu().numpy(), v_ref, TOL, f"interp_kick vel {scheme}")
        mx = mx.cpu().numpy()
        np.testing.assert_allclose(mx[0], orc.utils.max_abs(a_ref), rtol=1e-5)
        np.testing.assert_allclose(mx[1], orc.utils.max_abs(v_ref), rtol=1e-5)


# ---------------------------------------------------------------------------------- grids
def test_gradients_and_grid_ops_vs_golden(psc, golden):
    g = golden("grids")
    N = 16
    x = cases.scalar_grid(N, seed=11)
    u = cases.scalaron_field(N)
    for order in (2, 3, 5, 7):
        assert_close(psc.mesh.derivative(x, order), g[f"deriv{order}"], TOL, f"derivative{order}")
        for n in (1, 2):
            assert_close(psc.mesh.derivative_fR(x, u, np.float32(0.37), n, order), g[f"deriv{order}_fR{n}"],
                         TOL, f"derivative{order}_fR_n{n}")
            f = cases.vector_grid(N, seed=12)
            psc.mesh.add_derivative_fR(f, u, np.float32(0.37), n, order)
            ref = g[f"addderiv{order}_fR{n}"]
            if order == 7 and n == 1:  # reference reads out of bounds on the N-3 planes (see oracle test)
                m = np.ones((N, N, N), dtype=bool)
                m[N - 3, :, :] = m[:, N - 3, :] = m[:, :, N - 3] = False
                assert_close(f[m], ref[m], TOL, "add_derivative7_fR_n1")
            else:
                assert_close(f, ref, TOL, f"add_derivative{order}_fR_n{n}")
    with pytest.raises(NotImplementedError):
        psc.mesh.derivative(x, 4)
    with pytest.raises(NotImplementedError):
        psc.mesh.derivative_fR(x, u, np.float32(1), 3, 5)
    y = x.copy(); psc.utils.linear_operator_inplace(y, np.float32(0.193), np.float32(-0.193))
    assert_close(y, g["linop"], 1e-6, "linear_operator_inplace")
    y = x.copy(); psc.utils.prod_vector_scalar_inplace(y, np.float32(1.7))
    assert np.array_equal(y, g["scale"])
    y = x.copy(); psc.utils.linear_operator_vectors_inplace(y, np.float32(4), cases.scalar_grid(N, seed=13), np.float32(1))
    assert_close(y, g["lincomb"], 1e-6, "linear_operator_vectors_inplace")


@pytest.mark.parametrize("N", [16, 32])
def test_fourier_vs_golden(psc, golden, N):
    g = golden("grids")
    r = cases.density_contrast_rhs(N)
    spec_ref = g[f"rfft_N{N}"]
    spec = psc.fourier.fft_3D_real(r, 1)
    assert spec.shape == spec_ref.shape and spec.dtype == np.complex64
    assert_close(spec, spec_ref, 2e-5, "fft_3D_real (cuFFT vs pocketfft)")
    s = spec_ref.copy(); psc.fourier.inverse_laplacian(s)
    assert_close(s, g[f"green_plain_N{N}"], 1e-5, "inverse_laplacian")
    for p in (2, 3):
        s = spec_ref.copy(); psc.fourier.inverse_laplacian_compensated(s, p)
        assert_close(s, g[f"green_comp{p}_N{N}"], 3e-5, f"inverse_laplacian_compensated p={p}")
    s = spec_ref.copy(); psc.fourier.inverse_laplacian_7pt(s)
    assert_close(s, g[f"green_7pt_N{N}"], 1e-5, "inverse_laplacian_7pt")
    assert_close(psc.fourier.ifft_3D_real(g[f"green_7pt_N{N}"], 1), g[f"irfft_7pt_N{N}"], 1e-5, "ifft_3D_real")
    if N == 16:
        assert_close(psc.fourier.gradient_inverse_laplacian(spec_ref.copy()), g[f"gradgreen_N{N}"], 1e-5, "grad green")
        assert_close(psc.fourier.gradient_inverse_laplacian_compensated(spec_ref.copy(), 3),
                     g[f"gradgreen_comp3_N{N}"], 3e-5, "grad green compensated")
        assert_close(psc.fourier.ifft_3D_real_grad(g[f"gradgreen_comp3_N{N}"].copy(), 1), g[f"irfft_grad_N{N}"],
                     1e-5, "ifft_3D_real_grad")
    for p in (0, 2, 3):
        k, pk, nm = psc.fourier.fourier_grid_to_Pk(spec_ref.copy(), p)
        ref = g[f"pk_p{p}_N{N}"]
        assert np.array_equal(nm, ref[2])
        np.testing.assert_allclose(k, ref[0], rtol=1e-6)
        np.testing.assert_allclose(pk, ref[1], rtol=1e-4)   # north_star: P(k) within 1e-4


def test_multigrid_kernels_vs_golden(psc, golden):
    g = golden("multigrid")
    N = 16
    x = cases.scalar_grid(N, seed=21, smooth=True)
    b = cases.density_contrast_rhs(N, seed=22)
    L = psc.laplacian
    assert_close(L.operator(x), g["lap_operator"], 1e-5, "operator")
    assert_close(L.residual(x, b), g["lap_residual"], 1e-5, "residual")
    assert_close(L.restrict_residual(x, b), g["lap_restrict_residual"], 1e-5, "restrict_residual")
    np.testing.assert_allclose(L.residual_error(x, b), g["lap_residual_error"][0], rtol=1e-5)
    np.testing.assert_allclose(L.truncation_error(x), g["lap_truncation_error"][0], rtol=1e-4)
    assert_close(L.initialise_potential(b), g["lap_init"], 1e-6, "initialise_potential")
    y = x.copy(); L.gauss_seidel(y, b, np.float32(1.25))
    assert_close(y, g["lap_gs1"], TOL, "gauss_seidel x1")
    y = x.copy(); L.smoothing(y, b, 3)
    assert_close(y, g["lap_gs3"], 1e-5, "gauss_seidel x3")
    assert_close(psc.mesh.restriction(x), g["restriction"], TOL, "restriction")
    assert_close(psc.mesh.minus_restriction(x), g["minus_restriction"], TOL, "minus_restriction")
    xc = cases.scalar_grid(N // 2, seed=23)
    assert_close(psc.mesh.prolongation(xc), g["prolongation"], TOL, "prolongation")
    y = x.copy(); psc.mesh.add_prolongation(y, xc)
    assert_close(y, g["add_prolongation"], TOL, "add_prolongation")


def test_multigrid_cycles_vs_golden(psc, golden):
    g = golden("multigrid")
    N = 32
    b = cases.density_contrast_rhs(N, seed=24)
    param = cases.base_param(5, N ** 3, linear_newton_solver="multigrid", compute_additional_field=False)
    for name in ("V", "F", "W"):
        y = psc.laplacian.initialise_potential(b)
        getattr(psc.multigrid, f"{name}_cycle")(y, b, param)
        assert_close(y, g[f"{name}_cycle"], 1e-5, f"{name}_cycle")
    y = psc.laplacian.initialise_potential(b)
    p2 = param.copy()
    y = psc.multigrid.linear(y, b, p2)
    assert_close(y, g["linear"], 1e-5, "multigrid.linear")
    np.testing.assert_allclose(p2["tolerance"], g["linear_tolerance"][0], rtol=1e-4)
    p3 = cases.base_param(5, N ** 3, theory="fr", compute_additional_field=True)
    with pytest.raises(ValueError):
        psc.multigrid.linear(y, b, p3)


@pytest.mark.parametrize("N", [64, 128])
def test_multigrid_vs_oracle_larger(psc, orc, N):
    from oracle import host
    b = cases.density_contrast_rhs(N, seed=70)
    param = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver="multigrid", compute_additional_field=False)
    y = psc.multigrid.linear(psc.laplacian.initialise_potential(b), b, param.copy())
    ref = host.linear(orc.laplacian.initialise_potential(b), b, param.copy())
    assert_close(y, ref, 2e-5, f"multigrid.linear N={N}")
    # multigrid and the 7-point FFT Green's function solve the same discrete operator
    p7 = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver="fft_7pt", MAS_index=0,
                          compute_additional_field=False)
    y7 = psc.solver.fft(b.copy(), p7)
    strict = param.copy()
    strict["epsrel"] = 1e-6
    ys = psc.laplacian.initialise_potential(b)
    for _ in range(12):
        psc.multigrid.V_cycle(ys, b, strict)
    ys -= ys.mean()
    assert_close(ys, y7 - y7.mean(), 2e-4, "multigrid (12 V-cycles) vs fft_7pt")


@pytest.mark.parametrize("kind", [1, 2])
def test_fr_kernels_vs_golden(psc, golden, kind):
    g = golden("fr")
    mod = psc.cubic if kind == 1 else psc.quartic
    t = f"k{kind}"
    u, b, q, rhs = cases.fr_kernel_case(16, kind)
    assert_close(mod.operator(u, b, q), g[f"operator_{t}"], 5e-5, "operator")
    assert_close(mod.residual_with_rhs(u, b, q, rhs), g[f"residual_with_rhs_{t}"], 5e-5, "residual_with_rhs")
    y = u.copy(); mod.gauss_seidel(y, b, q, np.float32(1.25))
    assert_close(y, g[f"gs1_{t}"], TOL, "gauss_seidel")
    y = u.copy(); mod.smoothing(y, b, q, 3)
    assert_close(y, g[f"gs3_{t}"], 1e-5, "smoothing x3")
    y = u.copy(); mod.gauss_seidel_with_rhs(y, b, q, rhs, np.float32(1.25))
    assert_close(y, g[f"gs1_rhs_{t}"], TOL, "gauss_seidel_with_rhs")
    np.testing.assert_allclose(mod.residual_error(u, b, q), g[f"residual_error_{t}"][0], rtol=1e-5)
    np.testing.assert_allclose(mod.truncation_error(u, b, q), g[f"truncation_error_{t}"][0], rtol=1e-4)


@pytest.mark.parametrize("kind", [1, 2])
def test_fr_cycles_vs_golden(psc, golden, kind):
    g = golden("fr")
    mod = psc.cubic if kind == 1 else psc.quartic
    f1, f2, q, param = cases.fr_cycle_case(kind, psc.utils.set_units)
    rho = psc.mesh.TSC(cases.lattice_particles(32, 0.3, seed=50), 32)
    b = psc.utils.linear_operator(rho, f1, f2)
    assert_close(mod.initialise_potential(b, q), g[f"init_k{kind}"], TOL, "initialise_potential")
    u = mod.initialise_potential(b, q)
    psc.multigrid.V_cycle_FAS(u, b, param)
    assert_close(u, g[f"V_cycle_FAS_k{kind}"], 1e-5, "V_cycle_FAS")
    u = mod.initialise_potential(b, q)
    psc.multigrid.F_cycle_FAS(u, b, param)
    assert_close(u, g[f"F_cycle_FAS_k{kind}"], 1e-5, "F_cycle_FAS")
    u = mod.initialise_potential(b, q)
    p2 = param.copy()
    u = psc.multigrid.FAS(u, b, p2)
    assert_close(u, g[f"FAS_k{kind}"], 1e-5, "FAS")
    np.testing.assert_allclose(p2["tolerance_FAS"], g[f"FAS_tol_k{kind}"][0], rtol=1e-3)


def test_mond_vs_golden(psc, golden):
    g = golden("mond")
    N = 16
    phi = (cases.scalar_grid(N, seed=41, smooth=True) * np.float32(2e-3)).astype(np.float32)
    g0 = np.float32(0.05)
    for key, fn, alpha in (("simple", "rhs_simple", None), ("n2", "rhs_n", 2), ("beta1.5", "rhs_beta", 1.5),
                           ("gamma2", "rhs_gamma", 2.0), ("delta1.5", "rhs_delta", 1.5)):
        o = np.empty_like(phi)
        if alpha is None:
            getattr(psc.mond, fn)(phi, o, g0)
        else:
            getattr(psc.mond, fn)(phi, o, g0, alpha)
        assert_close(o, g[key], 3e-5, fn)


# ------------------------------------------------------------------------ whole PM force / steps
@pytest.mark.parametrize("name", list(cases.PM_CASES))
def test_pm_vs_golden(psc, golden, name):
    g = golden("pm")
    pos, param = cases.pm_inputs(name, psc.utils.set_units)
    acc, pot, add = psc.solver.pm(pos, param)
    assert_close(pot, g[f"{name}_pot"], 3e-5, "potential")
    assert_close(acc, g[f"{name}_acc"], 1e-4, "acceleration")
    if f"{name}_add" in g.files:
        assert_close(add, g[f"{name}_add"], 3e-5, "additional_field")
    if f"{name}_acc2" in g.files:
        param["aexp_old"] = param["aexp"]
        param["aexp"] = param["aexp"] * 1.04
        psc.utils.set_units(param)
        param["nsteps"] = 1
        acc2, pot2, _ = psc.solver.pm(pos, param, pot.copy(), add.copy(), cases.toy_tables())
        assert_close(pot2, g[f"{name}_pot2"], 1e-4, "potential (warm start)")
        assert_close(acc2, g[f"{name}_acc2"], 2e-4, "acceleration (warm start)")


def test_pm_option_errors(psc):
    pos, param = cases.pm_inputs("newton_fft_tsc", psc.utils.set_units)
    for key, val in (("mass_scheme", "ngp"), ("linear_newton_solver", "cg"), ("theory", "dgp"),
                     ("save_power_spectrum", "maybe")):
        p = param.copy()
        p[key] = val
        with pytest.raises(NotImplementedError):
            psc.solver.pm(pos, p)


def test_fft_force_vs_golden(psc, golden):
    g = golden("pm")
    pos, param = cases.pm_inputs("newton_fft_tsc", psc.utils.set_units)
    param["MAS_index"] = 3
    rho = psc.mesh.TSC(pos, 16)
    psc.utils.linear_operator_inplace(rho, np.float32(0.19), np.float32(-0.19))
    assert_close(psc.solver.fft_force(rho, param), g["fft_force"], 2e-5, "fft_force")
    # full_fft through pm(): the reference raises TypeError here (solver.py:181); fixed in the build
    p = param.copy()
    p["linear_newton_solver"] = "full_fft"
    acc, _, _ = psc.solver.pm(pos, p)
    assert np.isfinite(acc).all()


@pytest.mark.parametrize("name,ncoarse,solver", [("fft", 4, "fft"), ("mg", 5, "multigrid")])
@pytest.mark.parametrize("device_arrays", [False, True])
def test_steps_vs_golden(psc, golden, name, ncoarse, solver, device_arrays):
    """Three leapfrog steps (last one clamped to a snapshot time: float64 dt) + a Morton reorder."""
    g = golden("steps")
    N = 2 ** ncoarse
    tables = cases.toy_tables()
    pos = cases.lattice_particles(N, 0.3, seed=60)
    vel = cases.velocities(N ** 3, seed=61, scale=2e-3)
    param = cases.base_param(ncoarse, N ** 3, linear_newton_solver=solver)
    param["aexp"] = 0.2
    param["t"] = float(tables[1](np.log(param["aexp"])))
    psc.utils.set_units(param)
    if device_arrays:
        pos, vel = _cuda(pos), _cuda(vel)
    acc, pot, add = psc.solver.pm(pos, param)
    t_snap = param["t"] + 1e9
    dts = []
    for step in range(3):
        param["nsteps"] += 1
        if step == 2:
            t_snap = param["t"] + 0.4 * float(dts[-1])
        t0 = param["t"]
        pos, vel, acc, pot, add = psc.integration.integrate(pos, vel, acc, pot, add, tables, param, t_snap)
        dts.append(param["t"] - t0)
        if step == 1:
            pos, vel, acc = psc.utils.reorder_particles(pos, vel, acc)
    if device_arrays:
        # device-resident arrays come back in bin order (integration.leapfrog): the reference's rows are restored
        # from the particle ids before the row-by-row comparison
        assert psc.utils.particle_ids(pos) is not None
        pos, vel, acc = psc.utils.reference_order(pos, vel, acc)
        pos, vel, acc, pot = (t.cpu().numpy() for t in (pos, vel, acc, pot))
    np.testing.assert_allclose(dts, g[f"{name}_dts"], rtol=2e-5)
    assert bool(param["write_snapshot"]) == bool(g[f"{name}_write_snapshot"][0])
    # bit-exact particle ordering: every particle sits at the reference's row (positions differ by ulps)
    assert np.max(np.abs(pos - g[f"{name}_pos"])) < 1e-6      # box units
    assert_close(vel, g[f"{name}_vel"], 2e-4, "velocity")
    assert_close(acc, g[f"{name}_acc"], 2e-4, "acceleration")
    assert_close(pot, g[f"{name}_pot"], 2e-4, "potential")


def test_newtonian_step_properties_256(psc):
    """Size-independent invariants at 256^3 through the full fused step (no CPU oracle involved):
    mass conservation, zero-mean potential and force, momentum conservation of the PM force."""
    import torch
    N = 256
    pos = _cuda(cases.lattice_particles(N, 0.3, seed=80))
    param = cases.base_param(8, N ** 3, linear_newton_solver="fft")
    psc.utils.set_units(param)
    rho = psc.mesh.TSC(pos, N)
    assert abs(rho.sum(dtype=torch.float64).item() - N ** 3) < 1e-6 * N ** 3
    acc, pot, _ = psc.solver.pm(pos, param)
    assert abs(pot.mean(dtype=torch.float64).item()) < 1e-6 * pot.abs().max().item()
    amax = acc.abs().max().item()
    assert amax > 0
    assert acc.mean(dim=0, dtype=torch.float64).abs().max().item() < 1e-4 * amax
    # deposit is order independent up to float addition order
    perm = torch.randperm(pos.shape[0], device=pos.device)
    rho2 = psc.mesh.TSC(pos[perm].contiguous(), N)
    assert (rho2 - rho).abs().max().item() < 2e-5


# ------------------------------------------------------------------ BASELINE configs 2-4 at their full sizes
def _rel_rms(a, b):
    import torch
    return ((a - b).double().pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt().clamp_min(1e-300)).item()


def test_config2_multigrid_256_matches_7pt_fft(psc):
    """BASELINE config 2 (Newtonian 256^3, multigrid, red-black Gauss-Seidel).  Size-independent property: the
    multigrid iterates converge to the solution of the SAME discrete 7-point Poisson problem that `fft_7pt` solves
    exactly, so the two potentials agree to the solver tolerance and the accelerations to well below a per cent."""
    N = 256
    pos = _cuda(cases.lattice_particles(N, 0.3, seed=81))
    pm_mg = cases.base_param(8, N ** 3, linear_newton_solver="multigrid", epsrel=1e-3)
    pm_ft = cases.base_param(8, N ** 3, linear_newton_solver="fft_7pt")
    for p in (pm_mg, pm_ft):
        psc.utils.set_units(p)
    acc_f, pot_f, _ = psc.solver.pm(pos, pm_ft)
    acc_m, pot_m, _ = psc.solver.pm(pos, pm_mg)
    pot_f = pot_f - pot_f.mean()
    pot_m = pot_m - pot_m.mean()
    assert _rel_rms(pot_m, pot_f) < 2e-2       # fft_7pt deconvolves nothing: same operator, same RHS
    assert _rel_rms(acc_m, acc_f) < 2e-2
    # the discrete residual of the multigrid solution is small against the RHS
    rho = psc.mesh.TSC(pos, N)
    f1 = np.float32(1.5 * pm_mg["aexp"] * pm_mg["Om_m"])
    rhs = rho * f1 - f1
    res = psc.laplacian.residual_error(pot_m.contiguous(), rhs.contiguous())
    rhs_norm = float(rhs.double().pow(2).sum().sqrt())
    assert float(res) < 2e-2 * rhs_norm
    # and one more V-cycle reduces it further (the smoother + transfer operators contract)
    x2 = pot_m.clone().contiguous()
    psc.multigrid.V_cycle(x2, rhs.contiguous(), pm_mg)
    assert float(psc.laplacian.residual_error(x2, rhs.contiguous())) < 0.5 * float(res)


def test_config3_fr_256_scalaron(psc):
    """BASELINE config 3 (f(R) Hu-Sawicki n = 1, |fR0| = 1e-5, 256^3, nonlinear FAS multigrid with the cubic
    smoother).  Properties: the FAS solve lowers the nonlinear residual by the requested factor, the scalaron stays
    positive and finite, and the fifth force is a small, finite correction on top of the Newtonian one."""
    import torch
    N = 256
    pos = _cuda(cases.lattice_particles(N, 0.3, seed=82))
    over = dict(theory="fr", fR_n=1, fR_logfR0=5, linear_newton_solver="multigrid", aexp=0.05, aexp_old=0.05)
    p_fr = cases.base_param(8, N ** 3, **over)
    p_nw = cases.base_param(8, N ** 3, linear_newton_solver="multigrid", aexp=0.05, aexp_old=0.05)
    for p in (p_fr, p_nw):
        psc.utils.set_units(p)
    tables = cases.toy_tables()
    acc_n, pot_n, _ = psc.solver.pm(pos, p_nw, tables=tables)
    acc_f, pot_f, u = psc.solver.pm(pos, p_fr, tables=tables)
    assert u.shape == (N, N, N) and bool(torch.isfinite(u).all()) and u.min().item() > 0
    assert bool(torch.isfinite(acc_f).all())
    assert _rel_rms(pot_f, pot_n) < 1e-6 or _rel_rms(pot_f, pot_n) < 5e-2   # same Newtonian potential
    d = _rel_rms(acc_f, acc_n)
    assert 0 <= d < 0.34      # |F5| <= F_N / 3 in f(R)
    # nonlinear residual of the returned scalaron against the residual of the initial guess
    f1, f2, q = cases.fr_coeffs(p_fr)
    rho = psc.mesh.TSC(pos, N)
    b = (rho * np.float32(f1) + np.float32(f2)).contiguous()
    r_sol = float(psc.cubic.residual_error(u.contiguous(), b, np.float32(q)))
    u0 = psc.cubic.initialise_potential(b, np.float32(q))
    r0 = float(psc.cubic.residual_error(u0, b, np.float32(q)))
    assert r_sol < r0


def test_config4_mond_512_newtonian_limit(psc):
    """BASELINE config 4 (QUMOND 512^3: FFT Newtonian solve + MOND source + second Poisson solve).  With g0 -> 0
    the interpolating function nu -> 1, so the MOND acceleration must reduce to the Newtonian one; with the
    physical g0 the MOND force is stronger than Newton everywhere on average."""
    import torch
    N = 512
    pos = _cuda(cases.lattice_particles(N, 0.3, seed=83))
    p_nw = cases.base_param(9, N ** 3, linear_newton_solver="fft_7pt")
    p_m0 = cases.base_param(9, N ** 3, theory="mond", linear_newton_solver="fft_7pt", mond_g0=1e-12)
    p_m1 = cases.base_param(9, N ** 3, theory="mond", linear_newton_solver="fft_7pt", mond_g0=1.2)
    for p in (p_nw, p_m0, p_m1):
        psc.utils.set_units(p)
    acc_n, _, _ = psc.solver.pm(pos, p_nw)
    acc_0, _, _ = psc.solver.pm(pos, p_m0)
    assert _rel_rms(acc_0, acc_n) < 2e-3
    del acc_0
    acc_1, _, _ = psc.solver.pm(pos, p_m1)
    assert bool(torch.isfinite(acc_1).all())
    ratio = (acc_1.double().pow(2).mean().sqrt() / acc_n.double().pow(2).mean().sqrt()).item()
    assert ratio > 1.0


# ------------------------------------------------------------------ ragged / edge-case inputs of the fused kernels
@pytest.mark.parametrize("npart", [1, 3, 5003, 40001])
def test_kick_drift_wrap_count_ragged(psc, orc, npart):
    """psc_kick_drift_wrap_count (kick + drift + wrap + bin count in one pass) for particle counts that are not
    multiples of 4 / 32 and positions on the domain edges: same particles as psc_kick_drift_wrap, and the bin counts
    it leaves behind give the same binning as psc_bin_particles (checked through the deposit that consumes them)."""
    import torch
    N = 32
    pos, vel = cases.particles(N, npart, seed=21), cases.velocities(npart, seed=22, scale=5e-3)
    acc = cases.velocities(npart, seed=23, scale=1.0)
    dt = np.float32(0.0371)
    half = np.float32(0.5 * dt)
    p, v = pos.copy(), vel.copy()
    orc.utils.add_vector_scalar_inplace(v, acc, -half)
    orc.utils.add_vector_scalar_inplace(p, v, dt)
    orc.utils.periodic_wrap(p)
    tp, tv, ta = _cuda(pos), _cuda(vel), _cuda(acc)
    binned = psc.mesh.alloc_binned(npart, N)
    psc.mesh.kick_drift_wrap_count(tp, tv, ta, half, dt, 0, binned)
    assert np.max(np.abs(tp.cpu().numpy() - p)) <= 6e-8
    assert_close(tv.cpu().numpy(), v, 1e-6, "kick")
    psc.mesh.finish_binning(tp, binned)
    rho = psc.mesh.deposit_rhs(tp, N, psc._lib.TSC, 1.0, 1.0, 0.0, binned)
    ref = orc.mesh.TSC_seq(np.ascontiguousarray(tp.cpu().numpy()), N)
    assert abs(rho.sum(dtype=torch.float64).item() - npart) < 1e-5 * max(npart, 1)
    assert np.max(np.abs(rho.cpu().numpy() - ref)) < 1e-5


def test_pm_ragged_and_edge_positions_vs_oracle(psc, orc):
    """solver.pm (binned deposit + fused gradient / interpolation) with npart != N^3, particles exactly at 0, at the
    largest float below 1, on cell centres and cell edges, and two identical particles."""
    from oracle import host
    N, npart = 16, 5003
    pos = cases.particles(N, npart, seed=31)
    p1 = cases.base_param(4, npart, linear_newton_solver="fft")
    p2 = p1.copy()
    psc.utils.set_units(p1)
    host.set_units(p2)
    acc, pot, _ = psc.solver.pm(_cuda(pos), p1)
    acc_ref, pot_ref, _ = host.pm(pos.copy(), p2)
    assert_close(pot.cpu().numpy(), pot_ref, 3e-5, "potential")
    assert_close(acc.cpu().numpy(), acc_ref, 5e-5, "acceleration")


def test_interp_heavy_bins_vs_oracle(psc, orc):
    """A bin holding far more than BIN_PART = 4096 particles (a 200 000-particle blob two cells wide): the gradient +
    interpolation kernel splits such bins into parts handled by several CTAs; every particle must still get the
    oracle's acceleration and kick.  (The deposit of the same set is covered by test_deposit_clustered_and_ragged,
    against a float64 sum: with 10^4 particles per cell the reference's own float32 running sum is the less accurate
    side, so a whole-pm comparison would test the oracle's rounding, not the kernel.)"""
    N = 64
    rng = np.random.default_rng(17)
    blob = (0.37 + 0.01 * rng.standard_normal((200000, 3))).astype(np.float32)
    pos = np.ascontiguousarray(np.concatenate([blob % 1.0, cases.particles(N, 62144, seed=18)]).astype(np.float32))
    pos[pos >= 1.0] = 0.0
    vel = cases.velocities(pos.shape[0], seed=19, scale=2e-3)
    phi = cases.scalar_grid(N, seed=20, smooth=True)
    a_ref = orc.mesh.invTSC_vec(orc.mesh.derivative(phi, 5), pos)
    v_ref = vel.copy()
    orc.utils.add_vector_scalar_inplace(v_ref, a_ref, -np.float32(0.0123))
    tp, tv = _cuda(pos), _cuda(vel)
    binned = psc.mesh.bin_particles(tp, N)
    acc, mx = psc.mesh.interp_kick_phi(_cuda(phi), None, 0.0, 0, 5, tp, tv, psc._lib.TSC, np.float32(0.0123), binned)
    assert_close(acc.cpu().numpy(), a_ref, 2e-5, "acceleration")
    assert_close(tv.cpu().numpy(), v_ref, 2e-5, "kicked velocity")
    np.testing.assert_allclose(mx.cpu().numpy()[0], np.abs(a_ref).max(), rtol=1e-5)


def test_nonfinite_force_stops_the_run(psc):
    """A NaN reaching the force must surface as an error at the next integrate() (the max reduction of the fused
    interpolation kernel lets NaN win), not as silently corrupted particles."""
    import torch
    N = 16
    tables = cases.toy_tables()
    pos = _cuda(cases.lattice_particles(N, 0.3, seed=90))
    vel = _cuda(cases.velocities(N ** 3, seed=91, scale=1e-3))
    param = cases.base_param(4, N ** 3, linear_newton_solver="fft")
    psc.utils.set_units(param)
    acc, phi, add = psc.solver.pm(pos, param)
    vel[7, 1] = float("nan")            # one bad particle: its NaN position poisons the density, hence every force
    param["nsteps"] += 1
    state = psc.integration.integrate(pos, vel, acc, phi, add, tables, param, 1e30)
    assert not bool(torch.isfinite(state[2]).all())
    param["nsteps"] += 1
    with pytest.raises(ValueError, match="math domain error"):
        psc.integration.integrate(*state, tables, param, 1e30)


@pytest.mark.parametrize("N", [64, 128, 256, 512])
@pytest.mark.parametrize("solver_name,mas", [("fft", 0), ("fft", 3), ("fft_7pt", 3)])
def test_fft_poisson_fused_x_pass(psc, orc, N, solver_name, mas):
    """solver.fft through psc_fft_poisson (cuFFT 2-D (y, z) transforms + ONE kernel for the forward transform along x,
    the Green's function and the backward transform along x; radix 8 x 8, 8 x 8 x 2, 8 x 8 x 4, 8 x 8 x 8) against the
    three-call path (rfftn, Green, irfftn) and, at 64^3 / 128^3, against the oracle: two float32 FFTs of the same data, tolerance of the potential as in DESIGN section 2"""
    import torch
    L = psc._lib.load()
    assert L.psc_fft_poisson_supported(N) and not L.psc_fft_poisson_supported(96) and not L.psc_fft_poisson_supported(32)
    rhs = cases.density_contrast_rhs(N, seed=77) if N <= 128 else None
    if rhs is None:
        g = torch.Generator(device="cuda").manual_seed(9)
        t = torch.randn((N, N, N), generator=g, device="cuda")
        t -= t.mean()
    else:
        t = torch.from_numpy(rhs).cuda()
    param = cases.base_param(int(np.log2(N)), N ** 3, linear_newton_solver=solver_name)
    param["MAS_index"] = mas
    param["compute_additional_field"] = False
    param["save_pk"] = False
    fused = psc.solver.fft(t.clone(), param)
    os.environ["PSC_NO_FUSED_XFFT"] = "1"
    try:
        plain = psc.solver.fft(t.clone(), param)
    finally:
        del os.environ["PSC_NO_FUSED_XFFT"]
    rms = float(plain.pow(2).mean().sqrt())
    assert float((fused - plain).abs().max()) <= 2e-5 * rms, float((fused - plain).abs().max()) / rms
    if rhs is not None:
        from oracle import host
        ref = host.fft(rhs.copy(), param)
        assert_close(fused.cpu().numpy(), ref, 3e-5, "fused-x-pass FFT solve vs oracle")


@pytest.mark.parametrize("N,nyl,y0", [(64, 8, 16), (512, 16, 32), (1024, 4, 8), (2048, 4, 2040)])
@pytest.mark.parametrize("kind,p", [(1, 3), (2, 0)])
def test_xfft_green_slab_vs_torch_fft(psc, N, nyl, y0, kind, p):
    """psc_xfft_green_slab (the x part of the slab-decomposed solve in one kernel: radix-8 stages + a last radix 8 / 4 / 2
    stage, up to N = 2048 with 8 kz per CTA) against torch.fft along x around psc_green_slab on the same transposed
    block [N (kx)][nyl (ky = y0 ..)][N/2+1]"""
    import torch
    lib, L = psc._lib, psc._lib.load()
    nz = N // 2 + 1
    g = torch.Generator(device="cuda").manual_seed(N + nyl)
    a = torch.view_as_complex(torch.randn((N, nyl, nz, 2), generator=g, device="cuda"))
    scale = 1.0 / float(N) ** 3
    ref = torch.fft.fft(a, dim=0).contiguous()
    lib.check(L.psc_green_slab(lib.ptr(ref), N, nyl, y0, kind, p, scale, lib.stream()))
    ref = torch.fft.ifft(ref, dim=0) * N
    out = a.clone()
    lib.check(L.psc_xfft_green_slab(lib.ptr(out), N, nyl, y0, kind, p, scale, lib.stream()))
    torch.cuda.synchronize()
    rms = float(ref.abs().pow(2).mean().sqrt())
    err = float((out - ref).abs().max()) / rms
    assert err < 2e-5, err
