#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (mianbreton/pysco, Numba) on the
seeded inputs of tests/golden/cases.py.

Runs only in the build container (needs /root/reference and numba); the parity tests and the GPU
box read the committed .npz files, never /root/reference.

    python tests/golden/make_golden.py            # all groups
    python tests/golden/make_golden.py kernels pm # selected groups

The reference imports astropy (absent here) for three constants; tests/golden/_stubs/astropy
provides them (values = astropy's CODATA-2018 defaults).  nthreads = 1 everywhere: the reference
is bit-reproducible only single-threaded, and that selects TSC_seq and np.argsort.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYSCO_REFERENCE", "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_pysco")
os.environ.setdefault("NUMBA_NUM_THREADS", "8")
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, os.path.join(REF, "pysco"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import numba  # noqa: E402

import cases  # noqa: E402

numba.set_num_threads(1)

import utils, mesh, morton, fourier, laplacian, multigrid, cubic, quartic, mond, solver, integration  # noqa: E402,E401


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrs)} arrays")


def g_particles():
    out = {}
    for N, npart in ((16, 5000), (32, 3000)):
        pos = cases.particles(N, npart)
        tag = f"N{N}"
        out[f"keys_{tag}"] = morton.positions_to_keys(pos)
        out[f"argsort_{tag}"] = np.argsort(out[f"keys_{tag}"], kind="stable")
        out[f"tsc_{tag}"] = mesh.TSC_seq(pos, N)
        out[f"cic_{tag}"] = mesh.CIC(pos, N)
        out[f"ngp_{tag}"] = mesh.NGP(pos, N)
        f3 = cases.vector_grid(N)
        f1 = cases.scalar_grid(N)
        out[f"invtsc_vec_{tag}"] = mesh.invTSC_vec(f3, pos)
        out[f"invcic_vec_{tag}"] = mesh.invCIC_vec(f3, pos)
        out[f"invngp_vec_{tag}"] = mesh.invNGP_vec(f3, pos)
        out[f"invtsc_{tag}"] = mesh.invTSC(f1, pos)
        out[f"invcic_{tag}"] = mesh.invCIC(f1, pos)
        out[f"invngp_{tag}"] = mesh.invNGP(f1, pos)
    # reorder_particles (nthreads = 1 -> global np.argsort), utils.py:1019-1075.  Ties (pos[6] ==
    # pos[7]) carry identical rows so the unstable sort cannot change the result.
    pos = cases.particles(16, 5000)
    vel = cases.velocities(5000)
    acc = cases.velocities(5000, seed=44, scale=1.0)
    p2, v2, a2 = utils.reorder_particles(pos.copy(), vel.copy(), acc.copy())
    out["reorder_pos"], out["reorder_vel"], out["reorder_acc"] = p2, v2, a2
    # leapfrog-style vector ops
    y = vel.copy()
    utils.add_vector_scalar_inplace(y, acc, np.float32(-0.0123))
    out["axpy_f32"] = y
    y = pos.copy()
    utils.add_vector_scalar_inplace(y, vel, 0.731)  # python float -> float64 scalar path
    out["axpy_f64"] = y
    w = (pos * np.float32(3.0) - np.float32(1.0)).astype(np.float32)
    w.ravel()[:6] = [-1e-9, -2.0 ** -25, -2.0 ** -25 * 1.000002, -1e-7, 1.0, np.nextafter(np.float32(1), np.float32(0))]
    # one wrap pass only folds [-1, 2) into [0, 1): keep inputs in that range like the reference's callers
    w = np.clip(w, -0.999, 1.999).astype(np.float32)
    out["wrap_in"] = w.copy()
    utils.periodic_wrap(w)
    out["wrap_out"] = w
    out["max_abs"] = np.array([utils.max_abs(acc)], dtype=np.float32)
    save("particles", **out)


def g_grids():
    out = {}
    N = 16
    x = cases.scalar_grid(N, seed=11)
    for order in (2, 3, 5, 7):
        out[f"deriv{order}"] = mesh.derivative(x, order)
        u = cases.scalaron_field(N)
        for n in (1, 2):
            out[f"deriv{order}_fR{n}"] = mesh.derivative_fR(x, u, np.float32(0.37), n, order)
            f = cases.vector_grid(N, seed=12)
            mesh.add_derivative_fR(f, u, np.float32(0.37), n, order)
            out[f"addderiv{order}_fR{n}"] = f
    y = x.copy()
    utils.linear_operator_inplace(y, np.float32(0.193), np.float32(-0.193))
    out["linop"] = y
    y = x.copy()
    utils.prod_vector_scalar_inplace(y, np.float32(1.7))
    out["scale"] = y
    y = x.copy()
    utils.linear_operator_vectors_inplace(y, np.float32(4), cases.scalar_grid(N, seed=13), np.float32(1))
    out["lincomb"] = y
    # fourier
    for NN in (16, 32):
        r = cases.density_contrast_rhs(NN)
        spec = fourier.fft_3D_real(r, 1)
        out[f"rfft_N{NN}"] = spec
        s = spec.copy(); fourier.inverse_laplacian(s); out[f"green_plain_N{NN}"] = s
        for p in (2, 3):
            s = spec.copy(); fourier.inverse_laplacian_compensated(s, p); out[f"green_comp{p}_N{NN}"] = s
        s = spec.copy(); fourier.inverse_laplacian_7pt(s); out[f"green_7pt_N{NN}"] = s
        out[f"irfft_7pt_N{NN}"] = fourier.ifft_3D_real(s, 1)
        if NN == 16:
            out[f"gradgreen_N{NN}"] = fourier.gradient_inverse_laplacian(spec.copy())
            out[f"gradgreen_comp3_N{NN}"] = fourier.gradient_inverse_laplacian_compensated(spec.copy(), 3)
            out[f"irfft_grad_N{NN}"] = fourier.ifft_3D_real_grad(out[f"gradgreen_comp3_N{NN}"].copy(), 1)
        for p in (0, 2, 3):
            k, pk, nm = fourier.fourier_grid_to_Pk(spec.copy(), p)
            out[f"pk_p{p}_N{NN}"] = np.stack([k, pk, nm])
    save("grids", **out)


def g_multigrid():
    out = {}
    N = 16
    x = cases.scalar_grid(N, seed=21, smooth=True)
    b = cases.density_contrast_rhs(N, seed=22)
    out["lap_operator"] = laplacian.operator(x)
    out["lap_residual"] = laplacian.residual(x, b)
    out["lap_restrict_residual"] = laplacian.restrict_residual(x, b)
    out["lap_residual_error"] = np.array([laplacian.residual_error(x, b)], dtype=np.float32)
    out["lap_truncation_error"] = np.array([laplacian.truncation_error(x)], dtype=np.float32)
    out["lap_init"] = laplacian.initialise_potential(b)
    y = x.copy(); laplacian.gauss_seidel(y, b, np.float32(1.25)); out["lap_gs1"] = y
    y = x.copy(); laplacian.smoothing(y, b, 3); out["lap_gs3"] = y
    out["restriction"] = mesh.restriction(x)
    out["minus_restriction"] = mesh.minus_restriction(x)
    xc = cases.scalar_grid(N // 2, seed=23)
    out["prolongation"] = mesh.prolongation(xc)
    y = x.copy(); mesh.add_prolongation(y, xc); out["add_prolongation"] = y
    # cycles at N = 32 (ncoarse = 5: levels 32 -> 16 -> 8 -> 4)
    N = 32
    b = cases.density_contrast_rhs(N, seed=24)
    param = cases.base_param(5, N ** 3, linear_newton_solver="multigrid", compute_additional_field=False)
    for name, fn in (("V", multigrid.V_cycle), ("F", multigrid.F_cycle), ("W", multigrid.W_cycle)):
        y = laplacian.initialise_potential(b)
        fn(y, b, param)
        out[f"{name}_cycle"] = y
    y = laplacian.initialise_potential(b)
    p2 = param.copy()
    y = multigrid.linear(y, b, p2)
    out["linear"] = y
    out["linear_tolerance"] = np.array([p2["tolerance"]], dtype=np.float64)
    save("multigrid", **out)


def g_fr():
    out = {}
    N = 16
    for kind, mod in ((1, cubic), (2, quartic)):
        t = f"k{kind}"
        u, b, q, rhs = cases.fr_kernel_case(N, kind)
        out[f"operator_{t}"] = mod.operator(u, b, q)
        out[f"residual_with_rhs_{t}"] = mod.residual_with_rhs(u, b, q, rhs)
        y = u.copy(); mod.gauss_seidel(y, b, q, np.float32(1.25)); out[f"gs1_{t}"] = y
        y = u.copy(); mod.smoothing(y, b, q, 3); out[f"gs3_{t}"] = y
        y = u.copy(); mod.gauss_seidel_with_rhs(y, b, q, rhs, np.float32(1.25)); out[f"gs1_rhs_{t}"] = y
        out[f"residual_error_{t}"] = np.array([mod.residual_error(u, b, q)], dtype=np.float32)
        out[f"truncation_error_{t}"] = np.array([mod.truncation_error(u, b, q)], dtype=np.float32)
    # root solvers on a grid of (p, d1) / (p, q) covering the branches
    rng = np.random.default_rng(5)
    p = np.concatenate([rng.uniform(-3, 3, 400), [-3.0, 2.0, 1e-3, -1e-3]]).astype(np.float32)
    d = np.concatenate([rng.uniform(-2, 2, 400), [10.392305, 0.5, -0.1, 0.1]]).astype(np.float32)
    out["roots_p"], out["roots_d"] = p, d

    def _try(fn, a, c):
        # the reference raises ZeroDivisionError on the branches that take a fractional power of a
        # negative number (python error model); those inputs are recorded as NaN and not compared
        try:
            return fn(a, c)
        except ZeroDivisionError:
            return np.nan
    out["roots_cubic"] = np.array([_try(cubic.solution_cubic_equation, a, c) for a, c in zip(p, d)], dtype=np.float32)
    qq = (-np.abs(d) - np.float32(1e-3)).astype(np.float32)
    out["roots_q"] = qq
    out["roots_quartic"] = np.array([_try(quartic.solution_quartic_equation, a, c) for a, c in zip(p, qq)],
                                    dtype=np.float32)
    # initial guess, FAS V/F cycles and the full FAS solve, N = 32, real a = 0.05 coefficients
    for kind, mod in ((1, cubic), (2, quartic)):
        f1, f2, q, param = cases.fr_cycle_case(kind, utils.set_units)
        rho = mesh.TSC_seq(cases.lattice_particles(32, 0.3, seed=50), 32)
        b = utils.linear_operator(rho, f1, f2)
        out[f"init_k{kind}"] = mod.initialise_potential(b, q)
        u = mod.initialise_potential(b, q)
        multigrid.V_cycle_FAS(u, b, param)
        out[f"V_cycle_FAS_k{kind}"] = u
        u = mod.initialise_potential(b, q)
        multigrid.F_cycle_FAS(u, b, param)
        out[f"F_cycle_FAS_k{kind}"] = u
        u = mod.initialise_potential(b, q)
        p2 = param.copy()
        u = multigrid.FAS(u, b, p2)
        out[f"FAS_k{kind}"] = u
        out[f"FAS_tol_k{kind}"] = np.array([p2["tolerance_FAS"]], dtype=np.float64)
    save("fr", **out)


def g_mond():
    out = {}
    N = 16
    phi = (cases.scalar_grid(N, seed=41, smooth=True) * np.float32(2e-3)).astype(np.float32)
    g0 = np.float32(0.05)
    o = np.empty_like(phi); mond.rhs_simple(phi, o, g0); out["simple"] = o.copy()
    o = np.empty_like(phi); mond.rhs_n(phi, o, g0, 2); out["n2"] = o.copy()
    o = np.empty_like(phi); mond.rhs_beta(phi, o, g0, np.float32(1.5)); out["beta1.5"] = o.copy()
    o = np.empty_like(phi); mond.rhs_gamma(phi, o, g0, np.float32(2.0)); out["gamma2"] = o.copy()
    o = np.empty_like(phi); mond.rhs_delta(phi, o, g0, np.float32(1.5)); out["delta1.5"] = o.copy()
    save("mond", **out)


def g_pm():
    out = {}
    for name in cases.PM_CASES:
        pos, param = cases.pm_inputs(name, utils.set_units)
        acc, pot, add = solver.pm(pos, param)
        out[f"{name}_acc"], out[f"{name}_pot"] = acc, pot
        if len(add):
            out[f"{name}_add"] = add
        # second call warm-started from the first (exercises initialise_potential rescaling)
        if param["linear_newton_solver"] == "multigrid":
            param["aexp_old"] = param["aexp"]
            param["aexp"] = param["aexp"] * 1.04
            utils.set_units(param)
            param["nsteps"] = 1
            acc2, pot2, add2 = solver.pm(pos, param, pot.copy(), add.copy(), cases.toy_tables())
            out[f"{name}_acc2"], out[f"{name}_pot2"] = acc2, pot2
    # fft_force (full_fft arithmetic; reachable only at function level)
    pos, param = cases.pm_inputs("newton_fft_tsc", utils.set_units)
    param["MAS_index"] = 3
    rho = mesh.TSC_seq(pos, 16)
    utils.linear_operator_inplace(rho, np.float32(0.19), np.float32(-0.19))
    out["fft_force"] = solver.fft_force(rho, param)
    save("pm", **out)


def g_steps():
    """Three leapfrog steps through integration.integrate with analytic tables, incl. a snapshot
    clamp (float64 dt) on the last one, and a reorder."""
    out = {}
    for name, ncoarse, over in (("fft", 4, dict(linear_newton_solver="fft")),
                                ("mg", 5, dict(linear_newton_solver="multigrid"))):
        N = 2 ** ncoarse
        tables = cases.toy_tables()
        pos = cases.lattice_particles(N, 0.3, seed=60)
        vel = cases.velocities(N ** 3, seed=61, scale=2e-3)
        param = cases.base_param(ncoarse, N ** 3, **over)
        param["aexp"] = 0.2
        param["t"] = float(tables[1](np.log(param["aexp"])))
        utils.set_units(param)
        acc, pot, add = solver.pm(pos, param)
        t_snap = param["t"] + 1e9
        dts = []
        for step in range(3):
            param["nsteps"] += 1
            if step == 2:
                t_snap = param["t"] + 0.4 * float(dts[-1])
            t0 = param["t"]
            pos, vel, acc, pot, add = integration.integrate(pos, vel, acc, pot, add, tables, param, t_snap)
            dts.append(param["t"] - t0)
            if step == 1:
                pos, vel, acc = utils.reorder_particles(pos, vel, acc)
        out[f"{name}_pos"], out[f"{name}_vel"], out[f"{name}_acc"], out[f"{name}_pot"] = pos, vel, acc, pot
        out[f"{name}_dts"] = np.array(dts, dtype=np.float64)
        out[f"{name}_aexp"] = np.array([param["aexp"]], dtype=np.float64)
        out[f"{name}_write_snapshot"] = np.array([param["write_snapshot"]])
    save("steps", **out)


GROUPS = {"particles": g_particles, "grids": g_grids, "multigrid": g_multigrid, "fr": g_fr,
          "mond": g_mond, "pm": g_pm, "steps": g_steps}



def g_run():
    """Whole-run golden (BASELINE config 1 shape at 32^3): pysco.run, z = 49 -> 0, FFT and multigrid
    solvers, leapfrog, TSC, n_reorder = 50, P(k) at the last snapshot.  The initial conditions the
    reference generated (2LPT, seed 42) are stored so that the build starts from identical particles."""
    import shutil
    import pandas as pd
    sys.path.insert(0, REF)
    import pysco
    import initial_conditions, cosmotable  # noqa: E401
    out = {}
    for solver_name in ("fft", "multigrid"):
        base = f"/tmp/pysco_golden_run_{solver_name}/"
        shutil.rmtree(base, ignore_errors=True)
        param = cases.run_param(base, solver_name)
        # capture the ICs by generating them exactly as main.run does (main.py:105-112)
        p0 = pd.Series(dict(param))
        p0["write_snapshot"] = False
        p0["extra"] = "ics"
        os.makedirs(base + "/output_00000", exist_ok=True)
        p0["i_snap"] = 0
        tables = cosmotable.generate(p0)
        p0["aexp"] = 1.0 / (1 + p0["z_start"])
        utils.set_units(p0)
        p0["nsteps"] = 0
        pos0, vel0 = initial_conditions.generate(p0, tables)
        if solver_name == "fft":
            out["ic_pos"], out["ic_vel"] = pos0.copy(), vel0.copy()
        else:
            assert np.array_equal(out["ic_pos"], pos0)
        pysco.run(dict(param))
        import pyarrow.parquet as pq
        extra = f"newton_{solver_name}_ncoarse5"
        snap = f"{base}/output_00006/particles_{extra}.parquet"
        t = pq.read_table(snap)
        out[f"{solver_name}_pos"] = np.stack([np.asarray(t[c]) for c in ("x", "y", "z")], axis=1).astype(np.float32)
        out[f"{solver_name}_vel"] = np.stack([np.asarray(t[c]) for c in ("vx", "vy", "vz")], axis=1).astype(np.float32)
        pks = sorted(f for f in os.listdir(f"{base}/power") if f.endswith(".dat"))
        out[f"{solver_name}_pk_last"] = np.loadtxt(f"{base}/power/{pks[-1]}")
        out[f"{solver_name}_nsteps"] = np.array([int(pks[-1].split("_")[-1].split(".")[0])])
    save("run", **out)


GROUPS["run"] = g_run

def g_run64():
    """Whole run at 64^3 (FFT solver, z = 49 -> 0): the reference's ICs, the final P(k) file, and every 16th particle of
    the final snapshot (the full snapshot would be another 6 MB)."""
    import shutil
    import pandas as pd
    sys.path.insert(0, REF)
    import pysco
    import initial_conditions, cosmotable  # noqa: E401
    import pyarrow.parquet as pq
    out = {}
    base = "/tmp/pysco_golden_run64/"
    shutil.rmtree(base, ignore_errors=True)
    param = cases.run_param(base, "fft", ncoarse=6)
    p0 = pd.Series(dict(param))
    p0["write_snapshot"] = False
    p0["extra"] = "ics"
    os.makedirs(base + "/output_00000", exist_ok=True)
    p0["i_snap"] = 0
    tables = cosmotable.generate(p0)
    p0["aexp"] = 1.0 / (1 + p0["z_start"])
    utils.set_units(p0)
    p0["nsteps"] = 0
    pos0, vel0 = initial_conditions.generate(p0, tables)
    out["ic_pos"], out["ic_vel"] = pos0.copy(), vel0.copy()
    pysco.run(dict(param))
    snap = f"{base}/output_00006/particles_newton_fft_ncoarse6.parquet"
    t = pq.read_table(snap)
    out["fft_pos_16th"] = np.stack([np.asarray(t[c]) for c in ("x", "y", "z")], axis=1).astype(np.float32)[::16]
    out["fft_vel_16th"] = np.stack([np.asarray(t[c]) for c in ("vx", "vy", "vz")], axis=1).astype(np.float32)[::16]
    pks = sorted(f for f in os.listdir(f"{base}/power") if f.endswith(".dat"))
    out["fft_pk_last"] = np.loadtxt(f"{base}/power/{pks[-1]}")
    out["fft_nsteps"] = np.array([int(pks[-1].split("_")[-1].split(".")[0])])
    save("run64", **out)


GROUPS["run64"] = g_run64


def g_ics():
    """Initial conditions of the reference (initial_conditions.generate, nthreads = 1) at 16^3 for every LPT
    order / option, plus the table look-ups it used so that the build can be fed exactly the same growth factors."""
    import pandas as pd
    import cosmotable
    import initial_conditions as ic
    out = {}
    base = "/tmp/pysco_golden_ics"
    os.makedirs(base + "/output_00000", exist_ok=True)
    for name, over in cases.IC_CASES.items():
        param = pd.Series(cases.ic_param(base, **over))
        param["aexp"] = 1.0 / (1 + param["z_start"])
        utils.set_units(param)
        tables = cosmotable.generate(param)
        lna = np.log(1.0 / (1 + param["z_start"]))
        out[f"{name}_tables"] = np.array([float(tables[2](lna))] + [float(tables[3](0))]
                                         + [float(tables[i](lna)) for i in range(3, 13)])
        out[f"{name}_unit_t"] = np.array([param["unit_t"]])
        pos, vel = ic.generate(param, tables)
        out[f"{name}_pos"], out[f"{name}_vel"] = pos.astype(np.float32), vel.astype(np.float32)
    # host-side pieces, small enough for the CPU tier: white noise (half-spectrum part) and the density spectrum
    for N, seed in ((8, 5), (12, 6)):
        out[f"wn_N{N}"] = ic.white_noise_fourier(N, np.random.default_rng(seed))[:, :, :N // 2 + 1]
        out[f"wnfixed_N{N}"] = ic.white_noise_fourier_fixed(N, np.random.default_rng(seed), True)[:, :, :N // 2 + 1]
    param = pd.Series(cases.ic_param(base, npart=8 ** 3, seed=9))
    out["density_fourier_N8"] = ic.generate_density_fourier(param)[:, :, :5]
    save("ics", **out)


GROUPS["ics"] = g_ics

def g_euler():
    """Three Euler steps through integration.integrate (integration.py:121-189), the last one clamped to a snapshot
    time; FFT solver at 16^3."""
    out = {}
    N = 16
    tables = cases.toy_tables()
    pos = cases.lattice_particles(N, 0.3, seed=62)
    vel = cases.velocities(N ** 3, seed=63, scale=2e-3)
    param = cases.base_param(4, N ** 3, linear_newton_solver="fft", integrator="euler")
    param["aexp"] = 0.2
    param["t"] = float(tables[1](np.log(param["aexp"])))
    utils.set_units(param)
    acc, pot, add = solver.pm(pos, param)
    t_snap = param["t"] + 1e9
    dts = []
    for step in range(3):
        param["nsteps"] += 1
        if step == 2:
            t_snap = param["t"] + 0.4 * float(dts[-1])
        t0 = param["t"]
        pos, vel, acc, pot, add = integration.integrate(pos, vel, acc, pot, add, tables, param, t_snap)
        dts.append(param["t"] - t0)
    out["pos"], out["vel"], out["acc"], out["pot"] = pos, vel, acc, pot
    out["dts"] = np.array(dts, dtype=np.float64)
    out["write_snapshot"] = np.array([param["write_snapshot"]])
    save("euler", **out)


GROUPS["euler"] = g_euler


def g_cosmotable():
    """The reference's cosmotable.generate (cosmotable.py:18-110: supercomoving time, growth ODEs of 1-3LPT) evaluated
    at 24 scale factors for LCDM (examples/param.ini), w0-wa and the parametrized theory.  astropy is absent from this
    container: the Flatw0waCDM class the reference instantiates is the repo's restatement (tests/golden/_stubs), so
    these vectors pin everything cosmotable.py computes ON TOP of E(a) -- not astropy's E(a) itself."""
    import pandas as pd
    import cosmotable
    out = {}
    a = np.geomspace(1.0 / 150, 1.0, 24)
    out["aexp"] = a
    for name, over in cases.COSMO_CASES.items():
        param = pd.Series(cases.cosmo_param(**over))
        param["base"] = "/tmp/pysco_golden_cosmo"
        os.makedirs(param["base"], exist_ok=True)
        tables = cosmotable.generate(param)
        lna = np.log(a)
        t = tables[1](lna)
        out[f"{name}_tables"] = np.array([tables[0](t)] + [tb(lna) for tb in tables[1:]], dtype=np.float64)
        out[f"{name}_Om_r_Om_lambda"] = np.array([param["Om_r"], param["Om_lambda"]], dtype=np.float64)
    save("cosmotable", **out)


GROUPS["cosmotable"] = g_cosmotable


if __name__ == "__main__":
    todo = sys.argv[1:] or list(GROUPS)
    for g in todo:
        print("==", g, flush=True)
        GROUPS[g]()
