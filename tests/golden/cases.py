"""Deterministic inputs shared by tests/golden/make_golden.py (which feeds them to the unmodified
reference) and by the parity tests (which feed them to the oracle and to the CUDA path).

Inputs are regenerated from seeds so only the reference OUTPUTS are stored in tests/golden/.
"""
import numpy as np


def particles(N: int, npart: int, seed: int = 42, edge_cases: bool = True) -> np.ndarray:
    """Uniform random positions in [0,1) plus the edge cases the domain has: exact 0, the largest
    float32 below 1, exact cell centres (d == 0: CIC sign(0) == 0), exact cell edges, and the last
    cell (wrap-around of the +1 neighbour)."""
    rng = np.random.default_rng(seed)
    pos = rng.random((npart, 3), dtype=np.float32)
    if edge_cases and npart >= 16:
        h = np.float32(1.0 / N)
        pos[0] = (0.0, 0.0, 0.0)
        pos[1] = (np.nextafter(np.float32(1), np.float32(0)),) * 3
        pos[2] = (0.5 * h, 0.5 * h, 0.5 * h)                      # cell centre of cell 0
        pos[3] = (h, 2 * h, 3 * h)                                # exact cell edges
        pos[4] = (1 - 0.5 * h, 1 - 0.25 * h, 1 - 0.75 * h)        # last cell
        pos[5] = (0.25 * h, 1 - 0.25 * h, 0.5)                    # first/last cell mix
        pos[6] = pos[7] = (0.3, 0.6, 0.9)                         # identical particles (key ties)
    return np.ascontiguousarray(pos)


def lattice_particles(N: int, sigma_cells: float = 0.3, seed: int = 42) -> np.ndarray:
    """SURVEY 8(d) synthetic ICs: cell-centre lattice in lexicographic (i,j,k) order plus a
    Gaussian displacement of sigma_cells/N per axis, wrapped into [0,1)."""
    rng = np.random.default_rng(seed)
    g = (np.arange(N, dtype=np.float32) + np.float32(0.5)) / np.float32(N)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = pos + rng.standard_normal(pos.shape, dtype=np.float32) * np.float32(sigma_cells / N)
    pos = pos - np.floor(pos)
    pos[pos >= 1.0] = 0.0
    return np.ascontiguousarray(pos.astype(np.float32))


def velocities(npart: int, seed: int = 43, scale: float = 1e-3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.standard_normal((npart, 3), dtype=np.float32) * np.float32(scale))


def scalar_grid(N: int, seed: int = 7, smooth: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, N, N), dtype=np.float32)
    if smooth:  # low-pass so that stencils see a resolved field
        xk = np.fft.rfftn(x)
        k = np.fft.fftfreq(N)[:, None, None] ** 2 + np.fft.fftfreq(N)[None, :, None] ** 2 \
            + np.fft.rfftfreq(N)[None, None, :] ** 2
        x = np.fft.irfftn(xk * np.exp(-k * 40.0), s=(N, N, N), axes=(0, 1, 2)).astype(np.float32)
        x /= np.abs(x).max()
    return np.ascontiguousarray(x)


def vector_grid(N: int, seed: int = 8) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.standard_normal((N, N, N, 3), dtype=np.float32))


def density_contrast_rhs(N: int, seed: int = 9) -> np.ndarray:
    """Zero-mean right-hand side like 1.5*a*Om*(rho-1)."""
    x = scalar_grid(N, seed, smooth=True)
    x -= x.mean(dtype=np.float64).astype(np.float32)
    return np.ascontiguousarray(x.astype(np.float32))


def scalaron_field(N: int, seed: int = 10) -> np.ndarray:
    """Positive field of order one (f(R) u = sqrt(-f_R/f_R0)-like)."""
    x = scalar_grid(N, seed, smooth=True)
    return np.ascontiguousarray((1.0 + 0.2 * x).astype(np.float32))


def base_param(N_log2: int, npart: int, **over):
    """Parameter Series of examples/param.ini (reference) with test overrides."""
    import pandas as pd
    p = {
        "nthreads": 1, "theory": "newton", "fR_logfR0": 5, "fR_n": 1, "mond_function": "simple",
        "mond_g0": 1.2, "mond_scale_factor_exponent": 0, "mond_alpha": 1, "parametrized_mu0": -0.1,
        "H0": 72, "Om_m": 0.25733, "T_cmb": 2.726, "N_eff": 3.044, "w0": -1.0, "wa": 0.0,
        "boxlen": 100, "ncoarse": N_log2, "npart": npart, "z_start": 49, "seed": 42,
        "integrator": "leapfrog", "mass_scheme": "TSC", "n_reorder": 50, "Courant_factor": 1.0,
        "max_aexp_stepping": 10, "linear_newton_solver": "fft", "gradient_stencil_order": 5,
        "Npre": 2, "Npost": 1, "epsrel": 1e-2, "verbose": 0, "save_power_spectrum": "no",
        "aexp": 0.5, "aexp_old": 0.5, "t": -0.5, "nsteps": 0, "write_snapshot": False,
        # derived by cosmotable.generate in a full run (cosmotable.py:36-37)
        "Om_r": 8.0763e-05, "Om_lambda": 1.0 - 0.25733 - 8.0763e-05,
    }
    p.update(over)
    return pd.Series(p)


def toy_tables():
    """Analytic stand-ins for cosmotable's interpolators (Einstein-de Sitter-like supercomoving
    time: dt = dlna / (a^2 E), E = a^-1.5 -> t = -2 (a^-1/2 - 1)); D1 = a.  Index layout as in
    cosmotable.py:96-110: [lna(t), t(lna), H(lna), D1(lna), ...]."""
    from scipy.interpolate import interp1d
    lna = np.linspace(np.log(1.0 / 201), 0.0, 20001)
    a = np.exp(lna)
    t = -2.0 * (a ** -0.5 - 1.0)
    tabs = [interp1d(t, lna, fill_value="extrapolate"), interp1d(lna, t, fill_value="extrapolate"),
            interp1d(lna, 72.0 * a ** -1.5, fill_value="extrapolate"),
            interp1d(lna, a, fill_value="extrapolate")]
    return tabs + [interp1d(lna, np.ones_like(a), fill_value="extrapolate")] * 9


PM_CASES = {
    # name: (ncoarse, overrides)
    "newton_fft_tsc": (4, dict(linear_newton_solver="fft", mass_scheme="TSC")),
    "newton_fft_cic": (4, dict(linear_newton_solver="fft", mass_scheme="CIC", gradient_stencil_order=3)),
    "newton_fft7_tsc": (4, dict(linear_newton_solver="fft_7pt", gradient_stencil_order=7)),
    "newton_mg_tsc": (5, dict(linear_newton_solver="multigrid", gradient_stencil_order=5)),
    "param_fft_tsc": (4, dict(theory="parametrized", linear_newton_solver="fft", gradient_stencil_order=2)),
    # f(R): high redshift (screened regime).  At a = 0.5 on these tiny grids the reference's cubic
    # root solver takes a fractional power of a negative number and raises (cubic.py:196-197).
    "fr1_mg_tsc": (5, dict(theory="fr", fR_n=1, linear_newton_solver="multigrid", aexp=0.05, aexp_old=0.05)),
    "fr2_fft_tsc": (4, dict(theory="fr", fR_n=2, fR_logfR0=6, linear_newton_solver="fft", aexp=0.05, aexp_old=0.05)),
    "mond_fft7_tsc": (4, dict(theory="mond", linear_newton_solver="fft_7pt")),
    "mond_mg_cic": (5, dict(theory="mond", linear_newton_solver="multigrid", mass_scheme="CIC", mond_function="n", mond_alpha=2)),
    "newton_fft_tsc_np": (4, dict(linear_newton_solver="fft", npart_factor=2)),
}


def pm_inputs(name, set_units):
    ncoarse, over = PM_CASES[name]
    over = dict(over)
    N = 2 ** ncoarse
    fac = over.pop("npart_factor", 1)
    pos = lattice_particles(N, 0.3, seed=50)
    if fac != 1:
        pos = np.ascontiguousarray(np.concatenate([pos, particles(N, N ** 3 * (fac - 1), seed=51)]))
    param = base_param(ncoarse, pos.shape[0], **over)
    set_units(param)
    return pos, param


def fr_coeffs(param, c_light=299792458.0):
    """f1, f2, q of solver.get_additional_field (solver.py:326-347) for the given param."""
    a = param["aexp"]
    Rbar = 3 * param["Om_m"] * a ** (-3) + 12 * param["Om_lambda"]
    Rbar0 = 3 * param["Om_m"] + 12 * param["Om_lambda"]
    fR_a = -a ** 2 * ((Rbar0 / Rbar) ** (param["fR_n"] + 1)) * 10.0 ** (-param["fR_logfR0"])
    c2 = (c_light * 1e-3 * param["unit_t"] / (param["unit_l"] * a)) ** 2
    f1 = np.float32(a * param["Om_m"] / (c2 * 6)) / (-fR_a)
    f2 = np.float32(Rbar / 3 * a ** 4 - param["Om_m"] * a) / (6 * c2) / (-fR_a)
    q = np.float32(-a ** 4 * Rbar / (18 * c2)) / (-fR_a)
    return np.float32(f1), np.float32(f2), np.float32(q)


def fr_kernel_case(N, kind):
    """(u, b, q, rhs) for single-level cubic/quartic kernel tests: u ~ 1 with h^2 b << 1, i.e. the
    three-real-root (acos) branch of the cubic and the S-branch of the quartic."""
    u = scalaron_field(N, seed=31)
    b = (np.float32(2.0) * (np.float32(1.0) + np.float32(0.3) * density_contrast_rhs(N, seed=32))).astype(np.float32)
    q = np.float32(-2.0)
    rhs = (scalar_grid(N, seed=33, smooth=True) * np.float32(1e-4)).astype(np.float32)
    return u, np.ascontiguousarray(b), q, np.ascontiguousarray(rhs)


def fr_cycle_case(kind, set_units):
    """(b, q, param) for FAS-cycle tests at N = 32: coefficients of a real a = 0.05 f(R) step on the
    displaced-lattice density (screened regime on every level)."""
    N = 32
    param = base_param(5, N ** 3, theory="fr", fR_n=kind, aexp=0.05, aexp_old=0.05,
                       compute_additional_field=True, linear_newton_solver="multigrid")
    set_units(param)
    f1, f2, q = fr_coeffs(param)
    param["fR_q"] = q
    return f1, f2, q, param


def run_param(base, solver_name, ncoarse=5):
    """param of a whole pysco.run at (2^ncoarse)^3: examples/param.ini with the size / solver overrides
    (BASELINE config 1 shape)."""
    N = 2 ** ncoarse
    return {
        "nthreads": 1, "theory": "newton", "fR_logfR0": 5, "fR_n": 1, "mond_function": "simple", "mond_g0": 1.2,
        "mond_scale_factor_exponent": 0, "mond_alpha": 1, "parametrized_mu0": -0.1, "H0": 72, "Om_m": 0.25733,
        "T_cmb": 2.726, "N_eff": 3.044, "w0": -1.0, "wa": 0.0, "boxlen": 100, "ncoarse": ncoarse, "npart": N ** 3,
        "z_start": 49, "seed": 42, "position_ICS": "center", "fixed_ICS": False, "paired_ICS": False,
        "dealiased_ICS": False, "power_spectrum_file": "/root/reference/examples/pk_lcdmw7v2.dat",
        "initial_conditions": "2LPT", "base": base, "output_snapshot_format": "parquet",
        "z_out": "[10, 5, 2, 1, 0.5, 0]", "save_power_spectrum": "z_out", "integrator": "leapfrog",
        "mass_scheme": "TSC", "n_reorder": 50, "Courant_factor": 1.0, "max_aexp_stepping": 10,
        "linear_newton_solver": solver_name, "gradient_stencil_order": 5, "Npre": 2, "Npost": 1, "epsrel": 1e-2,
        "verbose": 0,
    }


# ---------------------------------------------------------------------------------------- initial conditions
def ic_pk_table():
    """A smooth LCDM-like P(k) table (k in h/Mpc, P in (Mpc/h)^3) written to a temporary file by the IC tests and by
    make_golden.py, so that neither needs the reference's data files at run time."""
    k = np.logspace(-4, 2, 400)
    q = k / 0.2
    T = np.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    return k, 2.0e4 * k ** 0.96 * T ** 2


def ic_pk_file(base):
    import os
    os.makedirs(base, exist_ok=True)
    path = os.path.join(base, "pk_test.dat")
    k, P = ic_pk_table()
    np.savetxt(path, np.c_[k, P])
    return path


def ic_param(base, **over):
    p = {
        "initial_conditions": "2LPT", "z_start": 49.0, "theory": "newton", "H0": 72.0, "Om_m": 0.25733,
        "T_cmb": 2.726, "N_eff": 3.044, "w0": -1.0, "wa": 0.0, "base": base, "extra": "test",
        "power_spectrum_file": ic_pk_file(base), "boxlen": 100, "npart": 16 ** 3, "seed": 42,
        "fixed_ICS": False, "paired_ICS": False, "dealiased_ICS": False, "position_ICS": "center",
        "output_snapshot_format": "parquet", "nthreads": 1, "parametrized_mu0": 0.0, "verbose": 0,
    }
    p.update(over)
    return p


IC_CASES = {
    "lpt1_edge": dict(initial_conditions="1LPT", position_ICS="edge"),
    "lpt2": dict(initial_conditions="2LPT"),
    "lpt3": dict(initial_conditions="3LPT", seed=7),
    "lpt2_fixed_paired": dict(initial_conditions="2LPT", fixed_ICS=True, paired_ICS=True, seed=3),
    "lpt3_dealiased": dict(initial_conditions="3LPT", dealiased_ICS=True, seed=11),
}


# background tables (cosmotable.generate): LCDM of examples/param.ini, a w0-wa model, the parametrized theory
COSMO_CASES = {
    "lcdm": dict(),
    "w0wa": dict(w0=-0.9, wa=0.1, Om_m=0.3),
    "parametrized": dict(theory="parametrized", parametrized_mu0=0.1),
}


def cosmo_param(**over):
    p = {"theory": "newton", "H0": 72.0, "Om_m": 0.25733, "T_cmb": 2.726, "N_eff": 3.044, "w0": -1.0, "wa": 0.0,
         "parametrized_mu0": 0.0, "evolution_table": "no", "base": "", "extra": "test"}
    p.update(over)
    return p
