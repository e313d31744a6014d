"""oracle/host.py -- TEST INFRASTRUCTURE ONLY (see oracle/pysco_oracle.c header).

CPU restatement of the reference's host-side sequencing of the PM step on NumPy arrays:
``solver.pm`` (solver.py:30-215), ``solver.fft`` (:453-522), ``solver.fft_force`` (:526-579),
``solver.rhs_poisson`` (:381-449), ``solver.get_additional_field`` (:285-378),
``solver.initialise_potential`` (:218-282), ``multigrid.linear/FAS/V|F|W_cycle[_FAS]``
(multigrid.py:23-858), ``integration.integrate/leapfrog/euler/dt_*`` (integration.py:17-358),
``utils.set_units`` (utils.py:167-196).  Kernels come from oracle.api (C restatement).
"""
import numpy as np

from .api import cubic, fourier, laplacian, mesh, mond, quartic, utils

# astropy.constants values the reference reads (utils.py:11, solver.py:14); CODATA-2018 / IAU-2015.
C_LIGHT = 299792458.0
PARSEC = 3.0856775814913673e16
G_NEWTON = 6.6743e-11

_EMPTY = np.empty(0, dtype=np.float32)


def set_units(param):
    mpc_to_km = 1e3 * PARSEC
    g = G_NEWTON * 1e-9
    H0 = param["H0"] / mpc_to_km
    rhoc = 3.0 * H0 ** 2 / (8.0 * np.pi * g)
    param["unit_l"] = param["aexp"] * param["boxlen"] * 100.0 / H0
    param["unit_t"] = param["aexp"] ** 2 / H0
    param["unit_d"] = param["Om_m"] * rhoc / param["aexp"] ** 3
    param["mpart"] = param["unit_d"] * param["unit_l"] ** 3 / param["npart"]


# ------------------------------------------------------------------------------- multigrid
def _fr_mod(param):
    n = param["fR_n"]
    if n == 1:
        return cubic
    if n == 2:
        return quartic
    raise NotImplementedError(f"Only f(R) with n = 1 and 2, currently {n=}")


def _is_scalaron(param):
    return bool(param["compute_additional_field"]) and param["theory"].casefold() == "fr"


def _coarsest(param, nlevel):
    return nlevel >= (param["ncoarse"] - 3)


def _cycle(kind, x, b, param, nlevel=0):
    """kind in 'V','F','W' -- multigrid.py:474-517, 583-638, 722-776."""
    laplacian.smoothing(x, b, param["Npre"])
    visits = [kind] if kind == "V" else [kind, "V" if kind == "F" else "W"]
    for n, sub in enumerate(visits):
        if n:
            laplacian.smoothing(x, b, param["Npre"])
        res_c = laplacian.restrict_residual(x, b)
        corr = laplacian.initialise_potential(res_c)
        if _coarsest(param, nlevel):
            laplacian.smoothing(corr, res_c, param["Npre"])
        else:
            _cycle(sub, corr, res_c, param, nlevel + 1)
        mesh.add_prolongation(x, corr)
    laplacian.smoothing(x, b, param["Npost"])


def V_cycle(x, b, param, nlevel=0):
    _cycle("V", x, b, param, nlevel)


def F_cycle(x, b, param, nlevel=0):
    _cycle("F", x, b, param, nlevel)


def W_cycle(x, b, param, nlevel=0):
    _cycle("W", x, b, param, nlevel)


def _fas_smooth(x, b, n, param, rhs):
    m = _fr_mod(param)
    q = np.float32(param["fR_q"])
    if len(rhs) == 0:
        m.smoothing(x, b, q, n)
    else:
        m.smoothing_with_rhs(x, b, q, n, rhs)


def _fas_restrict_residual(x, b, param, rhs):
    m = _fr_mod(param)
    q = np.float32(param["fR_q"])
    if len(rhs) == 0:
        return mesh.minus_restriction(m.operator(x, b, q))
    return mesh.restriction(m.residual_with_rhs(x, b, q, rhs))


def _cycle_FAS(kind, x, b, param, nlevel=0, rhs=_EMPTY):
    """multigrid.py:521-579 (V), 642-718 (F), 780-858 (W); scalaron (cubic/quartic) operators only."""
    m = _fr_mod(param)
    q = np.float32(param["fR_q"])
    _fas_smooth(x, b, param["Npre"], param, rhs)
    visits = [kind] if kind == "V" else [kind, "V" if kind == "F" else "W"]
    b_c = mesh.restriction(b)
    for n, sub in enumerate(visits):
        if n:
            _fas_smooth(x, b, param["Npre"], param, rhs)
        res_c = _fas_restrict_residual(x, b, param, rhs)
        x_c = mesh.restriction(x)
        corr = x_c.copy()
        L_c = m.operator(x_c, b_c, q)
        utils.linear_operator_vectors_inplace(res_c, np.float32(4), L_c, np.float32(1))
        if _coarsest(param, nlevel):
            _fas_smooth(corr, b_c, param["Npre"], param, res_c)
        else:
            _cycle_FAS(sub, corr, b_c, param, nlevel + 1, res_c)
        utils.add_vector_scalar_inplace(corr, x_c, np.float32(-1))
        mesh.add_prolongation(x, corr)
    _fas_smooth(x, b, param["Npost"], param, rhs)


def V_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY):
    _cycle_FAS("V", x, b, param, nlevel, rhs)


def F_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY):
    _cycle_FAS("F", x, b, param, nlevel, rhs)


def W_cycle_FAS(x, b, param, nlevel=0, rhs=_EMPTY):
    _cycle_FAS("W", x, b, param, nlevel, rhs)


def linear(x, b, param):
    """multigrid.py:23-83"""
    theory = param["theory"].casefold()
    if param["compute_additional_field"] and theory == "fr":
        raise ValueError("Linear should not be used for scalaron field")
    mond_pass = (not param["compute_additional_field"]) and theory == "mond"
    if ("tolerance" not in param) or (param["nsteps"] % 3) == 0:
        tol = param["epsrel"] * laplacian.truncation_error(x)
        param["tolerance_mond" if mond_pass else "tolerance"] = tol
    tol = param["tolerance_mond"] if mond_pass else param["tolerance"]
    err = 1e30
    while err > tol:
        V_cycle(x, b, param)
        e = laplacian.residual_error(x, b)
        if e < tol or err / e < 2:
            break
        err = e
    return x


def FAS(x, b, param):
    """multigrid.py:88-140 (f(R) scalaron path)"""
    m = _fr_mod(param)
    q = np.float32(param["fR_q"])
    if ("tolerance_FAS" not in param) or (param["nsteps"] % 3) == 0:
        param["tolerance_FAS"] = param["epsrel"] * m.truncation_error(x, b, q)
    tol = param["tolerance_FAS"]
    err = 1e30
    while err > tol:
        V_cycle_FAS(x, b, param)
        e = m.residual_error(x, b, q)
        if e < tol or err / e < 2:
            break
        err = e
    return x


# ---------------------------------------------------------------------------------- solver
def initialise_potential(potential, rhs, param, tables):
    """solver.py:218-282"""
    if len(potential) == 0:
        if _is_scalaron(param):
            return _fr_mod(param).initialise_potential(rhs, param["fR_q"])
        return laplacian.initialise_potential(rhs)
    if not param["compute_additional_field"]:
        scaling = (param["aexp"] * tables[3](np.log(param["aexp"]))
                   / (param["aexp_old"] * tables[3](np.log(param["aexp_old"]))))
        utils.prod_vector_scalar_inplace(potential, scaling)
    return potential


def _fr_background(param):
    a = param["aexp"]
    Rbar = 3 * param["Om_m"] * a ** (-3) + 12 * param["Om_lambda"]
    Rbar0 = 3 * param["Om_m"] + 12 * param["Om_lambda"]
    fR_a = -a ** 2 * ((Rbar0 / Rbar) ** (param["fR_n"] + 1)) * 10.0 ** (-param["fR_logfR0"])
    c2 = (C_LIGHT * 1e-3 * param["unit_t"] / (param["unit_l"] * a)) ** 2
    return Rbar, fR_a, c2


def rhs_poisson(density, additional_field, param):
    """solver.py:381-449 (in place on density)"""
    mond_pass = (param["compute_additional_field"] is False) and param["theory"].casefold() == "mond"
    if mond_pass:
        g0 = (param["mond_g0"] * 1e-3 * 1e-10 * param["unit_t"] ** 2 / param["unit_l"]
              * param["aexp"] ** (1 + param["mond_scale_factor_exponent"]))
        alpha = param["mond_alpha"]
        fn = param["mond_function"].casefold()
        if fn not in ("simple", "n", "beta", "gamma", "delta"):
            raise NotImplementedError(f"{fn=}")
        getattr(mond, "rhs_" + fn)(additional_field, density, g0, alpha)
    else:
        f1 = np.float32(1.5 * param["aexp"] * param["Om_m"] * param["parametrized_mu_z"])
        utils.linear_operator_inplace(density, f1, -f1)


def fft(rhs, param, pk_sink=None):
    """solver.py:453-522"""
    p = param["MAS_index"]
    spec = fourier.fft_3D_real(rhs, param["nthreads"])
    solver_name = param["linear_newton_solver"].casefold()
    mond_pass = (param["compute_additional_field"] is False) and param["theory"] == "mond"
    if "save_pk" in param and param["save_pk"] and not mond_pass:
        k, Pk, Nmodes = fourier.fourier_grid_to_Pk(spec, p)
        Pk *= ((param["boxlen"] / len(rhs) ** 2) ** 3 / (1.5 * param["aexp"] * param["Om_m"]) ** 2
               / param["parametrized_mu_z"] ** 2)
        k *= 2 * np.pi / param["boxlen"]
        if pk_sink is not None:
            pk_sink(k, Pk, Nmodes, param)
    if solver_name == "fft":
        if p == 0:
            fourier.inverse_laplacian(spec)
        else:
            fourier.inverse_laplacian_compensated(spec, p)
    elif solver_name == "fft_7pt":
        fourier.inverse_laplacian_7pt(spec)
    else:
        raise NotImplementedError(f"{solver_name=}")
    return fourier.ifft_3D_real(spec, param["nthreads"])


def fft_force(rhs, param):
    """solver.py:526-579"""
    p = param["MAS_index"]
    spec = fourier.fft_3D_real(rhs, param["nthreads"])
    if p == 0:
        force = fourier.gradient_inverse_laplacian(spec)
    else:
        force = fourier.gradient_inverse_laplacian_compensated(spec, p)
    return fourier.ifft_3D_real_grad(force, param["nthreads"])


def get_additional_field(additional_field, density, param, tables):
    """solver.py:285-378"""
    theory = param["theory"].casefold()
    if theory in ("newton", "parametrized"):
        return np.empty(0, dtype=np.float32)
    if theory == "fr":
        Rbar, fR_a, c2 = _fr_background(param)
        a = param["aexp"]
        f1 = np.float32(a * param["Om_m"] / (c2 * 6)) / (-fR_a)
        f2 = np.float32(Rbar / 3 * a ** 4 - param["Om_m"] * a) / (6 * c2) / (-fR_a)
        dens_term = utils.linear_operator(density, f1, f2)
        q = np.float32(-a ** 4 * Rbar / (18 * c2)) / (-fR_a)
        param["fR_q"] = q
        u = initialise_potential(additional_field, dens_term, param, tables)
        return FAS(u, dens_term, param)
    if theory == "mond":
        rhs_poisson(density, additional_field, param)
        name = param["linear_newton_solver"].casefold()
        if name == "multigrid":
            phi = initialise_potential(additional_field, density, param, tables)
            return linear(phi, density, param)
        if name == "fft_7pt":
            return fft(density, param)
        raise NotImplementedError(f"{name=}")
    raise NotImplementedError(f"{theory=}")


def pm(position, param, potential=_EMPTY, additional_field=_EMPTY, tables=(), pk_sink=None):
    """solver.py:30-215.  Canonical (bit-reproducible) oracle configuration: param["nthreads"] = 1."""
    N = 2 ** param["ncoarse"]
    scheme = param["mass_scheme"].casefold()
    theory = param["theory"].casefold()
    if scheme == "cic":
        param["MAS_index"] = 2
        density = mesh.CIC(position, N)
    elif scheme == "tsc":
        param["MAS_index"] = 3
        # solver.py:86-89: atomic (parallel) TSC from 4 threads up, sequential below
        density = mesh.TSC(position, N) if param["nthreads"] >= 4 else mesh.TSC_seq(position, N)
    else:
        raise NotImplementedError(f"{param['mass_scheme']=}")
    if theory == "parametrized":
        a = param["aexp"]
        evo = a ** (-3 * (1 + param["w0"] + param["wa"])) * np.exp(-3 * param["wa"] * (1 - a))
        olz = param["Om_lambda"] * evo / (param["Om_m"] * a ** (-3) + param["Om_r"] * a ** (-4)
                                         + param["Om_lambda"] * evo)
        param["parametrized_mu_z"] = np.float32(1 + param["parametrized_mu0"] * olz / param["Om_lambda"])
    else:
        param["parametrized_mu_z"] = np.float32(1)
    if N ** 3 != param["npart"]:
        utils.prod_vector_scalar_inplace(density, np.float32(N ** 3 / param["npart"]))
    save = param["save_power_spectrum"].casefold()
    if save == "yes":
        param["save_pk"] = True
    elif save == "z_out":
        param["save_pk"] = bool(param["write_snapshot"])
    elif save == "no":
        param["save_pk"] = False
    else:
        raise NotImplementedError(f"{save=}")
    name = param["linear_newton_solver"].casefold()
    if param["save_pk"] and name == "multigrid":
        spec = fourier.fft_3D_real(density, param["nthreads"])
        k, Pk, Nmodes = fourier.fourier_grid_to_Pk(spec, param["MAS_index"])
        Pk *= (param["boxlen"] / len(density) ** 2) ** 3
        k *= 2 * np.pi / param["boxlen"]
        if pk_sink is not None:
            pk_sink(k, Pk, Nmodes, param)
    param["compute_additional_field"] = True
    additional_field = get_additional_field(additional_field, density, param, tables)
    param["compute_additional_field"] = False
    rhs_poisson(density, additional_field, param)
    rhs = density
    if name == "multigrid":
        potential = initialise_potential(potential, rhs, param, tables)
        potential = linear(potential, rhs, param)
    elif name in ("fft", "fft_7pt"):
        potential = fft(rhs, param, pk_sink)
    elif name == "full_fft":
        pass
    else:
        raise NotImplementedError(f"{name=}")
    order = param["gradient_stencil_order"]
    if theory == "fr":
        _, fR_a, c2 = _fr_background(param)
        half_c2 = np.float32(0.5 * (-fR_a) * c2)
        if name == "full_fft":
            force = fft_force(rhs, param)
            mesh.add_derivative_fR(force, additional_field, half_c2, param["fR_n"], order)
        else:
            force = mesh.derivative_fR(potential, additional_field, half_c2, param["fR_n"], order)
    else:
        force = fft_force(rhs, param) if name == "full_fft" else mesh.derivative(potential, order)
    acceleration = mesh.invCIC_vec(force, position) if scheme == "cic" else mesh.invTSC_vec(force, position)
    return acceleration, potential, additional_field


# ----------------------------------------------------------------------------- integration
def dt_CFL_maxacc(acceleration, param):
    dx = np.float32(0.5 ** param["ncoarse"])
    return np.float32(param["Courant_factor"]) * np.sqrt(dx / utils.max_abs(acceleration))


def dt_CFL_maxvel(velocity, param):
    dx = np.float32(0.5 ** param["ncoarse"])
    return np.float32(param["Courant_factor"]) * dx / utils.max_abs(velocity)


def dt_weak_variation(func_t_a, param):
    f = 1.0 + 0.01 * param["max_aexp_stepping"]
    return np.float32(func_t_a(np.log(f * param["aexp"])) - func_t_a(np.log(param["aexp"])))


def leapfrog(position, velocity, acceleration, potential, additional_field, dt, tables, param):
    half_dt = np.float32(0.5 * dt)
    utils.add_vector_scalar_inplace(velocity, acceleration, -half_dt)
    utils.add_vector_scalar_inplace(position, velocity, dt)
    param["t"] += dt
    param["aexp_old"] = param["aexp"]
    param["aexp"] = np.exp(tables[0](param["t"]))
    set_units(param)
    utils.periodic_wrap(position)
    acceleration, potential, additional_field = pm(position, param, potential, additional_field, tables)
    utils.add_vector_scalar_inplace(velocity, acceleration, -half_dt)
    return position, velocity, acceleration, potential, additional_field


def euler(position, velocity, acceleration, potential, additional_field, dt, tables, param):
    utils.add_vector_scalar_inplace(position, velocity, dt)
    param["t"] += dt
    param["aexp_old"] = param["aexp"]
    param["aexp"] = np.exp(tables[0](param["t"]))
    set_units(param)
    utils.periodic_wrap(position)
    utils.add_vector_scalar_inplace(velocity, acceleration, -dt)
    acceleration, potential, additional_field = pm(position, param, potential, additional_field, tables)
    return position, velocity, acceleration, potential, additional_field


def integrate(position, velocity, acceleration, potential, additional_field, tables, param,
              t_snap_next=np.float32(0)):
    dt = np.min([dt_CFL_maxacc(acceleration, param), dt_CFL_maxvel(velocity, param),
                 dt_weak_variation(tables[1], param)])
    if (param["t"] + dt) > t_snap_next:
        dt = t_snap_next - param["t"]
        param["write_snapshot"] = True
    else:
        param["write_snapshot"] = False
    name = param["integrator"].casefold()
    if name == "leapfrog":
        return leapfrog(position, velocity, acceleration, potential, additional_field, dt, tables, param)
    if name == "euler":
        return euler(position, velocity, acceleration, potential, additional_field, dt, tables, param)
    raise NotImplementedError("ERROR: Integrator must be 'leapfrog' or 'euler'")
