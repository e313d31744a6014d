"""Mirror of pysco/solver.py: the particle-mesh force, re-hosted over the CUDA kernels.

pm                    solver.py:30-215
initialise_potential  solver.py:218-282
get_additional_field  solver.py:285-378
rhs_poisson           solver.py:381-449
fft / fft_force       solver.py:453-579   (fft_force: the reference calls it with a stray third
                                           argument from pm(), raising TypeError; fixed here)
force_3d              solver.py:582-639

Fusions relative to the reference's call sequence (same arithmetic, fewer passes over HBM):
  * TSC/CIC deposit + density rescale + `1.5 a Om mu (rho - 1)`   -> one kernel (Newtonian path)
  * Green's function x deconvolution x 1/N^3 (irfftn normalisation) -> one kernel
  * inverse interpolation + second half-kick + max|a|, max|v|      -> one kernel (via integration)
"""
import logging

import numpy as np
import torch

from . import _lib, cubic, distributed, fourier, iostream, laplacian, mesh, mond, multigrid, quartic, utils

_C_LIGHT = 299792458.0  # astropy.constants.c (solver.py:14)
_EMPTY = np.empty(0, dtype=np.float32)
_SCHEMES = {"cic": (_lib.CIC, 2), "tsc": (_lib.TSC, 3)}


def _theory(param):
    return param["theory"].casefold()


def _fr_background(param):
    a = param["aexp"]
    Rbar = 3 * param["Om_m"] * a ** (-3) + 12 * param["Om_lambda"]
    Rbar0 = 3 * param["Om_m"] + 12 * param["Om_lambda"]
    fR_a = -a ** 2 * ((Rbar0 / Rbar) ** (param["fR_n"] + 1)) * 10.0 ** (-param["fR_logfR0"])
    c2 = (_C_LIGHT * 1e-3 * param["unit_t"] / (param["unit_l"] * a)) ** 2  # m -> km -> BU
    return Rbar, fR_a, c2


def initialise_potential(potential, rhs, param, tables):
    """solver.py:218-282: first guess from the RHS, or previous potential rescaled by a*D1(a)."""
    if len(potential) == 0:
        logging.info("Assign potential from density field")
        if param["compute_additional_field"] and "fr" == _theory(param):
            q = param["fR_q"]
            if param["fR_n"] == 1:
                return cubic.initialise_potential(rhs, q)
            if param["fR_n"] == 2:
                return quartic.initialise_potential(rhs, q)
            raise NotImplementedError(f"Only f(R) with n = 1 and 2, currently {param['fR_n']=}")
        return laplacian.initialise_potential(rhs)
    logging.info("Rescale potential from previous step for Newtonian potential")
    if not param["compute_additional_field"]:
        scaling = (param["aexp"] * tables[3](np.log(param["aexp"]))
                   / (param["aexp_old"] * tables[3](np.log(param["aexp_old"]))))
        utils.prod_vector_scalar_inplace(potential, scaling)
    return potential


def rhs_poisson(density, additional_field, param) -> None:
    """solver.py:381-449 (in place on density)."""
    compute_MOND_potential = param["compute_additional_field"] is False and "mond" == _theory(param)
    if compute_MOND_potential:
        g0 = (param["mond_g0"] * 1e-3 * 1e-10 * param["unit_t"] ** 2 / param["unit_l"]
              * param["aexp"] ** (1 + param["mond_scale_factor_exponent"]))
        alpha = param["mond_alpha"]
        fn = param["mond_function"].casefold()
        if fn == "simple":
            mond.rhs_simple(additional_field, density, g0)
        elif fn == "n":
            mond.rhs_n(additional_field, density, g0, n=alpha)
        elif fn == "beta":
            mond.rhs_beta(additional_field, density, g0, beta=alpha)
        elif fn == "gamma":
            mond.rhs_gamma(additional_field, density, g0, gamma=alpha)
        elif fn == "delta":
            mond.rhs_delta(additional_field, density, g0, delta=alpha)
        else:
            raise NotImplementedError(f"MOND_FUNCTION={fn!r}, should be 'simple', 'n', 'beta', 'gamma' or 'delta'")
    else:
        f1 = np.float32(1.5 * param["aexp"] * param["Om_m"] * param["parametrized_mu_z"])
        utils.linear_operator_inplace(density, f1, -f1)


def _write_pk(spec, param, rhs_len, newtonian_rhs):
    k, Pk, Nmodes = fourier.fourier_grid_to_Pk(spec, param["MAS_index"])
    Pk *= (param["boxlen"] / rhs_len ** 2) ** 3
    if newtonian_rhs:
        Pk /= (1.5 * param["aexp"] * param["Om_m"]) ** 2 * param["parametrized_mu_z"] ** 2
    k *= 2 * np.pi / param["boxlen"]
    iostream.write_power_spectrum_to_ascii_file(k, Pk, Nmodes, param)


def fft(rhs, param):
    """solver.py:453-522: FFT Poisson solve.  rhs is consumed (the potential is written over it when
    it is a device tensor); Green multiply carries the 1/N^3 of the inverse transform."""
    c = _lib.Ctx()
    trhs = c.dev(rhs)
    N = trhs.shape[0]
    MAS_index = param["MAS_index"]
    name = param["linear_newton_solver"].casefold()
    compute_MOND_potential = param["compute_additional_field"] is False and param["theory"] == "mond".casefold()
    want_pk = "save_pk" in param and param["save_pk"] and not compute_MOND_potential
    scale = 1.0 / float(N) ** 3
    lib = _lib.load()
    if (not want_pk and name in ("fft", "fft_7pt") and lib.psc_fft_poisson_supported(N)
            and not __import__("os").environ.get("PSC_NO_FUSED_XFFT")):
        # cuFFT does the (y, z) transforms of the x planes; the transforms along x and the Green's function are one
        # kernel (csrc/fourier.cu xfft_green_kernel): 5 passes over the spectrum instead of 7 (N = 2^k, 64..2048)
        kind = _lib.GREEN_7PT if name == "fft_7pt" else (_lib.GREEN_PLAIN if MAS_index == 0 else _lib.GREEN_COMPENSATED)
        spec = _lib.empty((N, N, N // 2 + 1), torch.complex64)
        out = trhs if not c.np_mode else _lib.empty((N, N, N))
        _lib.check(lib.psc_fft_poisson(_lib.fft_plan(N), _lib.ptr(trhs), _lib.ptr(spec), _lib.ptr(out), kind,
                                       int(MAS_index), scale, _lib.stream()))
        return c.ret(out)
    spec = fourier.fft_3D_real(trhs, param["nthreads"])
    if want_pk:
        _write_pk(spec, param, N, True)
    if name == "fft":
        if MAS_index == 0:
            fourier.inverse_laplacian(spec, scale)
        else:
            fourier.inverse_laplacian_compensated(spec, MAS_index, scale)
    elif name == "fft_7pt":
        fourier.inverse_laplacian_7pt(spec, scale)
    else:
        raise NotImplementedError(f"LINEAR_NEWTON_SOLVER={name!r}, should be 'fft' or 'fft_7pt'")
    out = trhs if not c.np_mode else None
    pot = fourier.ifft_3D_real(spec, param["nthreads"], out=out, prescaled=True)
    return c.ret(pot)


def fft_force(rhs, param):
    """solver.py:526-579: spectral force -i k/k^2 (full_fft); returns AoS [N,N,N,3]."""
    c = _lib.Ctx()
    trhs = c.dev(rhs)
    N = trhs.shape[0]
    MAS_index = param["MAS_index"]
    spec = fourier.fft_3D_real(trhs, param["nthreads"])
    force_k = fourier.gradient_inverse_laplacian_compensated(spec, MAS_index, 1.0 / float(N) ** 3)
    if "save_pk" in param and param["save_pk"]:
        _write_pk(spec, param, N, True)
    del spec
    return c.ret(fourier.ifft_3D_real_grad(force_k, param["nthreads"], prescaled=True))


def get_additional_field(additional_field, density, param, tables):
    """solver.py:285-378"""
    THEORY = _theory(param)
    if THEORY in ("newton", "parametrized"):
        return _EMPTY
    if THEORY == "fr":
        Rbar, fR_a, c2 = _fr_background(param)
        a = param["aexp"]
        f1 = np.float32(a * param["Om_m"] / (c2 * 6)) / (-fR_a)
        f2 = np.float32(Rbar / 3 * a ** 4 - param["Om_m"] * a) / (6 * c2) / (-fR_a)
        dens_term = utils.linear_operator(density, f1, f2)
        q = np.float32(-a ** 4 * Rbar / (18 * c2)) / (-fR_a)
        param["fR_q"] = q
        u_scalaron = initialise_potential(additional_field, dens_term, param, tables)
        u_scalaron = multigrid.FAS(u_scalaron, dens_term, param)
        logging.info(f"{fR_a=}")
        return u_scalaron
    if THEORY == "mond":
        rhs_poisson(density, additional_field, param)
        name = param["linear_newton_solver"].casefold()
        if name == "multigrid":
            additional_field = initialise_potential(additional_field, density, param, tables)
            return multigrid.linear(additional_field, density, param)
        if name == "fft_7pt":
            # fft() consumes its input: density is re-used as the MOND RHS output afterwards
            return fft(density.clone() if isinstance(density, torch.Tensor) else density, param)
        raise NotImplementedError(f"{param['linear_newton_solver']=}, should be 'multigrid' or 'fft_7pt'")
    raise NotImplementedError(f"{param['theory']=}, should be 'newton', 'fr', 'parametrized' or 'mond'")


def _pm_device(position, param, potential, additional_field, tables, kick=None, counted=None):
    """pm() on device tensors.  kick = (velocity, half_dt) fuses the second leapfrog half-kick and the
    max reductions into the interpolation; returns (acc, potential, additional_field, maxima|None)."""
    ncells_1d = 2 ** (param["ncoarse"])
    MASS_SCHEME = param["mass_scheme"].casefold()
    THEORY = _theory(param)
    if MASS_SCHEME not in _SCHEMES:
        raise NotImplementedError(f"{param['mass_scheme']=}, should be 'CIC' or 'TSC'")
    scheme, param["MAS_index"] = _SCHEMES[MASS_SCHEME]

    if "parametrized" == THEORY:
        a = param["aexp"]
        evolution_term = a ** (-3 * (1 + param["w0"] + param["wa"])) * np.exp(-3 * param["wa"] * (1 - a))
        omega_lambda_z = (param["Om_lambda"] * evolution_term
                          / (param["Om_m"] * a ** (-3) + param["Om_r"] * a ** (-4)
                             + param["Om_lambda"] * evolution_term))
        param["parametrized_mu_z"] = np.float32(1 + param["parametrized_mu0"] * omega_lambda_z / param["Om_lambda"])
    else:
        param["parametrized_mu_z"] = np.float32(1)

    SAVE_POWER_SPECTRUM = param["save_power_spectrum"].casefold()
    if SAVE_POWER_SPECTRUM == "yes":
        param["save_pk"] = True
    elif SAVE_POWER_SPECTRUM == "z_out":
        param["save_pk"] = bool(param["write_snapshot"])
    elif SAVE_POWER_SPECTRUM == "no":
        param["save_pk"] = False
    else:
        raise NotImplementedError(f"{SAVE_POWER_SPECTRUM=}, should be 'yes', 'z_out' or 'no'")
    LINEAR_NEWTON_SOLVER = param["linear_newton_solver"].casefold()
    if LINEAR_NEWTON_SOLVER not in ("multigrid", "fft", "fft_7pt", "full_fft"):
        raise NotImplementedError(
            f"{param['linear_newton_solver']=}, should be multigrid, fft, fft_7pt or full_fft")

    conversion = np.float32(ncells_1d ** 3 / param["npart"]) if ncells_1d ** 3 != param["npart"] else np.float32(1)
    # one shadow binning of the particles per step, shared by the deposit and the interpolation
    # (`counted`: a Binned whose counts were already produced by the fused kick-drift-wrap of integration.leapfrog)
    if isinstance(counted, mesh.SortedBins):
        binned = counted        # the particle arrays are in bin order already (integration.leapfrog: mesh.step_sort)
    elif counted is not None:
        binned = mesh.finish_binning(position, counted)
    else:
        binned = mesh.bin_particles(position, ncells_1d) if mesh.can_bin(ncells_1d, position.shape[0]) else None
    pk_from_density = param["save_pk"] and "multigrid" == LINEAR_NEWTON_SOLVER
    fuse_rhs = THEORY in ("newton", "parametrized") and not pk_from_density
    if distributed.is_active():
        # particle-parallel / mesh-replicated: local counts -> all-reduce(sum) -> the affine maps
        density = mesh.deposit_rhs(position, ncells_1d, scheme, 1.0, 1.0, 0.0, binned)
        distributed.allreduce_sum_(density)
        if conversion != 1:
            utils.prod_vector_scalar_inplace(density, conversion)
        if fuse_rhs:
            f1 = np.float32(1.5 * param["aexp"] * param["Om_m"] * param["parametrized_mu_z"])
            utils.linear_operator_inplace(density, f1, -f1)
            rhs = density
            param["compute_additional_field"] = False
            additional_field = _EMPTY
    elif fuse_rhs:
        # deposit + rescale + 1.5 a Om mu (rho - 1) in one kernel
        f1 = np.float32(1.5 * param["aexp"] * param["Om_m"] * param["parametrized_mu_z"])
        rhs = mesh.deposit_rhs(position, ncells_1d, scheme, conversion, f1, -f1, binned)
        param["compute_additional_field"] = False
        additional_field = _EMPTY
    else:
        density = mesh.deposit_rhs(position, ncells_1d, scheme, conversion, 1.0, 0.0, binned)
    if not fuse_rhs:
        if pk_from_density:
            density_fourier = fourier.fft_3D_real(density, param["nthreads"])
            k, Pk, Nmodes = fourier.fourier_grid_to_Pk(density_fourier, param["MAS_index"])
            del density_fourier
            Pk *= (param["boxlen"] / ncells_1d ** 2) ** 3
            k *= 2 * np.pi / param["boxlen"]
            iostream.write_power_spectrum_to_ascii_file(k, Pk, Nmodes, param)
        param["compute_additional_field"] = True
        additional_field = get_additional_field(additional_field, density, param, tables)
        param["compute_additional_field"] = False
        rhs_poisson(density, additional_field, param)
        rhs = density
        del density

    if LINEAR_NEWTON_SOLVER == "multigrid":
        potential = initialise_potential(potential, rhs, param, tables)
        potential = multigrid.linear(potential, rhs, param)
    elif LINEAR_NEWTON_SOLVER in ("fft", "fft_7pt"):
        potential = fft(rhs, param)

    order = param["gradient_stencil_order"]
    velocity, half_dt = kick if kick is not None else (None, 0.0)
    fused = binned is not None and ncells_1d >= 16 and LINEAR_NEWTON_SOLVER != "full_fft"
    half_c2 = np.float32(0)
    if "fr" == THEORY:
        _, fR_a, c2 = _fr_background(param)
        half_c2 = np.float32(0.5 * (-fR_a) * c2)
    if fused:
        # gradient fused into the binned interpolation: no force grid
        if order not in (2, 3, 5, 7):
            raise NotImplementedError(f"Unsupported: gradient_order={order}")
        if "fr" == THEORY:
            if param["fR_n"] not in (1, 2):
                raise NotImplementedError(f"Unsupported: fR_n={param['fR_n']}")
            acceleration, maxima = mesh.interp_kick_phi(potential, additional_field, half_c2, param["fR_n"], order,
                                                        position, velocity, scheme, half_dt, binned)
        else:
            acceleration, maxima = mesh.interp_kick_phi(potential, None, 0.0, 0, order, position, velocity, scheme,
                                                        half_dt, binned)
    else:
        if "fr" == THEORY:
            if LINEAR_NEWTON_SOLVER == "full_fft":
                force = fft_force(rhs, param)
                mesh.add_derivative_fR(force, additional_field, half_c2, param["fR_n"], order)
            else:
                force = mesh.derivative_fR(potential, additional_field, half_c2, param["fR_n"], order, padded=True)
        else:
            if LINEAR_NEWTON_SOLVER == "full_fft":
                force = fft_force(rhs, param)
            else:
                force = mesh.derivative(potential, order, padded=True)
        acceleration, maxima = mesh.interp_kick(force, position, velocity, scheme, half_dt, binned)
        del force
    del rhs, binned
    return acceleration, potential, additional_field, (maxima if kick is not None else None)


def pm(position, param, potential=_EMPTY, additional_field=_EMPTY, tables=[]):
    """solver.py:30-215: particle-mesh acceleration.  Returns (acceleration [Np,3], potential [N^3],
    additional_field [N^3] or empty)."""
    c = _lib.Ctx()
    pos = c.dev(position)
    pot = c.dev(potential) if len(potential) else _EMPTY
    add = c.dev(additional_field) if len(additional_field) else _EMPTY
    acc, pot, add, _ = _pm_device(pos, param, pot, add, tables)
    return c.ret(acc), (c.ret(pot) if len(pot) else _EMPTY), (c.ret(add) if len(add) else _EMPTY)


def force_3d(rhs, param):
    """solver.py:582-639"""
    param["MAS_index"] = 0
    name = param["linear_newton_solver"].casefold()
    c = _lib.Ctx()
    trhs = c.dev(rhs)
    if name == "multigrid":
        param["compute_additional_field"] = False
        potential = initialise_potential(_EMPTY, trhs, param, [])
        potential = multigrid.linear(potential, trhs, param)
        force = mesh.derivative(potential, param["gradient_stencil_order"])
    elif name in ("fft", "fft_7pt"):
        potential = fft(trhs.clone(), param)
        force = mesh.derivative(potential, param["gradient_stencil_order"])
    elif name == "full_fft":
        force = fft_force(trhs, param)
    else:
        raise NotImplementedError(f"Unsupported LINEAR_NEWTON_SOLVER={name!r}")
    return c.ret(force)
