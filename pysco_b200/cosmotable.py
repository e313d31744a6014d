"""Background tables (reference pysco/cosmotable.py:18-383) without astropy.

generate(param) returns the same 13 interpolators, in the same order, as the reference:
[lna(t), t(lna), H(lna), D1, f1, D2, f2, D3a, f3a, D3b, f3b, D3c, f3c], and sets
param["Om_r"], param["Om_lambda"].  Host-side scalars only (three look-ups per step).
"""
import logging
import os

import numpy as np
from scipy.integrate import cumulative_trapezoid, solve_ivp
from scipy.interpolate import interp1d

from .cosmology import Flatw0waCDM


def _growth_rhs(lnaexp, y, cosmo, mu0):
    """Linear + 2nd/3rd-order LPT growth ODEs in ln a (cosmotable.py:139-247; Rampf & Buchert 2012)."""
    aexp = np.exp(lnaexp)
    z = 1.0 / aexp - 1
    Om_z = cosmo.Om(z)
    Or_z = cosmo.Ogamma(z) + cosmo.Onu(z)
    Ode_z = cosmo.Ode(z)
    mu = 1.0 if mu0 is None else 1 + (mu0 * Ode_z / cosmo.Ode0)
    beta = 1.5 * mu * Om_z
    gamma = 0.5 * (1.0 - 3.0 * Ode_z * (cosmo.w0 + cosmo.wa * (1.0 - aexp)) - Or_z)
    D1, dD1, D2, dD2, D3a, dD3a, D3b, dD3b, D3c, dD3c = y
    return np.array([
        dD1, -gamma * dD1 + beta * D1,
        dD2, -gamma * dD2 + beta * (D2 - D1 ** 2),
        dD3a, -gamma * dD3a + beta * (D3a - 2.0 * D1 ** 3),
        dD3b, -gamma * dD3b + beta * (D3b - 2.0 * D1 * (D2 - D1 ** 2)),
        dD3c, (1 - gamma) * dD3c + D2 * dD1 - D1 * dD2 - beta * D1 ** 3,
    ])


def compute_growth_functions(cosmo, param):
    """cosmotable.py:113-137 -> array [lna, d1, f1, d2, f2, d3a, f3a, d3b, f3b, d3c, f3c]"""
    lna_span = (np.log(1e-8), 0.0)
    a_eq = (cosmo.Ogamma0 + cosmo.Onu0) / cosmo.Om0
    if (cosmo.Ogamma0 + cosmo.Onu0) == 0:
        a_eq = 2e-7
    d1 = 3.0 / 5 * a_eq
    y0 = [d1, 0, -3.0 / 7 * d1 ** 2, 0, -1.0 / 3.0 * d1 ** 3, 0, 10.0 / 21.0 * d1 ** 3, 0, -1.0 / 7.0 * d1 ** 3, 0]
    lna = np.linspace(lna_span[0], lna_span[1], 100_000)
    mu0 = param["parametrized_mu0"] if param["theory"].casefold() == "parametrized" else None
    sol = solve_ivp(_growth_rhs, lna_span, y0, t_eval=lna, rtol=1e-13, atol=1e-13, args=(cosmo, mu0))
    y = sol.y
    return np.array([lna, y[0], y[1] / y[0], y[2], y[3] / y[2], y[4], y[5] / y[4], y[6], y[7] / y[6],
                     y[8], y[9] / y[8]])


def generate(param):
    """cosmotable.py:18-110"""
    cosmo = Flatw0waCDM(H0=param["H0"], Om0=param["Om_m"], Tcmb0=param["T_cmb"], Neff=param["N_eff"],
                        w0=param["w0"], wa=param["wa"])
    param["Om_r"] = cosmo.Ogamma0 + cosmo.Onu0
    param["Om_lambda"] = cosmo.Ode0
    lna = np.linspace(np.log(1.0 / 201), 0, 100_000)
    a = np.exp(lna)
    dlna = lna[1] - lna[0]
    E_array = cosmo.efunc(1.0 / a - 1)
    t_supercomoving = cumulative_trapezoid(dlna / (a ** 2 * E_array), initial=0)
    t_supercomoving -= t_supercomoving[-1]
    growth = compute_growth_functions(cosmo, param)
    growth = growth[:, growth[0] > lna[0]]
    lng = growth[0]
    if "base" in param and param["base"]:
        os.makedirs(param["base"], exist_ok=True)
        logging.warning(f"Write table in: {param['base']}/evolution_table_pysco.txt")
        np.savetxt(f"{param['base']}/evolution_table_pysco.txt",
                   np.c_[(a, E_array, t_supercomoving) + tuple(np.interp(lna, lng, g) for g in growth[1:])],
                   header="aexp, H/H0, t_supercomoving, dplus1, f1, dplus2, f2, dplus3a, f3a, dplus3b, f3b, dplus3c, f3c")
    return ([interp1d(t_supercomoving, lna, fill_value="extrapolate"),
             interp1d(lna, t_supercomoving, fill_value="extrapolate"),
             interp1d(lna, param["H0"] * E_array, fill_value="extrapolate")]
            + [interp1d(lng, g, fill_value="extrapolate") for g in growth[1:]])
