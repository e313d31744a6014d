"""Flat w0-wa CDM background with photons and massless neutrinos.

Dependency-free stand-in for the one astropy class the reference uses,
``astropy.cosmology.Flatw0waCDM(H0, Om0, Tcmb0, Neff, w0, wa)`` (reference call sites:
cosmotable.py:48-57, 64, 163-165, 252-256).  Exposes exactly the attributes and methods those call
sites read: ``Om0, Ogamma0, Onu0, Ode0, w0, wa, efunc, Om, Ogamma, Onu, Ode``.

Check value: examples/INFOS of the reference quotes Omega_r = 0.000080763 for H0 = 72,
T_cmb = 2.726, N_eff = 3.044 (this class gives 8.0763e-05).
"""
import numpy as np

_C = 299792458.0  # m/s
_G = 6.6743e-11  # m^3 kg^-1 s^-2 (CODATA 2018)
_SIGMA_SB = 5.670374419e-8  # W m^-2 K^-4 (CODATA 2018)
_MPC = 3.0856775814913673e22  # m


class Flatw0waCDM:
    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, w0=-1.0, wa=0.0):
        self.H0 = float(H0)
        self.Om0 = float(Om0)
        self.Tcmb0 = float(Tcmb0)
        self.Neff = float(Neff)
        self.w0 = float(w0)
        self.wa = float(wa)
        H0_s = self.H0 * 1e3 / _MPC
        rho_crit = 3.0 * H0_s ** 2 / (8.0 * np.pi * _G)
        rho_gamma = 4.0 * _SIGMA_SB / _C ** 3 * self.Tcmb0 ** 4
        self.Ogamma0 = rho_gamma / rho_crit
        # massless neutrinos: 7/8 (4/11)^(4/3) per species
        self.Onu0 = 0.875 * (4.0 / 11.0) ** (4.0 / 3.0) * self.Neff * self.Ogamma0
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0

    def _de_scale(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        return zp1 ** (3.0 * (1.0 + self.w0 + self.wa)) * np.exp(-3.0 * self.wa * (zp1 - 1.0) / zp1)

    def efunc(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=np.float64)
        return np.sqrt(self.Om0 * zp1 ** 3 + (self.Ogamma0 + self.Onu0) * zp1 ** 4
                       + self.Ode0 * self._de_scale(z))

    def Om(self, z):
        return self.Om0 * (1.0 + np.asarray(z, dtype=np.float64)) ** 3 / self.efunc(z) ** 2

    def Ogamma(self, z):
        return self.Ogamma0 * (1.0 + np.asarray(z, dtype=np.float64)) ** 4 / self.efunc(z) ** 2

    def Onu(self, z):
        return self.Onu0 * (1.0 + np.asarray(z, dtype=np.float64)) ** 4 / self.efunc(z) ** 2

    def Ode(self, z):
        return self.Ode0 * self._de_scale(z) / self.efunc(z) ** 2
