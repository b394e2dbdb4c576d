"""Drop-in for the reference's ``simpleDS/cosmo.py`` (same names and argument meaning) plus the
flat-LambdaCDM background it takes from astropy (``Planck15``, cosmo.py:12, 20).

astropy is not available, so the cosmology is carried as plain parameters
(``FlatLambdaCDM``: H0, Om0, Tcmb0, Neff, m_nu as astropy spells them); E(z) and the comoving
distance integral are evaluated in libsds.so (``sds_cosmo_evaluate``) -- there is no CPU
fallback.  Values are SI / Mpc floats or arrays; Quantities of ``simpleds_b200.units`` are
accepted where the reference demands them (``quantity_input``) and returned with unit tags.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .units import Quantity, Unit, is_quantity

C_KM_S = 299792.458                   # astropy.constants.c
C_M_S = 299792458.0
K_B = 1.380649e-23                    # astropy.constants.k_B (CODATA 2018)
_G_SI = 6.6743e-11
_SIGMA_SB = 5.6703744191844314e-08
_MPC_M = 3.0856775814913674e22
_KB_EV_K = 8.617333262145179e-05

f21 = Quantity(np.float64(1420405751.7667), "Hz")   # cosmo.py:16
_F21 = 1420405751.7667


class FlatLambdaCDM:
    """The slice of ``astropy.cosmology.FlatLambdaCDM`` the reference touches: ``H0``,
    ``efunc``, ``comoving_transverse_distance`` (and the density parameters behind them)."""

    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, m_nu=(0.0,), Ob0=None, name=None):
        self.name, self.Ob0 = name, Ob0
        self.H0, self.Om0, self.Tcmb0, self.Neff = float(H0), float(Om0), float(Tcmb0), float(Neff)
        h0_si = self.H0 * 1e3 / _MPC_M
        rho_c = 3.0 * h0_si**2 / (8.0 * np.pi * _G_SI)
        self.Ogamma0 = 4.0 * _SIGMA_SB / C_M_S**3 * self.Tcmb0**4 / rho_c
        m_nu = np.atleast_1d(np.asarray(m_nu, dtype=np.float64))
        nnu = int(np.floor(self.Neff))
        p = _lib.CosmoParams()
        p.h0_km_s_mpc, p.om0, p.ogamma0 = self.H0, self.Om0, self.Ogamma0
        if self.Tcmb0 > 0 and nnu > 0:
            if m_nu.size == 1:
                m_nu = np.full(nnu, m_nu[0])
            massive = m_nu[m_nu > 0]
            if massive.size > _lib.SDS_COSMO_MAX_NU:
                raise ValueError(f"at most {_lib.SDS_COSMO_MAX_NU} massive neutrino species")
            p.neff_per_nu = self.Neff / nnu
            p.nmassless_nu = float(np.sum(m_nu == 0))
            p.nmassive_nu = int(massive.size)
            for i, m in enumerate(massive):
                p.nu_y[i] = float(m / (_KB_EV_K * 0.7137658555036082 * self.Tcmb0))
            p.has_massive_nu_or_radiation = 1
            prefac, pw, invp, k = 0.22710731766, 1.83, 0.54644808743, 0.3173
            if massive.size == 0:
                nu0 = prefac * self.Neff
            else:
                nu_y = np.array([p.nu_y[i] for i in range(massive.size)])
                nu0 = prefac * p.neff_per_nu * (np.sum((1.0 + (k * nu_y) ** pw) ** invp) + p.nmassless_nu)
            self.Onu0 = self.Ogamma0 * nu0
        else:
            self.Onu0 = 0.0
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        p.ode0 = self.Ode0
        self._params = p

    # -- device evaluation ----------------------------------------------------------------
    def _evaluate(self, z, want_e, want_d):
        _lib.require_cuda()
        shape = np.shape(z)
        dev = torch.device("cuda", torch.cuda.current_device())
        zt = torch.as_tensor(np.ascontiguousarray(np.asarray(z, dtype=np.float64)).reshape(-1), device=dev)
        e = torch.empty_like(zt) if want_e else None
        d = torch.empty_like(zt) if want_d else None
        rc = _lib.load().sds_cosmo_evaluate(C.byref(self._params), _lib.ptr(zt), zt.numel(), _lib.ptr(e),
                                            _lib.ptr(d), _lib.stream_ptr())
        _lib.check(rc, "sds_cosmo_evaluate")
        out = [t.cpu().numpy().reshape(shape) if t is not None else None for t in (e, d)]
        return [np.float64(o) if o is not None and o.ndim == 0 else o for o in out]

    def efunc(self, z):
        return self._evaluate(z, True, False)[0]

    def comoving_distance(self, z):
        """Mpc."""
        return self._evaluate(z, False, True)[1]

    comoving_transverse_distance = comoving_distance  # flat universe


Planck15 = FlatLambdaCDM(67.74, 0.3075, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06), Ob0=0.0486,
                         name="Planck15")
WMAP9 = FlatLambdaCDM(69.32, 0.2865, Tcmb0=2.725, Neff=3.04, m_nu=(0.0,), Ob0=0.04628, name="WMAP9")


class _DefaultCosmology:
    """``astropy.cosmology.default_cosmology``: get() / set()."""

    def __init__(self):
        self._c = Planck15  # cosmo.py:20

    def get(self):
        return self._c

    def set(self, cosmo):
        if not isinstance(cosmo, FlatLambdaCDM):
            raise ValueError("default cosmology must be a FlatLambdaCDM instance")
        self._c = cosmo


default_cosmology = _DefaultCosmology()


def _need(x, unit, name, fn):
    if not is_quantity(x):
        raise TypeError(f"Argument '{name}' to function '{fn}' has no 'unit' attribute. "
                        "You should pass in an astropy Quantity instead.")
    if not x.unit == Unit(unit):
        raise ValueError(f"Argument '{name}' to function '{fn}' must be in units convertible to '{unit}'.")
    v = x.value
    return np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float64)


def calc_z(freq):
    """cosmo.py:23-39."""
    return _F21 / _need(freq, "Hz", "freq", "calc_z") - 1


def calc_freq(redshift):
    """cosmo.py:42-56."""
    return Quantity(_F21 / (1 + np.asarray(redshift, dtype=np.float64)), "Hz")


def u2kperp(u, z, cosmo=None):
    """cosmo.py:59-80 -> 1 / Mpc."""
    cosmo = default_cosmology.get() if cosmo is None else cosmo
    return Quantity(2 * np.pi * np.asarray(u, dtype=np.float64) / cosmo.comoving_transverse_distance(z), "Mpc^-1")


def kperp2u(kperp, z, cosmo=None):
    """cosmo.py:83-108."""
    cosmo = default_cosmology.get() if cosmo is None else cosmo
    return _need(kperp, "Mpc^-1", "kperp", "kperp2u") * cosmo.comoving_transverse_distance(z) / (2 * np.pi)


def eta2kparr(eta, z, cosmo=None):
    """cosmo.py:111-134 -> 1 / Mpc."""
    cosmo = default_cosmology.get() if cosmo is None else cosmo
    z = np.asarray(z, dtype=np.float64)
    k = _need(eta, "s", "eta", "eta2kparr") * (2 * np.pi * cosmo.H0 * _F21 * cosmo.efunc(z)) / (C_KM_S * (1 + z) ** 2)
    return Quantity(k, "Mpc^-1")


def kparr2eta(kparr, z, cosmo=None):
    """cosmo.py:137-160 -> s."""
    cosmo = default_cosmology.get() if cosmo is None else cosmo
    z = np.asarray(z, dtype=np.float64)
    eta = _need(kparr, "Mpc^-1", "kparr", "kparr2eta") * C_KM_S * (1 + z) ** 2 / (2 * np.pi * cosmo.H0 * _F21 * cosmo.efunc(z))
    return Quantity(eta, "s")


def X2Y(z, cosmo=None):
    """cosmo.py:163-186 -> Mpc^3 s / sr."""
    cosmo = default_cosmology.get() if cosmo is None else cosmo
    z = np.asarray(z, dtype=np.float64)
    e, d = cosmo._evaluate(z, True, True)
    return Quantity(C_KM_S * (1 + z) ** 2 * d**2 / (cosmo.H0 * _F21 * e), "Mpc^3 s sr^-1")
