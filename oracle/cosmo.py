"""CPU oracle for the cosmological normalisation.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against astropy: ``simpleDS/cosmo.py`` delegates the background cosmology to
``astropy.cosmology`` (``Planck15``, cosmo.py:12, 20), astropy (>= 5.0.4 per setup.py:28-34) is
not in this image and the reference's cosmology tests hold no literal values
(tests/test_cosmo.py compares against astropy itself).  What is restated here, value level:

* astropy's published ``FlatLambdaCDM`` algorithm -- ``efunc`` with photons and massive
  neutrinos (``FLRW.nu_relative_density``: the Komatsu et al. 2011 eq. 26 fit, p = 1.83,
  k = 0.3173), ``comoving_distance`` as ``scipy.integrate.quad`` of 1/E (the same library
  routine astropy calls) -- with the Planck15 parameters (H0 67.74, Om0 0.3075, Tcmb0 2.7255,
  Neff 3.046, m_nu (0, 0, 0.06) eV, Ob0 0.0486) and CODATA-2018 constants;
* ``simpleDS/cosmo.py:23-186`` (calc_z, calc_freq, u2kperp, kperp2u, eta2kparr, kparr2eta, X2Y);
* ``DelaySpectrum.update_cosmology`` (delay_spectrum.py:1201-1365): redshift, k_parallel,
  k_perpendicular, unit_conversion, thermal_conversion, little-h scaling.

Anchors that ARE checked (tests/test_cosmo.py): the reference tests' identities
(tests/test_cosmo.py:66-71, 131-140, 203-210: X2Y = (2 pi)^3 / (kperp(1)^2 kparr(1 s)), round
trips) and the WMAP9 comoving distances quoted in astropy's own documentation
(1916.07, 3363.07, 4451.75 Mpc at z = 0.5, 1, 1.5), which pin the massless-neutrino branch.
"""
import numpy as np
from scipy import integrate

C_KM_S = 299792.458
C_M_S = 299792458.0
K_B = 1.380649e-23            # J / K
G_SI = 6.6743e-11             # m^3 / (kg s^2)
SIGMA_SB = 5.6703744191844314e-08   # W / (m^2 K^4)
MPC_M = 3.0856775814913674e22
KB_EV_K = 8.617333262145179e-05     # eV / K
F21_HZ = 1420405751.7667      # cosmo.py:16


class FlatLambdaCDM:
    """astropy.cosmology.FlatLambdaCDM, the parts the reference touches."""

    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, m_nu=(0.0,), name=None):
        self.name, self.H0, self.Om0, self.Tcmb0, self.Neff = name, float(H0), float(Om0), float(Tcmb0), float(Neff)
        h0_si = self.H0 * 1e3 / MPC_M
        self.critical_density0 = 3.0 * h0_si**2 / (8.0 * np.pi * G_SI)          # kg / m^3
        self.Ogamma0 = 4.0 * SIGMA_SB / C_M_S**3 * self.Tcmb0**4 / self.critical_density0
        m_nu = np.atleast_1d(np.asarray(m_nu, dtype=float))
        self.Tnu0 = 0.7137658555036082 * self.Tcmb0
        nneutrinos = int(np.floor(self.Neff))
        if self.Tcmb0 > 0 and nneutrinos > 0:
            if m_nu.size == 1:
                m_nu = np.full(nneutrinos, m_nu[0])
            self.massive = m_nu[m_nu > 0]
            self.nmasslessnu = float(np.sum(m_nu == 0))
            self.neff_per_nu = self.Neff / nneutrinos
            self.nu_y = self.massive / (KB_EV_K * self.Tnu0)
            self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
            self.has_rad = True
        else:
            self.massive, self.nmasslessnu, self.neff_per_nu, self.nu_y = np.zeros(0), 0.0, 0.0, np.zeros(0)
            self.Onu0, self.has_rad = 0.0, False
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        self.hubble_distance = C_KM_S / self.H0   # Mpc

    def nu_relative_density(self, z):
        prefac = 0.22710731766
        z = np.asarray(z, dtype=float)
        if self.massive.size == 0:
            return prefac * self.Neff * np.ones_like(z)
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr_nu_y = self.nu_y / (1.0 + np.expand_dims(z, axis=-1))
        rel_mass = (1.0 + (k * curr_nu_y) ** p) ** invp
        return prefac * self.neff_per_nu * (rel_mass.sum(-1) + self.nmasslessnu)

    def efunc(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=float)
        orad = self.Ogamma0 * (1.0 + self.nu_relative_density(z)) if self.has_rad else 0.0
        return np.sqrt(zp1**3 * (orad * zp1 + self.Om0) + self.Ode0)

    def comoving_distance(self, z):
        f = np.vectorize(lambda zz: integrate.quad(lambda x: 1.0 / float(self.efunc(x)), 0.0, zz)[0])
        return self.hubble_distance * f(np.asarray(z, dtype=float))

    comoving_transverse_distance = comoving_distance  # flat


Planck15 = FlatLambdaCDM(67.74, 0.3075, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06), name="Planck15")
WMAP9 = FlatLambdaCDM(69.32, 0.2865, Tcmb0=2.725, Neff=3.04, m_nu=(0.0,), name="WMAP9")


def calc_z(freq_hz):                      # cosmo.py:23-39
    return F21_HZ / np.asarray(freq_hz, dtype=float) - 1


def calc_freq(redshift):                  # cosmo.py:42-56
    return F21_HZ / (1 + redshift)


def u2kperp(u, z, cosmo=Planck15):        # cosmo.py:59-80, 1 / Mpc
    return 2 * np.pi * u / cosmo.comoving_transverse_distance(z)


def kperp2u(kperp, z, cosmo=Planck15):    # cosmo.py:83-108
    return kperp * cosmo.comoving_transverse_distance(z) / (2 * np.pi)


def eta2kparr(eta_s, z, cosmo=Planck15):  # cosmo.py:111-134, 1 / Mpc
    return eta_s * (2 * np.pi * cosmo.H0 * F21_HZ * cosmo.efunc(z)) / (C_KM_S * (1 + z) ** 2)


def kparr2eta(kparr, z, cosmo=Planck15):  # cosmo.py:137-160, s
    return kparr * C_KM_S * (1 + z) ** 2 / (2 * np.pi * cosmo.H0 * F21_HZ * cosmo.efunc(z))


def X2Y(z, cosmo=Planck15):               # cosmo.py:163-186, Mpc^3 s / sr
    return C_KM_S * (1 + z) ** 2 * cosmo.comoving_transverse_distance(z) ** 2 / (cosmo.H0 * F21_HZ * cosmo.efunc(z))


def update_cosmology(freq_hz, delay_s, uvw_m, taper_values, beam_sq_area, power_unit="Jy^2 Hz^2",
                     cosmo=Planck15, littleh_units=False):
    """delay_spectrum.py:1201-1365, value level.  freq (S,F) Hz; delay (D,) s; uvw (3,) m;
    taper_values (F,); beam_sq_area (S,P,F) sr.  Returns dict."""
    freq_hz = np.asarray(freq_hz, dtype=float)
    S, F = freq_hz.shape
    redshift = calc_z(freq_hz).mean(axis=1)                                               # 1226
    k_parallel = eta2kparr(np.asarray(delay_s).reshape(1, -1), redshift.reshape(S, 1), cosmo)  # 1228-1232
    uvw_wave = np.linalg.norm(uvw_m) / (C_M_S / freq_hz.mean(axis=1))                     # 1234-1236
    k_perpendicular = u2kperp(uvw_wave, redshift, cosmo)                                  # 1237-1239
    x2y = X2Y(calc_z(freq_hz), cosmo).reshape(S, 1, F)
    w2 = np.asarray(taper_values, dtype=float).reshape(1, 1, F) ** 2
    f = freq_hz.reshape(S, 1, F)
    i4 = np.trapezoid(f**4 / x2y * w2 * beam_sq_area, x=f, axis=-1)                       # 1253-1268
    i0 = np.trapezoid(1.0 / x2y * w2 * beam_sq_area, x=f, axis=-1)                        # 1279-1292
    if power_unit == "Jy^2 Hz^2":
        unit_conversion = (C_M_S**2 / (2 * K_B)) ** 2 / i4 * 1e6 * 1e-52   # -> mK^2 Mpc^3 / (Jy^2 Hz^2)
    elif power_unit == "K^2 sr^2 Hz^2":
        unit_conversion = 1.0 / i0 * 1e6                                   # -> mK^2 Mpc^3 / (K^2 sr^2 Hz^2)
    else:
        unit_conversion = np.ones((S, beam_sq_area.shape[1]))
    thermal_conversion = 1.0 / i4 * 1e6                                    # 1318-1333, mK^2 Mpc^3 / (K^2 sr^2 Hz^6)
    out = dict(redshift=redshift, k_parallel=k_parallel, k_perpendicular=k_perpendicular,
               unit_conversion=unit_conversion, thermal_conversion=thermal_conversion, littleh_power=1.0)
    if littleh_units:                                                       # 1340-1365
        h = cosmo.H0 / 100.0
        out["k_parallel"] = k_parallel / h
        out["k_perpendicular"] = k_perpendicular / h
        out["littleh_power"] = h**3
    return out
