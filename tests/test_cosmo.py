"""Cosmological normalisation: cosmo.py functions, the flat-LambdaCDM background
and DelaySpectrum.update_cosmology / remove_cosmology.

astropy is absent, so parity against astropy itself is UNPINNED (oracle/cosmo.py explains what is
restated); what IS checked: (CPU) the oracle against the WMAP9 distances printed in astropy's own
documentation and the identities of the reference's tests/test_cosmo.py; (GPU) the drop-in --
E(z) and the distance integral run in libsds.so -- against the oracle to 1e-10, and
update_cosmology against the oracle's restatement of delay_spectrum.py:1201-1365.
"""
import numpy as np
import pytest

from oracle import cosmo as oc

RTOL = 1e-10


def test_oracle_wmap9_distances_match_astropy_documentation():
    # astropy docs, "cosmo.comoving_distance(np.array([0.5, 1.0, 1.5]))" with WMAP9
    got = oc.WMAP9.comoving_distance([0.5, 1.0, 1.5])
    np.testing.assert_allclose(got, [1916.06941724, 3363.07062107, 4451.7475201], rtol=2e-11)


def test_oracle_planck15_densities():
    c = oc.Planck15
    assert np.isclose(c.Ogamma0 + c.Onu0 + c.Om0 + c.Ode0, 1.0, rtol=0, atol=1e-15)
    assert np.isclose(c.Ode0, 0.69100983, atol=1e-8) and np.isclose(c.Ogamma0, 5.38926e-05, rtol=1e-5)
    assert np.isclose(float(c.efunc(0.0)), 1.0, rtol=0, atol=1e-15)
    assert np.isclose(c.nu_y[0], 0.06 / (8.617333262145179e-05 * 0.7137658555036082 * 2.7255))


def test_oracle_reference_identities():
    z = 7.0                                                    # tests/test_cosmo.py:203-210
    x = (2 * np.pi) ** 3 / (oc.u2kperp(1, z) ** 2 * oc.eta2kparr(1.0, z))
    assert np.isclose(x, oc.X2Y(z), rtol=1e-13)
    assert np.isclose(oc.calc_z(0.15e9), oc.F21_HZ / 0.15e9 - 1)   # :36-40
    assert np.isclose(oc.calc_freq(oc.calc_z(0.15e9)), 0.15e9)      # :43-47
    zz = 9.19508                                               # :131-193 round trip
    assert np.isclose(oc.kparr2eta(oc.eta2kparr(200e-9, zz), zz), 200e-9, rtol=1e-14)
    assert np.isclose(oc.kperp2u(oc.u2kperp(10, 7.6363125), 7.6363125), 10, rtol=1e-14)  # :66-113


@pytest.fixture(scope="module")
def sc(cuda_device):
    from simpleds_b200 import cosmo
    from simpleds_b200.units import Quantity
    return cosmo, Quantity


@pytest.mark.gpu
def test_gpu_background_matches_oracle(sc):
    cosmo, Quantity = sc
    z = np.concatenate([[0.0, 1e-3, 0.5, 1.0, 1.5], np.linspace(5.0, 30.0, 200)])
    for dev, ora in ((cosmo.Planck15, oc.Planck15), (cosmo.WMAP9, oc.WMAP9)):
        np.testing.assert_allclose(dev.efunc(z), ora.efunc(z), rtol=1e-13)
        np.testing.assert_allclose(dev.comoving_transverse_distance(z)[1:], ora.comoving_distance(z)[1:], rtol=RTOL)
        assert dev.comoving_distance(0.0) == 0.0
        assert np.isclose(dev.Ode0, ora.Ode0, rtol=0, atol=1e-15)
    np.testing.assert_allclose(cosmo.WMAP9.comoving_distance([0.5, 1.0, 1.5]),
                               [1916.06941724, 3363.07062107, 4451.7475201], rtol=2e-11)


@pytest.mark.gpu
def test_gpu_cosmo_functions(sc):
    """tests/test_cosmo.py restated on the drop-in."""
    cosmo, Quantity = sc
    f = Quantity(np.float64(0.15e9), "Hz")
    assert np.isclose(cosmo.calc_z(f), cosmo.f21.value / 0.15e9 - 1)
    assert np.isclose(cosmo.calc_freq(cosmo.calc_z(f)).value, 0.15e9) and str(cosmo.calc_freq(10).unit) == "Hz"
    with pytest.raises((TypeError, ValueError)):   # :23-33 wrong unit / not a Quantity
        cosmo.calc_z(Quantity(np.float64(0.15e9), "m"))
    with pytest.raises(TypeError):
        cosmo.calc_z(0.15e9)
    k = cosmo.u2kperp(10, 7.6363125)
    assert str(k.unit) == "Mpc^-1" and np.isclose(k.value, oc.u2kperp(10, 7.6363125), rtol=RTOL)
    assert np.isclose(cosmo.kperp2u(k, 7.6363125), 10, rtol=1e-13)
    with pytest.raises(TypeError):
        cosmo.kperp2u(0.1, 7.0)
    eta = Quantity(np.float64(200e-9), "s")
    kp = cosmo.eta2kparr(eta, 9.19508)
    assert str(kp.unit) == "Mpc^-1" and np.isclose(kp.value, oc.eta2kparr(200e-9, 9.19508), rtol=1e-13)
    back = cosmo.kparr2eta(kp, 9.19508)
    assert str(back.unit) == "s" and np.isclose(back.value, 200e-9, rtol=1e-13)
    with pytest.raises((TypeError, ValueError)):
        cosmo.eta2kparr(Quantity(np.float64(200e-9), "m"), 9.19508)
    x = cosmo.X2Y(7)
    assert str(x.unit) == "Mpc^3 s sr^-1"
    want = (2 * np.pi) ** 3 / (cosmo.u2kperp(1, 7).value ** 2 * cosmo.eta2kparr(Quantity(np.float64(1.0), "s"), 7).value)
    assert np.isclose(want, x.value, rtol=1e-12) and np.isclose(x.value, oc.X2Y(7.0), rtol=RTOL)


@pytest.mark.gpu
@pytest.mark.parametrize("vis_units,littleh", [("Jy", False), ("K str", False), ("Jy", True), ("uncalib", False)])
def test_gpu_update_cosmology_matches_oracle(sc, vis_units, littleh):
    cosmo, Quantity = sc
    from scipy.signal import windows

    from simpleds_b200.delay_spectrum import DelaySpectrum
    rng = np.random.default_rng(3)
    S, P, B, T, F = 2, 2, 3, 2, 32
    freq = np.stack([100e6 + 97656.25 * np.arange(F), 150e6 + 97656.25 * np.arange(F)])
    data = (rng.standard_normal((S, 1, P, B, T, F)) + 1j * rng.standard_normal((S, 1, P, B, T, F)))
    beam = rng.uniform(0.5, 0.9, (S, P, F))
    bsq = rng.uniform(0.2, 0.5, (S, P, F))
    uvw = np.array([14.6, 3.0, 0.1])
    ds = DelaySpectrum.from_arrays(data, np.zeros(data.shape, bool), np.ones(data.shape), freq,
                                   np.full((B, T), 10.7), np.array([-5, -6]), vis_units=vis_units,
                                   trcvr=144.0, uvw=uvw)
    ds.set_beam_area(beam, bsq)
    ds.calculate_delay_spectrum(littleh_units=littleh)
    cooked = ds.power_array.numpy().copy()
    thermal = ds.thermal_power.numpy().copy()
    unit = {"Jy": "Jy^2 Hz^2", "K str": "K^2 sr^2 Hz^2"}.get(vis_units, "none")
    ref = oc.update_cosmology(freq, ds.delay_array.value * 1e-9, uvw, windows.blackmanharris(F), bsq,
                              power_unit=unit, littleh_units=littleh)
    np.testing.assert_allclose(ds.redshift, ref["redshift"], rtol=1e-14)
    np.testing.assert_allclose(ds.k_parallel.value, ref["k_parallel"], rtol=1e-12)
    np.testing.assert_allclose(ds.k_perpendicular.value, ref["k_perpendicular"], rtol=RTOL)
    np.testing.assert_allclose(ds.unit_conversion.value, ref["unit_conversion"], rtol=RTOL)
    np.testing.assert_allclose(ds.thermal_conversion.value, ref["thermal_conversion"], rtol=RTOL)
    want_unit = "Hz^2" if vis_units == "uncalib" else ("Mpc^3 littleh^-3 mK^2" if littleh else "Mpc^3 mK^2")
    assert str(ds.power_array.unit) == want_unit == str(ds.noise_power.unit)
    assert str(ds.k_parallel.unit) == ("Mpc^-1 littleh" if littleh else "Mpc^-1")
    # the raw product is what remove_cosmology gives back (tests/test_delay_spectrum.py:1122-1147)
    ds.remove_cosmology()
    raw = ds.power_array.numpy()
    assert ds.unit_conversion is None and ds.k_parallel is None
    conv = ref["unit_conversion"].reshape(S, P, 1, 1, 1, 1) * ref["littleh_power"]
    np.testing.assert_allclose(cooked, raw * conv, rtol=1e-12)
    np.testing.assert_allclose(thermal, ds.thermal_power.numpy() * ref["thermal_conversion"].reshape(S, P, 1, 1, 1)
                               * ref["littleh_power"], rtol=1e-12)
    if vis_units != "uncalib":
        assert str(ds.power_array.unit) == unit.replace("Jy^2 Hz^2", "Hz^2 Jy^2").replace("K^2 sr^2 Hz^2", "Hz^2 K^2 sr^2")
    # a second cosmology re-normalises from the raw state; a non-cosmology is rejected
    ds.update_cosmology(cosmology=cosmo.WMAP9)
    ref9 = oc.update_cosmology(freq, ds.delay_array.value * 1e-9, uvw, windows.blackmanharris(F), bsq,
                               power_unit=unit, cosmo=oc.WMAP9)
    np.testing.assert_allclose(ds.power_array.numpy(), raw * ref9["unit_conversion"].reshape(S, P, 1, 1, 1, 1), rtol=1e-10)
    with pytest.raises(ValueError):
        ds.update_cosmology(cosmology="Planck15")
